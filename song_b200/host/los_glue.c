/* los_glue.c -- see los_glue.h */
#include "los_glue.h"

void los_config_release(los_config_buf *b)
{
  if (!b) return;
  free(b->l); free(b->m); free(b->lm_array); free(b->lm_src); free(b->tt_index_l); free(b->tt_index_m);
  memset(b, 0, sizeof *b);
}

int los_config_from_song(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels *pbs,
                         struct bessels2 *pbs2, double tau0, int tau_sampling_mode, int device, los_config_buf *out,
                         ErrorMsg errmsg)
{
  memset(out, 0, sizeof *out);
  song_los_config *c = &out->cfg;
  const int l_size = pbs->l_size, m_size = ppr2->m_size, L_size = pbs2->L_size;
  c->device = device;
  c->tau0 = tau0;
  c->tau_size = ppt2->tau_size; c->tau_sampling = ppt2->tau_sampling;
  c->tau_sampling_mode = tau_sampling_mode;
  c->tau_linstep_song = ppr2->tau_linstep_song;
  c->sources_k3_interpolation = (ppr2->sources_k3_interpolation == cubic_interpolation) ? SONG_LOS_INTERP_CUBIC : SONG_LOS_INTERP_LINEAR;
  c->sources_time_interpolation = (ppr2->sources_time_interpolation == cubic_interpolation) ? SONG_LOS_INTERP_CUBIC : SONG_LOS_INTERP_LINEAR;
  c->bessels_interpolation = (ppr->bessels_interpolation == cubic_interpolation) ? SONG_LOS_INTERP_CUBIC : SONG_LOS_INTERP_LINEAR;
  c->k3_extrapolation = (ppr->bispectra_k3_extrapolation == no_k3_extrapolation) ? SONG_LOS_NO_K3_EXTRAPOLATION : SONG_LOS_FLAT_K3_EXTRAPOLATION;
  c->xx_size = pbs2->xx_size; c->xx = pbs2->xx; c->xx_step = pbs2->xx_step;

  class_alloc(out->l, (size_t)l_size * sizeof(int), errmsg);
  class_alloc(out->m, (size_t)m_size * sizeof(int), errmsg);
  for (int i = 0; i < l_size; ++i) out->l[i] = pbs->l[i];                       /* transfer2_get_l_list, transfer2.c:1081-1103 */
  for (int i = 0; i < m_size; ++i) out->m[i] = ppr2->m[i];
  c->l_size = l_size; c->l = out->l; c->m_size = m_size; c->m = out->m;
  c->L_size = L_size;
  c->n_transfers = size_indexl_indexm(out->l, l_size, out->m, m_size);          /* transfer2.c:900 */

  /* lm_array[index_l][index_m] (transfer2.c:1133-1166) and lm(L,m) of the sources (perturbations2_macros.h:53) */
  class_alloc(out->lm_array, (size_t)l_size * m_size * sizeof(int), errmsg);
  class_alloc(out->lm_src, (size_t)L_size * m_size * sizeof(int), errmsg);
  for (int il = 0; il < l_size; ++il)
    for (int im = 0; im < m_size; ++im)
      out->lm_array[il * m_size + im] = (im <= ppr2->index_m_max[out->l[il]])
          ? multipole2offset_indexl_indexm(out->l[il], out->m[im], out->l, l_size, out->m, m_size) : -1;
  for (int L = 0; L < L_size; ++L)
    for (int im = 0; im < m_size; ++im)
      out->lm_src[L * m_size + im] = (out->m[im] <= L) ? ppt2->lm_array[L][im] : -1;     /* lm(L,m), as transfer2.c:2302 reads it */
  c->lm_array = out->lm_array; c->lm_src = out->lm_src;

  c->has_source_T = ppt2->has_source_T; c->has_source_E = ppt2->has_source_E; c->has_source_B = ppt2->has_source_B;
  c->index_tp2_T = ppt2->index_tp2_T; c->index_tp2_E = ppt2->index_tp2_E; c->index_tp2_B = ppt2->index_tp2_B;
  c->tp2_size = ppt2->tp2_size;
  c->l_max_los_t = ppr2->l_max_los_t; c->l_max_los_p = ppr2->l_max_los_p;

  /* transfer types: T block, then E, then B (transfer2.c:917-940), each n_transfers long */
  int tt = 0;
  c->index_tt2_T = c->index_tt2_E = c->index_tt2_B = 0;
  if (ppt2->has_source_T) { c->index_tt2_T = tt; tt += c->n_transfers; }
  if (ppt2->has_source_E) { c->index_tt2_E = tt; tt += c->n_transfers; }
  if (ppt2->has_source_B) { c->index_tt2_B = tt; tt += c->n_transfers; }
  c->tt2_size = tt;
  class_test(tt < 1, errmsg, "no second-order source type (T, E, B) is switched on: nothing to compute");
  class_alloc(out->tt_index_l, (size_t)tt * sizeof(int), errmsg);
  class_alloc(out->tt_index_m, (size_t)tt * sizeof(int), errmsg);
  for (int i = 0; i < tt; ++i) {                                                /* transfer2.c:1180-1282 */
    int il = -1, im = -1;
    class_test(offset2multipole_indexl_indexm(i % c->n_transfers, out->l, l_size, out->m, m_size, &il, &im) == _FAILURE_ || il < 0 || im < 0,
               errmsg, "index_tt=%d: no (index_l,index_m) for offset %d", i, i % c->n_transfers);
    out->tt_index_l[i] = il; out->tt_index_m[i] = im;
  }
  c->tt_index_l = out->tt_index_l; c->tt_index_m = out->tt_index_m;

  c->J_size = pbs2->J_size;
  c->index_J_TT = pbs2->has_J_TT ? pbs2->index_J_TT : -1;
  c->index_J_EE = pbs2->has_J_EE ? pbs2->index_J_EE : -1;
  c->index_J_EB = pbs2->has_J_EB ? pbs2->index_J_EB : -1;
  return _SUCCESS_;
}

int los_upload_host_projection(struct precision *ppr, struct precision2 *ppr2, struct bessels *pbs, struct bessels2 *pbs2,
                               song_los_ctx *ctx, ErrorMsg errmsg)
{
  /* ragged [J][L][l][m] rows -> flat list, jidx = ((J*L_size + L)*l_size + index_l)*m_size + index_m (bessel2.h:64-91) */
  const int l_size = pbs->l_size, m_size = ppr2->m_size;
  size_t n_j = (size_t)pbs2->J_size * pbs2->L_size * l_size * m_size;
  const double **rows, **ddrows; int *xmin, *xsize;
  class_calloc(rows, n_j, sizeof(double *), errmsg);
  class_calloc(ddrows, n_j, sizeof(double *), errmsg);
  class_calloc(xmin, n_j, sizeof(int), errmsg);
  class_calloc(xsize, n_j, sizeof(int), errmsg);
  const int cubic_J = (ppr->bessels_interpolation == cubic_interpolation);
  for (int J = 0; J < pbs2->J_size; ++J)
    for (int L = 0; L < pbs2->L_size; ++L)
      for (int index_l = 0; index_l < l_size; ++index_l) {
        int index_m_max = MIN(ppr2->index_m_max[pbs2->L[L]], ppr2->index_m_max[pbs->l[index_l]]);   /* bessel2.c:288 */
        for (int index_m = 0; index_m <= index_m_max; ++index_m) {
          size_t q = (((size_t)J * pbs2->L_size + L) * l_size + index_l) * m_size + index_m;
          rows[q] = pbs2->J_Llm_x[J][L][index_l][index_m];
          if (cubic_J) ddrows[q] = pbs2->ddJ_Llm_x[J][L][index_l][index_m];
          xmin[q] = pbs2->index_xmin_J[J][L][index_l][index_m];
          xsize[q] = pbs2->x_size_J[J][L][index_l][index_m];
        }
      }
  int rc = song_los_set_projection(ctx, rows, cubic_J ? ddrows : NULL, xmin, xsize);
  free(rows); free(ddrows); free(xmin); free(xsize);
  if (rc != SONG_LOS_SUCCESS) { snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(ctx)); return _FAILURE_; }
  return _SUCCESS_;
}
