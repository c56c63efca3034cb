/*
 * dropin_harness.c -- TEST HARNESS for the transfer2 drop-in (song_b200/host/transfer2_gpu.c).
 *
 * Plays the part of SONG's main/song.c and of the modules that run before transfer2 (input2, perturbations2,
 * bessel2): it fills the SONG structures -- declared by SONG's own headers, include/transfer2.h & co. -- from the
 * flat arrays of a synthetic workload, calls transfer2_init() exactly as main/song.c:106 does, and returns
 * ptr2->transfer flattened in transfers_%03d.dat order.  Modelled on test/print_transfers2.c:147-185.
 * perturb2_allocate/load/free_k1_level stand in for source/perturbations2.c:457-660 (out of scope).
 */
#include "transfer2.h"
#include "bessel2.h"

typedef struct harness_input {
  double tau0, tau_rec, rs_rec;
  int k_size; const double *k;
  int tau_size; const double *tau_sampling;
  int l_size; const int *l;
  int m_size; const int *m;
  int L_max, l_max_los_t, l_max_los_p;
  int has_T, has_E, has_B;
  int cubic_k3, cubic_bessel, k3_extrapolation;
  double extra_left, extra_right, q_linstep_song, smallest_allowed_variation;
  int xx_size; const double *xx; double xx_step, xx_max;
  int J_size, index_J_TT, index_J_EE, index_J_EB;
  const double *const *J_rows; const double *const *ddJ_rows; const int *index_xmin_J; const int *x_size_J;   /* flat jidx */
  /* pairs in the reference loop order (k1 major, k2 <= k1) */
  const int *k3_pt_size; const long long *pt_off; const double *k3_pt;
  const long long *src_off; const double *src;           /* packed [pair][tp][tau][k3] */
  int store_to_disk, load_from_disk; const char *transfers_dir;
  int verbose;
  int tau_sampling_mode;           /* 0 sources, 1 custom, 2 bessel (ptr2->tau_sampling, input2.c:1480-1500) */
  int cubic_time; double tau_linstep_song;
  /* bessel2 drop-in: when set, bessel2_init (main/song.c:100) builds struct bessels2 -- J_rows & co. above are ignored */
  int use_bessel2_init; double j_l1_cut, J_Llm_cut, bessel_tol_x_min;
  /* benchmark policy of the stand-in perturb2_free_k1_level: keep the sources for the next transfer2_init */
  int keep_sources;
} harness_input;

static int g_keep_sources = 0;

int perturb2_allocate_k1_level(struct perturbs2 *ppt2, int index_k1)
{
  snprintf(ppt2->error_message, _ERRORMSGSIZE_, "harness: sources are kept in memory, nothing to allocate for index_k1=%d", index_k1);
  return _FAILURE_;
}
int perturb2_load_sources_from_disk(struct perturbs2 *ppt2, int index_k1)
{
  snprintf(ppt2->error_message, _ERRORMSGSIZE_, "harness: no sources on disk (index_k1=%d)", index_k1);
  return _FAILURE_;
}
int perturb2_free_k1_level(struct perturbs2 *ppt2, int index_k1)
{
  /* as source/perturbations2.c:609-660: the k1 level of every source type is released */
  if (g_keep_sources) { ppt2->has_allocated_sources[index_k1] = _FALSE_; return _SUCCESS_; }   /* bench: refilling is perturbations2's job */
  for (int tp = 0; tp < ppt2->tp2_size; ++tp) {
    for (int k2 = 0; k2 <= index_k1; ++k2) { free(ppt2->sources[tp][index_k1][k2]); ppt2->sources[tp][index_k1][k2] = NULL; }
    free(ppt2->sources[tp][index_k1]); ppt2->sources[tp][index_k1] = NULL;
  }
  ppt2->has_allocated_sources[index_k1] = _FALSE_;
  return _SUCCESS_;
}

static int n_sources(int l_max, const int *m, int m_size)
{
  int n = 0;
  for (int L = 0; L <= l_max; ++L) for (int im = 0; im < m_size; ++im) if (m[im] <= L) ++n;
  return n;
}

typedef struct harness {
  struct precision pr; struct precision2 pr2; struct background ba; struct thermo th;
  struct perturbs pt; struct perturbs2 pt2; struct bessels bs; struct bessels2 bs2;
  struct transfers tr; struct transfers2 tr2;
  harness_input in;
  int bessel2_done, transfer2_done;
} harness;

static void free_sources(harness *h)
{
  struct perturbs2 *pt2 = &h->pt2;
  if (!pt2->sources) return;
  for (int t = 0; t < pt2->tp2_size; ++t) {
    for (int k1 = 0; k1 < h->in.k_size; ++k1) {
      if (!pt2->sources[t][k1]) continue;
      for (int k2 = 0; k2 <= k1; ++k2) free(pt2->sources[t][k1][k2]);
      free(pt2->sources[t][k1]);
    }
    free(pt2->sources[t]);
  }
  free(pt2->sources); pt2->sources = NULL;
}

/* ppt2->sources[tp][k1][k2] = a fresh copy of the packed input (what perturb2_init leaves in memory, perturbations2.c:82-439) */
int harness_fill_sources(harness *h)
{
  const harness_input *in = &h->in; struct perturbs2 *pt2 = &h->pt2;
  if (pt2->sources && g_keep_sources) {           /* still there: only the bookkeeping flag was cleared */
    for (int k1 = 0; k1 < in->k_size; ++k1) pt2->has_allocated_sources[k1] = _TRUE_;
    return 0;
  }
  free_sources(h);
  pt2->sources = malloc(sizeof(double ***) * pt2->tp2_size);
  for (int t = 0; t < pt2->tp2_size; ++t) pt2->sources[t] = calloc(in->k_size, sizeof(double **));
  #pragma omp parallel for schedule(dynamic, 1)
  for (int k1 = 0; k1 < in->k_size; ++k1) {
    int p = k1 * (k1 + 1) / 2;
    for (int t = 0; t < pt2->tp2_size; ++t) pt2->sources[t][k1] = malloc(sizeof(double *) * (k1 + 1));
    for (int k2 = 0; k2 <= k1; ++k2, ++p) {
      int np = in->k3_pt_size[p];
      for (int t = 0; t < pt2->tp2_size; ++t) {
        size_t n = (size_t)in->tau_size * np;
        pt2->sources[t][k1][k2] = malloc(sizeof(double) * n);
        memcpy(pt2->sources[t][k1][k2], in->src + in->src_off[p] + (long long)t * n, sizeof(double) * n);
      }
    }
    pt2->has_allocated_sources[k1] = _TRUE_;
  }
  return 0;
}

/* fills the structures the modules before transfer2 leave behind (input2, perturbations2, bessel, bessel2) */
harness *harness_create(const harness_input *in, char *errmsg, int errlen)
{
  harness *h = calloc(1, sizeof *h);
  if (!h) { snprintf(errmsg, errlen, "out of memory"); return NULL; }
  h->in = *in;
  g_keep_sources = in->keep_sources;
  struct precision *pr = &h->pr; struct precision2 *pr2 = &h->pr2; struct perturbs2 *pt2 = &h->pt2;
  struct bessels *bs = &h->bs; struct bessels2 *bs2 = &h->bs2; struct transfers2 *tr2 = &h->tr2;

  /* ---- input2 / precision ---- */
  pr->bispectra_k3_extrapolation = in->k3_extrapolation ? flat_k3_extrapolation : no_k3_extrapolation;
  pr->extra_k3_oscillations_left = in->extra_left; pr->extra_k3_oscillations_right = in->extra_right;
  pr->smallest_allowed_variation = in->smallest_allowed_variation;
  pr->bessels_interpolation = in->cubic_bessel ? cubic_interpolation : linear_interpolation;
  pr->bessel_tol_x_min = in->bessel_tol_x_min;
  pr2->m_size = in->m_size;
  for (int im = 0; im < in->m_size; ++im) pr2->m[im] = in->m[im];
  pr2->m_max_song = in->m[in->m_size - 1];
  for (int M = 0; M < _MAX_NUM_AZIMUTHAL_; ++M) { pr2->compute_m[M] = _FALSE_; pr2->index_m[M] = -1; }
  for (int im = 0; im < in->m_size; ++im) { pr2->compute_m[in->m[im]] = _TRUE_; pr2->index_m[in->m[im]] = im; }
  int l_top = in->l[in->l_size - 1];
  if (l_top < in->L_max) l_top = in->L_max;
  pr2->index_m_max = malloc(sizeof(int) * (l_top + 1));
  for (int l = 0; l <= l_top; ++l) { int idx = -1; for (int im = 0; im < in->m_size; ++im) if (in->m[im] <= l) idx = im; pr2->index_m_max[l] = idx; }
  pr2->l_max_los_t = in->l_max_los_t; pr2->l_max_los_p = in->l_max_los_p;
  pr2->q_linstep_song = in->q_linstep_song;
  pr2->sources_k3_interpolation = in->cubic_k3 ? cubic_interpolation : linear_interpolation;
  pr2->sources_time_interpolation = in->cubic_time ? cubic_interpolation : linear_interpolation;
  pr2->tau_linstep_song = in->tau_linstep_song;
  pr2->store_transfers_to_disk = in->store_to_disk; pr2->load_transfers_from_disk = in->load_from_disk;
  pr2->bessel_j_cut_song = in->j_l1_cut; pr2->bessel_J_cut_song = in->J_Llm_cut;

  /* ---- background / thermodynamics ---- */
  h->ba.conformal_age = in->tau0; h->th.tau_rec = in->tau_rec; h->th.rs_rec = in->rs_rec;

  /* ---- perturbations2: samplings and sources ---- */
  pt2->has_perturbations2 = _TRUE_; pt2->has_cmb_bispectra = _TRUE_;
  pt2->has_cmb_temperature = in->has_T; pt2->has_cmb_polarization_e = in->has_E; pt2->has_cmb_polarization_b = in->has_B;
  pt2->k_size = in->k_size; pt2->k = (double *)in->k;
  pt2->tau_size = in->tau_size; pt2->tau_sampling = (double *)in->tau_sampling;
  pt2->has_source_T = in->has_T; pt2->has_source_E = in->has_E; pt2->has_source_B = in->has_B;
  int nT = n_sources(in->l_max_los_t, in->m, in->m_size), nP = n_sources(in->l_max_los_p, in->m, in->m_size), tp = 0;
  if (in->has_T) { pt2->index_tp2_T = tp; tp += nT; }
  if (in->has_E) { pt2->index_tp2_E = tp; tp += nP; }
  if (in->has_B) { pt2->index_tp2_B = tp; tp += nP; }
  pt2->tp2_size = tp;
  pt2->lm_array = malloc(sizeof(int *) * (in->L_max + 1));
  for (int L = 0; L <= in->L_max; ++L) {
    pt2->lm_array[L] = malloc(sizeof(int) * in->m_size);
    for (int im = 0; im < in->m_size; ++im)
      pt2->lm_array[L][im] = (in->m[im] <= L) ? multipole2offset_l_indexm(L, in->m[im], (int *)in->m, in->m_size) : -1;
  }
  pt2->k3 = malloc(sizeof(double **) * in->k_size);
  pt2->k3_size = malloc(sizeof(int *) * in->k_size);
  pt2->has_allocated_sources = calloc(in->k_size, sizeof(short));
  for (int k1 = 0, p = 0; k1 < in->k_size; ++k1) {
    pt2->k3[k1] = malloc(sizeof(double *) * (k1 + 1));
    pt2->k3_size[k1] = malloc(sizeof(int) * (k1 + 1));
    for (int k2 = 0; k2 <= k1; ++k2, ++p) {
      pt2->k3_size[k1][k2] = in->k3_pt_size[p];
      pt2->k3[k1][k2] = (double *)(in->k3_pt + in->pt_off[p]);
    }
  }
  harness_fill_sources(h);

  /* ---- bessel ---- */
  bs->l_size = in->l_size; bs->l = (int *)in->l; bs->l_max = in->l[in->l_size - 1];
  bs->x_max = 0;
  bs2->L_max = in->L_max; bs2->xx_step = in->xx_step; bs2->xx_max = in->xx_max;      /* input2.c:1982-2016 */
  bs2->bessels2_verbose = in->verbose;
  if (in->use_bessel2_init) {
    /* ---- bessel2: the drop-in module builds everything else in struct bessels2 (main/song.c:100) ---- */
    if (bessel2_init(pr, pr2, pt2, bs, bs2) == _FAILURE_) {
      snprintf(errmsg, errlen, "%s", bs2->error_message);
      return NULL;
    }
    h->bessel2_done = 1;
  } else {
    /* ---- bessel2 as the reference's bessel2.c would leave it, from tables given by the caller ---- */
    bs2->L_size = in->L_max + 1;
    bs2->L = malloc(sizeof(int) * bs2->L_size);
    for (int L = 0; L < bs2->L_size; ++L) bs2->L[L] = L;
    bs2->m_size = in->m_size; bs2->m = (int *)in->m;
    bs2->xx = (double *)in->xx; bs2->xx_size = in->xx_size;
    bs2->J_size = in->J_size;
    bs2->has_J_TT = in->index_J_TT >= 0; bs2->has_J_EE = in->index_J_EE >= 0; bs2->has_J_EB = in->index_J_EB >= 0;
    bs2->index_J_TT = in->index_J_TT; bs2->index_J_EE = in->index_J_EE; bs2->index_J_EB = in->index_J_EB;
    bs2->J_Llm_x = malloc(sizeof(double ****) * in->J_size);
    bs2->ddJ_Llm_x = malloc(sizeof(double ****) * in->J_size);
    bs2->index_xmin_J = malloc(sizeof(int ***) * in->J_size);
    bs2->x_size_J = malloc(sizeof(int ***) * in->J_size);
    for (int J = 0; J < in->J_size; ++J) {
      bs2->J_Llm_x[J] = malloc(sizeof(double ***) * bs2->L_size); bs2->ddJ_Llm_x[J] = malloc(sizeof(double ***) * bs2->L_size);
      bs2->index_xmin_J[J] = malloc(sizeof(int **) * bs2->L_size); bs2->x_size_J[J] = malloc(sizeof(int **) * bs2->L_size);
      for (int L = 0; L < bs2->L_size; ++L) {
        bs2->J_Llm_x[J][L] = malloc(sizeof(double **) * in->l_size); bs2->ddJ_Llm_x[J][L] = malloc(sizeof(double **) * in->l_size);
        bs2->index_xmin_J[J][L] = malloc(sizeof(int *) * in->l_size); bs2->x_size_J[J][L] = malloc(sizeof(int *) * in->l_size);
        for (int il = 0; il < in->l_size; ++il) {
          bs2->J_Llm_x[J][L][il] = calloc(in->m_size, sizeof(double *)); bs2->ddJ_Llm_x[J][L][il] = calloc(in->m_size, sizeof(double *));
          bs2->index_xmin_J[J][L][il] = calloc(in->m_size, sizeof(int)); bs2->x_size_J[J][L][il] = calloc(in->m_size, sizeof(int));
          for (int im = 0; im < in->m_size; ++im) {
            size_t q = (((size_t)J * bs2->L_size + L) * in->l_size + il) * in->m_size + im;
            bs2->J_Llm_x[J][L][il][im] = (double *)in->J_rows[q];
            if (in->ddJ_rows) bs2->ddJ_Llm_x[J][L][il][im] = (double *)in->ddJ_rows[q];
            bs2->index_xmin_J[J][L][il][im] = in->index_xmin_J[q];
            bs2->x_size_J[J][L][il][im] = in->x_size_J[q];
          }
        }
      }
    }
  }

  /* ---- transfer2 (fields written by input2.c:1450-1522,1647) ---- */
  tr2->k_sampling = class_transfer2_k3_sampling;
  tr2->tau_sampling = in->tau_sampling_mode == 1 ? custom_tau_sampling : in->tau_sampling_mode == 2 ? bessel_tau_sampling : sources_tau_sampling;
  tr2->transfer2_verbose = in->verbose;
  if (in->transfers_dir) snprintf(tr2->transfers_dir, _FILENAMESIZE_, "%s", in->transfers_dir);
  return h;
}

/* transfer2_init exactly as main/song.c:106 calls it */
int harness_transfer2_init(harness *h, char *errmsg, int errlen)
{
  if (transfer2_init(&h->pr, &h->pr2, &h->ba, &h->th, &h->pt, &h->pt2, &h->bs, &h->bs2, &h->tr, &h->tr2) == _FAILURE_) {
    snprintf(errmsg, errlen, "%s", h->tr2.error_message);
    return 1;
  }
  h->transfer2_done = 1;
  return 0;
}

/* sum over every stored value, read through ptr2->transfer the way a consumer does (bispectra2.c:1298): the
   device->host read of a benchmark step */
double harness_checksum(harness *h)
{
  struct transfers2 *tr2 = &h->tr2;
  double total = 0;
  #pragma omp parallel for reduction(+ : total) schedule(dynamic, 1)
  for (int tt = 0; tt < tr2->tt2_size; ++tt)
    for (int k1 = 0; k1 < h->in.k_size; ++k1)
      for (int k2 = 0; k2 <= k1; ++k2) {
        const double *v = tr2->transfer[tt][k1][k2];
        double s = 0;
        for (int i = 0; i < tr2->k_size_k1k2[k1][k2]; ++i) s += fabs(v[i]);
        total += s;
      }
  return total;
}

/* the result the way bispectra2.c:1213,1298 reads it; sizes via out_sizes[3] = {tt2_size, total k3 values per type,
   count_memorised_transfers} */
int harness_read(harness *h, double *out, long long out_capacity, long long *out_sizes, int *labels_ok, char *errmsg, int errlen)
{
  const harness_input *in = &h->in; struct transfers2 *tr2 = &h->tr2;
  long long per_type = 0;
  for (int k1 = 0; k1 < in->k_size; ++k1) for (int k2 = 0; k2 <= k1; ++k2) per_type += tr2->k_size_k1k2[k1][k2];
  out_sizes[0] = tr2->tt2_size; out_sizes[1] = per_type; out_sizes[2] = tr2->count_memorised_transfers;
  *labels_ok = 1;
  if (out_capacity < per_type * tr2->tt2_size) { snprintf(errmsg, errlen, "output buffer too small"); return 1; }
  for (int tt = 0; tt < tr2->tt2_size; ++tt) {
    char expect[_MAX_LENGTH_LABEL_];
    const char *f = (in->has_T && tt >= tr2->index_tt2_T && tt < tr2->index_tt2_T + tr2->n_transfers) ? "T"
                  : (in->has_E && tt >= tr2->index_tt2_E && tt < tr2->index_tt2_E + tr2->n_transfers) ? "E" : "B";
    snprintf(expect, sizeof expect, "%s_%d_%d", f, tr2->l[tr2->corresponding_index_l[tt]], tr2->m[tr2->corresponding_index_m[tt]]);
    if (strcmp(expect, tr2->tt2_labels[tt]) != 0) *labels_ok = 0;
    if (in->store_to_disk || in->load_from_disk) {
      if (transfer2_load_transfers_from_disk(&h->pt2, tr2, tt) == _FAILURE_) { snprintf(errmsg, errlen, "%s", tr2->error_message); return 1; }
    }
    long long o = 0;
    for (int k1 = 0; k1 < in->k_size; ++k1)
      for (int k2 = 0; k2 <= k1; ++k2) {
        memcpy(out + tt * per_type + o, tr2->transfer[tt][k1][k2], sizeof(double) * tr2->k_size_k1k2[k1][k2]);
        o += tr2->k_size_k1k2[k1][k2];
      }
    if (in->store_to_disk || in->load_from_disk) transfer2_free_type_level(&h->pt2, tr2, tt);
  }
  return 0;
}

/* transfer2_free as main/song.c:168; the structures can then go through harness_fill_sources + harness_transfer2_init again */
int harness_transfer2_free(harness *h, char *errmsg, int errlen)
{
  if (!h->transfer2_done) return 0;
  h->transfer2_done = 0;
  if (transfer2_free(&h->pr2, &h->pt2, &h->tr2) == _FAILURE_) { snprintf(errmsg, errlen, "%s", h->tr2.error_message); return 1; }
  return 0;
}

/* the members of struct bessels2 a test wants to see (bessel2 drop-in): flat jidx tables as in song_los.h */
int harness_bessel2_tables(harness *h, int *index_xmin_J, int *x_size_J, double **J_rows, double **ddJ_rows, long long *sizes)
{
  struct bessels2 *bs2 = &h->bs2; const harness_input *in = &h->in;
  sizes[0] = bs2->J_size; sizes[1] = bs2->L_size; sizes[2] = bs2->xx_size; sizes[3] = bs2->l1_size; sizes[4] = bs2->x_size_max_J;
  sizes[5] = bs2->count_allocated_Js;
  if (!index_xmin_J) return 0;
  for (int J = 0; J < bs2->J_size; ++J)
    for (int L = 0; L < bs2->L_size; ++L)
      for (int il = 0; il < in->l_size; ++il) {
        int nm = MIN(h->pr2.index_m_max[bs2->L[L]], h->pr2.index_m_max[in->l[il]]) + 1;
        for (int im = 0; im < in->m_size; ++im) {
          size_t q = (((size_t)J * bs2->L_size + L) * in->l_size + il) * in->m_size + im;
          if (im >= nm) { index_xmin_J[q] = -1; x_size_J[q] = 0; if (J_rows) J_rows[q] = NULL; if (ddJ_rows) ddJ_rows[q] = NULL; continue; }
          index_xmin_J[q] = bs2->index_xmin_J[J][L][il][im]; x_size_J[q] = bs2->x_size_J[J][L][il][im];
          if (J_rows) J_rows[q] = bs2->J_Llm_x[J][L][il][im];
          if (ddJ_rows) ddJ_rows[q] = (h->pr.bessels_interpolation == cubic_interpolation) ? bs2->ddJ_Llm_x[J][L][il][im] : NULL;
        }
      }
  return 0;
}

/* bessel2_convolution through the structures of the harness (bispectra2.c:1326) */
int harness_bessel2_convolution(harness *h, double *kk, double *delta_kk, int k_size, double *f, double *g, int index_l1, double r,
                                double *integral)
{
  ErrorMsg msg;
  return bessel2_convolution(&h->pr, &h->bs2, kk, delta_kk, k_size, f, g, index_l1, r, integral, msg) == _FAILURE_;
}

void harness_set_verbose(harness *h, int verbose) { h->tr2.transfer2_verbose = verbose; }
const struct bessels2 *harness_bessels2(harness *h) { return &h->bs2; }
/* pbs2->j_l1[index_l1][0 .. x_size_l1) and its first grid index: the table bessel2_j_for_l1 leaves (bessel2.c:1802-1810) */
int harness_j_l1_row(harness *h, int index_l1, double *out, int n_max, int *index_xmin_l1)
{
  if (index_l1 < 0 || index_l1 >= h->bs2.l1_size || h->bs2.j_l1 == NULL) return -1;
  const int n = h->bs2.x_size_l1[index_l1];
  if (index_xmin_l1) *index_xmin_l1 = h->bs2.index_xmin_l1[index_l1];
  for (int i = 0; i < n && i < n_max; ++i) out[i] = h->bs2.j_l1[index_l1][i];
  return n;
}
/* Every member of struct transfers2 that transfer2_indices_of_transfers (transfer2.c:866-1285) and transfer2_get_k3_sizes
   (:1752-1834) leave behind, as text: called on the structures right after transfer2_init (with load_transfers_from_disk
   the module returns after exactly those two steps, :151-157, and needs no GPU).  Two modules that agree on this digest
   agree on the index tables, the labels, the grid sizes and the physical ranges. */
long long harness_indices_digest(harness *h, char *out, long long cap)
{
  struct transfers2 *t = &h->tr2; const harness_input *in = &h->in;
  long long n = 0;
#define DIG(...) do { if (n < cap) n += snprintf(out + n, (size_t)(cap - n), __VA_ARGS__); } while (0)
  DIG("tt2_size=%d n_transfers=%d T=%d E=%d B=%d nzE=%d nzB=%d l_size=%d m_size=%d k3_size_max=%d has_cls=%d\n", t->tt2_size, t->n_transfers,
      in->has_T ? t->index_tt2_T : -1, in->has_E ? t->index_tt2_E : -1, in->has_B ? t->index_tt2_B : -1,
      in->has_E ? t->n_nonzero_transfers_E : -1, in->has_B ? t->n_nonzero_transfers_B : -1, t->l_size, t->m_size, t->k3_size_max, (int)t->has_cls);
  DIG("tau0=%.17g tau_rec=%.17g rs_rec=%.17g\n", t->tau0, t->tau_rec, t->rs_rec);
  for (int i = 0; i < t->l_size; ++i) DIG("l%d ", t->l[i]);
  for (int i = 0; i < t->m_size; ++i) DIG("m%d ", t->m[i]);
  DIG("\n");
  for (int il = 0; il < t->l_size; ++il) for (int im = 0; im < t->m_size; ++im) if (t->m[im] <= t->l[il]) DIG("%d ", t->lm_array[il][im]);
  DIG("\n");
  for (int tt = 0; tt < t->tt2_size; ++tt)
    DIG("%s:%d:%d:%d:%d ", t->tt2_labels[tt], t->corresponding_index_l[tt], t->corresponding_index_m[tt], t->index_tt2_monopole[tt], t->index_pt2_monopole[tt]);
  DIG("\n");
  for (int k1 = 0; k1 < in->k_size; ++k1) for (int k2 = 0; k2 <= k1; ++k2)
    DIG("%d,%d,%d,%.17g,%.17g ", t->k_size_k1k2[k1][k2], t->k_physical_start_k1k2[k1][k2], t->k_physical_size_k1k2[k1][k2],
        t->k_min_k1k2[k1][k2], t->k_max_k1k2[k1][k2]);
  DIG("\n");
#undef DIG
  return n;
}

/* transfer2_get_k3_list of the module under test for pair (index_k1, index_k2) (what bispectra2.c:1255 and spectra2.c:451
   call): returns ptr2->k_size_k1k2 and fills out[0 .. size); needs transfer2_init (or transfer2_indices_of_transfers) first */
int harness_k3_list(harness *h, int index_k1, int index_k2, double *out, int n_max)
{
  if (!h->transfer2_done) return -1;
  const int n = h->tr2.k_size_k1k2[index_k1][index_k2];
  if (n > n_max) return -n;
  int dump = 0;
  if (transfer2_get_k3_list(&h->pr, &h->pr2, &h->pt2, &h->bs, &h->bs2, &h->tr2, index_k1, index_k2, out, &dump) == _FAILURE_) return -1;
  return n;
}
int harness_l1_info(harness *h, int *l1, double *x_min_l1) { for (int i = 0; i < h->bs2.l1_size; ++i) { l1[i] = h->bs2.l1[i]; x_min_l1[i] = h->bs2.x_min_l1[i]; } return h->bs2.l1_size; }

void harness_destroy(harness *h)
{
  if (!h) return;
  char msg[256];
  harness_transfer2_free(h, msg, sizeof msg);
  if (h->bessel2_done) bessel2_free(&h->pr, &h->pr2, &h->bs, &h->bs2);
  else if (h->bs2.J_Llm_x) {
    for (int J = 0; J < h->in.J_size; ++J) {
      for (int L = 0; L < h->bs2.L_size; ++L) {
        for (int il = 0; il < h->in.l_size; ++il) {
          free(h->bs2.J_Llm_x[J][L][il]); free(h->bs2.ddJ_Llm_x[J][L][il]); free(h->bs2.index_xmin_J[J][L][il]); free(h->bs2.x_size_J[J][L][il]);
        }
        free(h->bs2.J_Llm_x[J][L]); free(h->bs2.ddJ_Llm_x[J][L]); free(h->bs2.index_xmin_J[J][L]); free(h->bs2.x_size_J[J][L]);
      }
      free(h->bs2.J_Llm_x[J]); free(h->bs2.ddJ_Llm_x[J]); free(h->bs2.index_xmin_J[J]); free(h->bs2.x_size_J[J]);
    }
    free(h->bs2.J_Llm_x); free(h->bs2.ddJ_Llm_x); free(h->bs2.index_xmin_J); free(h->bs2.x_size_J); free(h->bs2.L);
  }
  g_keep_sources = 0;
  free_sources(h);
  for (int k1 = 0; k1 < h->in.k_size; ++k1) { free(h->pt2.k3[k1]); free(h->pt2.k3_size[k1]); }
  free(h->pt2.k3); free(h->pt2.k3_size); free(h->pt2.has_allocated_sources);
  for (int L = 0; L <= h->in.L_max; ++L) free(h->pt2.lm_array[L]);
  free(h->pt2.lm_array); free(h->pr2.index_m_max);
  free(h);
}

/* one shot: everything above in the order of main/song.c.  returns 0 on success; `out` receives, per transfer type, the
   [k1][k2][k3] values */
int harness_transfer2(const harness_input *in, double *out, long long out_capacity, long long *out_sizes, int *labels_ok,
                      char *errmsg, int errlen)
{
  harness *h = harness_create(in, errmsg, errlen);
  if (!h) return 1;
  int rc = harness_transfer2_init(h, errmsg, errlen);
  if (!rc) rc = harness_read(h, out, out_capacity, out_sizes, labels_ok, errmsg, errlen);
  if (!rc) rc = harness_transfer2_free(h, errmsg, errlen);
  harness_destroy(h);
  return rc;
}
