/*
 * ref_bispectra_harness.c -- variant oracle/_ref/libsong_ref_bisp.so: the reference's own
 * bispectra2_intrinsic_integrate_over_k3 (source/bispectra2.c:1107-1439) on the structures of the test harness.
 * The harness (tests/harness/dropin_harness.c, included here for its struct) has run bessel2_init and transfer2_init of
 * the reference; this file fills the few members of struct bispectra and struct bispectra_workspace_intrinsic the k3
 * integral reads -- what bispectra2_indices and bispectra2_intrinsic_workspace_init would leave there -- and calls it.
 * The remaining entry points of bispectra2.c reach into parts of the absent CLASS fork (r grid, 6j symbols, disk
 * format of the bispectra): stubs that fail loudly, none of them on the path of the k3 integral.  TEST INFRASTRUCTURE ONLY.
 */
#include "../../tests/harness/dropin_harness.c"
#include "bispectra2.h"
#include "spectra2.h"

#define REF_STUB(name) int name() { fprintf(stderr, "oracle/_ref: " #name " of the absent CLASS fork was called\n"); abort(); return _FAILURE_; }
REF_STUB(bessel_convolution)
REF_STUB(bispectra_get_r_grid)
REF_STUB(bispectra_store_to_disk)
REF_STUB(drc6j)
REF_STUB(is_triangular_int)

/* ---- spectra2_cls (source/spectra2.c:241-570) ----
   Two calls into the absent fork: primordial_spectrum_at_k (primordial.c) is answered from a table the test provides for
   the nodes of ppt2->k -- the only arguments spectra2_cls passes (:403, :437); spectra_cls_spline (spectra.c), the
   post-processing after the sum (:592-600), does nothing here. */
static const double *g_prim_k = NULL, *g_prim_p = NULL; static int g_prim_n = 0;
int primordial_spectrum_at_k(struct primordial *ppm, int index_md, int mode, double k, double *pk)
{
  (void)index_md; (void)mode;
  for (int i = 0; i < g_prim_n; ++i) if (g_prim_k[i] == k) { *pk = g_prim_p[i]; return _SUCCESS_; }
  snprintf(ppm->error_message, _ERRORMSGSIZE_, "primordial_spectrum_at_k stand-in: k=%.17g is not a node of the table", k);
  return _FAILURE_;
}
int spectra_cls_spline() { return _SUCCESS_; }

/* out[index_ct * l_size + index_l], index_ct: 0 = tt2, 1 = ee2, 2 = te2 (those requested, in this order) */
int ref_spectra2_cls(harness *h, const double *primordial, int has_tt2, int has_ee2, int has_te2, int rescale_cmb_sources,
                     double *out, char *errmsg, int errlen)
{
  struct spectra sp; struct primordial pm; struct bispectra bi;
  memset(&sp, 0, sizeof sp); memset(&pm, 0, sizeof pm); memset(&bi, 0, sizeof bi);
  g_prim_k = h->pt2.k; g_prim_p = primordial; g_prim_n = h->pt2.k_size;
  int ic_size[1] = { 1 };
  h->pt.index_md_scalars = 0; h->pt.has_scalars = _TRUE_; h->pt.md_size = 1; h->pt.ic_size = ic_size;
  h->pt2.rescale_cmb_sources = rescale_cmb_sources;
  int ct = 0;
  sp.has_tt2 = has_tt2; if (has_tt2) sp.index_ct_tt2 = ct++;
  sp.has_ee2 = has_ee2; if (has_ee2) sp.index_ct_ee2 = ct++;
  sp.has_te2 = has_te2; if (has_te2) sp.index_ct_te2 = ct++;
  sp.ct_size = ct;
  for (int i = 0; i < ct; ++i) sp.cl_type[i] = second_order;
  const int l_size = h->tr2.l_size;
  double *cl0 = calloc((size_t)l_size * ct, sizeof(double));      /* zeroed as spectra2_init does (:609-620) */
  double *cl[1] = { cl0 };
  sp.cl = cl;
  /* the tail of spectra2_cls (:606-628) clears C_l above the largest second-order l of the first-order l list: give it
     that list equal to the second-order one, so that nothing is above */
  int l_max_ct0[_MAX_NUM_CT_]; int *l_max_ct[1] = { l_max_ct0 }; int l_size_md[1] = { l_size };
  double *l_as_double = malloc(sizeof(double) * l_size);
  for (int il = 0; il < l_size; ++il) l_as_double[il] = h->tr2.l[il];
  sp.l_max_ct = l_max_ct; sp.l_size = l_size_md; sp.l = l_as_double;
  int rc = spectra2_cls(&h->pr, &h->pr2, &h->ba, &h->th, &h->pt, &h->pt2, &h->bs, &h->bs2, &h->tr, &h->tr2, &pm, NULL, &bi, &sp);
  if (rc == _FAILURE_) snprintf(errmsg, errlen, "%s", sp.error_message);
  else
    for (int il = 0; il < l_size; ++il)
      for (int i = 0; i < ct; ++i) out[i * l_size + il] = cl0[il * ct + i];
  free(cl0); free(l_as_double);
  h->pt.ic_size = NULL;
  return rc == _FAILURE_;
}

/* out[(index_l3 * r_size + index_r) * n_pairs + pair], pair = index_k1 (index_k1 + 1) / 2 + index_k2 */
int ref_integrate_over_k3(harness *h, int index_tt2_k3, int index_M3, int offset_L3, int abs_M3, double M3_coefficient,
                          const double *r, int r_size, double *out, char *errmsg, int errlen)
{
  struct bispectra bi; struct primordial pm; struct bispectra_workspace_intrinsic wb;
  memset(&bi, 0, sizeof bi); memset(&pm, 0, sizeof pm); memset(&wb, 0, sizeof wb);
  bi.l = h->tr2.l; bi.l_size = h->tr2.l_size; bi.bispectra_verbose = 0;
  wb.k_smooth_grid = h->pt2.k; wb.k_smooth_size = h->pt2.k_size;
  wb.r = (double *)r; wb.r_size = r_size;
  wb.abs_M3 = abs_M3;
  wb.M3_coefficient[index_M3] = M3_coefficient;
  int n_threads = 1;
#ifdef _OPENMP
  n_threads = omp_get_max_threads();
#endif
  wb.k3_grid = malloc(sizeof(double *) * n_threads); wb.delta_k3 = malloc(sizeof(double *) * n_threads);
  for (int t = 0; t < n_threads; ++t) {
    wb.k3_grid[t] = malloc(sizeof(double) * h->tr2.k3_size_max); wb.delta_k3[t] = malloc(sizeof(double) * h->tr2.k3_size_max);
  }
  int rc = bispectra2_intrinsic_integrate_over_k3(&h->pr, &h->pr2, &h->pt, &h->pt2, &h->bs, &h->bs2, &h->tr, &h->tr2, &pm, &bi,
                                                  index_tt2_k3, index_M3, offset_L3, &wb);
  if (rc == _FAILURE_) snprintf(errmsg, errlen, "%s", bi.error_message);
  else {
    const int nk = wb.k_smooth_size; const long n_pairs = (long)nk * (nk + 1) / 2;
    for (int il = 0; il < bi.l_size; ++il)
      for (int ir = 0; ir < r_size; ++ir)
        for (int k1 = 0, p = 0; k1 < nk; ++k1)
          for (int k2 = 0; k2 <= k1; ++k2, ++p)
            out[((long)il * r_size + ir) * n_pairs + p] = wb.integral_over_k3[il][ir][k1][k2];
  }
  if (wb.integral_over_k3)
    for (int il = 0; il < bi.l_size; ++il) {
      if (!wb.integral_over_k3[il]) continue;
      for (int ir = 0; ir < r_size; ++ir) {
        if (!wb.integral_over_k3[il][ir]) continue;
        for (int k1 = 0; k1 < wb.k_smooth_size; ++k1) free(wb.integral_over_k3[il][ir][k1]);
        free(wb.integral_over_k3[il][ir]);
      }
      free(wb.integral_over_k3[il]);
    }
  free(wb.integral_over_k3);
  for (int t = 0; t < n_threads; ++t) { free(wb.k3_grid[t]); free(wb.delta_k3[t]); }
  free(wb.k3_grid); free(wb.delta_k3);
  return rc == _FAILURE_;
}
