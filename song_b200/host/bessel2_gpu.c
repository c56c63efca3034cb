/*
 * bessel2_gpu.c -- drop-in replacement of SONG's source/bessel2.c: same exported symbols and struct semantics
 * (reference include/bessel2.h:161-260), with the assembly of the projection functions J_Llm(x) -- the loop nest of
 * bessel2.c:315-374 over bessel2_J_for_Llm / bessel2_J_Llm -- running on the GPUs through the C-ABI of include/song_los.h.
 *
 * Compiled against SONG's OWN headers (-I<song>/include -I<song>/class.git/include); the stand-ins under
 * song_b200/host/compat/ are used only when the CLASS fork is absent, as in this repository's tests.
 *
 * bessel2_init  host: x grid (bessel2.c:960-985), l1 list (:832-943), j_l1(x) with its lower cut (:1669-1815) and the two
 *               3j symbols of every (J,L,l,m) (:1131-1167) -- bessel2_host.c, bessel_j / drc3jj from CLASS;
 *               device: J_Llm(x), its lower cut index_xmin_J by bisection (:1202-1342) and, for cubic interpolation, the
 *               second derivatives (:395-435) -- song_los_build_projection on EVERY GPU the transfer stage will use, so that
 *               transfer2_init (main/song.c:106) finds the tables where it needs them;
 *               host again: the tables are read back once into the ragged pbs2->J_Llm_x / index_xmin_J / x_size_J / x_min_J
 *               members, because they are part of struct bessels2 (bessel2.h:64-91).
 * bessel2_free  releases both the host members and the engines on the devices.
 * bessel2_convolution, bessel2_l1_at_x(_linear)  host functions on pbs2->j_l1, called by bispectra2.c:1326,2192.
 *
 * Environment: SONG_LOS_DEVICES=<n> (default: all visible GPUs).
 */
#include "bessel2.h"
#include "arrays.h"
#include "song_los.h"
#include "bessel2_host.h"
#include "los_glue.h"
#include "los_module.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define B2_CUBIC(ppr) ((ppr)->bessels_interpolation == cubic_interpolation)

/* index_m runs over the m's allowed for both L and l (bessel2.c:288) */
static int m_count(struct precision2 *ppr2, struct bessels *pbs, struct bessels2 *pbs2, int index_L, int index_l)
{
  return MIN(ppr2->index_m_max[pbs2->L[index_L]], ppr2->index_m_max[pbs->l[index_l]]) + 1;
}

/* ------------------------------------------------------------------------------------------------ */
/*  grids                                                                                            */
/* ------------------------------------------------------------------------------------------------ */
int bessel2_get_xx_list(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels *pbs,
                        struct bessels2 *pbs2)
{
  /* at least as far as the first-order Bessel functions reach; linear sampling with step xx_step (bessel2.c:976-981) */
  pbs2->xx_max = MAX(pbs2->xx_max, pbs->x_max);
  pbs2->xx_size = pbs2->xx_max / pbs2->xx_step + 1;
  class_alloc(pbs2->xx, pbs2->xx_size * sizeof(double), pbs2->error_message);
  lin_space(pbs2->xx, 0, pbs2->xx_max, pbs2->xx_size);
  return _SUCCESS_;
}

int bessel2_get_l1_list(struct precision *ppr, struct precision2 *ppr2, struct bessels *pbs, struct bessels2 *pbs2)
{
  /* every l1 within L_max + m_max of a member of pbs->l, without abs() on the lower side (bessel2.c:873-897) */
  const int reach = pbs2->L_max + ppr2->m_max_song;
  const int l1_top = pbs->l[pbs->l_size - 1] + reach;
  char *wanted;
  class_calloc(wanted, l1_top + 1, sizeof(char), pbs2->error_message);
  for (int index_l = 0; index_l < pbs->l_size; ++index_l)
    for (int l1 = MAX(0, pbs->l[index_l] - reach); l1 <= pbs->l[index_l] + reach; ++l1) wanted[l1] = 1;
  pbs2->l1_size = 0;
  for (int l1 = 0; l1 <= l1_top; ++l1) pbs2->l1_size += wanted[l1];
  class_alloc(pbs2->l1, pbs2->l1_size * sizeof(int), pbs2->error_message);
  class_alloc(pbs2->index_l1, (l1_top + 1) * sizeof(int), pbs2->error_message);      /* l1 -> position, -1 when absent (:930-939) */
  for (int l1 = 0, n = 0; l1 <= l1_top; ++l1) {
    pbs2->index_l1[l1] = wanted[l1] ? n : -1;
    if (wanted[l1]) pbs2->l1[n++] = l1;
  }
  free(wanted);
  return _SUCCESS_;
}

/* ------------------------------------------------------------------------------------------------ */
/*  per-element functions of the reference                                                           */
/* ------------------------------------------------------------------------------------------------ */
int bessel2_j_for_l1(struct precision *ppr, struct precision2 *ppr2, struct bessels *pbs, struct bessels2 *pbs2, int index_l1)
{
  /* one l1 of the table bessel2_init fills in one sweep (bessel2.c:1669-1815); host code, as in the reference */
  const int l1 = pbs2->l1[index_l1];
  double x_min = 0, j;
  pbs2->index_xmin_l1[index_l1] = 0;
  if (l1 != 0) {
    double lo = 0, hi = l1 + 0.5;                                     /* j_l1 is suppressed below its first peak */
    class_call(bessel_j(pbs, l1, hi, &j), pbs->error_message, pbs2->error_message);
    class_test(j < pbs2->j_l1_cut, pbs2->error_message, "in bisection, wrong initial guess for x_min_up.");
    while ((hi - lo) / lo > ppr->bessel_tol_x_min) {
      if ((hi - lo) < ppr->smallest_allowed_variation) { hi = lo; break; }
      const double mid = 0.5 * (hi + lo);
      class_call(bessel_j(pbs, l1, mid, &j), pbs->error_message, pbs2->error_message);
      if (j >= pbs2->j_l1_cut) hi = mid; else lo = mid;
    }
    x_min = lo;
    if (x_min < pbs2->xx_max) {
      int index_x = 0;
      while (pbs2->xx[index_x] < x_min) ++index_x;
      pbs2->index_xmin_l1[index_l1] = (index_x != 0) ? index_x - 1 : 0;
    }
  }
  const int negligible = (x_min >= pbs2->xx_max);
  if (negligible) { pbs2->index_xmin_l1[index_l1] = pbs2->xx_size - 1; pbs2->x_size_l1[index_l1] = 1; }
  else pbs2->x_size_l1[index_l1] = pbs2->xx_size - pbs2->index_xmin_l1[index_l1];
  class_alloc(pbs2->j_l1[index_l1], sizeof(double) * pbs2->x_size_l1[index_l1], pbs2->error_message);
  #pragma omp atomic
  pbs2->count_allocated_Js += pbs2->x_size_l1[index_l1];
  if (B2_CUBIC(ppr)) {
    class_alloc(pbs2->ddj_l1[index_l1], sizeof(double) * pbs2->x_size_l1[index_l1], pbs2->error_message);
    #pragma omp atomic
    pbs2->count_allocated_Js += pbs2->x_size_l1[index_l1];
  }
  if (negligible) { pbs2->x_min_l1[index_l1] = pbs2->xx_max; pbs2->j_l1[index_l1][0] = 0.; return _SUCCESS_; }
  pbs2->x_min_l1[index_l1] = pbs2->xx[pbs2->index_xmin_l1[index_l1]];
  for (int i = 0; i < pbs2->x_size_l1[index_l1]; ++i)
    class_call(bessel_j(pbs, l1, pbs2->xx[pbs2->index_xmin_l1[index_l1] + i], &pbs2->j_l1[index_l1][i]),
               pbs->error_message, pbs2->error_message);
  return _SUCCESS_;
}

#define ON_DEVICE(name) class_stop(pbs2->error_message, \
  "%s is evaluated inside the CUDA kernels of this build (song_los_build_projection, include/song_los.h); it is not available as a host call", name)

int bessel2_J_for_Llm(struct precision *ppr, struct precision2 *ppr2, struct bessels *pbs, struct bessels2 *pbs2,
                      int index_J, int index_L, int index_l, int index_m)
{ ON_DEVICE("bessel2_J_for_Llm"); }
int bessel2_J_Llm(struct precision2 *ppr2, struct bessels *pbs, struct bessels2 *pbs2,
                  enum projection_function_types projection_function, int L, int l, int m, int index_x,
                  struct J_Llm_data *bessel_3j_data, double *J_Llm_x)
{ ON_DEVICE("bessel2_J_Llm"); }

/* j_l1(x) from the table: position of x counted from the table's own first point (bessel2.c:647-657) */
static inline double j_l1_interpolated(const struct bessels2 *pbs2, int index_l1, double x, int cubic)
{
  const double x_min = pbs2->x_min_l1[index_l1], step = pbs2->xx_step;
  const int index_x = (int)((x - x_min) / step);
  const double a = (x_min + step * (index_x + 1) - x) / step;
  const double *j = pbs2->j_l1[index_l1];
  if (!cubic) return a * j[index_x] + (1. - a) * j[index_x + 1];
  const double *ddj = pbs2->ddj_l1[index_l1];
  return a * j[index_x] + (1. - a) * (j[index_x + 1] - a * ((a + 1.) * ddj[index_x] + (2. - a) * ddj[index_x + 1]) * step * step / 6.0);
}

int bessel2_l1_at_x(struct bessels2 *pbs2, double x, int index_l1, double *j_l1)
{
  class_test(index_l1 > pbs2->l1_size, pbs2->error_message, "index_l1=%d>l1_size=%d; increase l_max.", index_l1, pbs2->l1_size);
  if (x < pbs2->x_min_l1[index_l1]) { *j_l1 = 0; return _SUCCESS_; }
  class_test(x > pbs2->xx_max, pbs2->error_message, "x=%e>xx_max=%e in bessel structure", x, pbs2->xx_max);
  *j_l1 = j_l1_interpolated(pbs2, index_l1, x, _TRUE_);
  return _SUCCESS_;
}

int bessel2_l1_at_x_linear(struct bessels2 *pbs2, double x, int index_l1, double *j_l1)
{
  class_test(index_l1 > pbs2->l1_size, pbs2->error_message, "index_l1=%d>l1_size=%d; increase l_max.", index_l1, pbs2->l1_size);
  if (x < pbs2->x_min_l1[index_l1]) { *j_l1 = 0; return _SUCCESS_; }
  class_test(x > pbs2->xx_max, pbs2->error_message, "x=%e>xx_max=%e in bessel structure", x, pbs2->xx_max);
  *j_l1 = j_l1_interpolated(pbs2, index_l1, x, _FALSE_);
  return _SUCCESS_;
}

/* 1/2 sum_k delta_k k^2 f(k) g(k) j_l1(k r): the k integrals of the bispectrum (bessel2.c:499-605; bispectra2.c:1326,2192).
   Terms are skipped, not added as zeros, where f or g vanish or x lies below the table (the reference's statement order). */
int bessel2_convolution(struct precision *ppr, struct bessels2 *pbs2, double *kk, double *delta_kk, int k_size, double *f,
                        double *g, int index_l, double r, double *integral, ErrorMsg error_message)
{
#ifdef DEBUG
  class_test((index_l < 0) || (index_l >= pbs2->l1_size), error_message, "index_l=%d out of bounds (l1_size=%d)", index_l, pbs2->l1_size);
  class_test((kk[k_size - 1] * r) > pbs2->xx[pbs2->xx_size - 1], error_message,
             "r*k_max=%g is larger than x_max=%g (index_l1=%d,r=%g,k_max=%g)", kk[k_size - 1] * r, pbs2->xx[pbs2->xx_size - 1], index_l, r, kk[k_size - 1]);
#endif
  const int cubic = B2_CUBIC(ppr);
  const double x_min = pbs2->x_min_l1[index_l];
  double sum = 0;
  for (int index_k = 0; index_k < k_size; ++index_k) {
    const double f_in_k = f[index_k];
    if (f_in_k == 0.) continue;
    double g_in_k = 1.;
    if (g != NULL) { g_in_k = g[index_k]; if (g_in_k == 0.) continue; }
    const double k = kk[index_k], x = k * r;
    if (x < x_min) continue;
    const double j = j_l1_interpolated(pbs2, index_l, x, cubic);
    const double integrand = k * k * j * f_in_k * g_in_k;
    sum += integrand * delta_kk[index_k];
  }
  *integral = sum * 0.5;
  return _SUCCESS_;
}

/* ------------------------------------------------------------------------------------------------ */
/*  bessel2_init (reference: bessel2.c:75-468)                                                       */
/* ------------------------------------------------------------------------------------------------ */
#ifndef SONG_B200_COMPAT_HEADERS
/* CLASS's bessel_j behind the provider signature of bessel2_host.h */
static double los_class_bessel_j(int l, double x, void *user)
{
  double j = 0.;
  bessel_j((struct bessels *)user, l, x, &j);
  return j;
}
/* the fork's drc3jj (tools/slatec_3j_C.c) behind the 3j provider signature */
static int los_class_drc3jj(double l2, double l3, double m2, double m3, double *l1min, double *l1max, double *thrcof, int ndim, void *user)
{
  return drc3jj(l2, l3, m2, m3, l1min, l1max, thrcof, ndim, ((struct bessels2 *)user)->error_message) == _FAILURE_;
}
#endif

int bessel2_init(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels *pbs, struct bessels2 *pbs2)
{
  if ((pbs->l_max == 0) || ((ppt2->has_cmb_spectra == _FALSE_) && (ppt2->has_cmb_bispectra == _FALSE_))
      || ((ppr2->load_transfers_from_disk == _TRUE_) && (ppr->load_bispectra_from_disk == _TRUE_))) {   /* :88-96 */
    if (pbs2->bessels2_verbose > 0) printf("Second-order Bessel module skipped.\n");
    return _SUCCESS_;
  }
  if (pbs2->bessels2_verbose > 0) printf("Computing 2nd-order projection functions (B200 line-of-sight engine)\n");
#ifdef _OPENMP
  const double t_start = omp_get_wtime();
#define B2_LAP(what) do { if (pbs2->bessels2_verbose > 1) printf("     * %-52s %8.1f ms since the start of bessel2_init\n", what, (omp_get_wtime() - t_start) * 1e3); } while (0)
#else
#define B2_LAP(what) do { } while (0)
#endif

  pbs2->count_allocated_Js = 0;
  pbs2->j_l1_cut = ppr2->bessel_j_cut_song;                                         /* :106-107 */
  pbs2->J_Llm_cut = ppr2->bessel_J_cut_song;
  class_call(bessel2_get_xx_list(ppr, ppr2, ppt2, pbs, pbs2), pbs2->error_message, pbs2->error_message);

  pbs2->m_size = ppr2->m_size;                                                      /* :116-128 */
  class_alloc(pbs2->m, pbs2->m_size * sizeof(int), pbs2->error_message);
  for (int index_m = 0; index_m < pbs2->m_size; ++index_m) pbs2->m[index_m] = ppr2->m[index_m];
  pbs2->L_size = pbs2->L_max + 1;
  class_alloc(pbs2->L, pbs2->L_size * sizeof(int), pbs2->error_message);
  for (int index_L = 0; index_L < pbs2->L_size; ++index_L) pbs2->L[index_L] = index_L;

  /* which projection functions: TT for temperature; EE and, with non-scalar modes, EB for polarisation (:150-178) */
  int index_J = 0;
  int J_types[3] = {SONG_LOS_J_TT, SONG_LOS_J_EE, SONG_LOS_J_EB};
  pbs2->has_J_TT = pbs2->has_J_EE = pbs2->has_J_EB = _FALSE_;
  if (ppt2->has_cmb_temperature == _TRUE_) { pbs2->has_J_TT = _TRUE_; J_types[index_J] = SONG_LOS_J_TT; pbs2->index_J_TT = index_J++; }
  if ((ppt2->has_cmb_polarization_e == _TRUE_) || (ppt2->has_cmb_polarization_b == _TRUE_)) {
    pbs2->has_J_EE = _TRUE_; J_types[index_J] = SONG_LOS_J_EE; pbs2->index_J_EE = index_J++;
    if (ppr2->m_max_song > 0) { pbs2->has_J_EB = _TRUE_; J_types[index_J] = SONG_LOS_J_EB; pbs2->index_J_EB = index_J++; }
  }
  pbs2->J_size = index_J;
  if ((pbs2->bessels2_verbose > 0) && (ppr2->load_transfers_from_disk == _FALSE_))
    printf(" -> will compute size_J=%d projection functions\n", pbs2->J_size);

  /* ---- host half: l1 list, j_l1(x), 3j symbols ---- */
#ifndef SONG_B200_COMPAT_HEADERS
  /* a SONG tree: CLASS's own bessel_j, called point by point as bessel2_j_for_l1 does (bessel2.c:1706,1724,1804), so that
     pbs2->j_l1 -- and everything assembled from it -- holds the reference's numbers (bessel2_host.h) */
  song_host_bessel2_set_provider(los_class_bessel_j, pbs);
  song_host_bessel2_set_3j_provider(los_class_drc3jj, pbs2);
#endif
  song_bessel2_host *hb = NULL;
  class_test(song_host_bessel2_prepare(pbs->l, pbs->l_size, pbs2->m, pbs2->m_size, pbs2->L_max, pbs2->xx, pbs2->xx_size, pbs2->xx_max,
                                       pbs2->J_size, J_types, pbs2->j_l1_cut, ppr->bessel_tol_x_min, ppr->smallest_allowed_variation,
                                       &hb, pbs2->error_message, _ERRORMSGSIZE_) != 0,
             pbs2->error_message, "j_l1 tables: %.1500s", pbs2->error_message);
  /* the j_l1 members of struct bessels2 (bessel2.h:96-106) take over the arrays; bispectra2.c reads them */
  pbs2->l1_size = hb->l1_size; pbs2->l1 = hb->l1; pbs2->index_l1 = hb->index_l1;
  pbs2->j_l1 = hb->j_l1; pbs2->index_xmin_l1 = hb->index_xmin_l1; pbs2->x_size_l1 = hb->x_size_l1; pbs2->x_min_l1 = hb->x_min_l1;
  pbs2->count_allocated_Js += hb->count_allocated_Js;
  B2_LAP("l1 list, j_l1(x) and 3j symbols (host)");
  if (B2_CUBIC(ppr)) {                                                              /* natural splines of j_l1, :439-463 */
    class_alloc(pbs2->ddj_l1, pbs2->l1_size * sizeof(double *), pbs2->error_message);
    int abort = _FALSE_;
    #pragma omp parallel for schedule(dynamic)
    for (int index_l1 = 0; index_l1 < pbs2->l1_size; ++index_l1) {
      pbs2->ddj_l1[index_l1] = calloc(pbs2->x_size_l1[index_l1], sizeof(double));
      if (pbs2->ddj_l1[index_l1] == NULL) { abort = _TRUE_; continue; }
      if (pbs2->x_size_l1[index_l1] < 3) continue;                                   /* a negligible j_l1: one zero */
      ErrorMsg local;
      if (array_spline_table_one_column(pbs2->xx + pbs2->index_xmin_l1[index_l1], pbs2->x_size_l1[index_l1], pbs2->j_l1[index_l1],
                                        1, 0, pbs2->ddj_l1[index_l1], _SPLINE_NATURAL_, local) == _FAILURE_) abort = _TRUE_;
    }
    class_test(abort == _TRUE_, pbs2->error_message, "could not spline the j_l1 table");
    for (int index_l1 = 0; index_l1 < pbs2->l1_size; ++index_l1) pbs2->count_allocated_Js += pbs2->x_size_l1[index_l1];
  }
#define B2_DROP_HOST() do { hb->l1 = NULL; hb->index_l1 = NULL; hb->j_l1 = NULL; hb->index_xmin_l1 = NULL; hb->x_size_l1 = NULL; \
                            hb->x_min_l1 = NULL; hb->l1_size = 0; song_host_bessel2_free(hb); } while (0)

  if (ppr2->load_transfers_from_disk == _TRUE_) {                                   /* :243-249 */
    if (pbs2->bessels2_verbose > 0) printf(" -> No second-order projection functions needed.\n");
    B2_DROP_HOST();
    return _SUCCESS_;
  }

  /* ---- device half: one engine per GPU, J_Llm(x) assembled on each ---- */
  int n_dev = 0;
  if (song_los_device_count(&n_dev) != SONG_LOS_SUCCESS || n_dev < 1) {
    B2_DROP_HOST();
    class_stop(pbs2->error_message, "%.1500s", song_los_last_error(NULL));           /* no usable GPU: there is no CPU fallback */
  }
  {
    const char *v = getenv("SONG_LOS_DEVICES");
    const int want = (v && *v) ? atoi(v) : 0;
    if (want > 0 && want < n_dev) n_dev = want;
    if (n_dev > LOS_MAX_DEVICES) n_dev = LOS_MAX_DEVICES;
    if ((ppr2->load_sources_from_disk == _TRUE_) || (ppr2->store_sources_to_disk == _TRUE_) || (ppr2->store_transfers_to_disk == _TRUE_))
      n_dev = 1;                                                                     /* transfer2_init keeps the k1 order of the files */
  }
  los_config_buf cb;
  if (los_config_from_song(ppr, ppr2, ppt2, pbs, pbs2, 0., SONG_LOS_TAU_SOURCES, 0, &cb, pbs2->error_message) == _FAILURE_) {
    B2_DROP_HOST();
    return _FAILURE_;
  }
  los_engines_drop(pbs2);                                                            /* a second bessel2_init on the same structure */
  los_engine_set *set = los_engines_register(pbs2);
  if (set == NULL) { los_config_release(&cb); B2_DROP_HOST(); class_stop(pbs2->error_message, "too many bessels2 structures hold GPU engines"); }
  int failed = _FALSE_;
  ErrorMsg first = "";
  #pragma omp parallel for num_threads(n_dev) schedule(static, 1) if (n_dev > 1)
  for (int d = 0; d < n_dev; ++d) {
    song_los_config cfg = cb.cfg;
    cfg.device = d;
    ErrorMsg local = "";
    int rc = song_los_create(&cfg, &set->ctx[d]);
    if (rc != SONG_LOS_SUCCESS) snprintf(local, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(NULL));
    else if (song_host_bessel2_upload(set->ctx[d], hb, J_types, pbs2->J_Llm_cut, pbs2->xx_max, local, _ERRORMSGSIZE_) != 0) rc = SONG_LOS_FAILURE;
    if (rc != SONG_LOS_SUCCESS) {
      #pragma omp critical(b2_engine_error)
      { if (!failed) { failed = _TRUE_; snprintf(first, _ERRORMSGSIZE_, "GPU %d: %.1500s", d, local); } }
    }
  }
  set->n_dev = n_dev;
  los_config_release(&cb);
  B2_DROP_HOST();
  if (failed) { los_engines_drop(pbs2); class_stop(pbs2->error_message, "%.1900s", first); }
  B2_LAP("J_Llm(x) assembled on the GPUs");

  /* ---- the ragged members of struct bessels2, read back from GPU 0 (layout of bessel2.c:277-311) ---- */
  const int l_size = pbs->l_size, m_size = pbs2->m_size, L_size = pbs2->L_size;
  const size_t n_jidx = (size_t)pbs2->J_size * L_size * l_size * m_size;
  int *xmin, *xsize; double **rows, **ddrows = NULL;
  class_calloc(xmin, n_jidx, sizeof(int), pbs2->error_message);
  class_calloc(xsize, n_jidx, sizeof(int), pbs2->error_message);
  class_calloc(rows, n_jidx, sizeof(double *), pbs2->error_message);
  if (song_los_get_projection(set->ctx[0], xmin, xsize, NULL) != SONG_LOS_SUCCESS) {
    snprintf(pbs2->error_message, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(set->ctx[0]));
    free(xmin); free(xsize); free(rows); los_engines_drop(pbs2);
    return _FAILURE_;
  }
  size_t total = 0;
  class_alloc(pbs2->index_xmin_J, pbs2->J_size * sizeof(int ***), pbs2->error_message);
  class_alloc(pbs2->x_min_J, pbs2->J_size * sizeof(double ***), pbs2->error_message);
  class_alloc(pbs2->x_size_J, pbs2->J_size * sizeof(int ***), pbs2->error_message);
  class_alloc(pbs2->J_Llm_x, pbs2->J_size * sizeof(double ****), pbs2->error_message);
  if (B2_CUBIC(ppr)) class_alloc(pbs2->ddJ_Llm_x, pbs2->J_size * sizeof(double ****), pbs2->error_message);
  for (int J = 0; J < pbs2->J_size; ++J) {
    class_alloc(pbs2->index_xmin_J[J], L_size * sizeof(int **), pbs2->error_message);
    class_alloc(pbs2->x_min_J[J], L_size * sizeof(double **), pbs2->error_message);
    class_alloc(pbs2->x_size_J[J], L_size * sizeof(int **), pbs2->error_message);
    class_alloc(pbs2->J_Llm_x[J], L_size * sizeof(double ***), pbs2->error_message);
    if (B2_CUBIC(ppr)) class_alloc(pbs2->ddJ_Llm_x[J], L_size * sizeof(double ***), pbs2->error_message);
    for (int index_L = 0; index_L < L_size; ++index_L) {
      class_alloc(pbs2->index_xmin_J[J][index_L], l_size * sizeof(int *), pbs2->error_message);
      class_alloc(pbs2->x_min_J[J][index_L], l_size * sizeof(double *), pbs2->error_message);
      class_alloc(pbs2->x_size_J[J][index_L], l_size * sizeof(int *), pbs2->error_message);
      class_alloc(pbs2->J_Llm_x[J][index_L], l_size * sizeof(double **), pbs2->error_message);
      if (B2_CUBIC(ppr)) class_alloc(pbs2->ddJ_Llm_x[J][index_L], l_size * sizeof(double **), pbs2->error_message);
      for (int index_l = 0; index_l < l_size; ++index_l) {
        const int nm = m_count(ppr2, pbs, pbs2, index_L, index_l);
        class_alloc(pbs2->index_xmin_J[J][index_L][index_l], nm * sizeof(int), pbs2->error_message);
        class_alloc(pbs2->x_min_J[J][index_L][index_l], nm * sizeof(double), pbs2->error_message);
        class_alloc(pbs2->x_size_J[J][index_L][index_l], nm * sizeof(int), pbs2->error_message);
        class_calloc(pbs2->J_Llm_x[J][index_L][index_l], nm, sizeof(double *), pbs2->error_message);
        if (B2_CUBIC(ppr)) class_calloc(pbs2->ddJ_Llm_x[J][index_L][index_l], nm, sizeof(double *), pbs2->error_message);
        for (int index_m = 0; index_m < nm; ++index_m) {
          const size_t q = (((size_t)J * L_size + index_L) * l_size + index_l) * m_size + index_m;
          pbs2->index_xmin_J[J][index_L][index_l][index_m] = xmin[q];
          pbs2->x_size_J[J][index_L][index_l][index_m] = xsize[q];
          pbs2->x_min_J[J][index_L][index_l][index_m] = (xmin[q] == pbs2->xx_size - 1) ? pbs2->xx_max : pbs2->xx[xmin[q]];   /* :1378-1382 */
          total += (size_t)xsize[q];
        }
      }
    }
  }
  /* the x level: ONE array for all tables, in the loop order above (the reference calloc's tens of thousands of rows,
     bessel2.c:1362); [0][0][0][0] is its base, which is what bessel2_free releases */
  double *values, *ddvalues = NULL;
  class_calloc(values, total > 0 ? total : 1, sizeof(double), pbs2->error_message);
  if (B2_CUBIC(ppr)) {
    class_calloc(ddvalues, total > 0 ? total : 1, sizeof(double), pbs2->error_message);
    class_calloc(ddrows, n_jidx, sizeof(double *), pbs2->error_message);
  }
  pbs2->x_size_max_J = 0;
  size_t at = 0;
  for (int J = 0; J < pbs2->J_size; ++J)
    for (int index_L = 0; index_L < L_size; ++index_L)
      for (int index_l = 0; index_l < l_size; ++index_l)
        for (int index_m = 0; index_m < m_count(ppr2, pbs, pbs2, index_L, index_l); ++index_m) {
          const size_t q = (((size_t)J * L_size + index_L) * l_size + index_l) * m_size + index_m;
          pbs2->J_Llm_x[J][index_L][index_l][index_m] = rows[q] = values + at;
          if (B2_CUBIC(ppr)) pbs2->ddJ_Llm_x[J][index_L][index_l][index_m] = ddrows[q] = ddvalues + at;
          at += (size_t)xsize[q];
          pbs2->x_size_max_J = MAX(pbs2->x_size_max_J, xsize[q]);                   /* :376-382 */
        }
  pbs2->count_allocated_Js += (long int)total * (B2_CUBIC(ppr) ? 2 : 1);
  int rc = song_los_get_projection(set->ctx[0], NULL, NULL, rows);
  if (rc == SONG_LOS_SUCCESS && B2_CUBIC(ppr)) rc = song_los_get_projection_dd(set->ctx[0], ddrows);
  free(xmin); free(xsize); free(rows); free(ddrows);
  if (rc != SONG_LOS_SUCCESS) {
    snprintf(pbs2->error_message, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(set->ctx[0]));
    los_engines_drop(pbs2);
    return _FAILURE_;
  }
  /* the kernels read the dense tables; the undensified copy only served the read-back */
  for (int d = 0; d < set->n_dev; ++d) song_los_drop_projection_rows(set->ctx[d]);
  B2_LAP("pbs2->J_Llm_x read back");
  if (pbs2->bessels2_verbose > 1)
    printf(" -> memory in use to store the 2nd-order projection functions: ~ %3g MB\n", 8 * pbs2->count_allocated_Js / 1e6);
  return _SUCCESS_;
}

/* ------------------------------------------------------------------------------------------------ */
/*  bessel2_free (reference: bessel2.c:735-818)                                                      */
/* ------------------------------------------------------------------------------------------------ */
int bessel2_free(struct precision *ppr, struct precision2 *ppr2, struct bessels *pbs, struct bessels2 *pbs2)
{
  los_engines_drop(pbs2);                                    /* CUDA contexts' tables and buffers on every GPU */
  if (ppr2->load_transfers_from_disk == _FALSE_) {
    /* the x level is one array per table kind (see bessel2_init) */
    free(pbs2->J_Llm_x[0][0][0][0]);
    if (B2_CUBIC(ppr)) free(pbs2->ddJ_Llm_x[0][0][0][0]);
    for (int J = 0; J < pbs2->J_size; ++J) {
      for (int index_L = 0; index_L < pbs2->L_size; ++index_L) {
        for (int index_l = 0; index_l < pbs->l_size; ++index_l) {
          free(pbs2->index_xmin_J[J][index_L][index_l]); free(pbs2->x_size_J[J][index_L][index_l]);
          free(pbs2->x_min_J[J][index_L][index_l]); free(pbs2->J_Llm_x[J][index_L][index_l]);
          if (B2_CUBIC(ppr)) free(pbs2->ddJ_Llm_x[J][index_L][index_l]);
        }
        free(pbs2->index_xmin_J[J][index_L]); free(pbs2->x_size_J[J][index_L]);
        free(pbs2->x_min_J[J][index_L]); free(pbs2->J_Llm_x[J][index_L]);
        if (B2_CUBIC(ppr)) free(pbs2->ddJ_Llm_x[J][index_L]);
      }
      free(pbs2->index_xmin_J[J]); free(pbs2->x_size_J[J]); free(pbs2->x_min_J[J]); free(pbs2->J_Llm_x[J]);
      if (B2_CUBIC(ppr)) free(pbs2->ddJ_Llm_x[J]);
    }
    free(pbs2->index_xmin_J); free(pbs2->x_size_J); free(pbs2->x_min_J); free(pbs2->J_Llm_x);
    if (B2_CUBIC(ppr)) free(pbs2->ddJ_Llm_x);
  }
  free(pbs2->l1); free(pbs2->index_l1); free(pbs2->xx);
  free(pbs2->x_size_l1); free(pbs2->x_min_l1); free(pbs2->index_xmin_l1);
  for (int index_l1 = 0; index_l1 < pbs2->l1_size; ++index_l1) free(pbs2->j_l1[index_l1]);
  free(pbs2->j_l1);
  if (B2_CUBIC(ppr)) {
    for (int index_l1 = 0; index_l1 < pbs2->l1_size; ++index_l1) free(pbs2->ddj_l1[index_l1]);
    free(pbs2->ddj_l1);
  }
  free(pbs2->L); free(pbs2->m);
  return _SUCCESS_;
}
