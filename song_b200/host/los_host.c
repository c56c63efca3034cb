/*
 * los_host.c -- host-side k3 integration grids of the transfer2 drop-in (see los_host.h).
 *
 * The grid is produced by one generator, k3_walk(), used both to count the points
 * (transfer2_get_k3_size, reference source/transfer2.c:1329-1551) and to write them
 * (transfer2_get_k3_list, :1566-1737).  The floating-point recurrence `k += k_step` and every
 * comparison are the reference's, so the grids are bit-identical to what bispectra2.c:1255 and
 * spectra2.c:451 re-derive.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "los_host.h"

#define SONG_PI 3.1415926535897932384626433832795e0

static int host_fail(char *errmsg, int errlen, const char *text)
{
  if (errmsg && errlen > 0) snprintf(errmsg, (size_t)errlen, "%s", text);
  return 1;
}

/* linear step of the fine grid, transfer2.c:1428-1448 */
static double k3_linear_step(const song_k3_params *p)
{
  if (p->k_sampling == 1) {
    if (p->old_run_step) return 2. * SONG_PI / p->rs_rec * p->q_linstep_song;
    return 2. * SONG_PI / (p->tau0 - p->tau_rec) * p->q_linstep_song;
  }
  return p->xx_step / p->tau0;
}

/* limits of the grid including the flat-extrapolation wings, transfer2.c:1354-1404 */
static void k3_limits(const song_k3_params *p, double k_min_pt, double k_max_pt,
                      double *k_min_tr, double *k_max_tr)
{
  double lo = k_min_pt, hi = k_max_pt;
  if (p->k3_extrapolation != 0) {
    double physical = k_max_pt - k_min_pt;
    double want_right = 2. * SONG_PI * p->extra_k3_oscillations_right / (p->tau0 - p->tau_rec);
    double want_left = 2. * SONG_PI * p->extra_k3_oscillations_left / (p->tau0 - p->tau_rec);
    if (physical < want_right) hi += (want_right - physical);
    if (hi > p->k_last) hi = p->k_last;
    if (physical < want_left) lo -= (want_left - physical);
    double floor_k = (p->k_first < k_min_pt) ? p->k_first : k_min_pt;
    if (lo < floor_k) lo = floor_k;
    if (p->extra_k3_oscillations_left < 0) lo = p->k_first;
    if (p->extra_k3_oscillations_right < 0) hi = p->k_last;
  }
  *k_min_tr = lo;
  *k_max_tr = hi;
}

/*
 * Emits the grid: [left wing by repeated k += step, snapped onto k_min_pt] + [source points
 * while their spacing is finer than step] + [k += step up to k_max_tr].  `out` may be NULL
 * (count only).  Returns the number of emitted points and the physical bookkeeping.
 */
static int k3_walk(const song_k3_params *p, const double *k3_pt, int np,
                   double k_min_tr, double k_max_tr, double *out, int cap,
                   int *n_out, int *phys_start, int *phys_size, int *n_pt_used)
{
  const double k_min_pt = k3_pt[0], k_max_pt = k3_pt[np - 1];
  const double step = k3_linear_step(p);
  if (step == 0) return 1;

  int n = 0;
  double k = k_min_tr;
#define EMIT(v) do { if (out && n < cap) out[n] = (v); ++n; } while (0)
  EMIT(k);
  if (p->k3_extrapolation != 0) {
    if (step < p->smallest_allowed_variation) return 2;
    while (k < k_min_pt) { k += step; EMIT(k); }                 /* :1486 / :1638 */
    if (k != k_min_pt) { k = k_min_pt; if (out && n - 1 < cap) out[n - 1] = k; }
  }
  *phys_start = n - 1;                                           /* :1498 */

  int ipt = 1;
  while (ipt < np && (k3_pt[ipt] - k) < step) {                  /* :1502 / :1653 */
    k = k3_pt[ipt];
    EMIT(k);
    ++ipt;
  }
  *n_pt_used = ipt;
  int physical = ipt;                                            /* :1509 */

  while (k < k_max_tr) {                                         /* :1513 / :1663 */
    k += step;
    EMIT(k);
    if (k <= k_max_pt && k >= k_min_pt) ++physical;
  }
  if (k > k_max_tr) --n;                                         /* :1525 */
#undef EMIT
  *n_out = n;
  *phys_size = physical;
  return 0;
}

int song_host_k3_size(const song_k3_params *p, const double *k3_pt, int np,
                      song_k3_range *r, char *errmsg, int errlen)
{
  char buf[256];
  if (np < 2) return host_fail(errmsg, errlen, "k3 grid of the sources has less than 2 points");
  k3_limits(p, k3_pt[0], k3_pt[np - 1], &r->k_min, &r->k_max);
  if (r->k_max > p->xx_max / p->tau0) {                          /* :1408 */
    snprintf(buf, sizeof buf, "xx_max is not large enough. The projection functions J(k*(tau0-tau)) do not cover "
             "the needed k-range (k_max=%g,tau0=%g,xx_max=%g). Maybe you are using too much extrapolation.",
             r->k_max, p->tau0, p->xx_max);
    return host_fail(errmsg, errlen, buf);
  }
  int used = 0;
  int rc = k3_walk(p, k3_pt, np, r->k_min, r->k_max, NULL, 0, &r->k_size, &r->k_physical_start,
                   &r->k_physical_size, &used);
  if (rc == 1) return host_fail(errmsg, errlen, "stop to avoid infinite loop");             /* :1450 */
  if (rc == 2) return host_fail(errmsg, errlen, "stopping to avoid segmentation fault, k3 step below smallest_allowed_variation");
  if (r->k_size < 2)                                                                        /* :1531 */
    return host_fail(errmsg, errlen, "found less than 2 valid values for the k-grid; k3_size_min must be at least 2");
  if (p->k3_extrapolation == 0 && (r->k_physical_size != r->k_size || r->k_physical_start != 0))
    return host_fail(errmsg, errlen, "when no extrapolation is requested, all k's should be physical, "
                     "i.e. satisfy the triangular condition.");                              /* :1538-1546 */
  return 0;
}

int song_host_k3_list(const song_k3_params *p, const double *k3_pt, int np,
                      const song_k3_range *r, double *k3, int *last_used_index_pt,
                      char *errmsg, int errlen)
{
  int n = 0, ps = 0, pz = 0, used = 0;
  int rc = k3_walk(p, k3_pt, np, r->k_min, r->k_max, k3, r->k_size, &n, &ps, &pz, &used);
  if (rc != 0) return host_fail(errmsg, errlen, "stop to avoid infinite loop");
  if (last_used_index_pt) *last_used_index_pt = used;
  if (n != r->k_size)
    return host_fail(errmsg, errlen, "bug in k list calculation: size differs from transfer2_get_k3_size");
  if (k3[r->k_size - 1] > r->k_max)                                                         /* :1670 */
    return host_fail(errmsg, errlen, "bug in k list calculation, k_max_transfers2 is larger than k_max_tr");
  for (int i = 0; i + 1 < r->k_size; ++i)                                                   /* :1727 */
    if (k3[i] >= k3[i + 1])
      return host_fail(errmsg, errlen, "bug in k list calculation, k3-grid is not strictly ascending");
  return 0;
}

int song_host_k3_sizes(const song_k3_params *p, int n_pairs, const int *k3_pt_size,
                       const int64_t *pt_off, const double *k3_pt, song_k3_range *ranges,
                       int *k3_size_max, int64_t *k3_size_sum, char *errmsg, int errlen)
{
  int mx = 0;
  int64_t sum = 0;
  for (int i = 0; i < n_pairs; ++i) {
    if (song_host_k3_size(p, k3_pt + pt_off[i], k3_pt_size[i], &ranges[i], errmsg, errlen)) return 1;
    if (ranges[i].k_size > mx) mx = ranges[i].k_size;
    sum += ranges[i].k_size;
  }
  if (k3_size_max) *k3_size_max = mx;                                                       /* :1816-1826 */
  if (k3_size_sum) *k3_size_sum = sum;
  return 0;
}

int song_host_k3_lists(const song_k3_params *p, int n_pairs, const int *k3_pt_size,
                       const int64_t *pt_off, const double *k3_pt, const song_k3_range *ranges,
                       const int64_t *tr_off, double *k3_tr, char *errmsg, int errlen)
{
  for (int i = 0; i < n_pairs; ++i)
    if (song_host_k3_list(p, k3_pt + pt_off[i], k3_pt_size[i], &ranges[i], k3_tr + tr_off[i], NULL,
                          errmsg, errlen))
      return 1;
  return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/*  multi-GPU partition of the (k1,k2) work: static LPT                               */
/* ------------------------------------------------------------------------------------------------ */
int song_host_lpt_schedule(const int64_t *cost, int n, int n_dev, int *device, int64_t *load)
{
  if (!cost || !device || n < 0 || n_dev < 1) return 1;
  int64_t *l = (int64_t *)calloc((size_t)n_dev, sizeof(int64_t));
  int *order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (!l || !order) { free(l); free(order); return 1; }
  for (int i = 0; i < n; ++i) order[i] = i;
  /* stable insertion sort by decreasing cost (n = number of k1 batches, small) -- for long pair lists a
     merge sort keeps the same order */
  if (n <= 2048) {
    for (int i = 1; i < n; ++i)
      for (int j = i; j > 0 && cost[order[j]] > cost[order[j - 1]]; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
  } else {
    int *tmp = (int *)malloc(sizeof(int) * (size_t)n);
    if (!tmp) { free(l); free(order); return 1; }
    for (int w = 1; w < n; w *= 2) {
      for (int lo = 0; lo < n; lo += 2 * w) {
        int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
        int a = lo, b = mid, o = lo;
        while (a < mid && b < hi) tmp[o++] = (cost[order[b]] > cost[order[a]]) ? order[b++] : order[a++];
        while (a < mid) tmp[o++] = order[a++];
        while (b < hi) tmp[o++] = order[b++];
      }
      int *sw = order; order = tmp; tmp = sw;
    }
    free(tmp);
  }
  for (int i = 0; i < n; ++i) {
    int best = 0;
    for (int d = 1; d < n_dev; ++d) if (l[d] < l[best]) best = d;
    device[order[i]] = best; l[best] += cost[order[i]];
  }
  if (load) for (int d = 0; d < n_dev; ++d) load[d] = l[d];
  free(l); free(order);
  return 0;
}

/* Static schedule at (k1,k2)-pair granularity.  The pair sequence (reference loop order, k1 major) is cut into
   n_dev * chunks_per_dev contiguous chunks of about equal cost and the chunks are assigned by LPT: one chunk per
   device gives every GPU a band of consecutive k1 levels (sources are loaded and freed level by level, transfer2.c:332-344,
   553); several chunks per device give every GPU the same mix of cheap low-k and expensive high-k pairs, so that the
   balance does not hinge on the cost model, while the results stay contiguous runs of the transfers_%03d.dat order. */
int song_host_schedule_pairs(const int64_t *cost, int n_pairs, int n_dev, int chunks_per_dev, int *device_of_pair, int64_t *load)
{
  if (!cost || !device_of_pair || n_pairs < 0 || n_dev < 1) return 1;
  if (chunks_per_dev < 1) chunks_per_dev = 1;
  int n_chunks = n_dev * chunks_per_dev;
  if (n_chunks > n_pairs) n_chunks = n_pairs > 0 ? n_pairs : 1;
  long double total = 0;
  for (int p = 0; p < n_pairs; ++p) total += (long double)cost[p];
  int64_t *ccost = (int64_t *)calloc((size_t)n_chunks, sizeof(int64_t));
  int *cdev = (int *)calloc((size_t)n_chunks, sizeof(int));
  int *chunk_of = (int *)malloc(sizeof(int) * (size_t)(n_pairs > 0 ? n_pairs : 1));
  if (!ccost || !cdev || !chunk_of) { free(ccost); free(cdev); free(chunk_of); return 1; }
  /* cut where the running cost crosses the next multiple of total/n_chunks (prefix rule: no drift) */
  long double acc = 0;
  int c = 0;
  for (int p = 0; p < n_pairs; ++p) {
    /* the pairs left must still fill the chunks left */
    while (c + 1 < n_chunks && ccost[c] > 0 && (acc >= total * (c + 1) / n_chunks || n_pairs - p <= n_chunks - 1 - c)) ++c;
    chunk_of[p] = c; ccost[c] += cost[p]; acc += (long double)cost[p];
  }
  int rc = song_host_lpt_schedule(ccost, n_chunks, n_dev, cdev, load);
  if (!rc) for (int p = 0; p < n_pairs; ++p) device_of_pair[p] = cdev[chunk_of[p]];
  free(ccost); free(cdev); free(chunk_of);
  return rc;
}

/* Cost model of the static schedule.  Every k3 value of a pair costs tt2_size integrals over the same tau and L ranges
   (SURVEY.md 8e), but the tensor-core kernel skips the multipole columns whose J_Llm(x) is still below the cut at
   x = k3 (tau0 - tau) << l: about min(1, k3 tau0 / l_max) of the columns are active.  cost = k3 values x (0.65 + 0.35 x
   active fraction at the middle of the pair's k3 range), in units of 1/1024 k3 value.  The weights are measured: on the
   lmax3000 run the lowest-k quarter of the pairs costs 0.74 of the rest per k3 value (a stronger correction, 0.35 + active,
   left one of two GPUs with 1089 ms of kernels against 901 ms). */
int song_host_pair_costs(int n_pairs, const int *k3_tr_size, const double *k_min, const double *k_max, double tau0, int l_max,
                         int64_t *cost)
{
  if (!k3_tr_size || !cost || n_pairs < 0) return 1;
  for (int p = 0; p < n_pairs; ++p) {
    double active = 1.0;
    if (k_min && k_max && l_max > 0) {
      active = 0.5 * (k_min[p] + k_max[p]) * tau0 / (double)l_max;
      if (active > 1.0) active = 1.0;
      if (active < 0.0) active = 0.0;
    }
    cost[p] = (int64_t)((double)k3_tr_size[p] * (0.65 + 0.35 * active) * 1024.0);
    if (cost[p] < 1) cost[p] = 1;
  }
  return 0;
}

/* ---- k3 bands (los_host.h) ---- */
/* Cost model of the FAST path, microseconds, fitted to the per-GPU kernel times of configs[3] on 8 GPUs (2 pair groups x
   4 bands) and cross-checked against the one-, two- and eight-GPU runs by pairs (all within 3 %):
     a k3 value costs   c_interp + (c0 + c1 a) (1 + rho0 / rho)
       a   = min(1, k3 tau0 / l_max), the share of multipole columns not skipped at x << l,
       rho = k3 values of the GPU per row of the projection table around this k3 (the DMMA kernel's tile holds the rows of
             192 consecutive values: the fewer values share a row, the more rows a tile fetches and the more of its 8-value
             groups straddle two rows),
     a coarse k3 point of a pair the GPU touches costs c_point (its spline, and its row staged by the interpolation kernel). */
static const double MODEL_C_INTERP = 0.0167, MODEL_C0 = 0.070, MODEL_C1 = 0.098, MODEL_RHO0 = 31.0, MODEL_C_POINT = 0.245;
enum { MODEL_DENSITY_ROWS = 50 };          /* the density is estimated over windows of this many table rows */

static double model_active(double k3, double tau0, int l_max)
{
  double a = (l_max > 0) ? k3 * tau0 / (double)l_max : 1.0;
  return a > 1.0 ? 1.0 : (a < 0.0 ? 0.0 : a);
}

/* density table of the selected pairs: values per table row in windows of MODEL_DENSITY_ROWS rows; row of a value =
   k3 * rows_per_k.  Returns the table (caller frees) and its length. */
static double *model_density(int n_pairs, const int *k3_tr_size, const int64_t *tr_off, const double *k3_tr,
                             const unsigned char *pair_in_group, double rows_per_k, int64_t *n_win_out)
{
  double kmax = 0.;
  for (int p = 0; p < n_pairs; ++p)
    if ((!pair_in_group || pair_in_group[p]) && k3_tr_size[p] > 0 && k3_tr[tr_off[p] + k3_tr_size[p] - 1] > kmax)
      kmax = k3_tr[tr_off[p] + k3_tr_size[p] - 1];
  const int64_t n_win = (int64_t)(kmax * rows_per_k / MODEL_DENSITY_ROWS) + 2;
  double *rho = (double *)calloc((size_t)n_win, sizeof(double));
  if (!rho) return NULL;
  for (int p = 0; p < n_pairs; ++p) {
    if (pair_in_group && !pair_in_group[p]) continue;
    const double *k = k3_tr + tr_off[p];
    for (int i = 0; i < k3_tr_size[p]; ++i) {
      int64_t w = (int64_t)(k[i] * rows_per_k / MODEL_DENSITY_ROWS);
      if (w < 0) w = 0;
      if (w >= n_win) w = n_win - 1;
      rho[w] += 1.0 / MODEL_DENSITY_ROWS;
    }
  }
  *n_win_out = n_win;
  return rho;
}

static double model_unit_cost(double k3, double tau0, int l_max, const double *rho, int64_t n_win, double rows_per_k, double share)
{
  int64_t w = (int64_t)(k3 * rows_per_k / MODEL_DENSITY_ROWS);
  if (w < 0) w = 0;
  if (w >= n_win) w = n_win - 1;
  const double r = rho[w] * share;
  return MODEL_C_INTERP + (MODEL_C0 + MODEL_C1 * model_active(k3, tau0, l_max)) * (1.0 + (r > 0. ? MODEL_RHO0 / r : 0.));
}

int song_host_pair_costs_model(int n_pairs, const int *k3_tr_size, const int64_t *tr_off, const double *k3_tr, const int *k3_pt_size,
                               double tau0, int l_max, double rows_per_k, int n_groups, int64_t *cost)
{
  if (!k3_tr_size || !tr_off || !k3_tr || !k3_pt_size || !cost || n_pairs < 0 || n_groups < 1 || !(rows_per_k > 0.)) return 1;
  int64_t n_win = 0;
  double *rho = model_density(n_pairs, k3_tr_size, tr_off, k3_tr, NULL, rows_per_k, &n_win);
  if (!rho) return 1;
  /* a group of an n_groups-way split by interleaved chunks holds about 1/n_groups of the run's values at every k3 */
  const double share = 1.0 / (double)n_groups;
  for (int p = 0; p < n_pairs; ++p) {
    const double *k = k3_tr + tr_off[p];
    double c = MODEL_C_POINT * (double)k3_pt_size[p];
    for (int i = 0; i < k3_tr_size[p]; ++i) c += model_unit_cost(k[i], tau0, l_max, rho, n_win, rows_per_k, share);
    cost[p] = (int64_t)(c * 1024.0);
    if (cost[p] < 1) cost[p] = 1;
  }
  free(rho);
  return 0;
}

int song_host_band_edges(int n_pairs, const int *k3_tr_size, const int64_t *tr_off, const double *k3_tr, const int *k3_pt_size,
                         const unsigned char *pair_in_group, double tau0, int l_max, double rows_per_k, int n_bands, double *edges)
{
  if (!k3_tr_size || !tr_off || !k3_tr || !k3_pt_size || n_pairs < 0 || n_bands < 1 || (n_bands > 1 && !edges) || !(rows_per_k > 0.)) return 1;
  if (n_bands == 1) return 0;
  if (n_bands > 64) return 1;
  /* cost histogram over the k3 axis (linear bins: the fine grids are linear in k3), then the quantiles */
  enum { NBINS = 1 << 16 };
  double kmin = 0., kmax = 0.; int any = 0;
  for (int p = 0; p < n_pairs; ++p) {
    if ((pair_in_group && !pair_in_group[p]) || k3_tr_size[p] < 1) continue;
    const double a = k3_tr[tr_off[p]], b = k3_tr[tr_off[p] + k3_tr_size[p] - 1];
    if (!any || a < kmin) kmin = a;
    if (!any || b > kmax) kmax = b;
    any = 1;
  }
  if (!any || !(kmax > kmin)) { for (int r = 0; r + 1 < n_bands; ++r) edges[r] = kmax; return 0; }
  int64_t n_win = 0;
  double *rho = model_density(n_pairs, k3_tr_size, tr_off, k3_tr, pair_in_group, rows_per_k, &n_win);
  double *hist = (double *)calloc(NBINS, sizeof(double));
  if (!hist || !rho) { free(hist); free(rho); return 1; }
  const double scale = (double)NBINS / (kmax - kmin);
  double total = 0.;
  for (int p = 0; p < n_pairs; ++p) {
    if (pair_in_group && !pair_in_group[p]) continue;
    const double *k = k3_tr + tr_off[p];
    for (int i = 0; i < k3_tr_size[p]; ++i) {
      int bin = (int)((k[i] - kmin) * scale);
      if (bin < 0) bin = 0;
      if (bin >= NBINS) bin = NBINS - 1;
      const double c = model_unit_cost(k[i], tau0, l_max, rho, n_win, rows_per_k, 1.0);
      hist[bin] += c; total += c;
    }
  }
  free(rho);
  /* Every band should cost the same INCLUDING the coarse points of the pairs it touches, which depend on the edges: three
     rounds of  edges from the targets -> points per band -> targets. */
  double point_cost[64];
  for (int r = 0; r < n_bands; ++r) point_cost[r] = 0.;
  for (int round = 0; round < 3; ++round) {
    double psum = 0.;
    for (int r = 0; r < n_bands; ++r) psum += point_cost[r];
    const double per_band = (total + psum) / (double)n_bands;
    double target = 0., acc = 0.; int r = 0;
    target = per_band - point_cost[0];
    for (int bin = 0; bin < NBINS && r + 1 < n_bands; ++bin) {
      acc += hist[bin];
      while (r + 1 < n_bands && acc >= target) { edges[r++] = kmin + (double)(bin + 1) / scale; if (r < n_bands) target += per_band - point_cost[r]; }
    }
    for (; r + 1 < n_bands; ++r) edges[r] = kmax;
    for (int b = 0; b < n_bands; ++b) point_cost[b] = 0.;
    for (int p = 0; p < n_pairs; ++p) {
      if ((pair_in_group && !pair_in_group[p]) || k3_tr_size[p] < 1) continue;
      const double *k = k3_tr + tr_off[p];
      const int n = k3_tr_size[p];
      int lo_prev = 0;
      for (int b = 0; b < n_bands; ++b) {
        int lo = lo_prev, hi = n;
        if (b + 1 < n_bands) { while (lo < hi) { const int mid = (lo + hi) >> 1; if (k[mid] < edges[b]) lo = mid + 1; else hi = mid; } }
        else lo = n;
        if (lo > lo_prev) point_cost[b] += MODEL_C_POINT * (double)k3_pt_size[p];       /* the band holds values of this pair */
        lo_prev = lo;
      }
    }
  }
  free(hist);
  return 0;
}

int song_host_band_cuts(int n_pairs, const int *k3_tr_size, const int64_t *tr_off, const double *k3_tr,
                        int n_bands, const double *edges, int *cut)
{
  if (!k3_tr_size || !tr_off || !k3_tr || !cut || n_pairs < 0 || n_bands < 1 || (n_bands > 1 && !edges)) return 1;
  for (int p = 0; p < n_pairs; ++p) {
    const double *k = k3_tr + tr_off[p];
    const int n = k3_tr_size[p];
    int *c = cut + (size_t)p * (size_t)(n_bands + 1);
    c[0] = 0; c[n_bands] = n;
    for (int r = 1; r < n_bands; ++r) {
      int lo = c[r - 1], hi = n;                       /* first index with k >= edges[r-1]; edges ascend */
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (k[mid] < edges[r - 1]) lo = mid + 1; else hi = mid; }
      c[r] = lo;
    }
  }
  return 0;
}

int song_host_band_count(int64_t total_units, int xx_size, int n_dev, int dense_per_row)
{
  if (n_dev <= 1 || xx_size < 1 || dense_per_row < 1) return 1;
  for (int bands = 1; bands <= n_dev; ++bands) {
    if (n_dev % bands) continue;
    if ((double)total_units / (double)n_dev * (double)bands / (double)xx_size >= (double)dense_per_row) return bands;
  }
  return n_dev;
}
