/*
 * bessel2_host.c -- host half of the bessel2 drop-in (see bessel2_host.h).
 * bessel_j / drc3jj resolve to CLASS's own functions in a SONG tree and to class_compat.c here.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bessel2_host.h"
#include "class_compat.h"

/* j_l(x) provider (bessel2_host.h) */
static song_host_bessel_fn g_bessel_fn = NULL;
static void *g_bessel_user = NULL;
void song_host_bessel2_set_provider(song_host_bessel_fn fn, void *user) { g_bessel_fn = fn; g_bessel_user = user; }
static double jl_point(int l, double x) { return g_bessel_fn ? g_bessel_fn(l, x, g_bessel_user) : compat_bessel_j(l, x); }
static song_host_3j_fn g_3j_fn = NULL;
static void *g_3j_user = NULL;
void song_host_bessel2_set_3j_provider(song_host_3j_fn fn, void *user) { g_3j_fn = fn; g_3j_user = user; }
static int threej(double l2, double l3, double m2, double m3, double *lo, double *hi, double *f, int ndim)
{
  return g_3j_fn ? g_3j_fn(l2, l3, m2, m3, lo, hi, f, ndim, g_3j_user) : compat_drc3jj(l2, l3, m2, m3, lo, hi, f, ndim);
}

static int b2_fail(char *errmsg, int errlen, const char *text)
{
  if (errmsg && errlen > 0) snprintf(errmsg, (size_t)errlen, "%s", text);
  return 1;
}

void song_host_bessel2_free(song_bessel2_host *t)
{
  if (!t) return;
  if (t->j_l1) for (int i = 0; i < t->l1_size; ++i) free(t->j_l1[i]);
  free(t->j_l1); free(t->l1); free(t->index_l1); free(t->index_xmin_l1); free(t->x_size_l1); free(t->x_min_l1);
  free(t->first_3j); free(t->second_3j); free(t->l1_min); free(t->l1_count); free(t->index_l1_of_l1min);
  free(t);
}

int song_host_bessel2_prepare(const int *l, int l_size, const int *m, int m_size, int L_max,
                              const double *xx, int xx_size, double xx_max,
                              int J_size, const int *J_types,
                              double j_l1_cut, double tol_x_min, double smallest_var,
                              song_bessel2_host **out, char *errmsg, int errlen)
{
  song_bessel2_host *t = (song_bessel2_host *)calloc(1, sizeof *t);
  if (!t) return b2_fail(errmsg, errlen, "out of memory");
  int m_max = m[m_size - 1];
  int L_size = L_max + 1;

  /* ---- l1 list: every l1 within +-(L_max + m_max) of a member of l (bessel2.c:873-922) ---- */
  int reach = L_max + m_max;
  int l1_top = l[l_size - 1] + reach;
  char *keep = (char *)calloc((size_t)l1_top + 1, 1);
  for (int il = 0; il < l_size; ++il) {
    int lo = l[il] - reach; if (lo < 0) lo = 0;
    for (int c = lo; c <= l[il] + reach; ++c) keep[c] = 1;
  }
  for (int c = 0; c <= l1_top; ++c) t->l1_size += keep[c];
  t->l1 = (int *)malloc((size_t)t->l1_size * sizeof(int));
  t->index_l1_len = l1_top + 1;
  t->index_l1 = (int *)malloc((size_t)t->index_l1_len * sizeof(int));
  for (int c = 0, i = 0; c <= l1_top; ++c) {
    t->index_l1[c] = keep[c] ? i : -1;                               /* bessel2.c:930-939 */
    if (keep[c]) t->l1[i++] = c;
  }
  free(keep);
  int n1 = t->l1_size;
  t->j_l1 = (double **)calloc((size_t)n1, sizeof(double *));
  t->index_xmin_l1 = (int *)malloc((size_t)n1 * sizeof(int));
  t->x_size_l1 = (int *)malloc((size_t)n1 * sizeof(int));
  t->x_min_l1 = (double *)malloc((size_t)n1 * sizeof(double));
  char *all_negligible = (char *)calloc((size_t)n1, 1);

  /* ---- lower cut of every j_l1 by bisection on x (bessel2.c:1686-1762) ---- */
  int bad = 0;
  #pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < n1; ++i) {
    int l1 = t->l1[i];
    double x_min = 0;
    t->index_xmin_l1[i] = 0;
    if (l1 != 0) {
      double lo = 0, hi = (double)l1 + 0.5;
      if (jl_point(l1, hi) < j_l1_cut) { bad = 1; continue; }  /* :1713 */
      while ((hi - lo) / lo > tol_x_min) {
        if ((hi - lo) < smallest_var) { hi = lo; break; }
        double mid = 0.5 * (hi + lo);
        if (jl_point(l1, mid) >= j_l1_cut) hi = mid; else lo = mid;
      }
      x_min = lo;
      if (x_min < xx_max) {
        int ix = 0;
        while (xx[ix] < x_min) ix++;
        if (ix != 0) --ix;                                             /* :1748 */
        t->index_xmin_l1[i] = ix;
      }
    }
    if (x_min >= xx_max) {                                             /* :1756-1759 */
      t->index_xmin_l1[i] = xx_size - 1; t->x_size_l1[i] = 1; t->x_min_l1[i] = xx_max; all_negligible[i] = 1;
    } else {
      t->x_size_l1[i] = xx_size - t->index_xmin_l1[i];
      t->x_min_l1[i] = xx[t->index_xmin_l1[i]];
    }
    t->j_l1[i] = (double *)calloc((size_t)t->x_size_l1[i], sizeof(double));
  }
  if (bad) { free(all_negligible); song_host_bessel2_free(t); return b2_fail(errmsg, errlen, "in bisection, wrong initial guess for x_min_up."); }
  for (int i = 0; i < n1; ++i) t->count_allocated_Js += t->x_size_l1[i];

  /* ---- fill j_l1(xx) (bessel2.c:1802-1810) ---- */
  const char *pp = getenv("SONG_LOS_JL1_PER_POINT");
  if (g_bessel_fn != NULL || (pp != NULL && pp[0] == '1')) {
    /* the reference's call pattern: the provider once per (l1, x) */
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n1; ++i) {
      if (all_negligible[i]) continue;
      for (int r = 0; r < t->x_size_l1[i]; ++r) t->j_l1[i][r] = jl_point(t->l1[i], xx[t->index_xmin_l1[i] + r]);
    }
  } else {
  /* default provider: one recurrence per x serves every l1 */
  int lmax = t->l1[n1 - 1];
  #pragma omp parallel
  {
    double *col = (double *)malloc((size_t)(lmax + 1) * sizeof(double));
    #pragma omp for schedule(dynamic, 32)
    for (int ix = 0; ix < xx_size; ++ix) {
      compat_sph_bessel_table(lmax, xx[ix], col);
      for (int i = 0; i < n1; ++i) {
        if (all_negligible[i]) continue;
        int r = ix - t->index_xmin_l1[i];
        if (r >= 0) t->j_l1[i][r] = col[t->l1[i]];
      }
    }
    free(col);
  }
  }
  free(all_negligible);

  /* ---- the two 3j symbols of every (J,L,l,m) table (bessel2.c:1074-1179) ---- */
  t->ndim = 2 * L_max + 1;
  t->n_jidx = J_size * L_size * l_size * m_size;
  size_t nn = (size_t)t->n_jidx;
  t->first_3j = (double *)calloc(nn * t->ndim, sizeof(double));
  t->second_3j = (double *)calloc(nn * t->ndim, sizeof(double));
  t->l1_min = (int *)calloc(nn, sizeof(int));
  t->l1_count = (int *)calloc(nn, sizeof(int));
  t->index_l1_of_l1min = (int *)calloc(nn, sizeof(int));
  for (int J = 0; J < J_size; ++J) {
    int pf = J_types[J];
    int S = (pf == SONG_LOS_J_TT) ? 0 : 2;
    #pragma omp parallel for collapse(2) schedule(dynamic, 8)
    for (int L = 0; L < L_size; ++L)
      for (int il = 0; il < l_size; ++il)
        for (int im = 0; im < m_size; ++im) {
          int lv = l[il], mv = m[im];
          size_t q = (((size_t)J * L_size + L) * l_size + il) * m_size + im;
          int lim = (L < lv) ? L : lv;
          if (mv > lim) continue;                                      /* not allocated, :288 */
          if (S > lv || S > L || (pf == SONG_LOS_J_EB && mv == 0)) continue;   /* :1090-1100 -> l1_count 0 */
          double lo, hi;
          double *f1 = t->first_3j + q * t->ndim, *f2 = t->second_3j + q * t->ndim;
          if (threej(L, lv, S, -S, &lo, &hi, f1, t->ndim)) { bad = 1; continue; }      /* :1131 */
          int l1min = (int)(lo + 1e-6), l1max = (int)(hi + 1e-6);
          int cnt = l1max - l1min + 1;
          if ((S != 0) || (mv != 0)) {
            if (threej(L, lv, -mv, mv, &lo, &hi, f2, t->ndim)) { bad = 1; continue; }  /* :1154 */
          } else {
            memcpy(f2, f1, (size_t)cnt * sizeof(double));                                    /* :1165 */
          }
          if (l1min >= t->index_l1_len || t->index_l1[l1min] < 0) { bad = 1; continue; }     /* :1177 */
          t->l1_min[q] = l1min; t->l1_count[q] = cnt; t->index_l1_of_l1min[q] = t->index_l1[l1min];
        }
  }
  if (bad) { song_host_bessel2_free(t); return b2_fail(errmsg, errlen, "error in ell_1 list or in the 3j recursion"); }
  *out = t;
  return 0;
}

int song_host_bessel2_upload(song_los_ctx *ctx, const song_bessel2_host *t, const int *J_types,
                             double J_Llm_cut, double xx_max, char *errmsg, int errlen)
{
  int rc = song_los_build_projection(ctx, t->l1_size, t->l1, (const double *const *)t->j_l1, t->index_xmin_l1,
                                     t->x_size_l1, t->n_jidx, t->first_3j, t->second_3j, t->l1_min,
                                     t->l1_count, t->index_l1_of_l1min, J_types, J_Llm_cut, xx_max);
  if (rc) return b2_fail(errmsg, errlen, song_los_last_error(ctx));
  return 0;
}
