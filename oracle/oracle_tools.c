/*
 * oracle_tools.c -- restatements of the CLASS-fork helpers the hot path calls.
 * TEST INFRASTRUCTURE ONLY (see los_oracle.h).  Helpers of the absent CLASS fork, UPSTREAM-RECALLED: pinned by known-answer tests only (tests/test_oracle_kat.py).
 *
 * Every function here is UPSTREAM-RECALLED: the sources (class.git/tools/arrays.c,
 * song_tools.c, slatec_3j_C.c, source/bessel.c) are absent from /root/reference, so the
 * algorithms are restated from the published CLASS / SLATEC / Numerical-Recipes methods and
 * from how SONG calls them (call sites quoted per function).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "los_oracle.h"

static char g_err[1024];
const char *oracle_last_error(void) { return g_err; }
#define ORACLE_ERR(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return ORACLE_FAIL; } while (0)

/* ---------------------------------------------------------------------------------------
 * Cubic-spline second derivatives for y_size columns sharing one abscissa.
 * Call site: transfer2.c:2639-2648 (x=k3_pt, y=sources[tp][k1][k2], y_size=tau_size) and
 * transfer2.c:2822-2831 (y_size=1).  Layout y[index_y*x_size + index_x] (the sources(tau,k)
 * macro, transfer2.h:47-51).  Algorithm: Numerical-Recipes tridiagonal sweep; with
 * _SPLINE_EST_DERIV_ the end slopes come from the parabola through the first/last three points.
 * ------------------------------------------------------------------------------------- */
int oracle_spline_table_columns(const double *x, int n, const double *y, int y_size,
                                double *ddy, int mode)
{
  if (n < 3) ORACLE_ERR("spline needs at least 3 points, got %d", n);
  double *u = (double *)malloc((size_t)(n - 1) * sizeof(double));
  if (!u) ORACLE_ERR("out of memory");

  for (int iy = 0; iy < y_size; ++iy) {
    const double *yy = y + (size_t)iy * n;
    double *dd = ddy + (size_t)iy * n;

    if (mode == ORACLE_SPLINE_NATURAL) {
      dd[0] = 0.0;
      u[0] = 0.0;
    } else {
      double dy_first = ((x[2] - x[0]) * (x[2] - x[0]) * (yy[1] - yy[0])
                         - (x[1] - x[0]) * (x[1] - x[0]) * (yy[2] - yy[0]))
                        / ((x[2] - x[0]) * (x[1] - x[0]) * (x[2] - x[1]));
      dd[0] = -0.5;
      u[0] = (3. / (x[1] - x[0])) * ((yy[1] - yy[0]) / (x[1] - x[0]) - dy_first);
    }

    for (int i = 1; i < n - 1; ++i) {
      double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
      double p = sig * dd[i - 1] + 2.0;
      dd[i] = (sig - 1.0) / p;
      double ui = (yy[i + 1] - yy[i]) / (x[i + 1] - x[i]) - (yy[i] - yy[i - 1]) / (x[i] - x[i - 1]);
      u[i] = (6.0 * ui / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
    }

    double qn, un;
    if (mode == ORACLE_SPLINE_NATURAL) {
      qn = 0.0;
      un = 0.0;
    } else {
      double dy_last = ((x[n - 3] - x[n - 1]) * (x[n - 3] - x[n - 1]) * (yy[n - 2] - yy[n - 1])
                        - (x[n - 2] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (yy[n - 3] - yy[n - 1]))
                       / ((x[n - 3] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (x[n - 3] - x[n - 2]));
      qn = 0.5;
      un = (3. / (x[n - 1] - x[n - 2])) * (dy_last - (yy[n - 1] - yy[n - 2]) / (x[n - 1] - x[n - 2]));
    }
    dd[n - 1] = (un - qn * u[n - 2]) / (qn * dd[n - 2] + 1.0);

    for (int i = n - 2; i >= 0; --i)
      dd[i] = dd[i] * dd[i + 1] + u[i];
  }
  free(u);
  return ORACLE_OK;
}

/* Call site: bessel2.c:417-426, :445-454 (always y_size=1, index_y=0). */
int oracle_spline_table_one_column(const double *x, int n, const double *y, int y_size,
                                   int index_y, double *ddy, int mode)
{
  (void)y_size;
  return oracle_spline_table_columns(x, n, y + (size_t)index_y * n, 1, ddy + (size_t)index_y * n, mode);
}

/* Call site: bessel2.c:981, perturbations2.c:1626,2166.  Endpoints exact. */
void oracle_lin_space(double *x, double x_min, double x_max, int n)
{
  for (int j = 0; j < n - 1; ++j)
    x[j] = x_min + j * (x_max - x_min) / (n - 1);
  x[n - 1] = x_max;
}

/* ---------------------------------------------------------------------------------------
 * (l,m) <-> flat offset, l-major / m-minor, only m <= l allowed.
 * Call sites: transfer2.c:900,1154,1207; perturbations2.c:854,1159; macro lm(l,m) of
 * perturbations2_macros.h:53.
 * ------------------------------------------------------------------------------------- */
int oracle_size_l_indexm(int l_max, const int *m, int m_size)
{
  int count = 0;
  for (int l = 0; l <= l_max; ++l)
    for (int im = 0; im < m_size; ++im)
      if (m[im] <= l) ++count;
  return count;
}

int oracle_multipole2offset_l_indexm(int L, int M, const int *m, int m_size)
{
  int offset = 0;
  for (int l = 0; l <= L; ++l)
    for (int im = 0; im < m_size; ++im) {
      if (m[im] > l) continue;
      if (l == L && m[im] == M) return offset;
      ++offset;
    }
  return -1;
}

int oracle_size_indexl_indexm(const int *l, int l_size, const int *m, int m_size)
{
  int count = 0;
  for (int il = 0; il < l_size; ++il)
    for (int im = 0; im < m_size; ++im)
      if (m[im] <= l[il]) ++count;
  return count;
}

int oracle_multipole2offset_indexl_indexm(int L, int M, const int *l, int l_size,
                                          const int *m, int m_size)
{
  int offset = 0;
  for (int il = 0; il < l_size; ++il)
    for (int im = 0; im < m_size; ++im) {
      if (m[im] > l[il]) continue;
      if (l[il] == L && m[im] == M) return offset;
      ++offset;
    }
  return -1;
}

int oracle_offset2multipole_indexl_indexm(int target, const int *l, int l_size,
                                          const int *m, int m_size, int *index_l, int *index_m)
{
  int offset = 0;
  for (int il = 0; il < l_size; ++il)
    for (int im = 0; im < m_size; ++im) {
      if (m[im] > l[il]) continue;
      if (offset == target) { *index_l = il; *index_m = im; return ORACLE_OK; }
      ++offset;
    }
  *index_l = -1; *index_m = -1;
  ORACLE_ERR("offset %d out of range", target);
}

/* ---------------------------------------------------------------------------------------
 * Wigner 3j symbols ( l1 l2 l3 ; m1 m2 m3 ), m1 = -m2-m3, for every allowed l1.
 * Call sites: bessel2.c:1131 drc3jj(L,l,S,-S) and :1154 drc3jj(L,l,-m,m).
 * Algorithm (SLATEC DRC3JJ, Schulten & Gordon 1975): three-term recursion in l1
 *   l1 A(l1+1) f(l1+1) + B(l1) f(l1) + (l1+1) A(l1) f(l1-1) = 0,
 *   A(l1)^2 = (l1^2-(l2-l3)^2) ((l2+l3+1)^2-l1^2) (l1^2-m1^2),
 *   B(l1)   = -(2 l1+1) [ l2(l2+1) m1 - l3(l3+1) m1 - l1(l1+1)(m3-m2) ],
 * run upward from l1min while it is stable (|C1| decreasing), downward from l1max for the
 * rest, matched on three overlapping values, normalised by sum (2 l1+1) f^2 = 1, with the
 * sign of f(l1max) fixed to (-1)^(l2-l3-m1).
 * ------------------------------------------------------------------------------------- */
static double a_sq(double l1, double l2, double l3, double m1)
{
  return (l1 * l1 - (l2 - l3) * (l2 - l3)) * ((l2 + l3 + 1.) * (l2 + l3 + 1.) - l1 * l1) * (l1 * l1 - m1 * m1);
}

int oracle_drc3jj(double l2, double l3, double m2, double m3,
                  double *l1min_out, double *l1max_out, double *thrcof, int ndim)
{
  const double eps = 0.01, big = 1e150, inv_big = 1e-150;
  double m1 = -m2 - m3;

  if (l2 - fabs(m2) + eps < 0. || l3 - fabs(m3) + eps < 0.)
    ORACLE_ERR("drc3jj: l2-|m2| or l3-|m3| negative");
  if (fmod(l2 + fabs(m2) + eps, 1.) >= eps + eps || fmod(l3 + fabs(m3) + eps, 1.) >= eps + eps)
    ORACLE_ERR("drc3jj: l2+|m2| or l3+|m3| not integer");

  double l1min = fmax(fabs(l2 - l3), fabs(m1));
  double l1max = l2 + l3;
  *l1min_out = l1min;
  *l1max_out = l1max;
  if (fmod(l1max - l1min + eps, 1.) >= eps + eps) ORACLE_ERR("drc3jj: l1max-l1min not integer");
  if (l1min >= l1max + eps) ORACLE_ERR("drc3jj: l1min > l1max");

  int n = (int)(l1max - l1min + 1. + eps);
  if (n > ndim) ORACLE_ERR("drc3jj: ndim=%d too small for %d values", ndim, n);

  if (n == 1) {
    int e = (int)(fabs(l2 + m2 - l3 + m3) + eps);
    thrcof[0] = ((e % 2 == 0) ? 1. : -1.) / sqrt(l1min + l2 + l3 + 1.);
    return ORACLE_OK;
  }

  /* ---- upward recursion ---- */
  int nf;                       /* number of values produced by the upward leg */
  double c1old = 0., oldfac = 0.;
  thrcof[0] = 1.;
  {
    int i = 1;
    for (;;) {
      double l1 = l1min + i;
      double newfac = sqrt(a_sq(l1, l2, l3, m1));
      double c1, denom;
      if (l1 < 1. + eps) {
        /* l1min = 0: recursion divided through by (l1-1) */
        denom = newfac;
        c1 = -(2. * l1 - 1.) * l1 * (m3 - m2) / newfac;
      } else {
        double dv = -l2 * (l2 + 1.) * m1 + l3 * (l3 + 1.) * m1 + l1 * (l1 - 1.) * (m3 - m2);
        denom = (l1 - 1.) * newfac;
        c1 = -(2. * l1 - 1.) * dv / denom;
      }
      if (i == 1) {
        thrcof[1] = c1 * thrcof[0];
      } else {
        double c2 = -l1 * oldfac / denom;
        thrcof[i] = c1 * thrcof[i - 1] + c2 * thrcof[i - 2];
      }
      oldfac = newfac;
      if (fabs(thrcof[i]) > big)
        for (int j = 0; j <= i; ++j) thrcof[j] *= inv_big;
      if (i == n - 1) { nf = n; break; }
      if (i >= 2 && !(c1old > fabs(c1))) { nf = i + 1; break; }
      c1old = fabs(c1);
      ++i;
    }
  }

  /* ---- downward recursion and matching ---- */
  if (nf < n) {
    double f1 = thrcof[nf - 3], f2 = thrcof[nf - 2], f3 = thrcof[nf - 1];
    int lo = nf - 3;            /* lowest index produced by the downward leg */
    double *g = (double *)malloc((size_t)n * sizeof(double));
    if (!g) ORACLE_ERR("out of memory");
    g[n - 1] = 1.;
    for (int i = n - 2; i >= lo; --i) {
      /* relation centred on l1 = l1min+i+1: solve for f(l1-1) */
      double l1 = l1min + i + 1;
      double a_up = sqrt(fmax(a_sq(l1 + 1., l2, l3, m1), 0.));
      double a_mid = sqrt(a_sq(l1, l2, l3, m1));
      double b = -(2. * l1 + 1.) * (l2 * (l2 + 1.) * m1 - l3 * (l3 + 1.) * m1 - l1 * (l1 + 1.) * (m3 - m2));
      double up = (i + 2 <= n - 1) ? g[i + 2] : 0.;
      g[i] = -(l1 * a_up * up + b * g[i + 1]) / ((l1 + 1.) * a_mid);
      if (fabs(g[i]) > big)
        for (int j = i; j < n; ++j) g[j] *= inv_big;
    }
    double ratio = (f1 * g[lo] + f2 * g[lo + 1] + f3 * g[lo + 2]) / (f1 * f1 + f2 * f2 + f3 * f3);
    for (int i = 0; i < lo; ++i) thrcof[i] *= ratio;
    for (int i = lo; i < n; ++i) thrcof[i] = g[i];
    free(g);
  }

  /* ---- normalisation and phase ---- */
  double sum = 0.;
  for (int i = 0; i < n; ++i) sum += (2. * (l1min + i) + 1.) * thrcof[i] * thrcof[i];
  double cnorm = 1. / sqrt(sum);
  int e = (int)(fabs(l2 - l3 - m1) + eps);
  double sign2 = (e % 2 == 0) ? 1. : -1.;
  double sign1 = (thrcof[n - 1] >= 0.) ? 1. : -1.;
  if (sign1 * sign2 < 0.) cnorm = -cnorm;
  for (int i = 0; i < n; ++i) thrcof[i] *= cnorm;
  return ORACLE_OK;
}

/* ---------------------------------------------------------------------------------------
 * Spherical Bessel functions.  The reference calls CLASS's bessel_j (bessel2.c:1706,1724,
 * 1804), an O(1) approximation accurate to ~1e-5 whose source is absent; j_l1 is therefore an
 * INPUT of the parity tests (same table fed to oracle and GPU) and this generator only has to
 * be a good j_l: upward recurrence for l <= x, Miller's downward recurrence above, matched at
 * l = floor(x).  Relative accuracy ~1e-13.
 * ------------------------------------------------------------------------------------- */
static void sph_bessel_all(int lmax, double x, double *out)
{
  if (x == 0.) {
    out[0] = 1.;
    for (int l = 1; l <= lmax; ++l) out[l] = 0.;
    return;
  }
  int lm = (int)floor(x);
  if (lm > lmax) lm = lmax;
  double j0, j1;
  if (fabs(x) < 1e-2) {
    double x2 = x * x;
    j0 = 1. - x2 / 6. * (1. - x2 / 20. * (1. - x2 / 42.));
    j1 = x / 3. * (1. - x2 / 10. * (1. - x2 / 28. * (1. - x2 / 54.)));
  } else {
    j0 = sin(x) / x;
    j1 = (sin(x) / x - cos(x)) / x;
  }
  out[0] = j0;
  if (lmax >= 1 && lm >= 1) out[1] = j1;
  for (int l = 1; l < lm; ++l)
    out[l + 1] = (2. * l + 1.) / x * out[l] - out[l - 1];
  if (lm == lmax) return;

  /* Miller: downward from well above lmax to lm, tracking rescalings */
  int nstart = lmax + 30 + (int)sqrt(100. * (lmax + 1.));
  double tp1 = 0., t = 1e-30;
  int scale = 0;
  int *sc = (int *)malloc((size_t)(lmax + 1) * sizeof(int));
  for (int l = nstart; l > lm; --l) {
    double tm1 = (2. * l + 1.) / x * t - tp1;   /* value at l-1 */
    tp1 = t; t = tm1;
    if (fabs(t) > 1e200) { t *= 1e-200; tp1 *= 1e-200; ++scale; }
    if (l - 1 <= lmax) { out[l - 1] = t; sc[l - 1] = scale; }
  }
  /* now out[lm] holds the unnormalised value; the true one is known */
  double true_lm = (lm == 0) ? j0 : ((lm == 1) ? j1 : 0.);
  if (lm >= 2) {
    /* recompute upward value at lm (it was overwritten) */
    double a = j0, b = j1;
    for (int l = 1; l < lm; ++l) { double c = (2. * l + 1.) / x * b - a; a = b; b = c; }
    true_lm = b;
  }
  double norm = true_lm / out[lm];
  int sm = sc[lm];
  for (int l = lm; l <= lmax; ++l) {
    int ds = sm - sc[l];
    double v = out[l] * norm;
    if (ds == 0) out[l] = v;
    else if (ds == 1) out[l] = v * 1e-200;
    else out[l] = 0.;
  }
  out[lm] = true_lm;
  free(sc);
}

double oracle_sph_bessel_j(int l, double x)
{
  double *tmp = (double *)malloc((size_t)(l + 1) * sizeof(double));
  sph_bessel_all(l, x, tmp);
  double v = tmp[l];
  free(tmp);
  return v;
}

/* exported for oracle_bessel2.c */
void oracle_sph_bessel_all(int lmax, double x, double *out) { sph_bessel_all(lmax, x, out); }
