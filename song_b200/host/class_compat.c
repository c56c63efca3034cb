/*
 * class_compat.c -- stand-ins for the CLASS-fork functions the bessel2 drop-in calls.
 *
 * In a real SONG tree these symbols come from class.git (source/bessel.c: bessel_j;
 * tools/slatec_3j_C.c: drc3jj; tools/arrays.c: lin_space) and this file is NOT linked.  The
 * reference checkout has an empty class.git, so the standalone build (tests, bench.py) needs
 * providers with the same calling semantics.  They are inputs generators, not part of the hot
 * path: parity of the J tables is defined for the same j_l1 table and the same 3j symbols.
 *
 *   compat_sph_bessel_table  j_l(x) for l = 0..lmax at one x, Miller downward recurrence started
 *                            above max(lmax, x), normalised with j_0 or j_1 (whichever is larger)
 *   compat_bessel_j          single (l, x) value            (CLASS: bessel_j(pbs, l, x, &j))
 *   compat_drc3jj            Wigner 3j for all l1           (SLATEC: drc3jj(l2,l3,m2,m3,...))
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "class_compat.h"

void compat_lin_space(double *x, double x_min, double x_max, int n)
{
  for (int j = 0; j < n - 1; ++j) x[j] = x_min + j * (x_max - x_min) / (n - 1);
  x[n - 1] = x_max;
}

void compat_sph_bessel_table(int lmax, double x, double *out)
{
  if (x == 0.) {
    out[0] = 1.;
    for (int l = 1; l <= lmax; ++l) out[l] = 0.;
    return;
  }
  double j0, j1;
  if (fabs(x) < 1e-2) {
    double x2 = x * x;
    j0 = 1. - x2 / 6. * (1. - x2 / 20. * (1. - x2 / 42.));
    j1 = x / 3. * (1. - x2 / 10. * (1. - x2 / 28. * (1. - x2 / 54.)));
  } else {
    j0 = sin(x) / x;
    j1 = (sin(x) / x - cos(x)) / x;
  }
  double top = (x > lmax) ? x : (double)lmax;
  int nstart = (int)top + 40 + (int)sqrt(120. * (top + 1.));
  /* pass 1: scale exponent at l = 1 and l = 0 relative to the start (values are rescaled on
     the fly, the number of rescalings before each stored l is remembered) */
  int *resc = (int *)malloc((size_t)(lmax + 2) * sizeof(int));
  double up = 0., cur = 1e-200;
  int count = 0;
  for (int l = nstart; l >= 1; --l) {
    double down = (2. * l + 1.) / x * cur - up;     /* j_{l-1} from j_l, j_{l+1} */
    up = cur; cur = down;
    if (fabs(cur) > 1e150) { cur *= 1e-150; up *= 1e-150; ++count; }
    if (l - 1 <= lmax) { out[l - 1] = cur; resc[l - 1] = count; }
  }
  /* unnormalised u_l = out[l] * 1e150^resc[l]; normalise on the larger of j0, j1 */
  int ref = (fabs(j0) >= fabs(j1) || lmax == 0) ? 0 : 1;
  double truth = ref ? j1 : j0;
  double norm = truth / out[ref];
  int cref = resc[ref];
  for (int l = 0; l <= lmax; ++l) {
    int d = cref - resc[l];                        /* >= 0: rescalings between l and ref */
    double v = out[l] * norm;
    while (d > 0 && v != 0.) { v *= 1e-150; --d; }
    out[l] = v;
  }
  out[ref] = truth;
  free(resc);
}

double compat_bessel_j(int l, double x)
{
  double *t = (double *)malloc((size_t)(l + 1) * sizeof(double));
  compat_sph_bessel_table(l, x, t);
  double v = t[l];
  free(t);
  return v;
}

/* ---------------------------------------------------------------------------------------------
 * 3j symbols by the three-term recursion in l1 (Schulten & Gordon 1975; the algorithm behind
 * SLATEC DRC3JJ), organised as: long-double upward sweep from l1min and downward sweep from
 * l1max, joined at the largest of the three central values, normalised with
 * sum (2 l1+1) f^2 = 1 and sign(f(l1max)) = (-1)^(l2-l3-m1).
 * ------------------------------------------------------------------------------------------- */
static long double A2(long double l1, long double l2, long double l3, long double m1)
{
  return (l1 * l1 - (l2 - l3) * (l2 - l3)) * ((l2 + l3 + 1) * (l2 + l3 + 1) - l1 * l1) * (l1 * l1 - m1 * m1);
}
static long double Bc(long double l1, long double l2, long double l3, long double m1, long double m2, long double m3)
{
  return -(2 * l1 + 1) * (l2 * (l2 + 1) * m1 - l3 * (l3 + 1) * m1 - l1 * (l1 + 1) * (m3 - m2));
}

int compat_drc3jj(double l2, double l3, double m2, double m3, double *l1min_out, double *l1max_out,
                  double *thrcof, int ndim)
{
  long double m1 = -(long double)m2 - m3;
  double l1min = fmax(fabs(l2 - l3), fabs((double)m1));
  double l1max = l2 + l3;
  *l1min_out = l1min; *l1max_out = l1max;
  if (l2 < fabs(m2) - 1e-9 || l3 < fabs(m3) - 1e-9 || l1max < l1min - 1e-9) return 1;
  int n = (int)floor(l1max - l1min + 1.5);
  if (n > ndim) return 2;
  if (n == 1) {
    int e = (int)floor(fabs(l2 + m2 - l3 + m3) + 0.5);
    thrcof[0] = ((e % 2) ? -1. : 1.) / sqrt(l1min + l2 + l3 + 1.);
    return 0;
  }
  long double *f = (long double *)malloc((size_t)n * sizeof(long double));
  long double *g = (long double *)malloc((size_t)n * sizeof(long double));
  int mid = n / 2;
  /* upward: f[i+1] from f[i], f[i-1] using the relation centred on l1 = l1min + i */
  f[0] = 1.L;
  for (int i = 0; i < n - 1 && i <= mid + 1; ++i) {
    long double l1 = (long double)l1min + i;
    long double aup = sqrtl(fmaxl(A2(l1 + 1, l2, l3, m1), 0.L));
    long double prev = (i > 0) ? f[i - 1] : 0.L;
    if (l1 < 0.5L) {
      /* l1 = 0 (l2 = l3, m1 = 0): relation divided by l1 (l1+1), see SLATEC's special case */
      long double a1 = sqrtl(((long double)l2 + l3 + 1) * ((long double)l2 + l3 + 1) - 1.L);
      f[1] = -((long double)m3 - m2) / a1 * f[0];
    } else {
      long double alo = sqrtl(fmaxl(A2(l1, l2, l3, m1), 0.L));
      f[i + 1] = -(Bc(l1, l2, l3, m1, m2, m3) * f[i] + (l1 + 1) * alo * prev) / (l1 * aup);
    }
    if (fabsl(f[i + 1]) > 1e300L) for (int j = 0; j <= i + 1; ++j) f[j] *= 1e-300L;
  }
  int nf = (n - 1 < mid + 2) ? n : mid + 3;       /* entries of f that are valid */
  /* downward: g[i-1] from g[i], g[i+1] */
  g[n - 1] = 1.L;
  int glow = (mid - 1 > 0) ? mid - 1 : 0;
  for (int i = n - 1; i > glow; --i) {
    long double l1 = (long double)l1min + i;
    long double aup = sqrtl(fmaxl(A2(l1 + 1, l2, l3, m1), 0.L));
    long double alo = sqrtl(fmaxl(A2(l1, l2, l3, m1), 0.L));
    long double nxt = (i + 1 <= n - 1) ? g[i + 1] : 0.L;
    if (alo == 0.L) { g[i - 1] = 0.L; continue; }
    g[i - 1] = -(l1 * aup * nxt + Bc(l1, l2, l3, m1, m2, m3) * g[i]) / ((l1 + 1) * alo);
    if (fabsl(g[i - 1]) > 1e300L) for (int j = i - 1; j < n; ++j) g[j] *= 1e-300L;
  }
  /* join on the largest central value */
  int jn = mid;
  for (int c = mid - 1; c <= mid + 1; ++c)
    if (c >= glow && c < nf && c < n && fabsl(f[c]) > fabsl(f[jn])) jn = c;
  long double ratio = (f[jn] != 0.L) ? g[jn] / f[jn] : 0.L;
  long double sum = 0.L;
  for (int i = 0; i < n; ++i) {
    long double v = (i < jn) ? f[i] * ratio : g[i];
    f[i] = v;
    sum += (2 * ((long double)l1min + i) + 1) * v * v;
  }
  long double cn = 1.L / sqrtl(sum);
  int e = (int)floor(fabs(l2 - l3 - (double)m1) + 0.5);
  long double want = (e % 2) ? -1.L : 1.L;
  if ((f[n - 1] < 0 ? -1.L : 1.L) * want < 0) cn = -cn;
  for (int i = 0; i < n; ++i) thrcof[i] = (double)(f[i] * cn);
  free(f); free(g);
  return 0;
}

/* Tridiagonal system of the interpolating cubic spline (continuity of the first derivative at the inner knots), solved by
   forward elimination / back substitution; the workspace holds the eliminated right-hand sides. */
int compat_spline_second_derivatives(const double *x, int n, const double *y, double *ddy, int est_deriv)
{
  if (n < 3) return 1;
  double *rhs = (double *)malloc((size_t)n * sizeof(double));
  if (!rhs) return 1;
  if (est_deriv) {
    const double d1 = x[1] - x[0], d2 = x[2] - x[0];
    if (d1 == 0. || d2 == 0. || d1 == d2) { free(rhs); return 1; }
    /* denominator as upstream CLASS writes it (tools/arrays.c): (x2-x0)(x1-x0)(x2-x1), not (d2-d1) -- one ulp apart */
    const double slope0 = (d2 * d2 * (y[1] - y[0]) - d1 * d1 * (y[2] - y[0])) / (d2 * d1 * (x[2] - x[1]));
    ddy[0] = -0.5;
    rhs[0] = (3. / d1) * ((y[1] - y[0]) / d1 - slope0);
  } else { ddy[0] = 0.; rhs[0] = 0.; }
  for (int i = 1; i < n - 1; ++i) {
    const double hl = x[i] - x[i - 1], hr = x[i + 1] - x[i], span = x[i + 1] - x[i - 1];
    if (hl == 0. || hr == 0.) { free(rhs); return 1; }
    const double w = hl / span, piv = w * ddy[i - 1] + 2.0;
    ddy[i] = (w - 1.0) / piv;
    rhs[i] = (6.0 * ((y[i + 1] - y[i]) / hr - (y[i] - y[i - 1]) / hl) / span - w * rhs[i - 1]) / piv;
  }
  double qn = 0., un = 0.;
  if (est_deriv) {
    const double e1 = x[n - 2] - x[n - 1], e2 = x[n - 3] - x[n - 1];
    const double slopeN = (e2 * e2 * (y[n - 2] - y[n - 1]) - e1 * e1 * (y[n - 3] - y[n - 1])) / (e2 * e1 * (x[n - 3] - x[n - 2]));
    qn = 0.5;
    un = (3. / (x[n - 1] - x[n - 2])) * (slopeN - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
  }
  ddy[n - 1] = (un - qn * rhs[n - 2]) / (qn * ddy[n - 2] + 1.0);
  for (int i = n - 2; i >= 0; --i) ddy[i] = ddy[i] * ddy[i + 1] + rhs[i];
  free(rhs);
  return 0;
}
