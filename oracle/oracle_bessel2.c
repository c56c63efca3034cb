/*
 * oracle_bessel2.c -- restatement of the projection-function tables of SONG's bessel2.c.
 * TEST INFRASTRUCTURE ONLY (see los_oracle.h).  PINNED: bit-identical to the reference's own source/bessel2.c run here on the same j_l(x) / 3j numbers (tests/test_reference_pin.py).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "los_oracle.h"

void oracle_sph_bessel_all(int lmax, double x, double *out);

#define ALT_SIGN(n) (((n) % 2 == 0) ? 1 : -1)   /* ALTERNATING_SIGN, test/print_matrix_inverse.c:5 */
#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))

/* bessel2.c:979: xx_size = xx_max/xx_step + 1, truncated to int */
int oracle_bessel2_xx_size(double xx_max, double xx_step)
{
  return (int)(xx_max / xx_step + 1);
}

/* bessel2.c:832-943.  Keep every l1 with l-Lm <= l1 <= l+Lm for some l in the list
   ("It is important to omit abs()", :889), Lm = L_max + m_max_song (:873). */
int oracle_bessel2_l1_list(const int *l, int l_size, int Lm, int *l1)
{
  int l1_max = l[l_size - 1] + Lm;
  int count = 0;
  for (int c = 0; c <= l1_max; ++c) {
    int keep = 0;
    for (int il = 0; il < l_size; ++il)
      if (c >= l[il] - Lm && c <= l[il] + Lm) keep = 1;
    if (keep) l1[count++] = c;
  }
  return count;
}

void oracle_jl1_free(oracle_jl1 *t)
{
  if (!t) return;
  if (t->j_l1) for (int i = 0; i < t->l1_size; ++i) free(t->j_l1[i]);
  free(t->j_l1); free(t->l1); free(t->index_l1);
  free(t->index_xmin_l1); free(t->x_size_l1); free(t->x_min_l1);
  memset(t, 0, sizeof *t);
}

/* bessel2.c:1669-1815 applied to every l1 (loop of bessel2.c:205-214).
   bessel_j -> oracle_sph_bessel_j (see oracle_tools.c for why this is not CLASS's). */
int oracle_bessel2_j_for_all_l1(const int *l, int l_size, int L_max, int m_max,
                                const double *xx, int xx_size, double xx_max,
                                double j_l1_cut, double bessel_tol_x_min,
                                double smallest_allowed_variation, oracle_jl1 *out)
{
  memset(out, 0, sizeof *out);
  int Lm = L_max + m_max;
  int cap = l[l_size - 1] + Lm + 1;
  out->l1 = (int *)malloc((size_t)cap * sizeof(int));
  out->l1_size = oracle_bessel2_l1_list(l, l_size, Lm, out->l1);
  int n1 = out->l1_size;

  /* inverse map, bessel2.c:930-939 */
  out->index_l1_len = out->l1[n1 - 1] + 1;
  out->index_l1 = (int *)malloc((size_t)out->index_l1_len * sizeof(int));
  for (int c = 0; c < out->index_l1_len; ++c) out->index_l1[c] = -1;
  for (int i = 0; i < n1; ++i) out->index_l1[out->l1[i]] = i;

  out->j_l1 = (double **)calloc((size_t)n1, sizeof(double *));
  out->index_xmin_l1 = (int *)malloc((size_t)n1 * sizeof(int));
  out->x_size_l1 = (int *)malloc((size_t)n1 * sizeof(int));
  out->x_min_l1 = (double *)malloc((size_t)n1 * sizeof(double));
  char *negligible = (char *)calloc((size_t)n1, 1);

  /* per-l1 lower cut, bessel2.c:1686-1762 */
  for (int i = 0; i < n1; ++i) {
    int l1 = out->l1[i];
    double x_min;
    if (l1 == 0) {
      x_min = 0;
      out->index_xmin_l1[i] = 0;
    } else {
      double x_min_down = 0;
      double x_min_up = (double)l1 + 0.5;
      double j = oracle_sph_bessel_j(l1, x_min_up);
      if (j < j_l1_cut) return ORACLE_FAIL;          /* "wrong initial guess for x_min_up" :1713 */
      x_min = 0;
      while ((x_min_up - x_min_down) / x_min_down > bessel_tol_x_min) {
        if ((x_min_up - x_min_down) < smallest_allowed_variation) {
          x_min = x_min_up = x_min_down;
          break;
        }
        j = oracle_sph_bessel_j(l1, 0.5 * (x_min_up + x_min_down));
        if (j >= j_l1_cut)
          x_min_up = 0.5 * (x_min_up + x_min_down);
        else
          x_min_down = 0.5 * (x_min_up + x_min_down);
      }
      x_min = x_min_down;
      if (x_min < xx_max) {
        int ix = 0;
        while (xx[ix] < x_min) ix++;
        if (ix != 0) --ix;                            /* store from the index preceding x_min :1748 */
        out->index_xmin_l1[i] = ix;
      }
    }
    if (x_min >= xx_max) {
      out->index_xmin_l1[i] = xx_size - 1;
      out->x_size_l1[i] = 1;
      out->x_min_l1[i] = xx_max;
      negligible[i] = 1;
    } else {
      out->x_size_l1[i] = xx_size - out->index_xmin_l1[i];
      out->x_min_l1[i] = xx[out->index_xmin_l1[i]];
    }
    out->j_l1[i] = (double *)calloc((size_t)out->x_size_l1[i], sizeof(double));
  }

  /* fill: j_l1[i][ix - index_xmin_l1[i]] = j_{l1}(xx[ix]), bessel2.c:1802-1810; done x-major so
     that one recurrence per x serves every l1 */
  int lmax = out->l1[n1 - 1];
  #pragma omp parallel
  {
    double *col = (double *)malloc((size_t)(lmax + 1) * sizeof(double));
    #pragma omp for schedule(dynamic, 64)
    for (int ix = 0; ix < xx_size; ++ix) {
      oracle_sph_bessel_all(lmax, xx[ix], col);
      for (int i = 0; i < n1; ++i) {
        if (negligible[i]) continue;                   /* all-negligible row stays 0 :1788-1792 */
        int r = ix - out->index_xmin_l1[i];
        if (r >= 0) out->j_l1[i][r] = col[out->l1[i]];
      }
    }
    free(col);
  }
  free(negligible);
  return ORACLE_OK;
}

/* bessel2.c:1460-1588: J_Llm at grid index index_x from the j_l1 table and the two 3j arrays */
static double J_Llm_at(int pf, int L, int l, int m, int index_x, const oracle_jl1 *jl,
                       const double *first_3j, const double *second_3j, int l1_size,
                       int index_l1_min)
{
  double J = 0;
  for (int i = 0; i < l1_size; ++i) {
    if (first_3j[i] == 0) continue;                               /* :1489 */
    int row = index_l1_min + i;
    int r = index_x - jl->index_xmin_l1[row];
    int l1 = jl->l1[row];
    if (r < 0) continue;                                          /* :1499 */
    int is_even = ((l - l1 - L) % 2 == 0);
    double pre;
    if (pf == ORACLE_J_TT || pf == ORACLE_J_EE) {
      if (!is_even) continue;
      pre = ALT_SIGN((l - l1 - L) / 2);                           /* :1521 */
    } else {
      if (is_even) continue;
      pre = ALT_SIGN((l - l1 - L - 1) / 2);                       /* :1527 */
    }
    double j_l1_x = jl->j_l1[row][r];
    J += pre * (2 * l1 + 1) * first_3j[i] * second_3j[i] * j_l1_x; /* :1543-1546 */
  }
  J *= ALT_SIGN(m) * (2 * l + 1);                                 /* :1555 */
  return J;
}

/* bessel2.c:1041-1437 */
int oracle_bessel2_J_for_Llm(int pf, int L, int l, int m, int L_max,
                             const oracle_jl1 *jl, const double *xx, int xx_size, double xx_max,
                             double J_Llm_cut,
                             double **row, int *index_xmin_out, int *x_size_out, double *x_min_out)
{
  int index_xmin_J = 0;
  double x_min_J = 0;
  int S = (pf == ORACLE_J_TT) ? 0 : 2;                            /* :1074-1085 */
  int spin_constraint = (abs(S) > l) || (abs(S) > L);             /* :1090 */
  int EB_constraint = (pf == ORACLE_J_EB) && (m == 0);            /* :1093 */
  int ndim = 2 * L_max + 1;
  double *first_3j = NULL, *second_3j = NULL;
  int l1_min = 0, l1_size = 0, index_l1_min = 0;

  if (spin_constraint || EB_constraint) {
    index_xmin_J = xx_size - 1;                                   /* :1098 */
    goto store;
  }

  first_3j = (double *)calloc((size_t)ndim, sizeof(double));
  second_3j = (double *)calloc((size_t)ndim, sizeof(double));
  {
    double lo, hi;
    if (oracle_drc3jj(L, l, S, -S, &lo, &hi, first_3j, ndim) != ORACLE_OK) goto fail;   /* :1131 */
    l1_min = (int)(lo + 1e-6);
    int l1_max = (int)(hi + 1e-6);
    l1_size = l1_max - l1_min + 1;
    if ((S != 0) || (m != 0)) {
      if (oracle_drc3jj(L, l, -m, m, &lo, &hi, second_3j, ndim) != ORACLE_OK) goto fail; /* :1154 */
    } else {
      for (int i = 0; i < l1_size; ++i) second_3j[i] = first_3j[i];                     /* :1165 */
    }
    if (l1_min >= jl->index_l1_len || jl->index_l1[l1_min] < 0) goto fail;              /* :1177 */
    index_l1_min = jl->index_l1[l1_min];
  }

  if (L == l) {                                                   /* :1202-1206 */
    index_xmin_J = 0;
    x_min_J = xx[0];
    goto store;
  }

  {
    int down = 0, up;
    double x_up = l1_min + 0.5;                                   /* :1225 */
    if (x_up >= xx_max) {
      up = xx_size - 1;
      x_up = xx_max;
    } else {
      up = 0;
      while (xx[up] < x_up) up++;                                 /* :1243 */
      x_up = xx[up];
    }
    double J = J_Llm_at(pf, L, l, m, up, jl, first_3j, second_3j, l1_size, index_l1_min);
    if (fabs(J) < J_Llm_cut) {                                    /* :1276 */
      index_xmin_J = xx_size - 1;
      if (x_up < xx_max) goto fail;                               /* :1287 */
      goto store;
    }
    double x_down = xx[0];
    while (down != up - 1) {                                      /* :1303 */
      int mid = (up + down) / 2;
      J = J_Llm_at(pf, L, l, m, mid, jl, first_3j, second_3j, l1_size, index_l1_min);
      if (fabs(J) >= J_Llm_cut) { up = mid; }
      else { down = mid; x_down = xx[mid]; }
    }
    index_xmin_J = down;                                          /* :1341 */
    x_min_J = x_down;
  }

store:
  {
    int x_size = xx_size - index_xmin_J;                          /* :1357 */
    double *r = (double *)calloc((size_t)x_size, sizeof(double)); /* :1362 */
    if (index_xmin_J == xx_size - 1) {                            /* :1378-1388 */
      x_min_J = xx_max;
      r[0] = 0.;
    } else {
      for (int ix = index_xmin_J; ix < xx_size; ++ix)             /* :1400-1423 */
        r[ix - index_xmin_J] = J_Llm_at(pf, L, l, m, ix, jl, first_3j, second_3j, l1_size, index_l1_min);
    }
    *row = r;
    *index_xmin_out = index_xmin_J;
    *x_size_out = x_size;
    *x_min_out = x_min_J;
  }
  free(first_3j); free(second_3j);
  return ORACLE_OK;

fail:
  free(first_3j); free(second_3j);
  return ORACLE_FAIL;
}

/* bessel2.c:499-605: 1/2 sum_k delta_k k^2 f g j_l1(k r), linear (:575-579) or cubic (:581-588) interpolation in the
   j_l1 table of one l1 (row j, second derivatives ddj, first abscissa x_min_l1, step xx_step).  g may be NULL (:547-556).
   Statement order of the reference: zero f or g and x below the table skip the term (:541-563). */
int oracle_bessel2_convolution(const double *kk, const double *delta_kk, int k_size, const double *f, const double *g,
                               const double *j, const double *ddj, double x_min_l1, double xx_step, int cubic, double r,
                               double *integral)
{
  *integral = 0;
  for (int index_k = 0; index_k < k_size; ++index_k) {
    double f_in_k = f[index_k];
    if (f_in_k == 0.) continue;
    double g_in_k = 1.;
    if (g != NULL) {
      g_in_k = g[index_k];
      if (g_in_k == 0.) continue;
    }
    double k = kk[index_k];
    double jv;
    double x = k * r;
    if (x < x_min_l1) continue;
    int index_x = (int)((x - x_min_l1) / xx_step);
    double a = (x_min_l1 + xx_step * (index_x + 1) - x) / xx_step;
    if (!cubic)
      jv = a * j[index_x] + (1. - a) * j[index_x + 1];
    else
      jv = a * j[index_x] + (1. - a) * (j[index_x + 1] - a * ((a + 1.) * ddj[index_x] + (2. - a) * ddj[index_x + 1])
                                                            * xx_step * xx_step / 6.0);
    double integrand = k * k * jv * f_in_k * g_in_k;
    *integral += integrand * delta_kk[index_k];
  }
  (*integral) *= 0.5;
  return 0;
}
