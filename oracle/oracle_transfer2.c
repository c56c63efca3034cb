/*
 * oracle_transfer2.c -- restatement of the compute phase of SONG's source/transfer2.c.
 * TEST INFRASTRUCTURE ONLY (see los_oracle.h).  PINNED: bit-identical to the reference's own source/transfer2.c run here (tests/test_reference_pin.py).
 *
 * Statement order inside each expression follows the reference (C left-to-right); build with
 * -ffp-contract=off.  Struct-pointer plumbing is replaced by the flat oracle_setup record.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "los_oracle.h"

#define PI_ 3.1415926535897932384626433832795e0   /* _PI_ of CLASS common.h */
#define DMIN(a, b) (((a) < (b)) ? (a) : (b))
#define DMAX(a, b) (((a) > (b)) ? (a) : (b))

int oracle_num_threads(void)
{
  int n = 1;
#ifdef _OPENMP
  #pragma omp parallel
  {
    #pragma omp single
    n = omp_get_num_threads();
  }
#endif
  return n;
}

/* linear step of the k3 grid, transfer2.c:1421-1452 / :1596-1612 */
static double k3_step_max(const oracle_setup *s)
{
  double k_step_max = 0;
  if (s->k_sampling == ORACLE_K3_CLASS_SAMPLING) {
    k_step_max = 2. * PI_ / (s->tau0 - s->tau_rec) * s->q_linstep_song;
    if (s->old_run_step)
      k_step_max = 2. * PI_ / s->rs_rec * s->q_linstep_song;
  } else if (s->k_sampling == ORACLE_K3_BESSEL_SAMPLING) {
    k_step_max = s->xx_step / s->tau0;
  }
  return k_step_max;
}

/* transfer2.c:1329-1551 */
int oracle_get_k3_size(const oracle_setup *s, const double *k3_pt, int k_pt_size,
                       oracle_k3_info *info)
{
  double k_min_pt = k3_pt[0];
  double k_max_pt = k3_pt[k_pt_size - 1];
  double k_min_tr = k_min_pt;
  double k_max_tr = k_max_pt;

  if (s->k3_extrapolation != ORACLE_NO_K3_EXTRAPOLATION) {                     /* :1365 */
    double physical_k3_range = k_max_pt - k_min_pt;
    double extended_k3_range_right = 2. * PI_ * s->extra_k3_oscillations_right / (s->tau0 - s->tau_rec);
    double extended_k3_range_left = 2. * PI_ * s->extra_k3_oscillations_left / (s->tau0 - s->tau_rec);
    if (physical_k3_range < extended_k3_range_right)
      k_max_tr += (extended_k3_range_right - physical_k3_range);
    k_max_tr = DMIN(k_max_tr, s->k[s->k_size - 1]);
    if (physical_k3_range < extended_k3_range_left)
      k_min_tr -= (extended_k3_range_left - physical_k3_range);
    k_min_tr = DMAX(k_min_tr, DMIN(s->k[0], k_min_pt));
    if (s->extra_k3_oscillations_left < 0) k_min_tr = s->k[0];
    if (s->extra_k3_oscillations_right < 0) k_max_tr = s->k[s->k_size - 1];
  }
  info->k_min = k_min_tr;
  info->k_max = k_max_tr;

  if (k_max_tr > s->xx_max / s->tau0) return ORACLE_FAIL;                      /* :1408 */

  double k_step_max = k3_step_max(s);
  if (k_step_max == 0) return ORACLE_FAIL;                                     /* :1450 */

  int index_k_pt = 0;
  int index_k_tr = 0;
  double k = k_min_tr;
  index_k_tr++;

  if (s->k3_extrapolation != ORACLE_NO_K3_EXTRAPOLATION) {                     /* :1471 */
    double k_step = k_step_max;
    if (k_step < s->smallest_allowed_variation) return ORACLE_FAIL;
    while (k < k_min_pt) {
      k += k_step;
      index_k_tr++;
    }
    if (k != k_min_pt) k = k_min_pt;
  }
  index_k_pt++;

  info->k_physical_start = index_k_tr - 1;                                     /* :1498 */

  while ((index_k_pt < k_pt_size) && ((k3_pt[index_k_pt] - k) < k_step_max)) { /* :1502 */
    k = k3_pt[index_k_pt];
    index_k_pt++;
    index_k_tr++;
  }
  info->k_physical_size = index_k_pt;                                          /* :1509 */

  while (k < k_max_tr) {                                                       /* :1513 */
    k += k_step_max;
    index_k_tr++;
    if ((k <= k_max_pt) && (k >= k_min_pt)) info->k_physical_size++;
  }
  if (k > k_max_tr)
    info->k_size = index_k_tr - 1;
  else
    info->k_size = index_k_tr;

  if (info->k_size < 2) return ORACLE_FAIL;                                    /* :1531 */
  if (s->k3_extrapolation == ORACLE_NO_K3_EXTRAPOLATION &&
      (info->k_physical_size != info->k_size || info->k_physical_start != 0))
    return ORACLE_FAIL;                                                        /* :1538-1546 */
  return ORACLE_OK;
}

/* transfer2.c:1566-1737 */
int oracle_get_k3_list(const oracle_setup *s, const double *k3_pt, int k_pt_size,
                       const oracle_k3_info *info, double *k3, int *last_used_index_pt)
{
  double k_min_pt = k3_pt[0];
  int k_tr_size = info->k_size;
  double k_min_tr = info->k_min;
  double k_max_tr = info->k_max;
  double k_step_max = k3_step_max(s);
  if (k_step_max == 0) return ORACLE_FAIL;

  int index_k_pt = 0;
  int index_k_tr = 0;
  double k = k_min_tr;
  k3[0] = k_min_tr;
  index_k_tr++;

  if (s->k3_extrapolation != ORACLE_NO_K3_EXTRAPOLATION) {                     /* :1624 */
    double k_step = k_step_max;
    if (k_step < s->smallest_allowed_variation) return ORACLE_FAIL;
    while (k < k_min_pt) {
      k += k_step;
      k3[index_k_tr] = k;
      index_k_tr++;
    }
    if (k != k_min_pt) {
      k = k_min_pt;
      k3[index_k_tr - 1] = k;
    }
  }
  index_k_pt++;

  while ((index_k_pt < k_pt_size) && ((k3_pt[index_k_pt] - k) < k_step_max)) { /* :1653 */
    k = k3_pt[index_k_pt];
    k3[index_k_tr] = k;
    index_k_pt++;
    index_k_tr++;
  }
  if (last_used_index_pt) *last_used_index_pt = index_k_pt;

  while ((index_k_tr < k_tr_size) && (k < k_max_tr)) {                         /* :1663 */
    k += k_step_max;
    k3[index_k_tr] = k;
    index_k_tr++;
  }
  if (k3[k_tr_size - 1] > k_max_tr) return ORACLE_FAIL;                        /* :1670 */

  /* DEBUG checks of :1711-1733 */
  for (int i = 0; i < k_tr_size - 1; ++i)
    if (k3[i] >= k3[i + 1]) return ORACLE_FAIL;
  return ORACLE_OK;
}

/* transfer2.c:185-203: upper bound on the number of integration points */
int oracle_time_grid_size_max(const oracle_setup *s, double k_max_global)
{
  if (s->tau_sampling_mode == ORACLE_TAU_SOURCES) return s->tau_size;
  if (s->tau_sampling_mode == ORACLE_TAU_CUSTOM) {
    double tau_step_min = 2 * PI_ / k_max_global * s->tau_linstep_song;
    double tau_max = s->tau_sampling[s->tau_size - 1];
    return s->tau_size + (int)ceil(tau_max / tau_step_min) + 1;
  }
  return s->tau_size + s->xx_size;
}

/* transfer2.c:2366-2585 */
int oracle_get_time_grid(const oracle_setup *s, double k, double *tau_grid, int *index_tau_left,
                         double *tau0_minus_tau, double *delta_tau, int *tau_grid_size_out)
{
  int tau_size_pt = s->tau_size;
  const double *tau_pt = s->tau_sampling;
  double tau_min_pt = tau_pt[0];
  double tau_step_max = 0;
  int tau_grid_size;

  if (s->tau_sampling_mode == ORACLE_TAU_SOURCES) {                            /* :2405-2413 */
    tau_grid_size = tau_size_pt;
    for (int i = 0; i < tau_grid_size; ++i) {
      tau_grid[i] = tau_pt[i];
      if (index_tau_left) index_tau_left[i] = (i < tau_size_pt - 1) ? i : tau_size_pt - 2;
    }
    goto fill_delta_tau;
  } else if (s->tau_sampling_mode == ORACLE_TAU_CUSTOM) {
    tau_step_max = 2 * PI_ / k * s->tau_linstep_song;                          /* :2422 */
  } else if (s->tau_sampling_mode == ORACLE_TAU_BESSEL) {
    tau_step_max = s->xx_step / k;                                             /* :2432 */
  }
  if (tau_step_max == 0) return ORACLE_FAIL;

  {
    int index_tau_tr = 0, index_tau_pt = 0;
    double tau = tau_min_pt;
    tau_grid[0] = tau;
    index_tau_tr++;
    index_tau_pt++;
    index_tau_left[0] = 0;
    while (index_tau_pt < tau_size_pt) {                                       /* :2491 */
      while ((tau_pt[index_tau_pt] - tau) > tau_step_max) {
        tau += tau_step_max;
        tau_grid[index_tau_tr] = tau;
        index_tau_left[index_tau_tr] = index_tau_pt;
        index_tau_tr++;
      }
      tau = tau_pt[index_tau_pt];
      tau_grid[index_tau_tr] = tau;
      index_tau_left[index_tau_tr] = index_tau_pt;
      index_tau_tr++;
      index_tau_pt++;
    }
    tau_grid_size = index_tau_tr;
    for (int i = 0; i < tau_grid_size; ++i)                                    /* :2519-2521 */
      if (index_tau_left[i] == tau_size_pt - 1) index_tau_left[i] = tau_size_pt - 2;
  }

fill_delta_tau:
  for (int i = 0; i < tau_grid_size; ++i)                                      /* :2571 */
    tau0_minus_tau[i] = s->tau0 - tau_grid[i];
  delta_tau[0] = tau_grid[1] - tau_grid[0];                                    /* :2575 */
  for (int i = 1; i < tau_grid_size - 1; ++i)
    delta_tau[i] = tau_grid[i + 1] - tau_grid[i - 1];
  delta_tau[tau_grid_size - 1] = tau_grid[tau_grid_size - 1] - tau_grid[tau_grid_size - 2];
  *tau_grid_size_out = tau_grid_size;
  return ORACLE_OK;
}

/* transfer2.c:2612-2763 */
int oracle_interpolate_sources_in_k(const oracle_setup *s, const double *k_pt, int k_pt_size,
                                    const double *src, const oracle_k3_info *info,
                                    const double *k_tr, double *sources_k_spline,
                                    double *interp)
{
  int k_tr_size = info->k_size;
  int tau_size = s->tau_size;
#define SRC(itau, ik) src[(size_t)(itau) * k_pt_size + (ik)]                   /* transfer2.h:47-51 */

  if (s->sources_k3_interpolation == ORACLE_INTERP_CUBIC)
    if (oracle_spline_table_columns(k_pt, k_pt_size, src, tau_size, sources_k_spline,
                                    ORACLE_SPLINE_EST_DERIV) != ORACLE_OK)     /* :2639 */
      return ORACLE_FAIL;

  int physical_size = info->k_physical_size;
  int first_physical_index = info->k_physical_start;
  int last_physical_index = first_physical_index + physical_size - 1;

  int index_k = 0;
  double h = k_pt[index_k + 1] - k_pt[index_k];

  for (int index_k_tr = first_physical_index; index_k_tr <= last_physical_index; ++index_k_tr) {
    while (((index_k + 1) < k_pt_size) && (k_pt[index_k + 1] < k_tr[index_k_tr])) {   /* :2683 */
      index_k++;
      h = k_pt[index_k + 1] - k_pt[index_k];
    }
    if (h == 0) return ORACLE_FAIL;
    double b = (k_tr[index_k_tr] - k_pt[index_k]) / h;
    double a = 1 - b;

    if (s->sources_k3_interpolation == ORACLE_INTERP_LINEAR) {
      for (int index_tau = 0; index_tau < tau_size; index_tau++)
        interp[(size_t)index_k_tr * tau_size + index_tau] =
          a * SRC(index_tau, index_k) + b * SRC(index_tau, index_k + 1);
    } else {
      for (int index_tau = 0; index_tau < tau_size; index_tau++)
        interp[(size_t)index_k_tr * tau_size + index_tau] =
          a * SRC(index_tau, index_k) + b * SRC(index_tau, index_k + 1)
          + ((a * a * a - a) * sources_k_spline[(size_t)index_tau * k_pt_size + index_k]
             + (b * b * b - b) * sources_k_spline[(size_t)index_tau * k_pt_size + index_k + 1]) * h * h / 6.0;
    }
  }

  if (s->k3_extrapolation != ORACLE_NO_K3_EXTRAPOLATION) {                     /* :2740-2758 */
    for (int index_k_tr = last_physical_index + 1; index_k_tr < k_tr_size; ++index_k_tr)
      if (s->k3_extrapolation == ORACLE_FLAT_K3_EXTRAPOLATION)
        for (int index_tau = 0; index_tau < tau_size; index_tau++)
          interp[(size_t)index_k_tr * tau_size + index_tau] =
            interp[(size_t)last_physical_index * tau_size + index_tau];
    for (int index_k_tr = 0; index_k_tr < first_physical_index; ++index_k_tr)
      if (s->k3_extrapolation == ORACLE_FLAT_K3_EXTRAPOLATION)
        for (int index_tau = 0; index_tau < tau_size; index_tau++)
          interp[(size_t)index_k_tr * tau_size + index_tau] =
            interp[(size_t)first_physical_index * tau_size + index_tau];
  }
#undef SRC
  return ORACLE_OK;
}

/* transfer2.c:2783-2908 */
int oracle_interpolate_sources_in_time(const oracle_setup *s, int index_k,
                                       const double *interp_k,
                                       const double *tau_tr, const int *index_tau_left,
                                       int tau_size_tr, double *sources_time_spline,
                                       double *out)
{
  int tau_size_pt = s->tau_size;
  const double *tau_pt = s->tau_sampling;
  const double *row = interp_k + (size_t)index_k * tau_size_pt;

  if (s->tau_sampling_mode == ORACLE_TAU_SOURCES) {                            /* :2808-2814 */
    for (int i = 0; i < tau_size_tr; ++i) out[i] = row[i];
    return ORACLE_OK;
  }
  if (s->sources_time_interpolation == ORACLE_INTERP_CUBIC)
    if (oracle_spline_table_columns(tau_pt, tau_size_pt, row, 1, sources_time_spline,
                                    ORACLE_SPLINE_EST_DERIV) != ORACLE_OK)     /* :2822 */
      return ORACLE_FAIL;

  for (int index_tau_tr = 0; index_tau_tr < tau_size_tr; ++index_tau_tr) {
    int index_tau = index_tau_left[index_tau_tr];
    double h = tau_pt[index_tau + 1] - tau_pt[index_tau];
    if (h == 0.) return ORACLE_FAIL;
    double b = (tau_tr[index_tau_tr] - tau_pt[index_tau]) / h;
    double a = 1 - b;
    if (s->sources_time_interpolation == ORACLE_INTERP_LINEAR) {
      out[index_tau_tr] = a * row[index_tau] + b * row[index_tau + 1];
    } else {
      out[index_tau_tr] = a * row[index_tau] + b * row[index_tau + 1]
        + ((a * a * a - a) * sources_time_spline[index_tau]
           + (b * b * b - b) * sources_time_spline[index_tau + 1]) * h * h / 6.0;
    }
  }
  return ORACLE_OK;
}

/* transfer2.c:2140-2341 */
int oracle_integrate(const oracle_setup *s, double k, int index_l, int index_m,
                     const double *const *interpolated_sources_in_time,
                     int index_J, int index_source_monopole, int L_max,
                     const double *tau0_minus_tau, const double *delta_tau, int tau_grid_size,
                     double *integral, long *n_terms)
{
  int l = s->l[index_l];
  int m = s->m[index_m];
  long terms = 0;
  *integral = 0;

  if (k * tau0_minus_tau[0] > s->xx_max) return ORACLE_FAIL;                   /* :2172 (DEBUG) */

  for (int index_tau = 0; index_tau < tau_grid_size; ++index_tau) {
    double x = k * tau0_minus_tau[index_tau];                                  /* :2195 */
    int index_x = (int)(x / s->xx_step);                                       /* :2199 */
    double a_J = (s->xx[index_x + 1] - x) / s->xx_step;                        /* :2202 */
    double integrand = 0;

    for (int index_L = 0; index_L <= L_max; ++index_L) {
      int L = index_L;                                                         /* pbs2->L[i]=i */
      if (abs(m) > ((L < l) ? L : l)) continue;                                /* :2217 */
      size_t jidx = (((size_t)index_J * s->L_size + index_L) * s->l_size + index_l) * s->m_size + index_m;
      int index_x_min = s->index_xmin_J[jidx];
      if (index_x < index_x_min) continue;                                     /* :2232 */
      int index_x_in_J = index_x - index_x_min;
      if (index_x_in_J >= s->x_size_J[jidx]) return ORACLE_FAIL;               /* :2241-2245 (DEBUG) */

      const double *J = s->J_rows[jidx] + index_x_in_J;
      double J_Llm_left = J[0];
      double J_Llm_right = J[1];
      double J_Llm;
      if (s->bessels_interpolation == ORACLE_INTERP_LINEAR) {
        J_Llm = a_J * J_Llm_left + (1 - a_J) * J_Llm_right;                    /* :2260 */
      } else {
        const double *ddJ = s->ddJ_rows[jidx] + index_x_in_J;
        double ddJ_Llm_left = ddJ[0];
        double ddJ_Llm_right = ddJ[1];
        J_Llm = (a_J * J_Llm_left +
                 (1. - a_J) * (J_Llm_right
                 - a_J * ((a_J + 1.) * ddJ_Llm_left
                 + (2. - a_J) * ddJ_Llm_right)
                 * s->xx_step * s->xx_step / 6.0));                            /* :2266-2270 */
      }
      int lm = s->lm_src[index_L * s->m_size + index_m];
      double source_Lm = interpolated_sources_in_time[index_source_monopole + lm][index_tau];
      integrand += J_Llm * source_Lm;                                          /* :2305 */
      ++terms;
    }
    *integral += integrand * delta_tau[index_tau];                             /* :2317 */
  }
  *integral *= 0.5;                                                            /* :2329 */
  if (isnan(*integral)) return ORACLE_FAIL;                                    /* :2334 */
  if (n_terms) *n_terms += terms;
  return ORACLE_OK;
}

/* transfer2.c:1861-2116 */
int oracle_compute(const oracle_setup *s, double k, int index_tt,
                   const double *const *src_t,
                   const double *tau0_minus_tau, const double *delta_tau, int tau_grid_size,
                   double *transfer_out, long *n_terms)
{
  int index_l = s->tt_index_l[index_tt];
  int index_m = s->tt_index_m[index_tt];
  double transfer = 0;

  if (s->has_source_T && index_tt >= s->index_tt2_T && index_tt < s->index_tt2_T + s->n_transfers) {
    if (oracle_integrate(s, k, index_l, index_m, src_t, s->index_J_TT, s->index_tp2_T,
                         s->l_max_los_t, tau0_minus_tau, delta_tau, tau_grid_size,
                         &transfer, n_terms) != ORACLE_OK) return ORACLE_FAIL;
  }
  else if (s->has_source_E && index_tt >= s->index_tt2_E && index_tt < s->index_tt2_E + s->n_transfers) {
    double direct_contribution = 0, mixing_contribution = 0;
    if (oracle_integrate(s, k, index_l, index_m, src_t, s->index_J_EE, s->index_tp2_E,
                         s->l_max_los_p, tau0_minus_tau, delta_tau, tau_grid_size,
                         &direct_contribution, n_terms) != ORACLE_OK) return ORACLE_FAIL;
    if (s->m[index_m] != 0)                                                    /* :1961 */
      if (oracle_integrate(s, k, index_l, index_m, src_t, s->index_J_EB, s->index_tp2_B,
                           s->l_max_los_p, tau0_minus_tau, delta_tau, tau_grid_size,
                           &mixing_contribution, n_terms) != ORACLE_OK) return ORACLE_FAIL;
    transfer = direct_contribution + mixing_contribution;                      /* :1986 */
  }
  else if (s->has_source_B && index_tt >= s->index_tt2_B && index_tt < s->index_tt2_B + s->n_transfers) {
    double direct_contribution = 0, mixing_contribution = 0;
    if (oracle_integrate(s, k, index_l, index_m, src_t, s->index_J_EE, s->index_tp2_B,
                         s->l_max_los_p, tau0_minus_tau, delta_tau, tau_grid_size,
                         &direct_contribution, n_terms) != ORACLE_OK) return ORACLE_FAIL;
    if (oracle_integrate(s, k, index_l, index_m, src_t, s->index_J_EB, s->index_tp2_E,
                         s->l_max_los_p, tau0_minus_tau, delta_tau, tau_grid_size,
                         &mixing_contribution, n_terms) != ORACLE_OK) return ORACLE_FAIL;
    transfer = direct_contribution - mixing_contribution;                      /* :2059 */
  }

  transfer /= 4.;                                                              /* :2092 */
  transfer /= (2. * s->l[index_l] + 1);                                        /* :2095 */
  transfer /= 2.;                                                              /* :2098 */
  *transfer_out = transfer;
  return ORACLE_OK;
}

/* body of the (k1,k2) loop, transfer2.c:348-547 */
int oracle_transfer2_pair(const oracle_setup *s, const double *k3_pt, int k3_pt_size,
                          const double *const *sources, const oracle_k3_info *info,
                          const double *k3_grid, double *const *out, long *n_terms)
{
  int tp2 = s->tp2_size, tau_size = s->tau_size, k_size = info->k_size;
  int status = ORACLE_OK;
  long terms_total = 0;
  double **sources_k_spline = (double **)calloc((size_t)tp2, sizeof(double *));
  double **interp_k = (double **)calloc((size_t)tp2, sizeof(double *));

  /* serial block, transfer2.c:401-431 */
  for (int tp = 0; tp < tp2; ++tp) {
    sources_k_spline[tp] = (double *)malloc((size_t)k3_pt_size * tau_size * sizeof(double));
    interp_k[tp] = (double *)malloc((size_t)k_size * tau_size * sizeof(double));
    if (oracle_interpolate_sources_in_k(s, k3_pt, k3_pt_size, sources[tp], info, k3_grid,
                                        sources_k_spline[tp], interp_k[tp]) != ORACLE_OK)
      status = ORACLE_FAIL;
  }

  int tau_max = oracle_time_grid_size_max(s, s->k[s->k_size - 1]);
  if (s->tau_sampling_mode != ORACLE_TAU_SOURCES) {
    /* transfer2.c:193 uses the largest k_max_k1k2; the global k_max bounds it */
    tau_max = oracle_time_grid_size_max(s, DMAX(info->k_max, s->k[s->k_size - 1]));
  }

  if (status == ORACLE_OK) {
    /* parallel region, transfer2.c:436-539 */
    #pragma omp parallel reduction(+ : terms_total)
    {
      double *tau_grid = (double *)malloc((size_t)tau_max * sizeof(double));
      double *tau0_minus_tau = (double *)malloc((size_t)tau_max * sizeof(double));
      double *delta_tau = (double *)malloc((size_t)tau_max * sizeof(double));
      int *index_tau_left = (int *)malloc((size_t)tau_max * sizeof(int));
      double *time_spline = (double *)malloc((size_t)tau_max * sizeof(double));
      double **src_t = (double **)malloc((size_t)tp2 * sizeof(double *));
      for (int tp = 0; tp < tp2; ++tp) src_t[tp] = (double *)malloc((size_t)tau_max * sizeof(double));

      #pragma omp for schedule(static)
      for (int index_k = 0; index_k < k_size; ++index_k) {
        double k = k3_grid[index_k];
        int tau_grid_size = 0;
        if (oracle_get_time_grid(s, k, tau_grid, index_tau_left, tau0_minus_tau, delta_tau,
                                 &tau_grid_size) != ORACLE_OK) { status = ORACLE_FAIL; continue; }
        for (int tp = 0; tp < tp2; ++tp)
          if (oracle_interpolate_sources_in_time(s, index_k, interp_k[tp], tau_grid, index_tau_left,
                                                 tau_grid_size, time_spline, src_t[tp]) != ORACLE_OK)
            status = ORACLE_FAIL;
        for (int index_tt = 0; index_tt < s->tt2_size; ++index_tt) {
          double t;
          long nt = 0;
          if (oracle_compute(s, k, index_tt, (const double *const *)src_t, tau0_minus_tau, delta_tau,
                             tau_grid_size, &t, &nt) != ORACLE_OK) { status = ORACLE_FAIL; continue; }
          out[index_tt][index_k] = t;                                          /* :2107 */
          terms_total += nt;
        }
      }
      for (int tp = 0; tp < tp2; ++tp) free(src_t[tp]);
      free(src_t); free(time_spline); free(index_tau_left);
      free(delta_tau); free(tau0_minus_tau); free(tau_grid);
    }
  }

  for (int tp = 0; tp < tp2; ++tp) { free(sources_k_spline[tp]); free(interp_k[tp]); }
  free(sources_k_spline); free(interp_k);
  if (n_terms) *n_terms += terms_total;
  return status;
}

/* serial pair loop, transfer2.c:315,348 */
int oracle_transfer2_run(const oracle_setup *s, int n_pairs, const int *k3_pt_size,
                         const long *pt_off, const double *k3_pt, const long *src_off,
                         const double *src, const oracle_k3_info *info, const long *tr_off,
                         const double *k3_tr, long n_units_total, double *out, long *n_terms)
{
  int tp2 = s->tp2_size;
  const double **sources = (const double **)malloc((size_t)tp2 * sizeof(double *));
  double **outp = (double **)malloc((size_t)s->tt2_size * sizeof(double *));
  int status = ORACLE_OK;
  for (int p = 0; p < n_pairs && status == ORACLE_OK; ++p) {
    int np = k3_pt_size[p];
    for (int tp = 0; tp < tp2; ++tp)
      sources[tp] = src + src_off[p] + (size_t)tp * s->tau_size * np;
    for (int tt = 0; tt < s->tt2_size; ++tt)
      outp[tt] = out + (size_t)tt * n_units_total + tr_off[p];
    status = oracle_transfer2_pair(s, k3_pt + pt_off[p], np, sources, &info[p],
                                   k3_tr + tr_off[p], outp, n_terms);
  }
  free(sources); free(outp);
  return status;
}
