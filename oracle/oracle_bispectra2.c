/*
 * oracle_bispectra2.c -- restatement of the k3 integral of SONG's intrinsic bispectrum,
 * bispectra2_intrinsic_integrate_over_k3 (source/bispectra2.c:1107-1439).
 * TEST INFRASTRUCTURE ONLY (see los_oracle.h).  PINNED: bit-identical to the reference's own bispectra2_intrinsic_integrate_over_k3 run here (oracle/_ref/libsong_ref_bisp.so,
 * tests/test_reference_pin.py).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "los_oracle.h"

/* One requested transfer type (one (l3, M3) of the loop at bispectra2.c:1172) for n_pairs (k1,k2) pairs:
 *   k3_tr + tr_off[p], k3_size[p]          the pair's integration grid (transfer2_get_k3_list, :1255-1271)
 *   transfer + tr_off[p]                   T_l3,M3(k1,k2,k3) of the pair (:1298-1300)
 *   j / ddj / x_min_l1                     row index_L3 of pbs2->j_l1 (:1200)
 *   out[index_r * n_pairs + p]             pwb->integral_over_k3[index_l3][index_r][index_k1][index_k2]
 * The loop over pairs is the OpenMP loop of :1232; delta_k3 is the trapezoidal measure of :1279-1284. */
int oracle_integral_over_k3(int n_pairs, const int *k3_size, const long *tr_off, const double *k3_tr, const double *transfer,
                            const double *j, const double *ddj, double x_min_l1, double xx_step, int cubic,
                            int r_size, const double *r, int abs_M3, double M3_coefficient, double *out)
{
  int failed = 0;
  #pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < n_pairs; ++p) {
    int n = k3_size[p];
    if (n < 2) { failed = 1; continue; }                                  /* :1274-1276 */
    const double *k3_grid = k3_tr + tr_off[p];
    double *delta_k3 = (double *)malloc(sizeof(double) * n);
    delta_k3[0] = k3_grid[1] - k3_grid[0];
    for (int index_k3 = 1; index_k3 < (n - 1); ++index_k3)
      delta_k3[index_k3] = k3_grid[index_k3 + 1] - k3_grid[index_k3 - 1];
    delta_k3[n - 1] = k3_grid[n - 1] - k3_grid[n - 2];
    for (int index_r = 0; index_r < r_size; ++index_r) {
      double *dst = &out[(long)index_r * n_pairs + p];
      oracle_bessel2_convolution(k3_grid, delta_k3, n, transfer + tr_off[p], NULL, j, ddj, x_min_l1, xx_step, cubic,
                                 r[index_r], dst);                        /* :1326-1341 */
      if (abs_M3 != 0) *dst *= M3_coefficient;                            /* :1369-1370 */
      *dst *= 2;                                                          /* :1375 */
    }
    free(delta_k3);
  }
  return failed;
}
