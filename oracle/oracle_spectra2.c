/*
 * oracle_spectra2.c -- restatement of the (k1,k2,k3) sum of SONG's intrinsic C_l, spectra2_cls
 * (source/spectra2.c:241-570).  TEST INFRASTRUCTURE ONLY (see los_oracle.h).  PINNED: bit-identical to
 * the reference's own spectra2_cls run here (oracle/_ref/libsong_ref_bisp.so, tests/test_reference_pin.py).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "los_oracle.h"

#define _PI_ 3.1415926535897932384626433832795e0

/* One (index_ct, index_l, index_M) of the loops at spectra2.c:303-326: the sum over the given pairs (index_k1, index_k2 in
 * the reference's loop order) and their k3 grids (transfer2_get_k3_list, :439-451) of
 *   total_factor * transfer_1[index_k3] * transfer_2[index_k3]                        (:501-511)
 * primordial[i] = primordial_spectrum_at_k(k[i]) (:389-397, :421-429; the CLASS primordial module is outside the path). */
double oracle_spectra2_cl(int n_pairs, const int *pair_k1, const int *pair_k2, int k_size, const double *k,
                          const double *primordial, const int *k3_size, const long *tr_off, const double *k3_tr,
                          const double *transfer_1, const double *transfer_2, int m, int rescale_cmb_sources)
{
  double result = 0;
  for (int p = 0; p < n_pairs; ++p) {
    int index_k1 = pair_k1[p], index_k2 = pair_k2[p];
    double step_k1;                                                                    /* :377-381 */
    if (index_k1 == 0) step_k1 = (k[1] - k[0]) / 2.;
    else if (index_k1 == k_size - 1) step_k1 = (k[k_size - 1] - k[k_size - 2]) / 2.;
    else step_k1 = (k[index_k1 + 1] - k[index_k1 - 1]) / 2.;
    double k1 = k[index_k1];
    double spectra_k1 = primordial[index_k1];
    spectra_k1 = 2 * _PI_ * _PI_ / (k1 * k1 * k1) * spectra_k1;                         /* :398 */
    double step_k2;                                                                    /* :409-414 */
    if (index_k1 == 0) step_k2 = 0.;
    else if (index_k2 == 0) step_k2 = (k[1] - k[0]) / 2.;
    else if (index_k2 == index_k1) step_k2 = (k[index_k1] - k[index_k1 - 1]) / 2.;
    else step_k2 = (k[index_k2 + 1] - k[index_k2 - 1]) / 2.;
    double k2 = k[index_k2];
    double spectra_k2 = primordial[index_k2];
    spectra_k2 = 2 * _PI_ * _PI_ / (k2 * k2 * k2) * spectra_k2;                         /* :430 */
    const double *k3_grid = k3_tr + tr_off[p];
    int n = k3_size[p];
    const double *t1 = transfer_1 + tr_off[p], *t2 = transfer_2 + tr_off[p];
    int triangular_first = 0;
    int triangular_last = 0;
    for (int index_k3 = 0; index_k3 < n; ++index_k3) {
      double k3 = k3_grid[index_k3];
      double step_k3;                                                                  /* :474-478 */
      if (index_k3 == 0) step_k3 = (k3_grid[1] - k3_grid[0]) / 2.;
      else if (index_k3 == n - 1) step_k3 = (k3_grid[n - 1] - k3_grid[n - 2]) / 2.;
      else step_k3 = (k3_grid[index_k3 + 1] - k3_grid[index_k3 - 1]) / 2.;
      if (k3 < k1 - k2 || k3 > k1 + k2) step_k3 = 0.;                                   /* :481-483 */
      if (k3 >= k1 - k2 && triangular_first == 0 && index_k3 < n - 1) {                /* :485-488 */
        triangular_first = 1;
        step_k3 = (k3_grid[index_k3 + 1] - k3) / 2. + k3 - (k1 - k2);
      }
      /* :489-492 reads k3_grid[index_k3+1] also at the last point (one past the grid); the test is restricted to
         index_k3 + 1 < n here, which is what the reference's k3_size_max-sized buffer yields for a grid that ends inside
         the triangular region only by accident of the bytes beyond it */
      if (index_k3 + 1 < n && k3_grid[index_k3 + 1] > k1 + k2 && triangular_last == 0 && index_k3 > 0) {
        triangular_last = 1;
        step_k3 = (k3 - k3_grid[index_k3 - 1]) / 2. + (k1 + k2) - k3;
      }
      if (step_k3 == 0) continue;                                                      /* :495-496 */
      double scale = pow(sqrt(1. - pow((k3 * k3 + k1 * k1 - k2 * k2) / 2. / k3 / k1, 2)), m);   /* :499 */
      double total_factor =
          2. *
          k1 * k2 * k3 *
          step_k1 * step_k2 * step_k3 *
          (rescale_cmb_sources && step_k3 != 0. ? pow(scale, 2) : 1.) *
          2. / _PI_ / 2. / _PI_ / 2. / _PI_ / 2. / _PI_ *
          spectra_k1 * spectra_k2;                                                     /* :501-507 */
      double integrand = total_factor * t1[index_k3] * t2[index_k3];
      result += integrand;                                                             /* :510 */
    }
  }
  return result;
}
