// los_common.h -- device-side parameter blocks shared by the LOS kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace songlos {

// Dense projection table.  Row ix holds, for every column c = (index_J*L_size + L)*n_transfers + lm,
// the pair {A, B} = {J[ix], J[ix+1]} when ix >= index_xmin_J(c) and {0, 0} otherwise, which is
// exactly what the skip test of transfer2.c:2232 leaves in the sum.
struct ProjTable {
  const double2 *J;      // [xx_size][n_cols]
  const double2 *ddJ;    // same layout, cubic Bessel interpolation only (transfer2.c:2262-2271)
  int64_t n_cols;
  const unsigned char *block_nonzero;  // [J_size*L_size]: 0 if every column of (J,L) is empty
};

struct TypeTables {
  int tt2_size, n_transfers, L_size, m_size, l_size;
  const int *tt_type;    // 0=T 1=E 2=B                  [tt2_size]
  const int *tt_lm;      // column inside the type block [tt2_size]
  const int *tt_l;       // multipole value l            [tt2_size]
  const int *tt_m;       // azimuthal value m            [tt2_size]
  const int *tt_im;      // index_m                      [tt2_size]
  const int *lm_src;     // [L_size][m_size]
  int index_J_TT, index_J_EE, index_J_EB;
  int index_tp2_T, index_tp2_E, index_tp2_B;
  int l_max_los_t, l_max_los_p;
  int tp2_size;
};

struct TimeTables {
  int tau_size;
  const double *tau;            // ppt2->tau_sampling
  const double *tau0_minus_tau; // transfer2.c:2572
  const double *delta_tau;      // transfer2.c:2575-2580
  double tau0;
  int mode;                     // SONG_LOS_TAU_*
  double tau_linstep_song;
};

struct Batch {
  int n_pairs;
  const int *np, *nt, *phys_start, *phys_size;
  const int64_t *pt_off, *tr_off, *src_off;
  const double *k3_pt, *k3_tr, *src, *dd;
  const double *phys_k_first, *phys_k_last;   // k3 the wings of the flat extrapolation are interpolated at (transfer2.c:2743-2757)
  int64_t n_units;
  double *out;                  // [tt2_size][n_units]
};

struct ExactParams {
  ProjTable proj;
  TypeTables ty;
  TimeTables tm;
  Batch b;
  const double *xx; int xx_size; double xx_step;
  int cubic_k3, cubic_J, cubic_time, flat_extrap;
  const int2 *tiles;            // (pair, first unit) per CTA
  int tau_chunk;                // tau points staged per pass
  int *err;                     // device error flags
  // custom / bessel tau sampling scratch (per CTA unit slot)
  double *scratch; int64_t scratch_stride; int tau_grid_max;
};

enum { ERR_X_OUT_OF_RANGE = 1, ERR_H_ZERO = 2, ERR_NAN = 4, ERR_TAU_GRID_OVERFLOW = 8 };

constexpr int SPLINE_MAX_NP = 256;          // k3 grids up to this length: spline workspace in thread-local memory
constexpr int EXACT_UNITS_PER_CTA = 8;     // one warp per k3 value
constexpr int EXACT_THREADS = 256;
constexpr int EXACT_MAX_TT_PASSES = 16;    // 16*32 = 512 transfer types per unit pass

}  // namespace songlos
