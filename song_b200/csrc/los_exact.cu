// los_exact.cu -- EXACT-mode kernels: the reference's statement order, no FMA contraction.
//
// This translation unit is compiled with -fmad=false, so every `a*b + c` below is a separate
// IEEE multiply and add evaluated left to right, as gcc -ffp-contract=off does for the oracle.
// Division and int conversion are IEEE-correct on sm_100a, hence the results are bit-identical
// to the reference statement order (transfer2.c line numbers quoted per kernel).
#include "los_common.h"

namespace songlos {

// ---------------------------------------------------------------------------------------------
// K2: second derivatives in k3 of every (pair, source type, tau) column.
// Replaces the serial array_spline_table_columns call of transfer2.c:2639-2648
// (_SPLINE_EST_DERIV_); one thread per column, columns of a pair share the abscissa k3_pt.
// ---------------------------------------------------------------------------------------------


__global__ void __launch_bounds__(128) spline_k3_kernel(Batch b, int tau_size, int tp2_size,
                                                         double *dd_out, double *u_scratch, int *err)
{
  int64_t cols_per_pair = (int64_t)tp2_size * tau_size;
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols_per_pair * b.n_pairs) return;
  int p = (int)(c / cols_per_pair);
  int64_t col = c - (int64_t)p * cols_per_pair;
  int n = b.np[p];
  const double *x = b.k3_pt + b.pt_off[p];
  const double *y = b.src + b.src_off[p] + col * n;
  double *dd = dd_out + b.src_off[p] + col * n;
  // right-hand sides of the elimination: thread-local for the usual grid lengths, in global scratch (laid out like dd)
  // for longer k3 grids, so that any k3_size works as in the reference
  double u_local[SPLINE_MAX_NP];
  double *u = u_local;
  if (n < 3) { atomicOr(err, ERR_H_ZERO); return; }
  if (n > SPLINE_MAX_NP) {
    if (u_scratch == nullptr) { atomicOr(err, ERR_H_ZERO); return; }
    u = u_scratch + b.src_off[p] + col * n;
  }

  double dy_first = ((x[2] - x[0]) * (x[2] - x[0]) * (y[1] - y[0])
                     - (x[1] - x[0]) * (x[1] - x[0]) * (y[2] - y[0]))
                    / ((x[2] - x[0]) * (x[1] - x[0]) * (x[2] - x[1]));
  double prev_dd = -0.5;
  dd[0] = prev_dd;
  u[0] = (3. / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - dy_first);

  for (int i = 1; i < n - 1; ++i) {
    double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    double pp = sig * prev_dd + 2.0;
    prev_dd = (sig - 1.0) / pp;
    dd[i] = prev_dd;
    double ui = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
    u[i] = (6.0 * ui / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / pp;
  }

  double dy_last = ((x[n - 3] - x[n - 1]) * (x[n - 3] - x[n - 1]) * (y[n - 2] - y[n - 1])
                    - (x[n - 2] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (y[n - 3] - y[n - 1]))
                   / ((x[n - 3] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (x[n - 3] - x[n - 2]));
  double qn = 0.5;
  double un = (3. / (x[n - 1] - x[n - 2])) * (dy_last - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
  double next = (un - qn * u[n - 2]) / (qn * prev_dd + 1.0);
  dd[n - 1] = next;
  for (int i = n - 2; i >= 0; --i) {
    next = dd[i] * next + u[i];
    dd[i] = next;
  }
}

// ---------------------------------------------------------------------------------------------
// K1 (exact): one CTA per (pair, 8 consecutive k3 values); one warp per k3 value, lanes over
// transfer types.  Per tau chunk the CTA first evaluates the k3 interpolation of every source
// type for its 8 k3 values into shared memory (transfer2.c:2681-2757, never materialised in
// HBM), together with index_x and a_J (transfer2.c:2195-2202, shared by all l), then every lane
// runs the tau / L loops of transfer2_integrate (transfer2.c:2192-2326) for its transfer type.
// ---------------------------------------------------------------------------------------------
struct UnitInfo {
  double k;        // k3 value of the unit (argument of J)
  double a, b, h;  // spline weights of the (clamped) interpolation point
  int j;           // left bracket in k3_pt
};

__device__ __forceinline__ double lerp_J(const ProjTable &proj, int cubic_J, int64_t rowcol,
                                         double a_J, double xx_step)
{
  double2 jj = __ldg(&proj.J[rowcol]);
  if (!cubic_J)
    return a_J * jj.x + (1 - a_J) * jj.y;                                  // transfer2.c:2260
  double2 dj = __ldg(&proj.ddJ[rowcol]);
  return (a_J * jj.x + (1. - a_J) * (jj.y - a_J * ((a_J + 1.) * dj.x + (2. - a_J) * dj.y)
                                     * xx_step * xx_step / 6.0));          // transfer2.c:2266-2270
}

__global__ void __launch_bounds__(EXACT_THREADS) los_exact_kernel(ExactParams P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int KT = EXACT_UNITS_PER_CTA;
  const int TC = P.tau_chunk;
  const int tp2 = P.ty.tp2_size;
  double *Sk = reinterpret_cast<double *>(smem_raw);                 // [KT][TC][tp2]
  double *aJs = Sk + (size_t)KT * TC * tp2;                          // [KT][TC]
  int *ixs = reinterpret_cast<int *>(aJs + (size_t)KT * TC);         // [KT][TC]
  __shared__ UnitInfo unit[EXACT_UNITS_PER_CTA];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int2 tile = P.tiles[blockIdx.x];
  const int p = tile.x, u0 = tile.y;
  const int np = P.b.np[p], nt = P.b.nt[p];
  const int first = P.b.phys_start[p], last = first + P.b.phys_size[p] - 1;
  const double *kp = P.b.k3_pt + P.b.pt_off[p];
  const double *kt = P.b.k3_tr + P.b.tr_off[p];
  const double *src = P.b.src + P.b.src_off[p];
  const double *dd = P.b.dd + P.b.src_off[p];
  const int nunits = min(KT, nt - u0);
  const int tau_size = P.tm.tau_size;

  if (tid < nunits) {
    int it = u0 + tid;
    double ke = kt[it];
    if (P.flat_extrap) ke = (it < first) ? P.b.phys_k_first[p] : (it > last) ? P.b.phys_k_last[p] : ke;   // transfer2.c:2743-2757
    int j = 0;
    while ((j + 1) < np && kp[j + 1] < ke) ++j;                           // transfer2.c:2683 (strict <)
    if (j > np - 2) j = np - 2;
    double h = kp[j + 1] - kp[j];
    if (h == 0) atomicOr(P.err, ERR_H_ZERO);
    double b = (ke - kp[j]) / h;                                          // transfer2.c:2690
    UnitInfo ui;
    ui.k = kt[it]; ui.b = b; ui.a = 1 - b; ui.h = h; ui.j = j;
    unit[tid] = ui;
  }
  __syncthreads();

  double acc_direct[EXACT_MAX_TT_PASSES], acc_mixing[EXACT_MAX_TT_PASSES];
  const int npass = (P.ty.tt2_size + 31) / 32;
  for (int q = 0; q < EXACT_MAX_TT_PASSES; ++q) { acc_direct[q] = 0; acc_mixing[q] = 0; }

  for (int c0 = 0; c0 < tau_size; c0 += TC) {
    const int tc = min(TC, tau_size - c0);

    // ---- stage A: x, index_x, a_J per (unit, tau) ----
    for (int idx = tid; idx < nunits * tc; idx += EXACT_THREADS) {
      int u = idx / tc, itau = idx - u * tc;
      double x = unit[u].k * P.tm.tau0_minus_tau[c0 + itau];              // transfer2.c:2195
      int index_x = (int)(x / P.xx_step);                                 // transfer2.c:2199
      if (index_x < 0 || index_x + 1 >= P.xx_size) {
        atomicOr(P.err, ERR_X_OUT_OF_RANGE);                              // transfer2.c:2172
        index_x = 0;
      }
      aJs[u * TC + itau] = (P.xx[index_x + 1] - x) / P.xx_step;           // transfer2.c:2202
      ixs[u * TC + itau] = index_x;
    }
    // ---- stage A: k3 interpolation of the sources ----
    for (int idx = tid; idx < nunits * tc * tp2; idx += EXACT_THREADS) {
      int itau = idx % tc;
      int r = idx / tc;
      int tp = r % tp2;
      int u = r / tp2;
      const UnitInfo ui = unit[u];
      int64_t o = ((int64_t)tp * tau_size + (c0 + itau)) * np + ui.j;
      double s0 = src[o], s1 = src[o + 1];
      double v;
      if (!P.cubic_k3) {
        v = ui.a * s0 + ui.b * s1;                                        // transfer2.c:2697
      } else {
        double a = ui.a, b = ui.b, h = ui.h;
        v = a * s0 + b * s1
            + ((a * a * a - a) * dd[o] + (b * b * b - b) * dd[o + 1]) * h * h / 6.0;  // :2703-2705
      }
      Sk[((size_t)u * TC + itau) * tp2 + tp] = v;
    }
    __syncthreads();

    // ---- stage B: the LOS integral, one warp per k3, lanes over transfer types ----
    if (warp < nunits) {
      for (int q = 0; q < npass; ++q) {
        int tt = q * 32 + lane;
        if (tt >= P.ty.tt2_size) continue;
        const int type = P.ty.tt_type[tt];
        const int lm = P.ty.tt_lm[tt];
        const int im = P.ty.tt_im[tt];
        const int mval = P.ty.tt_m[tt];
        int J1, tp1, J2 = -1, tp2b = -1, Lmax;
        if (type == 0)      { J1 = P.ty.index_J_TT; tp1 = P.ty.index_tp2_T; Lmax = P.ty.l_max_los_t; }
        else if (type == 1) { J1 = P.ty.index_J_EE; tp1 = P.ty.index_tp2_E; Lmax = P.ty.l_max_los_p;
                              if (mval != 0) { J2 = P.ty.index_J_EB; tp2b = P.ty.index_tp2_B; } }   // :1961
        else                { J1 = P.ty.index_J_EE; tp1 = P.ty.index_tp2_B; Lmax = P.ty.l_max_los_p;
                              J2 = P.ty.index_J_EB; tp2b = P.ty.index_tp2_E; }                      // :2034
        double integral1 = acc_direct[q], integral2 = acc_mixing[q];
        for (int itau = 0; itau < tc; ++itau) {
          const double a_J = aJs[warp * TC + itau];
          const int64_t row = (int64_t)ixs[warp * TC + itau] * P.proj.n_cols;
          const double *S = Sk + ((size_t)warp * TC + itau) * tp2;
          double integrand = 0;
          for (int L = 0; L <= Lmax; ++L) {
            int ls = P.ty.lm_src[L * P.ty.m_size + im];
            if (ls < 0) continue;                                         // |m| > L, :2217
            if (!P.proj.block_nonzero[J1 * P.ty.L_size + L]) continue;    // all-zero (J,L) block
            double Jv = lerp_J(P.proj, P.cubic_J, row + ((int64_t)J1 * P.ty.L_size + L) * P.ty.n_transfers + lm,
                               a_J, P.xx_step);
            integrand += Jv * S[tp1 + ls];                                // :2305
          }
          integral1 += integrand * P.tm.delta_tau[c0 + itau];             // :2317
          if (J2 >= 0) {
            double integrand2 = 0;
            for (int L = 0; L <= Lmax; ++L) {
              int ls = P.ty.lm_src[L * P.ty.m_size + im];
              if (ls < 0) continue;
              if (!P.proj.block_nonzero[J2 * P.ty.L_size + L]) continue;
              double Jv = lerp_J(P.proj, P.cubic_J, row + ((int64_t)J2 * P.ty.L_size + L) * P.ty.n_transfers + lm,
                                 a_J, P.xx_step);
              integrand2 += Jv * S[tp2b + ls];
            }
            integral2 += integrand2 * P.tm.delta_tau[c0 + itau];
          }
        }
        acc_direct[q] = integral1;
        acc_mixing[q] = integral2;
      }
    }
    __syncthreads();
  }

  // ---- epilogue: trapezoid 1/2, E/B mixing, the three divisions of transfer2.c:2092-2098 ----
  if (warp < nunits) {
    for (int q = 0; q < npass; ++q) {
      int tt = q * 32 + lane;
      if (tt >= P.ty.tt2_size) continue;
      const int type = P.ty.tt_type[tt];
      double direct = acc_direct[q] * 0.5;                                // :2329
      double mixing = acc_mixing[q] * 0.5;
      double t;
      if (type == 0) t = direct;
      else if (type == 1) t = direct + mixing;                            // :1986
      else t = direct - mixing;                                           // :2059
      t /= 4.;
      t /= (2. * P.ty.tt_l[tt] + 1);
      t /= 2.;
      if (t != t) atomicOr(P.err, ERR_NAN);                               // :2334
      P.b.out[(int64_t)tt * P.b.n_units + P.b.tr_off[p] + u0 + warp] = t; // :2107
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1 (exact, custom / bessel time sampling): transfer2_get_time_grid builds a k-dependent integration
// grid (transfer2.c:2416-2524: the source times plus extra points so that no step exceeds
// 2 pi/k * tau_linstep_song or xx_step/k) and transfer2_interpolate_sources_in_time resamples the
// k3-interpolated sources on it (transfer2.c:2816-2880, linear or cubic spline in tau).
// Persistent CTAs, 8 consecutive k3 of one pair at a time as above; per tile
//   1. k3 interpolation of every source type at EVERY source time        -> rowS (global scratch)
//   2. one thread per k3 value walks the grid construction               -> grid, left (scratch)
//   3. time splines of the rows, one thread per (k3, source type)         -> rowD (scratch)
//   4. the tau / L loops in chunks of the unit's own grid, sources and trapezoid weights resampled
//      into shared memory first.
// Same statement order as the reference: bit-identical to the oracle.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EXACT_THREADS) los_exact_resample_kernel(ExactParams P, int n_tiles)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int KT = EXACT_UNITS_PER_CTA;
  const int TC = P.tau_chunk;
  const int tp2 = P.ty.tp2_size;
  const int tau_size = P.tm.tau_size;
  const int gmax = P.tau_grid_max;
  double *Sk = reinterpret_cast<double *>(smem_raw);                 // [KT][TC][tp2]
  double *aJs = Sk + (size_t)KT * TC * tp2;                          // [KT][TC]
  double *dts = aJs + (size_t)KT * TC;                               // [KT][TC] trapezoid weights
  int *ixs = reinterpret_cast<int *>(dts + (size_t)KT * TC);         // [KT][TC]
  __shared__ UnitInfo unit[EXACT_UNITS_PER_CTA];
  __shared__ int gsize[EXACT_UNITS_PER_CTA];

  // per-CTA scratch in global memory
  double *rowS = P.scratch + (int64_t)blockIdx.x * P.scratch_stride;  // [KT][tp2][tau_size]
  double *rowD = rowS + (size_t)KT * tp2 * tau_size;                  // time-spline second derivatives
  double *rowU = rowD + (size_t)KT * tp2 * tau_size;                  // tridiagonal work array
  double *grid = rowU + (size_t)KT * tp2 * tau_size;                  // [KT][gmax]
  int *left = reinterpret_cast<int *>(grid + (size_t)KT * gmax);      // [KT][gmax]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double *tau_pt = P.tm.tau;
  const int npass = (P.ty.tt2_size + 31) / 32;

  for (int tile_i = blockIdx.x; tile_i < n_tiles; tile_i += gridDim.x) {
    const int2 tile = P.tiles[tile_i];
    const int p = tile.x, u0 = tile.y;
    const int np = P.b.np[p], nt = P.b.nt[p];
    const int first = P.b.phys_start[p], last = first + P.b.phys_size[p] - 1;
    const double *kp = P.b.k3_pt + P.b.pt_off[p];
    const double *kt = P.b.k3_tr + P.b.tr_off[p];
    const double *src = P.b.src + P.b.src_off[p];
    const double *dd = P.b.dd + P.b.src_off[p];
    const int nunits = min(KT, nt - u0);

    if (tid < nunits) {
      int it = u0 + tid;
      double ke = kt[it];
      if (P.flat_extrap) ke = (it < first) ? P.b.phys_k_first[p] : (it > last) ? P.b.phys_k_last[p] : ke;   // transfer2.c:2743-2757
      int j = 0;
      while ((j + 1) < np && kp[j + 1] < ke) ++j;                         // transfer2.c:2683 (strict <)
      if (j > np - 2) j = np - 2;
      double h = kp[j + 1] - kp[j];
      if (h == 0) atomicOr(P.err, ERR_H_ZERO);
      double b = (ke - kp[j]) / h;                                        // transfer2.c:2690
      UnitInfo ui;
      ui.k = kt[it]; ui.b = b; ui.a = 1 - b; ui.h = h; ui.j = j;
      unit[tid] = ui;

      // ---- 2. the integration grid of this k (transfer2.c:2416-2524) ----
      const double k = ui.k;
      double tau_step_max = (P.tm.mode == 1) ? 2 * 3.1415926535897932384626433832795 / k * P.tm.tau_linstep_song   // :2422
                                             : P.xx_step / k;                                                        // :2432
      double *g = grid + (size_t)tid * gmax;
      int *lf = left + (size_t)tid * gmax;
      int n = 0;
      if (!(tau_step_max > 0)) { atomicOr(P.err, ERR_TAU_GRID_OVERFLOW); gsize[tid] = 0; }
      else {
        double tau = tau_pt[0];
        g[0] = tau; lf[0] = 0; n = 1;
        int index_tau_pt = 1;
        bool overflow = false;
        while (index_tau_pt < tau_size && !overflow) {                    // :2491
          while ((tau_pt[index_tau_pt] - tau) > tau_step_max) {
            tau += tau_step_max;
            if (n >= gmax) { overflow = true; break; }
            g[n] = tau; lf[n] = index_tau_pt; ++n;
          }
          if (overflow || n >= gmax) { overflow = true; break; }
          tau = tau_pt[index_tau_pt];
          g[n] = tau; lf[n] = index_tau_pt; ++n;
          ++index_tau_pt;
        }
        if (overflow) { atomicOr(P.err, ERR_TAU_GRID_OVERFLOW); n = 0; }
        for (int i = 0; i < n; ++i) if (lf[i] == tau_size - 1) lf[i] = tau_size - 2;   // :2519-2521
        gsize[tid] = n;
      }
    }
    __syncthreads();

    // ---- 1. k3 interpolation at every source time ----
    for (int idx = tid; idx < nunits * tp2 * tau_size; idx += EXACT_THREADS) {
      int itau = idx % tau_size;
      int r = idx / tau_size;
      int tp = r % tp2;
      int u = r / tp2;
      const UnitInfo ui = unit[u];
      int64_t o = ((int64_t)tp * tau_size + itau) * np + ui.j;
      double s0 = src[o], s1 = src[o + 1];
      double v;
      if (!P.cubic_k3) v = ui.a * s0 + ui.b * s1;                         // transfer2.c:2697
      else {
        double a = ui.a, b = ui.b, h = ui.h;
        v = a * s0 + b * s1 + ((a * a * a - a) * dd[o] + (b * b * b - b) * dd[o + 1]) * h * h / 6.0;   // :2703-2705
      }
      rowS[((size_t)u * tp2 + tp) * tau_size + itau] = v;
    }
    __syncthreads();

    // ---- 3. time splines (transfer2.c:2822, array_spline_table_columns with one column, _SPLINE_EST_DERIV_) ----
    if (P.cubic_time) {
      for (int c = tid; c < nunits * tp2; c += EXACT_THREADS) {
        const double *x = tau_pt;
        const double *y = rowS + (size_t)c * tau_size;
        double *d2 = rowD + (size_t)c * tau_size;
        double *uu = rowU + (size_t)c * tau_size;
        const int n = tau_size;
        double dy_first = ((x[2] - x[0]) * (x[2] - x[0]) * (y[1] - y[0]) - (x[1] - x[0]) * (x[1] - x[0]) * (y[2] - y[0]))
                          / ((x[2] - x[0]) * (x[1] - x[0]) * (x[2] - x[1]));
        double prev_dd = -0.5;
        d2[0] = prev_dd;
        uu[0] = (3. / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - dy_first);
        for (int i = 1; i < n - 1; ++i) {
          double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
          double pp = sig * prev_dd + 2.0;
          prev_dd = (sig - 1.0) / pp;
          d2[i] = prev_dd;
          double ui = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
          uu[i] = (6.0 * ui / (x[i + 1] - x[i - 1]) - sig * uu[i - 1]) / pp;
        }
        double dy_last = ((x[n - 3] - x[n - 1]) * (x[n - 3] - x[n - 1]) * (y[n - 2] - y[n - 1])
                          - (x[n - 2] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (y[n - 3] - y[n - 1]))
                         / ((x[n - 3] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (x[n - 3] - x[n - 2]));
        double qn = 0.5;
        double un = (3. / (x[n - 1] - x[n - 2])) * (dy_last - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
        double next = (un - qn * uu[n - 2]) / (qn * prev_dd + 1.0);
        d2[n - 1] = next;
        for (int i = n - 2; i >= 0; --i) { next = d2[i] * next + uu[i]; d2[i] = next; }
      }
      __syncthreads();
    }

    double acc_direct[EXACT_MAX_TT_PASSES], acc_mixing[EXACT_MAX_TT_PASSES];
    for (int q = 0; q < EXACT_MAX_TT_PASSES; ++q) { acc_direct[q] = 0; acc_mixing[q] = 0; }
    int gmax_tile = 0;
    for (int u = 0; u < nunits; ++u) gmax_tile = max(gmax_tile, gsize[u]);

    for (int c0 = 0; c0 < gmax_tile; c0 += TC) {
      // ---- stage A: x, index_x, a_J, trapezoid weight per (unit, grid point) ----
      for (int idx = tid; idx < nunits * TC; idx += EXACT_THREADS) {
        int u = idx / TC, i = idx - u * TC;
        const int gi = c0 + i, n = gsize[u];
        if (gi >= n) continue;
        const double *g = grid + (size_t)u * gmax;
        double x = unit[u].k * (P.tm.tau0 - g[gi]);                       // transfer2.c:2572, 2195
        int index_x = (int)(x / P.xx_step);                               // transfer2.c:2199
        if (index_x < 0 || index_x + 1 >= P.xx_size) { atomicOr(P.err, ERR_X_OUT_OF_RANGE); index_x = 0; }   // :2172
        aJs[u * TC + i] = (P.xx[index_x + 1] - x) / P.xx_step;            // transfer2.c:2202
        ixs[u * TC + i] = index_x;
        double dt;                                                        // transfer2.c:2575-2580
        if (gi == 0) dt = g[1] - g[0];
        else if (gi == n - 1) dt = g[n - 1] - g[n - 2];
        else dt = g[gi + 1] - g[gi - 1];
        dts[u * TC + i] = dt;
      }
      // ---- stage A: sources resampled in time (transfer2.c:2866-2880) ----
      for (int idx = tid; idx < nunits * TC * tp2; idx += EXACT_THREADS) {
        int i = idx % TC;
        int r = idx / TC;
        int tp = r % tp2;
        int u = r / tp2;
        const int gi = c0 + i;
        if (gi >= gsize[u]) continue;
        const double tau = grid[(size_t)u * gmax + gi];
        const int il = left[(size_t)u * gmax + gi];
        const double *row = rowS + ((size_t)u * tp2 + tp) * tau_size;
        double h = tau_pt[il + 1] - tau_pt[il];
        if (h == 0.) atomicOr(P.err, ERR_H_ZERO);
        double b = (tau - tau_pt[il]) / h;
        double a = 1 - b;
        double v;
        if (!P.cubic_time) v = a * row[il] + b * row[il + 1];
        else {
          const double *d2 = rowD + ((size_t)u * tp2 + tp) * tau_size;
          v = a * row[il] + b * row[il + 1] + ((a * a * a - a) * d2[il] + (b * b * b - b) * d2[il + 1]) * h * h / 6.0;
        }
        Sk[((size_t)u * TC + i) * tp2 + tp] = v;
      }
      __syncthreads();

      // ---- stage B: the LOS integral, one warp per k3, lanes over transfer types ----
      if (warp < nunits) {
        const int tc = max(0, min(TC, gsize[warp] - c0));
        for (int q = 0; q < npass; ++q) {
          int tt = q * 32 + lane;
          if (tt >= P.ty.tt2_size) continue;
          const int type = P.ty.tt_type[tt];
          const int lm = P.ty.tt_lm[tt];
          const int im = P.ty.tt_im[tt];
          const int mval = P.ty.tt_m[tt];
          int J1, tp1, J2 = -1, tp2b = -1, Lmax;
          if (type == 0)      { J1 = P.ty.index_J_TT; tp1 = P.ty.index_tp2_T; Lmax = P.ty.l_max_los_t; }
          else if (type == 1) { J1 = P.ty.index_J_EE; tp1 = P.ty.index_tp2_E; Lmax = P.ty.l_max_los_p;
                                if (mval != 0) { J2 = P.ty.index_J_EB; tp2b = P.ty.index_tp2_B; } }   // :1961
          else                { J1 = P.ty.index_J_EE; tp1 = P.ty.index_tp2_B; Lmax = P.ty.l_max_los_p;
                                J2 = P.ty.index_J_EB; tp2b = P.ty.index_tp2_E; }                      // :2034
          double integral1 = acc_direct[q], integral2 = acc_mixing[q];
          for (int itau = 0; itau < tc; ++itau) {
            const double a_J = aJs[warp * TC + itau];
            const int64_t row = (int64_t)ixs[warp * TC + itau] * P.proj.n_cols;
            const double *S = Sk + ((size_t)warp * TC + itau) * tp2;
            const double dt = dts[warp * TC + itau];
            double integrand = 0;
            for (int L = 0; L <= Lmax; ++L) {
              int ls = P.ty.lm_src[L * P.ty.m_size + im];
              if (ls < 0) continue;                                       // |m| > L, :2217
              if (!P.proj.block_nonzero[J1 * P.ty.L_size + L]) continue;  // all-zero (J,L) block
              double Jv = lerp_J(P.proj, P.cubic_J, row + ((int64_t)J1 * P.ty.L_size + L) * P.ty.n_transfers + lm, a_J, P.xx_step);
              integrand += Jv * S[tp1 + ls];                              // :2305
            }
            integral1 += integrand * dt;                                  // :2317
            if (J2 >= 0) {
              double integrand2 = 0;
              for (int L = 0; L <= Lmax; ++L) {
                int ls = P.ty.lm_src[L * P.ty.m_size + im];
                if (ls < 0) continue;
                if (!P.proj.block_nonzero[J2 * P.ty.L_size + L]) continue;
                double Jv = lerp_J(P.proj, P.cubic_J, row + ((int64_t)J2 * P.ty.L_size + L) * P.ty.n_transfers + lm, a_J, P.xx_step);
                integrand2 += Jv * S[tp2b + ls];
              }
              integral2 += integrand2 * dt;
            }
          }
          acc_direct[q] = integral1;
          acc_mixing[q] = integral2;
        }
      }
      __syncthreads();
    }

    // ---- epilogue: trapezoid 1/2, E/B mixing, the three divisions of transfer2.c:2092-2098 ----
    if (warp < nunits) {
      for (int q = 0; q < npass; ++q) {
        int tt = q * 32 + lane;
        if (tt >= P.ty.tt2_size) continue;
        const int type = P.ty.tt_type[tt];
        double direct = acc_direct[q] * 0.5;                              // :2329
        double mixing = acc_mixing[q] * 0.5;
        double t;
        if (type == 0) t = direct;
        else if (type == 1) t = direct + mixing;                          // :1986
        else t = direct - mixing;                                         // :2059
        t /= 4.;
        t /= (2. * P.ty.tt_l[tt] + 1);
        t /= 2.;
        if (t != t) atomicOr(P.err, ERR_NAN);                             // :2334
        P.b.out[(int64_t)tt * P.b.n_units + P.b.tr_off[p] + u0 + warp] = t;   // :2107
      }
    }
    __syncthreads();
  }
}

// launch wrappers (called from song_los_api.cu) ------------------------------------------------
cudaError_t launch_spline_k3(const Batch &b, int tau_size, int tp2_size, double *dd, double *u_scratch, int *err,
                             cudaStream_t st)
{
  int64_t cols = (int64_t)tp2_size * tau_size * b.n_pairs;
  if (cols == 0) return cudaSuccess;
  int threads = 128;
  int64_t blocks = (cols + threads - 1) / threads;
  spline_k3_kernel<<<(unsigned)blocks, threads, 0, st>>>(b, tau_size, tp2_size, dd, u_scratch, err);
  return cudaGetLastError();
}

size_t exact_smem_bytes(int tau_chunk, int tp2_size)
{
  return (size_t)EXACT_UNITS_PER_CTA * tau_chunk * ((size_t)tp2_size * 8 + 8 + 4) + 16;
}

cudaError_t launch_los_exact(const ExactParams &P, int n_tiles, cudaStream_t st)
{
  if (n_tiles == 0) return cudaSuccess;
  size_t smem = exact_smem_bytes(P.tau_chunk, P.ty.tp2_size);
  cudaError_t e = cudaFuncSetAttribute(los_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  los_exact_kernel<<<n_tiles, EXACT_THREADS, smem, st>>>(P);
  return cudaGetLastError();
}

// doubles of global scratch one CTA of the resampling kernel needs
int64_t exact_resample_scratch_doubles(int tau_size, int tp2_size, int tau_grid_max)
{
  return (int64_t)EXACT_UNITS_PER_CTA * ((int64_t)3 * tp2_size * tau_size + tau_grid_max + (tau_grid_max + 1) / 2) + 16;
}

cudaError_t launch_los_exact_resample(const ExactParams &P, int n_tiles, int n_ctas, cudaStream_t st)
{
  if (n_tiles == 0) return cudaSuccess;
  size_t smem = exact_smem_bytes(P.tau_chunk, P.ty.tp2_size) + (size_t)EXACT_UNITS_PER_CTA * P.tau_chunk * 8;
  cudaError_t e = cudaFuncSetAttribute(los_exact_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  los_exact_resample_kernel<<<n_ctas, EXACT_THREADS, smem, st>>>(P, n_tiles);
  return cudaGetLastError();
}

}  // namespace songlos
