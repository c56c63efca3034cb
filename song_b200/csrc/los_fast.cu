// los_fast.cu -- FAST mode: k3-sorted, FP64 tensor-core (DMMA m8n8k4) line-of-sight kernel.
//
// Why this shape ("FAST kernel"):
//   T_l(k3) = 1/2 sum_tau w_tau sum_L [a J_L,l(ix) + (1-a) J_L,l(ix+1)] S_L(k3,tau),
//   ix = floor(k3 (tau0-tau)/dx).  In the reference's (k1,k2)-major order neighbouring k3 values
//   hit table rows ~12 apart, so every term needs its own two J doubles: 16 B per 2 FMA, which
//   binds at the L1/L2 gather bandwidth, 10-30x below the FP64 pipe.  Sorting ALL k3 values of a
//   batch makes ~100-250 consecutive units share the same row index at a given tau; a tile of
//   units then needs ONE pair of rows per tau and the sum over (tau, L) becomes a small dense
//   product  C[unit, l] += A[unit, (L,s)] * B[(L,s), l]  with A = w {a, 1-a} S and B = {J[ix],
//   J[ix+1]} -- exactly the 8x8x4 FP64 MMA.  DMMA is the only FP64 tensor instruction on sm_100a
//   (tcgen05 has no f64 kind); measured 37.0 TFLOP/s on B200 vs 34.0 for plain DFMA.
//
// Pipeline of one song_los_compute(FAST):
//   spline_T      k3 second derivatives, written transposed  CS/CD[pair][j][tau][tp]
//   edge_rows     interpolated sources at the first/last physical k3 (flat extrapolation wings)
//   sort          cub radix sort of (k3, unit)
//   prep_units    per sorted unit: bracket, spline weights, row offsets
//   los_fast      the DMMA kernel: stage A operands per tau chunk in shared memory, B fragments
//                 straight from the dense table (L1/L2 resident), accumulators in registers
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../include/song_los.h"
#include "los_common.h"
#include "los_fast.h"

namespace songlos {

// ---------------------------------------------------------------------------------------------
// device structures
// ---------------------------------------------------------------------------------------------
constexpr int FAST_MAX_BLOCKS = 12;
constexpr int FAST_MAX_TERMS = 64;
constexpr int FAST_THREADS = 256;
constexpr int FAST_RC = 8;          // B fragments (8 columns each) per warp
constexpr int AS_PAD = 4;           // doubles of padding per A row (bank spread of the staging stores)

struct FastBlock {
  int ncols, ncols8;                // real / padded column count
  int nterms, nkc;                  // real K terms, K chunks of two terms
  int64_t jq_off;                   // offset of the block inside a table row (doubles)
  int col_off;                      // offset into col_tt / col_l
  int term_off;                     // offset into term_tp / term_sign
  int log2_lanes;                   // producer lanes per (unit,tau): 2^log2_lanes >= 2*nkc
};

struct UnitPrep {
  double k;                         // k3 (argument of the projection functions)
  double a, b, a3, b3, h6;          // spline weights: a, b, a^3-a, b^3-b, h^2/6
  int64_t off;                      // CS/CD offset of row j, or edge-row offset for wing units
  int wing;                         // 1: flat-extrapolated unit (value = edge row)
  int u;                            // original unit index (output position)
};

struct FastParams {
  int n_blocks;
  FastBlock blk[FAST_MAX_BLOCKS];
  const int *col_tt, *col_l;        // per column: transfer type index, multipole l
  const int *term_tp;               // per term: source type index
  const double *term_sign;          // per term: +1 / -1 (B-mode mixing)
  const double *Jq; int64_t jq_row; // dense tensor-core table and its row length (doubles)
  const UnitPrep *units; int64_t n_units;
  const double *CS, *CD, *E;        // transposed sources, second derivatives, edge rows
  int64_t rowlen;                   // tau_size * tp2_size
  int tau_size, tp2_size, tau_chunk, cubic_k3, nkc_max;
  const double *tau0_minus_tau, *delta_tau, *xx;
  double xx_step; int xx_size;
  double *out; int64_t out_stride;  // [tt][n_units]
  int *err;
};

// ---------------------------------------------------------------------------------------------
// dense table for the MMA:  Jq[ix][block][kc][col][4] = {A(e0), B(e0), A(e1), B(e1)},
// e0 = 2 kc, e1 = 2 kc + 1 terms of the block; A/B as in ProjTable (masked by index_xmin_J)
// ---------------------------------------------------------------------------------------------
__global__ void build_jq_kernel(const double2 *J, int64_t n_cols_dense, int xx_size,
                                int64_t jq_row, FastBlock blk, const int *term_col0, const int *col_lm,
                                double *Jq)
{
  // one thread per (kc, col) of this block, grid.y = ix
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int ix = blockIdx.y;
  if (i >= (int64_t)blk.nkc * blk.ncols8) return;
  int kc = (int)(i / blk.ncols8), c = (int)(i % blk.ncols8);
  double v[4] = {0., 0., 0., 0.};
  if (c < blk.ncols) {
    int lm = col_lm[blk.col_off + c];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int e = 2 * kc + s;
      if (e < blk.nterms) {
        int64_t dense_col = (int64_t)term_col0[blk.term_off + e] + lm;   // (J*L_size + L)*n_transfers + lm
        double2 jj = J[(int64_t)ix * n_cols_dense + dense_col];
        v[2 * s] = jj.x; v[2 * s + 1] = jj.y;
      }
    }
  }
  double *dst = Jq + (int64_t)ix * jq_row + blk.jq_off + ((int64_t)kc * blk.ncols8 + c) * 4;
  reinterpret_cast<double2 *>(dst)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2 *>(dst)[1] = make_double2(v[2], v[3]);
}

// ---------------------------------------------------------------------------------------------
// spline_T: one thread per (pair, tau, tp) column; writes CS / CD transposed [pair][j][tau][tp]
// (same algorithm as array_spline_table_columns, _SPLINE_EST_DERIV_; transfer2.c:2639)
// ---------------------------------------------------------------------------------------------
constexpr int FAST_SPLINE_MAX_NP = 256;

__global__ void __launch_bounds__(128) spline_transposed_kernel(Batch b, int tau_size, int tp2_size, int cubic,
                                                                 double *CS, double *CD, int *err)
{
  int64_t rowlen = (int64_t)tp2_size * tau_size;
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= rowlen * b.n_pairs) return;
  int p = (int)(c / rowlen);
  int64_t col = c - (int64_t)p * rowlen;           // col = tau*tp2 + tp   (tp fastest)
  int tau = (int)(col / tp2_size), tp = (int)(col % tp2_size);
  int n = b.np[p];
  const double *x = b.k3_pt + b.pt_off[p];
  const double *y = b.src + b.src_off[p] + ((int64_t)tp * tau_size + tau) * n;
  double *cs = CS + b.src_off[p] + col;
  double *cd = CD + b.src_off[p] + col;
  for (int i = 0; i < n; ++i) cs[(int64_t)i * rowlen] = y[i];
  if (!cubic) return;
  if (n > FAST_SPLINE_MAX_NP || n < 3) { atomicOr(err, ERR_H_ZERO); return; }
  double u[FAST_SPLINE_MAX_NP], dd[FAST_SPLINE_MAX_NP];
  double dy_first = ((x[2] - x[0]) * (x[2] - x[0]) * (y[1] - y[0]) - (x[1] - x[0]) * (x[1] - x[0]) * (y[2] - y[0]))
                    / ((x[2] - x[0]) * (x[1] - x[0]) * (x[2] - x[1]));
  dd[0] = -0.5;
  u[0] = (3. / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - dy_first);
  for (int i = 1; i < n - 1; ++i) {
    double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    double pp = sig * dd[i - 1] + 2.0;
    dd[i] = (sig - 1.0) / pp;
    double ui = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
    u[i] = (6.0 * ui / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / pp;
  }
  double dy_last = ((x[n - 3] - x[n - 1]) * (x[n - 3] - x[n - 1]) * (y[n - 2] - y[n - 1])
                    - (x[n - 2] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (y[n - 3] - y[n - 1]))
                   / ((x[n - 3] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (x[n - 3] - x[n - 2]));
  double un = (3. / (x[n - 1] - x[n - 2])) * (dy_last - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
  double next = (un - 0.5 * u[n - 2]) / (0.5 * dd[n - 2] + 1.0);
  cd[(int64_t)(n - 1) * rowlen] = next;
  for (int i = n - 2; i >= 0; --i) {
    next = dd[i] * next + u[i];
    cd[(int64_t)i * rowlen] = next;
  }
}

// bracket and weights of one interpolation point (transfer2.c:2683-2691)
__device__ __forceinline__ void bracket(const double *kp, int np, double ke, int &j, double &a, double &b, double &h)
{
  // j = number of interior coarse points kp[i], 1 <= i <= np-2, strictly below ke
  int lo = 1, hi = np - 1;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (kp[mid] < ke) lo = mid + 1; else hi = mid; }
  j = lo - 1;
  if (j > np - 2) j = np - 2;
  h = kp[j + 1] - kp[j];
  b = (ke - kp[j]) / h;
  a = 1 - b;
}

__device__ __forceinline__ double interp_value(const double *CS, const double *CD, int64_t off, int64_t rowlen,
                                               int64_t col, double a, double b, double a3, double b3, double h6,
                                               int cubic)
{
  double v = a * CS[off + col] + b * CS[off + rowlen + col];
  if (cubic) v += (a3 * CD[off + col] + b3 * CD[off + rowlen + col]) * h6;     // transfer2.c:2703-2705
  return v;
}

// edge_rows: E[pair][side][col], side 0 = first physical k3, 1 = last (transfer2.c:2743-2757)
__global__ void edge_rows_kernel(Batch b, int64_t rowlen, int cubic, const double *CS, const double *CD,
                                 double *E, int *err)
{
  int p = blockIdx.y;
  int side = blockIdx.z;
  const double *kp = b.k3_pt + b.pt_off[p];
  const double *kt = b.k3_tr + b.tr_off[p];
  int first = b.phys_start[p], last = first + b.phys_size[p] - 1;
  double ke = kt[side ? last : first];
  int j; double a, bb, h;
  bracket(kp, b.np[p], ke, j, a, bb, h);
  if (h == 0) { if (threadIdx.x == 0) atomicOr(err, ERR_H_ZERO); return; }
  double a3 = a * a * a - a, b3 = bb * bb * bb - bb, h6 = h * h / 6.0;
  int64_t off = b.src_off[p] + (int64_t)j * rowlen;
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < rowlen; col += (int64_t)gridDim.x * blockDim.x)
    E[((int64_t)p * 2 + side) * rowlen + col] = interp_value(CS, CD, off, rowlen, col, a, bb, a3, b3, h6, cubic);
}

__global__ void iota_kernel(int *p, int64_t n)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}

// prep_units: everything a sorted unit needs, once per unit
__global__ void prep_units_kernel(Batch b, const int *perm, const int *unit_pair, int64_t rowlen,
                                  int flat_extrap, UnitPrep *out, int *err)
{
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= b.n_units) return;
  int u = perm[s];
  int p = unit_pair[u];
  int i = u - (int)b.tr_off[p];
  const double *kp = b.k3_pt + b.pt_off[p];
  const double *kt = b.k3_tr + b.tr_off[p];
  int first = b.phys_start[p], last = first + b.phys_size[p] - 1;
  UnitPrep up;
  up.k = kt[i];
  up.u = u;
  up.wing = 0;
  if (flat_extrap && (i < first || i > last)) {
    up.wing = 1;
    up.a = up.b = up.a3 = up.b3 = up.h6 = 0;
    up.off = ((int64_t)p * 2 + (i > last ? 1 : 0)) * rowlen;
  } else {
    int j; double a, bb, h;
    bracket(kp, b.np[p], kt[i], j, a, bb, h);
    if (h == 0) atomicOr(err, ERR_H_ZERO);
    up.a = a; up.b = bb; up.a3 = a * a * a - a; up.b3 = bb * bb * bb - bb; up.h6 = h * h / 6.0;
    up.off = b.src_off[p] + (int64_t)j * rowlen;
  }
  out[s] = up;
}

// ---------------------------------------------------------------------------------------------
// the DMMA kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int WM, int WN>
__global__ void __launch_bounds__(FAST_THREADS, 2) los_fast_kernel(const __grid_constant__ FastParams P)
{
  constexpr int MT = 16 * WM;                    // units per CTA
  constexpr int AROW = MT * 4 + AS_PAD;          // doubles per (tau, kc) row of A operands
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: units[MT] | coef[TC][MT] (double2) | ixs[TC][MT] (int) | As[TC][nkc_max][AROW]
  UnitPrep *su = reinterpret_cast<UnitPrep *>(smem_raw);
  double2 *coef = reinterpret_cast<double2 *>(su + MT);
  const int TC = P.tau_chunk;
  int *ixs = reinterpret_cast<int *>(coef + (size_t)TC * MT);
  double *As = reinterpret_cast<double *>(ixs + (size_t)TC * MT);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp / WN, warp_n = warp % WN;
  const int64_t s0 = (int64_t)blockIdx.x * MT;
  const int64_t nleft = P.n_units - s0;
  const int nvalid = nleft < (int64_t)MT ? (int)nleft : MT;

  for (int n = tid; n < MT; n += FAST_THREADS) {
    UnitPrep up;
    if (n < nvalid) up = P.units[s0 + n];
    else { up = P.units[s0 + nvalid - 1]; up.u = -1; }        // tail: replicate, never stored
    su[n] = up;
  }
  __syncthreads();

  const int g0 = warp_m * 2, g1 = g0 + 1;      // the two 8-unit groups of this warp
  const int um = lane >> 2;                    // unit inside a group held by this lane (A and C rows)

  for (int ib = 0; ib < P.n_blocks; ++ib) {
    const FastBlock B = P.blk[ib];
    const int nsb = B.ncols8 >> 3;             // 8-column sub-blocks of the block
    int nb = 0;                                // sub-blocks owned by this warp: warp_n, warp_n+WN, ...
    for (int j = 0; j < FAST_RC; ++j) nb += (warp_n + j * WN < nsb) ? 1 : 0;
    // more than FAST_RC*WN sub-blocks: handled in several column passes
    for (int sb0 = 0; sb0 < nsb; sb0 += FAST_RC * WN) {
      int nbp = 0;
      for (int j = 0; j < FAST_RC; ++j) nbp += (sb0 + warp_n + j * WN < nsb) ? 1 : 0;
      double acc0[FAST_RC][2], acc1[FAST_RC][2];
#pragma unroll
      for (int j = 0; j < FAST_RC; ++j) { acc0[j][0] = acc0[j][1] = 0.; acc1[j][0] = acc1[j][1] = 0.; }

      for (int c0 = 0; c0 < P.tau_size; c0 += TC) {
        const int tc = min(TC, P.tau_size - c0);
        __syncthreads();                       // previous chunk fully consumed
        // ---- stage 1: row index and quadrature x lerp weights per (tau, unit) ----
        for (int idx = tid; idx < tc * MT; idx += FAST_THREADS) {
          int t = idx / MT, n = idx - t * MT;
          double x = su[n].k * P.tau0_minus_tau[c0 + t];               // transfer2.c:2195
          int index_x = (int)(x / P.xx_step);                          // transfer2.c:2199
          if (index_x < 0 || index_x + 1 >= P.xx_size) { atomicOr(P.err, ERR_X_OUT_OF_RANGE); index_x = 0; }
          double a_J = (P.xx[index_x + 1] - x) / P.xx_step;            // transfer2.c:2202
          double w = P.delta_tau[c0 + t];
          coef[t * MT + n] = make_double2(w * a_J, w * (1 - a_J));
          ixs[t * MT + n] = index_x;
        }
        __syncthreads();
        // ---- stage 2: A operands  As[t][kc][unit*4 + 2*(e&1) + s] = coef_s * sign * S ----
        const int nt2 = B.nkc * 2;             // term slots including the padding one
        for (int idx = tid; idx < tc * MT * nt2; idx += FAST_THREADS) {
          int e = idx % nt2;
          int r = idx / nt2;
          int t = r % tc, n = r / tc;
          double v = 0.;
          if (e < B.nterms && n < nvalid) {
            const UnitPrep &up = su[n];
            int64_t col = (int64_t)(c0 + t) * P.tp2_size + P.term_tp[B.term_off + e];
            if (up.wing) v = P.E[up.off + col];
            else v = interp_value(P.CS, P.CD, up.off, P.rowlen, col, up.a, up.b, up.a3, up.b3, up.h6, P.cubic_k3);
            v *= P.term_sign[B.term_off + e];
          }
          double2 cf = coef[t * MT + n];
          double *dst = As + ((size_t)t * B.nkc + (e >> 1)) * AROW + n * 4 + 2 * (e & 1);
          *reinterpret_cast<double2 *>(dst) = make_double2(cf.x * v, cf.y * v);
        }
        __syncthreads();

        // ---- MMA: for each tau, rows ix..ix(+1) of the table against the staged A operands ----
        if (nbp > 0) {
          for (int t = 0; t < tc; ++t) {
            const int ix0 = ixs[t * MT + g0 * 8 + um];
            const int ix1 = ixs[t * MT + g1 * 8 + um];
            const int lo = __reduce_min_sync(0xffffffffu, min(ix0, ix1));
            const int hi = __reduce_max_sync(0xffffffffu, max(ix0, ix1));
            const double *arow = As + (size_t)t * B.nkc * AROW + lane;
            for (int v = lo; v <= hi; ++v) {
              const bool m0 = (ix0 == v), m1 = (ix1 == v);
              const double *brow = P.Jq + (int64_t)v * P.jq_row + B.jq_off + lane + (int64_t)(sb0 + warp_n) * 32;
              for (int kc = 0; kc < B.nkc; ++kc) {
                double a0 = arow[(size_t)kc * AROW + g0 * 32];
                double a1 = arow[(size_t)kc * AROW + g1 * 32];
                if (!m0) a0 = 0.;
                if (!m1) a1 = 0.;
                const double *bp = brow + (int64_t)kc * B.ncols8 * 4;
                double bf[FAST_RC];
#pragma unroll
                for (int j = 0; j < FAST_RC; ++j)
                  bf[j] = (j < nbp) ? __ldg(bp + (int64_t)j * WN * 32) : 0.;
#pragma unroll
                for (int j = 0; j < FAST_RC; ++j)
                  if (j < nbp) {
                    dmma884(acc0[j][0], acc0[j][1], a0, bf[j]);
                    dmma884(acc1[j][0], acc1[j][1], a1, bf[j]);
                  }
              }
            }
          }
        }
      }

      // ---- epilogue: 1/2 of the trapezoid and the divisions of transfer2.c:2092-2098 ----
      const int cpair = 2 * (lane & 3);
#pragma unroll
      for (int j = 0; j < FAST_RC; ++j) {
        if (j >= nbp) continue;
        int cbase = (sb0 + warp_n + j * WN) * 8 + cpair;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int c = cbase + h;
          if (c >= B.ncols) continue;
          int tt = P.col_tt[B.col_off + c];
          double lfac = 2. * P.col_l[B.col_off + c] + 1;
          int u0 = su[g0 * 8 + um].u, u1 = su[g1 * 8 + um].u;
          double t0 = acc0[j][h] * 0.5; t0 /= 4.; t0 /= lfac; t0 /= 2.;
          double t1 = acc1[j][h] * 0.5; t1 /= 4.; t1 /= lfac; t1 /= 2.;
          if (u0 >= 0) { if (t0 != t0) atomicOr(P.err, ERR_NAN); P.out[(int64_t)tt * P.out_stride + u0] = t0; }
          if (u1 >= 0) { if (t1 != t1) atomicOr(P.err, ERR_NAN); P.out[(int64_t)tt * P.out_stride + u1] = t1; }
        }
      }
    }
    (void)nb;
  }
}


// ---------------------------------------------------------------------------------------------
// Warp-specialised variant: 8 consumer warps issue DMMAs, 4 producer warps stage the A operands
// of the NEXT tau chunk into the other half of a double buffer.  Hand-off through named barriers
// (FULL/EMPTY per buffer), so the tensor pipe never waits on the L2 latency of the staging loads.
// ---------------------------------------------------------------------------------------------
constexpr int WS_CONSUMERS = 256, WS_PRODUCERS = 128, WS_THREADS = WS_CONSUMERS + WS_PRODUCERS;
constexpr int BAR_FULL0 = 1, BAR_EMPTY0 = 3, BAR_PROD = 5;
constexpr int WS_UNR = 4;

__device__ __forceinline__ void bar_sync_named(int id, int count)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive_named(int id, int count)
{
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int WM, int WN>
__global__ void __launch_bounds__(WS_THREADS, 1) los_fast_ws_kernel(const __grid_constant__ FastParams P)
{
  constexpr int MT = 16 * WM;
  constexpr int LOG_MT = (MT == 128) ? 7 : (MT == 64 ? 6 : 5);
  constexpr int AROW = MT * 4 + AS_PAD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TC = P.tau_chunk;
  const int nkc_max = P.nkc_max;
  UnitPrep *su = reinterpret_cast<UnitPrep *>(smem_raw);
  double2 *coef = reinterpret_cast<double2 *>(su + MT);                 // [2][TC][MT]
  int *ixs = reinterpret_cast<int *>(coef + (size_t)2 * TC * MT);       // [2][TC][MT]
  double *As = reinterpret_cast<double *>(ixs + (size_t)2 * TC * MT);   // [2][TC][nkc_max][AROW]
  const size_t as_buf = (size_t)TC * nkc_max * AROW;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t s0 = (int64_t)blockIdx.x * MT;
  const int64_t nleft = P.n_units - s0;
  const int nvalid = nleft < (int64_t)MT ? (int)nleft : MT;

  for (int n = tid; n < MT; n += WS_THREADS) {
    UnitPrep up;
    if (n < nvalid) up = P.units[s0 + n];
    else { up = P.units[s0 + nvalid - 1]; up.u = -1; }
    su[n] = up;
  }
  __syncthreads();

  if (warp >= WS_CONSUMERS / 32) {
    // =========================== producers ===========================
    const int ptid = tid - WS_CONSUMERS, pw = warp - WS_CONSUMERS / 32;
    int item = 0;
    for (int ib = 0; ib < P.n_blocks; ++ib) {
      const FastBlock B = P.blk[ib];
      const int nsb = B.ncols8 >> 3;
      const int nt2 = B.nkc * 2;
      const int lg = B.log2_lanes;               // lanes per (unit,tau): 2^lg >= nt2
      const int gpw = 32 >> lg;                  // (unit,tau) groups per warp
      const int sub = lane >> lg, e = lane & ((1 << lg) - 1);
      const bool e_real = e < B.nterms, e_slot = e < nt2;
      const int tp = e_real ? P.term_tp[B.term_off + e] : 0;
      const double sgn = e_real ? P.term_sign[B.term_off + e] : 0.;
      for (int sb0 = 0; sb0 < nsb; sb0 += FAST_RC * WN) {
        for (int c0 = 0; c0 < P.tau_size; c0 += TC, ++item) {
          const int tc = min(TC, P.tau_size - c0);
          const int buf = item & 1;
          if (item >= 2) bar_sync_named(BAR_EMPTY0 + buf, WS_THREADS);
          double2 *cf = coef + (size_t)buf * TC * MT;
          int *ixb = ixs + (size_t)buf * TC * MT;
          double *Ab = As + (size_t)buf * as_buf;
          // ---- stage 1: row index and quadrature x lerp weights per (tau, unit) ----
          for (int idx = ptid; idx < tc * MT; idx += WS_PRODUCERS) {
            int t = idx >> LOG_MT, n = idx & (MT - 1);
            double x = su[n].k * P.tau0_minus_tau[c0 + t];                 // transfer2.c:2195
            int index_x = (int)(x / P.xx_step);                            // transfer2.c:2199
            if (index_x < 0 || index_x + 1 >= P.xx_size) { atomicOr(P.err, ERR_X_OUT_OF_RANGE); index_x = 0; }
            double a_J = (P.xx[index_x + 1] - x) / P.xx_step;              // transfer2.c:2202
            double w = P.delta_tau[c0 + t];
            cf[idx] = make_double2(w * a_J, w * (1 - a_J));
            ixb[idx] = index_x;
          }
          bar_sync_named(BAR_PROD, WS_PRODUCERS);
          // ---- stage 2: A operands, WS_UNR (unit,tau) pairs per lane group in flight ----
          const int ntot = tc * MT;
          const int stride = WS_PRODUCERS / 32 * gpw;
          for (int r0 = pw * gpw + sub; r0 < ntot; r0 += stride * WS_UNR) {
            double y0[WS_UNR], y1[WS_UNR], d0[WS_UNR], d1[WS_UNR];
#pragma unroll
            for (int u = 0; u < WS_UNR; ++u) {
              const int r = r0 + u * stride;
              y0[u] = y1[u] = d0[u] = d1[u] = 0.;
              if (r < ntot && e_real) {
                const int n = r & (MT - 1), t = r >> LOG_MT;
                if (n < nvalid) {
                  const UnitPrep &up = su[n];
                  const int64_t col = (int64_t)(c0 + t) * P.tp2_size + tp;
                  if (up.wing) y0[u] = __ldg(P.E + up.off + col);
                  else {
                    y0[u] = __ldg(P.CS + up.off + col);
                    y1[u] = __ldg(P.CS + up.off + P.rowlen + col);
                    if (P.cubic_k3) { d0[u] = __ldg(P.CD + up.off + col); d1[u] = __ldg(P.CD + up.off + P.rowlen + col); }
                  }
                }
              }
            }
#pragma unroll
            for (int u = 0; u < WS_UNR; ++u) {
              const int r = r0 + u * stride;
              if (r < ntot && e_slot) {
                const int n = r & (MT - 1), t = r >> LOG_MT;
                const UnitPrep &up = su[n];
                double v = up.wing ? y0[u] : (up.a * y0[u] + up.b * y1[u] + (up.a3 * d0[u] + up.b3 * d1[u]) * up.h6);
                v *= sgn;
                const double2 c2 = cf[r];
                double *dst = Ab + ((size_t)t * B.nkc + (e >> 1)) * AROW + n * 4 + 2 * (e & 1);
                *reinterpret_cast<double2 *>(dst) = make_double2(c2.x * v, c2.y * v);
              }
            }
          }
          bar_arrive_named(BAR_FULL0 + buf, WS_THREADS);
        }
      }
    }
    return;
  }

  // =========================== consumers ===========================
  const int warp_m = warp / WN, warp_n = warp % WN;
  const int g0 = warp_m * 2, g1 = g0 + 1;
  const int um = lane >> 2;
  int item = 0;
  for (int ib = 0; ib < P.n_blocks; ++ib) {
    const FastBlock B = P.blk[ib];
    const int nsb = B.ncols8 >> 3;
    for (int sb0 = 0; sb0 < nsb; sb0 += FAST_RC * WN) {
      int nbp = 0;
      for (int j = 0; j < FAST_RC; ++j) nbp += (sb0 + warp_n + j * WN < nsb) ? 1 : 0;
      double acc0[FAST_RC][2], acc1[FAST_RC][2];
#pragma unroll
      for (int j = 0; j < FAST_RC; ++j) { acc0[j][0] = acc0[j][1] = 0.; acc1[j][0] = acc1[j][1] = 0.; }

      for (int c0 = 0; c0 < P.tau_size; c0 += TC, ++item) {
        const int tc = min(TC, P.tau_size - c0);
        const int buf = item & 1;
        const int *ixb = ixs + (size_t)buf * TC * MT;
        const double *Ab = As + (size_t)buf * as_buf;
        bar_sync_named(BAR_FULL0 + buf, WS_THREADS);
        if (nbp > 0) {
          for (int t = 0; t < tc; ++t) {
            const int ix0 = ixb[t * MT + g0 * 8 + um];
            const int ix1 = ixb[t * MT + g1 * 8 + um];
            const int lo = __reduce_min_sync(0xffffffffu, min(ix0, ix1));
            const int hi = __reduce_max_sync(0xffffffffu, max(ix0, ix1));
            const double *arow = Ab + (size_t)t * B.nkc * AROW + lane;
            for (int v = lo; v <= hi; ++v) {
              const bool m0 = (ix0 == v), m1 = (ix1 == v);
              const bool any0 = __any_sync(0xffffffffu, m0), any1 = __any_sync(0xffffffffu, m1);
              if (!any0 && !any1) continue;
              const double *brow = P.Jq + (int64_t)v * P.jq_row + B.jq_off + lane + (int64_t)(sb0 + warp_n) * 32;
              for (int kc = 0; kc < B.nkc; ++kc) {
                const double *bp = brow + (int64_t)kc * B.ncols8 * 4;
                double bf[FAST_RC];
#pragma unroll
                for (int j = 0; j < FAST_RC; ++j)
                  bf[j] = (j < nbp) ? __ldg(bp + (int64_t)j * WN * 32) : 0.;
                if (any0) {
                  double a0 = arow[(size_t)kc * AROW + g0 * 32];
                  if (!m0) a0 = 0.;
#pragma unroll
                  for (int j = 0; j < FAST_RC; ++j)
                    if (j < nbp) dmma884(acc0[j][0], acc0[j][1], a0, bf[j]);
                }
                if (any1) {
                  double a1 = arow[(size_t)kc * AROW + g1 * 32];
                  if (!m1) a1 = 0.;
#pragma unroll
                  for (int j = 0; j < FAST_RC; ++j)
                    if (j < nbp) dmma884(acc1[j][0], acc1[j][1], a1, bf[j]);
                }
              }
            }
          }
        }
        bar_arrive_named(BAR_EMPTY0 + buf, WS_THREADS);
      }

      const int cpair = 2 * (lane & 3);
#pragma unroll
      for (int j = 0; j < FAST_RC; ++j) {
        if (j >= nbp) continue;
        int cbase = (sb0 + warp_n + j * WN) * 8 + cpair;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int c = cbase + h;
          if (c >= B.ncols) continue;
          int tt = P.col_tt[B.col_off + c];
          double lfac = 2. * P.col_l[B.col_off + c] + 1;
          int u0 = su[g0 * 8 + um].u, u1 = su[g1 * 8 + um].u;
          double t0 = acc0[j][h] * 0.5; t0 /= 4.; t0 /= lfac; t0 /= 2.;
          double t1 = acc1[j][h] * 0.5; t1 /= 4.; t1 /= lfac; t1 /= 2.;
          if (u0 >= 0) { if (t0 != t0) atomicOr(P.err, ERR_NAN); P.out[(int64_t)tt * P.out_stride + u0] = t0; }
          if (u1 >= 0) { if (t1 != t1) atomicOr(P.err, ERR_NAN); P.out[(int64_t)tt * P.out_stride + u1] = t1; }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Buf {
  T *p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct FastState {
  bool table_ready = false;
  int n_blocks = 0;
  FastBlock blk[FAST_MAX_BLOCKS];
  int64_t jq_row = 0;
  int wm = 8, wn = 1;
  Buf<double> Jq, CS, CD, E, keys_out;
  Buf<int> col_tt, col_l, col_lm, term_tp, term_col0, perm_in, perm, unit_pair;
  Buf<double> term_sign;
  Buf<UnitPrep> units;
  Buf<unsigned char> cub_tmp;
  cudaEvent_t ev[6]{};
  bool ev_ready = false;
};

FastState *fast_state_create() { return new FastState(); }

void fast_state_destroy(FastState *f)
{
  if (!f) return;
  f->Jq.release(); f->CS.release(); f->CD.release(); f->E.release(); f->keys_out.release();
  f->col_tt.release(); f->col_l.release(); f->col_lm.release(); f->term_tp.release(); f->term_col0.release();
  f->perm_in.release(); f->perm.release(); f->unit_pair.release(); f->term_sign.release();
  f->units.release(); f->cub_tmp.release();
  if (f->ev_ready) for (auto &e : f->ev) cudaEventDestroy(e);
  delete f;
}

void fast_invalidate_table(FastState *f) { if (f) f->table_ready = false; }

#define FCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  err = std::string(#call) + " failed: " + cudaGetErrorString(e_); return 1; } } while (0)

// Column blocks and K terms (the dispatch of transfer2_compute, transfer2.c:1885-2061, turned
// into data): block = (field X, index_m); its terms are the (J, L, source type, sign) products.
static int build_layout(FastState *fs, const FastHostLayout &h, const ExactParams &P, cudaStream_t st,
                        std::string &err)
{
  std::vector<int> col_tt, col_l, col_lm, term_tp, term_col0;
  std::vector<double> term_sign;
  fs->n_blocks = 0;
  int64_t jq_off = 0;
  int max_sb = 0;
  for (int type = 0; type < 3; ++type) {
    bool on = (type == 0 && h.has_T) || (type == 1 && h.has_E) || (type == 2 && h.has_B);
    if (!on) continue;
    int tt_base = type == 0 ? h.index_tt2_T : (type == 1 ? h.index_tt2_E : h.index_tt2_B);
    int Lmax = type == 0 ? h.l_max_los_t : h.l_max_los_p;
    for (int im = 0; im < h.m_size; ++im) {
      int mval = h.m[im];
      if (fs->n_blocks >= FAST_MAX_BLOCKS) { err = "too many (field,m) blocks for the FAST kernel"; return 1; }
      FastBlock B{};
      B.col_off = (int)col_tt.size();
      B.term_off = (int)term_tp.size();
      for (int il = 0; il < h.l_size; ++il) {
        int lm = h.lm_array[(size_t)il * h.m_size + im];
        if (lm < 0) continue;
        col_tt.push_back(tt_base + lm); col_l.push_back(h.l[il]); col_lm.push_back(lm);
      }
      B.ncols = (int)col_tt.size() - B.col_off;
      if (B.ncols == 0) continue;
      // segments: (J index, source base, sign)
      struct Seg { int J, tp; double sign; };
      std::vector<Seg> segs;
      if (type == 0) segs.push_back({h.index_J_TT, h.index_tp2_T, 1.});
      if (type == 1) { segs.push_back({h.index_J_EE, h.index_tp2_E, 1.});
                       if (mval != 0) segs.push_back({h.index_J_EB, h.index_tp2_B, 1.}); }     // :1961
      if (type == 2) { segs.push_back({h.index_J_EE, h.index_tp2_B, 1.});
                       segs.push_back({h.index_J_EB, h.index_tp2_E, -1.}); }                   // :2059
      for (const Seg &sg : segs) {
        if (sg.J < 0) continue;
        for (int L = 0; L <= Lmax; ++L) {
          int ls = h.lm_src[(size_t)L * h.m_size + im];
          if (ls < 0) continue;                                   // |m| > L
          if (!h.block_nonzero[(size_t)sg.J * h.L_size + L]) continue;   // spin constraint: empty (J,L)
          term_tp.push_back(sg.tp + ls);
          term_sign.push_back(sg.sign);
          term_col0.push_back((sg.J * h.L_size + L) * h.n_transfers);
        }
      }
      B.nterms = (int)term_tp.size() - B.term_off;
      if (B.nterms > FAST_MAX_TERMS) { err = "too many K terms for the FAST kernel"; return 1; }
      if (B.nterms == 0) {                      // nothing contributes: outputs are exactly zero
        term_tp.push_back(0); term_sign.push_back(0.); term_col0.push_back(0); B.nterms = 1;
      }
      B.nkc = (B.nterms + 1) / 2;
      B.log2_lanes = 0;
      while ((1 << B.log2_lanes) < 2 * B.nkc) ++B.log2_lanes;
      if (B.log2_lanes > 5) { err = "too many K terms for one warp of producers"; return 1; }
      B.ncols8 = (B.ncols + 7) / 8 * 8;
      B.jq_off = jq_off;
      jq_off += (int64_t)B.nkc * B.ncols8 * 4;
      max_sb = std::max(max_sb, B.ncols8 / 8);
      fs->blk[fs->n_blocks++] = B;
    }
  }
  fs->jq_row = jq_off;
  // warp grid: WN column groups so that each warp owns at most FAST_RC sub-blocks
  if (max_sb <= FAST_RC) { fs->wm = 8; fs->wn = 1; }
  else if (max_sb <= 2 * FAST_RC) { fs->wm = 4; fs->wn = 2; }
  else { fs->wm = 2; fs->wn = 4; }

  auto up = [&](auto &buf, const auto &vec) -> cudaError_t {
    cudaError_t e = buf.reserve(vec.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice, st);
  };
  FCK(up(fs->col_tt, col_tt)); FCK(up(fs->col_l, col_l)); FCK(up(fs->col_lm, col_lm));
  FCK(up(fs->term_tp, term_tp)); FCK(up(fs->term_col0, term_col0)); FCK(up(fs->term_sign, term_sign));
  FCK(fs->Jq.reserve((size_t)h.xx_size * fs->jq_row));
  for (int ib = 0; ib < fs->n_blocks; ++ib) {
    const FastBlock &B = fs->blk[ib];
    int64_t n = (int64_t)B.nkc * B.ncols8;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)h.xx_size);
    build_jq_kernel<<<grid, 128, 0, st>>>(P.proj.J, P.proj.n_cols, h.xx_size, fs->jq_row, B, fs->term_col0.p,
                                          fs->col_lm.p, fs->Jq.p);
    FCK(cudaGetLastError());
  }
  FCK(cudaStreamSynchronize(st));
  fs->table_ready = true;
  return 0;
}

static size_t fast_smem_bytes(int MT, int TC, int nkc_max)
{
  return (size_t)MT * sizeof(UnitPrep) + (size_t)TC * MT * (sizeof(double2) + sizeof(int))
       + (size_t)TC * nkc_max * (MT * 4 + AS_PAD) * sizeof(double);
}

template <int WM, int WN>
static cudaError_t launch_fast(const FastParams &FP, int64_t n_tiles, size_t smem, cudaStream_t st)
{
  cudaError_t e = cudaFuncSetAttribute(los_fast_kernel<WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  los_fast_kernel<WM, WN><<<(unsigned)n_tiles, FAST_THREADS, smem, st>>>(FP);
  return cudaGetLastError();
}

int fast_compute(FastState *fs, const FastHostLayout &h, const ExactParams &P, const int *h_unit_pair,
                 int device_sms, cudaStream_t st, song_los_timing *tm, std::string &err)
{
  (void)device_sms;
  if (P.cubic_J) { err = "FAST mode supports linear Bessel interpolation only; use SONG_LOS_MODE_EXACT"; return 1; }
  if (!fs->ev_ready) { for (auto &e : fs->ev) FCK(cudaEventCreate(&e)); fs->ev_ready = true; }
  if (!fs->table_ready && build_layout(fs, h, P, st, err)) return 1;
  const Batch &b = P.b;
  const int64_t U = b.n_units;
  const int64_t rowlen = (int64_t)P.tm.tau_size * P.ty.tp2_size;
  if (U >= (int64_t)1 << 31) { err = "more than 2^31 k3 units in one batch"; return 1; }

  // ---- phase 0: transposed sources + second derivatives, edge rows ----
  FCK(cudaEventRecord(fs->ev[0], st));
  int64_t n_src = h.n_src_total;
  FCK(fs->CS.reserve((size_t)n_src));
  if (P.cubic_k3) FCK(fs->CD.reserve((size_t)n_src));
  {
    int64_t cols = rowlen * b.n_pairs;
    spline_transposed_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(b, P.tm.tau_size, P.ty.tp2_size,
                                                                           P.cubic_k3, fs->CS.p, fs->CD.p, P.err);
    FCK(cudaGetLastError());
    tm->n_launches++;
  }
  if (P.flat_extrap) {
    FCK(fs->E.reserve((size_t)b.n_pairs * 2 * rowlen));
    dim3 grid((unsigned)std::min<int64_t>((rowlen + 255) / 256, 64), (unsigned)b.n_pairs, 2);
    // gridDim.y is limited to 65535 pairs per launch
    if (b.n_pairs > 65535) { err = "more than 65535 pairs in one batch"; return 1; }
    edge_rows_kernel<<<grid, 256, 0, st>>>(b, rowlen, P.cubic_k3, fs->CS.p, fs->CD.p, fs->E.p, P.err);
    FCK(cudaGetLastError());
    tm->n_launches++;
  }
  FCK(cudaEventRecord(fs->ev[1], st));

  // ---- phase 1: global k3 order ----
  FCK(fs->perm_in.reserve((size_t)U)); FCK(fs->perm.reserve((size_t)U)); FCK(fs->keys_out.reserve((size_t)U));
  FCK(fs->unit_pair.reserve((size_t)U));
  FCK(cudaMemcpyAsync(fs->unit_pair.p, h_unit_pair, (size_t)U * sizeof(int), cudaMemcpyHostToDevice, st));
  {
    iota_kernel<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(fs->perm_in.p, U);
    FCK(cudaGetLastError());
    size_t tmp_bytes = 0;
    FCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, b.k3_tr, fs->keys_out.p, fs->perm_in.p, fs->perm.p,
                                        (int)U, 0, 64, st));
    FCK(fs->cub_tmp.reserve(tmp_bytes));
    FCK(cub::DeviceRadixSort::SortPairs(fs->cub_tmp.p, tmp_bytes, b.k3_tr, fs->keys_out.p, fs->perm_in.p, fs->perm.p,
                                        (int)U, 0, 64, st));
    tm->n_launches += 4;
  }
  FCK(cudaEventRecord(fs->ev[2], st));
  FCK(fs->units.reserve((size_t)U));
  prep_units_kernel<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(b, fs->perm.p, fs->unit_pair.p, rowlen, P.flat_extrap,
                                                                fs->units.p, P.err);
  FCK(cudaGetLastError());
  tm->n_launches++;
  FCK(cudaEventRecord(fs->ev[3], st));

  // ---- phase 2: the DMMA kernel ----
  FastParams FP{};
  FP.n_blocks = fs->n_blocks;
  int nkc_max = 0;
  for (int i = 0; i < fs->n_blocks; ++i) { FP.blk[i] = fs->blk[i]; nkc_max = std::max(nkc_max, fs->blk[i].nkc); }
  FP.col_tt = fs->col_tt.p; FP.col_l = fs->col_l.p; FP.term_tp = fs->term_tp.p; FP.term_sign = fs->term_sign.p;
  FP.Jq = fs->Jq.p; FP.jq_row = fs->jq_row;
  FP.units = fs->units.p; FP.n_units = U;
  FP.CS = fs->CS.p; FP.CD = fs->CD.p; FP.E = fs->E.p; FP.rowlen = rowlen;
  FP.tau_size = P.tm.tau_size; FP.tp2_size = P.ty.tp2_size; FP.cubic_k3 = P.cubic_k3;
  FP.tau0_minus_tau = P.tm.tau0_minus_tau; FP.delta_tau = P.tm.delta_tau; FP.xx = P.xx;
  FP.xx_step = P.xx_step; FP.xx_size = P.xx_size;
  FP.out = b.out; FP.out_stride = U; FP.err = P.err;

  const int MT = 16 * fs->wm;
  FP.nkc_max = nkc_max;
  int64_t n_tiles = (U + MT - 1) / MT;
  static const int variant = []() { const char *v = getenv("SONG_LOS_FAST_VARIANT"); return v ? atoi(v) : 1; }();
  if (variant == 1) {
    // warp-specialised: one CTA per SM, double-buffered operands, as many tau samples per chunk as fit
    auto ws_smem = [&](int TC) {
      return (size_t)MT * sizeof(UnitPrep) + (size_t)2 * TC * MT * (sizeof(double2) + sizeof(int))
           + (size_t)2 * TC * nkc_max * (MT * 4 + AS_PAD) * sizeof(double);
    };
    int TC = 1;
    while (TC < P.tm.tau_size && TC < 8 && ws_smem(TC + 1) <= 200 * 1024) ++TC;
    FP.tau_chunk = TC;
    size_t smem = ws_smem(TC);
    if (smem > 225 * 1024) { err = "FAST kernel: shared-memory stage does not fit"; return 1; }
    if (n_tiles > 0) {
      cudaError_t e;
      if (fs->wm == 8) { e = cudaFuncSetAttribute(los_fast_ws_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                         if (e == cudaSuccess) { los_fast_ws_kernel<8, 1><<<(unsigned)n_tiles, WS_THREADS, smem, st>>>(FP); e = cudaGetLastError(); } }
      else if (fs->wm == 4) { e = cudaFuncSetAttribute(los_fast_ws_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                         if (e == cudaSuccess) { los_fast_ws_kernel<4, 2><<<(unsigned)n_tiles, WS_THREADS, smem, st>>>(FP); e = cudaGetLastError(); } }
      else { e = cudaFuncSetAttribute(los_fast_ws_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                         if (e == cudaSuccess) { los_fast_ws_kernel<2, 4><<<(unsigned)n_tiles, WS_THREADS, smem, st>>>(FP); e = cudaGetLastError(); } }
      FCK(e);
      tm->n_launches++;
    }
  } else {
    // single-role variant (kept for A/B measurements): two CTAs per SM overlap staging and MMA
    int TC = 1;
    while (TC < P.tm.tau_size && fast_smem_bytes(MT, TC + 1, nkc_max) <= 100 * 1024) ++TC;
    FP.tau_chunk = TC;
    size_t smem = fast_smem_bytes(MT, TC, nkc_max);
    if (smem > 220 * 1024) { err = "FAST kernel: shared-memory stage does not fit"; return 1; }
    if (n_tiles > 0) {
      if (fs->wm == 8) FCK((launch_fast<8, 1>(FP, n_tiles, smem, st)));
      else if (fs->wm == 4) FCK((launch_fast<4, 2>(FP, n_tiles, smem, st)));
      else FCK((launch_fast<2, 4>(FP, n_tiles, smem, st)));
      tm->n_launches++;
    }
  }
  FCK(cudaEventRecord(fs->ev[4], st));
  FCK(cudaEventSynchronize(fs->ev[4]));
  float ms;
  FCK(cudaEventElapsedTime(&ms, fs->ev[0], fs->ev[1])); tm->ms_spline = ms;
  FCK(cudaEventElapsedTime(&ms, fs->ev[1], fs->ev[2])); tm->ms_sort = ms;
  FCK(cudaEventElapsedTime(&ms, fs->ev[2], fs->ev[3])); tm->ms_interp = ms;
  FCK(cudaEventElapsedTime(&ms, fs->ev[3], fs->ev[4])); tm->ms_integrate = ms;
  FCK(cudaEventElapsedTime(&ms, fs->ev[0], fs->ev[4])); tm->ms_total = ms;
  return 0;
}

}  // namespace songlos
