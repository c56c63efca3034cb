// los_tc.cu -- FAST mode: k3-sorted FP64 tensor-core (DMMA m8n8k4) line-of-sight kernel with a
// TMA-bulk / mbarrier operand pipeline.
//
// The sum being evaluated (reference transfer2_integrate, source/transfer2.c:2140-2341, and the
// field dispatch of transfer2_compute, :1861-2116):
//   T_l(k3) = 1/2 sum_tau w_tau sum_L [a J_L,l(ix) + (1-a) J_L,l(ix+1)] S_L(k3,tau),
//   ix = floor(k3 (tau0-tau)/dx),  a = (xx[ix+1] - k3 (tau0-tau))/dx.
// In the reference's (k1,k2)-major order neighbouring k3 values hit table rows ~12 apart, so every
// term needs its own two J doubles.  Sorting ALL k3 values of a batch makes ~100-250 consecutive
// units share the same row at a given tau; a tile of units then needs 2-3 table rows per tau and the
// sum over (tau, L) becomes a small dense product
//        C[unit, l] += A[unit, (L,s)] * B_ix[(L,s), l],   A = w {a, 1-a} S,  B_ix = {J[ix], J[ix+1]}
// which is the 8x8x4 FP64 MMA (the only FP64 tensor instruction of sm_100a; tcgen05 has no f64 kind).
//
// Pipeline of one song_los_compute(FAST):
//   spline_csd    k3 second derivatives, written per interval: CSD[pair][j][tau][tp] = {y_j, dd_j, y_j+1, dd_j+1}
//                 (one 256-bit load per (unit, tau, source type) in the kernel)
//   edge_rows     interpolated sources at the first/last physical k3 (flat-extrapolation wings)
//   sort          cub radix sort of (k3, unit)
//   prep_units    per sorted unit: bracket, spline weights, source row pointer
//   los_tc        per CTA one tile of MT sorted units.  4 producer warps: one thread issues
//                 cp.async.bulk (TMA 1-D) copies of the needed table rows into a shared-memory ring,
//                 all of them evaluate the k3 spline and the quadrature/lerp weights and write the A
//                 operands into the same ring stage; 8 consumer warps wait on the stage's mbarrier,
//                 read conflict-free 256-byte fragments and issue DMMAs into register accumulators.
//                 Columns whose J is still below the reference's cut at this row (x << l) are skipped.
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
constexpr int TC_MAX_BLOCKS = 12;
constexpr int TC_MAX_PASSES = 24;
constexpr int TC_MAX_TERMS = 64;          // K terms of one block (2 per kc)
constexpr int TC_KCS_MAX = 6;             // kc's (pairs of terms) per pipeline stage
constexpr int TC_CONSUMERS = 256, TC_PRODUCERS = 128, TC_THREADS = TC_CONSUMERS + TC_PRODUCERS;
constexpr int TC_APAD = 4;                // doubles of padding per A row

struct TcBlock {
  int ncols, nsb;                         // real columns, 8-column sub-blocks
  int nterms, nkc;                        // real K terms, pairs of terms
  int col_off, term_off;                  // offsets into col_* / term_* tables
};

struct TcPass {                           // one column pass of a block = one accumulator lifetime
  int blk, sb0, nsb;                      // block, first sub-block, sub-blocks in the pass
  int nq, kcs;                            // K chunks and kc's per chunk (last chunk may be shorter)
  int64_t jq_off;                         // offset of the pass inside a table row (doubles)
};

struct TcUnit {
  double k;                               // k3 (argument of the projection functions)
  double a, b, a3h, b3h;                  // spline weights: a, b, (a^3-a) h^2/6, (b^3-b) h^2/6
  const double4 *base;                    // interval j of the pair at (tau=0, tp=0); wing units: edge row
  int u;                                  // original unit index (output position), -1 = padding
  int pad;
};

struct TcParams {
  int n_blocks, n_pass;
  TcBlock blk[TC_MAX_BLOCKS];
  TcPass pass[TC_MAX_PASSES];
  int term_tp[TC_MAX_TERMS * 2];          // per term: source type index          (indexed term_off + e)
  double term_sign[TC_MAX_TERMS * 2];     // per term: +1 / -1 (B-mode mixing)
  const int *col_tt, *col_l;              // per column: transfer type index, multipole l
  const double *Jq; int64_t jq_row;       // tensor-core table and its row length (doubles)
  const int *nact;                        // [n_blocks][xx_size] active sub-block prefix per row
  const TcUnit *units; int64_t n_units;
  int64_t rowlen;                         // tau_size * tp2_size
  int tau_size, tp2_size;
  const double *tau0_minus_tau, *delta_tau, *xx;
  double inv_xx_step; int xx_size;
  int n_stages, rmax;                     // ring depth, table rows per stage
  int hdr_bytes, a_bytes, b_row_bytes, stage_bytes;  // shared-memory geometry of one stage
  double *out; int64_t out_stride;        // [tt][n_units]
  int *err;
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk copy (TMA)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "TC_WAIT:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra TC_DONE;\n"
      " bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ double4 ldg256(const double4 *p)
{
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// table for the MMA.  Row ix is laid out in the order the kernel consumes it:
//   [pass][chunk q][sub-block sb][kc in chunk][col in sub-block (8)][4] = {A(e0), B(e0), A(e1), B(e1)}
// with e0 = 2 kc, e1 = 2 kc + 1 the two terms of kc and {A, B} = {J[ix], J[ix+1]} masked by
// index_xmin_J exactly like the dense table (skip test of transfer2.c:2232).  A warp's B fragment of
// (sb, kc) is 32 consecutive doubles; the active sub-block prefix of a (row, pass, chunk) is one
// contiguous piece, i.e. one bulk copy.
// ---------------------------------------------------------------------------------------------
__global__ void tc_build_table_kernel(const double2 *J, int64_t n_cols_dense, int64_t jq_row, TcPass ps, TcBlock blk,
                                      const int *term_col0, const int *col_lm, double *Jq)
{
  // one thread per (kc, column of the pass), grid.y = ix
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int ix = blockIdx.y;
  int ncolp = ps.nsb * 8;
  if (i >= (int64_t)blk.nkc * ncolp) return;
  int kc = (int)(i / ncolp), cp = (int)(i % ncolp);
  int c = ps.sb0 * 8 + cp;                         // column inside the block
  double v[4] = {0., 0., 0., 0.};
  if (c < blk.ncols) {
    int lm = col_lm[blk.col_off + c];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int e = 2 * kc + s;
      if (e < blk.nterms) {
        double2 jj = J[(int64_t)ix * n_cols_dense + term_col0[blk.term_off + e] + lm];
        v[2 * s] = jj.x; v[2 * s + 1] = jj.y;
      }
    }
  }
  int q = kc / ps.kcs, kcq = kc - q * ps.kcs;
  int kcs_q = min(ps.kcs, blk.nkc - q * ps.kcs);
  int sb = cp >> 3, c8 = cp & 7;
  double *dst = Jq + (int64_t)ix * jq_row + ps.jq_off + (int64_t)ps.nsb * (q * ps.kcs) * 32
              + ((int64_t)sb * kcs_q + kcq) * 32 + c8 * 4;
  reinterpret_cast<double2 *>(dst)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2 *>(dst)[1] = make_double2(v[2], v[3]);
}

// nact[blk][ix] = 1 + last sub-block of the block with a non-zero entry in row ix
__global__ void tc_nact_kernel(const double2 *J, int64_t n_cols_dense, int xx_size, TcBlock blk, int ib,
                               const int *term_col0, const int *col_lm, int *nact)
{
  int ix = blockIdx.x;
  int best = 0;
  for (int c = threadIdx.x; c < blk.ncols; c += blockDim.x) {
    int lm = col_lm[blk.col_off + c];
    bool on = false;
    for (int e = 0; e < blk.nterms && !on; ++e) {
      double2 jj = J[(int64_t)ix * n_cols_dense + term_col0[blk.term_off + e] + lm];
      on = (jj.x != 0.) || (jj.y != 0.);
    }
    if (on) best = max(best, c / 8 + 1);
  }
  best = __reduce_max_sync(0xffffffffu, best);
  if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(&nact[(int64_t)ib * xx_size + ix], best);
}

// running maximum over ix: the prefix used for a tile must cover every row the tile touches
__global__ void tc_nact_cummax_kernel(int *nact, int xx_size, int n_blocks)
{
  int ib = blockIdx.x * blockDim.x + threadIdx.x;
  if (ib >= n_blocks) return;
  int m = 0;
  for (int ix = 0; ix < xx_size; ++ix) {
    m = max(m, nact[(int64_t)ib * xx_size + ix]);
    nact[(int64_t)ib * xx_size + ix] = m;
  }
}

// ---------------------------------------------------------------------------------------------
// spline_csd: one thread per (pair, tau, tp) column; writes, for every k3 interval j of the pair,
// CSD[pair][j][tau][tp] = {y_j, dd_j, y_j+1, dd_j+1}   (pair p starts at entry src_off[p] - p*rowlen)
// (same algorithm as array_spline_table_columns, _SPLINE_EST_DERIV_; transfer2.c:2639)
// ---------------------------------------------------------------------------------------------
constexpr int TC_SPLINE_MAX_NP = 256;

__global__ void __launch_bounds__(128) tc_spline_csd_kernel(Batch b, int tau_size, int tp2_size, int cubic,
                                                            double4 *CSD, int *err)
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
  double4 *cs = CSD + (b.src_off[p] - (int64_t)p * rowlen) + col;
  if (!cubic) {
    for (int i = 0; i + 1 < n; ++i) cs[(int64_t)i * rowlen] = make_double4(y[i], 0., y[i + 1], 0.);
    return;
  }
  if (n > TC_SPLINE_MAX_NP || n < 3) { atomicOr(err, ERR_H_ZERO); return; }
  double u[TC_SPLINE_MAX_NP], dd[TC_SPLINE_MAX_NP];
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
  double next = (un - 0.5 * u[n - 2]) / (0.5 * dd[n - 2] + 1.0);     // dd of point i+1
  for (int i = n - 2; i >= 0; --i) {
    double cur = dd[i] * next + u[i];                                  // dd of point i
    cs[(int64_t)i * rowlen] = make_double4(y[i], cur, y[i + 1], next);
    next = cur;
  }
}

// bracket and weights of one interpolation point (transfer2.c:2683-2691)
__device__ __forceinline__ void tc_bracket(const double *kp, int np, double ke, int &j, double &a, double &b, double &h)
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

// edge rows: E[pair][side][col] = {S(k3 edge), 0, 0, 0}; side 0 = first physical k3, 1 = last
// (flat extrapolation, transfer2.c:2743-2757).  Wing units read them with weights a=1, b=a3h=b3h=0.
__global__ void tc_edge_rows_kernel(Batch b, int64_t rowlen, const double4 *CSD, double4 *E, int *err)
{
  int p = blockIdx.y;
  int side = blockIdx.z;
  const double *kp = b.k3_pt + b.pt_off[p];
  const double *kt = b.k3_tr + b.tr_off[p];
  int first = b.phys_start[p], last = first + b.phys_size[p] - 1;
  double ke = kt[side ? last : first];
  int j; double a, bb, h;
  tc_bracket(kp, b.np[p], ke, j, a, bb, h);
  if (h == 0) { if (threadIdx.x == 0) atomicOr(err, ERR_H_ZERO); return; }
  double a3h = (a * a * a - a) * (h * h / 6.0), b3h = (bb * bb * bb - bb) * (h * h / 6.0);
  const double4 *r0 = CSD + (b.src_off[p] - (int64_t)p * rowlen) + (int64_t)j * rowlen;
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < rowlen; col += (int64_t)gridDim.x * blockDim.x) {
    double4 s4 = r0[col];
    E[((int64_t)p * 2 + side) * rowlen + col] = make_double4(a * s4.x + bb * s4.z + a3h * s4.y + b3h * s4.w, 0., 0., 0.);
  }
}

__global__ void tc_iota_kernel(int *p, int64_t n)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}

// prep_units: everything a sorted unit needs, once per unit
__global__ void tc_prep_units_kernel(Batch b, const int *perm, const int *unit_pair, int64_t rowlen,
                                     int flat_extrap, const double4 *CSD, const double4 *E, TcUnit *out, int *err)
{
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= b.n_units) return;
  int u = perm[s];
  int p = unit_pair[u];
  int i = u - (int)b.tr_off[p];
  const double *kp = b.k3_pt + b.pt_off[p];
  const double *kt = b.k3_tr + b.tr_off[p];
  int first = b.phys_start[p], last = first + b.phys_size[p] - 1;
  TcUnit up;
  up.k = kt[i];
  up.u = u;
  up.pad = 0;
  if (flat_extrap && (i < first || i > last)) {
    up.a = 1.; up.b = 0.; up.a3h = 0.; up.b3h = 0.;
    up.base = E + ((int64_t)p * 2 + (i > last ? 1 : 0)) * rowlen;
  } else {
    int j; double a, bb, h;
    tc_bracket(kp, b.np[p], kt[i], j, a, bb, h);
    if (h == 0) atomicOr(err, ERR_H_ZERO);
    double h6 = h * h / 6.0;
    up.a = a; up.b = bb; up.a3h = (a * a * a - a) * h6; up.b3h = (bb * bb * bb - bb) * h6;
    up.base = CSD + (b.src_off[p] - (int64_t)p * rowlen) + (int64_t)j * rowlen;
  }
  out[s] = up;
}

// ---------------------------------------------------------------------------------------------
// the DMMA kernel
// ---------------------------------------------------------------------------------------------
// shared memory:  [bars: full[NS], empty[NS]] [unit ids: MT ints] [stage 0] [stage 1] ...
// stage:          header {meta int4 {row_lo, nrows, nact, row_hi}, ixs[MT]} | A[KCS][AROW] | B[rmax][b_row]
//                 (each part padded to 128 bytes: bulk-copy destinations must be 16-byte aligned)
struct TcStagePtrs {
  int *meta; int *ixs; double *A; double *B;
};

template <int MT>
__device__ __forceinline__ TcStagePtrs tc_stage(unsigned char *stage0, int s, const TcParams &P)
{
  unsigned char *base = stage0 + (size_t)s * P.stage_bytes;
  TcStagePtrs r;
  r.meta = reinterpret_cast<int *>(base);
  r.ixs = reinterpret_cast<int *>(base + 16);
  r.A = reinterpret_cast<double *>(base + P.hdr_bytes);
  r.B = reinterpret_cast<double *>(base + P.hdr_bytes + P.a_bytes);
  return r;
}

// ---- consumer inner loops -------------------------------------------------------------------
// One table row against the staged A operands of a chunk.  NACT (active sub-blocks of this warp) is
// a template parameter so that the body is straight-line code: per kc 1-2 A fragments and NACT B
// fragments (LDS.64, 256 contiguous bytes per warp) feed 1-2 x NACT DMMAs; the fragments of kc+1 are
// loaded into the other register set before the DMMAs of kc are issued (ping-pong, no moves).

// both 8-unit groups of the warp have units on this row
template <int NACT, int NW>
__device__ __forceinline__ void tc_mma_both(double (&acc0)[NW][2], double (&acc1)[NW][2],
                                            const double *a0p, const double *a1p, int arow, bool m0, bool m1,
                                            const double *bp, int kcs, int b_sb_stride)
{
  double pa0, pa1, pb[NACT], qa0, qa1, qb[NACT];
#define TC_LOAD(A0, A1, BF, KC) do { \
    A0 = a0p[(size_t)(KC) * arow]; A1 = a1p[(size_t)(KC) * arow]; \
    _Pragma("unroll") for (int j = 0; j < NACT; ++j) BF[j] = bp[(size_t)j * b_sb_stride + (KC) * 32]; } while (0)
#define TC_MMA(A0, A1, BF) do { \
    const double x0 = m0 ? A0 : 0., x1 = m1 ? A1 : 0.; \
    _Pragma("unroll") for (int j = 0; j < NACT; ++j) dmma884(acc0[j][0], acc0[j][1], x0, BF[j]); \
    _Pragma("unroll") for (int j = 0; j < NACT; ++j) dmma884(acc1[j][0], acc1[j][1], x1, BF[j]); } while (0)
  TC_LOAD(pa0, pa1, pb, 0);
  int kc = 0;
  for (; kc + 2 <= kcs; kc += 2) {
    TC_LOAD(qa0, qa1, qb, kc + 1);
    TC_MMA(pa0, pa1, pb);
    if (kc + 2 < kcs) TC_LOAD(pa0, pa1, pb, kc + 2);
    TC_MMA(qa0, qa1, qb);
  }
  if (kc < kcs) TC_MMA(pa0, pa1, pb);
#undef TC_LOAD
#undef TC_MMA
}

// only one of the two groups has units on this row
template <int NACT, int NW>
__device__ __forceinline__ void tc_mma_one(double (&acc)[NW][2], const double *ap, int arow, bool m,
                                           const double *bp, int kcs, int b_sb_stride)
{
  double pa, pb[NACT], qa, qb[NACT];
#define TC_LOAD(A0, BF, KC) do { \
    A0 = ap[(size_t)(KC) * arow]; \
    _Pragma("unroll") for (int j = 0; j < NACT; ++j) BF[j] = bp[(size_t)j * b_sb_stride + (KC) * 32]; } while (0)
#define TC_MMA(A0, BF) do { \
    const double x = m ? A0 : 0.; \
    _Pragma("unroll") for (int j = 0; j < NACT; ++j) dmma884(acc[j][0], acc[j][1], x, BF[j]); } while (0)
  TC_LOAD(pa, pb, 0);
  int kc = 0;
  for (; kc + 2 <= kcs; kc += 2) {
    TC_LOAD(qa, qb, kc + 1);
    TC_MMA(pa, pb);
    if (kc + 2 < kcs) TC_LOAD(pa, pb, kc + 2);
    TC_MMA(qa, qb);
  }
  if (kc < kcs) TC_MMA(pa, pb);
#undef TC_LOAD
#undef TC_MMA
}

template <int NACT, int NW>
__device__ __forceinline__ void tc_mma_row_smem(double (&acc0)[NW][2], double (&acc1)[NW][2],
                                                const double *a0p, const double *a1p, int arow,
                                                bool m0, bool m1, bool any0, bool any1,
                                                const double *bp, int kcs, int b_sb_stride)
{
  if (any0 && any1) tc_mma_both<NACT, NW>(acc0, acc1, a0p, a1p, arow, m0, m1, bp, kcs, b_sb_stride);
  else if (any0) tc_mma_one<NACT, NW>(acc0, a0p, arow, m0, bp, kcs, b_sb_stride);
  else tc_mma_one<NACT, NW>(acc1, a1p, arow, m1, bp, kcs, b_sb_stride);
}

// rows outside the staged window (sparse k3 regions): B fragments straight from the table in global
// memory; generic in the number of active sub-blocks (rare path, kept small)
template <int NW>
__device__ __forceinline__ void tc_mma_row_global(double (&acc0)[NW][2], double (&acc1)[NW][2],
                                               const double *a0p, const double *a1p, int arow,
                                               bool m0, bool m1, const double *bp, int kcs, int b_sb_stride, int nactw)
{
  for (int kc = 0; kc < kcs; ++kc) {
    double a0 = a0p[(size_t)kc * arow], a1 = a1p[(size_t)kc * arow];
    if (!m0) a0 = 0.;
    if (!m1) a1 = 0.;
#pragma unroll
    for (int j = 0; j < NW; ++j)
      if (j < nactw) {
        const double bf = __ldg(bp + (size_t)j * b_sb_stride + kc * 32);
        dmma884(acc0[j][0], acc0[j][1], a0, bf);
        dmma884(acc1[j][0], acc1[j][1], a1, bf);
      }
  }
}

template <int WM, int WN, int NW>
__global__ void __launch_bounds__(TC_THREADS, 1) los_tc_kernel(const __grid_constant__ TcParams P)
{
  constexpr int MT = 16 * WM;                    // units per CTA
  constexpr int AROW = MT * 4 + TC_APAD;         // doubles per kc row of A operands
  constexpr int TPU = TC_PRODUCERS / MT;         // producer threads per unit (1, 2 or 4)
  static_assert(WM * WN == TC_CONSUMERS / 32, "8 consumer warps");
  static_assert(TPU >= 1 && TPU * MT == TC_PRODUCERS, "producer mapping");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NS = P.n_stages;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);            // full[NS], empty[NS]
  int *unit_id = reinterpret_cast<int *>(smem_raw + 16 * 8);           // up to 8 stages of barriers
  unsigned char *stage0 = smem_raw + 128 + MT * 4;                     // 128-byte aligned for MT % 32 == 0

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t s0 = (int64_t)blockIdx.x * MT;
  const int64_t nleft = P.n_units - s0;
  const int nvalid = nleft < (int64_t)MT ? (int)nleft : MT;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&bars[s]), TC_PRODUCERS + 1);          // producers' A writes + the copy issuer
      mbar_init(smem_u32(&bars[8 + s]), TC_CONSUMERS / 32);     // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int n = tid; n < MT; n += TC_THREADS) unit_id[n] = (n < nvalid) ? P.units[s0 + n].u : -1;
  __syncthreads();

  if (warp >= TC_CONSUMERS / 32) {
    // =========================== producers ===========================
    // thread <-> (unit n, share `sub` of the kc's of a chunk); stage sequence: pass, chunk q, tau
    const int ptid = tid - TC_CONSUMERS;
    const int n = ptid % MT, sub = ptid / MT;
    constexpr int NKC_T = (TC_KCS_MAX + TPU - 1) / TPU;          // kc's per thread per stage
    const TcUnit up = P.units[s0 + (n < nvalid ? n : nvalid - 1)];   // tail: replicate, never stored
    const double k_first = P.units[s0].k, k_last = P.units[s0 + nvalid - 1].k;
    int it = 0;
    for (int ip = 0; ip < P.n_pass; ++ip) {
      const TcPass ps = P.pass[ip];
      const TcBlock B = P.blk[ps.blk];
      const int *nact_b = P.nact + (int64_t)ps.blk * P.xx_size;
      for (int q = 0; q < ps.nq; ++q) {
        const int kc0 = q * ps.kcs;
        const int kcs_q = min(ps.kcs, B.nkc - kc0);
        // source types and signs of this thread's terms, hoisted out of the tau loop; slots past
        // the last term read term 0 with weight 0 (unconditional loads, no predication)
        int tpo[NKC_T * 2];
        unsigned real_mask = 0, neg_mask = 0;                    // bit per slot: term exists / enters with sign -1
        int nmine = 0;
#pragma unroll
        for (int i = 0; i < NKC_T; ++i) {
          const int kcq = sub + i * TPU;
          if (kcq < kcs_q) nmine = i + 1;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int e = 2 * (kc0 + kcq) + h;
            const bool real = (kcq < kcs_q) && (e < B.nterms);
            tpo[2 * i + h] = real ? P.term_tp[B.term_off + e] : P.term_tp[B.term_off];
            if (real) { real_mask |= 1u << (2 * i + h); if (P.term_sign[B.term_off + e] < 0.) neg_mask |= 1u << (2 * i + h); }
          }
        }
        const uint32_t bytes_per_sb = (uint32_t)kcs_q * 256u;
        const int64_t q_off = ps.jq_off + (int64_t)ps.nsb * kc0 * 32;
        for (int t = 0; t < P.tau_size; ++t, ++it) {
          const int s = it % NS;
          const uint32_t full = smem_u32(&bars[s]), empty = smem_u32(&bars[8 + s]);
          const TcStagePtrs st = tc_stage<MT>(stage0, s, P);
          const double d = __ldg(P.tau0_minus_tau + t), w = __ldg(P.delta_tau + t);
          // ---- source loads first: {y_j, dd_j, y_j+1, dd_j+1} of every term of this thread ----
          const double4 *src_t = up.base + (int64_t)t * P.tp2_size;
          double4 y[NKC_T * 2];
#pragma unroll
          for (int i = 0; i < NKC_T; ++i)
            if (i < nmine) { y[2 * i] = ldg256(src_t + tpo[2 * i]); y[2 * i + 1] = ldg256(src_t + tpo[2 * i + 1]); }
          // ---- row index and quadrature x lerp weights of this unit (transfer2.c:2195-2202) ----
          const double x = up.k * d;
          int ix = (int)(x * P.inv_xx_step);
          if (ix < 0 || ix + 1 >= P.xx_size) { atomicOr(P.err, ERR_X_OUT_OF_RANGE); ix = 0; }
          const double a_J = (__ldg(P.xx + ix + 1) - x) * P.inv_xx_step;
          const double c0 = w * a_J, c1 = w - c0;
          int row_lo = 0, row_hi = 0, nact = 0;
          if (ptid == 0) {
            // table rows of the tile at this tau: units are k3-sorted, so rows are [row(first), row(last)]
            row_lo = (int)(k_first * d * P.inv_xx_step); row_hi = (int)(k_last * d * P.inv_xx_step);
            row_lo = max(0, min(row_lo, P.xx_size - 2)); row_hi = max(row_lo, min(row_hi, P.xx_size - 2));
            nact = max(0, min(__ldg(nact_b + row_hi) - ps.sb0, ps.nsb));
          }
          if (it >= NS) mbar_wait(empty, ((it / NS) - 1) & 1);
          if (ptid == 0) {
            const int nrows = min(row_hi - row_lo + 1, P.rmax);
            const uint32_t bytes_row = (uint32_t)nact * bytes_per_sb;
            st.meta[0] = row_lo; st.meta[1] = nrows; st.meta[2] = nact; st.meta[3] = row_hi;
            if (bytes_row > 0) {
              mbar_arrive_expect_tx(full, bytes_row * nrows);
              const double *src = P.Jq + (int64_t)row_lo * P.jq_row + q_off;
              for (int r = 0; r < nrows; ++r)
                bulk_g2s(smem_u32(st.B) + r * P.b_row_bytes, src + (int64_t)r * P.jq_row, bytes_row, full);
            } else {
              mbar_arrive(full);
            }
          }
          if (sub == 0) st.ixs[n] = ix;
          // ---- A operands: v = S(k3, tau, term) by the k3 spline (transfer2.c:2697-2705), stored
          //      as {c0 v, c1 v} per term, two terms per kc ----
#pragma unroll
          for (int i = 0; i < NKC_T; ++i) {
            if (i < nmine) {
              double v[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const double4 yy = y[2 * i + h];
                double val = up.a * yy.x + up.b * yy.z + up.a3h * yy.y + up.b3h * yy.w;
                if (neg_mask >> (2 * i + h) & 1u) val = -val;
                v[h] = (real_mask >> (2 * i + h) & 1u) ? val : 0.;
              }
              double *dst = st.A + (size_t)(sub + i * TPU) * AROW + n * 4;
              reinterpret_cast<double2 *>(dst)[0] = make_double2(c0 * v[0], c1 * v[0]);
              reinterpret_cast<double2 *>(dst)[1] = make_double2(c0 * v[1], c1 * v[1]);
            }
          }
          mbar_arrive(full);
        }
      }
    }
    return;
  }

  // =========================== consumers ===========================
  const int warp_m = warp / WN, warp_n = (warp + warp_m) % WN;      // rotate column groups over the SM sub-partitions
  const int g0 = warp_m * 2, g1 = g0 + 1;      // the two 8-unit groups of this warp
  const int um = lane >> 2;                    // unit inside a group held by this lane (A and C rows)
  int it = 0;
  for (int ip = 0; ip < P.n_pass; ++ip) {
    const TcPass ps = P.pass[ip];
    const TcBlock B = P.blk[ps.blk];
    double acc0[NW][2], acc1[NW][2];
#pragma unroll
    for (int j = 0; j < NW; ++j) { acc0[j][0] = acc0[j][1] = 0.; acc1[j][0] = acc1[j][1] = 0.; }

    for (int q = 0; q < ps.nq; ++q) {
      const int kc0 = q * ps.kcs;
      const int kcs_q = min(ps.kcs, B.nkc - kc0);
      for (int t = 0; t < P.tau_size; ++t, ++it) {
        const int s = it % NS;
        const TcStagePtrs st = tc_stage<MT>(stage0, s, P);
        mbar_wait(smem_u32(&bars[s]), (it / NS) & 1);
        const int row_lo = st.meta[0], nrows = st.meta[1], nact = st.meta[2];
        // owned sub-blocks: warp_n, warp_n + WN, ... ; active ones form a prefix of that list
        const int nactw = (nact > warp_n) ? min(NW, (nact - warp_n + WN - 1) / WN) : 0;
        if (nactw > 0) {
          const int ix0 = st.ixs[g0 * 8 + um];
          const int ix1 = st.ixs[g1 * 8 + um];
          const double *a0p = st.A + g0 * 32 + lane;
          const double *a1p = st.A + g1 * 32 + lane;
          const int b_sb_stride = WN * kcs_q * 32;
          int v = __reduce_min_sync(0xffffffffu, min(ix0, ix1));
          while (v != 0x7fffffff) {
            const bool m0 = (ix0 == v), m1 = (ix1 == v);
            const bool any0 = __any_sync(0xffffffffu, m0), any1 = __any_sync(0xffffffffu, m1);
            const int rb = v - row_lo;
            if (rb < nrows) {
              const double *bp = st.B + (size_t)rb * (P.b_row_bytes / 8) + (size_t)warp_n * kcs_q * 32 + lane;
#define TC_CASE(N) case N: if constexpr (N <= NW) tc_mma_row_smem<N, NW>(acc0, acc1, a0p, a1p, AROW, m0, m1, any0, any1, bp, kcs_q, b_sb_stride); break;
              switch (nactw) {
                TC_CASE(1) TC_CASE(2) TC_CASE(3) TC_CASE(4) TC_CASE(5) TC_CASE(6) TC_CASE(7) TC_CASE(8) TC_CASE(9) TC_CASE(10)
              }
#undef TC_CASE
            } else {
              const double *bp = P.Jq + (int64_t)v * P.jq_row + ps.jq_off + (int64_t)ps.nsb * kc0 * 32
                               + (size_t)warp_n * kcs_q * 32 + lane;
              tc_mma_row_global<NW>(acc0, acc1, a0p, a1p, AROW, m0, m1, bp, kcs_q, b_sb_stride, nactw);
            }
            const int n0 = (ix0 > v) ? ix0 : 0x7fffffff, n1 = (ix1 > v) ? ix1 : 0x7fffffff;
            v = __reduce_min_sync(0xffffffffu, min(n0, n1));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[8 + s]));
      }
    }

    // ---- epilogue: 1/2 of the trapezoid and the divisions of transfer2.c:2092-2098 ----
    const int cpair = 2 * (lane & 3);
    const int u0 = unit_id[g0 * 8 + um], u1 = unit_id[g1 * 8 + um];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const int sb = warp_n + j * WN;
      if (sb >= ps.nsb) continue;
      const int cbase = (ps.sb0 + sb) * 8 + cpair;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = cbase + h;
        if (c >= B.ncols) continue;
        const int tt = P.col_tt[B.col_off + c];
        const double lfac = 2. * P.col_l[B.col_off + c] + 1;
        double t0 = acc0[j][h] * 0.5; t0 /= 4.; t0 /= lfac; t0 /= 2.;
        double t1 = acc1[j][h] * 0.5; t1 /= 4.; t1 /= lfac; t1 /= 2.;
        if (u0 >= 0) { if (t0 != t0) atomicOr(P.err, ERR_NAN); P.out[(int64_t)tt * P.out_stride + u0] = t0; }
        if (u1 >= 0) { if (t1 != t1) atomicOr(P.err, ERR_NAN); P.out[(int64_t)tt * P.out_stride + u1] = t1; }
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
  int n_blocks = 0, n_pass = 0;
  TcBlock blk[TC_MAX_BLOCKS];
  TcPass pass[TC_MAX_PASSES];
  std::vector<int> term_tp; std::vector<double> term_sign;
  int64_t jq_row = 0;
  int wm = 8, wn = 1, nw = 8;
  int kcs_max = 1, nsb_pass_max = 1;
  Buf<double> Jq, keys_out;
  Buf<double4> CSD, E;
  Buf<int> col_tt, col_l, col_lm, term_col0, nact, perm_in, perm, unit_pair;
  Buf<TcUnit> units;
  Buf<unsigned char> cub_tmp;
  TcParams tp{};                 // kernel parameter block (about 2 KB: kept off the stack)
  cudaEvent_t ev[6]{};
  bool ev_ready = false;
};

FastState *fast_state_create() { return new FastState(); }

void fast_state_destroy(FastState *f)
{
  if (!f) return;
  f->Jq.release(); f->keys_out.release(); f->CSD.release(); f->E.release();
  f->col_tt.release(); f->col_l.release(); f->col_lm.release(); f->term_col0.release(); f->nact.release();
  f->perm_in.release(); f->perm.release(); f->unit_pair.release(); f->units.release(); f->cub_tmp.release();
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
  std::vector<int> col_tt, col_l, col_lm, term_col0;
  fs->term_tp.clear(); fs->term_sign.clear();
  fs->n_blocks = 0;
  int max_sb = 0, max_kc = 0;
  for (int type = 0; type < 3; ++type) {
    bool on = (type == 0 && h.has_T) || (type == 1 && h.has_E) || (type == 2 && h.has_B);
    if (!on) continue;
    int tt_base = type == 0 ? h.index_tt2_T : (type == 1 ? h.index_tt2_E : h.index_tt2_B);
    int Lmax = type == 0 ? h.l_max_los_t : h.l_max_los_p;
    for (int im = 0; im < h.m_size; ++im) {
      int mval = h.m[im];
      if (fs->n_blocks >= TC_MAX_BLOCKS) { err = "too many (field,m) blocks for the FAST kernel"; return 1; }
      TcBlock B{};
      B.col_off = (int)col_tt.size();
      B.term_off = (int)fs->term_tp.size();
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
          fs->term_tp.push_back(sg.tp + ls);
          fs->term_sign.push_back(sg.sign);
          term_col0.push_back((sg.J * h.L_size + L) * h.n_transfers);
        }
      }
      B.nterms = (int)fs->term_tp.size() - B.term_off;
      if (B.nterms > TC_MAX_TERMS) { err = "too many K terms for the FAST kernel"; return 1; }
      if (B.nterms == 0) {                      // nothing contributes: outputs are exactly zero
        fs->term_tp.push_back(0); fs->term_sign.push_back(0.); term_col0.push_back(0); B.nterms = 1;
      }
      B.nkc = (B.nterms + 1) / 2;
      B.nsb = (B.ncols + 7) / 8;
      max_sb = std::max(max_sb, B.nsb);
      max_kc = std::max(max_kc, B.nkc);
      fs->blk[fs->n_blocks++] = B;
    }
  }
  if ((int)fs->term_tp.size() > TC_MAX_TERMS * 2) { err = "too many K terms in total for the FAST kernel"; return 1; }

  // warp grid (WM x WN consumer warps, NW sub-blocks per warp): widest tile of units that still
  // covers the columns of a block in one pass
  if (max_sb <= 8) { fs->wm = 8; fs->wn = 1; fs->nw = 8; }
  else if (max_sb <= 10) { fs->wm = 8; fs->wn = 1; fs->nw = 10; }
  else if (max_sb <= 16) { fs->wm = 4; fs->wn = 2; fs->nw = 8; }
  else if (max_sb <= 20) { fs->wm = 4; fs->wn = 2; fs->nw = 10; }
  else if (max_sb <= 32) { fs->wm = 2; fs->wn = 4; fs->nw = 8; }
  else { fs->wm = 2; fs->wn = 4; fs->nw = 10; }
  if (const char *v = getenv("SONG_LOS_TC_SHAPE")) {          // A/B measurements: "wm,wn,nw"
    int a, b, c;
    if (sscanf(v, "%d,%d,%d", &a, &b, &c) == 3 && a * b == 8 && (c == 8 || c == 10) && (a == 8 || a == 4 || a == 2)) {
      fs->wm = a; fs->wn = b; fs->nw = c;
    }
  }
  const int sb_per_pass = fs->wn * fs->nw;
  int kcs_cap = TC_KCS_MAX;
  if (const char *v = getenv("SONG_LOS_TC_KCS")) kcs_cap = std::max(1, std::min(TC_KCS_MAX, atoi(v)));

  // passes and the row layout
  fs->n_pass = 0;
  int64_t off = 0;
  fs->kcs_max = 1; fs->nsb_pass_max = 1;
  for (int ib = 0; ib < fs->n_blocks; ++ib) {
    const TcBlock &B = fs->blk[ib];
    int npass_b = (B.nsb + sb_per_pass - 1) / sb_per_pass;
    int per = (B.nsb + npass_b - 1) / npass_b;                // balanced passes
    for (int sb0 = 0; sb0 < B.nsb; sb0 += per) {
      if (fs->n_pass >= TC_MAX_PASSES) { err = "too many column passes for the FAST kernel"; return 1; }
      TcPass ps{};
      ps.blk = ib; ps.sb0 = sb0; ps.nsb = std::min(per, B.nsb - sb0);
      ps.nq = (B.nkc + kcs_cap - 1) / kcs_cap;
      ps.kcs = (B.nkc + ps.nq - 1) / ps.nq;
      ps.jq_off = off;
      off += (int64_t)ps.nsb * B.nkc * 32;
      fs->kcs_max = std::max(fs->kcs_max, ps.kcs);
      fs->nsb_pass_max = std::max(fs->nsb_pass_max, ps.nsb);
      fs->pass[fs->n_pass++] = ps;
    }
  }
  fs->jq_row = off;
  (void)max_kc;

  auto up = [&](auto &buf, const auto &vec) -> cudaError_t {
    cudaError_t e = buf.reserve(vec.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice, st);
  };
  FCK(up(fs->col_tt, col_tt)); FCK(up(fs->col_l, col_l)); FCK(up(fs->col_lm, col_lm));
  FCK(up(fs->term_col0, term_col0));
  FCK(fs->Jq.reserve((size_t)h.xx_size * fs->jq_row));
  for (int ip = 0; ip < fs->n_pass; ++ip) {
    const TcPass &ps = fs->pass[ip];
    const TcBlock &B = fs->blk[ps.blk];
    int64_t n = (int64_t)B.nkc * ps.nsb * 8;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)h.xx_size);
    tc_build_table_kernel<<<grid, 128, 0, st>>>(P.proj.J, P.proj.n_cols, fs->jq_row, ps, B, fs->term_col0.p,
                                                fs->col_lm.p, fs->Jq.p);
    FCK(cudaGetLastError());
  }
  FCK(fs->nact.reserve((size_t)fs->n_blocks * h.xx_size));
  FCK(cudaMemsetAsync(fs->nact.p, 0, (size_t)fs->n_blocks * h.xx_size * sizeof(int), st));
  for (int ib = 0; ib < fs->n_blocks; ++ib) {
    tc_nact_kernel<<<(unsigned)h.xx_size, 64, 0, st>>>(P.proj.J, P.proj.n_cols, h.xx_size, fs->blk[ib], ib,
                                                        fs->term_col0.p, fs->col_lm.p, fs->nact.p);
    FCK(cudaGetLastError());
  }
  tc_nact_cummax_kernel<<<1, 32, 0, st>>>(fs->nact.p, h.xx_size, fs->n_blocks);
  FCK(cudaGetLastError());
  FCK(cudaStreamSynchronize(st));
  fs->table_ready = true;
  return 0;
}

template <int WM, int WN, int NW>
static cudaError_t launch_tc(const TcParams &TP, int64_t n_tiles, size_t smem, cudaStream_t st)
{
  cudaError_t e = cudaFuncSetAttribute(los_tc_kernel<WM, WN, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  los_tc_kernel<WM, WN, NW><<<(unsigned)n_tiles, TC_THREADS, smem, st>>>(TP);
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
  if (b.n_pairs > 65535) { err = "more than 65535 pairs in one batch"; return 1; }

  // ---- phase 0: transposed {y, dd} sources, edge rows ----
  FCK(cudaEventRecord(fs->ev[0], st));
  const int64_t n_src = h.n_src_total;
  FCK(fs->CSD.reserve((size_t)n_src));                           // (np - 1) intervals per pair: n_src is an upper bound
  {
    int64_t cols = rowlen * b.n_pairs;
    if (cols > 0) {
      tc_spline_csd_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(b, P.tm.tau_size, P.ty.tp2_size,
                                                                          P.cubic_k3, fs->CSD.p, P.err);
      FCK(cudaGetLastError());
      tm->n_launches++;
    }
  }
  FCK(fs->E.reserve((size_t)b.n_pairs * 2 * rowlen));
  if (P.flat_extrap && b.n_pairs > 0) {
    dim3 grid((unsigned)std::min<int64_t>((rowlen + 255) / 256, 64), (unsigned)b.n_pairs, 2);
    tc_edge_rows_kernel<<<grid, 256, 0, st>>>(b, rowlen, fs->CSD.p, fs->E.p, P.err);
    FCK(cudaGetLastError());
    tm->n_launches++;
  }
  FCK(cudaEventRecord(fs->ev[1], st));

  // ---- phase 1: global k3 order ----
  FCK(fs->perm_in.reserve((size_t)U)); FCK(fs->perm.reserve((size_t)U)); FCK(fs->keys_out.reserve((size_t)U));
  FCK(fs->unit_pair.reserve((size_t)U));
  FCK(cudaMemcpyAsync(fs->unit_pair.p, h_unit_pair, (size_t)U * sizeof(int), cudaMemcpyHostToDevice, st));
  if (U > 0) {
    tc_iota_kernel<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(fs->perm_in.p, U);
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
  if (U > 0) {
    tc_prep_units_kernel<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(b, fs->perm.p, fs->unit_pair.p, rowlen,
                                                                     P.flat_extrap, fs->CSD.p, fs->E.p, fs->units.p, P.err);
    FCK(cudaGetLastError());
    tm->n_launches++;
  }
  FCK(cudaEventRecord(fs->ev[3], st));

  // ---- phase 2: the DMMA kernel ----
  TcParams &TP = fs->tp;
  TP = TcParams{};
  TP.n_blocks = fs->n_blocks; TP.n_pass = fs->n_pass;
  for (int i = 0; i < fs->n_blocks; ++i) TP.blk[i] = fs->blk[i];
  for (int i = 0; i < fs->n_pass; ++i) TP.pass[i] = fs->pass[i];
  for (size_t i = 0; i < fs->term_tp.size(); ++i) { TP.term_tp[i] = fs->term_tp[i]; TP.term_sign[i] = fs->term_sign[i]; }
  TP.col_tt = fs->col_tt.p; TP.col_l = fs->col_l.p;
  TP.Jq = fs->Jq.p; TP.jq_row = fs->jq_row; TP.nact = fs->nact.p;
  TP.units = fs->units.p; TP.n_units = U;
  TP.rowlen = rowlen; TP.tau_size = P.tm.tau_size; TP.tp2_size = P.ty.tp2_size;
  TP.tau0_minus_tau = P.tm.tau0_minus_tau; TP.delta_tau = P.tm.delta_tau; TP.xx = P.xx;
  TP.inv_xx_step = 1.0 / P.xx_step; TP.xx_size = P.xx_size;
  TP.out = b.out; TP.out_stride = U; TP.err = P.err;

  const int MT = 16 * fs->wm;
  const int arow = MT * 4 + TC_APAD;
  auto r128 = [](size_t x) { return (x + 127) / 128 * 128; };
  TP.hdr_bytes = (int)r128(16 + (size_t)MT * 4);
  TP.a_bytes = (int)r128((size_t)fs->kcs_max * arow * 8);
  TP.b_row_bytes = fs->nsb_pass_max * fs->kcs_max * 256;
  int rmax = (MT >= 128) ? 4 : (MT >= 64 ? 3 : 2);
  if (const char *v = getenv("SONG_LOS_TC_RMAX")) rmax = std::max(1, std::min(8, atoi(v)));
  const size_t fixed = 128 + (size_t)MT * 4;
  const size_t budget = 224 * 1024;
  auto stage_bytes = [&](int r) { return (size_t)TP.hdr_bytes + TP.a_bytes + (size_t)r * TP.b_row_bytes; };
  while (rmax > 1 && (budget - fixed) / stage_bytes(rmax) < 3) --rmax;
  int ns = (int)std::min<size_t>(8, (budget - fixed) / stage_bytes(rmax));
  if (const char *v = getenv("SONG_LOS_TC_STAGES")) ns = std::max(2, std::min(ns, atoi(v)));
  if (ns < 2) { err = "FAST kernel: shared-memory stage does not fit"; return 1; }
  TP.rmax = rmax; TP.n_stages = ns;
  TP.stage_bytes = (int)stage_bytes(rmax);
  const size_t smem = fixed + (size_t)ns * TP.stage_bytes;

  const int64_t n_tiles = (U + MT - 1) / MT;
  if (n_tiles > 0) {
    cudaError_t e = cudaErrorInvalidValue;
    const int key = fs->wm * 100 + fs->nw;
    switch (key) {
      case 808:  e = launch_tc<8, 1, 8>(TP, n_tiles, smem, st); break;
      case 810:  e = launch_tc<8, 1, 10>(TP, n_tiles, smem, st); break;
      case 408:  e = launch_tc<4, 2, 8>(TP, n_tiles, smem, st); break;
      case 410:  e = launch_tc<4, 2, 10>(TP, n_tiles, smem, st); break;
      case 208:  e = launch_tc<2, 4, 8>(TP, n_tiles, smem, st); break;
      case 210:  e = launch_tc<2, 4, 10>(TP, n_tiles, smem, st); break;
    }
    FCK(e);
    tm->n_launches++;
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
