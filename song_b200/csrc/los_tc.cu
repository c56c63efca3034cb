// los_tc.cu -- FAST mode: k3-sorted FP64 tensor-core (DMMA m8n8k4) line-of-sight kernel fed by a
// TMA / mbarrier operand pipeline.
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
//   spline        k3 second derivatives of every (pair, tau, source type) column (transfer2.c:2639): the grid-only part
//                 of the elimination once per pair (spline_coef), then a division-free sweep per column (spline_csd_coef),
//                 written transposed: CSD[pair][j][tau][tp] = {y, dd}
//   sort          cub radix sort of (k3, unit), on a second stream beside the splines
//   prep_units    per unit: bracket j, spline weights (transfer2.c:2683-2691), sorted position; flat
//                 extrapolation = weights of the first/last physical k3 (transfer2.c:2743-2757)
//   then, per (column block, K chunk) -- ONE source slab, reused, so that a batch can hold many k3 values:
//   interp_units  one CTA per (k1,k2) pair walks the pair's units in k3 order, so that the two CSD rows
//                 of an interval are read once for all the units inside it, and writes the interpolated
//                 sources of the chunk in the order the tensor-core kernel consumes them:
//                 S[sorted unit][tau][slot]   (slot = term of the chunk)
//   los_tc        persistent CTA per SM, tiles of MT sorted units.  Producer warp(s): a single thread issues, per
//                 tau, ONE cp.async.bulk.tensor (TMA, 3-D box {slots, 1 tau, MT units}) for the sources
//                 and one cp.async.bulk per needed table row into a shared-memory ring; the lanes
//                 compute row index and quadrature x lerp weights of every unit.  8 or 12 consumer warps
//                 wait on the stage's mbarrier, read 256-byte fragments and issue DMMAs into register
//                 accumulators.  Columns whose J is still below the reference's cut at this row (x << l) are
//                 skipped.  Later chunks of a block start from the partial sums the previous launch left in
//                 the result array; the last one applies 1/(16 (2l+1)).
#include <cub/cub.cuh>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <algorithm>
#include <cmath>
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
constexpr int TC_MAX_BLOCKS = FAST_MAX_BLOCKS;
constexpr int TC_MAX_GROUPS = 12;         // (block, K chunk): one source slab and one tensor map each
constexpr int TC_MAX_PASSES = 24;
constexpr int TC_MAX_TERMS = 64;          // K terms of one block (2 per kc)
constexpr int TC_KCS_MAX = 6;             // kc's (pairs of terms) per pipeline stage
// WM x WN consumer warps and NP producer warps are template parameters of the kernel
constexpr int TC_MAX_STAGES = 8;

struct TcBlock {
  int ncols, nsb;                         // real columns, 8-column sub-blocks
  int nterms, nkc;                        // real K terms, pairs of terms
  int col_off, term_off;                  // offsets into col_* / term_* tables
};

struct TcGroup {                          // K chunk q of a block: kc in [kc0, kc0 + kcs)
  int blk, kc0, kcs;
  int tpp;                                // doubles per (unit, tau) row of the slab: 2*kcs
  int64_t s_off;                          // offset of the slab inside the S buffer (doubles)
};

struct TcPass {                           // one column pass of a block = one accumulator lifetime
  int blk, sb0, nsb;                      // block, first sub-block, sub-blocks in the pass
  int g0, nq;                             // first group and number of K chunks of the block
  int64_t jq_off;                         // offset of the pass inside a table row (doubles)
};

struct TcParams {
  CUtensorMap tmap[TC_MAX_GROUPS];        // S[group]: dims {tpp, tau, units}, box {tpp, 1, MT}
  int n_blocks, n_groups, n_pass;
  TcBlock blk[TC_MAX_BLOCKS];
  TcGroup grp[TC_MAX_GROUPS];
  TcPass pass[TC_MAX_PASSES];
  const int *col_tt; const double *col_scale;   // per column: transfer type index, 1 / (16 (2l+1))
  const double *Jq; int64_t jq_row;       // tensor-core table and its row length (doubles)
  const int *nact;                        // [n_blocks][xx_size] active sub-block prefix per row
  const double *k_sorted; const int *perm; int64_t n_units;   // sorted k3 and original unit index
  int sm_count;
  int tau_size;
  const double *tau0_minus_tau, *delta_tau, *xx;
  double inv_xx_step; int xx_size;
  double xx_step_ref;                     // pbs2->xx_step itself: divisor of transfer2.c:2199 where x * inv_xx_step cannot decide
  double xx_step; int xx_uniform;         // xx[i] == i * xx_step BIT FOR BIT (xx_step = the grid's own delta): no table read for a_J
  int n_cols_total, cols_pad;             // columns of all blocks; padded to 32 (shared-memory copies)
  int n_stages, rmax;                     // ring depth, table rows per stage
  int hdr_bytes, a_bytes, b_row_bytes, stage_bytes;  // shared-memory geometry of one stage
  double *out; int64_t out_stride;        // [tt][n_units]
  int *err;
  int pass_begin, pass_end;               // passes of this launch (one column block)
  int q_fixed;                            // K chunk of this launch (-1: all chunks of the block)
  int acc_in, acc_final;                  // accumulators start from the partial sums in `out` / are scaled and final
  int dbg;                                // timing experiments only (results invalid): 4 = no MMAs, 8 = no source copies, 16 = no table copies
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk copies (TMA), FP64 MMA
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
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// L2 eviction policies for the two operand streams of the kernel.  The source slab is read exactly once per launch
// (86 GB at configs[3]) while the table rows of a wave of tiles (~20 MB) are re-read by every CTA of the wave and at
// neighbouring tau: without hints the stream evicts the rows and half of the row fetches go to DRAM (ncu: 300 GB
// read per launch for 86 GB of slab).
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
// 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// 3-D tensor (TMA descriptor) box copy global -> shared
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// index_x of transfer2.c:2199, (int)(x / xx_step), without a division on the common path.  x * (1/xx_step) differs from
// the correctly rounded quotient by a few ulp (< 1e-11 for x/xx_step < 1e5), which can change the integer part only
// when the product lies that close to an integer; within 1e-9 of one the division itself decides, so the row index is
// the reference's for every x (a unit sitting exactly on a grid point included).
__device__ __noinline__ int tc_index_x_div(double x, double step) { return (int)(x / step); }   // rare: kept out of line

__device__ __forceinline__ int tc_index_x(double x, double inv_step, double step)
{
  const double q = x * inv_step;
  int ix = (int)q;
  const double f = q - (double)ix;
  if (f < 1e-9 || f > 1.0 - 1e-9) ix = tc_index_x_div(x, step);
  return ix;
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
                                      int kcs, const int *term_col0, const int *col_lm, double *Jq)
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
  int q = kc / kcs, kcq = kc - q * kcs;
  int kcs_q = min(kcs, blk.nkc - q * kcs);
  int sb = cp >> 3, c8 = cp & 7;
  double *dst = Jq + (int64_t)ix * jq_row + ps.jq_off + (int64_t)ps.nsb * (q * kcs) * 32
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
// transposed sources CSD[pair][j][tau][tp] = {y, dd}: dd = second derivative in k3 of every (pair, tau, tp) column
// (array_spline_table_columns with _SPLINE_EST_DERIV_, transfer2.c:2639), one thread per column
// ---------------------------------------------------------------------------------------------
constexpr int TC_SPLINE_MAX_NP = 256;      // columns up to this length keep their elimination workspace in registers / local memory

// linear k3 interpolation (quick_song_run): the transposed copy alone, dd = 0
__global__ void __launch_bounds__(128) tc_transpose_csd_kernel(Batch b, int tau_size, int tp2_size, double2 *CSD)
{
  int64_t rowlen = (int64_t)tp2_size * tau_size;
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= rowlen * b.n_pairs) return;
  int p = (int)(c / rowlen);
  int64_t col = c - (int64_t)p * rowlen;           // col = tau*tp2 + tp   (tp fastest)
  int tau = (int)(col / tp2_size), tp = (int)(col % tp2_size);
  int n = b.np[p];
  const double *y = b.src + b.src_off[p] + ((int64_t)tp * tau_size + tau) * n;
  double2 *cs = CSD + b.src_off[p] + col;
  for (int i = 0; i < n; ++i) cs[(int64_t)i * rowlen] = make_double2(y[i], 0.);
}

// The part of the tridiagonal elimination that depends on the k3 grid only is the same for the tp2 x tau_size columns
// of a pair: one thread per pair computes it once (reciprocal spacings, sig, 1/pp, the dd coefficients and the weights
// of the end-derivative estimates); the column kernel below is then free of divisions.
struct TcSplineCoef { double rh, r6, sig, rpp, ddc; };     // per coarse k3 point, indexed pt_off[p] + i
struct TcSplineEnds { double fa, fb, la, lb, rbot; };      // per pair

__global__ void tc_spline_coef_kernel(Batch b, TcSplineCoef *coef, TcSplineEnds *ends, int *err)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= b.n_pairs) return;
  const int n = b.np[p];
  const double *x = b.k3_pt + b.pt_off[p];
  TcSplineCoef *c = coef + b.pt_off[p];
  if (n < 3) { atomicOr(err, ERR_H_ZERO); return; }
  double ddp = -0.5;
  c[0].ddc = ddp; c[0].sig = 0.; c[0].rpp = 0.; c[0].r6 = 0.;
  for (int i = 0; i + 1 < n; ++i) c[i].rh = 1.0 / (x[i + 1] - x[i]);
  c[n - 1].rh = 0.;
  for (int i = 1; i + 1 < n; ++i) {
    const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    const double pp = sig * ddp + 2.0;
    ddp = (sig - 1.0) / pp;
    c[i].sig = sig; c[i].rpp = 1.0 / pp; c[i].ddc = ddp; c[i].r6 = 6.0 / (x[i + 1] - x[i - 1]);
  }
  TcSplineEnds e;
  const double a = x[2] - x[0], cc = x[1] - x[0], g = x[2] - x[1];
  e.fa = a * a / (a * cc * g); e.fb = cc * cc / (a * cc * g);
  const double al = x[n - 3] - x[n - 1], cl = x[n - 2] - x[n - 1], gl = x[n - 3] - x[n - 2];
  e.la = al * al / (al * cl * gl); e.lb = cl * cl / (al * cl * gl);
  e.rbot = 1.0 / (0.5 * c[n - 2].ddc + 1.0);
  ends[p] = e;
}

__global__ void __launch_bounds__(128) tc_spline_csd_coef_kernel(Batch b, int tau_size, int tp2_size, const TcSplineCoef *coef,
                                                                 const TcSplineEnds *ends, double2 *CSD)
{
  int64_t rowlen = (int64_t)tp2_size * tau_size;
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= rowlen * b.n_pairs) return;
  int p = (int)(c / rowlen);
  int64_t col = c - (int64_t)p * rowlen;           // col = tau*tp2 + tp   (tp fastest)
  int tau = (int)(col / tp2_size), tp = (int)(col % tp2_size);
  int n = b.np[p];
  if (n < 3) return;                               // flagged by the coefficient kernel
  const double *y = b.src + b.src_off[p] + ((int64_t)tp * tau_size + tau) * n;
  const TcSplineCoef *k = coef + b.pt_off[p];
  const TcSplineEnds e = ends[p];
  double2 *cs = CSD + b.src_off[p] + col;
  // right-hand sides of the elimination: in local memory for the usual column lengths; longer columns park them in the
  // dd slot of the output, which the back substitution overwrites in the same order (any k3_size works, as in the reference)
  const bool big = n > TC_SPLINE_MAX_NP;
  double u[TC_SPLINE_MAX_NP];
#define TC_U_ST(I, V) do { if (big) cs[(int64_t)(I) * rowlen].y = (V); else u[(I)] = (V); } while (0)
#define TC_U_LD(I) (big ? cs[(int64_t)(I) * rowlen].y : u[(I)])
  double ylo = y[0], ymid = y[1];
  const double dy_first = e.fa * (ymid - ylo) - e.fb * (y[2] - ylo);
  double up = (3. * k[0].rh) * ((ymid - ylo) * k[0].rh - dy_first);
  TC_U_ST(0, up);
  double rh_prev = k[0].rh;
  for (int i = 1; i + 1 < n; ++i) {
    const TcSplineCoef ki = k[i];
    const double yhi = y[i + 1];
    const double ui = (yhi - ymid) * ki.rh - (ymid - ylo) * rh_prev;
    up = (ki.r6 * ui - ki.sig * up) * ki.rpp;
    TC_U_ST(i, up);
    ylo = ymid; ymid = yhi; rh_prev = ki.rh;
  }
  // here ylo = y[n-2], ymid = y[n-1], up = u[n-2]
  const double dy_last = e.la * (ylo - ymid) - e.lb * (y[n - 3] - ymid);
  const double un = (3. * k[n - 2].rh) * (dy_last - (ymid - ylo) * k[n - 2].rh);
  double next = (un - 0.5 * up) * e.rbot;
  cs[(int64_t)(n - 1) * rowlen] = make_double2(ymid, next);
  for (int i = n - 2; i >= 0; --i) {
    next = k[i].ddc * next + TC_U_LD(i);
    cs[(int64_t)i * rowlen] = make_double2(y[i], next);
  }
#undef TC_U_ST
#undef TC_U_LD
}

// unit -> pair map: one CTA per pair writes its run of the pair-major unit sequence
__global__ void tc_unit_pair_kernel(Batch b, int *unit_pair)
{
  const int p = blockIdx.x;
  int *dst = unit_pair + b.tr_off[p];
  for (int i = threadIdx.x; i < b.nt[p]; i += blockDim.x) dst[i] = p;
}

__global__ void tc_iota_kernel(int *p, int64_t n)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}

// prep_units: per unit in ORIGINAL order (pair-major, k3 ascending inside a pair): the k3 bracket and
// spline weights of transfer2.c:2683-2691 and the position of the unit in the k3-sorted sequence.
// Flat extrapolation (transfer2.c:2743-2757): wing units take the weights of the first / last
// physical k3 of their pair.
struct TcUnitW { double a, b, a3h, b3h; };          // a, b, (a^3-a) h^2/6, (b^3-b) h^2/6

__global__ void tc_prep_units_kernel(Batch b, const int *perm, const int *unit_pair, int flat_extrap,
                                     TcUnitW *uw, int2 *ujp, int *jcount, int *err)
{
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // sorted position
  if (s >= b.n_units) return;
  int u = perm[s];
  int p = unit_pair[u];
  int i = u - (int)b.tr_off[p];
  const double *kp = b.k3_pt + b.pt_off[p];
  const double *kt = b.k3_tr + b.tr_off[p];
  int np = b.np[p];
  int first = b.phys_start[p], last = first + b.phys_size[p] - 1;
  double ke = kt[i];
  if (flat_extrap) ke = (i < first) ? b.phys_k_first[p] : (i > last) ? b.phys_k_last[p] : ke;   // k3 of the first / last physical point
  // j = number of interior coarse points kp[i], 1 <= i <= np-2, strictly below ke (strict <, :2683)
  int lo = 1, hi = np - 1;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (kp[mid] < ke) lo = mid + 1; else hi = mid; }
  int j = min(lo - 1, np - 2);
  double h = kp[j + 1] - kp[j];
  if (h == 0) atomicOr(err, ERR_H_ZERO);
  double bb = (ke - kp[j]) / h, a = 1 - bb, h6 = h * h / 6.0;
  TcUnitW w; w.a = a; w.b = bb; w.a3h = (a * a * a - a) * h6; w.b3h = (bb * bb * bb - bb) * h6;
  uw[u] = w;
  ujp[u] = make_int2(j, (int)s);
  atomicAdd(&jcount[b.pt_off[p] + j], 1);
}

// interp_units: CTA = (pair, chunk of TAUC tau samples).  The CTA walks the k3 intervals of the pair;
// per interval it stages the one new CSD row in shared memory (each row is read once per CTA) and its
// warps take the units inside the interval round-robin.  A warp evaluates the spline of its unit for
// every (group, tau, slot) column of the chunk -- 4 FMAs each -- and writes them as one contiguous
// run per group of S[group][sorted unit][tau][slot]: full 256-byte stores, sequential per unit.
constexpr int TC_SLOTS = 2 * TC_KCS_MAX + 4;
constexpr int TC_INTERP_THREADS = 384;      // measured: 2.58 / 2.37 / 2.31 / 2.53 ms at 128 / 256 / 384 / 512 threads
struct TcInterp {
  int n_groups;
  int tauc;                               // tau samples per CTA
  int col_stride;                         // tauc * (largest tpp): row length of the column map
  int grp_tpp[TC_MAX_GROUPS];             // slots per (unit, tau) row of the group
  int64_t grp_s_off[TC_MAX_GROUPS];       // slab offsets
  const int *slot_tp;                     // per (group, slot): source type, -1 = padding   [n_groups][TC_SLOTS]
  const double *slot_sign;
  const int *jcount;                      // units per (pair, interval): indexed pt_off[p] + j
};

__global__ void __launch_bounds__(512) tc_interp_units_kernel(Batch b, TcInterp I, int tau_size, int tp2_size,
                                                              const double2 *CSD, const TcUnitW *uw, const int2 *ujp, double *S)
{
  extern __shared__ __align__(16) unsigned char interp_smem[];
  const int p = blockIdx.x;
  const int tau0 = blockIdx.y * I.tauc;
  const int ntau = min(I.tauc, tau_size - tau0);
  const int chunk = ntau * tp2_size;                                  // double2 entries of a row chunk
  double2 *rows = reinterpret_cast<double2 *>(interp_smem);           // [2][tauc * tp2]
  int *s_tp = reinterpret_cast<int *>(rows + 2 * (size_t)I.tauc * tp2_size);   // [n_groups][TC_SLOTS]
  double *s_sign = reinterpret_cast<double *>(s_tp + TC_MAX_GROUPS * TC_SLOTS);
  // per (group, column of the chunk): index of the staged {y, dd} entry (tau_local * tp2 + tp) and the slot,
  // packed as idx | slot << 24; -1 for padding slots.  Replaces a division per element in the inner loop.
  int *s_col = reinterpret_cast<int *>(s_sign + TC_MAX_GROUPS * TC_SLOTS);   // [n_groups][tauc * tpp_max]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x / 32;
  for (int i = tid; i < I.n_groups * TC_SLOTS; i += blockDim.x) { s_tp[i] = I.slot_tp[i]; s_sign[i] = I.slot_sign[i]; }
  for (int g = 0; g < I.n_groups; ++g) {
    const int tpp = I.grp_tpp[g];
    for (int c = tid; c < ntau * tpp; c += blockDim.x) {
      const int tl = c / tpp, slot = c - tl * tpp;
      const int tp = I.slot_tp[g * TC_SLOTS + slot];
      s_col[g * I.col_stride + c] = (tp >= 0) ? ((tl * tp2_size + tp) | (slot << 24)) : -1;
    }
  }
  const int np = b.np[p];
  const int64_t u0 = b.tr_off[p];
  const int64_t rowlen = (int64_t)tau_size * tp2_size;
  const double2 *cs = CSD + b.src_off[p] + (int64_t)tau0 * tp2_size;
  const int *jcount = I.jcount + b.pt_off[p];
  for (int c = tid; c < chunk; c += blockDim.x) rows[c] = cs[c];      // row 0
  int start = 0;
  for (int j = 0; j + 1 < np; ++j) {
    double2 *r0 = rows + (size_t)(j & 1) * I.tauc * tp2_size, *r1 = rows + (size_t)((j + 1) & 1) * I.tauc * tp2_size;
    for (int c = tid; c < chunk; c += blockDim.x) r1[c] = cs[(int64_t)(j + 1) * rowlen + c];
    __syncthreads();
    const int cnt = jcount[j];
    for (int i = start + warp; i < start + cnt; i += nwarps) {
      const TcUnitW w = uw[u0 + i];
      const int64_t pos = ujp[u0 + i].y;
      for (int g = 0; g < I.n_groups; ++g) {
        const int tpp = I.grp_tpp[g];
        double *dst = S + I.grp_s_off[g] + (pos * tau_size + tau0) * tpp;
        const int ncol = ntau * tpp;
        const int *colmap = s_col + g * I.col_stride;
        const double *sgn = s_sign + g * TC_SLOTS;
        for (int c = lane; c < ncol; c += 32) {
          const int m = colmap[c];
          double v = 0.;
          if (m >= 0) {
            const int e = m & 0xffffff;
            const double2 a0 = r0[e], a1 = r1[e];
            // transfer2.c:2697 (linear: dd = 0) / :2703-2705 (cubic)
            v = sgn[m >> 24] * (w.a * a0.x + w.b * a1.x + w.a3h * a0.y + w.b3h * a1.y);
          }
          dst[c] = v;
        }
      }
    }
    start += cnt;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// the DMMA kernel
// ---------------------------------------------------------------------------------------------
// shared memory:  [bars: full[8], empty[8]] [tau tables] [producer row table] [epilogue column tables] [stages...]
// stage:          header {meta int4 {row_lo, nrows, nact, row_hi}, ixs[MT], cc[MT] double2 {c0, c1}}
//                 | S[MT][tpp] (TMA box) | B[rmax][b_row]     (each part padded to 128 bytes)
struct TcStagePtrs {
  int *meta; int *ixs; double2 *cc; double *S; double *B;
};

template <int MT>
__device__ __forceinline__ TcStagePtrs tc_stage(unsigned char *stage0, int s, const TcParams &P)
{
  unsigned char *base = stage0 + (size_t)s * P.stage_bytes;
  TcStagePtrs r;
  r.meta = reinterpret_cast<int *>(base);
  r.ixs = reinterpret_cast<int *>(base + 16);
  r.cc = reinterpret_cast<double2 *>(base + 16 + MT * 4);
  r.S = reinterpret_cast<double *>(base + P.hdr_bytes);
  r.B = reinterpret_cast<double *>(base + P.hdr_bytes + P.a_bytes);
  return r;
}

// ---- consumer inner loops -------------------------------------------------------------------
// One table row against the staged sources of a chunk.  NACT (active sub-blocks of this warp) is a
// template parameter so that the body is straight-line code: per kc 1-2 A fragments (source value x
// per-unit weight, zero weight for units on another row) and NACT B fragments (LDS.64, 256 contiguous
// bytes per warp) feed 1-2 x NACT DMMAs; the fragments of kc+1 are loaded into the other register
// set before the DMMAs of kc are issued (ping-pong, no moves).

// both 8-unit groups of the warp have units on this row
template <int NACT, int NW, bool PP>
__device__ __forceinline__ void tc_mma_both(double (&acc0)[NW][2], double (&acc1)[NW][2],
                                            const double *a0p, const double *a1p, double w0, double w1,
                                            const double *bp, int kcs, int b_sb_stride)
{
  if constexpr (!PP) {
    // single-buffered: with three or four warps per SM sub-partition the other warps cover the LDS latency
    // B fragments in two halves when a warp owns many sub-blocks: fewer live registers (NW = 10 at 128 registers)
    constexpr int H0 = (NACT > 8) ? (NACT + 1) / 2 : NACT;
    for (int kc = 0; kc < kcs; ++kc) {
      const double x0 = w0 * a0p[2 * kc], x1 = w1 * a1p[2 * kc];
      {
        double bf[H0];
#pragma unroll
        for (int j = 0; j < H0; ++j) bf[j] = bp[(size_t)j * b_sb_stride + kc * 32];
#pragma unroll
        for (int j = 0; j < H0; ++j) dmma884(acc0[j][0], acc0[j][1], x0, bf[j]);
#pragma unroll
        for (int j = 0; j < H0; ++j) dmma884(acc1[j][0], acc1[j][1], x1, bf[j]);
      }
      if constexpr (H0 < NACT) {
        double bf[NACT - H0];
#pragma unroll
        for (int j = H0; j < NACT; ++j) bf[j - H0] = bp[(size_t)j * b_sb_stride + kc * 32];
#pragma unroll
        for (int j = H0; j < NACT; ++j) dmma884(acc0[j][0], acc0[j][1], x0, bf[j - H0]);
#pragma unroll
        for (int j = H0; j < NACT; ++j) dmma884(acc1[j][0], acc1[j][1], x1, bf[j - H0]);
      }
    }
    return;
  }
  double pa0, pa1, pb[NACT], qa0, qa1, qb[NACT];
#define TC_LOAD(A0, A1, BF, KC) do { \
    A0 = a0p[2 * (KC)]; A1 = a1p[2 * (KC)]; \
    _Pragma("unroll") for (int j = 0; j < NACT; ++j) BF[j] = bp[(size_t)j * b_sb_stride + (KC) * 32]; } while (0)
#define TC_MMA(A0, A1, BF) do { \
    const double x0 = w0 * A0, x1 = w1 * A1; \
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
template <int NACT, int NW, bool PP>
__device__ __forceinline__ void tc_mma_one(double (&acc)[NW][2], const double *ap, double w,
                                           const double *bp, int kcs, int b_sb_stride)
{
  if constexpr (!PP) {
    for (int kc = 0; kc < kcs; ++kc) {
      const double x = w * ap[2 * kc];
      double bf[NACT];
#pragma unroll
      for (int j = 0; j < NACT; ++j) bf[j] = bp[(size_t)j * b_sb_stride + kc * 32];
#pragma unroll
      for (int j = 0; j < NACT; ++j) dmma884(acc[j][0], acc[j][1], x, bf[j]);
    }
    return;
  }
  double pa, pb[NACT], qa, qb[NACT];
#define TC_LOAD(A0, BF, KC) do { \
    A0 = ap[2 * (KC)]; \
    _Pragma("unroll") for (int j = 0; j < NACT; ++j) BF[j] = bp[(size_t)j * b_sb_stride + (KC) * 32]; } while (0)
#define TC_MMA(A0, BF) do { \
    const double x = w * A0; \
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

template <int NACT, int NW, bool PP>
__device__ __forceinline__ void tc_mma_row_smem(double (&acc0)[NW][2], double (&acc1)[NW][2],
                                                const double *a0p, const double *a1p, double w0, double w1,
                                                bool any0, bool any1, const double *bp, int kcs, int b_sb_stride)
{
  if (any0 && any1) tc_mma_both<NACT, NW, PP>(acc0, acc1, a0p, a1p, w0, w1, bp, kcs, b_sb_stride);
  else if (any0) tc_mma_one<NACT, NW, PP>(acc0, a0p, w0, bp, kcs, b_sb_stride);
  else tc_mma_one<NACT, NW, PP>(acc1, a1p, w1, bp, kcs, b_sb_stride);
}

// rows outside the staged window (sparse k3 regions): B fragments straight from the table in global
// memory; generic in the number of active sub-blocks (rare path, kept small)
template <int NW>
__device__ __forceinline__ void tc_mma_row_global(double (&acc0)[NW][2], double (&acc1)[NW][2],
                                                  const double *a0p, const double *a1p, double w0, double w1,
                                                  const double *bp, int kcs, int b_sb_stride, int nactw)
{
  for (int kc = 0; kc < kcs; ++kc) {
    const double a0 = w0 * a0p[2 * kc], a1 = w1 * a1p[2 * kc];
#pragma unroll
    for (int j = 0; j < NW; ++j)
      if (j < nactw) {
        const double bf = __ldg(bp + (size_t)j * b_sb_stride + kc * 32);
        dmma884(acc0[j][0], acc0[j][1], a0, bf);
        dmma884(acc1[j][0], acc1[j][1], a1, bf);
      }
  }
}

// All table rows of one stage for one consumer warp.  The units of a tile are k3-sorted, so the row
// index ix is non-decreasing with the unit: the rows of the warp's 16 units are [lo0, hi1] with group 0
// covering [lo0, hi0] and group 1 [lo1, hi1] (all four read as shared-memory broadcasts, warp-uniform).
// Dense k3 regions (a handful of rows per warp) step through the rows without any warp reduction or
// vote; sparse regions (rows far apart) find the next occupied row with a reduction.
constexpr int TC_DENSE_SPAN = 3;

template <int NACT, int NW, bool PP>
__device__ __forceinline__ void tc_stage_rows(double (&acc0)[NW][2], double (&acc1)[NW][2],
                                              const double *a0p, const double *a1p, int ix0, int ix1, double cw0, double cw1,
                                              int lo0, int hi0, int lo1, int hi1, int row_lo, int nrows,
                                              const double *bs, int b_row_doubles, const double *bg, int64_t jq_row,
                                              int kcs, int b_sb_stride)
{
  if (hi1 - lo0 <= TC_DENSE_SPAN) {
    for (int v = lo0; v <= hi1; ++v) {
      const bool any0 = (v <= hi0), any1 = (v >= lo1);
      if (!any0 && !any1) continue;
      const double w0 = (ix0 == v) ? cw0 : 0., w1 = (ix1 == v) ? cw1 : 0.;
      const int rb = v - row_lo;
      if (rb < nrows) tc_mma_row_smem<NACT, NW, PP>(acc0, acc1, a0p, a1p, w0, w1, any0, any1, bs + (size_t)rb * b_row_doubles, kcs, b_sb_stride);
      else tc_mma_row_global<NW>(acc0, acc1, a0p, a1p, w0, w1, bg + (int64_t)v * jq_row, kcs, b_sb_stride, NACT);
    }
  } else {
    int v = lo0;
    while (v != 0x7fffffff) {
      const bool m0 = (ix0 == v), m1 = (ix1 == v);
      const double w0 = m0 ? cw0 : 0., w1 = m1 ? cw1 : 0.;
      const int rb = v - row_lo;
      if (rb < nrows) tc_mma_row_smem<NACT, NW, PP>(acc0, acc1, a0p, a1p, w0, w1, v <= hi0, v >= lo1, bs + (size_t)rb * b_row_doubles, kcs, b_sb_stride);
      else tc_mma_row_global<NW>(acc0, acc1, a0p, a1p, w0, w1, bg + (int64_t)v * jq_row, kcs, b_sb_stride, NACT);
      const int n0 = (ix0 > v) ? ix0 : 0x7fffffff, n1 = (ix1 > v) ? ix1 : 0x7fffffff;
      v = __reduce_min_sync(0xffffffffu, min(n0, n1));
    }
  }
}

template <int WM, int WN, int NW, int NP>
__global__ void __launch_bounds__((WM * WN + NP) * 32, 1) los_tc_kernel(const __grid_constant__ TcParams P)
{
  constexpr int MT = 16 * WM;                    // units per tile
  constexpr int UPL = MT / (32 * NP);            // units per producer lane
  constexpr int TC_CONSUMERS = WM * WN * 32, TC_THREADS = TC_CONSUMERS + NP * 32;
  static_assert(MT % (32 * NP) == 0, "producer lanes own MT/(32 NP) units each");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NS = P.n_stages;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);             // full[8], empty[8]
  double *tau_d = reinterpret_cast<double *>(smem_raw + 128);           // tau0 - tau       [tau_pad]
  const int tau_pad = (P.tau_size + 15) & ~15;
  double *tau_w = tau_d + tau_pad;                                      // trapezoid weight [tau_pad]
  int4 *prow = reinterpret_cast<int4 *>(tau_w + tau_pad);               // producer: {row_lo, nrows, nact, row_hi} per tau
  double *col_scale_s = reinterpret_cast<double *>(prow + tau_pad);     // epilogue: 1/(16(2l+1)) per column [cols_pad]
  int *col_tt_s = reinterpret_cast<int *>(col_scale_s + P.cols_pad);    //           transfer type per column [cols_pad]
  unsigned char *stage0 = smem_raw + 128 + (size_t)tau_pad * 32 + (size_t)P.cols_pad * 12;   // 128-byte aligned

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (P.n_units + MT - 1) / MT;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&bars[s]), NP + 1);                    // one arrival per producer warp + the copy issuer's expect_tx
      mbar_init(smem_u32(&bars[8 + s]), TC_CONSUMERS / 32);     // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int t = tid; t < P.tau_size; t += TC_THREADS) { tau_d[t] = P.tau0_minus_tau[t]; tau_w[t] = P.delta_tau[t]; }
  for (int c = tid; c < P.n_cols_total; c += TC_THREADS) { col_scale_s[c] = P.col_scale[c]; col_tt_s[c] = P.col_tt[c]; }
  __syncthreads();

  // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  (the k3-sorted sequence is sampled
  // with a stride, so every CTA sees the same mix of cheap low-k and expensive high-k tiles).  The
  // ring of stages runs across tile boundaries: the producer warp is already filling the stages of
  // the next tile while the consumers drain and store the current one.
  int s = 0; uint32_t ph = 0;                   // ring slot and its phase parity, advanced per stage
#define TC_ADVANCE() do { if (++s == NS) { s = 0; ph ^= 1u; } } while (0)

  if (warp >= TC_CONSUMERS / 32) {
    // =========================== producer warps ===========================
    // NP warps share the per-unit header work of a stage (row index and quadrature x lerp weights of every
    // unit: lane <-> units pw*32 + lane + 32*NP*i); lane 0 of the first one also issues the copies.  With the
    // uniform xx grid of bessel2_get_xx_list nothing here waits on global memory: per unit and stage
    //   x = k3 (tau0-tau),  ix = (int)(x/dx),  a_J = (xx[ix+1] - x)/dx,  c0 = w a_J,  c1 = w - c0
    // (transfer2.c:2195-2202, xx[ix+1] recomputed as (ix+1) delta).  Stage sequence per tile: pass, K chunk, tau.
    const int pw = warp - TC_CONSUMERS / 32;
    // (dbg 64 / 128: the reversed and the neutral policies, for A/B timing only)
    const uint64_t pol_stream = (P.dbg & 128) ? l2_policy_evict_normal() : (P.dbg & 64) ? l2_policy_evict_last() : l2_policy_evict_first();   // slab: read once
    const uint64_t pol_keep = (P.dbg & 128) ? l2_policy_evict_normal() : (P.dbg & 64) ? l2_policy_evict_first() : l2_policy_evict_last();     // table rows: shared by the wave
    bool first_round = true;                    // the first NS stages need no empty-slot wait
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t s0 = tile * MT;
      const int64_t nleft = P.n_units - s0;
      const int nvalid = nleft < (int64_t)MT ? (int)nleft : MT;
      double kk[UPL];
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        const int n = pw * 32 + lane + 32 * NP * i;
        const int64_t sidx = s0 + (n < nvalid ? n : nvalid - 1);      // tail: replicate the last unit, never stored
        kk[i] = __ldg(P.k_sorted + sidx);
      }
      double k_first = 0., k_last = 0.;
      if (pw == 0) { k_first = __ldg(P.k_sorted + s0); k_last = __ldg(P.k_sorted + s0 + nvalid - 1); }
      // x, row index and xx[ix+1] of the NEXT stage are prepared one stage ahead: when the xx grid is not
      // i * delta bit for bit, xx[ix+1] is a table read whose latency then overlaps the wait for a free slot
      int ix_n[UPL]; double xxn_n[UPL], x_n[UPL]; bool bad_n[UPL];
#define TC_PREFETCH(T) do { \
        const double d_ = tau_d[T]; \
        _Pragma("unroll") for (int i = 0; i < UPL; ++i) { \
          const double x_ = kk[i] * d_;                                 /* transfer2.c:2195 */ \
          int ix_ = tc_index_x(x_, P.inv_xx_step, P.xx_step_ref);       /* transfer2.c:2199 */ \
          bad_n[i] = (ix_ < 0 || ix_ + 1 >= P.xx_size); \
          if (bad_n[i]) { atomicOr(P.err, ERR_X_OUT_OF_RANGE); ix_ = max(0, min(ix_, P.xx_size - 2)); } \
          x_n[i] = x_; ix_n[i] = ix_; \
          xxn_n[i] = P.xx_uniform ? (double)(ix_ + 1) * P.xx_step : __ldg(P.xx + ix_ + 1); } } while (0)
      TC_PREFETCH(0);
      for (int ip = P.pass_begin; ip < P.pass_end; ++ip) {
        const TcPass ps = P.pass[ip];
        if (pw == 0) {
          // Table rows of the tile at every tau (units are k3-sorted: rows span [row(first), row(last)]) and the
          // active column prefix of this pass there: one parallel round of global reads per (tile, pass), so that
          // no table read sits on the per-stage critical path of the copy issuer.
          const int *nact_b = P.nact + (int64_t)ps.blk * P.xx_size;
          __syncwarp();
          for (int t = lane; t < P.tau_size; t += 32) {
            const double d = tau_d[t];
            int row_lo = tc_index_x(k_first * d, P.inv_xx_step, P.xx_step_ref), row_hi = tc_index_x(k_last * d, P.inv_xx_step, P.xx_step_ref);
            row_lo = max(0, min(row_lo, P.xx_size - 2)); row_hi = max(row_lo, min(row_hi, P.xx_size - 2));
            const int nact = max(0, min(__ldg(nact_b + row_hi) - ps.sb0, ps.nsb));
            prow[t] = make_int4(row_lo, min(row_hi - row_lo + 1, P.rmax), nact, row_hi);
          }
          __syncwarp();
        }
        for (int q = (P.q_fixed >= 0 ? P.q_fixed : 0); q < (P.q_fixed >= 0 ? P.q_fixed + 1 : ps.nq); ++q) {
          const TcGroup G = P.grp[ps.g0 + q];
          const uint32_t bytes_per_sb = (uint32_t)G.kcs * 256u;
          const uint32_t s_bytes = (uint32_t)MT * G.tpp * 8u;
          const int64_t q_off = ps.jq_off + (int64_t)ps.nsb * G.kc0 * 32;
          const CUtensorMap *tmap = &P.tmap[ps.g0 + q];
          for (int t = 0; t < P.tau_size; ++t) {
            const uint32_t full = smem_u32(&bars[s]), empty = smem_u32(&bars[8 + s]);
            const TcStagePtrs st = tc_stage<MT>(stage0, s, P);
            const double w = tau_w[t];
            int ix[UPL]; double c0[UPL], c1[UPL];
#pragma unroll
            for (int i = 0; i < UPL; ++i) {
              ix[i] = ix_n[i];
              const double a_J = bad_n[i] ? 0. : (xxn_n[i] - x_n[i]) * P.inv_xx_step;     // transfer2.c:2202
              c0[i] = w * a_J; c1[i] = w - c0[i];                             // weights of J[ix], J[ix+1]
            }
            { const int tn = (t + 1 < P.tau_size) ? t + 1 : 0; TC_PREFETCH(tn); }   // header inputs of the next stage
            if (!first_round) mbar_wait(empty, ph ^ 1u);
            if (pw == 0 && lane == 0) {
              const int4 pr = prow[t];
              const int nrows = pr.y;
              const uint32_t bytes_row = (uint32_t)pr.z * bytes_per_sb;
              *reinterpret_cast<int4 *>(st.meta) = pr;
              mbar_arrive_expect_tx(full, ((P.dbg & 8) ? 0u : s_bytes) + ((P.dbg & 16) ? 0u : bytes_row * nrows));
              if (!(P.dbg & 8)) tma_load_3d(smem_u32(st.S), tmap, 0, t, (int)s0, full, pol_stream);
              const double *src = P.Jq + (int64_t)pr.x * P.jq_row + q_off;
              if (bytes_row > 0 && !(P.dbg & 16))
                for (int r = 0; r < nrows; ++r)
                  bulk_g2s(smem_u32(st.B) + r * P.b_row_bytes, src + (int64_t)r * P.jq_row, bytes_row, full, pol_keep);
            }
#pragma unroll
            for (int i = 0; i < UPL; ++i) {
              const int n = pw * 32 + lane + 32 * NP * i;
              st.ixs[n] = ix[i];
              st.cc[n] = make_double2(c0[i], c1[i]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full);
            if (++s == NS) { s = 0; ph ^= 1u; first_round = false; }
          }
        }
      }
#undef TC_PREFETCH
    }
    return;
  }

  // =========================== consumers ===========================
  const int warp_m = warp / WN, warp_n = (warp + warp_m) % WN;      // rotate column groups over the SM sub-partitions
  const int g0 = warp_m * 2, g1 = g0 + 1;      // the two 8-unit groups of this warp
  const int um = lane >> 2;                    // unit inside a group held by this lane (A and C rows)
  const int ksel = lane & 1, tsel = (lane & 3) >> 1;   // weight {c0,c1} and term of the kc held by this lane
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // original (pair-major) index of this lane's two units, read where it is needed (start of a continued chunk, epilogue):
    // a tile lasts tens of thousands of cycles, two scattered loads per lane are free, and no shared buffer has to
    // outlive the ring (a producer running ahead could otherwise overwrite it when a tile has fewer stages than the ring)
    for (int ip = P.pass_begin; ip < P.pass_end; ++ip) {
      const TcPass ps = P.pass[ip];
      const TcBlock B = P.blk[ps.blk];
      double acc0[NW][2], acc1[NW][2];
#pragma unroll
      for (int j = 0; j < NW; ++j) { acc0[j][0] = acc0[j][1] = 0.; acc1[j][0] = acc1[j][1] = 0.; }
      if (P.acc_in) {
        // later K chunk of the block: continue from the unscaled partial sums the previous launch left in `out`
        const int cpair = 2 * (lane & 3);
        const int64_t su0 = tile * MT + g0 * 8 + um, su1 = su0 + 8;
        const int u0 = (su0 < P.n_units) ? __ldg(P.perm + su0) : -1, u1 = (su1 < P.n_units) ? __ldg(P.perm + su1) : -1;
#pragma unroll
        for (int j = 0; j < NW; ++j) {
          const int sb = warp_n + j * WN;
          if (sb >= ps.nsb) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = (ps.sb0 + sb) * 8 + cpair + h;
            if (c >= B.ncols) continue;
            const int tt = col_tt_s[B.col_off + c];
            if (u0 >= 0) acc0[j][h] = P.out[(int64_t)tt * P.out_stride + u0];
            if (u1 >= 0) acc1[j][h] = P.out[(int64_t)tt * P.out_stride + u1];
          }
        }
      }

      for (int q = (P.q_fixed >= 0 ? P.q_fixed : 0); q < (P.q_fixed >= 0 ? P.q_fixed + 1 : ps.nq); ++q) {
        const TcGroup G = P.grp[ps.g0 + q];
        const int kcs_q = G.kcs;
        const int b_sb_stride = WN * kcs_q * 32;
        const int a_off0 = (g0 * 8 + um) * G.tpp + tsel, a_off1 = (g1 * 8 + um) * G.tpp + tsel;
        // The header of the NEXT stage (row window, row indices and weights of this lane's units) is fetched
        // before the MMAs of the current one whenever that stage has already landed, so that its barrier
        // test and shared-memory latencies are covered by the DMMA stream.
        // (register budget: only the (12,1,7) grid has room for the second header without spilling in the MMA loop;
        //  measured 7.05 -> 6.84 ms there, slower elsewhere)
        constexpr bool PREFETCH_HDR = (WM == 12 && WN == 1 && NW == 7);
        bool have_next = false;
        int4 n_meta = make_int4(0, 0, 0, 0);
        int n_ix0 = 0, n_ix1 = 0, n_lo0 = 0, n_hi0 = 0, n_lo1 = 0, n_hi1 = 0;
        double n_cw0 = 0., n_cw1 = 0.;
#define TC_LOAD_HDR(ST, META, IX0, IX1, LO0, HI0, LO1, HI1, CW0, CW1) do { \
          META = *reinterpret_cast<const int4 *>((ST).meta);      /* {row_lo, nrows, nact, row_hi} */ \
          IX0 = (ST).ixs[g0 * 8 + um]; IX1 = (ST).ixs[g1 * 8 + um]; \
          LO0 = (ST).ixs[g0 * 8]; HI0 = (ST).ixs[g0 * 8 + 7]; LO1 = (ST).ixs[g1 * 8]; HI1 = (ST).ixs[g1 * 8 + 7]; \
          const double2 cc0_ = (ST).cc[g0 * 8 + um], cc1_ = (ST).cc[g1 * 8 + um]; \
          CW0 = ksel ? cc0_.y : cc0_.x; CW1 = ksel ? cc1_.y : cc1_.x; } while (0)
        for (int t = 0; t < P.tau_size; ++t) {
          const TcStagePtrs st = tc_stage<MT>(stage0, s, P);
          int4 meta; int ix0, ix1, lo0, hi0, lo1, hi1; double cw0, cw1;
          if (have_next) {
            meta = n_meta; ix0 = n_ix0; ix1 = n_ix1; lo0 = n_lo0; hi0 = n_hi0; lo1 = n_lo1; hi1 = n_hi1; cw0 = n_cw0; cw1 = n_cw1;
          } else {
            mbar_wait(smem_u32(&bars[s]), ph);
            TC_LOAD_HDR(st, meta, ix0, ix1, lo0, hi0, lo1, hi1, cw0, cw1);
          }
          have_next = false;
          if (PREFETCH_HDR && t + 1 < P.tau_size && !(P.dbg & 32)) {
            const int s1 = (s + 1 == NS) ? 0 : s + 1;
            const uint32_t ph1 = (s + 1 == NS) ? (ph ^ 1u) : ph;
            if (__all_sync(0xffffffffu, mbar_test(smem_u32(&bars[s1]), ph1))) {
              const TcStagePtrs st1 = tc_stage<MT>(stage0, s1, P);
              TC_LOAD_HDR(st1, n_meta, n_ix0, n_ix1, n_lo0, n_hi0, n_lo1, n_hi1, n_cw0, n_cw1);
              have_next = true;
            }
          }
          const int nact = meta.z;
          // owned sub-blocks: warp_n, warp_n + WN, ... ; active ones form a prefix of that list
          const int nactw = (nact > warp_n) ? min(NW, (nact - warp_n + WN - 1) / WN) : 0;
          if (nactw > 0 && !(P.dbg & 4)) {
            // A fragment of kc: S[unit][2 kc + tsel] x weight
            const double *a0p = st.S + a_off0;
            const double *a1p = st.S + a_off1;
            const double *bs = st.B + (size_t)warp_n * kcs_q * 32 + lane;
            const double *bg = P.Jq + ps.jq_off + (int64_t)ps.nsb * G.kc0 * 32 + (size_t)warp_n * kcs_q * 32 + lane;
#define TC_CASE(N) case N: if constexpr (N <= NW) tc_stage_rows<N, NW, (WM * WN <= 8)>(acc0, acc1, a0p, a1p, ix0, ix1, cw0, cw1, lo0, hi0, lo1, hi1, \
                                       meta.x, meta.y, bs, P.b_row_bytes / 8, bg, P.jq_row, kcs_q, b_sb_stride); break;
            switch (nactw) {
              TC_CASE(1) TC_CASE(2) TC_CASE(3) TC_CASE(4) TC_CASE(5) TC_CASE(6) TC_CASE(7) TC_CASE(8) TC_CASE(9) TC_CASE(10)
            }
#undef TC_CASE
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars[8 + s]));
          TC_ADVANCE();
        }
#undef TC_LOAD_HDR
      }

      // ---- epilogue: the 1/2 of the trapezoid and the divisions by 4, (2l+1), 2 of
      //      transfer2.c:2092-2098 as one multiplication by 1 / (16 (2l+1)) ----
      const int cpair = 2 * (lane & 3);
      const int64_t su0 = tile * MT + g0 * 8 + um, su1 = su0 + 8;
      const int u0 = (su0 < P.n_units) ? __ldg(P.perm + su0) : -1, u1 = (su1 < P.n_units) ? __ldg(P.perm + su1) : -1;
#pragma unroll
      for (int j = 0; j < NW; ++j) {
        const int sb = warp_n + j * WN;
        if (sb >= ps.nsb) continue;
        const int cbase = (ps.sb0 + sb) * 8 + cpair;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = cbase + h;
          if (c >= B.ncols) continue;
          const int tt = col_tt_s[B.col_off + c];
          const double scale = col_scale_s[B.col_off + c];
          const double t0 = P.acc_final ? acc0[j][h] * scale : acc0[j][h], t1 = P.acc_final ? acc1[j][h] * scale : acc1[j][h];
          if (u0 >= 0) { if (t0 != t0) atomicOr(P.err, ERR_NAN); P.out[(int64_t)tt * P.out_stride + u0] = t0; }
          if (u1 >= 0) { if (t1 != t1) atomicOr(P.err, ERR_NAN); P.out[(int64_t)tt * P.out_stride + u1] = t1; }
        }
      }
    }
  }
#undef TC_ADVANCE
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
  int n_blocks = 0, n_groups = 0, n_pass = 0;
  TcBlock blk[TC_MAX_BLOCKS];
  TcGroup grp[TC_MAX_GROUPS];
  TcPass pass[TC_MAX_PASSES];
  int64_t jq_row = 0;
  int64_t s_doubles_per_unit = 0;        // sum over groups of tau_size * tpp
  int wm = 8, wn = 1, nw = 8;
  int kcs_max = 1, nsb_pass_max = 1, tpp_max = 4, n_cols_total = 0, tau_size = 0;
  int blk_tt_lo[TC_MAX_BLOCKS] = {0}, blk_tt_n[TC_MAX_BLOCKS] = {0};   // rows of `out` a block owns when they are one run (else n = 0)
  Buf<double> Jq, keys_out, S, slot_sign, col_scale;
  Buf<double2> CSD;
  Buf<int> col_tt, col_l, col_lm, term_col0, nact, perm_in, perm, unit_pair, slot_tp, jcount;
  Buf<TcUnitW> uw; Buf<int2> ujp;
  Buf<TcSplineCoef> spl_coef; Buf<TcSplineEnds> spl_ends;
  Buf<unsigned char> cub_tmp;
  TcParams tp{};                 // kernel parameter block (about 3 KB: kept off the stack)
  TcInterp interp{};
  PFN_cuTensorMapEncodeTiled encode = nullptr;
  cudaStream_t st2 = nullptr; cudaEvent_t ev_join = nullptr;     // second stream: the k3 sort runs beside the splines
};

FastState *fast_state_create() { return new FastState(); }

void fast_state_destroy(FastState *f)
{
  if (!f) return;
  f->Jq.release(); f->keys_out.release(); f->S.release(); f->slot_sign.release(); f->col_scale.release(); f->CSD.release();
  f->col_tt.release(); f->col_l.release(); f->col_lm.release(); f->term_col0.release(); f->nact.release();
  f->perm_in.release(); f->perm.release(); f->unit_pair.release(); f->slot_tp.release();
  f->uw.release(); f->ujp.release(); f->cub_tmp.release(); f->jcount.release(); f->spl_coef.release(); f->spl_ends.release();
  if (f->st2) cudaStreamDestroy(f->st2);
  if (f->ev_join) cudaEventDestroy(f->ev_join);
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
  std::vector<int> col_tt, col_l, col_lm, term_col0, term_tp;
  std::vector<double> term_sign;
  fs->n_blocks = 0;
  int max_sb = 0;
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
      B.term_off = (int)term_tp.size();
      for (int il = 0; il < h.l_size; ++il) {
        int lm = h.lm_array[(size_t)il * h.m_size + im];
        if (lm < 0) continue;
        col_tt.push_back(tt_base + lm); col_l.push_back(h.l[il]); col_lm.push_back(lm);
      }
      B.ncols = (int)col_tt.size() - B.col_off;
      if (B.ncols == 0) continue;
      {
        bool run = true;
        for (int c = 1; c < B.ncols; ++c) run = run && (col_tt[B.col_off + c] == col_tt[B.col_off + c - 1] + 1);
        fs->blk_tt_lo[fs->n_blocks] = col_tt[B.col_off];
        fs->blk_tt_n[fs->n_blocks] = run ? B.ncols : 0;
      }
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
      if (B.nterms > TC_MAX_TERMS) { err = "too many K terms for the FAST kernel"; return 1; }
      if (B.nterms == 0) {                      // nothing contributes: outputs are exactly zero
        term_tp.push_back(0); term_sign.push_back(0.); term_col0.push_back(0); B.nterms = 1;
      }
      B.nkc = (B.nterms + 1) / 2;
      B.nsb = (B.ncols + 7) / 8;
      max_sb = std::max(max_sb, B.nsb);
      fs->blk[fs->n_blocks++] = B;
    }
  }

  // warp grid (WM x WN consumer warps, NW sub-blocks per warp): widest tile of units that still
  // covers the columns of a block in one pass
  // (measured on B200: sn_t_10percent 7.0 ms with (12,1,8) vs 7.3 ms with (8,1,8); sn_pol_10percent 14.8 ms with
  //  (12,1,10) vs 16.0 ms with (6,2,5) and 16.6 ms with (8,1,10);
  //  sn_pol_5percent_lmax3000 357 ms with (4,3,7) vs 367 ms with (2,4,8) on a 1.5 M k3 sample; 192-unit tiles in three
  //  column passes, (12,1,7), won on a uniformly sampled 3.7 M k3 batch (616 vs 648 ms) but lost on the full run, whose
  //  batches are consecutive k1 levels (1448 vs 1387 ms): not adopted)
  if (max_sb <= 7) { fs->wm = 12; fs->wn = 1; fs->nw = 7; }     // three consumer warps per SM sub-partition
  else if (max_sb <= 8) { fs->wm = 12; fs->wn = 1; fs->nw = 8; }
  else if (max_sb <= 10) { fs->wm = 12; fs->wn = 1; fs->nw = 10; }
  else if (max_sb <= 16) { fs->wm = 4; fs->wn = 2; fs->nw = 8; }
  else if (max_sb <= 20) { fs->wm = 4; fs->wn = 2; fs->nw = 10; }
  else if (max_sb <= 21) { fs->wm = 12; fs->wn = 1; fs->nw = 7; }    // three column passes of 7: 1367 ms vs 1411 ms with (4,3,7) in one pass
                                                                      // on the full lmax3000 run, 287 vs 326 ms on its 1/8 shard (round 2)
  else if (max_sb <= 32) { fs->wm = 4; fs->wn = 2; fs->nw = 8; }     // two column passes: 731 ms vs 855 ms with (2,4,8) on the lmax4000 shard
  else { fs->wm = 2; fs->wn = 4; fs->nw = 10; }
  if (const char *v = getenv("SONG_LOS_TC_SHAPE")) {          // A/B measurements: "wm,wn,nw"
    int a, b, c;
    if (sscanf(v, "%d,%d,%d", &a, &b, &c) == 3) {
      fs->wm = a; fs->wn = b; fs->nw = c;
    }
  }
  const int sb_per_pass = fs->wn * fs->nw;
  int kcs_cap = TC_KCS_MAX;
  if (const char *v = getenv("SONG_LOS_TC_KCS")) kcs_cap = std::max(1, std::min(TC_KCS_MAX, atoi(v)));

  // K chunks (groups) of every block, column passes, and the row layout of the table
  fs->n_groups = 0; fs->n_pass = 0;
  int64_t off = 0, s_per_unit = 0;
  fs->kcs_max = 1; fs->nsb_pass_max = 1; fs->tpp_max = 4;
  std::vector<int> slot_tp((size_t)TC_MAX_GROUPS * TC_SLOTS, -1);
  std::vector<double> slot_sign((size_t)TC_MAX_GROUPS * TC_SLOTS, 0.);
  std::vector<int> blk_kcs(fs->n_blocks);
  for (int ib = 0; ib < fs->n_blocks; ++ib) {
    const TcBlock &B = fs->blk[ib];
    int nq = (B.nkc + kcs_cap - 1) / kcs_cap;
    int kcs = (B.nkc + nq - 1) / nq;
    nq = (B.nkc + kcs - 1) / kcs;
    blk_kcs[ib] = kcs;
    const int g_first = fs->n_groups;
    for (int q = 0; q < nq; ++q) {
      if (fs->n_groups >= TC_MAX_GROUPS) { err = "too many K chunks for the FAST kernel"; return 1; }
      TcGroup G{};
      G.blk = ib; G.kc0 = q * kcs; G.kcs = std::min(kcs, B.nkc - G.kc0);
      // row of 2*kcs doubles (16-byte multiples: TMA stride rule).  Odd kcs: conflict-free A-fragment reads; even kcs has
      // 2- to 4-way conflicts there, but padding the rows by a pair cost more in slab traffic than it gained (measured)
      G.tpp = 2 * G.kcs;
      for (int sl = 0; sl < 2 * G.kcs; ++sl) {
        int e = 2 * G.kc0 + sl;
        if (e < B.nterms) { slot_tp[(size_t)fs->n_groups * TC_SLOTS + sl] = term_tp[B.term_off + e];
                            slot_sign[(size_t)fs->n_groups * TC_SLOTS + sl] = term_sign[B.term_off + e]; }
      }
      s_per_unit += (int64_t)h.tau_size * G.tpp;
      fs->kcs_max = std::max(fs->kcs_max, G.kcs);
      fs->tpp_max = std::max(fs->tpp_max, G.tpp);
      fs->grp[fs->n_groups++] = G;
    }
    int npass_b = (B.nsb + sb_per_pass - 1) / sb_per_pass;
    int per = (B.nsb + npass_b - 1) / npass_b;                // balanced passes
    for (int sb0 = 0; sb0 < B.nsb; sb0 += per) {
      if (fs->n_pass >= TC_MAX_PASSES) { err = "too many column passes for the FAST kernel"; return 1; }
      TcPass ps{};
      ps.blk = ib; ps.sb0 = sb0; ps.nsb = std::min(per, B.nsb - sb0);
      ps.g0 = g_first; ps.nq = nq;
      ps.jq_off = off;
      off += (int64_t)ps.nsb * B.nkc * 32;
      fs->nsb_pass_max = std::max(fs->nsb_pass_max, ps.nsb);
      fs->pass[fs->n_pass++] = ps;
    }
  }
  fs->jq_row = off;
  fs->tau_size = h.tau_size;
  fs->s_doubles_per_unit = s_per_unit;

  auto up = [&](auto &buf, const auto &vec) -> cudaError_t {
    cudaError_t e = buf.reserve(vec.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice, st);
  };
  fs->n_cols_total = (int)col_tt.size();
  std::vector<double> col_scale(col_l.size());
  for (size_t i = 0; i < col_l.size(); ++i) col_scale[i] = 1.0 / (16.0 * (2.0 * col_l[i] + 1.0));
  FCK(up(fs->col_tt, col_tt)); FCK(up(fs->col_l, col_l)); FCK(up(fs->col_lm, col_lm)); FCK(up(fs->col_scale, col_scale));
  FCK(up(fs->term_col0, term_col0)); FCK(up(fs->slot_tp, slot_tp)); FCK(up(fs->slot_sign, slot_sign));
  FCK(fs->Jq.reserve((size_t)h.xx_size * fs->jq_row));
  for (int ip = 0; ip < fs->n_pass; ++ip) {
    const TcPass &ps = fs->pass[ip];
    const TcBlock &B = fs->blk[ps.blk];
    int64_t n = (int64_t)B.nkc * ps.nsb * 8;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)h.xx_size);
    tc_build_table_kernel<<<grid, 128, 0, st>>>(P.proj.J, P.proj.n_cols, fs->jq_row, ps, B, blk_kcs[ps.blk],
                                                fs->term_col0.p, fs->col_lm.p, fs->Jq.p);
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

  if (!fs->encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FCK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { err = "cuTensorMapEncodeTiled is not available in this driver"; return 1; }
    fs->encode = (PFN_cuTensorMapEncodeTiled)fn;
  }
  fs->table_ready = true;
  return 0;
}

template <int WM, int WN, int NW, int NP>
static cudaError_t launch_tc(const TcParams &TP, int64_t n_ctas, size_t smem, cudaStream_t st)
{
  cudaError_t e = cudaFuncSetAttribute(los_tc_kernel<WM, WN, NW, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  los_tc_kernel<WM, WN, NW, NP><<<(unsigned)n_ctas, (WM * WN + NP) * 32, smem, st>>>(TP);
  return cudaGetLastError();
}

int fast_prepare(FastState *fs, const FastHostLayout &h, const ExactParams &P, cudaStream_t st, std::string &err)
{
  if (P.cubic_J) { err = "FAST mode supports linear Bessel interpolation only; use SONG_LOS_MODE_EXACT"; return 1; }
  if (!fs->table_ready && build_layout(fs, h, P, st, err)) return 1;
  return 0;
}

// device bytes the FAST path needs per k3 unit beyond the batch inputs and the result: the source
// slab S plus the per-unit sort / weight arrays
int64_t fast_bytes_per_unit(const FastState *fs)
{
  return (int64_t)fs->tau_size * fs->tpp_max * 8 + 8 /*keys*/ + 3 * 4 /*perm_in, perm, unit_pair*/ + (int64_t)sizeof(TcUnitW) + 8 /*ujp*/ + 16 /*cub*/;
}

// bytes of per-batch intermediates the FAST path already holds: reused by the next batch, so they count as available
// when the capacity of a batch is computed (song_los_max_batch_units)
int64_t fast_held_bytes(const FastState *fs)
{
  if (!fs) return 0;
  size_t b = fs->S.cap * sizeof(double) + fs->CSD.cap * sizeof(double2) + fs->keys_out.cap * sizeof(double);
  b += (fs->perm_in.cap + fs->perm.cap + fs->unit_pair.cap + fs->jcount.cap) * sizeof(int);
  b += fs->uw.cap * sizeof(TcUnitW) + fs->ujp.cap * sizeof(int2) + fs->cub_tmp.cap;
  b += fs->spl_coef.cap * sizeof(TcSplineCoef) + fs->spl_ends.cap * sizeof(TcSplineEnds);
  return (int64_t)b;
}

// Allocates every intermediate of the FAST path for batches of up to U k3 values / n_src source doubles /
// n_pt coarse k3 points, so that no (device-synchronising) allocation happens between pipelined batches.
int fast_reserve(FastState *fs, int64_t U, int64_t n_src, int64_t n_pt, std::string &err)
{
  if (!fs->table_ready) { err = "fast_reserve: layout not built"; return 1; }
  FCK(fs->CSD.reserve((size_t)n_src));
  FCK(fs->perm_in.reserve((size_t)U)); FCK(fs->perm.reserve((size_t)U)); FCK(fs->keys_out.reserve((size_t)U));
  FCK(fs->unit_pair.reserve((size_t)U));
  size_t tmp_bytes = 0;
  FCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const double *)nullptr, fs->keys_out.p, fs->perm_in.p, fs->perm.p,
                                      (int)U, 0, 64, (cudaStream_t)0));
  FCK(fs->cub_tmp.reserve(tmp_bytes));
  FCK(fs->uw.reserve((size_t)U)); FCK(fs->ujp.reserve((size_t)U));
  FCK(fs->S.reserve((size_t)U * fs->tau_size * fs->tpp_max));
  FCK(fs->jcount.reserve((size_t)n_pt));
  FCK(fs->spl_coef.reserve((size_t)n_pt)); FCK(fs->spl_ends.reserve((size_t)n_pt));    // n_pairs <= n_pt / 2
  return 0;
}

// Enqueues the whole FAST pipeline on `st` and returns without waiting; ev[0..4] (caller-owned) are
// recorded at the phase boundaries and read later by fast_collect.
int fast_enqueue(FastState *fs, const FastHostLayout &h, const ExactParams &P, int device_sms, cudaStream_t st,
                 cudaEvent_t *ev, int *n_launches, int *n_chunks_out, FastEarlyOut *early, std::string &err)
{
  if (early) early->n_ranges = 0;
  if (fast_prepare(fs, h, P, st, err)) return 1;
  const Batch &b = P.b;
  const int64_t U = b.n_units;
  const int tau_size = P.tm.tau_size, tp2 = P.ty.tp2_size;
  const int64_t rowlen = (int64_t)tau_size * tp2;
  if (U >= (int64_t)1 << 31) { err = "more than 2^31 k3 units in one batch"; return 1; }
  // ---- tile shape and row window of this batch ----
  // The k3-sorted units of a batch sweep `rows_spanned` table rows, rho = U / rows_spanned of them per row.  A tile of MT
  // units needs MT/rho + 1 rows per stage; rows beyond the `rmax` staged in shared memory are read from L2 by the
  // consumer warps (slow).  Dense batches (the whole run on one GPU: rho ~ 90) want the widest tile and 4 rows; the
  // smaller batches of a pipelined or multi-GPU run want more rows per stage, or a narrower tile when the widest one
  // cannot hold its rows with a ring of 3 stages.  All shapes of a family read the same table layout (one column pass of
  // at most 8 sub-blocks), so the choice is made per batch.  Measured on configs[1] (los_tc ms, default -> chosen):
  // 800 k units 4.63 -> 3.91, 400 k 3.44 -> 2.59, 200 k 2.61 -> 2.48; lmax3000 1/4 shard 416 -> 376 (tests/big/ab_knobs.py).
  int wm = fs->wm, wn = fs->wn, nw = fs->nw, rmax_sel = 0;
  {
    const double rho = (h.rows_spanned > 1. && U > 0) ? (double)U / h.rows_spanned : 1e9;
    const size_t fixed0 = 128 + (size_t)((tau_size + 15) & ~15) * 32 + (size_t)((fs->n_cols_total + 31) & ~31) * 12;
    auto r128b = [](size_t x) { return (x + 127) / 128 * 128; };
    auto cap_of = [&](int mt) {        // most rows per stage that still leave a ring of 3 stages
      int cap = 0;
      for (int r = 1; r <= 8; ++r) {
        const size_t stage = r128b(16 + (size_t)mt * 20) + r128b((size_t)mt * fs->tpp_max * 8) + (size_t)r * fs->nsb_pass_max * fs->kcs_max * 256;
        if ((224 * 1024 - fixed0) / stage >= 3) cap = r;
      }
      return cap;
    };
    struct Shape { int wm, wn, nw; };
    std::vector<Shape> family{{fs->wm, fs->wn, fs->nw}};
    // (blocks covered in several column passes keep their shape: narrower tiles were slower there, 361 vs 287 ms on the
    //  lmax3000 1/8 shard; only the row window adapts)
    const bool narrow_ok = (fs->wm == 12 && fs->wn == 1 && fs->nsb_pass_max <= 8 && fs->n_pass == fs->n_blocks && !getenv("SONG_LOS_TC_SHAPE"));
    if (narrow_ok) { family.push_back({8, 1, 8}); family.push_back({4, 2, 8}); }
    bool chosen = false;
    for (const Shape &sh : family) {
      const int mt = 16 * sh.wm, cap = cap_of(mt);
      const double need = (double)mt / rho;
      if (need + 0.5 <= (double)cap) {
        const int dflt = (mt >= 128) ? 4 : (mt >= 64 ? 3 : 2);
        wm = sh.wm; wn = sh.wn; nw = sh.nw;
        rmax_sel = std::max(std::min(dflt, cap), std::min(cap, (int)std::ceil(need + 1.5)));
        chosen = true;
        break;
      }
    }
    if (!chosen && narrow_ok) { wm = 8; wn = 1; nw = 8; }     // very sparse batches: every shape reads most rows from L2
    if (getenv("SONG_LOS_TC_NO_ADAPT")) { wm = fs->wm; wn = fs->wn; nw = fs->nw; rmax_sel = 0; }
  }
  const int MT = 16 * wm;
  const int64_t n_tiles = (U + MT - 1) / MT;

  // ---- phase 0 and 1 side by side: the global k3 order (second stream) does not depend on the splines ----
  FCK(cudaEventRecord(ev[0], st));
  if (!fs->st2) { FCK(cudaStreamCreateWithFlags(&fs->st2, cudaStreamNonBlocking)); FCK(cudaEventCreateWithFlags(&fs->ev_join, cudaEventDisableTiming)); }
  FCK(fs->CSD.reserve((size_t)h.n_src_total));
  FCK(fs->perm_in.reserve((size_t)U)); FCK(fs->perm.reserve((size_t)U)); FCK(fs->keys_out.reserve((size_t)U));
  FCK(fs->unit_pair.reserve((size_t)U));
  if (U > 0) {
    cudaStream_t s2 = fs->st2;
    FCK(cudaStreamWaitEvent(s2, ev[0], 0));
    tc_unit_pair_kernel<<<(unsigned)b.n_pairs, 128, 0, s2>>>(b, fs->unit_pair.p);
    FCK(cudaGetLastError());
    tc_iota_kernel<<<(unsigned)((U + 255) / 256), 256, 0, s2>>>(fs->perm_in.p, U);
    FCK(cudaGetLastError());
    size_t tmp_bytes = 0;
    FCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, b.k3_tr, fs->keys_out.p, fs->perm_in.p, fs->perm.p,
                                        (int)U, 0, 64, s2));
    FCK(fs->cub_tmp.reserve(tmp_bytes));
    FCK(cub::DeviceRadixSort::SortPairs(fs->cub_tmp.p, tmp_bytes, b.k3_tr, fs->keys_out.p, fs->perm_in.p, fs->perm.p,
                                        (int)U, 0, 64, s2));
    FCK(cudaEventRecord(fs->ev_join, s2));
    *n_launches += 2 + 10;                 // unit_pair, iota + cub (histogram, scan, 8 onesweep passes over 64-bit keys)
  }
  // ---- phase 0: transposed {y, dd} sources ----
  {
    int64_t cols = rowlen * b.n_pairs;
    if (cols > 0 && P.cubic_k3) {
      FCK(fs->spl_coef.reserve((size_t)h.n_pt_total)); FCK(fs->spl_ends.reserve((size_t)b.n_pairs));
      tc_spline_coef_kernel<<<(unsigned)((b.n_pairs + 63) / 64), 64, 0, st>>>(b, fs->spl_coef.p, fs->spl_ends.p, P.err);
      FCK(cudaGetLastError());
      tc_spline_csd_coef_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(b, tau_size, tp2, fs->spl_coef.p, fs->spl_ends.p, fs->CSD.p);
      FCK(cudaGetLastError());
      *n_launches += 2;
    } else if (cols > 0) {
      tc_transpose_csd_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(b, tau_size, tp2, fs->CSD.p);
      FCK(cudaGetLastError());
      ++*n_launches;
    }
  }
  FCK(cudaEventRecord(ev[1], st));
  if (U > 0) FCK(cudaStreamWaitEvent(st, fs->ev_join, 0));      // ms_sort = what is left of the sort after the splines
  FCK(cudaEventRecord(ev[2], st));

  // ---- phase 2: per-unit k3 brackets and spline weights ----
  FCK(fs->uw.reserve((size_t)U)); FCK(fs->ujp.reserve((size_t)U));
  FCK(fs->S.reserve((size_t)U * tau_size * fs->tpp_max));
  FCK(fs->jcount.reserve((size_t)h.n_pt_total));
  if (U > 0) {
    FCK(cudaMemsetAsync(fs->jcount.p, 0, (size_t)h.n_pt_total * sizeof(int), st));
    tc_prep_units_kernel<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(b, fs->perm.p, fs->unit_pair.p, P.flat_extrap,
                                                                     fs->uw.p, fs->ujp.p, fs->jcount.p, P.err);
    FCK(cudaGetLastError());
    ++*n_launches;
  }
  FCK(cudaEventRecord(ev[3], st));

  // ---- phase 3: per (column block, K chunk): interpolated sources of the chunk, then the DMMA kernel ----
  // One source slab S (the largest chunk's size) is reused by every chunk: the slab is what limits the number of
  // k3 values per batch, and the kernel's efficiency grows with that number (more k3 values per table row).
  // Later chunks of a block start from the unscaled partial sums the previous launch left in `out`.
  TcParams &TP = fs->tp;
  TP = TcParams{};
  TP.n_blocks = fs->n_blocks; TP.n_groups = fs->n_groups; TP.n_pass = fs->n_pass;
  for (int i = 0; i < fs->n_blocks; ++i) TP.blk[i] = fs->blk[i];
  for (int i = 0; i < fs->n_groups; ++i) { TP.grp[i] = fs->grp[i]; TP.grp[i].s_off = 0; }
  for (int i = 0; i < fs->n_pass; ++i) TP.pass[i] = fs->pass[i];
  TP.col_tt = fs->col_tt.p; TP.col_scale = fs->col_scale.p;
  TP.Jq = fs->Jq.p; TP.jq_row = fs->jq_row; TP.nact = fs->nact.p;
  TP.k_sorted = fs->keys_out.p; TP.perm = fs->perm.p; TP.n_units = U;
  TP.tau_size = tau_size;
  TP.tau0_minus_tau = P.tm.tau0_minus_tau; TP.delta_tau = P.tm.delta_tau; TP.xx = P.xx;
  TP.inv_xx_step = 1.0 / P.xx_step; TP.xx_size = P.xx_size; TP.xx_step_ref = P.xx_step;
  TP.xx_step = h.xx_delta; TP.xx_uniform = h.xx_uniform;
  if (getenv("SONG_LOS_TC_XX_TABLE")) TP.xx_uniform = 0;
  TP.out = b.out; TP.out_stride = U; TP.err = P.err;
  if (const char *v = getenv("SONG_LOS_TC_DBG")) TP.dbg = atoi(v);

  auto r128 = [](size_t x) { return (x + 127) / 128 * 128; };
  TP.hdr_bytes = (int)r128(16 + (size_t)MT * 4 + (size_t)MT * 16);
  TP.a_bytes = (int)r128((size_t)MT * fs->tpp_max * 8);
  TP.b_row_bytes = fs->nsb_pass_max * fs->kcs_max * 256;
  int rmax = (MT >= 128) ? 4 : (MT >= 64 ? 3 : 2);
  if (rmax_sel > 0) rmax = rmax_sel;
  if (const char *v = getenv("SONG_LOS_TC_RMAX")) rmax = std::max(1, std::min(8, atoi(v)));
  TP.n_cols_total = fs->n_cols_total; TP.cols_pad = (fs->n_cols_total + 31) & ~31;
  const size_t fixed = 128 + (size_t)((tau_size + 15) & ~15) * 32 + (size_t)TP.cols_pad * 12;
  const size_t budget = 224 * 1024;
  auto stage_bytes = [&](int r) { return (size_t)TP.hdr_bytes + TP.a_bytes + (size_t)r * TP.b_row_bytes; };
  while (rmax > 1 && (budget - fixed) / stage_bytes(rmax) < 3) --rmax;
  int ns = (int)std::min<size_t>(TC_MAX_STAGES, (budget - fixed) / stage_bytes(rmax));
  if (const char *v = getenv("SONG_LOS_TC_STAGES")) ns = std::max(2, std::min(ns, atoi(v)));
  if (ns < 2) { err = "FAST kernel: shared-memory stage does not fit"; return 1; }
  TP.rmax = rmax; TP.n_stages = ns;
  TP.stage_bytes = (int)stage_bytes(rmax);
  const size_t smem = fixed + (size_t)ns * TP.stage_bytes;

  TcInterp &I = fs->interp;
  int interp_threads = TC_INTERP_THREADS;
  if (const char *v = getenv("SONG_LOS_TC_INTERP_THREADS")) interp_threads = std::max(64, std::min(512, atoi(v) / 32 * 32));
  // tau samples per CTA: two staged rows of tauc*tp2 {y, dd} entries within ~64 KB
  int tauc = std::max(1, std::min(tau_size, (int)((64 * 1024) / (2 * 16 * (size_t)tp2))));
  if (const char *v = getenv("SONG_LOS_TC_TAUC")) tauc = std::max(1, std::min(tauc, atoi(v)));
  int n_chunks = 0;
  for (int ib = 0; ib < fs->n_blocks && U > 0; ++ib) {
    int pb = fs->n_pass, pe = 0;
    for (int ip = 0; ip < fs->n_pass; ++ip) if (fs->pass[ip].blk == ib) { pb = std::min(pb, ip); pe = std::max(pe, ip + 1); }
    if (pb >= pe) continue;
    const int g_first = fs->pass[pb].g0, nq = fs->pass[pb].nq;
    for (int q = 0; q < nq; ++q, ++n_chunks) {
      const int g = g_first + q;
      const TcGroup &G = fs->grp[g];
      if (5 + 2 * n_chunks + 1 >= FAST_MAX_EVENTS) { err = "FAST path: too many K chunks for the timing events"; return 1; }
      FCK(cudaEventRecord(ev[5 + 2 * n_chunks], st));
      // interpolated, sign-folded sources of this chunk: S[sorted unit][tau][slot]
      I = TcInterp{};
      I.n_groups = 1; I.tauc = tauc;
      I.grp_tpp[0] = G.tpp; I.grp_s_off[0] = 0;
      I.slot_tp = fs->slot_tp.p + (size_t)g * TC_SLOTS; I.slot_sign = fs->slot_sign.p + (size_t)g * TC_SLOTS;
      I.jcount = fs->jcount.p;
      I.col_stride = I.tauc * G.tpp;
      dim3 grid((unsigned)b.n_pairs, (unsigned)((tau_size + I.tauc - 1) / I.tauc));
      const size_t ismem = (size_t)2 * I.tauc * tp2 * 16 + (size_t)TC_MAX_GROUPS * TC_SLOTS * 12 + (size_t)I.col_stride * 4;
      FCK(cudaFuncSetAttribute(tc_interp_units_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ismem));
      tc_interp_units_kernel<<<grid, interp_threads, ismem, st>>>(b, I, tau_size, tp2, fs->CSD.p, fs->uw.p, fs->ujp.p, fs->S.p);
      FCK(cudaGetLastError());
      ++*n_launches;
      FCK(cudaEventRecord(ev[6 + 2 * n_chunks], st));

      {
        cuuint64_t gdim[3] = {(cuuint64_t)G.tpp, (cuuint64_t)tau_size, (cuuint64_t)U};
        cuuint64_t gstr[2] = {(cuuint64_t)G.tpp * 8, (cuuint64_t)G.tpp * 8 * (cuuint64_t)tau_size};
        cuuint32_t box[3] = {(cuuint32_t)G.tpp, 1u, (cuuint32_t)MT};
        cuuint32_t estr[3] = {1u, 1u, 1u};
        CUresult r = fs->encode(&TP.tmap[g], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, fs->S.p, gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r); return 1; }
      }
      TP.pass_begin = pb; TP.pass_end = pe; TP.q_fixed = q;
      TP.acc_in = (q > 0); TP.acc_final = (q == nq - 1);
      cudaError_t e = cudaErrorInvalidValue;
      const int key = wm * 10000 + wn * 100 + nw;
      const int64_t n_ctas = std::min<int64_t>(n_tiles, std::max(1, device_sms));     // persistent: one CTA per SM
      const int np1 = getenv("SONG_LOS_TC_NP1") != nullptr;     // A/B: a single producer warp
      switch (key) {
        case 80108:  e = np1 ? launch_tc<8, 1, 8, 1>(TP, n_ctas, smem, st) : launch_tc<8, 1, 8, 4>(TP, n_ctas, smem, st); break;
        case 80110:  e = launch_tc<8, 1, 10, 4>(TP, n_ctas, smem, st); break;
        case 40208:  e = launch_tc<4, 2, 8, 2>(TP, n_ctas, smem, st); break;
        case 40210:  e = launch_tc<4, 2, 10, 2>(TP, n_ctas, smem, st); break;
        case 20408:  e = launch_tc<2, 4, 8, 1>(TP, n_ctas, smem, st); break;
        case 20410:  e = launch_tc<2, 4, 10, 1>(TP, n_ctas, smem, st); break;
        // three consumer warps per SM sub-partition
        case 120107: e = np1 ? launch_tc<12, 1, 7, 1>(TP, n_ctas, smem, st) : launch_tc<12, 1, 7, 3>(TP, n_ctas, smem, st); break;
        case 120108: e = np1 ? launch_tc<12, 1, 8, 3>(TP, n_ctas, smem, st) : launch_tc<12, 1, 8, 1>(TP, n_ctas, smem, st); break;
        case 120110: e = launch_tc<12, 1, 10, 1>(TP, n_ctas, smem, st); break;
        case 60205:  e = launch_tc<6, 2, 5, 1>(TP, n_ctas, smem, st); break;
        case 60208:  e = launch_tc<6, 2, 8, 1>(TP, n_ctas, smem, st); break;
        case 40307:  e = launch_tc<4, 3, 7, 1>(TP, n_ctas, smem, st); break;
        case 40308:  e = launch_tc<4, 3, 8, 1>(TP, n_ctas, smem, st); break;
        case 80205:  e = launch_tc<8, 2, 5, 2>(TP, n_ctas, smem, st); break;
        default: err = "FAST kernel: no instantiation for the requested warp grid"; return 1;
      }
      FCK(e);
      ++*n_launches;
    }
    // the block's rows are final: send them home under the kernels of the next block
    if (early && fs->blk_tt_n[ib] > 0 && early->n_ranges < FAST_MAX_BLOCKS && !getenv("SONG_LOS_NO_EARLY_D2H")) {
      const int k = early->n_ranges;
      FCK(cudaEventRecord(early->ev[k], st));
      FCK(cudaStreamWaitEvent(early->st_d2h, early->ev[k], 0));
      FCK(cudaMemcpyAsync(early->host + (size_t)fs->blk_tt_lo[ib] * (size_t)U, b.out + (size_t)fs->blk_tt_lo[ib] * (size_t)U,
                          (size_t)fs->blk_tt_n[ib] * (size_t)U * sizeof(double), cudaMemcpyDeviceToHost, early->st_d2h));
      early->tt_lo[k] = fs->blk_tt_lo[ib]; early->tt_n[k] = fs->blk_tt_n[ib];
      early->n_ranges = k + 1;
    }
  }
  FCK(cudaEventRecord(ev[4], st));
  *n_chunks_out = n_chunks;
  return 0;
}

int fast_collect(cudaEvent_t *ev, int n_chunks, song_los_timing *tm, std::string &err)
{
  FCK(cudaEventSynchronize(ev[4]));
  float ms;
  FCK(cudaEventElapsedTime(&ms, ev[0], ev[1])); tm->ms_spline = ms;
  FCK(cudaEventElapsedTime(&ms, ev[1], ev[2])); tm->ms_sort = ms;
  FCK(cudaEventElapsedTime(&ms, ev[2], ev[3])); tm->ms_interp = ms;            // prep_units
  tm->ms_integrate = 0.f;
  for (int i = 0; i < n_chunks; ++i) {
    FCK(cudaEventElapsedTime(&ms, ev[5 + 2 * i], ev[6 + 2 * i])); tm->ms_interp += ms;
    FCK(cudaEventElapsedTime(&ms, ev[6 + 2 * i], i + 1 < n_chunks ? ev[5 + 2 * (i + 1)] : ev[4])); tm->ms_integrate += ms;
  }
  FCK(cudaEventElapsedTime(&ms, ev[0], ev[4])); tm->ms_total = ms;
  return 0;
}

}  // namespace songlos
