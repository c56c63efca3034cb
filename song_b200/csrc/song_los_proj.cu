// song_los_proj.cu -- device assembly of the projection functions J_Llm(x)
// (reference: bessel2_J_for_Llm + bessel2_J_Llm, source/bessel2.c:1041-1588).
//
// Compiled with -fmad=false: the sum over l1 is accumulated in the reference's order with
// separate multiply and add, so the tables are bit-identical to what the host loop of
// bessel2.c:1481-1555 produces from the same j_l1 table and the same 3j symbols.
//
//   proj_rows_kernel    J_Llm(xx[ix]) for every ix and every (J,L,l,m)        (bessel2.c:1400-1423)
//   proj_xmin_kernel    first grid index where |J| >= J_Llm_cut, by the reference's bisection on
//                       grid indices                                          (bessel2.c:1202-1342)
//   proj_dd_kernel      second derivatives of every J_Llm row for the cubic interpolation of
//                       transfer2.c:2262-2271 (bessel2.c:395-435: array_spline_table_one_column,
//                       _SPLINE_EST_DERIV_ for l = 2, _SPLINE_NATURAL_ otherwise)
//   proj_dense_kernel   masked {J[ix], J[ix+1]} table used by the LOS kernels (transfer2.c:2232)
#include <cuda_runtime.h>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/song_los.h"
#include "los_common.h"
#include "song_los_ctx.h"

namespace songlos {

struct ProjBuild {
  int xx_size, ndim;                // ndim = 2*L_max+1
  int n_jidx;
  const double *jl;                 // dense j_l1 table [l1_size][xx_size], zero below index_xmin_l1
  const int *jl_xmin;               // index_xmin_l1
  const int *l1;                    // l1 values
  const double *coef;               // [n_jidx][ndim]  pre*(2 l1+1)*first_3j*second_3j, 0 = skipped term
  const int *row0;                  // [n_jidx] index of l1_min in the l1 list
  const int *nterm;                 // [n_jidx] number of l1 terms (0: table vanishes identically)
  const double *outer;              // [n_jidx] ALTERNATING_SIGN(m)*(2l+1)
  const int64_t *col_of;            // [n_jidx] dense column, -1 if unused
};

__global__ void proj_rows_kernel(ProjBuild B, double *rows /* [n_jidx][xx_size] */)
{
  int ix = blockIdx.x * blockDim.x + threadIdx.x;
  int q = blockIdx.y;
  if (ix >= B.xx_size) return;
  int n = B.nterm[q];
  double J = 0;
  const double *cf = B.coef + (size_t)q * B.ndim;
  int r0 = B.row0[q];
  for (int i = 0; i < n; ++i) {
    double c = cf[i];
    if (c == 0) continue;                                   // 3j == 0 or wrong parity, bessel2.c:1489,1516-1530
    int row = r0 + i;
    if (ix - B.jl_xmin[row] < 0) continue;                  // j_l1 below its own cut, bessel2.c:1499
    J += c * B.jl[(size_t)row * B.xx_size + ix];            // bessel2.c:1543-1546
  }
  J *= B.outer[q];                                          // bessel2.c:1555
  rows[(size_t)q * B.xx_size + ix] = J;
}

__global__ void proj_xmin_kernel(ProjBuild B, const double *rows, const double *xx, double xx_max,
                                 double J_cut, const int *L_of, const int *l_of, const int *l1min_of,
                                 int *index_xmin, int *x_size, int *err)
{
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B.n_jidx) return;
  if (B.col_of[q] < 0) { index_xmin[q] = B.xx_size - 1; x_size[q] = 0; return; }   // not allocated
  int xmin;
  const double *row = rows + (size_t)q * B.xx_size;
  if (B.nterm[q] == 0) {
    xmin = B.xx_size - 1;                                   // spin / EB constraint, bessel2.c:1097-1100
  } else if (L_of[q] == l_of[q]) {
    xmin = 0;                                               // bessel2.c:1202-1206
  } else {
    int down = 0, up;
    double x_up = l1min_of[q] + 0.5;                        // bessel2.c:1225
    if (x_up >= xx_max) { up = B.xx_size - 1; x_up = xx_max; }
    else { up = 0; while (xx[up] < x_up) up++; x_up = xx[up]; }
    if (fabs(row[up]) < J_cut) {                            // bessel2.c:1276
      xmin = B.xx_size - 1;
      if (x_up < xx_max) atomicOr(err, 16);                 // "wrong initial guess for x_min_up"
    } else {
      while (down != up - 1) {                              // bessel2.c:1303
        int mid = (up + down) / 2;
        if (fabs(row[mid]) >= J_cut) up = mid; else down = mid;
      }
      xmin = down;
    }
  }
  index_xmin[q] = xmin;
  x_size[q] = B.xx_size - xmin;                             // bessel2.c:1357
}

// one thread per (J,L,l,m) row: the tridiagonal sweep is sequential in x (up to xx_size points)
__global__ void proj_dd_kernel(ProjBuild B, const double *rows, const double *xx, const int *index_xmin, const int *x_size,
                               const int *l_of, double *ddrows, double *work)
{
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B.n_jidx || B.col_of[q] < 0) return;
  const int xmin = index_xmin[q];
  const int n = (xmin == B.xx_size - 1) ? 1 : x_size[q];    // all-negligible marker: a single zero (bessel2.c:1378-1388)
  if (n < 3) return;                                       // ddrows is zero-initialised
  const double *x = xx + xmin;
  const double *y = rows + (size_t)q * B.xx_size + xmin;
  double *dd = ddrows + (size_t)q * B.xx_size + xmin;
  double *u = work + (size_t)q * B.xx_size;
  const bool est = (l_of[q] == 2);                         // bessel2.c:414
  if (!est) { dd[0] = 0.0; u[0] = 0.0; }
  else {
    double dy_first = ((x[2] - x[0]) * (x[2] - x[0]) * (y[1] - y[0]) - (x[1] - x[0]) * (x[1] - x[0]) * (y[2] - y[0]))
                      / ((x[2] - x[0]) * (x[1] - x[0]) * (x[2] - x[1]));
    dd[0] = -0.5;
    u[0] = (3. / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - dy_first);
  }
  for (int i = 1; i < n - 1; ++i) {
    double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    double p = sig * dd[i - 1] + 2.0;
    dd[i] = (sig - 1.0) / p;
    double ui = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
    u[i] = (6.0 * ui / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
  }
  double qn, un;
  if (!est) { qn = 0.0; un = 0.0; }
  else {
    double dy_last = ((x[n - 3] - x[n - 1]) * (x[n - 3] - x[n - 1]) * (y[n - 2] - y[n - 1])
                      - (x[n - 2] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (y[n - 3] - y[n - 1]))
                     / ((x[n - 3] - x[n - 1]) * (x[n - 2] - x[n - 1]) * (x[n - 3] - x[n - 2]));
    qn = 0.5;
    un = (3. / (x[n - 1] - x[n - 2])) * (dy_last - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
  }
  dd[n - 1] = (un - qn * u[n - 2]) / (qn * dd[n - 2] + 1.0);
  for (int i = n - 2; i >= 0; --i) dd[i] = dd[i] * dd[i + 1] + u[i];
}

__global__ void proj_dense_kernel(ProjBuild B, const double *rows, const int *index_xmin,
                                  int64_t n_cols, double2 *dense)
{
  int ix = blockIdx.x * blockDim.x + threadIdx.x;
  int q = blockIdx.y;
  if (ix >= B.xx_size) return;
  int64_t col = B.col_of[q];
  if (col < 0) return;
  int xmin = index_xmin[q];
  double2 v = make_double2(0., 0.);
  if (xmin != B.xx_size - 1 && ix >= xmin) {                // all-negligible marker stays empty
    const double *row = rows + (size_t)q * B.xx_size;
    v.x = row[ix];
    if (ix + 1 < B.xx_size) v.y = row[ix + 1];
  }
  dense[(size_t)ix * n_cols + col] = v;
}

}  // namespace songlos

using namespace songlos;

#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return ctx_fail(ctx, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

extern "C" int song_los_build_projection(song_los_ctx *ctx, int l1_size, const int *l1,
                                         const double *const *j_l1, const int *index_xmin_l1,
                                         const int *x_size_l1,
                                         int n3j, const double *first_3j, const double *second_3j,
                                         const int *l1_min, const int *l1_count, const int *index_l1_of_l1min,
                                         const int *J_type_of_index_J, double J_Llm_cut, double xx_max)
{
  if (!ctx || !l1 || !j_l1 || !index_xmin_l1 || !x_size_l1 || !first_3j || !second_3j || !l1_min ||
      !l1_count || !index_l1_of_l1min || !J_type_of_index_J)
    return ctx_fail(ctx, "song_los_build_projection: NULL argument");
  PCK(cudaSetDevice(ctx->device));
  const song_los_config &c = ctx->cfg;
  const int n_jidx = c.J_size * c.L_size * c.l_size * c.m_size;
  if (n3j != n_jidx) return ctx_fail(ctx, "song_los_build_projection: expected %d 3j blocks, got %d", n_jidx, n3j);
  const bool cubic_J = (c.bessels_interpolation == SONG_LOS_INTERP_CUBIC);
  const int ndim = 2 * (c.L_size - 1) + 1;
  const int xs = c.xx_size;
  cudaStream_t st = ctx->stream;

  // dense j_l1 table
  std::vector<double> jl((size_t)l1_size * xs, 0.0);
  for (int i = 0; i < l1_size; ++i) {
    if (index_xmin_l1[i] < 0 || index_xmin_l1[i] + x_size_l1[i] > xs)
      return ctx_fail(ctx, "song_los_build_projection: j_l1 row %d outside the x grid", i);
    std::memcpy(jl.data() + (size_t)i * xs + index_xmin_l1[i], j_l1[i], (size_t)x_size_l1[i] * sizeof(double));
  }
  // per-(J,L,l,m) coefficient of every l1 term, in the reference's multiplication order
  std::vector<double> coef((size_t)n_jidx * ndim, 0.0), outer((size_t)n_jidx, 0.0);
  std::vector<int> row0((size_t)n_jidx, 0), nterm((size_t)n_jidx, 0), L_of((size_t)n_jidx), l_of((size_t)n_jidx),
      l1min_of((size_t)n_jidx, 0);
  std::vector<int64_t> col_of((size_t)n_jidx, -1);
  ctx->n_cols = (int64_t)c.J_size * c.L_size * c.n_transfers;
  for (int J = 0; J < c.J_size; ++J)
    for (int L = 0; L < c.L_size; ++L)
      for (int il = 0; il < c.l_size; ++il)
        for (int im = 0; im < c.m_size; ++im) {
          size_t q = (((size_t)J * c.L_size + L) * c.l_size + il) * c.m_size + im;
          int l = c.l[il], m = c.m[im];
          L_of[q] = L; l_of[q] = l;
          int lm = c.lm_array[(size_t)il * c.m_size + im];
          if (lm < 0 || m > L) continue;                       // not allocated (bessel2.c:288)
          col_of[q] = ((int64_t)J * c.L_size + L) * c.n_transfers + lm;
          int n = l1_count[q];
          if (n == 0) continue;                                // vanishing table
          if (n > ndim) return ctx_fail(ctx, "song_los_build_projection: %d l1 terms exceed 2*L_max+1", n);
          int pf = J_type_of_index_J[J];
          row0[q] = index_l1_of_l1min[q];
          l1min_of[q] = l1_min[q];
          nterm[q] = n;
          if (row0[q] < 0 || row0[q] + n > l1_size) return ctx_fail(ctx, "error in ell_1 list");   // bessel2.c:1177
          for (int i = 0; i < n; ++i) {
            double f1 = first_3j[q * ndim + i], f2 = second_3j[q * ndim + i];
            if (f1 == 0) continue;                             // bessel2.c:1489
            int l1v = l1[row0[q] + i];
            int is_even = ((l - l1v - L) % 2 == 0);
            double pre;
            if (pf == SONG_LOS_J_TT || pf == SONG_LOS_J_EE) {
              if (!is_even) continue;
              pre = (((l - l1v - L) / 2) % 2 == 0) ? 1 : -1;   // bessel2.c:1521
            } else {
              if (is_even) continue;
              pre = (((l - l1v - L - 1) / 2) % 2 == 0) ? 1 : -1;   // bessel2.c:1527
            }
            coef[q * ndim + i] = pre * (2 * l1v + 1) * f1 * f2;    // bessel2.c:1543-1545
          }
          outer[q] = ((m % 2 == 0) ? 1 : -1) * (2 * l + 1);        // bessel2.c:1555
        }

  double *d_jl = nullptr, *d_coef = nullptr, *d_outer = nullptr, *d_rows = nullptr;
  int *d_jlx = nullptr, *d_l1 = nullptr, *d_row0 = nullptr, *d_nterm = nullptr, *d_L = nullptr, *d_l = nullptr,
      *d_l1min = nullptr, *d_xmin = nullptr, *d_xsize = nullptr;
  int64_t *d_col = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_jl); cudaFree(d_coef); cudaFree(d_outer); cudaFree(d_rows); cudaFree(d_jlx); cudaFree(d_l1);
    cudaFree(d_row0); cudaFree(d_nterm); cudaFree(d_L); cudaFree(d_l); cudaFree(d_l1min); cudaFree(d_xmin);
    cudaFree(d_xsize); cudaFree(d_col);
  };
#define PUP(dst, vec) do { cudaError_t e_ = cudaMalloc(&dst, (vec).size() * sizeof((vec)[0])); \
    if (e_ == cudaSuccess) e_ = cudaMemcpyAsync(dst, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice, st); \
    if (e_ != cudaSuccess) { cleanup(); return ctx_fail(ctx, "song_los_build_projection: upload failed: %s", cudaGetErrorString(e_)); } } while (0)
  std::vector<int> jlx(index_xmin_l1, index_xmin_l1 + l1_size), l1v(l1, l1 + l1_size);
  PUP(d_jl, jl); PUP(d_coef, coef); PUP(d_outer, outer); PUP(d_jlx, jlx); PUP(d_l1, l1v); PUP(d_row0, row0);
  PUP(d_nterm, nterm); PUP(d_L, L_of); PUP(d_l, l_of); PUP(d_l1min, l1min_of); PUP(d_col, col_of);
  cudaError_t e = cudaMalloc(&d_rows, (size_t)n_jidx * xs * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_xmin, (size_t)n_jidx * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&d_xsize, (size_t)n_jidx * sizeof(int));
  if (e == cudaSuccess) e = ctx->d_J.reserve((size_t)xs * ctx->n_cols);
  if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_J.p, 0, (size_t)xs * ctx->n_cols * sizeof(double2), st);
  if (e != cudaSuccess) { cleanup(); return ctx_fail(ctx, "song_los_build_projection: allocation failed: %s", cudaGetErrorString(e)); }

  ProjBuild B{xs, ndim, n_jidx, d_jl, d_jlx, d_l1, d_coef, d_row0, d_nterm, d_outer, d_col};
  dim3 grid((unsigned)((xs + 127) / 128), (unsigned)n_jidx);
  if (n_jidx > 65535) {   // split the y dimension
    cleanup();
    return ctx_fail(ctx, "song_los_build_projection: more than 65535 (J,L,l,m) tables in one call");
  }
  proj_rows_kernel<<<grid, 128, 0, st>>>(B, d_rows);
  proj_xmin_kernel<<<(n_jidx + 127) / 128, 128, 0, st>>>(B, d_rows, ctx->d_xx.p, xx_max, J_Llm_cut, d_L, d_l, d_l1min,
                                                        d_xmin, d_xsize, ctx->d_err.p);
  proj_dense_kernel<<<grid, 128, 0, st>>>(B, d_rows, d_xmin, ctx->n_cols, ctx->d_J.p);
  e = cudaGetLastError();
  double *d_ddrows = nullptr, *d_work = nullptr;
  if (e == cudaSuccess && cubic_J) {
    // second derivatives of every row (bessel2.c:395-435), then the same masked dense layout as J
    e = cudaMalloc(&d_ddrows, (size_t)n_jidx * xs * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_work, (size_t)n_jidx * xs * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_ddrows, 0, (size_t)n_jidx * xs * sizeof(double), st);
    if (e == cudaSuccess) e = ctx->d_ddJ.reserve((size_t)xs * ctx->n_cols);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_ddJ.p, 0, (size_t)xs * ctx->n_cols * sizeof(double2), st);
    if (e == cudaSuccess) {
      proj_dd_kernel<<<(n_jidx + 63) / 64, 64, 0, st>>>(B, d_rows, ctx->d_xx.p, d_xmin, d_xsize, d_l, d_ddrows, d_work);
      proj_dense_kernel<<<grid, 128, 0, st>>>(B, d_ddrows, d_xmin, ctx->n_cols, ctx->d_ddJ.p);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_work);
    if (e != cudaSuccess) cudaFree(d_ddrows);
    else { if (ctx->d_proj_ddrows) cudaFree(ctx->d_proj_ddrows); ctx->d_proj_ddrows = d_ddrows; }
  }
  // host mirrors: index tables + ragged rows for song_los_get_projection
  ctx->h_jidx_xmin.resize((size_t)n_jidx); ctx->h_jidx_xsize.resize((size_t)n_jidx);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_jidx_xmin.data(), d_xmin, (size_t)n_jidx * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_jidx_xsize.data(), d_xsize, (size_t)n_jidx * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cleanup(); return ctx_fail(ctx, "song_los_build_projection: kernels failed: %s", cudaGetErrorString(e)); }
  int herr = 0;
  cudaMemcpy(&herr, ctx->d_err.p, sizeof(int), cudaMemcpyDeviceToHost);
  if (herr) {
    cudaMemset(ctx->d_err.p, 0, sizeof(int));
    cleanup();
    return ctx_fail(ctx, "in bisection, wrong initial guess for x_min_up");              // bessel2.c:1287
  }
  // keep the full rows on the device for get_projection (freed there or at destroy)
  if (ctx->d_proj_rows) cudaFree(ctx->d_proj_rows);
  ctx->d_proj_rows = d_rows; d_rows = nullptr;
  ctx->proj_rows_n = n_jidx;
  std::vector<unsigned char> bnz((size_t)c.J_size * c.L_size, 0);
  for (int q = 0; q < n_jidx; ++q)
    if (col_of[(size_t)q] >= 0 && ctx->h_jidx_xmin[(size_t)q] != xs - 1)
      bnz[(size_t)(col_of[(size_t)q] / c.n_transfers)] = 1;
  ctx->h_block_nonzero = bnz;
  e = ctx->d_block_nonzero.reserve(bnz.size());
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_block_nonzero.p, bnz.data(), bnz.size(), cudaMemcpyHostToDevice);
  cleanup();
  if (e != cudaSuccess) return ctx_fail(ctx, "song_los_build_projection: %s", cudaGetErrorString(e));
  for (int q = 0; q < n_jidx; ++q)
    if (col_of[(size_t)q] < 0) { ctx->h_jidx_xmin[(size_t)q] = xs - 1; ctx->h_jidx_xsize[(size_t)q] = 0; }
  fast_invalidate_table(ctx->fast);
  ctx->have_proj = true;
  return SONG_LOS_SUCCESS;
}

static int get_rows(song_los_ctx *ctx, const double *d_rows, double *const *out_rows);

// The undensified rows are kept for song_los_get_projection only; a host module that has read them back (bessel2_init)
// returns their memory (n_jidx x xx_size doubles, GBs at l_max >= 3000) to the batches.
extern "C" int song_los_drop_projection_rows(song_los_ctx *ctx)
{
  if (!ctx) return SONG_LOS_FAILURE;
  PCK(cudaSetDevice(ctx->device));
  if (ctx->d_proj_rows) { PCK(cudaFree(ctx->d_proj_rows)); ctx->d_proj_rows = nullptr; }
  if (ctx->d_proj_ddrows) { PCK(cudaFree(ctx->d_proj_ddrows)); ctx->d_proj_ddrows = nullptr; }
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_get_projection_dd(song_los_ctx *ctx, double *const *ddJ_rows)
{
  if (!ctx || !ddJ_rows) return ctx_fail(ctx, "song_los_get_projection_dd: NULL argument");
  PCK(cudaSetDevice(ctx->device));
  if (!ctx->d_proj_ddrows) return ctx_fail(ctx, "song_los_get_projection_dd: no device-built second derivatives (bessels_interpolation is not cubic)");
  return get_rows(ctx, ctx->d_proj_ddrows, ddJ_rows);
}

static int get_rows(song_los_ctx *ctx, const double *d_rows, double *const *out_rows)
{
  const int n = ctx->proj_rows_n, xs = ctx->cfg.xx_size;
  for (int q = 0; q < n; ++q) {
    int sz = ctx->h_jidx_xsize[(size_t)q], xm = ctx->h_jidx_xmin[(size_t)q];
    if (sz <= 0 || !out_rows[q]) continue;
    if (xm == xs - 1) { out_rows[q][0] = 0.; continue; }       // bessel2.c:1378-1388
    PCK(cudaMemcpyAsync(out_rows[q], d_rows + (size_t)q * xs + xm, (size_t)sz * sizeof(double),
                        cudaMemcpyDeviceToHost, ctx->stream));
  }
  PCK(cudaStreamSynchronize(ctx->stream));
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_get_projection(song_los_ctx *ctx, int *index_xmin_J, int *x_size_J, double *const *J_rows)
{
  if (!ctx) return SONG_LOS_FAILURE;
  PCK(cudaSetDevice(ctx->device));
  if (!ctx->d_proj_rows) return ctx_fail(ctx, "song_los_get_projection: no device-built projection functions");
  const int n = ctx->proj_rows_n, xs = ctx->cfg.xx_size;
  if (index_xmin_J) std::memcpy(index_xmin_J, ctx->h_jidx_xmin.data(), (size_t)n * sizeof(int));
  if (x_size_J) std::memcpy(x_size_J, ctx->h_jidx_xsize.data(), (size_t)n * sizeof(int));
  if (J_rows) {
    for (int q = 0; q < n; ++q) {
      int sz = ctx->h_jidx_xsize[(size_t)q], xm = ctx->h_jidx_xmin[(size_t)q];
      if (sz <= 0 || !J_rows[q]) continue;
      if (xm == xs - 1) { J_rows[q][0] = 0.; continue; }       // bessel2.c:1378-1388
      PCK(cudaMemcpyAsync(J_rows[q], ctx->d_proj_rows + (size_t)q * xs + xm, (size_t)sz * sizeof(double),
                          cudaMemcpyDeviceToHost, ctx->stream));
    }
    PCK(cudaStreamSynchronize(ctx->stream));
  }
  return SONG_LOS_SUCCESS;
}
