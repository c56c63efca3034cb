// los_k3int.cu -- the k3 integral of the intrinsic bispectrum on the transfer functions still in device memory.
//
// Reference: bispectra2_intrinsic_integrate_over_k3 (source/bispectra2.c:1107-1439), whose inner loop is
// bessel2_convolution (source/bessel2.c:499-605):
//
//   I_l3(r, k1, k2) = 2 c_M3 * 1/2 sum_k3 delta_k3 * k3^2 * T_l3,M3(k1,k2,k3) * j_L3(k3 r)
//
// with the trapezoidal measure delta_k3 of bispectra2.c:1279-1284 on the pair's own k3 grid (the grid of
// transfer2_get_k3_list, already on the device as k3_tr), j_L3 interpolated linearly or with the cubic formula in the
// pbs2->j_l1 table (bessel2.c:571-588), terms skipped where T vanishes or x = k3 r lies below the table (:541-563).
// It consumes T_lm(k1,k2,k3) where song_los_compute left it, so a caller that only needs I_l3 (bispectra2.c) copies
// r_size numbers per (l3, k1, k2) to the host instead of k3_size.
//
// One CTA per (requested transfer type, pair); thread <-> r.  Every thread walks the pair's k3 values in the reference's
// order with the reference's statements (compiled -fmad=false): the result is bit-identical to the CPU loop.  The k3 and
// T values of a step are the same for the whole CTA (broadcast loads), the table reads of neighbouring r fall within a
// few hundred bytes of each other.
#include <cuda_runtime.h>
#include <vector>
#include "song_los_ctx.h"

namespace songlos {

struct K3IntParams {
  int n_req, n_pairs, r_size;
  const int *req_tt, *req_row;            // per request: transfer type, row of the j_l1 table
  const double *req_coef;                 // per request: M3 coefficient; applied when req_has_coef (abs_M3 != 0)
  const int *req_has_coef;
  const double *r;
  const double *jl, *ddjl;                // packed rows
  const int64_t *row_off; const double *row_xmin; const int *row_size;
  double xx_step; int cubic;
  const int *nt; const int64_t *tr_off; const double *k3_tr;
  const double *T; int64_t n_units;       // [tt2_size][n_units]
  double *out;                            // [n_req][r_size][n_pairs]
  int *err;
};

__global__ void __launch_bounds__(128) k3_integral_kernel(K3IntParams P)
{
  const int p = blockIdx.x, q = blockIdx.y;
  const int n = P.nt[p];
  const double *kk = P.k3_tr + P.tr_off[p];
  const double *f = P.T + (int64_t)P.req_tt[q] * P.n_units + P.tr_off[p];
  const int row = P.req_row[q];
  const double *j = P.jl + P.row_off[row];
  const double *ddj = P.cubic ? P.ddjl + P.row_off[row] : nullptr;
  const double x_min = P.row_xmin[row], step = P.xx_step;
  const int jsize = P.row_size[row];
  for (int ir = threadIdx.x; ir < P.r_size; ir += blockDim.x) {
    const double r = P.r[ir];
    double integral = 0;
    for (int i = 0; i < n; ++i) {
      const double f_in_k = f[i];
      if (f_in_k == 0.) continue;                                        // bessel2.c:541-545
      const double g_in_k = 1.;
      const double k = kk[i];
      const double x = k * r;
      if (x < x_min) continue;                                            // :562
      // trapezoidal measure of bispectra2.c:1279-1284
      const double delta = (i == 0) ? kk[1] - kk[0] : (i == n - 1) ? kk[n - 1] - kk[n - 2] : kk[i + 1] - kk[i - 1];
      const int index_x = (int)((x - x_min) / step);                      // :565
      if (index_x + 1 >= jsize) { atomicOr(P.err, ERR_X_OUT_OF_RANGE); continue; }
      const double a = (x_min + step * (index_x + 1) - x) / step;         // :566
      double jv;
      if (!P.cubic) jv = a * j[index_x] + (1. - a) * j[index_x + 1];      // :571-574
      else jv = a * j[index_x] + (1. - a) * (j[index_x + 1] - a * ((a + 1.) * ddj[index_x] + (2. - a) * ddj[index_x + 1]) * step * step / 6.0);
      const double integrand = k * k * jv * f_in_k * g_in_k;              // :592
      integral += integrand * delta;                                      // :595
    }
    integral *= 0.5;                                                      // :601
    if (P.req_has_coef[q]) integral *= P.req_coef[q];                     // bispectra2.c:1369-1370
    integral *= 2;                                                        // :1375
    P.out[((int64_t)q * P.r_size + ir) * P.n_pairs + p] = integral;
  }
}

}  // namespace songlos

using namespace songlos;

#define KCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(ctx, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

template <typename T>
static cudaError_t up(DevBuf<T> &d, const T *h, size_t n, cudaStream_t st)
{
  cudaError_t e = d.reserve(n);
  if (e != cudaSuccess || n == 0) return e;
  return cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, st);
}

extern "C" int song_los_set_bessel_table(song_los_ctx *ctx, int n_rows, const double *const *j_rows,
                                         const double *const *ddj_rows, const double *x_min, const int *x_size)
{
  if (!ctx || n_rows < 1 || !j_rows || !x_min || !x_size) return ctx_fail(ctx, "song_los_set_bessel_table: NULL argument");
  KCK(cudaSetDevice(ctx->device));
  const bool cubic = ctx->cfg.bessels_interpolation == SONG_LOS_INTERP_CUBIC;
  if (cubic && !ddj_rows) return ctx_fail(ctx, "song_los_set_bessel_table: cubic Bessel interpolation needs pbs2->ddj_l1");
  std::vector<int64_t> off((size_t)n_rows);
  int64_t total = 0;
  for (int i = 0; i < n_rows; ++i) {
    if (x_size[i] < 1 || !j_rows[i] || (cubic && !ddj_rows[i])) return ctx_fail(ctx, "song_los_set_bessel_table: row %d is empty", i);
    off[i] = total; total += x_size[i];
  }
  std::vector<double> flat((size_t)total), ddflat(cubic ? (size_t)total : 0);
  for (int i = 0; i < n_rows; ++i) {
    std::copy(j_rows[i], j_rows[i] + x_size[i], flat.begin() + off[i]);
    if (cubic) std::copy(ddj_rows[i], ddj_rows[i] + x_size[i], ddflat.begin() + off[i]);
  }
  cudaStream_t st = ctx->stream;
  KCK(up(ctx->d_jl1, flat.data(), flat.size(), st));
  if (cubic) KCK(up(ctx->d_ddjl1, ddflat.data(), ddflat.size(), st));
  KCK(up(ctx->d_jl1_off, off.data(), off.size(), st));
  KCK(up(ctx->d_jl1_xmin, x_min, (size_t)n_rows, st));
  KCK(up(ctx->d_jl1_size, x_size, (size_t)n_rows, st));
  KCK(cudaStreamSynchronize(st));
  ctx->jl1_rows = n_rows;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_integrate_k3(song_los_ctx *ctx, const song_los_k3_request *rq, double *out)
{
  if (!ctx || !rq || !out) return ctx_fail(ctx, "song_los_integrate_k3: NULL argument");
  KCK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight) return ctx_fail(ctx, "song_los_integrate_k3: submitted batches are still in flight; call song_los_wait first");
  song_los_ctx::Slot &S = ctx->slot[ctx->cur];
  if (!S.have_result) return ctx_fail(ctx, "song_los_integrate_k3: no transfer functions on the device (song_los_compute first)");
  if (ctx->jl1_rows < 1) return ctx_fail(ctx, "song_los_integrate_k3: j_l1 table not set (song_los_set_bessel_table)");
  if (rq->n_requests < 1 || rq->r_size < 1 || !rq->tt || !rq->index_l1 || !rq->r)
    return ctx_fail(ctx, "song_los_integrate_k3: empty request");
  std::vector<int> has((size_t)rq->n_requests);
  std::vector<double> coef((size_t)rq->n_requests, 1.0);
  for (int q = 0; q < rq->n_requests; ++q) {
    if (rq->tt[q] < 0 || rq->tt[q] >= ctx->cfg.tt2_size) return ctx_fail(ctx, "song_los_integrate_k3: transfer type %d outside [0,%d)", rq->tt[q], ctx->cfg.tt2_size);
    if (rq->index_l1[q] < 0 || rq->index_l1[q] >= ctx->jl1_rows) return ctx_fail(ctx, "song_los_integrate_k3: index_l1=%d out of bounds (l1_size=%d)", rq->index_l1[q], ctx->jl1_rows);
    has[q] = (rq->M3_coefficient != nullptr && rq->abs_M3 != 0);
    if (has[q]) coef[q] = rq->M3_coefficient[q];
  }
  cudaStream_t st = ctx->stream;
  KCK(up(ctx->d_k3req_tt, rq->tt, (size_t)rq->n_requests, st));
  KCK(up(ctx->d_k3req_row, rq->index_l1, (size_t)rq->n_requests, st));
  KCK(up(ctx->d_k3req_has, has.data(), has.size(), st));
  KCK(up(ctx->d_k3req_coef, coef.data(), coef.size(), st));
  KCK(up(ctx->d_k3_r, rq->r, (size_t)rq->r_size, st));
  const size_t n_out = (size_t)rq->n_requests * rq->r_size * S.hb.n_pairs;
  KCK(ctx->d_k3_out.reserve(n_out));
  K3IntParams P;
  P.n_req = rq->n_requests; P.n_pairs = S.hb.n_pairs; P.r_size = rq->r_size;
  P.req_tt = ctx->d_k3req_tt.p; P.req_row = ctx->d_k3req_row.p; P.req_coef = ctx->d_k3req_coef.p; P.req_has_coef = ctx->d_k3req_has.p;
  P.r = ctx->d_k3_r.p; P.jl = ctx->d_jl1.p; P.ddjl = ctx->d_ddjl1.p;
  P.row_off = ctx->d_jl1_off.p; P.row_xmin = ctx->d_jl1_xmin.p; P.row_size = ctx->d_jl1_size.p;
  P.xx_step = ctx->cfg.xx_step; P.cubic = ctx->cfg.bessels_interpolation == SONG_LOS_INTERP_CUBIC;
  P.nt = S.d_nt; P.tr_off = S.d_tr_off; P.k3_tr = S.d_k3_tr.p; P.T = S.d_out.p; P.n_units = S.hb.n_units;
  P.out = ctx->d_k3_out.p; P.err = ctx->d_err.p;
  KCK(cudaMemsetAsync(ctx->d_err.p, 0, sizeof(int), st));
  cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
  KCK(cudaEventRecord(e0, st));
  dim3 grid((unsigned)S.hb.n_pairs, (unsigned)rq->n_requests);
  if (S.hb.n_pairs > 0) {
    k3_integral_kernel<<<grid, 128, 0, st>>>(P);
    KCK(cudaGetLastError());
  }
  KCK(cudaEventRecord(e1, st));
  int h_err = 0;
  KCK(cudaMemcpyAsync(&h_err, ctx->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  KCK(cudaMemcpyAsync(out, ctx->d_k3_out.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
  KCK(cudaStreamSynchronize(st));
  KCK(cudaEventElapsedTime(&ctx->k3_ms, e0, e1));
  if (h_err & ERR_X_OUT_OF_RANGE)
    return ctx_fail(ctx, "r*k_max is larger than x_max of the j_l1 table (bessel2.c:520-523); increase xx_max");
  return SONG_LOS_SUCCESS;
}

extern "C" float song_los_last_k3_ms(const song_los_ctx *ctx) { return ctx ? ctx->k3_ms : 0.f; }
