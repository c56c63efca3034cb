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

// ---------------------------------------------------------------------------------------------
// spectra2_cls (source/spectra2.c:241-570): the (k1,k2,k3) sum of the intrinsic C_l on the device-resident T
//   C_l^{XY} = sum_{k1, k2 <= k1, k3} total_factor(k1,k2,k3,m) T^X_lm(k1,k2,k3) T^Y_lm(k1,k2,k3)
// step_k3 of a pair depends on the k3 before it (the two triangular-edge flags, :468-492), so one thread per pair walks
// its grid once and stores the steps; the products are then summed per (request, pair) in the reference's k3 order and
// the pairs of the batch are added in pair order.
// ---------------------------------------------------------------------------------------------
struct ClParams {
  int n_req, n_pairs;
  const int *tt1, *tt2, *m;
  int rescale;
  const double *k, *prim; int k_size;
  const int *pair_k1, *pair_k2;
  const int *nt; const int64_t *tr_off; const double *k3_tr;
  const double *T; int64_t n_units;
  double *step_k3;                        // [n_units]
  double *partial;                        // [n_req][n_pairs]
  double *out;                            // [n_req]
};

__global__ void cl_step_k3_kernel(ClParams P)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.n_pairs) return;
  const int n = P.nt[p];
  const double *k3_grid = P.k3_tr + P.tr_off[p];
  double *step = P.step_k3 + P.tr_off[p];
  const double k1 = P.k[P.pair_k1[p]], k2 = P.k[P.pair_k2[p]];
  int triangular_first = 0, triangular_last = 0;
  for (int i = 0; i < n; ++i) {
    const double k3 = k3_grid[i];
    double step_k3;
    if (i == 0) step_k3 = (k3_grid[1] - k3_grid[0]) / 2.;
    else if (i == n - 1) step_k3 = (k3_grid[n - 1] - k3_grid[n - 2]) / 2.;
    else step_k3 = (k3_grid[i + 1] - k3_grid[i - 1]) / 2.;
    if (k3 < k1 - k2 || k3 > k1 + k2) step_k3 = 0.;                                   // :481-483
    if (k3 >= k1 - k2 && triangular_first == 0 && i < n - 1) {                       // :485-488
      triangular_first = 1;
      step_k3 = (k3_grid[i + 1] - k3) / 2. + k3 - (k1 - k2);
    }
    if (i + 1 < n && k3_grid[i + 1] > k1 + k2 && triangular_last == 0 && i > 0) {    // :489-492 (the reference reads one past the
      triangular_last = 1;                                                            //  grid at its last point; never true there)
      step_k3 = (k3 - k3_grid[i - 1]) / 2. + (k1 + k2) - k3;
    }
    step[i] = step_k3;
  }
}

__global__ void __launch_bounds__(128) cl_pair_sum_kernel(ClParams P)
{
  const int p = blockIdx.x;
  const int n = P.nt[p];
  const int ik1 = P.pair_k1[p], ik2 = P.pair_k2[p];
  const double k1 = P.k[ik1], k2 = P.k[ik2];
  const double _PI_ = 3.1415926535897932384626433832795e0;
  // trapezoidal steps in k1 and k2 (:377-381, :409-414) and the primordial factors (:398, :430)
  double step_k1, step_k2;
  if (ik1 == 0) step_k1 = (P.k[1] - P.k[0]) / 2.;
  else if (ik1 == P.k_size - 1) step_k1 = (P.k[P.k_size - 1] - P.k[P.k_size - 2]) / 2.;
  else step_k1 = (P.k[ik1 + 1] - P.k[ik1 - 1]) / 2.;
  if (ik1 == 0) step_k2 = 0.;
  else if (ik2 == 0) step_k2 = (P.k[1] - P.k[0]) / 2.;
  else if (ik2 == ik1) step_k2 = (P.k[ik1] - P.k[ik1 - 1]) / 2.;
  else step_k2 = (P.k[ik2 + 1] - P.k[ik2 - 1]) / 2.;
  const double spectra_k1 = 2 * _PI_ * _PI_ / (k1 * k1 * k1) * P.prim[ik1];
  const double spectra_k2 = 2 * _PI_ * _PI_ / (k2 * k2 * k2) * P.prim[ik2];
  const double *k3_grid = P.k3_tr + P.tr_off[p];
  const double *step = P.step_k3 + P.tr_off[p];
  for (int q = threadIdx.x; q < P.n_req; q += blockDim.x) {
    const double *t1 = P.T + (int64_t)P.tt1[q] * P.n_units + P.tr_off[p];
    const double *t2 = P.T + (int64_t)P.tt2[q] * P.n_units + P.tr_off[p];
    const int m = P.m[q];
    double sum = 0.;
    for (int i = 0; i < n; ++i) {
      const double step_k3 = step[i];
      if (step_k3 == 0) continue;                                                       // :495-496
      const double k3 = k3_grid[i];
      double resc = 1.;
      if (P.rescale && m != 0) {                                                        // :499, :505 (pow(x, 0) = 1 exactly)
        const double c = (k3 * k3 + k1 * k1 - k2 * k2) / 2. / k3 / k1;
        const double scale = pow(sqrt(1. - c * c), (double)m);
        resc = scale * scale;
      }
      const double total_factor = 2. * k1 * k2 * k3 * step_k1 * step_k2 * step_k3 * resc
                                * 2. / _PI_ / 2. / _PI_ / 2. / _PI_ / 2. / _PI_ * spectra_k1 * spectra_k2;   // :501-507
      sum += total_factor * t1[i] * t2[i];                                              // :509-511
    }
    P.partial[(int64_t)q * P.n_pairs + p] = sum;
  }
}

__global__ void cl_final_sum_kernel(ClParams P)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P.n_req) return;
  double sum = 0.;
  for (int p = 0; p < P.n_pairs; ++p) sum += P.partial[(int64_t)q * P.n_pairs + p];   // pair order = the reference's k1, k2 loops
  P.out[q] = sum;
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

extern "C" int song_los_cl_sum(song_los_ctx *ctx, const song_los_cl_request *rq, double *out)
{
  if (!ctx || !rq || !out) return ctx_fail(ctx, "song_los_cl_sum: NULL argument");
  KCK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight) return ctx_fail(ctx, "song_los_cl_sum: submitted batches are still in flight; call song_los_wait first");
  song_los_ctx::Slot &S = ctx->slot[ctx->cur];
  if (!S.have_result) return ctx_fail(ctx, "song_los_cl_sum: no transfer functions on the device (song_los_compute first)");
  if (rq->n_requests < 1 || !rq->tt1 || !rq->tt2 || !rq->m || rq->k_size < 2 || !rq->k || !rq->primordial || !rq->pair_k1 || !rq->pair_k2)
    return ctx_fail(ctx, "song_los_cl_sum: empty request");
  const int np = S.hb.n_pairs;
  for (int q = 0; q < rq->n_requests; ++q)
    if (rq->tt1[q] < 0 || rq->tt1[q] >= ctx->cfg.tt2_size || rq->tt2[q] < 0 || rq->tt2[q] >= ctx->cfg.tt2_size)
      return ctx_fail(ctx, "song_los_cl_sum: transfer type outside [0,%d)", ctx->cfg.tt2_size);
  for (int p = 0; p < np; ++p)
    if (rq->pair_k1[p] < 0 || rq->pair_k1[p] >= rq->k_size || rq->pair_k2[p] < 0 || rq->pair_k2[p] > rq->pair_k1[p])
      return ctx_fail(ctx, "song_los_cl_sum: pair %d has (index_k1,index_k2)=(%d,%d), outside the k grid of size %d", p, rq->pair_k1[p], rq->pair_k2[p], rq->k_size);
  cudaStream_t st = ctx->stream;
  KCK(up(ctx->d_k3req_tt, rq->tt1, (size_t)rq->n_requests, st));
  KCK(up(ctx->d_k3req_row, rq->tt2, (size_t)rq->n_requests, st));
  KCK(up(ctx->d_k3req_has, rq->m, (size_t)rq->n_requests, st));
  KCK(up(ctx->d_cl_k, rq->k, (size_t)rq->k_size, st));
  KCK(up(ctx->d_cl_prim, rq->primordial, (size_t)rq->k_size, st));
  KCK(up(ctx->d_cl_k1, rq->pair_k1, (size_t)np, st));
  KCK(up(ctx->d_cl_k2, rq->pair_k2, (size_t)np, st));
  KCK(ctx->d_cl_step.reserve((size_t)S.hb.n_units));
  KCK(ctx->d_k3_out.reserve((size_t)rq->n_requests * (np + 1)));
  ClParams P;
  P.n_req = rq->n_requests; P.n_pairs = np;
  P.tt1 = ctx->d_k3req_tt.p; P.tt2 = ctx->d_k3req_row.p; P.m = ctx->d_k3req_has.p; P.rescale = rq->rescale_cmb_sources;
  P.k = ctx->d_cl_k.p; P.prim = ctx->d_cl_prim.p; P.k_size = rq->k_size;
  P.pair_k1 = ctx->d_cl_k1.p; P.pair_k2 = ctx->d_cl_k2.p;
  P.nt = S.d_nt; P.tr_off = S.d_tr_off; P.k3_tr = S.d_k3_tr.p; P.T = S.d_out.p; P.n_units = S.hb.n_units;
  P.step_k3 = ctx->d_cl_step.p; P.partial = ctx->d_k3_out.p; P.out = ctx->d_k3_out.p + (size_t)rq->n_requests * np;
  cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
  KCK(cudaEventRecord(e0, st));
  if (np > 0) {
    cl_step_k3_kernel<<<(unsigned)((np + 63) / 64), 64, 0, st>>>(P);
    KCK(cudaGetLastError());
    cl_pair_sum_kernel<<<(unsigned)np, 128, 0, st>>>(P);
    KCK(cudaGetLastError());
  }
  cl_final_sum_kernel<<<(unsigned)((rq->n_requests + 63) / 64), 64, 0, st>>>(P);
  KCK(cudaGetLastError());
  KCK(cudaEventRecord(e1, st));
  KCK(cudaMemcpyAsync(out, P.out, (size_t)rq->n_requests * sizeof(double), cudaMemcpyDeviceToHost, st));
  KCK(cudaStreamSynchronize(st));
  KCK(cudaEventElapsedTime(&ctx->k3_ms, e0, e1));
  return SONG_LOS_SUCCESS;
}

extern "C" float song_los_last_k3_ms(const song_los_ctx *ctx) { return ctx ? ctx->k3_ms : 0.f; }
