// song_los_api.cu -- C-ABI of the LOS engine (include/song_los.h): device memory, table
// construction, launches, timing.  No CPU fallback: every path needs a CUDA device.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/song_los.h"
#include "los_common.h"
#include "los_fast.h"
#include "song_los_ctx.h"

namespace songlos {
// los_exact.cu
cudaError_t launch_spline_k3(const Batch &b, int tau_size, int tp2_size, double *dd, double *u_scratch, int *err, cudaStream_t st);
cudaError_t launch_los_exact(const ExactParams &P, int n_tiles, cudaStream_t st);
cudaError_t launch_los_exact_resample(const ExactParams &P, int n_tiles, int n_ctas, cudaStream_t st);
int64_t exact_resample_scratch_doubles(int tau_size, int tp2_size, int tau_grid_max);
size_t exact_smem_bytes(int tau_chunk, int tp2_size);
// los_peaks.cu
cudaError_t launch_dfma_probe(double *out, int iters, int blocks, int threads, cudaStream_t st);
cudaError_t launch_dmma_probe(double *out, int iters, int blocks, int threads, cudaStream_t st);
}  // namespace songlos

using namespace songlos;

static std::string g_create_error;
static std::mutex g_mutex;

int ctx_fail(song_los_ctx *c, const char *fmt, ...)
{
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (c) c->err = buf; else { std::lock_guard<std::mutex> g(g_mutex); g_create_error = buf; }
  return SONG_LOS_FAILURE;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return ctx_fail(ctx, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

template <typename T>
static cudaError_t upload(DevBuf<T> &d, const T *h, size_t n, cudaStream_t st)
{
  cudaError_t e = d.reserve(n);
  if (e != cudaSuccess || n == 0) return e;
  return cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, st);
}

extern "C" int song_los_version(void) { return 200; }

extern "C" int song_los_device_count(int *n)
{
  if (!n) return SONG_LOS_FAILURE;
  *n = 0;
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) { *n = 0; return ctx_fail(nullptr, "song_los_device_count: %s; there is no CPU fallback", cudaGetErrorString(e)); }
  return SONG_LOS_SUCCESS;
}

// tau grid, tau0 - tau and the trapezoid measure of the sources time sampling (transfer2.c:2567-2580) on the device
static cudaError_t upload_time_tables(song_los_ctx *ctx)
{
  const song_los_config &c = ctx->cfg;
  std::vector<double> d(c.tau_size), w(c.tau_size);
  for (int i = 0; i < c.tau_size; ++i) d[i] = c.tau0 - ctx->tau[i];
  w[0] = ctx->tau[1] - ctx->tau[0];
  for (int i = 1; i < c.tau_size - 1; ++i) w[i] = ctx->tau[i + 1] - ctx->tau[i - 1];
  w[c.tau_size - 1] = ctx->tau[c.tau_size - 1] - ctx->tau[c.tau_size - 2];
  cudaStream_t st = ctx->stream;
  cudaError_t e = upload(ctx->d_tau, ctx->tau.data(), ctx->tau.size(), st);
  if (e == cudaSuccess) e = upload(ctx->d_tau0_minus_tau, d.data(), d.size(), st);
  if (e == cudaSuccess) e = upload(ctx->d_delta_tau, w.data(), w.size(), st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // d and w live on this stack frame
  return e;
}

extern "C" const char *song_los_last_error(const song_los_ctx *ctx)
{
  if (ctx) return ctx->err.c_str();
  return g_create_error.c_str();
}

extern "C" int song_los_create(const song_los_config *cfg, song_los_ctx **out)
{
  song_los_ctx *ctx = nullptr;
  if (!cfg || !out) return ctx_fail(nullptr, "song_los_create: NULL argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return ctx_fail(nullptr, "song_los_create: no CUDA device available (%s); there is no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev)
    return ctx_fail(nullptr, "song_los_create: device %d out of range (have %d)", cfg->device, ndev);
  if (cfg->tau_size < 2) return ctx_fail(nullptr, "song_los_create: tau_size=%d < 2", cfg->tau_size);
  if (cfg->xx_size < 2 || cfg->xx_step <= 0) return ctx_fail(nullptr, "song_los_create: bad xx grid");
  if (cfg->tt2_size < 1 || cfg->tp2_size < 1 || cfg->l_size < 1 || cfg->m_size < 1 || cfg->L_size < 1)
    return ctx_fail(nullptr, "song_los_create: empty index tables");
  if (cfg->tt2_size > 32 * EXACT_MAX_TT_PASSES)
    return ctx_fail(nullptr, "song_los_create: tt2_size=%d exceeds the supported %d", cfg->tt2_size, 32 * EXACT_MAX_TT_PASSES);
  if (cfg->l_max_los_t >= cfg->L_size || cfg->l_max_los_p >= cfg->L_size)
    return ctx_fail(nullptr, "song_los_create: l_max_los exceeds L_size-1");

  ctx = new song_los_ctx();
  ctx->device = cfg->device;
  ctx->cfg = *cfg;
  if (cudaSetDevice(ctx->device) != cudaSuccess) { delete ctx; return ctx_fail(nullptr, "cudaSetDevice failed"); }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, ctx->device);
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx; return ctx_fail(nullptr, "cudaStreamCreate failed");
  }
  ctx->stream = ctx->own_stream;
  for (auto &ev : ctx->ev) cudaEventCreate(&ev);
  bool slots_ok = cudaStreamCreateWithFlags(&ctx->st_h2d, cudaStreamNonBlocking) == cudaSuccess
               && cudaStreamCreateWithFlags(&ctx->st_d2h, cudaStreamNonBlocking) == cudaSuccess;
  for (auto &S : ctx->slot) {
    slots_ok = slots_ok && cudaEventCreate(&S.ev_h2d) == cudaSuccess && cudaEventCreate(&S.ev_compute) == cudaSuccess
            && cudaEventCreate(&S.ev_d2h) == cudaSuccess && cudaEventCreate(&S.ev_h2d_start) == cudaSuccess
            && cudaEventCreate(&S.ev_d2h_start) == cudaSuccess
            && cudaHostAlloc((void **)&S.h_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
    for (auto &e : S.ev_t) slots_ok = slots_ok && cudaEventCreate(&e) == cudaSuccess;
    if (S.h_err) *S.h_err = 0;
  }
  if (!slots_ok) { song_los_destroy(ctx); return ctx_fail(nullptr, "song_los_create: stream/event creation failed"); }

  ctx->tau.assign(cfg->tau_sampling, cfg->tau_sampling + cfg->tau_size);
  ctx->xx.assign(cfg->xx, cfg->xx + cfg->xx_size);
  // bessel2.c:976-981 builds xx with lin_space(0, xx_max, xx_size): xx[i] = i * delta.  If one of the candidate deltas
  // reproduces the whole grid BIT FOR BIT the FAST kernel recomputes xx[ix+1] instead of reading it (transfer2.c:2202).
  ctx->xx_uniform = 0;
  {
    const int n = cfg->xx_size;
    const double cand[3] = {cfg->xx_step, ctx->xx[1] - ctx->xx[0], (ctx->xx[n - 1] - ctx->xx[0]) / (n - 1)};
    for (double delta : cand) {
      bool same = (ctx->xx[0] == 0.);
      for (int i = 0; same && i < n; ++i) same = (ctx->xx[i] == (double)i * delta);
      if (same) { ctx->xx_uniform = 1; ctx->xx_delta = delta; break; }
    }
  }
  ctx->l.assign(cfg->l, cfg->l + cfg->l_size);
  ctx->m.assign(cfg->m, cfg->m + cfg->m_size);
  ctx->lm_array.assign(cfg->lm_array, cfg->lm_array + (size_t)cfg->l_size * cfg->m_size);
  ctx->lm_src.assign(cfg->lm_src, cfg->lm_src + (size_t)cfg->L_size * cfg->m_size);
  ctx->tt_index_l.assign(cfg->tt_index_l, cfg->tt_index_l + cfg->tt2_size);
  ctx->tt_index_m.assign(cfg->tt_index_m, cfg->tt_index_m + cfg->tt2_size);
  ctx->cfg.tau_sampling = ctx->tau.data(); ctx->cfg.xx = ctx->xx.data();
  ctx->cfg.l = ctx->l.data(); ctx->cfg.m = ctx->m.data();
  ctx->cfg.lm_array = ctx->lm_array.data(); ctx->cfg.lm_src = ctx->lm_src.data();
  ctx->cfg.tt_index_l = ctx->tt_index_l.data(); ctx->cfg.tt_index_m = ctx->tt_index_m.data();

  // per-transfer-type tables (the dispatch of transfer2_compute, transfer2.c:1885-2061)
  int ntt = cfg->tt2_size;
  ctx->tt_type.resize(ntt); ctx->tt_lm.resize(ntt); ctx->tt_l.resize(ntt);
  ctx->tt_m.resize(ntt); ctx->tt_im.resize(ntt);
  for (int tt = 0; tt < ntt; ++tt) {
    int type = -1, base = 0;
    if (cfg->has_source_T && tt >= cfg->index_tt2_T && tt < cfg->index_tt2_T + cfg->n_transfers) { type = 0; base = cfg->index_tt2_T; }
    else if (cfg->has_source_E && tt >= cfg->index_tt2_E && tt < cfg->index_tt2_E + cfg->n_transfers) { type = 1; base = cfg->index_tt2_E; }
    else if (cfg->has_source_B && tt >= cfg->index_tt2_B && tt < cfg->index_tt2_B + cfg->n_transfers) { type = 2; base = cfg->index_tt2_B; }
    if (type < 0) { song_los_destroy(ctx); return ctx_fail(nullptr, "song_los_create: transfer type %d belongs to no field", tt); }
    int il = cfg->tt_index_l[tt], im = cfg->tt_index_m[tt];
    if (il < 0 || il >= cfg->l_size || im < 0 || im >= cfg->m_size) {
      song_los_destroy(ctx); return ctx_fail(nullptr, "song_los_create: (index_l,index_m) of tt=%d out of bounds", tt);
    }
    ctx->tt_type[tt] = type;
    ctx->tt_lm[tt] = tt - base;
    ctx->tt_l[tt] = cfg->l[il];
    ctx->tt_m[tt] = cfg->m[im];
    ctx->tt_im[tt] = im;
    if (cfg->lm_array[(size_t)il * cfg->m_size + im] != tt - base) {
      song_los_destroy(ctx); return ctx_fail(nullptr, "song_los_create: lm_array inconsistent with corresponding_index at tt=%d", tt);
    }
  }

  cudaStream_t st = ctx->stream;
  bool ok = upload_time_tables(ctx) == cudaSuccess
         && upload(ctx->d_xx, ctx->xx.data(), ctx->xx.size(), st) == cudaSuccess
         && upload(ctx->d_tt_type, ctx->tt_type.data(), (size_t)ntt, st) == cudaSuccess
         && upload(ctx->d_tt_lm, ctx->tt_lm.data(), (size_t)ntt, st) == cudaSuccess
         && upload(ctx->d_tt_l, ctx->tt_l.data(), (size_t)ntt, st) == cudaSuccess
         && upload(ctx->d_tt_m, ctx->tt_m.data(), (size_t)ntt, st) == cudaSuccess
         && upload(ctx->d_tt_im, ctx->tt_im.data(), (size_t)ntt, st) == cudaSuccess
         && upload(ctx->d_lm_src, ctx->lm_src.data(), ctx->lm_src.size(), st) == cudaSuccess
         && ctx->d_err.reserve(1) == cudaSuccess
         && cudaMemsetAsync(ctx->d_err.p, 0, sizeof(int), st) == cudaSuccess
         && cudaStreamSynchronize(st) == cudaSuccess;
  if (!ok) { song_los_destroy(ctx); return ctx_fail(nullptr, "song_los_create: device allocation/upload failed: %s", cudaGetErrorString(cudaGetLastError())); }
  ctx->fast = fast_state_create();
  *out = ctx;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_destroy(song_los_ctx *ctx)
{
  if (!ctx) return SONG_LOS_SUCCESS;
  cudaSetDevice(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  ctx->d_tau.release(); ctx->d_tau0_minus_tau.release(); ctx->d_delta_tau.release(); ctx->d_xx.release();
  ctx->d_tt_type.release(); ctx->d_tt_lm.release(); ctx->d_tt_l.release(); ctx->d_tt_m.release();
  ctx->d_tt_im.release(); ctx->d_lm_src.release(); ctx->d_J.release(); ctx->d_ddJ.release();
  ctx->d_block_nonzero.release(); ctx->d_ragged.release(); ctx->d_col_off.release();
  ctx->d_col_xmin.release(); ctx->d_col_xsize.release();
  cudaDeviceSynchronize();
  for (auto &S : ctx->slot) {
    S.d_meta.release(); S.d_k3_pt.release(); S.d_k3_tr.release(); S.d_src.release(); S.d_out.release(); S.d_tiles.release();
    if (S.h_meta) cudaFreeHost(S.h_meta);
    if (S.h_err) cudaFreeHost(S.h_err);
    for (cudaEvent_t e : {S.ev_h2d, S.ev_compute, S.ev_d2h, S.ev_h2d_start, S.ev_d2h_start}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : S.ev_blk) if (e) cudaEventDestroy(e);
    for (auto &e : S.ev_t) if (e) cudaEventDestroy(e);
  }
  if (ctx->st_h2d) cudaStreamDestroy(ctx->st_h2d);
  if (ctx->st_d2h) cudaStreamDestroy(ctx->st_d2h);
  ctx->d_dd.release(); ctx->d_spline_u.release(); ctx->d_err.release(); ctx->d_scratch.release();
  ctx->d_jl1.release(); ctx->d_ddjl1.release(); ctx->d_jl1_xmin.release(); ctx->d_k3req_coef.release(); ctx->d_k3_r.release();
  ctx->d_k3_out.release(); ctx->d_jl1_off.release(); ctx->d_jl1_size.release(); ctx->d_k3req_tt.release();
  ctx->d_k3req_row.release(); ctx->d_k3req_has.release();
  ctx->d_cl_k.release(); ctx->d_cl_prim.release(); ctx->d_cl_step.release(); ctx->d_cl_k1.release(); ctx->d_cl_k2.release();
  if (ctx->fast) fast_state_destroy(ctx->fast);
  if (ctx->d_proj_rows) cudaFree(ctx->d_proj_rows);
  if (ctx->d_proj_ddrows) cudaFree(ctx->d_proj_ddrows);
  for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_set_stream(song_los_ctx *ctx, void *cuda_stream)
{
  if (!ctx) return SONG_LOS_FAILURE;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_set_time(song_los_ctx *ctx, double tau0, int tau_sampling_mode, double tau_linstep_song)
{
  if (!ctx) return SONG_LOS_FAILURE;
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight) return ctx_fail(ctx, "song_los_set_time: submitted batches are still in flight");
  if (tau_sampling_mode != SONG_LOS_TAU_SOURCES && tau_sampling_mode != SONG_LOS_TAU_CUSTOM && tau_sampling_mode != SONG_LOS_TAU_BESSEL)
    return ctx_fail(ctx, "song_los_set_time: unknown time sampling %d", tau_sampling_mode);
  ctx->cfg.tau0 = tau0; ctx->cfg.tau_sampling_mode = tau_sampling_mode; ctx->cfg.tau_linstep_song = tau_linstep_song;
  CK(upload_time_tables(ctx));
  return SONG_LOS_SUCCESS;
}

// Does the context describe the run `cfg` would create (everything except the device and the members song_los_set_time
// replaces)?  transfer2_init asks this before reusing an engine bessel2_init left on the device.
extern "C" int song_los_config_matches(const song_los_ctx *ctx, const song_los_config *b)
{
  if (!ctx || !b) return 0;
  const song_los_config &a = ctx->cfg;
  bool same = a.tau_size == b->tau_size && a.sources_k3_interpolation == b->sources_k3_interpolation
           && a.sources_time_interpolation == b->sources_time_interpolation && a.bessels_interpolation == b->bessels_interpolation
           && a.k3_extrapolation == b->k3_extrapolation && a.xx_size == b->xx_size && a.xx_step == b->xx_step
           && a.l_size == b->l_size && a.m_size == b->m_size && a.L_size == b->L_size && a.n_transfers == b->n_transfers
           && a.has_source_T == b->has_source_T && a.has_source_E == b->has_source_E && a.has_source_B == b->has_source_B
           && a.tp2_size == b->tp2_size && a.l_max_los_t == b->l_max_los_t && a.l_max_los_p == b->l_max_los_p
           && a.tt2_size == b->tt2_size && a.J_size == b->J_size && a.index_J_TT == b->index_J_TT
           && a.index_J_EE == b->index_J_EE && a.index_J_EB == b->index_J_EB;
  if (same && a.has_source_T) same = a.index_tp2_T == b->index_tp2_T && a.index_tt2_T == b->index_tt2_T;
  if (same && a.has_source_E) same = a.index_tp2_E == b->index_tp2_E && a.index_tt2_E == b->index_tt2_E;
  if (same && a.has_source_B) same = a.index_tp2_B == b->index_tp2_B && a.index_tt2_B == b->index_tt2_B;
  if (!same) return 0;
  auto eq = [](const auto &v, const auto *p) { for (size_t i = 0; i < v.size(); ++i) if (v[i] != p[i]) return false; return true; };
  return eq(ctx->tau, b->tau_sampling) && eq(ctx->xx, b->xx) && eq(ctx->l, b->l) && eq(ctx->m, b->m)
      && eq(ctx->lm_array, b->lm_array) && eq(ctx->lm_src, b->lm_src) && eq(ctx->tt_index_l, b->tt_index_l)
      && eq(ctx->tt_index_m, b->tt_index_m) ? 1 : 0;
}

extern "C" int song_los_select_slot(song_los_ctx *ctx, int slot)
{
  if (!ctx) return SONG_LOS_FAILURE;
  if (slot < 0 || slot >= song_los_ctx::N_SLOTS) return ctx_fail(ctx, "song_los_select_slot: slot %d outside [0,%d)", slot, (int)song_los_ctx::N_SLOTS);
  if (ctx->n_in_flight) return ctx_fail(ctx, "song_los_select_slot: submitted batches are still in flight");
  ctx->cur = slot;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_device_info(song_los_ctx *ctx, int64_t *free_bytes, int64_t *total_bytes,
                                    int *sm_count, char *name, int name_len)
{
  if (!ctx) return SONG_LOS_FAILURE;
  CK(cudaSetDevice(ctx->device));
  size_t f = 0, t = 0;
  CK(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  if (sm_count) *sm_count = ctx->sm_count;
  if (name && name_len > 0) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
    snprintf(name, (size_t)name_len, "%s", prop.name);
  }
  return SONG_LOS_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// Projection tables
// ---------------------------------------------------------------------------------------------
namespace songlos {
__global__ void densify_projection_kernel(const double *ragged, const int64_t *col_off,
                                          const int *col_xmin, const int *col_xsize,
                                          int64_t n_cols, int xx_size, double2 *dense)
{
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int ix = blockIdx.y;
  if (c >= n_cols) return;
  double2 v = make_double2(0., 0.);
  int xsize = col_xsize[c];
  int xmin = col_xmin[c];
  if (xsize > 0 && ix >= xmin) {                      // skip test of transfer2.c:2232
    int r = ix - xmin;
    const double *row = ragged + col_off[c];
    if (r < xsize) v.x = row[r];
    if (r + 1 < xsize) v.y = row[r + 1];
  }
  dense[(int64_t)ix * n_cols + c] = v;
}
}  // namespace songlos

// column c of the dense table <-> (index_J, L, lm);  lm <-> (index_l, index_m) through lm_array
static int densify(song_los_ctx *ctx, const std::vector<double> &flat, DevBuf<double2> &dense)
{
  cudaStream_t st = ctx->stream;
  CK(upload(ctx->d_ragged, flat.data(), flat.size(), st));
  CK(upload(ctx->d_col_off, ctx->h_col_off.data(), ctx->h_col_off.size(), st));
  CK(upload(ctx->d_col_xmin, ctx->h_col_xmin.data(), ctx->h_col_xmin.size(), st));
  CK(upload(ctx->d_col_xsize, ctx->h_col_xsize.data(), ctx->h_col_xsize.size(), st));
  CK(dense.reserve((size_t)ctx->cfg.xx_size * ctx->n_cols));
  dim3 grid((unsigned)((ctx->n_cols + 127) / 128), (unsigned)ctx->cfg.xx_size);
  densify_projection_kernel<<<grid, 128, 0, st>>>(ctx->d_ragged.p, ctx->d_col_off.p, ctx->d_col_xmin.p,
                                                  ctx->d_col_xsize.p, ctx->n_cols, ctx->cfg.xx_size, dense.p);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_set_projection(song_los_ctx *ctx, const double *const *J_rows,
                                       const double *const *ddJ_rows,
                                       const int *index_xmin_J, const int *x_size_J)
{
  if (!ctx || !J_rows || !index_xmin_J || !x_size_J) return ctx_fail(ctx, "song_los_set_projection: NULL argument");
  CK(cudaSetDevice(ctx->device));
  const song_los_config &c = ctx->cfg;
  if (c.bessels_interpolation == SONG_LOS_INTERP_CUBIC && !ddJ_rows)
    return ctx_fail(ctx, "song_los_set_projection: cubic Bessel interpolation needs ddJ rows");
  ctx->n_cols = (int64_t)c.J_size * c.L_size * c.n_transfers;
  ctx->h_col_off.assign((size_t)ctx->n_cols, 0);
  ctx->h_col_xmin.assign((size_t)ctx->n_cols, c.xx_size);
  ctx->h_col_xsize.assign((size_t)ctx->n_cols, 0);
  std::vector<unsigned char> block_nonzero((size_t)c.J_size * c.L_size, 0);

  // pass 1: sizes
  int64_t total = 0;
  std::vector<int64_t> src_jidx((size_t)ctx->n_cols, -1);
  for (int J = 0; J < c.J_size; ++J)
    for (int L = 0; L < c.L_size; ++L)
      for (int il = 0; il < c.l_size; ++il)
        for (int im = 0; im < c.m_size; ++im) {
          int lm = c.lm_array[(size_t)il * c.m_size + im];
          if (lm < 0) continue;
          size_t jidx = (((size_t)J * c.L_size + L) * c.l_size + il) * c.m_size + im;
          if (!J_rows[jidx]) continue;
          int xs = x_size_J[jidx], xm = index_xmin_J[jidx];
          if (xs < 1 || xm < 0 || xm + xs > c.xx_size)
            return ctx_fail(ctx, "song_los_set_projection: row (J=%d,L=%d,il=%d,im=%d) has xmin=%d size=%d outside xx_size=%d",
                        J, L, il, im, xm, xs, c.xx_size);
          int64_t col = ((int64_t)J * c.L_size + L) * c.n_transfers + lm;
          // a row holding the single value 0 at the last grid point is the reference's
          // "negligible everywhere" marker (bessel2.c:1378-1388): keep the column empty
          if (xs == 1 && xm == c.xx_size - 1) continue;
          ctx->h_col_off[(size_t)col] = total;
          ctx->h_col_xmin[(size_t)col] = xm;
          ctx->h_col_xsize[(size_t)col] = xs;
          src_jidx[(size_t)col] = (int64_t)jidx;
          total += xs;
          block_nonzero[(size_t)J * c.L_size + L] = 1;
        }
  std::vector<double> flat((size_t)total);
  for (int64_t col = 0; col < ctx->n_cols; ++col)
    if (src_jidx[(size_t)col] >= 0)
      std::memcpy(flat.data() + ctx->h_col_off[(size_t)col], J_rows[src_jidx[(size_t)col]],
                  (size_t)ctx->h_col_xsize[(size_t)col] * sizeof(double));
  if (densify(ctx, flat, ctx->d_J) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
  if (c.bessels_interpolation == SONG_LOS_INTERP_CUBIC) {
    for (int64_t col = 0; col < ctx->n_cols; ++col)
      if (src_jidx[(size_t)col] >= 0) {
        if (!ddJ_rows[src_jidx[(size_t)col]]) return ctx_fail(ctx, "song_los_set_projection: missing ddJ row");
        std::memcpy(flat.data() + ctx->h_col_off[(size_t)col], ddJ_rows[src_jidx[(size_t)col]],
                    (size_t)ctx->h_col_xsize[(size_t)col] * sizeof(double));
      }
    if (densify(ctx, flat, ctx->d_ddJ) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
    // restore the J values in the ragged mirror
    for (int64_t col = 0; col < ctx->n_cols; ++col)
      if (src_jidx[(size_t)col] >= 0)
        std::memcpy(flat.data() + ctx->h_col_off[(size_t)col], J_rows[src_jidx[(size_t)col]],
                    (size_t)ctx->h_col_xsize[(size_t)col] * sizeof(double));
    CK(upload(ctx->d_ragged, flat.data(), flat.size(), ctx->stream));
  }
  CK(upload(ctx->d_block_nonzero, block_nonzero.data(), block_nonzero.size(), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->h_block_nonzero = block_nonzero;
  fast_invalidate_table(ctx->fast);
  ctx->have_proj = true;
  return SONG_LOS_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// Batches.  Every batch lives in one of three slots (inputs + result on the device, pinned staging of
// the per-pair arrays on the host).  Three streams: host->device copies, kernels, device->host copies.
// A slot is reused only after song_los_wait has seen its device->host copy finish, so no cross-slot
// event waits are needed beyond  h2d -> compute -> d2h  inside a slot.
// ---------------------------------------------------------------------------------------------
namespace songlos {
// device error flags -> page-locked host word (device-accessible under unified addressing), and reset
__global__ void flush_err_kernel(int *d_err, int *h_err) { *h_err = *d_err; *d_err = 0; }

// CTA work list of the exact kernel, (pair, first k3 index), built on the device
__global__ void fill_tiles_kernel(const int *nt, const int *tile_off, int n_pairs, int2 *tiles)
{
  const int p = blockIdx.x;
  const int n = (nt[p] + EXACT_UNITS_PER_CTA - 1) / EXACT_UNITS_PER_CTA;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tiles[tile_off[p] + i] = make_int2(p, i * EXACT_UNITS_PER_CTA);
}
}  // namespace songlos

static int validate_batch(song_los_ctx *ctx, const song_los_batch *b)
{
  const song_los_config &c = ctx->cfg;
  if (b->n_pairs < 0) return ctx_fail(ctx, "song_los_upload_pairs: negative n_pairs");
  // consistency checks (the reference's class_test of transfer2.c:1766-1771 and layout checks)
  int64_t pt = 0, tr = 0, so = 0;
  if ((b->phys_k_first == nullptr) != (b->phys_k_last == nullptr))
    return ctx_fail(ctx, "song_los_upload_pairs: phys_k_first and phys_k_last come together");
  const bool sliced = (b->phys_k_first != nullptr);     // slices of fine grids: the physical range may lie outside them
  const int need = (c.sources_k3_interpolation == SONG_LOS_INTERP_CUBIC) ? 4 : 2;
  for (int p = 0; p < b->n_pairs; ++p) {
    int np = b->k3_pt_size[p], nt = b->k3_tr_size[p];
    if (np < need)
      return ctx_fail(ctx, "pair %d: cannot do interpolation in k3 with just k3_size=%d values. Increase k3_size.", p, np);
    if (nt < 1) return ctx_fail(ctx, "pair %d: empty k3 grid", p);
    if (b->phys_size[p] < 1 || (!sliced && (b->phys_start[p] < 0 || b->phys_start[p] + b->phys_size[p] > nt)))
      return ctx_fail(ctx, "pair %d: physical range [%d,+%d) outside the k3 grid of %d points", p, b->phys_start[p], b->phys_size[p], nt);
    if (c.k3_extrapolation == SONG_LOS_NO_K3_EXTRAPOLATION && !sliced && (b->phys_start[p] != 0 || b->phys_size[p] != nt))
      return ctx_fail(ctx, "pair %d: when no extrapolation is requested, all k's should be physical", p);
    if (b->pt_off[p] != pt || b->tr_off[p] != tr || b->src_off[p] != so)
      return ctx_fail(ctx, "pair %d: offsets are not the running sums of the sizes (packed layout expected)", p);
    pt += np; tr += nt; so += (int64_t)c.tp2_size * c.tau_size * np;
  }
  if (pt != b->n_pt_total || tr != b->n_units || so != b->n_src_total)
    return ctx_fail(ctx, "song_los_upload_pairs: totals do not match the per-pair sizes");
  if (b->n_units >= ((int64_t)1 << 31)) return ctx_fail(ctx, "song_los_upload_pairs: more than 2^31 k3 values in one batch");
  return SONG_LOS_SUCCESS;
}

// host->device copies of one batch into slot S, enqueued on the copy stream; records S.ev_h2d
static int enqueue_upload(song_los_ctx *ctx, song_los_ctx::Slot &S, const song_los_batch *b)
{
  const song_los_config &c = ctx->cfg;
  if (validate_batch(ctx, b) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
  const size_t n = (size_t)b->n_pairs;
  // packed per-pair arrays: 3 x int64, 2 x double, then 5 x int, one pinned staging buffer, one copy
  const size_t meta_bytes = n * (3 * sizeof(int64_t) + 2 * sizeof(double) + 5 * sizeof(int));
  if (meta_bytes > S.h_meta_cap) {
    if (S.h_meta) cudaFreeHost(S.h_meta);
    S.h_meta = nullptr; S.h_meta_cap = 0;
    CK(cudaHostAlloc((void **)&S.h_meta, std::max<size_t>(meta_bytes, 64), cudaHostAllocDefault));
    S.h_meta_cap = std::max<size_t>(meta_bytes, 64);
  }
  CK(S.d_meta.reserve(std::max<size_t>(meta_bytes, 64)));
  int64_t *h64 = reinterpret_cast<int64_t *>(S.h_meta);
  double *hkf = reinterpret_cast<double *>(S.h_meta + n * 3 * sizeof(int64_t)), *hkl = hkf + n;
  int *h32 = reinterpret_cast<int *>(S.h_meta + n * 5 * sizeof(int64_t));
  for (size_t p = 0; p < n; ++p) {       // k3 of the first / last physical point (transfer2.c:2743-2757)
    hkf[p] = b->phys_k_first ? b->phys_k_first[p] : b->k3_tr[b->tr_off[p] + b->phys_start[p]];
    hkl[p] = b->phys_k_last ? b->phys_k_last[p] : b->k3_tr[b->tr_off[p] + b->phys_start[p] + b->phys_size[p] - 1];
  }
  std::memcpy(h64, b->pt_off, n * 8); std::memcpy(h64 + n, b->tr_off, n * 8); std::memcpy(h64 + 2 * n, b->src_off, n * 8);
  std::memcpy(h32, b->k3_pt_size, n * 4); std::memcpy(h32 + n, b->k3_tr_size, n * 4);
  std::memcpy(h32 + 2 * n, b->phys_start, n * 4); std::memcpy(h32 + 3 * n, b->phys_size, n * 4);
  int64_t n_tiles = 0;
  S.max_np = 0;
  for (size_t p = 0; p < n; ++p) {
    S.max_np = std::max(S.max_np, b->k3_pt_size[p]);
    h32[4 * n + p] = (int)n_tiles;
    n_tiles += (b->k3_tr_size[p] + EXACT_UNITS_PER_CTA - 1) / EXACT_UNITS_PER_CTA;
  }
  if (n_tiles >= ((int64_t)1 << 31)) return ctx_fail(ctx, "song_los_upload_pairs: too many k3 tiles in one batch");
  S.n_tiles = (int)n_tiles; S.tiles_ready = false;
  int64_t *d64 = reinterpret_cast<int64_t *>(S.d_meta.p);
  int *d32 = reinterpret_cast<int *>(S.d_meta.p + n * 5 * sizeof(int64_t));
  S.d_phys_k_first = reinterpret_cast<double *>(d64 + 3 * n); S.d_phys_k_last = S.d_phys_k_first + n;
  S.d_pt_off = d64; S.d_tr_off = d64 + n; S.d_src_off = d64 + 2 * n;
  S.d_np = d32; S.d_nt = d32 + n; S.d_phys_start = d32 + 2 * n; S.d_phys_size = d32 + 3 * n; S.d_tile_off = d32 + 4 * n;

  CK(S.d_k3_pt.reserve((size_t)b->n_pt_total));
  CK(S.d_k3_tr.reserve((size_t)b->n_units));
  CK(S.d_src.reserve((size_t)b->n_src_total));
  CK(S.d_out.reserve((size_t)b->n_units * c.tt2_size));
  cudaStream_t st = ctx->st_h2d;
  int64_t staged = 0;                      // leading part of the sources already enqueued by song_los_stage_sources
  if (S.staged_total > 0) {
    if (S.staged_total != b->n_src_total)
      return ctx_fail(ctx, "song_los_submit: %lld source doubles were announced to song_los_stage_sources, the batch has %lld",
                      (long long)S.staged_total, (long long)b->n_src_total);
    staged = S.staged_src;
  } else {
    CK(cudaEventRecord(S.ev_h2d_start, st));
  }
  S.staged_src = S.staged_total = 0;
  if (meta_bytes) CK(cudaMemcpyAsync(S.d_meta.p, S.h_meta, meta_bytes, cudaMemcpyHostToDevice, st));
  if (b->n_pt_total) CK(cudaMemcpyAsync(S.d_k3_pt.p, b->k3_pt, (size_t)b->n_pt_total * 8, cudaMemcpyHostToDevice, st));
  if (b->n_units) CK(cudaMemcpyAsync(S.d_k3_tr.p, b->k3_tr, (size_t)b->n_units * 8, cudaMemcpyHostToDevice, st));
  if (b->n_src_total > staged)
    CK(cudaMemcpyAsync(S.d_src.p + staged, b->src + staged, (size_t)(b->n_src_total - staged) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(S.ev_h2d, st));

  // largest k3 of the batch: every pair's grid is ascending (transfer2.c:1727), so its last value is the pair's maximum
  S.k3_max = 0.;
  for (int p = 0; p < b->n_pairs; ++p)
    if (b->k3_tr_size[p] > 0) S.k3_max = std::max(S.k3_max, b->k3_tr[b->tr_off[p] + b->k3_tr_size[p] - 1]);
  S.hb = *b;
  S.hb.k3_pt_size = S.hb.k3_tr_size = S.hb.phys_start = S.hb.phys_size = nullptr;
  S.hb.pt_off = S.hb.tr_off = S.hb.src_off = nullptr;
  S.hb.k3_pt = S.hb.k3_tr = S.hb.src = nullptr;
  S.hb.phys_k_first = S.hb.phys_k_last = nullptr;
  S.have_batch = true;
  S.have_result = false;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_upload_pairs(song_los_ctx *ctx, const song_los_batch *b)
{
  if (!ctx || !b) return ctx_fail(ctx, "song_los_upload_pairs: NULL argument");
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight) return ctx_fail(ctx, "song_los_upload_pairs: %d submitted batches are still in flight; call song_los_wait first", ctx->n_in_flight);
  song_los_ctx::Slot &S = ctx->slot[ctx->cur];
  if (enqueue_upload(ctx, S, b) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
  CK(cudaEventSynchronize(S.ev_h2d));
  return SONG_LOS_SUCCESS;
}

static void fill_params(song_los_ctx *ctx, const song_los_ctx::Slot &S, ExactParams &P)
{
  const song_los_config &c = ctx->cfg;
  std::memset(&P, 0, sizeof P);
  P.proj.J = ctx->d_J.p; P.proj.ddJ = ctx->d_ddJ.p; P.proj.n_cols = ctx->n_cols;
  P.proj.block_nonzero = ctx->d_block_nonzero.p;
  P.ty.tt2_size = c.tt2_size; P.ty.n_transfers = c.n_transfers; P.ty.L_size = c.L_size;
  P.ty.m_size = c.m_size; P.ty.l_size = c.l_size;
  P.ty.tt_type = ctx->d_tt_type.p; P.ty.tt_lm = ctx->d_tt_lm.p; P.ty.tt_l = ctx->d_tt_l.p;
  P.ty.tt_m = ctx->d_tt_m.p; P.ty.tt_im = ctx->d_tt_im.p; P.ty.lm_src = ctx->d_lm_src.p;
  P.ty.index_J_TT = c.index_J_TT; P.ty.index_J_EE = c.index_J_EE; P.ty.index_J_EB = c.index_J_EB;
  P.ty.index_tp2_T = c.index_tp2_T; P.ty.index_tp2_E = c.index_tp2_E; P.ty.index_tp2_B = c.index_tp2_B;
  P.ty.l_max_los_t = c.l_max_los_t; P.ty.l_max_los_p = c.l_max_los_p; P.ty.tp2_size = c.tp2_size;
  P.tm.tau_size = c.tau_size; P.tm.tau = ctx->d_tau.p; P.tm.tau0_minus_tau = ctx->d_tau0_minus_tau.p;
  P.tm.delta_tau = ctx->d_delta_tau.p; P.tm.tau0 = c.tau0; P.tm.mode = c.tau_sampling_mode;
  P.tm.tau_linstep_song = c.tau_linstep_song;
  P.b.n_pairs = S.hb.n_pairs; P.b.np = S.d_np; P.b.nt = S.d_nt;
  P.b.phys_start = S.d_phys_start; P.b.phys_size = S.d_phys_size;
  P.b.pt_off = S.d_pt_off; P.b.tr_off = S.d_tr_off; P.b.src_off = S.d_src_off;
  P.b.phys_k_first = S.d_phys_k_first; P.b.phys_k_last = S.d_phys_k_last;
  P.b.k3_pt = S.d_k3_pt.p; P.b.k3_tr = S.d_k3_tr.p; P.b.src = S.d_src.p; P.b.dd = ctx->d_dd.p;
  P.b.n_units = S.hb.n_units; P.b.out = S.d_out.p;
  P.xx = ctx->d_xx.p; P.xx_size = c.xx_size; P.xx_step = c.xx_step;
  P.cubic_k3 = (c.sources_k3_interpolation == SONG_LOS_INTERP_CUBIC);
  P.cubic_J = (c.bessels_interpolation == SONG_LOS_INTERP_CUBIC);
  P.cubic_time = (c.sources_time_interpolation == SONG_LOS_INTERP_CUBIC);
  P.flat_extrap = (c.k3_extrapolation == SONG_LOS_FLAT_K3_EXTRAPOLATION);
  P.tiles = S.d_tiles.p;
  P.err = ctx->d_err.p;
}

static void fill_fast_layout(song_los_ctx *ctx, const song_los_ctx::Slot &S, FastHostLayout &h)
{
  const song_los_config &c = ctx->cfg;
  h = FastHostLayout{};
  h.has_T = c.has_source_T; h.has_E = c.has_source_E; h.has_B = c.has_source_B;
  h.index_tt2_T = c.index_tt2_T; h.index_tt2_E = c.index_tt2_E; h.index_tt2_B = c.index_tt2_B;
  h.index_tp2_T = c.index_tp2_T; h.index_tp2_E = c.index_tp2_E; h.index_tp2_B = c.index_tp2_B;
  h.index_J_TT = c.index_J_TT; h.index_J_EE = c.index_J_EE; h.index_J_EB = c.index_J_EB;
  h.l_max_los_t = c.l_max_los_t; h.l_max_los_p = c.l_max_los_p;
  h.l_size = c.l_size; h.m_size = c.m_size; h.L_size = c.L_size; h.n_transfers = c.n_transfers;
  h.xx_size = c.xx_size; h.l = c.l; h.m = c.m; h.lm_array = c.lm_array; h.lm_src = c.lm_src;
  h.block_nonzero = ctx->h_block_nonzero.data(); h.n_src_total = S.hb.n_src_total;
  h.tau_size = c.tau_size; h.n_pt_total = S.hb.n_pt_total;
  h.xx_uniform = ctx->xx_uniform; h.xx_delta = ctx->xx_delta;
  h.max_np = S.max_np;
  // table rows the batch's k3 values sweep at a mid-run time: x = k3 (tau0 - tau) / xx_step (transfer2.c:2195-2199)
  h.rows_spanned = (S.k3_max > 0. && c.xx_step > 0.)
                 ? S.k3_max * (c.tau0 - ctx->tau[c.tau_size / 2]) / c.xx_step + 1. : (double)c.xx_size;
}

static int report_device_errors(song_los_ctx *ctx, int h)
{
  if (h & ERR_X_OUT_OF_RANGE)
    return ctx_fail(ctx, "not enough J's: k*(tau0-tau) exceeds xx_max (transfer2.c:2172); increase l_max or decrease k_max");
  if (h & ERR_H_ZERO) return ctx_fail(ctx, "stop to avoid division by zero in the k3 interpolation (two equal k3 values, or fewer than 3 of them)");
  if (h & ERR_TAU_GRID_OVERFLOW) return ctx_fail(ctx, "time integration grid larger than the allocated maximum");
  if (h & ERR_NAN) return ctx_fail(ctx, "found nan in second-order transfer function");
  return SONG_LOS_SUCCESS;
}

// kernels of one batch on the compute stream, after the slot's upload; records S.ev_compute.  No host wait.
static int enqueue_compute(song_los_ctx *ctx, song_los_ctx::Slot &S, int mode, double *early_host = nullptr)
{
  S.early = songlos::FastEarlyOut{};
  if (!ctx->have_proj) return ctx_fail(ctx, "song_los_compute: projection functions not set (bessel2 first)");
  if (!S.have_batch) return ctx_fail(ctx, "song_los_compute: no pairs uploaded");
  const song_los_config &c = ctx->cfg;
  if (mode != SONG_LOS_MODE_EXACT && mode != SONG_LOS_MODE_FAST) return ctx_fail(ctx, "song_los_compute: unknown mode %d", mode);
  if (c.tau_sampling_mode != SONG_LOS_TAU_SOURCES) mode = SONG_LOS_MODE_EXACT;     // k-dependent time grids: exact kernels only
  song_los_timing &t = S.t;
  t = song_los_timing{};
  t.n_integrals = S.hb.n_units * c.tt2_size;
  {
    int64_t nT = 0, nP = 0;
    for (int tt = 0; tt < c.tt2_size; ++tt) (ctx->tt_type[tt] == 0 ? nT : nP)++;
    t.n_terms_upper = S.hb.n_units * c.tau_size * (nT * (c.l_max_los_t + 1) + nP * (c.l_max_los_p + 1));
  }
  cudaStream_t st = ctx->stream;
  CK(cudaStreamWaitEvent(st, S.ev_h2d, 0));
  if (mode == SONG_LOS_MODE_EXACT && c.sources_k3_interpolation == SONG_LOS_INTERP_CUBIC)
    CK(ctx->d_dd.reserve((size_t)S.hb.n_src_total));
  ExactParams P;
  fill_params(ctx, S, P);
  S.fast_timed = false;
  if (mode == SONG_LOS_MODE_EXACT) {
    CK(cudaEventRecord(S.ev_t[0], st));
    if (!S.tiles_ready && S.n_tiles > 0) {
      CK(S.d_tiles.reserve((size_t)S.n_tiles));
      P.tiles = S.d_tiles.p;
      fill_tiles_kernel<<<(unsigned)S.hb.n_pairs, 64, 0, st>>>(S.d_nt, S.d_tile_off, S.hb.n_pairs, S.d_tiles.p);
      CK(cudaGetLastError());
      S.tiles_ready = true; t.n_launches++;
    }
    if (P.cubic_k3) {
      double *u_scratch = nullptr;                 // only k3 grids longer than SPLINE_MAX_NP need it
      if (S.max_np > SPLINE_MAX_NP) { CK(ctx->d_spline_u.reserve((size_t)S.hb.n_src_total)); u_scratch = ctx->d_spline_u.p; }
      CK(launch_spline_k3(P.b, c.tau_size, c.tp2_size, ctx->d_dd.p, u_scratch, ctx->d_err.p, st));
      t.n_launches++;
    }
    CK(cudaEventRecord(S.ev_t[1], st));
    // tau chunk: as many time samples as fit in ~96 KB of shared memory
    size_t per_tau = exact_smem_bytes(1, c.tp2_size) - 16;
    int chunk = (int)std::min<size_t>((size_t)c.tau_size, (96 * 1024) / per_tau);
    if (chunk < 1) return ctx_fail(ctx, "song_los_compute: tp2_size=%d too large for the shared-memory stage", c.tp2_size);
    P.tau_chunk = chunk;
    if (c.tau_sampling_mode == SONG_LOS_TAU_SOURCES) {
      CK(launch_los_exact(P, S.n_tiles, st));
    } else {
      // upper bound on the points of a time grid, transfer2.c:185-203
      int gmax;
      if (c.tau_sampling_mode == SONG_LOS_TAU_CUSTOM) {
        if (!(c.tau_linstep_song > 0) || !(S.k3_max > 0)) return ctx_fail(ctx, "stopping to prevent segmentation fault (tau_step_max == 0)");
        double tau_step_min = 2 * 3.1415926535897932384626433832795 / S.k3_max * c.tau_linstep_song;
        gmax = c.tau_size + (int)std::ceil(ctx->tau[c.tau_size - 1] / tau_step_min) + 1;
      } else {
        gmax = c.tau_size + c.xx_size;
      }
      P.tau_grid_max = gmax;
      P.tau_chunk = (int)std::min<size_t>((size_t)gmax, (96 * 1024) / (per_tau + 8));
      const int n_ctas = std::min(S.n_tiles, ctx->sm_count * 2);
      P.scratch_stride = exact_resample_scratch_doubles(c.tau_size, c.tp2_size, gmax);
      CK(ctx->d_scratch.reserve((size_t)n_ctas * P.scratch_stride));
      P.scratch = ctx->d_scratch.p;
      CK(launch_los_exact_resample(P, S.n_tiles, n_ctas, st));
    }
    t.n_launches++;
    CK(cudaEventRecord(S.ev_t[2], st));
  } else {
    std::string e;
    FastHostLayout h;
    fill_fast_layout(ctx, S, h);
    songlos::FastEarlyOut *early = nullptr;
    if (early_host != nullptr) {
      for (cudaEvent_t &ev : S.ev_blk) if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      S.early.st_d2h = ctx->st_d2h; S.early.host = early_host; S.early.ev = S.ev_blk;
      early = &S.early;
    }
    if (fast_enqueue(ctx->fast, h, P, ctx->sm_count, st, S.ev_t, &t.n_launches, &S.n_chunks, early, e) != 0)
      return ctx_fail(ctx, "%s", e.c_str());
    S.fast_timed = true;
  }
  // error flags of this batch: written to the slot's page-locked word by a one-thread kernel, then cleared for the next
  // batch.  (Not a cudaMemcpyAsync: a device->host copy on the compute stream queues on the same copy engine as the
  // result rows already travelling home and stalled the next batch's kernels behind several GB of them: 87 ms per batch.)
  flush_err_kernel<<<1, 1, 0, st>>>(ctx->d_err.p, S.h_err);
  CK(cudaGetLastError());
  CK(cudaEventRecord(S.ev_compute, st));
  return SONG_LOS_SUCCESS;
}

// waits for the slot's kernels, fills the timing and converts device error flags
static int finish_compute(song_los_ctx *ctx, song_los_ctx::Slot &S)
{
  CK(cudaEventSynchronize(S.ev_compute));
  song_los_timing &t = S.t;
  if (S.fast_timed) {
    std::string e;
    if (fast_collect(S.ev_t, S.n_chunks, &t, e) != 0) return ctx_fail(ctx, "%s", e.c_str());
  } else {
    CK(cudaEventElapsedTime(&t.ms_spline, S.ev_t[0], S.ev_t[1]));
    CK(cudaEventElapsedTime(&t.ms_integrate, S.ev_t[1], S.ev_t[2]));
    CK(cudaEventElapsedTime(&t.ms_total, S.ev_t[0], S.ev_t[2]));
  }
  CK(cudaEventElapsedTime(&t.ms_h2d, S.ev_h2d_start, S.ev_h2d));
  if (*S.h_err) return report_device_errors(ctx, *S.h_err);
  S.have_result = true;
  return SONG_LOS_SUCCESS;
}

static int enqueue_download(song_los_ctx *ctx, song_los_ctx::Slot &S, double *transfer)
{
  const size_t U = (size_t)S.hb.n_units;
  const int tt2 = ctx->cfg.tt2_size;
  CK(cudaStreamWaitEvent(ctx->st_d2h, S.ev_compute, 0));
  const bool early = (S.early.n_ranges > 0 && S.early.host == transfer);
  CK(cudaEventRecord(S.ev_d2h_start, ctx->st_d2h));          // ms_d2h: the rows that were not sent early
  // rows the FAST path has not sent already (whole column blocks, enqueued as soon as they were final)
  std::vector<unsigned char> sent((size_t)tt2, 0);
  if (early)
    for (int k = 0; k < S.early.n_ranges; ++k)
      for (int tt = S.early.tt_lo[k]; tt < S.early.tt_lo[k] + S.early.tt_n[k] && tt < tt2; ++tt) sent[(size_t)tt] = 1;
  for (int tt = 0; tt < tt2 && U > 0;) {
    if (sent[(size_t)tt]) { ++tt; continue; }
    int end = tt;
    while (end < tt2 && !sent[(size_t)end]) ++end;
    CK(cudaMemcpyAsync(transfer + (size_t)tt * U, S.d_out.p + (size_t)tt * U, (size_t)(end - tt) * U * sizeof(double), cudaMemcpyDeviceToHost, ctx->st_d2h));
    tt = end;
  }
  CK(cudaEventRecord(S.ev_d2h, ctx->st_d2h));
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_compute(song_los_ctx *ctx, int mode, song_los_timing *tm)
{
  if (!ctx) return SONG_LOS_FAILURE;
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight) return ctx_fail(ctx, "song_los_compute: %d submitted batches are still in flight; call song_los_wait first", ctx->n_in_flight);
  song_los_ctx::Slot &S = ctx->slot[ctx->cur];
  if (enqueue_compute(ctx, S, mode) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
  if (finish_compute(ctx, S) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
  if (tm) *tm = S.t;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_download(song_los_ctx *ctx, double *transfer)
{
  if (!ctx || !transfer) return ctx_fail(ctx, "song_los_download: NULL argument");
  CK(cudaSetDevice(ctx->device));
  song_los_ctx::Slot &S = ctx->slot[ctx->cur];
  if (!S.have_result) return ctx_fail(ctx, "song_los_download: nothing computed yet");
  if (enqueue_download(ctx, S, transfer) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
  CK(cudaEventSynchronize(S.ev_d2h));
  return SONG_LOS_SUCCESS;
}

// The slot the next song_los_submit fills: the one after the newest batch in flight.
static int next_submit_slot(const song_los_ctx *ctx)
{
  if (ctx->stage_slot >= 0) return ctx->stage_slot;
  return (ctx->n_in_flight == 0) ? ctx->cur : (ctx->cur + 1) % song_los_ctx::N_SLOTS;
}

// Sources of the NEXT batch, part by part: the host packs ppt2->sources into its page-locked staging in pieces and hands
// every finished piece over at once, so that the host->device copy runs under the packing of the following pieces instead
// of after all of them.  The pieces must be consecutive from offset 0 of the batch's packed `src`; what has not been
// staged when song_los_submit is called is copied there.
extern "C" int song_los_stage_sources(song_los_ctx *ctx, const double *src_part, int64_t offset, int64_t n, int64_t n_src_total)
{
  if (!ctx || !src_part || offset < 0 || n < 0 || offset + n > n_src_total) return ctx_fail(ctx, "song_los_stage_sources: bad arguments");
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight >= song_los_ctx::N_SLOTS)
    return ctx_fail(ctx, "song_los_stage_sources: %d batches already in flight; call song_los_wait first", ctx->n_in_flight);
  const int is = next_submit_slot(ctx);
  song_los_ctx::Slot &S = ctx->slot[is];
  if (S.staged_total == 0) {
    if (offset != 0) return ctx_fail(ctx, "song_los_stage_sources: the first part must start at offset 0");
    CK(S.d_src.reserve((size_t)n_src_total));
    CK(cudaEventRecord(S.ev_h2d_start, ctx->st_h2d));
    S.staged_total = n_src_total; S.staged_src = 0;
    ctx->stage_slot = is;
  }
  if (n_src_total != S.staged_total || offset != S.staged_src)
    return ctx_fail(ctx, "song_los_stage_sources: parts must be consecutive (expected offset %lld of %lld, got %lld of %lld)",
                    (long long)S.staged_src, (long long)S.staged_total, (long long)offset, (long long)n_src_total);
  if (n) CK(cudaMemcpyAsync(S.d_src.p + offset, src_part, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->st_h2d));
  S.staged_src += n;
  return SONG_LOS_SUCCESS;
}

// Asynchronous path: enqueue upload, kernels and download of one batch and return.  Up to three batches
// may be in flight; they complete in submission order through song_los_wait.
extern "C" int song_los_submit(song_los_ctx *ctx, const song_los_batch *batch, int mode, double *transfer)
{
  if (!ctx || !batch || !transfer) return ctx_fail(ctx, "song_los_submit: NULL argument");
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight >= song_los_ctx::N_SLOTS)
    return ctx_fail(ctx, "song_los_submit: %d batches already in flight; call song_los_wait first", ctx->n_in_flight);
  const int is = next_submit_slot(ctx);
  ctx->stage_slot = -1;
  song_los_ctx::Slot &S = ctx->slot[is];
  if (enqueue_upload(ctx, S, batch) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
  if (enqueue_compute(ctx, S, mode, transfer) != SONG_LOS_SUCCESS) { cudaDeviceSynchronize(); return SONG_LOS_FAILURE; }
  if (enqueue_download(ctx, S, transfer) != SONG_LOS_SUCCESS) { cudaDeviceSynchronize(); return SONG_LOS_FAILURE; }
  S.in_flight = true;
  if (ctx->n_in_flight == 0) ctx->oldest = is;
  ctx->n_in_flight++;
  ctx->cur = is;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_wait(song_los_ctx *ctx, song_los_timing *tm)
{
  if (!ctx) return SONG_LOS_FAILURE;
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight == 0) return ctx_fail(ctx, "song_los_wait: nothing submitted");
  song_los_ctx::Slot &S = ctx->slot[ctx->oldest];
  S.in_flight = false;
  ctx->n_in_flight--;
  ctx->oldest = (ctx->oldest + 1) % song_los_ctx::N_SLOTS;
  if (finish_compute(ctx, S) != SONG_LOS_SUCCESS) { cudaDeviceSynchronize(); return SONG_LOS_FAILURE; }
  CK(cudaEventSynchronize(S.ev_d2h));
  CK(cudaEventElapsedTime(&S.t.ms_d2h, S.ev_compute, S.ev_d2h));     // what is left of the copy after the last kernel
  if (tm) *tm = S.t;
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_transfer_pairs(song_los_ctx *ctx, const song_los_batch *batch, int mode,
                                       double *transfer, song_los_timing *timing)
{
  if (ctx && ctx->n_in_flight) return ctx_fail(ctx, "song_los_transfer_pairs: submitted batches are still in flight; call song_los_wait first");
  if (song_los_submit(ctx, batch, mode, transfer) != SONG_LOS_SUCCESS) return SONG_LOS_FAILURE;
  return song_los_wait(ctx, timing);
}

// Pinned host memory for the batch inputs and the result: copies from/to it run asynchronously at
// full PCIe rate, pageable memory works too but serialises the copies with the host.
extern "C" int song_los_host_alloc(void **ptr, size_t bytes)
{
  if (!ptr) return SONG_LOS_FAILURE;
  *ptr = nullptr;
  cudaError_t e = cudaHostAlloc(ptr, std::max<size_t>(bytes, 1), cudaHostAllocPortable);
  if (e != cudaSuccess) { ctx_fail(nullptr, "song_los_host_alloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e)); *ptr = nullptr; return SONG_LOS_FAILURE; }
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_host_free(void *ptr)
{
  if (ptr && cudaFreeHost(ptr) != cudaSuccess) return SONG_LOS_FAILURE;
  return SONG_LOS_SUCCESS;
}

// Pre-allocates the device buffers of all slots and the intermediates of one compute for batches of up to
// the given sizes: cudaMalloc/cudaFree synchronise the whole device, so growing a buffer between pipelined
// batches would serialise the copies of one batch with the kernels of another.
extern "C" int song_los_reserve(song_los_ctx *ctx, int mode, int n_in_flight, int max_pairs, int64_t max_units, int64_t max_pt, int64_t max_src)
{
  if (!ctx) return SONG_LOS_FAILURE;
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_in_flight) return ctx_fail(ctx, "song_los_reserve: submitted batches are still in flight");
  const song_los_config &c = ctx->cfg;
  // one batch at a time stays in the current slot; any deeper pipeline cycles through all of them
  const int n_slots = (n_in_flight <= 1) ? 1 : (int)song_los_ctx::N_SLOTS;
  for (int is = 0; is < n_slots; ++is) {
    // the slots a pipeline of this depth cycles through start at the current one
    song_los_ctx::Slot &S = ctx->slot[(ctx->cur + is) % song_los_ctx::N_SLOTS];
    const size_t meta_bytes = std::max<size_t>((size_t)max_pairs * (3 * sizeof(int64_t) + 5 * sizeof(int)), 64);
    if (meta_bytes > S.h_meta_cap) {
      if (S.h_meta) cudaFreeHost(S.h_meta);
      S.h_meta = nullptr; S.h_meta_cap = 0;
      CK(cudaHostAlloc((void **)&S.h_meta, meta_bytes, cudaHostAllocDefault));
      S.h_meta_cap = meta_bytes;
    }
    CK(S.d_meta.reserve(meta_bytes));
    CK(S.d_k3_pt.reserve((size_t)max_pt)); CK(S.d_k3_tr.reserve((size_t)max_units));
    CK(S.d_src.reserve((size_t)max_src)); CK(S.d_out.reserve((size_t)max_units * c.tt2_size));
  }
  if (mode == SONG_LOS_MODE_FAST) {
    if (!ctx->have_proj) return ctx_fail(ctx, "song_los_reserve: projection functions not set");
    std::string e;
    song_los_ctx::Slot dummy;
    ExactParams P; fill_params(ctx, dummy, P);
    FastHostLayout h; fill_fast_layout(ctx, dummy, h);
    if (fast_prepare(ctx->fast, h, P, ctx->stream, e) != 0) return ctx_fail(ctx, "%s", e.c_str());
    if (fast_reserve(ctx->fast, max_units, max_src, max_pt, e) != 0) return ctx_fail(ctx, "song_los_reserve: %s", e.c_str());
  } else {
    if (c.sources_k3_interpolation == SONG_LOS_INTERP_CUBIC) CK(ctx->d_dd.reserve((size_t)max_src));
    for (auto &S : ctx->slot) CK(S.d_tiles.reserve((size_t)(max_units / EXACT_UNITS_PER_CTA + max_pairs + 1)));
  }
  return SONG_LOS_SUCCESS;
}

// Largest number of k3 values one batch may hold so that three batches in flight plus the intermediates
// of one compute fit in the free device memory (with a 10% margin).  per-unit bytes: result rows and
// fine grid of two slots + FAST intermediates; sources are accounted with the caller's ratio.
extern "C" int song_los_max_batch_units(song_los_ctx *ctx, int mode, double src_doubles_per_unit, int64_t *units)
{
  return song_los_max_resident_units(ctx, mode, src_doubles_per_unit, song_los_ctx::N_SLOTS, units);
}

// The same bound for `n_slots` batches of that size held at once (1..3): a run whose batches stay resident in the slots
// (bench.py's `value` leg, or a consumer that integrates over k3 on the device) needs fewer than the pipeline's three.
extern "C" int song_los_max_resident_units(song_los_ctx *ctx, int mode, double src_doubles_per_unit, int n_slots, int64_t *units)
{
  if (!ctx || !units) return ctx_fail(ctx, "song_los_max_batch_units: NULL argument");
  if (n_slots < 1 || n_slots > song_los_ctx::N_SLOTS) return ctx_fail(ctx, "song_los_max_resident_units: n_slots must be 1..%d", song_los_ctx::N_SLOTS);
  CK(cudaSetDevice(ctx->device));
  if (!ctx->have_proj) return ctx_fail(ctx, "song_los_max_batch_units: projection functions not set");
  const song_los_config &c = ctx->cfg;
  size_t fr = 0, tot = 0;
  CK(cudaMemGetInfo(&fr, &tot));
  // memory already held by the batch buffers is reusable
  size_t held = 0;
  for (auto &S : ctx->slot) held += (S.d_k3_pt.cap + S.d_k3_tr.cap + S.d_src.cap + S.d_out.cap) * 8 + S.d_tiles.cap * 8;
  held += ctx->d_dd.cap * 8;
  double per_unit = (double)n_slots * (8.0 * c.tt2_size + 8.0 + 8.0 * src_doubles_per_unit);
  if (mode == SONG_LOS_MODE_FAST) {
    std::string e;
    song_los_ctx::Slot dummy;
    ExactParams P; fill_params(ctx, dummy, P);
    FastHostLayout h; fill_fast_layout(ctx, dummy, h);
    if (fast_prepare(ctx->fast, h, P, ctx->stream, e) != 0) return ctx_fail(ctx, "%s", e.c_str());
    size_t fr2 = 0; CK(cudaMemGetInfo(&fr2, &tot)); fr = fr2;
    per_unit += (double)fast_bytes_per_unit(ctx->fast) + 16.0 * src_doubles_per_unit;   // S, sort arrays, {y,dd} sources
    held += (size_t)fast_held_bytes(ctx->fast);          // the previous batch's intermediates are reused (a second call of
                                                         // transfer2_init must find the capacity of the first)
  } else {
    per_unit += 8.0 * src_doubles_per_unit + 8.0 / EXACT_UNITS_PER_CTA;                 // dd, tiles
  }
  double budget = 0.90 * ((double)fr + (double)held);
  int64_t u = (int64_t)(budget / per_unit);
  *units = std::max<int64_t>(1, std::min<int64_t>(u, ((int64_t)1 << 31) - 1));
  return SONG_LOS_SUCCESS;
}

extern "C" uint64_t song_los_device_result(song_los_ctx *ctx, int64_t *n_doubles)
{
  if (!ctx || !ctx->slot[ctx->cur].have_result) { if (n_doubles) *n_doubles = 0; return 0; }
  if (n_doubles) *n_doubles = ctx->slot[ctx->cur].hb.n_units * ctx->cfg.tt2_size;
  return (uint64_t)(uintptr_t)ctx->slot[ctx->cur].d_out.p;
}

extern "C" int song_los_copy_result(song_los_ctx *ctx, uint64_t device_dst, int64_t n_doubles)
{
  if (!ctx || !device_dst) return ctx_fail(ctx, "song_los_copy_result: NULL argument");
  CK(cudaSetDevice(ctx->device));
  song_los_ctx::Slot &S = ctx->slot[ctx->cur];
  if (!S.have_result) return ctx_fail(ctx, "song_los_copy_result: nothing computed yet");
  int64_t n = S.hb.n_units * ctx->cfg.tt2_size;
  if (n_doubles != n) return ctx_fail(ctx, "song_los_copy_result: expected %lld doubles, got %lld", (long long)n, (long long)n_doubles);
  CK(cudaMemcpyAsync((void *)(uintptr_t)device_dst, S.d_out.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SONG_LOS_SUCCESS;
}

extern "C" int song_los_measure_peaks(song_los_ctx *ctx, double *dfma_tflops, double *dmma_tflops, double *copy_gbs)
{
  if (!ctx) return SONG_LOS_FAILURE;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (int which = 0; which < 2; ++which) {
    double *dst = which ? dmma_tflops : dfma_tflops;
    if (!dst) continue;
    const int threads = 256, blocks = ctx->sm_count * 8, iters = 4096;
    DevBuf<double> sink;
    CK(sink.reserve((size_t)blocks * threads));
    auto launch = [&](int it) { return which ? launch_dmma_probe(sink.p, it, blocks, threads, st) : launch_dfma_probe(sink.p, it, blocks, threads, st); };
    CK(launch(64));      // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(ctx->ev[3], st));
      CK(launch(iters));
      CK(cudaEventRecord(ctx->ev[4], st));
      CK(cudaEventSynchronize(ctx->ev[4]));
      float ms; CK(cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]));
      best = std::min(best, ms);
    }
    // DFMA: 16 chains x 2 flop per thread-iteration; DMMA: 8 MMAs x 512 flop per warp-iteration
    double flops = which ? 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32)
                         : 2.0 * 16.0 * iters * (double)blocks * threads;
    *dst = flops / (best * 1e-3) / 1e12;
    sink.release();
  }
  if (copy_gbs) {
    size_t n = (size_t)1 << 28;   // 2 GiB of doubles per buffer
    DevBuf<double> a, b;
    CK(a.reserve(n)); CK(b.reserve(n));
    CK(cudaMemsetAsync(a.p, 0, n * 8, st));
    CK(cudaMemcpyAsync(b.p, a.p, n * 8, cudaMemcpyDeviceToDevice, st));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(ctx->ev[3], st));
      CK(cudaMemcpyAsync(b.p, a.p, n * 8, cudaMemcpyDeviceToDevice, st));
      CK(cudaEventRecord(ctx->ev[4], st));
      CK(cudaEventSynchronize(ctx->ev[4]));
      float ms; CK(cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]));
      best = std::min(best, ms);
    }
    *copy_gbs = 2.0 * n * 8 / (best * 1e-3) / 1e9;
    a.release(); b.release();
  }
  return SONG_LOS_SUCCESS;
}

