// song_los_ctx.h -- the engine context shared by the C-ABI translation units (internal).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/song_los.h"
#include "los_common.h"
#include "los_fast.h"

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
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

struct song_los_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::string err;

  // configuration (host copies)
  song_los_config cfg{};
  std::vector<double> tau, xx;
  int xx_uniform = 0; double xx_delta = 0.;   // xx[i] == i*xx_delta bit for bit
  std::vector<int> l, m, lm_array, lm_src, tt_index_l, tt_index_m;
  std::vector<int> tt_type, tt_lm, tt_l, tt_m, tt_im;

  // static device tables
  DevBuf<double> d_tau, d_tau0_minus_tau, d_delta_tau, d_xx;
  DevBuf<int> d_tt_type, d_tt_lm, d_tt_l, d_tt_m, d_tt_im, d_lm_src;
  DevBuf<double2> d_J, d_ddJ;
  DevBuf<unsigned char> d_block_nonzero;
  int64_t n_cols = 0;
  bool have_proj = false;
  // ragged mirror kept for song_los_get_projection
  DevBuf<double> d_ragged; DevBuf<int64_t> d_col_off; DevBuf<int> d_col_xmin, d_col_xsize;
  std::vector<int> h_col_xmin, h_col_xsize; std::vector<int64_t> h_col_off;

  // batches: three slots, so that the host->device copy of one batch and the device->host copy of
  // another overlap the kernels of a third stage (song_los_submit / song_los_wait)
  struct Slot {
    song_los_batch hb{};          // scalar fields only are valid
    unsigned char *h_meta = nullptr; size_t h_meta_cap = 0;     // pinned staging of the per-pair arrays
    DevBuf<unsigned char> d_meta;                               // packed per-pair arrays on the device
    int *d_np = nullptr, *d_nt = nullptr, *d_phys_start = nullptr, *d_phys_size = nullptr, *d_tile_off = nullptr;
    int64_t *d_pt_off = nullptr, *d_tr_off = nullptr, *d_src_off = nullptr;
    double *d_phys_k_first = nullptr, *d_phys_k_last = nullptr;
    DevBuf<double> d_k3_pt, d_k3_tr, d_src, d_out;
    DevBuf<int2> d_tiles; int n_tiles = 0; bool tiles_ready = false;
    int max_np = 0;               // largest coarse k3 grid of the batch
    double k3_max = 0.;           // largest k3 of the batch (custom/bessel time sampling: grid size bound)
    cudaEvent_t ev_h2d = nullptr, ev_compute = nullptr, ev_d2h = nullptr, ev_h2d_start = nullptr, ev_d2h_start = nullptr;
    cudaEvent_t ev_t[songlos::FAST_MAX_EVENTS]{};   // phase boundaries of this slot's compute
    cudaEvent_t ev_blk[songlos::FAST_MAX_BLOCKS]{}; // column blocks whose rows travel home early (song_los_submit)
    songlos::FastEarlyOut early{};                  // rows already enqueued for the device->host copy
    int n_chunks = 0;             // (column block, K chunk) launches of the last FAST compute
    int *h_err = nullptr;         // pinned copy of the device error flags after this slot's compute
    bool have_batch = false, have_result = false, in_flight = false, fast_timed = false, with_d2h = false;
    int64_t staged_src = 0, staged_total = 0;   // song_los_stage_sources: leading source doubles already on their way to d_src
    song_los_timing t{};
  };
  static constexpr int N_SLOTS = 3;
  Slot slot[N_SLOTS];
  int cur = 0;                    // slot of the synchronous upload/compute/download calls and of the last submit
  int n_in_flight = 0, oldest = 0;
  int stage_slot = -1;            // slot song_los_stage_sources is filling for the next song_los_submit
  cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
  std::vector<unsigned char> h_block_nonzero;
  DevBuf<double> d_dd;
  DevBuf<double> d_spline_u;      // EXACT spline workspace for k3 grids longer than SPLINE_MAX_NP
  DevBuf<int> d_err;
  DevBuf<double> d_scratch;

  songlos::FastState *fast = nullptr;
  // device-built projection functions (song_los_build_projection)
  double *d_proj_rows = nullptr, *d_proj_ddrows = nullptr; int proj_rows_n = 0;
  std::vector<int> h_jidx_xmin, h_jidx_xsize;
  cudaEvent_t ev[8]{};

  // k3 integral of the intrinsic bispectrum (los_k3int.cu): j_l1 table and request buffers
  int jl1_rows = 0;
  DevBuf<double> d_jl1, d_ddjl1, d_jl1_xmin, d_k3req_coef, d_k3_r, d_k3_out;
  DevBuf<int64_t> d_jl1_off;
  DevBuf<int> d_jl1_size, d_k3req_tt, d_k3req_row, d_k3req_has;
  DevBuf<double> d_cl_k, d_cl_prim, d_cl_step;   // spectra2_cls sum: k grid, primordial spectrum, k3 steps
  DevBuf<int> d_cl_k1, d_cl_k2;
  float k3_ms = 0.f;
};


int ctx_fail(song_los_ctx *c, const char *fmt, ...);

