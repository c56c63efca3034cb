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

  // batch
  song_los_batch hb{};          // scalar fields only are valid
  std::vector<int> h_np, h_nt, h_unit_pair;
  std::vector<int64_t> h_tr_off;
  std::vector<unsigned char> h_block_nonzero;
  DevBuf<int> d_np, d_nt, d_phys_start, d_phys_size;
  DevBuf<int64_t> d_pt_off, d_tr_off, d_src_off;
  DevBuf<double> d_k3_pt, d_k3_tr, d_src, d_dd, d_out;
  DevBuf<int2> d_tiles; int n_tiles = 0;
  DevBuf<int> d_err;
  DevBuf<double> d_scratch;
  bool have_batch = false, have_result = false;

  songlos::FastState *fast = nullptr;
  // device-built projection functions (song_los_build_projection)
  double *d_proj_rows = nullptr; int proj_rows_n = 0;
  std::vector<int> h_jidx_xmin, h_jidx_xsize;
  cudaEvent_t ev[8]{};
};


int ctx_fail(song_los_ctx *c, const char *fmt, ...);

