// los_fast.h -- interface between the C-ABI layer and the FAST (DMMA) path (internal).
#pragma once
#include <string>
#include <vector>
#include "../../include/song_los.h"
#include "los_common.h"

namespace songlos {

// host copies of the index tables the FAST layout is derived from
struct FastHostLayout {
  int has_T, has_E, has_B;
  int index_tt2_T, index_tt2_E, index_tt2_B;
  int index_tp2_T, index_tp2_E, index_tp2_B;
  int index_J_TT, index_J_EE, index_J_EB;
  int l_max_los_t, l_max_los_p;
  int l_size, m_size, L_size, n_transfers, xx_size;
  const int *l, *m, *lm_array, *lm_src;
  const unsigned char *block_nonzero;   // [J_size*L_size]
  int64_t n_src_total;
  int tau_size;
  int64_t n_pt_total;                   // coarse k3 points of the batch
  int max_np;                           // longest coarse k3 grid of the batch
  int xx_uniform; double xx_delta;      // xx[i] == i*xx_delta bit for bit (bessel2_get_xx_list's lin_space from 0)
  double rows_spanned;                  // table rows the batch's k3 range covers at a mid-run tau (k3 density = n_units / rows)
};

struct FastState;
FastState *fast_state_create();
void fast_state_destroy(FastState *);
void fast_invalidate_table(FastState *);
int fast_prepare(FastState *fs, const FastHostLayout &h, const ExactParams &P, cudaStream_t st, std::string &err);
int64_t fast_bytes_per_unit(const FastState *fs);
int64_t fast_held_bytes(const FastState *fs);
int fast_reserve(FastState *fs, int64_t n_units, int64_t n_src, int64_t n_pt, std::string &err);
constexpr int FAST_MAX_EVENTS = 32;     // 5 phase boundaries + 2 per (column block, K chunk) launch
constexpr int FAST_MAX_BLOCKS = 12;     // (field, m) column blocks
// Pipelined runs (song_los_submit): the rows [tt_lo, tt_lo + tt_n) of a column block are final once the block's last K
// chunk has run, so their device->host copy is enqueued on the copy stream right there, under the kernels of the next
// block, instead of after the whole batch.  fast_enqueue fills the ranges it has taken care of.
struct FastEarlyOut {
  cudaStream_t st_d2h; double *host; cudaEvent_t *ev;     // ev: [FAST_MAX_BLOCKS], caller-owned
  int n_ranges; int tt_lo[FAST_MAX_BLOCKS], tt_n[FAST_MAX_BLOCKS];
};
int fast_enqueue(FastState *fs, const FastHostLayout &h, const ExactParams &P, int device_sms, cudaStream_t st,
                 cudaEvent_t *ev /* [FAST_MAX_EVENTS] */, int *n_launches, int *n_chunks, FastEarlyOut *early, std::string &err);
int fast_collect(cudaEvent_t *ev, int n_chunks, song_los_timing *tm, std::string &err);

}  // namespace songlos
