/*
 * los_host.h -- host-side C pieces of the transfer2 drop-in that do not touch the GPU:
 * the k3 integration grids (kept on the host so that bispectra2.c / spectra2.c, which re-derive
 * them through transfer2_get_k3_list, see bit-identical values) and the trapezoid time grid.
 * Flat-argument forms; the struct-based SONG prototypes in transfer2_gpu.c forward to these.
 */
#ifndef SONG_LOS_HOST_H
#define SONG_LOS_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* the scalars transfer2_get_k3_size / _list read from ppr, ppr2, ptr2, pbs2, ppt2 */
typedef struct song_k3_params {
  double tau0, tau_rec, rs_rec;          /* ptr2->tau0, tau_rec, rs_rec                 */
  double k_first, k_last;                /* ppt2->k[0], ppt2->k[k_size-1]               */
  int    k3_extrapolation;               /* ppr->bispectra_k3_extrapolation (0 = none)  */
  double extra_k3_oscillations_left;     /* ppr->extra_k3_oscillations_left             */
  double extra_k3_oscillations_right;    /* ppr->extra_k3_oscillations_right            */
  int    k_sampling;                     /* ptr2->k_sampling: 0 bessel, 1 class/smart   */
  double q_linstep_song;                 /* ppr2->q_linstep_song                        */
  int    old_run_step;                   /* ppr->load_run && ppr2->old_run              */
  double smallest_allowed_variation;     /* ppr->smallest_allowed_variation             */
  double xx_max, xx_step;                /* pbs2->xx_max, pbs2->xx_step                 */
} song_k3_params;

typedef struct song_k3_range {
  double k_min, k_max;                   /* ptr2->k_min_k1k2, k_max_k1k2                */
  int k_size;                            /* ptr2->k_size_k1k2                           */
  int k_physical_start;                  /* ptr2->k_physical_start_k1k2                 */
  int k_physical_size;                   /* ptr2->k_physical_size_k1k2                  */
} song_k3_range;

/* reference: transfer2_get_k3_size, source/transfer2.c:1329-1551.  Returns 0 or 1 (+ message). */
int song_host_k3_size(const song_k3_params *p, const double *k3_pt, int k3_pt_size,
                      song_k3_range *out, char *errmsg, int errlen);
/* reference: transfer2_get_k3_list, source/transfer2.c:1566-1737 */
int song_host_k3_list(const song_k3_params *p, const double *k3_pt, int k3_pt_size,
                      const song_k3_range *range, double *k3, int *last_used_index_pt,
                      char *errmsg, int errlen);
/* all pairs at once (transfer2_get_k3_sizes, source/transfer2.c:1752-1834): fills ranges[] and
   returns k3_size_max / sum through the pointers */
int song_host_k3_sizes(const song_k3_params *p, int n_pairs, const int *k3_pt_size,
                       const int64_t *pt_off, const double *k3_pt, song_k3_range *ranges,
                       int *k3_size_max, int64_t *k3_size_sum, char *errmsg, int errlen);
int song_host_k3_lists(const song_k3_params *p, int n_pairs, const int *k3_pt_size,
                       const int64_t *pt_off, const double *k3_pt, const song_k3_range *ranges,
                       const int64_t *tr_off, double *k3_tr, char *errmsg, int errlen);

/* Static longest-processing-time schedule of `n` work items (k1 batches, or (k1,k2) pairs) over `n_dev`
   GPUs: items by decreasing cost (ties: lower index first), each to the least loaded device (ties: lower
   device).  cost[i] = k3 values of item i x tt2_size x tau_size x (L_max+1) up to a common factor
   (SURVEY.md 8e).  Writes device[i] and, when load != NULL, the per-device totals.  Returns 0, or 1 on
   bad arguments. */
int song_host_lpt_schedule(const int64_t *cost, int n, int n_dev, int *device, int64_t *load);

/* The static multi-GPU schedule of the (k1,k2) pairs (north_star: "partitioned across the GPUs by (k1,k2) pair with a
   load-balanced static schedule"): the pair sequence in reference order is cut into n_dev * chunks_per_dev contiguous
   chunks of about equal cost, LPT assigns the chunks.  Writes device_of_pair[p] and, when load != NULL, the per-device
   cost.  Used by transfer2_init (transfer2_gpu.c) and by the multi-rank driver (song_b200/shard.py). */
int song_host_schedule_pairs(const int64_t *cost, int n_pairs, int n_dev, int chunks_per_dev, int *device_of_pair, int64_t *load);

/* Cost of every (k1,k2) pair for the static schedule: k3 values x (fixed part + fraction of the multipole columns that
   are active at the pair's k3), in 1/1024 k3 value.  k_min / k_max: ptr2->k_min_k1k2 / k_max_k1k2 flattened in pair order
   (NULL: every column counted).  l_max: largest multipole of the run. */
int song_host_pair_costs(int n_pairs, const int *k3_tr_size, const double *k_min, const double *k_max, double tau0, int l_max,
                         int64_t *cost);

/* ---- k3 bands: the second axis of the multi-GPU schedule ----
   The tensor-core kernel works on the k3-sorted values of a batch and its efficiency follows the number of values that
   share a row of the projection table.  Splitting a run by pairs alone divides that density by the number of GPUs;
   splitting the k3 axis into bands keeps it: a band holds, of EVERY pair of its group, the fine-grid values in
   [edge[r], edge[r+1]) -- fewer table rows for the same number of values.  The price is that the coarse sources of a
   pair are needed by every band it crosses (the k3 spline is global in the pair).  n_dev GPUs = n_groups pair groups
   (song_host_schedule_pairs) x n_bands bands.

   song_host_band_edges: the n_bands - 1 interior edges that split the predicted time of the selected pairs evenly over the
   bands, under the measured cost model of los_host.c (a k3 value costs more where few values share a table row and where
   many multipole columns are active; a coarse point costs its spline and its staging on every band its pair crosses).
   rows_per_k = (tau0 - a mid-run tau) / xx_step converts k3 into rows of the projection table.  pair_in_group: NULL = all
   pairs, else only pairs with pair_in_group[p] != 0 are counted.
   song_host_pair_costs_model: the same model per pair, for the split of the pairs into n_groups groups
   (song_host_schedule_pairs); units of 1/1024 microsecond.
   song_host_band_cuts: cut[p * (n_bands + 1) + r] = first fine-grid index of pair p with k3 >= edge[r]
   (cut[.. + 0] = 0, cut[.. + n_bands] = k3_tr_size[p]); grids are ascending (transfer2.c:1727). */
int song_host_band_edges(int n_pairs, const int *k3_tr_size, const int64_t *tr_off, const double *k3_tr, const int *k3_pt_size,
                         const unsigned char *pair_in_group, double tau0, int l_max, double rows_per_k, int n_bands, double *edges);
int song_host_pair_costs_model(int n_pairs, const int *k3_tr_size, const int64_t *tr_off, const double *k3_tr, const int *k3_pt_size,
                               double tau0, int l_max, double rows_per_k, int n_groups, int64_t *cost);
int song_host_band_cuts(int n_pairs, const int *k3_tr_size, const int64_t *tr_off, const double *k3_tr,
                        int n_bands, const double *edges, int *cut);
/* how many k3 bands for n_dev GPUs: the smallest divisor of n_dev that keeps `dense_per_row` fine-grid values per row of
   the projection table on every GPU (total_units / n_dev * bands / xx_size), 1 when pairs alone are dense enough */
int song_host_band_count(int64_t total_units, int xx_size, int n_dev, int dense_per_row);

#ifdef __cplusplus
}
#endif
#endif
