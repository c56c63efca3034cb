/*
 * song_los.h -- C-ABI of the B200 line-of-sight (LOS) engine behind SONG's transfer2 module.
 *
 * This is the "thin C-ABI layer" between SONG's host C code and the CUDA kernels: plain C,
 * plain pointers and sizes, no CUDA or torch types.  The drop-in host module
 * (song_b200/host/transfer2_gpu.c) implements SONG's own module API -- transfer2_init() and
 * friends, reference include/transfer2.h:327-541 -- on top of these entry points;
 * song_b200/host/bessel2_gpu.c does the same for SONG's bessel2 module (bessel2_init, bessel2_free,
 * bessel2_convolution, ...; reference include/bessel2.h:161-260).
 *
 * What each entry point replaces in the reference (coccoinomane/song @ 2ea44ecf):
 *
 *   song_los_create            per-run state built at transfer2.c:128-281 (tau0, workspaces) and
 *                              the index tables of transfer2_indices_of_transfers, :866-1067
 *   song_los_set_projection    consumes pbs2->J_Llm_x / index_xmin_J / x_size_J as filled by
 *                              bessel2_init, bessel2.c:315-374 (ragged -> dense device table)
 *   song_los_build_projection  the J_Llm(x) assembly itself, bessel2_J_for_Llm + bessel2_J_Llm,
 *                              bessel2.c:1041-1588, from a host j_l1 table and host 3j symbols, and for
 *                              cubic Bessel interpolation their second derivatives, bessel2.c:395-435
 *   song_los_get_projection    writes the device-built tables back in bessels2 layout
 *   song_los_upload_pairs      the inputs of the (k1,k2) loop body: ppt2->k3[k1][k2],
 *                              ppt2->sources[tp][k1][k2] (transfer2.h:47-51 layout) and the fine
 *                              grid of transfer2_get_k3_list, transfer2.c:1566-1737
 *   song_los_compute           transfer2.c:401-539 for every uploaded pair: k3 spline
 *                              (transfer2_interpolate_sources_in_k :2612), time grid (:2366),
 *                              time interpolation (:2783) and every transfer2_compute /
 *                              transfer2_integrate call (:1861, :2140)
 *   song_los_download          ptr2->transfer[tt][k1][k2][k3] for the uploaded pairs, in the
 *                              [tt][k1][k2][k3] order of transfers_%03d.dat (transfer2.c:2940-2957)
 *   song_los_transfer_pairs    upload + compute + download in one call with HOST buffers (the
 *                              end-to-end path a SONG run takes)
 *   song_los_submit / _wait    the same, asynchronous and double-buffered: the k1 loop of
 *                              transfer2_init (transfer2.c:315-570) as a pipeline in which the copy of
 *                              batch n+1 to the device and of batch n-1 to the host overlap the
 *                              kernels of batch n (the k1 levels are independent, transfer2.c:549-551)
 *   song_los_stage_sources     pieces of the next batch's packed sources, copied to the device while the host packs the
 *                              rest (consecutive from offset 0; song_los_submit copies what is left)
 *   song_los_host_alloc/_free  page-locked staging memory for ppt2->sources / ptr2->transfer slabs
 *   song_los_max_batch_units   how many k3 values of consecutive k1 levels fit in one batch (three batches in flight)
 *   song_los_max_resident_units  the same bound when only n_slots batches are held at once
 *   song_los_reserve           device buffers for the largest batch, allocated before the k1 loop starts
 *   song_los_device_count      how many GPUs the (k1,k2) pairs can be spread over (the k1 modes are independent,
 *                              transfer2.c:549-551)
 *   song_los_set_time          the members of the run that only transfer2_init knows: pba->conformal_age
 *                              (transfer2.c:128) and ptr2->tau_sampling (input2.c:1480-1500); lets bessel2_init
 *                              create the engines and transfer2_init reuse them with their tables in place
 *   song_los_drop_projection_rows   frees the device copy of pbs2->J_Llm_x once bessel2_init has read it back
 *   song_los_select_slot       which of the three batch slots the synchronous calls use (several batches resident)
 *   song_los_set_bessel_table / song_los_integrate_k3   the k3 integral of bispectra2_intrinsic_integrate_over_k3
 *                              (bispectra2.c:1279-1375 with bessel2_convolution, bessel2.c:499-605) on the transfer
 *                              functions still in device memory
 *   song_los_cl_sum            the (k1,k2,k3) sum of spectra2_cls (spectra2.c:241-570), likewise
 *
 * Error convention (reference: _SUCCESS_/_FAILURE_ + ErrorMsg): every function
 * returns SONG_LOS_SUCCESS (0) or SONG_LOS_FAILURE (1); the text is available from
 * song_los_last_error() and is what the host module snprintf's into ptr2->error_message.
 * There is NO CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef SONG_LOS_H
#define SONG_LOS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SONG_LOS_SUCCESS 0
#define SONG_LOS_FAILURE 1
#define SONG_LOS_MAX_IN_FLIGHT 3

/* mirrors of the reference enums, passed as ints */
enum { SONG_LOS_TAU_SOURCES = 0, SONG_LOS_TAU_CUSTOM = 1, SONG_LOS_TAU_BESSEL = 2 }; /* transfer2.h:27-32 */
enum { SONG_LOS_INTERP_LINEAR = 0, SONG_LOS_INTERP_CUBIC = 1 };
enum { SONG_LOS_NO_K3_EXTRAPOLATION = 0, SONG_LOS_FLAT_K3_EXTRAPOLATION = 1 };
enum { SONG_LOS_J_TT = 0, SONG_LOS_J_EE = 1, SONG_LOS_J_EB = 2 };                   /* bessel2.h:12-16 */

/* Arithmetic mode of song_los_compute.
   EXACT: one thread per (k3,l) walks tau and L in the reference's order with no FMA
          contraction; bit-identical to the reference statement order (and to the oracle).
   FAST : k3-sorted, register-tiled kernel; reassociates the L and tau sums and contracts FMAs
          (max relative deviation ~1e-13 of the per-(l,m) maximum, tolerance stated in tests).
          Requires the sources time sampling and linear Bessel interpolation; custom / bessel
          time sampling (k-dependent grids) is computed by the EXACT kernels in either mode. */
enum { SONG_LOS_MODE_EXACT = 0, SONG_LOS_MODE_FAST = 1 };

/*
 * Run-wide description.  Flattening of every [J][L][l][m] table (bessel2.h:64-91):
 *     jidx = ((index_J*L_size + index_L)*l_size + index_l)*m_size + index_m
 * Entries with |m| > min(L,l) are not allocated by the reference: pass NULL rows for them.
 * All arrays are copied; the caller may free them after song_los_create returns.
 */
typedef struct song_los_config {
  int device;                      /* CUDA device ordinal */

  double tau0;                     /* pba->conformal_age, transfer2.c:128 */
  int tau_size;                    /* ppt2->tau_size */
  const double *tau_sampling;      /* ppt2->tau_sampling */
  int tau_sampling_mode;           /* ptr2->tau_sampling */
  double tau_linstep_song;         /* ppr2->tau_linstep_song (custom sampling only) */

  int sources_k3_interpolation;    /* ppr2->sources_k3_interpolation */
  int sources_time_interpolation;  /* ppr2->sources_time_interpolation */
  int bessels_interpolation;       /* ppr->bessels_interpolation */
  int k3_extrapolation;            /* ppr->bispectra_k3_extrapolation */

  int xx_size;                     /* pbs2->xx_size */
  const double *xx;                /* pbs2->xx (used verbatim: transfer2.c:2202) */
  double xx_step;                  /* pbs2->xx_step */

  int l_size;  const int *l;       /* ptr2->l */
  int m_size;  const int *m;       /* ptr2->m */
  int L_size;                      /* pbs2->L_size, L[i] = i */
  int n_transfers;                 /* ptr2->n_transfers */
  const int *lm_array;             /* ptr2->lm_array as [l_size][m_size], -1 where m > l */
  const int *lm_src;               /* lm(L,m) of perturbations2_macros.h:53 as [L_size][m_size], -1 where m > L */

  int has_source_T, has_source_E, has_source_B;     /* ppt2->has_source_X */
  int index_tp2_T, index_tp2_E, index_tp2_B;        /* ppt2->index_tp2_X  */
  int tp2_size;                                     /* ppt2->tp2_size     */
  int l_max_los_t, l_max_los_p;                     /* ppr2->l_max_los_t/p */

  int index_tt2_T, index_tt2_E, index_tt2_B;        /* ptr2->index_tt2_X  */
  int tt2_size;                                     /* ptr2->tt2_size     */
  const int *tt_index_l;           /* ptr2->corresponding_index_l [tt2_size] */
  const int *tt_index_m;           /* ptr2->corresponding_index_m [tt2_size] */

  int J_size;                      /* pbs2->J_size */
  int index_J_TT, index_J_EE, index_J_EB;           /* -1 when the table is absent */
} song_los_config;

/* One batch of (k1,k2) pairs, packed.  Pair p:
 *   coarse grid   k3_pt + pt_off[p],  k3_pt_size[p] doubles      (ppt2->k3[k1][k2])
 *   sources       src + src_off[p] + tp*tau_size*np + itau*np + j (ppt2->sources[tp][k1][k2])
 *   fine grid     k3_tr + tr_off[p],  k3_tr_size[p] doubles      (transfer2_get_k3_list)
 *   physical range [phys_start[p], phys_start[p]+phys_size[p])   (ptr2->k_physical_*_k1k2)
 * Output unit index of (p, i) is tr_off[p] + i; the transfer array is [tt2_size][n_units].
 *
 * A batch may hold only a SLICE [a, b) of a pair's fine grid (the k3 bands of the multi-GPU schedule,
 * song_host_band_edges): k3_tr then holds the b-a values of the slice, phys_start[p] is given relative to the
 * slice (ptr2->k_physical_start_k1k2 - a: it may be negative or beyond the slice), and phys_k_first / phys_k_last
 * carry the k3 of the pair's first / last physical point, which the flat extrapolation of transfer2.c:2743-2757
 * interpolates the sources at and which may lie outside the slice.  Both are NULL for whole grids. */
typedef struct song_los_batch {
  int n_pairs;
  const int *k3_pt_size;
  const int *k3_tr_size;
  const int *phys_start;
  const int *phys_size;
  const int64_t *pt_off;
  const int64_t *tr_off;
  const int64_t *src_off;
  int64_t n_pt_total;              /* sum of k3_pt_size */
  int64_t n_units;                 /* sum of k3_tr_size */
  int64_t n_src_total;             /* doubles in src */
  const double *k3_pt;
  const double *k3_tr;
  const double *src;
  const double *phys_k_first;      /* optional [n_pairs]: k3 of the first physical point of the pair's FULL grid */
  const double *phys_k_last;       /* optional [n_pairs]: k3 of the last one; both or neither */
} song_los_batch;

/* Device-side timings of the last song_los_compute, in milliseconds (CUDA events on the
   engine's stream), plus exact work counters. */
typedef struct song_los_timing {
  float ms_total;
  float ms_spline;                 /* k3 second derivatives (transfer2.c:2639)        */
  float ms_sort;                   /* FAST: k3 ordering of the units                  */
  float ms_interp;                 /* FAST: source interpolation into the sorted slab */
  float ms_integrate;              /* the LOS kernel proper                           */
  float ms_reorder;                /* FAST: [unit][tt] -> [tt][unit]                  */
  int   n_launches;                /* kernels launched by the call                    */
  int64_t n_integrals;             /* stored T values = n_units * tt2_size            */
  int64_t n_terms_upper;           /* n_integrals * tau_size * (L_max+1) upper bound  */
  float ms_h2d;                    /* host->device copy of the batch (copy stream)    */
  float ms_d2h;                    /* device->host copy of the result that follows the last kernel (song_los_wait);
                                      column blocks finished earlier are already on their way (FAST mode) */
} song_los_timing;

typedef struct song_los_ctx song_los_ctx;

/* Number of CUDA devices visible to the process (0 and SONG_LOS_FAILURE when there is none). */
int  song_los_device_count(int *n_devices);

int  song_los_create(const song_los_config *cfg, song_los_ctx **ctx);
/* Replaces cfg.tau0, cfg.tau_sampling_mode and cfg.tau_linstep_song of an existing context (the time tables are
   rebuilt; projection tables, buffers and the tensor-core table stay). */
int  song_los_set_time(song_los_ctx *ctx, double tau0, int tau_sampling_mode, double tau_linstep_song);
/* 1 when `ctx` was created from a configuration equal to `cfg` in everything but the device and the three members
   song_los_set_time replaces, 0 otherwise. */
int  song_los_config_matches(const song_los_ctx *ctx, const song_los_config *cfg);
int  song_los_destroy(song_los_ctx *ctx);
const char *song_los_last_error(const song_los_ctx *ctx);   /* ctx may be NULL: last create error */

/* ragged host tables -> dense device table.  ddJ_rows may be NULL unless bessels_interpolation
   is cubic.  Arrays have J_size*L_size*l_size*m_size entries. */
int  song_los_set_projection(song_los_ctx *ctx, const double *const *J_rows,
                             const double *const *ddJ_rows,
                             const int *index_xmin_J, const int *x_size_J);

/* Device assembly of J_Llm(x) (bessel2.c:1041-1588).
 *   j_l1 rows: j_l1[i] has x_size_l1[i] doubles starting at grid index index_xmin_l1[i]
 *   three_j  : per jidx, first_3j and second_3j as [n_jidx][2*L_max+1] (zero padded),
 *              l1_min[jidx], l1_size[jidx]; l1_size[jidx] = 0 marks a vanishing table
 *              (spin / EB constraint, bessel2.c:1090-1100)
 *   index_l1_of_l1min[jidx]: position of l1_min in the l1 list (bessel2.c:1175)           */
int  song_los_build_projection(song_los_ctx *ctx, int l1_size, const int *l1,
                               const double *const *j_l1, const int *index_xmin_l1,
                               const int *x_size_l1,
                               int n3j, const double *first_3j, const double *second_3j,
                               const int *l1_min, const int *l1_count, const int *index_l1_of_l1min,
                               const int *J_type_of_index_J, double J_Llm_cut, double xx_max);

/* Sizes then contents of the device-built tables, bessels2 layout.  Call with rows == NULL to
   obtain index_xmin_J / x_size_J first, allocate, then call again with rows. */
int  song_los_get_projection(song_los_ctx *ctx, int *index_xmin_J, int *x_size_J,
                             double *const *J_rows);
/* pbs2->ddJ_Llm_x of the device-built tables (cubic Bessel interpolation only; bessel2.c:395-435),
   rows sized and offset like J_rows. */
int  song_los_get_projection_dd(song_los_ctx *ctx, double *const *ddJ_rows);

/* Frees the undensified device rows behind song_los_get_projection / _dd (the dense tables the kernels read stay). */
int  song_los_drop_projection_rows(song_los_ctx *ctx);

/* The synchronous calls below work on one of SONG_LOS_MAX_IN_FLIGHT batch slots (default 0): a caller that keeps
   several batches resident on the device uploads each into its own slot and selects it before compute / download. */
int  song_los_select_slot(song_los_ctx *ctx, int slot);

int  song_los_upload_pairs(song_los_ctx *ctx, const song_los_batch *batch);
int  song_los_compute(song_los_ctx *ctx, int mode, song_los_timing *timing);
int  song_los_download(song_los_ctx *ctx, double *transfer /* [tt2_size][n_units] */);
int  song_los_transfer_pairs(song_los_ctx *ctx, const song_los_batch *batch, int mode,
                             double *transfer, song_los_timing *timing);

/* Asynchronous pipeline.  song_los_submit validates the batch, enqueues its host->device copy, its
 * kernels and the device->host copy of the result into `transfer` on three CUDA streams and returns.
 * At most SONG_LOS_MAX_IN_FLIGHT (3) batches may be in flight -- one being copied in, one in the kernels, one
 * being copied out; song_los_wait blocks until the OLDEST one is complete (result
 * in `transfer`, device error flags converted to the error convention) and reports its timing.
 * The batch arrays and `transfer` must stay valid until the matching song_los_wait returns; use
 * song_los_host_alloc memory for k3_pt / k3_tr / src / transfer so that the copies overlap.
 * The synchronous calls above refuse to run while submissions are in flight. */
int  song_los_submit(song_los_ctx *ctx, const song_los_batch *batch, int mode, double *transfer);
int  song_los_wait(song_los_ctx *ctx, song_los_timing *timing);

/* Page-locked host memory (cudaHostAlloc, portable across devices). */
int  song_los_host_alloc(void **ptr, size_t bytes);
int  song_los_host_free(void *ptr);

/* Upper bound on the k3 values of one batch such that three batches in flight and the intermediates of
 * one compute fit in the free device memory.  src_doubles_per_unit = (source doubles of the run) /
 * (k3 values of the run), i.e. tp2_size * tau_size * <k3_pt> / <k3_tr>.  Needs the projection tables. */
int  song_los_max_batch_units(song_los_ctx *ctx, int mode, double src_doubles_per_unit, int64_t *units);
int  song_los_stage_sources(song_los_ctx *ctx, const double *src_part, int64_t offset, int64_t n, int64_t n_src_total);
int  song_los_max_resident_units(song_los_ctx *ctx, int mode, double src_doubles_per_unit, int n_slots, int64_t *units);

/* Pre-allocates all device buffers for a pipeline of `n_in_flight` (1..SONG_LOS_MAX_IN_FLIGHT) batches of up to the
 * given sizes (pairs, k3 values, coarse k3 points, source doubles).  Optional, but without it a batch larger than every earlier one makes the
 * engine grow its buffers, and cudaMalloc/cudaFree stall the pipeline. */
int  song_los_reserve(song_los_ctx *ctx, int mode, int n_in_flight, int max_pairs, int64_t max_units, int64_t max_pt,
                      int64_t max_src);

/* Device pointer of the [tt2_size][n_units] result of the last compute (for the NCCL gather
   done by the multi-GPU driver), as an integer address; 0 before the first compute. */
uint64_t song_los_device_result(song_los_ctx *ctx, int64_t *n_doubles);

/* Copies the result of the last compute into caller-owned DEVICE memory (device-to-device, on the
   engine's stream, synchronised before returning): the hand-off to an NCCL gather. */
int  song_los_copy_result(song_los_ctx *ctx, uint64_t device_dst, int64_t n_doubles);

/* ---- the k3 integral of the intrinsic bispectrum on the device ----
 * bispectra2_intrinsic_integrate_over_k3 (source/bispectra2.c:1107-1439) integrates every T_l3,M3(k1,k2,k3) over k3
 * against j_L3(k3 r) with bessel2_convolution (source/bessel2.c:499-605).  song_los_integrate_k3 does that on the
 * transfer functions song_los_compute left in device memory (current slot), for the pairs of that batch:
 *     out[q][index_r][pair] = 2 * M3_coefficient[q] * 1/2 sum_k3 delta_k3 k3^2 T_tt[q](pair, k3) j_{l1[index_l1[q]]}(k3 r)
 * (the coefficient is applied only when abs_M3 != 0, bispectra2.c:1369; delta_k3 = trapezoidal measure of the pair's
 * own k3 grid, :1279-1284).  Bit-identical to the reference's statement order. */
typedef struct song_los_k3_request {
  int n_requests;                  /* (l3, M3) combinations */
  const int *tt;                   /* index_tt2_k3 + ptr2->lm_array[index_l3][index_M3]            [n_requests] */
  const int *index_l1;             /* pbs2->index_l1[L3], L3 = |l3 - |M3|| + offset_L3 (:1183,1200) [n_requests] */
  int abs_M3;                      /* pwb->abs_M3 */
  const double *M3_coefficient;    /* pwb->M3_coefficient[index_M3] per request; may be NULL when abs_M3 == 0 */
  int r_size; const double *r;     /* pwb->r */
} song_los_k3_request;

/* pbs2->j_l1 (and ddj_l1 for cubic Bessel interpolation), x_min_l1, x_size_l1 of bessel2_init (bessel2.h:96-106). */
int  song_los_set_bessel_table(song_los_ctx *ctx, int l1_size, const double *const *j_l1, const double *const *ddj_l1,
                               const double *x_min_l1, const int *x_size_l1);
int  song_los_integrate_k3(song_los_ctx *ctx, const song_los_k3_request *request, double *out /* [n_requests][r_size][n_pairs] */);
float song_los_last_k3_ms(const song_los_ctx *ctx);   /* device time of the last song_los_integrate_k3 kernel */

/* ---- the (k1,k2,k3) sum of the intrinsic C_l ----
 * spectra2_cls (source/spectra2.c:241-570) sums total_factor(k1,k2,k3,m) T^X_lm T^Y_lm over the nodes of the transfer
 * functions, with the trapezoidal steps of :377-381, :409-414, :471-492 (triangular edges included) and the primordial
 * spectra of :389-398, :421-430.  song_los_cl_sum returns the contribution of the pairs of the batch in the current slot;
 * the caller adds the batches (and GPUs).  The sum runs over k3 in the reference's order inside a pair and over the pairs
 * in the reference's order; it is not the reference's single running sum, so agreement is to rounding (1e-13), and for
 * m != 0 pow() is the device's. */
typedef struct song_los_cl_request {
  int n_requests;                  /* (index_ct, index_l, index_M) combinations */
  const int *tt1, *tt2;            /* index_tt_1, index_tt_2 (spectra2.c:336-353)                 [n_requests] */
  const int *m;                    /* ptr2->m[index_M] of each request                           [n_requests] */
  int rescale_cmb_sources;         /* ppt2->rescale_cmb_sources (:505) */
  int k_size; const double *k;     /* ppt2->k */
  const double *primordial;        /* primordial_spectrum_at_k (ppm, index_md, linear, k[i]) per k index [k_size] */
  const int *pair_k1, *pair_k2;    /* index_k1, index_k2 of every pair of the batch */
} song_los_cl_request;
int  song_los_cl_sum(song_los_ctx *ctx, const song_los_cl_request *request, double *out /* [n_requests] */);

/* Runs the engine's kernels on an external CUDA stream (a cudaStream_t passed as void*),
   so that a caller's own events see the work.  NULL restores the engine's stream. */
int  song_los_set_stream(song_los_ctx *ctx, void *cuda_stream);

/* Roofline denominators measured on the device the context lives on: dependency-free DFMA loop
   and DMMA (m8n8k4) loop -> FP64 TFLOP/s; device-to-device copy -> GB/s (read+write bytes). */
int  song_los_measure_peaks(song_los_ctx *ctx, double *dfma_tflops, double *dmma_tflops, double *copy_gbs);

/* Free bytes / total bytes / SM count of the context's device. */
int  song_los_device_info(song_los_ctx *ctx, int64_t *free_bytes, int64_t *total_bytes,
                          int *sm_count, char *name, int name_len);

int  song_los_version(void);

#ifdef __cplusplus
}
#endif
#endif
