/*
 * transfer2_gpu.c -- drop-in replacement of SONG's source/transfer2.c: same exported symbols and struct
 * semantics (reference include/transfer2.h:327-541), with the (k1,k2) loop body -- k3 interpolation of the
 * sources, time grid, and every transfer2_compute / transfer2_integrate call (reference
 * source/transfer2.c:401-539) -- running on B200 GPUs through the C-ABI of include/song_los.h.
 *
 * Compiled against SONG's OWN headers (-I<song>/include -I<song>/class.git/include); the stand-ins under
 * song_b200/host/compat/ are used only when the CLASS fork is absent, as in this repository's tests.
 *
 * What stays on the host, bit-identical to the reference: the index tables (transfer2.c:866-1285), the k3
 * integration grids (transfer2.c:1329-1737, through los_host.c, because bispectra2.c:1255 and spectra2.c:451
 * re-derive them), memory management of ptr2->transfer (transfer2.c:636-691, 2976-3104) and the disk format
 * (transfer2.c:705-763, 2929-2963: raw doubles, [k1][k2][k3] per transfer type).
 *
 * Batching: the reference handles one (k1,k2) pair at a time; here runs of consecutive pairs are packed into GPU
 * batches, because the tensor-core kernel wants many k3 values per table row.  The pairs are spread over the visible
 * GPUs by a static load-balanced schedule at pair granularity (song_host_schedule_pairs; transfer2.c:549-551: the k1
 * modes are independent); every GPU has one host thread that runs its batches as a three-deep pipeline
 * (song_los_submit / song_los_wait): inputs are packed into page-locked staging, the result of a batch is copied by the
 * device straight into the slab of ptr2->transfer that holds its pairs (los_module.h) -- that copy is the "gather to
 * the host".  Side effects keep the reference's per-k1 meaning: sources of a k1 level are loaded before use and freed
 * once its last pair is packed (transfer2.c:332-344, 553), transfers are stored and freed per k1 in k1 order
 * (transfer2.c:559-567).  The engines (CUDA contexts, projection tables) are those bessel2_init left on the devices
 * (bessel2_gpu.c); with the reference's bessel2.c they are built here from pbs2->J_Llm_x.
 *
 * Environment: SONG_LOS_MODE=fast|exact (default fast; cubic Bessel interpolation forces exact),
 * SONG_LOS_DEVICES=<n> (default: all visible), SONG_LOS_BATCH_UNITS, SONG_LOS_MIN_BATCHES, SONG_LOS_CHUNKS_PER_DEVICE,
 * SONG_LOS_HOST_THREADS, SONG_LOS_PINNED_TRANSFERS=0|1, SONG_LOS_PINNED_POOL=0|1.
 */
#ifndef _XOPEN_SOURCE
#define _XOPEN_SOURCE 700          /* pwrite, fileno (the writer thread of store_transfers_to_disk) */
#endif
#include "transfer2.h"
#include "song_los.h"
#include "los_host.h"
#include "los_glue.h"
#include "los_module.h"

#ifdef _OPENMP
#include <omp.h>
#endif
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------------ */
/*  small helpers                                                                                    */
/* ------------------------------------------------------------------------------------------------ */
static void k3_params_of(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2,
                         struct bessels2 *pbs2, struct transfers2 *ptr2, song_k3_params *kp)
{
  kp->tau0 = ptr2->tau0; kp->tau_rec = ptr2->tau_rec; kp->rs_rec = ptr2->rs_rec;
  kp->k_first = ppt2->k[0]; kp->k_last = ppt2->k[ppt2->k_size - 1];
  kp->k3_extrapolation = (ppr->bispectra_k3_extrapolation != no_k3_extrapolation);
  kp->extra_k3_oscillations_left = ppr->extra_k3_oscillations_left;
  kp->extra_k3_oscillations_right = ppr->extra_k3_oscillations_right;
  kp->k_sampling = (ptr2->k_sampling == class_transfer2_k3_sampling) ? 1 : 0;
  kp->q_linstep_song = ppr2->q_linstep_song;
  kp->old_run_step = (ppr->load_run == _TRUE_) && (ppr2->old_run == _TRUE_);
  kp->smallest_allowed_variation = ppr->smallest_allowed_variation;
  kp->xx_max = pbs2->xx_max; kp->xx_step = pbs2->xx_step;
}

static int env_int(const char *name, int fallback)
{
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : fallback;
}

/* ------------------------------------------------------------------------------------------------ */
/*  index tables (reference: transfer2.c:866-1285)                                                   */
/* ------------------------------------------------------------------------------------------------ */
int transfer2_get_l_list(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2,
                         struct bessels *pbs, struct bessels2 *pbs2, struct transfers2 *ptr2)
{
  /* the multipoles are those of the Bessel module (transfer2.c:1081-1103) */
  ptr2->l_size = pbs->l_size;
  class_alloc(ptr2->l, ptr2->l_size * sizeof(int), ptr2->error_message);
  for (int index_l = 0; index_l < ptr2->l_size; ++index_l) ptr2->l[index_l] = pbs->l[index_l];
  return _SUCCESS_;
}

int transfer2_get_lm_lists(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2,
                           struct bessels *pbs, struct bessels2 *pbs2, struct transfers2 *ptr2)
{
  /* (l,m) -> offset inside a field block (transfer2.c:1133-1166) */
  class_alloc(ptr2->lm_array, ptr2->l_size * sizeof(int *), ptr2->error_message);
  for (int index_l = 0; index_l < ptr2->l_size; ++index_l) {
    int l = ptr2->l[index_l];
    class_calloc(ptr2->lm_array[index_l], ppr2->index_m_max[l] + 1, sizeof(int), ptr2->error_message);
    for (int index_m = 0; index_m <= ppr2->index_m_max[l]; ++index_m)
      ptr2->lm_array[index_l][index_m] =
          multipole2offset_indexl_indexm(l, ptr2->m[index_m], ptr2->l, ptr2->l_size, ptr2->m, ptr2->m_size);
  }
  /* transfer type -> (field, index_l, index_m) and labels (transfer2.c:1180-1282) */
  class_alloc(ptr2->corresponding_index_l, ptr2->tt2_size * sizeof(int), ptr2->error_message);
  class_alloc(ptr2->corresponding_index_m, ptr2->tt2_size * sizeof(int), ptr2->error_message);
  class_alloc(ptr2->index_tt2_monopole, ptr2->tt2_size * sizeof(int), ptr2->error_message);
  class_alloc(ptr2->index_pt2_monopole, ptr2->tt2_size * sizeof(int), ptr2->error_message);
  for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt) {
    int index_l = -1, index_m = -1;
    const char *field = NULL;
    if (ppt2->has_source_T == _TRUE_ && index_tt >= ptr2->index_tt2_T && index_tt < ptr2->index_tt2_T + ptr2->n_transfers) {
      ptr2->index_pt2_monopole[index_tt] = ppt2->index_tp2_T; ptr2->index_tt2_monopole[index_tt] = ptr2->index_tt2_T; field = "T";
    }
    else if (ppt2->has_source_E == _TRUE_ && index_tt >= ptr2->index_tt2_E && index_tt < ptr2->index_tt2_E + ptr2->n_transfers) {
      ptr2->index_pt2_monopole[index_tt] = ppt2->index_tp2_E; ptr2->index_tt2_monopole[index_tt] = ptr2->index_tt2_E; field = "E";
    }
    else if (ppt2->has_source_B == _TRUE_ && index_tt >= ptr2->index_tt2_B && index_tt < ptr2->index_tt2_B + ptr2->n_transfers) {
      ptr2->index_pt2_monopole[index_tt] = ppt2->index_tp2_B; ptr2->index_tt2_monopole[index_tt] = ptr2->index_tt2_B; field = "B";
    }
    if (field != NULL) {
      offset2multipole_indexl_indexm(index_tt - ptr2->index_tt2_monopole[index_tt], ptr2->l, ptr2->l_size,
                                     ptr2->m, ptr2->m_size, &index_l, &index_m);
      if (index_l >= 0 && index_m >= 0 && index_l < ptr2->l_size && index_m < ptr2->m_size)
        sprintf(ptr2->tt2_labels[index_tt], "%s_%d_%d", field, ptr2->l[index_l], ptr2->m[index_m]);
    }
    class_test((index_l >= ptr2->l_size) || (index_m >= ptr2->m_size) || (index_l < 0) || (index_m < 0),
               ptr2->error_message,
               "index_tt=%d: result (index_l,index_m)=(%d,%d) is out of bounds index_l=[%d,%d], index_m=[%d,%d]\n",
               index_tt, index_l, index_m, 0, ptr2->l_size - 1, 0, ptr2->m_size - 1);
    ptr2->corresponding_index_l[index_tt] = index_l;
    ptr2->corresponding_index_m[index_tt] = index_m;
  }
  return _SUCCESS_;
}

/* fine k3 grid of one pair: size (transfer2.c:1329-1551) and values (transfer2.c:1566-1737) */
int transfer2_get_k3_size(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2,
                          struct bessels *pbs, struct bessels2 *pbs2, struct transfers2 *ptr2,
                          int index_k1, int index_k2)
{
  song_k3_params kp; song_k3_range r;
  k3_params_of(ppr, ppr2, ppt2, pbs2, ptr2, &kp);
  class_test(song_host_k3_size(&kp, ppt2->k3[index_k1][index_k2], ppt2->k3_size[index_k1][index_k2], &r,
                               ptr2->error_message, _ERRORMSGSIZE_) != 0,
             ptr2->error_message, "k3 grid of (index_k1,index_k2)=(%d,%d): %.1000s", index_k1, index_k2, ptr2->error_message);
  ptr2->k_min_k1k2[index_k1][index_k2] = r.k_min;
  ptr2->k_max_k1k2[index_k1][index_k2] = r.k_max;
  ptr2->k_size_k1k2[index_k1][index_k2] = r.k_size;
  ptr2->k_physical_start_k1k2[index_k1][index_k2] = r.k_physical_start;
  ptr2->k_physical_size_k1k2[index_k1][index_k2] = r.k_physical_size;
  return _SUCCESS_;
}

int transfer2_get_k3_list(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2,
                          struct bessels *pbs, struct bessels2 *pbs2, struct transfers2 *ptr2,
                          int index_k1, int index_k2, double *k3, int *last_used_index_pt)
{
  song_k3_params kp; song_k3_range r;
  k3_params_of(ppr, ppr2, ppt2, pbs2, ptr2, &kp);
  r.k_min = ptr2->k_min_k1k2[index_k1][index_k2];
  r.k_max = ptr2->k_max_k1k2[index_k1][index_k2];
  r.k_size = ptr2->k_size_k1k2[index_k1][index_k2];
  r.k_physical_start = ptr2->k_physical_start_k1k2[index_k1][index_k2];
  r.k_physical_size = ptr2->k_physical_size_k1k2[index_k1][index_k2];
  int last = 0;
  class_test(song_host_k3_list(&kp, ppt2->k3[index_k1][index_k2], ppt2->k3_size[index_k1][index_k2], &r, k3, &last,
                               ptr2->error_message, _ERRORMSGSIZE_) != 0,
             ptr2->error_message, "k3 list of (index_k1,index_k2)=(%d,%d): %.1000s", index_k1, index_k2, ptr2->error_message);
  if (last_used_index_pt != NULL) *last_used_index_pt = last;
  return _SUCCESS_;
}

int transfer2_get_k3_sizes(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2,
                           struct bessels *pbs, struct bessels2 *pbs2, struct transfers *ptr, struct transfers2 *ptr2)
{
  /* transfer2.c:1752-1834 */
  int k1_size = ppt2->k_size;
  class_alloc(ptr2->k_size_k1k2, k1_size * sizeof(int *), ptr2->error_message);
  class_alloc(ptr2->k_physical_start_k1k2, k1_size * sizeof(int *), ptr2->error_message);
  class_alloc(ptr2->k_physical_size_k1k2, k1_size * sizeof(int *), ptr2->error_message);
  class_alloc(ptr2->k_min_k1k2, k1_size * sizeof(double *), ptr2->error_message);
  class_alloc(ptr2->k_max_k1k2, k1_size * sizeof(double *), ptr2->error_message);
  ptr2->k3_size_max = 0;
  for (int index_k1 = 0; index_k1 < k1_size; ++index_k1) {
    int k2_size = index_k1 + 1;
    class_alloc(ptr2->k_size_k1k2[index_k1], k2_size * sizeof(int), ptr2->error_message);
    class_alloc(ptr2->k_physical_start_k1k2[index_k1], k2_size * sizeof(int), ptr2->error_message);
    class_alloc(ptr2->k_physical_size_k1k2[index_k1], k2_size * sizeof(int), ptr2->error_message);
    class_alloc(ptr2->k_min_k1k2[index_k1], k2_size * sizeof(double), ptr2->error_message);
    class_alloc(ptr2->k_max_k1k2[index_k1], k2_size * sizeof(double), ptr2->error_message);
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) {
      /* cubic interpolation needs at least 4 source points (transfer2.c:1766-1771) */
      class_test((ppr2->sources_k3_interpolation == cubic_interpolation) && (ppt2->k3_size[index_k1][index_k2] < 4),
                 ptr2->error_message,
                 "index_k1=%d, index_k2=%d: cannot do cubic interpolation in k3 with just k3_size=%d values. Increase k3_size.",
                 index_k1, index_k2, ppt2->k3_size[index_k1][index_k2]);
      class_call(transfer2_get_k3_size(ppr, ppr2, ppt2, pbs, pbs2, ptr2, index_k1, index_k2),
                 ptr2->error_message, ptr2->error_message);
      ptr2->k3_size_max = MAX(ptr2->k3_size_max, ptr2->k_size_k1k2[index_k1][index_k2]);
    }
  }
  return _SUCCESS_;
}

int transfer2_indices_of_transfers(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2,
                                   struct bessels *pbs, struct bessels2 *pbs2, struct transfers *ptr, struct transfers2 *ptr2)
{
  class_call(transfer2_get_l_list(ppr, ppr2, ppt2, pbs, pbs2, ptr2), ptr2->error_message, ptr2->error_message);
  ptr2->m_size = ppr2->m_size;
  class_alloc(ptr2->m, ptr2->m_size * sizeof(int), ptr2->error_message);
  for (int index_m = 0; index_m < ptr2->m_size; ++index_m) ptr2->m[index_m] = ppr2->m[index_m];

  ptr2->n_transfers = size_indexl_indexm(ptr2->l, ptr2->l_size, ppr2->m, ppr2->m_size);
  ptr2->n_nonzero_transfers_E = ptr2->n_transfers;                     /* transfer2.c:903-915 (debug only) */
  if (ppr2->compute_m[0]) ptr2->n_nonzero_transfers_E -= 2;
  if (ppr2->compute_m[1]) ptr2->n_nonzero_transfers_E -= 1;
  ptr2->n_nonzero_transfers_B = ptr2->n_transfers;
  if (ppr2->compute_m[0]) ptr2->n_nonzero_transfers_B -= ptr2->l_size;
  if (ppr2->compute_m[1]) ptr2->n_nonzero_transfers_B -= 1;

  int index_tt = 0;                                                   /* T block, then E, then B (:917-940) */
  if (ppt2->has_source_T) { ptr2->index_tt2_T = index_tt; index_tt += ptr2->n_transfers; }
  if (ppt2->has_source_E) { ptr2->index_tt2_E = index_tt; index_tt += ptr2->n_transfers; }
  if (ppt2->has_source_B) { ptr2->index_tt2_B = index_tt; index_tt += ptr2->n_transfers; }
  ptr2->tt2_size = index_tt;
  class_calloc(ptr2->tt2_labels, ptr2->tt2_size * _MAX_LENGTH_LABEL_, sizeof(char), ptr2->error_message);
  if (ptr2->transfer2_verbose > 1)
    printf(" -> will compute tt2_size=%d transfer functions (%d per field)\n", ptr2->tt2_size, ptr2->n_transfers);

  class_call(transfer2_get_lm_lists(ppr, ppr2, ppt2, pbs, pbs2, ptr2), ptr2->error_message, ptr2->error_message);
  class_call(transfer2_get_k3_sizes(ppr, ppr2, ppt2, pbs, pbs2, ptr, ptr2), ptr2->error_message, ptr2->error_message);

  ptr2->count_memorised_transfers = 0;
  ptr2->count_allocated_transfers = 0;
  class_alloc(ptr2->transfer, ptr2->tt2_size * sizeof(double ***), ptr2->error_message);
  for (int itt = 0; itt < ptr2->tt2_size; ++itt)
    class_alloc(ptr2->transfer[itt], ppt2->k_size * sizeof(double **), ptr2->error_message);
  class_calloc(ptr2->has_allocated_transfers, ppt2->k_size, sizeof(short), ptr2->error_message);

  if ((ppr2->store_transfers_to_disk == _TRUE_) || (ppr2->load_transfers_from_disk == _TRUE_)) {   /* :1043-1062 */
    class_alloc(ptr2->transfers_files, ptr2->tt2_size * sizeof(FILE *), ptr2->error_message);
    class_alloc(ptr2->transfers_paths, ptr2->tt2_size * sizeof(char *), ptr2->error_message);
    for (int itt = 0; itt < ptr2->tt2_size; ++itt) {
      class_alloc(ptr2->transfers_paths[itt], _FILENAMESIZE_ * sizeof(char), ptr2->error_message);
      snprintf(ptr2->transfers_paths[itt], _FILENAMESIZE_, "%s/transfers_%03d.dat", ptr2->transfers_dir, itt);
      if (ppr2->store_transfers_to_disk == _TRUE_)
        class_open(ptr2->transfers_files[itt], ptr2->transfers_paths[itt], "wb", ptr2->error_message);
    }
  }
  return _SUCCESS_;
}

/* ------------------------------------------------------------------------------------------------ */
/*  memory management and disk format of ptr2->transfer                                              */
/* ------------------------------------------------------------------------------------------------ */
/* The reference allocates ptr2->transfer[tt][k1][k2] one array at a time (tt2_size * (k1+1) mallocs per k1 level,
   transfer2.c:2976-3032).  Here the pointers of all levels live in one block and the values in large page-locked slabs,
   [tt2_size][k3 values of a run of pairs] each (los_module.h): a GPU batch copies its result straight into its slab.
   The exported functions keep the reference's meaning and guards. */
static los_transfer_store *store_of(struct perturbs2 *ppt2, struct transfers2 *ptr2)
{
  los_transfer_store *s = los_store_find(ptr2);
  if (s == NULL) s = los_store_create(ptr2, ppt2->k_size, ptr2->tt2_size);
  return s;
}

static int pinned_transfers(void) { return env_int("SONG_LOS_PINNED_TRANSFERS", 1) != 0; }

/* Values of the pairs (k1a,k2a) ... (k1b,k2b), reference loop order, as ONE slab [tt2_size][units of the run]; points
   ptr2->transfer[tt][k1][k2] at its rows.  *base receives the slab, *n_units the k3 values of the run. */
static int allocate_pair_run(struct perturbs2 *ppt2, struct transfers2 *ptr2, int k1a, int k2a, int k1b, int k2b,
                             double **base, int64_t *n_units)
{
  los_transfer_store *st = store_of(ppt2, ptr2);
  class_test(st == NULL, ptr2->error_message, "could not set up the storage of ptr2->transfer");
  int64_t units = 0;
  for (int k1 = k1a; k1 <= k1b; ++k1)
    for (int k2 = (k1 == k1a ? k2a : 0); k2 <= (k1 == k1b ? k2b : k1); ++k2) units += ptr2->k_size_k1k2[k1][k2];
  double *slab = NULL;
  int id = los_store_new_slab(st, (size_t)units * ptr2->tt2_size, pinned_transfers(), &slab);
  class_test(id < 0, ptr2->error_message, "could not allocate %g MB for the transfer functions of k1 levels %d-%d",
             (double)units * ptr2->tt2_size * 8e-6, k1a, k1b);
  for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt) {
    double *row = slab + (size_t)index_tt * (size_t)units;
    for (int k1 = k1a; k1 <= k1b; ++k1) {
      double **ptrs = los_store_row(st, index_tt, k1);
      ptr2->transfer[index_tt][k1] = ptrs;
      for (int k2 = (k1 == k1a ? k2a : 0); k2 <= (k1 == k1b ? k2b : k1); ++k2) { ptrs[k2] = row; row += ptr2->k_size_k1k2[k1][k2]; }
    }
  }
  for (int k1 = k1a; k1 <= k1b; ++k1) { los_store_link(st, k1, id); ptr2->has_allocated_transfers[k1] = _TRUE_; }
  const long int count = (long int)units * ptr2->tt2_size;
  #pragma omp atomic
  ptr2->count_allocated_transfers += count;
  if (base) *base = slab;
  if (n_units) *n_units = units;
  return _SUCCESS_;
}

int transfer2_allocate_k1_level(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_k1)
{
  class_test(ptr2->has_allocated_transfers[index_k1] == _TRUE_, ptr2->error_message,
             "the index_k1=%d level of ptr2->transfers is already allocated, stop to prevent error", index_k1);
  return allocate_pair_run(ppt2, ptr2, index_k1, 0, index_k1, index_k1, NULL, NULL);
}

int transfer2_free_k1_level(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_k1)
{
  class_test(ptr2->has_allocated_transfers[index_k1] == _FALSE_, ptr2->error_message,
             "the index_k1=%d level of ptr2->transfers is already free, stop to prevent error", index_k1);
  los_transfer_store *st = los_store_find(ptr2);
  if (st != NULL) los_store_release_level(st, index_k1);      /* a slab goes back when its last level is freed */
  ptr2->has_allocated_transfers[index_k1] = _FALSE_;
  return _SUCCESS_;
}

/* One transfer type for all (k1,k2,k3) -- the layout of a transfers_%03d.dat file (transfer2.c:636-691): one array of
   values and one of pointers per type; transfer2_free_type_level recovers both from [0][0] and [0]. */
int transfer2_allocate_type_level(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_tt)
{
  const int k1_size = ppt2->k_size;
  size_t n_pairs = (size_t)k1_size * (k1_size + 1) / 2, per_type = 0;
  for (int index_k1 = 0; index_k1 < k1_size; ++index_k1)
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) per_type += ptr2->k_size_k1k2[index_k1][index_k2];
  double *values, **ptrs;
  class_alloc(ptr2->transfer[index_tt], k1_size * sizeof(double **), ptr2->error_message);
  class_alloc(ptrs, n_pairs * sizeof(double *), ptr2->error_message);
  class_alloc(values, (per_type > 0 ? per_type : 1) * sizeof(double), ptr2->error_message);
  for (int index_k1 = 0; index_k1 < k1_size; ++index_k1) {
    ptr2->transfer[index_tt][index_k1] = ptrs + (size_t)index_k1 * (index_k1 + 1) / 2;
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) {
      ptr2->transfer[index_tt][index_k1][index_k2] = values;
      values += ptr2->k_size_k1k2[index_k1][index_k2];
    }
  }
  #pragma omp atomic
  ptr2->count_allocated_transfers += (long int)per_type;
  return _SUCCESS_;
}

int transfer2_free_type_level(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_tt)
{
  free(ptr2->transfer[index_tt][0][0]);
  free(ptr2->transfer[index_tt][0]);
  free(ptr2->transfer[index_tt]);
  return _SUCCESS_;
}

int transfer2_store_transfers_to_disk(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_k1)
{
  /* one file per transfer type, appended per k1 in [k2][k3] order (transfer2.c:2940-2957) */
  for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt)
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) {
      size_t n = (size_t)ptr2->k_size_k1k2[index_k1][index_k2];
      class_test(fwrite(ptr2->transfer[index_tt][index_k1][index_k2], sizeof(double), n, ptr2->transfers_files[index_tt]) != n,
                 ptr2->error_message, "could not write to '%s'", ptr2->transfers_paths[index_tt]);
    }
  return _SUCCESS_;
}

int transfer2_load_transfers_from_disk(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_tt)
{
  class_call(transfer2_allocate_type_level(ppt2, ptr2, index_tt), ptr2->error_message, ptr2->error_message);
  class_open(ptr2->transfers_files[index_tt], ptr2->transfers_paths[index_tt], "rb", ptr2->error_message);
  /* the file holds [k1][k2][k3] (transfer2.c:725-751), which is the order of the type's array: one read */
  size_t n_to_read = 0;
  for (int index_k1 = 0; index_k1 < ppt2->k_size; ++index_k1)
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) n_to_read += ptr2->k_size_k1k2[index_k1][index_k2];
  size_t n_read = fread(ptr2->transfer[index_tt][0][0], sizeof(double), n_to_read, ptr2->transfers_files[index_tt]);
  fclose(ptr2->transfers_files[index_tt]);
  class_test(n_read != n_to_read, ptr2->error_message,
             "Could not read in '%s' file, read %zu entries but expected %zu (index_tt=%d)",
             ptr2->transfers_paths[index_tt], n_read, n_to_read, index_tt);
  #pragma omp atomic
  ptr2->count_memorised_transfers += (long int)n_to_read;
  return _SUCCESS_;
}

int transfer2_free(struct precision2 *ppr2, struct perturbs2 *ppt2, struct transfers2 *ptr2)
{
  if (ptr2->has_cls == _TRUE_) {
    free(ptr2->tt2_labels);
    if ((ppr2->store_transfers_to_disk == _FALSE_) && (ppr2->load_transfers_from_disk == _FALSE_)) {
      for (int index_k1 = 0; index_k1 < ppt2->k_size; ++index_k1)
        if (ptr2->has_allocated_transfers[index_k1] == _TRUE_)
          class_call(transfer2_free_k1_level(ppt2, ptr2, index_k1), ptr2->error_message, ptr2->error_message);
      for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt) free(ptr2->transfer[index_tt]);
    }
    los_store_drop(ptr2);
    free(ptr2->transfer);
    free(ptr2->has_allocated_transfers);
    for (int index_k1 = 0; index_k1 < ppt2->k_size; ++index_k1) {
      free(ptr2->k_size_k1k2[index_k1]); free(ptr2->k_physical_start_k1k2[index_k1]);
      free(ptr2->k_physical_size_k1k2[index_k1]); free(ptr2->k_min_k1k2[index_k1]); free(ptr2->k_max_k1k2[index_k1]);
    }
    free(ptr2->k_size_k1k2); free(ptr2->k_physical_start_k1k2); free(ptr2->k_physical_size_k1k2);
    free(ptr2->k_min_k1k2); free(ptr2->k_max_k1k2);
    free(ptr2->l); free(ptr2->m);
    free(ptr2->corresponding_index_l); free(ptr2->corresponding_index_m);
    free(ptr2->index_tt2_monopole); free(ptr2->index_pt2_monopole);
    for (int index_l = 0; index_l < ptr2->l_size; ++index_l) free(ptr2->lm_array[index_l]);
    free(ptr2->lm_array);
    if ((ppr2->store_transfers_to_disk == _TRUE_) || (ppr2->load_transfers_from_disk == _TRUE_)) {
      for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt) free(ptr2->transfers_paths[index_tt]);
      free(ptr2->transfers_files); free(ptr2->transfers_paths);
    }
  }
  return _SUCCESS_;
}

/* ------------------------------------------------------------------------------------------------ */
/*  the per-element functions of the reference run inside the GPU kernels                            */
/* ------------------------------------------------------------------------------------------------ */
#define ON_DEVICE(name) class_stop(ptr2->error_message, \
  "%s is evaluated inside the CUDA kernels of this build (include/song_los.h); it is not available as a host call", name)

int transfer2_interpolate_sources_in_k(struct precision *ppr, struct precision2 *ppr2, struct perturbs *ppt,
      struct perturbs2 *ppt2, struct bessels *pbs, struct bessels2 *pbs2, struct transfers2 *ptr2, int index_k1,
      int index_k2, int index_tp, double *k_grid, double *splines, double *interpolated_sources)
{ ON_DEVICE("transfer2_interpolate_sources_in_k"); }
int transfer2_get_time_grid(struct precision *ppr, struct precision2 *ppr2, struct perturbs *ppt, struct perturbs2 *ppt2,
      struct bessels *pbs, struct bessels2 *pbs2, struct transfers2 *ptr2, int index_k1, int index_k2, int index_k,
      struct transfer2_workspace *pw)
{ ON_DEVICE("transfer2_get_time_grid"); }
int transfer2_interpolate_sources_in_time(struct precision *ppr, struct precision2 *ppr2, struct perturbs *ppt,
      struct perturbs2 *ppt2, struct bessels *pbs, struct bessels2 *pbs2, struct transfers2 *ptr2, int index_tp2,
      double *interpolated_sources_in_k, double *sources_time_spline, double *interpolated_sources_in_time,
      struct transfer2_workspace *pw)
{ ON_DEVICE("transfer2_interpolate_sources_in_time"); }
int transfer2_compute(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels *pbs,
      struct bessels2 *pbs2, struct transfers2 *ptr2, int index_k1, int index_k2, int index_k, int index_l, int index_m,
      int index_tt, double **interpolated_sources_in_time, struct transfer2_workspace *pw)
{ ON_DEVICE("transfer2_compute"); }
int transfer2_integrate(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels *pbs,
      struct bessels2 *pbs2, struct transfers2 *ptr2, int index_k1, int index_k2, int index_k, int index_l, int index_m,
      double **interpolated_sources_in_time, int index_J, int index_source_monopole, struct transfer2_workspace *pw,
      double *integral)
{ ON_DEVICE("transfer2_integrate"); }

/* ------------------------------------------------------------------------------------------------ */
/*  GPU batches                                                                                      */
/* ------------------------------------------------------------------------------------------------ */
/* One GPU batch = a run of consecutive (k1,k2) pairs in the reference loop order (pair p = k1(k1+1)/2 + k2). */
typedef struct {
  int p_first, p_last;             /* pairs of the batch */
  int64_t n_units, n_pt, n_src;    /* k3 values, coarse k3 points, source doubles */
  int device;
  double *slab;                    /* [tt2_size][n_units]: where ptr2->transfer of these pairs lives */
} los_pair_batch;

/* Page-locked staging of the INPUTS of one batch in flight (the result needs none: it is copied into its slab).
   SONG_LOS_MAX_IN_FLIGHT of these per GPU, used in turn. */
typedef struct {
  int *np, *nt, *ps, *pz; int64_t *pt_off, *tr_off, *src_off;
  double *k3_pt, *k3_tr, *src;
  unsigned char *block; size_t block_bytes;
  int n_pairs;
  const los_pair_batch *pb;        /* NULL when the staging is free */
} los_staging;

static void staging_free(los_staging *s) { los_pinned_put(s->block); memset(s, 0, sizeof *s); }

/* one page-locked block for the largest batch of a device, carved into the arrays of song_los_batch */
static int staging_alloc(los_staging *s, int max_pairs, int64_t max_pt, int64_t max_units, int64_t max_src)
{
  memset(s, 0, sizeof *s);
  const size_t a_int = ((size_t)max_pairs * sizeof(int) + 63) & ~(size_t)63, a_i64 = ((size_t)max_pairs * sizeof(int64_t) + 63) & ~(size_t)63;
  const size_t a_pt = ((size_t)max_pt * 8 + 63) & ~(size_t)63, a_tr = ((size_t)max_units * 8 + 63) & ~(size_t)63, a_src = (size_t)max_src * 8 + 64;
  s->block_bytes = 4 * a_int + 3 * a_i64 + a_pt + a_tr + a_src;
  s->block = (unsigned char *)los_pinned_get(s->block_bytes);
  if (s->block == NULL) return _FAILURE_;
  unsigned char *p = s->block;
  s->np = (int *)p; p += a_int; s->nt = (int *)p; p += a_int; s->ps = (int *)p; p += a_int; s->pz = (int *)p; p += a_int;
  s->pt_off = (int64_t *)p; p += a_i64; s->tr_off = (int64_t *)p; p += a_i64; s->src_off = (int64_t *)p; p += a_i64;
  s->k3_pt = (double *)p; p += a_pt; s->k3_tr = (double *)p; p += a_tr; s->src = (double *)p;
  return _SUCCESS_;
}

/* what a run of the k1 loop shares between the host threads of the GPUs */
typedef struct {
  struct precision *ppr; struct precision2 *ppr2; struct perturbs2 *ppt2; struct bessels *pbs; struct bessels2 *pbs2;
  struct transfers2 *ptr2;
  int n_pairs; int *pair_k1, *pair_k2;     /* pair -> (index_k1, index_k2) */
  int *pairs_to_pack;                      /* per k1 level: pairs whose sources have not been packed yet (-> perturb2_free_k1_level) */
  int mode, host_threads, sources_on_disk, store_to_disk;
  int next_level_to_load;                        /* ordered runs (sources read from disk level by level) */
  double t_start;                                /* transfer2_verbose > 2: lap times of every batch */
  struct los_writer *writer;                     /* store_transfers_to_disk: the thread that writes finished batches */
} los_run;

static double los_now(void)
{
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.;
#endif
}

/* ---- store_transfers_to_disk (transfer2.c:559-567, 2929-2963) without stalling the GPUs ----
   transfers_%03d.dat holds, per transfer type, the values in [k1][k2][k3] order, i.e. in pair order; the result of a
   batch is, per type, one contiguous run of that file.  A writer thread therefore puts every finished batch at its offset
   with one pwrite per type, in whatever order the GPUs finish, while the host threads of the GPUs go on packing and
   submitting; a k1 level is freed as the reference does (:565) once every pair of it is on disk. */
typedef struct los_write_job { const los_pair_batch *pb; struct los_write_job *next; } los_write_job;
typedef struct los_writer {
  pthread_t thread; pthread_mutex_t lock; pthread_cond_t wake;
  los_write_job *head, *tail; int closing, failed;
  los_run *R; int64_t *unit_offset_of_pair;   /* position of every pair in a type's file, in doubles */
  int *pairs_to_write;                        /* per k1 level */
  ErrorMsg error;
} los_writer;

static int writer_store_batch(los_writer *W, const los_pair_batch *pb)
{
  los_run *R = W->R; struct transfers2 *ptr2 = R->ptr2;
  const off_t at = (off_t)W->unit_offset_of_pair[pb->p_first] * (off_t)sizeof(double);
  const size_t bytes = (size_t)pb->n_units * sizeof(double);
  for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt) {
    const char *src = (const char *)(pb->slab + (size_t)index_tt * (size_t)pb->n_units);
    const int fd = fileno(ptr2->transfers_files[index_tt]);
    size_t done = 0;
    while (done < bytes) {
      ssize_t w = pwrite(fd, src + done, bytes - done, at + (off_t)done);
      if (w <= 0) { snprintf(W->error, _ERRORMSGSIZE_, "could not write to '%s'", ptr2->transfers_paths[index_tt]); return _FAILURE_; }
      done += (size_t)w;
    }
  }
  /* levels whose last pair is now on disk */
  const int k1a = R->pair_k1[pb->p_first], k1b = R->pair_k1[pb->p_last];
  for (int k1 = k1a; k1 <= k1b; ++k1) {
    const int lo = (k1 == k1a) ? R->pair_k2[pb->p_first] : 0, hi = (k1 == k1b) ? R->pair_k2[pb->p_last] : k1;
    W->pairs_to_write[k1] -= hi - lo + 1;                      /* only this thread touches the counters */
    if (W->pairs_to_write[k1] == 0 && transfer2_free_k1_level(R->ppt2, ptr2, k1) == _FAILURE_) {
      snprintf(W->error, _ERRORMSGSIZE_, "%.1500s", ptr2->error_message); return _FAILURE_;
    }
  }
  return _SUCCESS_;
}

static void *writer_main(void *arg)
{
  los_writer *W = (los_writer *)arg;
  for (;;) {
    pthread_mutex_lock(&W->lock);
    while (W->head == NULL && !W->closing) pthread_cond_wait(&W->wake, &W->lock);
    los_write_job *job = W->head;
    if (job != NULL) { W->head = job->next; if (W->head == NULL) W->tail = NULL; }
    pthread_mutex_unlock(&W->lock);
    if (job == NULL) return NULL;                               /* closing and drained */
    if (!W->failed && writer_store_batch(W, job->pb) == _FAILURE_) W->failed = 1;
    free(job);
  }
}

static los_writer *writer_start(los_run *R)
{
  los_writer *W = (los_writer *)calloc(1, sizeof *W);
  if (W == NULL) return NULL;
  W->R = R;
  W->unit_offset_of_pair = (int64_t *)malloc(sizeof(int64_t) * (size_t)R->n_pairs);
  W->pairs_to_write = (int *)malloc(sizeof(int) * (size_t)R->ppt2->k_size);
  if (!W->unit_offset_of_pair || !W->pairs_to_write) { free(W->unit_offset_of_pair); free(W->pairs_to_write); free(W); return NULL; }
  int64_t off = 0;
  for (int p = 0; p < R->n_pairs; ++p) { W->unit_offset_of_pair[p] = off; off += R->ptr2->k_size_k1k2[R->pair_k1[p]][R->pair_k2[p]]; }
  for (int k1 = 0; k1 < R->ppt2->k_size; ++k1) W->pairs_to_write[k1] = k1 + 1;
  pthread_mutex_init(&W->lock, NULL); pthread_cond_init(&W->wake, NULL);
  if (pthread_create(&W->thread, NULL, writer_main, W) != 0) { free(W->unit_offset_of_pair); free(W->pairs_to_write); free(W); return NULL; }
  return W;
}

static int writer_push(los_writer *W, const los_pair_batch *pb)
{
  los_write_job *job = (los_write_job *)malloc(sizeof *job);
  if (job == NULL) return _FAILURE_;
  job->pb = pb; job->next = NULL;
  pthread_mutex_lock(&W->lock);
  if (W->tail) W->tail->next = job; else W->head = job;
  W->tail = job;
  pthread_cond_signal(&W->wake);
  pthread_mutex_unlock(&W->lock);
  return _SUCCESS_;
}

/* waits for the queue to drain; returns _FAILURE_ (message in errmsg) if a write failed */
static int writer_finish(los_writer *W, char *errmsg)
{
  pthread_mutex_lock(&W->lock);
  W->closing = 1;
  pthread_cond_signal(&W->wake);
  pthread_mutex_unlock(&W->lock);
  pthread_join(W->thread, NULL);
  const int failed = W->failed;
  if (failed) snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", W->error);
  pthread_mutex_destroy(&W->lock); pthread_cond_destroy(&W->wake);
  free(W->unit_offset_of_pair); free(W->pairs_to_write); free(W);
  return failed ? _FAILURE_ : _SUCCESS_;
}

/* packs the pairs of `pb` into the staging and enqueues the batch on the GPU (song_los_submit): the kernels' result goes
   straight to the batch's slab of ptr2->transfer */
static int submit_batch(los_run *R, song_los_ctx *ctx, los_pair_batch *pb, los_staging *s, char *errmsg)
{
  struct perturbs2 *ppt2 = R->ppt2; struct transfers2 *ptr2 = R->ptr2;
  const int tp2 = ppt2->tp2_size, ntau = ppt2->tau_size;
  const int k1a = R->pair_k1[pb->p_first], k1b = R->pair_k1[pb->p_last];
  const double t0 = los_now();

  /* sources kept on disk arrive one k1 level at a time (transfer2.c:332-344) */
  if (R->sources_on_disk)
    for (; R->next_level_to_load <= k1b; ++R->next_level_to_load)
      if (perturb2_allocate_k1_level(ppt2, R->next_level_to_load) == _FAILURE_ || perturb2_load_sources_from_disk(ppt2, R->next_level_to_load) == _FAILURE_) {
        snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", ppt2->error_message); return _FAILURE_;
      }

  /* ptr2->transfer of these pairs: one slab, which is also where the device->host copy of the batch lands */
  int64_t units = 0;
  if (allocate_pair_run(ppt2, ptr2, k1a, R->pair_k2[pb->p_first], k1b, R->pair_k2[pb->p_last], &pb->slab, &units) == _FAILURE_) {
    snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", ptr2->error_message); return _FAILURE_;
  }

  const double t1 = los_now();
  /* sizes and offsets first (serial), then every pair is packed independently by the host threads of this GPU */
  s->n_pairs = pb->p_last - pb->p_first + 1;
  int64_t o_pt = 0, o_tr = 0, o_src = 0;
  for (int q = 0; q < s->n_pairs; ++q) {
    const int k1 = R->pair_k1[pb->p_first + q], k2 = R->pair_k2[pb->p_first + q];
    s->np[q] = ppt2->k3_size[k1][k2]; s->nt[q] = ptr2->k_size_k1k2[k1][k2];
    s->ps[q] = ptr2->k_physical_start_k1k2[k1][k2]; s->pz[q] = ptr2->k_physical_size_k1k2[k1][k2];
    s->pt_off[q] = o_pt; s->tr_off[q] = o_tr; s->src_off[q] = o_src;
    o_pt += s->np[q]; o_tr += s->nt[q]; o_src += (int64_t)tp2 * ntau * s->np[q];
  }
  /* Packed in pieces of ~256 MB of sources: a finished piece starts its host->device copy at once
     (song_los_stage_sources), under the packing of the next ones; the last piece travels with song_los_submit. */
  int failed = 0;
  const int64_t piece_doubles = (int64_t)env_int("SONG_LOS_STAGE_MB", 256) * 131072;
  for (int q0 = 0; q0 < s->n_pairs && !failed;) {
    int q1 = q0 + 1;
    while (q1 < s->n_pairs && (piece_doubles <= 0 || s->src_off[q1] - s->src_off[q0] < piece_doubles)) ++q1;
    #pragma omp parallel for schedule(dynamic, 8) num_threads(R->host_threads) if (R->host_threads > 1)
    for (int q = q0; q < q1; ++q) {
      const int k1 = R->pair_k1[pb->p_first + q], k2 = R->pair_k2[pb->p_first + q];
      memcpy(s->k3_pt + s->pt_off[q], ppt2->k3[k1][k2], sizeof(double) * s->np[q]);
      if (transfer2_get_k3_list(R->ppr, R->ppr2, ppt2, R->pbs, R->pbs2, ptr2, k1, k2, s->k3_tr + s->tr_off[q], NULL) == _FAILURE_) {
        #pragma omp atomic write
        failed = 1;
        continue;
      }
      /* ppt2->sources[tp][k1][k2][tau*k3_size + k3] (transfer2.h:47-51) -> packed [pair][tp][tau][k3] */
      for (int tp = 0; tp < tp2; ++tp)
        memcpy(s->src + s->src_off[q] + (int64_t)tp * ntau * s->np[q], ppt2->sources[tp][k1][k2], sizeof(double) * (size_t)ntau * s->np[q]);
    }
    if (!failed && q1 < s->n_pairs && piece_doubles > 0 &&
        song_los_stage_sources(ctx, s->src + s->src_off[q0], s->src_off[q0], s->src_off[q1] - s->src_off[q0], o_src) != SONG_LOS_SUCCESS) {
      snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(ctx)); return _FAILURE_;
    }
    q0 = q1;
  }
  if (failed) { snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", ptr2->error_message); return _FAILURE_; }

  const double t2 = los_now();
  song_los_batch b;
  memset(&b, 0, sizeof b);
  b.n_pairs = s->n_pairs; b.k3_pt_size = s->np; b.k3_tr_size = s->nt; b.phys_start = s->ps; b.phys_size = s->pz;
  b.pt_off = s->pt_off; b.tr_off = s->tr_off; b.src_off = s->src_off;
  b.n_pt_total = o_pt; b.n_units = o_tr; b.n_src_total = o_src;
  b.k3_pt = s->k3_pt; b.k3_tr = s->k3_tr; b.src = s->src;
  if (song_los_submit(ctx, &b, R->mode, pb->slab) != SONG_LOS_SUCCESS) {
    snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(ctx)); return _FAILURE_;
  }
  s->pb = pb;
  if (ptr2->transfer2_verbose > 2)
    printf("       GPU %d batch of %ld k3 values: allocate %.2f ms, pack %.2f ms (%.1f GB/s), submit %.2f ms; submitted at %.1f ms\n",
           pb->device, (long)pb->n_units, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (double)o_src * 8e-9 / (t2 - t1 > 0 ? t2 - t1 : 1), (los_now() - t2) * 1e3,
           (los_now() - R->t_start) * 1e3);

  /* the sources of these pairs now live in the staging: a k1 level whose last pair has been packed is freed as the
     reference does after its k2 loop (transfer2.c:553) */
  for (int k1 = k1a; k1 <= k1b; ++k1) {
    const int lo = (k1 == k1a) ? R->pair_k2[pb->p_first] : 0, hi = (k1 == k1b) ? R->pair_k2[pb->p_last] : k1;
    int left;
    #pragma omp atomic capture
    left = R->pairs_to_pack[k1] -= (hi - lo + 1);
    if (left == 0) {
      int rc;
      #pragma omp critical(los_free_sources)
      rc = perturb2_free_k1_level(ppt2, k1);
      if (rc == _FAILURE_) { snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", ppt2->error_message); return _FAILURE_; }
    }
  }
  return _SUCCESS_;
}

/* waits for the oldest batch in flight -- its result is then in ptr2->transfer (transfer2.c:2107) -- and applies the
   per-k1 side effects of transfer2.c:559-567 */
static int finish_batch(los_run *R, song_los_ctx *ctx, los_staging *s, char *errmsg)
{
  struct transfers2 *ptr2 = R->ptr2;
  const los_pair_batch *pb = s->pb;
  s->pb = NULL;
  const double t0 = los_now();
  song_los_timing tm;
  if (song_los_wait(ctx, &tm) != SONG_LOS_SUCCESS) {
    snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(ctx)); return _FAILURE_;
  }
  if (ptr2->transfer2_verbose > 2)
    printf("       GPU %d batch of %ld k3 values done at %.1f ms (waited %.2f ms): h2d %.2f ms, kernels %.2f ms (LOS %.2f), d2h %.2f ms\n",
           pb->device, (long)pb->n_units, (los_now() - R->t_start) * 1e3, (los_now() - t0) * 1e3, tm.ms_h2d, tm.ms_total, tm.ms_integrate, tm.ms_d2h);
  const long int stored = (long int)pb->n_units * ptr2->tt2_size;
  #pragma omp atomic
  ptr2->count_memorised_transfers += stored;                                    /* transfer2.c:2110-2111 */
  if (R->store_to_disk && writer_push(R->writer, pb) == _FAILURE_) {           /* :559-567: written and freed by the writer thread */
    snprintf(errmsg, _ERRORMSGSIZE_, "out of memory in the disk writer queue"); return _FAILURE_;
  }
  return _SUCCESS_;
}

/* The engines of a run: those bessel2_init left on the devices with the projection tables in place (bessel2_gpu.c), or --
   when pbs2 comes from the reference's bessel2.c -- a private set built from pbs2->J_Llm_x. */
static int acquire_engines(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels *pbs,
                           struct bessels2 *pbs2, struct transfers2 *ptr2, int n_dev_wanted, int tau_mode,
                           los_engine_set *priv, los_engine_set **out, int *is_private)
{
  los_config_buf cb;
  class_call(los_config_from_song(ppr, ppr2, ppt2, pbs, pbs2, ptr2->tau0, tau_mode, 0, &cb, ptr2->error_message),
             ptr2->error_message, ptr2->error_message);
  los_engine_set *set = los_engines_find(pbs2);
  if (set != NULL && set->n_dev > 0 && song_los_config_matches(set->ctx[0], &cb.cfg)) {
    int ok = _TRUE_;
    for (int d = 0; d < set->n_dev && ok; ++d)
      if (song_los_set_time(set->ctx[d], ptr2->tau0, tau_mode, ppr2->tau_linstep_song) != SONG_LOS_SUCCESS) {
        snprintf(ptr2->error_message, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(set->ctx[d])); ok = _FALSE_;
      }
    los_config_release(&cb);
    if (!ok) return _FAILURE_;
    *out = set; *is_private = _FALSE_;
    return _SUCCESS_;
  }
  /* no engines for this pbs2 (reference bessel2_init): create them and upload its tables */
  memset(priv, 0, sizeof *priv);
  if (n_dev_wanted > LOS_MAX_DEVICES) n_dev_wanted = LOS_MAX_DEVICES;
  int failed = _FALSE_;
  ErrorMsg first = "";
  #pragma omp parallel for num_threads(n_dev_wanted) schedule(static, 1) if (n_dev_wanted > 1)
  for (int d = 0; d < n_dev_wanted; ++d) {
    song_los_config cfg = cb.cfg;
    cfg.device = d;
    ErrorMsg local = "";
    int rc = song_los_create(&cfg, &priv->ctx[d]);
    if (rc != SONG_LOS_SUCCESS) snprintf(local, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(NULL));
    else if (los_upload_host_projection(ppr, ppr2, pbs, pbs2, priv->ctx[d], local) == _FAILURE_) rc = SONG_LOS_FAILURE;
    if (rc != SONG_LOS_SUCCESS) {
      #pragma omp critical(los_engine_error)
      { if (!failed) { failed = _TRUE_; snprintf(first, _ERRORMSGSIZE_, "GPU %d: %.1500s", d, local); } }
    }
  }
  los_config_release(&cb);
  priv->n_dev = n_dev_wanted;
  if (failed) {
    for (int d = 0; d < n_dev_wanted; ++d) if (priv->ctx[d]) song_los_destroy(priv->ctx[d]);
    memset(priv, 0, sizeof *priv);
    class_stop(ptr2->error_message, "%.1900s", first);
  }
  *out = priv; *is_private = _TRUE_;
  return _SUCCESS_;
}

/* ------------------------------------------------------------------------------------------------ */
/*  transfer2_init (reference: transfer2.c:79-621)                                                   */
/* ------------------------------------------------------------------------------------------------ */
int transfer2_init(struct precision *ppr, struct precision2 *ppr2, struct background *pba, struct thermo *pth,
                   struct perturbs *ppt, struct perturbs2 *ppt2, struct bessels *pbs, struct bessels2 *pbs2,
                   struct transfers *ptr, struct transfers2 *ptr2)
{
  if (!ppt2->has_perturbations2 ||
      (!ppt2->has_cmb_spectra && !ppt2->has_cmb_bispectra && !ptr2->stop_at_transfers2)) {   /* :99-110 */
    ptr2->has_cls = _FALSE_;
    if (ptr2->transfer2_verbose > 0) printf("Second-order transfer module skipped.\n");
    return _SUCCESS_;
  }
  ptr2->has_cls = _TRUE_;
  if (ptr2->transfer2_verbose > 0) printf("Computing second-order transfer functions (B200 line-of-sight engine)\n");
#ifdef _OPENMP
  const double t_start = omp_get_wtime();
#define LAP(what) do { if (ptr2->transfer2_verbose > 1) printf("     * %-52s %8.1f ms since the start of transfer2_init\n", what, (omp_get_wtime() - t_start) * 1e3); } while (0)
#else
#define LAP(what) do { } while (0)
#endif
  if ((ppt2->k_size < 4) && (ptr2->stop_at_transfers2 == _FALSE_))
    printf("# WARNING: cannot do cubic interpolation in k1 with less than 4 values. "
           "Increase k_size, or do not trust results for 1st-order transfer functions.\n");

  ptr2->tau0 = pba->conformal_age;                                                            /* :128-130 */
  ptr2->tau_rec = pth->tau_rec;
  ptr2->rs_rec = pth->rs_rec;

  class_call(transfer2_indices_of_transfers(ppr, ppr2, ppt2, pbs, pbs2, ptr, ptr2), ptr2->error_message, ptr2->error_message);
  LAP("index tables and k3 grid sizes");

  if (ppr2->load_transfers_from_disk == _TRUE_) {                                             /* :151-157 */
    if (ptr2->transfer2_verbose > 0) printf(" -> leaving transfer2 module; transfer functions will be read from disk\n");
    return _SUCCESS_;
  }

  /* arithmetic mode */
  int mode = SONG_LOS_MODE_FAST;
  const char *mode_env = getenv("SONG_LOS_MODE");
  if ((mode_env && strcmp(mode_env, "exact") == 0) || (ppr->bessels_interpolation == cubic_interpolation))
    mode = SONG_LOS_MODE_EXACT;
  const int tau_mode = (ptr2->tau_sampling == sources_tau_sampling) ? SONG_LOS_TAU_SOURCES
                     : (ptr2->tau_sampling == bessel_tau_sampling) ? SONG_LOS_TAU_BESSEL : SONG_LOS_TAU_CUSTOM;

  /* devices: every visible GPU (SONG_LOS_DEVICES caps the number), one host thread each.  A run that reads its sources
     from disk level by level (perturb2_load_sources_from_disk, reference code) keeps the k1 order and uses one GPU;
     transfers written to disk do not constrain the schedule (see the writer thread above) */
  const int sources_on_disk = (ppr2->load_sources_from_disk == _TRUE_) || (ppr2->store_sources_to_disk == _TRUE_);
  const int ordered = sources_on_disk;
  int n_visible = 0;
  class_test(song_los_device_count(&n_visible) != SONG_LOS_SUCCESS || n_visible < 1, ptr2->error_message,
             "%.1500s", song_los_last_error(NULL));                                            /* no usable GPU: there is no CPU fallback */
  int n_dev = env_int("SONG_LOS_DEVICES", 0);
  if (n_dev <= 0 || n_dev > n_visible) n_dev = n_visible;
  if (n_dev > LOS_MAX_DEVICES) n_dev = LOS_MAX_DEVICES;
  if (ordered) n_dev = 1;

  los_engine_set private_set, *engines = NULL;
  int engines_private = _FALSE_;
  class_call(acquire_engines(ppr, ppr2, ppt2, pbs, pbs2, ptr2, n_dev, tau_mode, &private_set, &engines, &engines_private),
             ptr2->error_message, ptr2->error_message);
  if (engines->n_dev < n_dev) n_dev = engines->n_dev;
  LAP(engines_private ? "engines created from pbs2->J_Llm_x" : "engines of bessel2_init taken over");
#define RELEASE_ENGINES() do { if (engines_private) for (int d_ = 0; d_ < engines->n_dev; ++d_) song_los_destroy(engines->ctx[d_]); } while (0)

  /* ---- the pairs and their static schedule (transfer2.c:549-551: the k1 modes are independent) ---- */
  los_run R;
  memset(&R, 0, sizeof R);
  R.ppr = ppr; R.ppr2 = ppr2; R.ppt2 = ppt2; R.pbs = pbs; R.pbs2 = pbs2; R.ptr2 = ptr2;
  R.t_start = los_now();
#ifdef _OPENMP
  R.t_start = t_start;
#endif
  R.mode = mode; R.sources_on_disk = sources_on_disk; R.store_to_disk = (ppr2->store_transfers_to_disk == _TRUE_);
  R.n_pairs = ppt2->k_size * (ppt2->k_size + 1) / 2;
  int64_t *cost; int *device_of_pair;
  class_alloc(R.pair_k1, R.n_pairs * sizeof(int), ptr2->error_message);
  class_alloc(R.pair_k2, R.n_pairs * sizeof(int), ptr2->error_message);
  class_alloc(R.pairs_to_pack, ppt2->k_size * sizeof(int), ptr2->error_message);
  class_alloc(cost, R.n_pairs * sizeof(int64_t), ptr2->error_message);
  class_alloc(device_of_pair, R.n_pairs * sizeof(int), ptr2->error_message);
  int64_t total_units = 0, total_pt = 0;
  int *units_of_pair; double *kmin_of_pair, *kmax_of_pair;
  class_alloc(units_of_pair, R.n_pairs * sizeof(int), ptr2->error_message);
  class_alloc(kmin_of_pair, R.n_pairs * sizeof(double), ptr2->error_message);
  class_alloc(kmax_of_pair, R.n_pairs * sizeof(double), ptr2->error_message);
  for (int k1 = 0, p = 0; k1 < ppt2->k_size; ++k1) {
    R.pairs_to_pack[k1] = k1 + 1;
    for (int k2 = 0; k2 <= k1; ++k2, ++p) {
      R.pair_k1[p] = k1; R.pair_k2[p] = k2;
      units_of_pair[p] = ptr2->k_size_k1k2[k1][k2];
      kmin_of_pair[p] = ptr2->k_min_k1k2[k1][k2]; kmax_of_pair[p] = ptr2->k_max_k1k2[k1][k2];
      total_units += ptr2->k_size_k1k2[k1][k2]; total_pt += ppt2->k3_size[k1][k2];
    }
  }
  /* cost of a pair: its k3 values, weighted by the share of the multipole columns that are active at its k3 (los_host.c) */
  song_host_pair_costs(R.n_pairs, units_of_pair, kmin_of_pair, kmax_of_pair, ptr2->tau0, ptr2->l[ptr2->l_size - 1], cost);
  free(kmin_of_pair); free(kmax_of_pair);
  LAP("pair costs");

  /* Batch size.  The tensor-core kernel works on the k3-sorted values of a batch: the more of them share a row of the
     projection table (pbs2->xx_size rows), the fewer partly filled MMAs it issues -- a batch with 40 k3 values per row
     runs within ~10 % of the full run's rate, one with 10 per row at half of it.  Small batches, on the other hand, let
     the copies of one batch overlap the kernels of another.  So: SONG_LOS_MIN_BATCHES (4) batches per GPU when they stay
     above 40 values per row, fewer and larger ones otherwise; never more than fits in device memory. */
  int64_t cap_units = 0;
  const double src_per_unit = (double)ppt2->tp2_size * ppt2->tau_size * (double)total_pt / (double)(total_units > 0 ? total_units : 1);
  if (song_los_max_batch_units(engines->ctx[0], mode, src_per_unit, &cap_units) != SONG_LOS_SUCCESS) {
    snprintf(ptr2->error_message, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(engines->ctx[0]));
    RELEASE_ENGINES();
    return _FAILURE_;
  }
  LAP("device capacity (song_los_max_batch_units)");
  const int64_t dense_units = (int64_t)env_int("SONG_LOS_DENSE_K3_PER_ROW", 40) * pbs2->xx_size;
  int64_t batch_units = env_int("SONG_LOS_BATCH_UNITS", 0);
  if (batch_units <= 0) {
    int min_batches = env_int("SONG_LOS_MIN_BATCHES", 4);
    if (min_batches < 1) min_batches = 1;
    batch_units = (total_units / n_dev + min_batches - 1) / min_batches;
    if (batch_units < dense_units) batch_units = dense_units;
  }
  if (batch_units > cap_units) batch_units = cap_units;

  /* several chunks per GPU give every GPU the same mix of low-k and high-k pairs (the cost model then matters less);
     not more than keep a chunk dense */
  int chunks_per_dev = env_int("SONG_LOS_CHUNKS_PER_DEVICE", 0);
  if (chunks_per_dev <= 0) {
    chunks_per_dev = (n_dev > 1) ? (int)((total_units / n_dev) / (dense_units > 0 ? dense_units : 1)) : 1;
    if (chunks_per_dev > 3) chunks_per_dev = 3;
    if (chunks_per_dev < 1) chunks_per_dev = 1;
  }
  if (ordered) chunks_per_dev = 1;
  if (song_host_schedule_pairs(cost, R.n_pairs, n_dev, chunks_per_dev, device_of_pair, NULL) != 0) {
    RELEASE_ENGINES();
    class_stop(ptr2->error_message, "could not schedule %d (k1,k2) pairs on %d GPUs", R.n_pairs, n_dev);
  }
  for (int p = 0; p < R.n_pairs; ++p) cost[p] = units_of_pair[p];       /* from here on: k3 values (batch sizes) */
  free(units_of_pair);

  los_pair_batch *batches;
  class_alloc(batches, R.n_pairs * sizeof(los_pair_batch), ptr2->error_message);
  int n_batches = 0;
  enum { LOS_MAX_PARTS = 64 };
  int device_started[LOS_MAX_DEVICES];
  memset(device_started, 0, sizeof device_started);
  const int ramp_enabled = env_int("SONG_LOS_RAMP", 1) && !ordered;
  for (int p = 0; p < R.n_pairs;) {
    /* the run of consecutive pairs of one device starting at p, cut into equal parts of at most batch_units */
    int q = p; int64_t run_units = 0;
    while (q < R.n_pairs && device_of_pair[q] == device_of_pair[p]) run_units += cost[q++];
    /* about batch_units each (rounded to the nearest count: a part never falls far below the dense size), and what fits */
    int parts = (int)((2 * run_units + batch_units) / (2 * batch_units));
    const int parts_min = (int)((run_units + cap_units - 1) / cap_units);
    if (parts < parts_min) parts = parts_min;
    if (parts < 1) parts = 1;
    /* Sizes of the parts.  Nothing of a GPU's first batch overlaps anything: its packing and host->device copy are pure
       latency (0.23 s of a 2.2 s call at configs[3] with four equal batches), and so is the device->host copy of the last
       one.  The first run of a device therefore starts with a small batch and grows: weights 1 : 2 : 3.5 : 3.5 : ... -- the
       packing of batch n+1 (with its host->device copy beside it, song_los_stage_sources: 0.115 us per k3 value on the
       bench host) about fits under the kernels of batch n (0.15 us per k3 value for the cheap low-k pairs of a first batch,
       0.23-0.25 later), and the large batches keep the k3 density.
       SONG_LOS_RAMP=0: equal parts. */
    double bound[LOS_MAX_PARTS + 1];
    if (parts > LOS_MAX_PARTS) parts = LOS_MAX_PARTS;
    const int ramp = ramp_enabled && !device_started[device_of_pair[p]] && parts >= 2;
    device_started[device_of_pair[p]] = 1;
    {
      double w[LOS_MAX_PARTS];
      for (int i = 0; i < parts; ++i) { w[i] = ramp ? (i == 0 ? 1. : i == 1 ? 2. : 3.5) : 1.; }
      /* sizes min(cap, lambda w[i]) that add up to the run: a part that would exceed the device's capacity is clipped
         and the others grow (parts >= parts_min: the capacity suffices) */
      double lo = 0., hi = (double)run_units + 1.;
      for (int it = 0; it < 60; ++it) {
        const double lambda = 0.5 * (lo + hi);
        double sum = 0.;
        for (int i = 0; i < parts; ++i) sum += (lambda * w[i] < 0.98 * (double)cap_units) ? lambda * w[i] : 0.98 * (double)cap_units;
        if (sum < (double)run_units) lo = lambda; else hi = lambda;
      }
      bound[0] = 0.;
      for (int i = 0; i < parts; ++i) bound[i + 1] = bound[i] + ((hi * w[i] < 0.98 * (double)cap_units) ? hi * w[i] : 0.98 * (double)cap_units);
      bound[parts] = (double)run_units;
    }
    int64_t acc = 0; int part = 0;
    los_pair_batch pb; memset(&pb, 0, sizeof pb); pb.p_first = p; pb.device = device_of_pair[p];
    for (int i = p; i < q; ++i) {
      acc += cost[i];
      pb.n_units += cost[i]; pb.n_pt += ppt2->k3_size[R.pair_k1[i]][R.pair_k2[i]];
      pb.p_last = i;
      if (i + 1 == q || (part + 1 < parts && (double)acc >= bound[part + 1])) {
        pb.n_src = (int64_t)ppt2->tp2_size * ppt2->tau_size * pb.n_pt;
        batches[n_batches++] = pb;
        memset(&pb, 0, sizeof pb); pb.p_first = i + 1; pb.device = device_of_pair[p];
        ++part;
      }
    }
    p = q;
  }
  if (ptr2->transfer2_verbose > 0)
    printf(" -> %d (k1,k2) pairs in %d batches (up to %ld k3 values each) on %d GPU(s), %s arithmetic\n", R.n_pairs, n_batches,
           (long)batch_units, n_dev, mode == SONG_LOS_MODE_FAST ? "tensor-core" : "reference-order");
  LAP("batches cut");
  LAP("batches cut");
  los_transfer_store *st = store_of(ppt2, ptr2);
  if (st == NULL) { RELEASE_ENGINES(); class_stop(ptr2->error_message, "could not set up the storage of ptr2->transfer"); }
  LAP("pair schedule and batches");

  /* Per GPU: a pipeline of up to SONG_LOS_MAX_IN_FLIGHT batches.  While the kernels of batch n run, the host packs batch
     n+1 into the next staging and submits it (its host->device copy overlaps), and the result of batch n-1 travels
     back into ptr2->transfer.  Host threads per GPU for the packing: SONG_LOS_HOST_THREADS (default: an even share of
     the processors minus one -- with every processor packing, the driver's own threads are descheduled and the call
     time becomes erratic: measured 37 ms with 8 of 16, 37-98 ms with 16 of 16 at configs[1]). */
  int host_threads = env_int("SONG_LOS_HOST_THREADS", 0);
#ifdef _OPENMP
  if (host_threads <= 0) host_threads = omp_get_num_procs() / (n_dev > 0 ? n_dev : 1) - 1;   /* one processor per GPU is left to the CUDA driver's threads */
  if (n_dev > 1) {
    /* one thread per GPU, each with its own team of packing threads.  Runtimes older than OpenMP 5.0 (the libgomp.so.1
       that PyTorch wheels ship and that wins the soname lookup in a process that imported torch first) serialise the
       inner teams unless nesting is switched on explicitly: measured 5 GB/s instead of 29 GB/s of packing per GPU */
    omp_set_max_active_levels(2);
#if defined(__GNUC__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wdeprecated-declarations"
#endif
    omp_set_nested(1);
#if defined(__GNUC__)
#pragma GCC diagnostic pop
#endif
  }
#endif
  if (host_threads < 1) host_threads = 1;
  if (host_threads > 32) host_threads = 32;
  R.host_threads = host_threads;
  int abort_flag = _FALSE_;
  ErrorMsg first_error = "";
  if (R.store_to_disk) {
    R.writer = writer_start(&R);
    if (R.writer == NULL) { RELEASE_ENGINES(); class_stop(ptr2->error_message, "could not start the disk writer thread"); }
  }
  #pragma omp parallel for num_threads(n_dev) schedule(static, 1) if (n_dev > 1)
  for (int d = 0; d < n_dev; ++d) {
    ErrorMsg local = "";
    int ok = _TRUE_;
    song_los_ctx *ctx = engines->ctx[d];
    /* staging and device buffers sized for the largest batch of this device */
    int max_pairs = 1, n_mine = 0; int64_t max_pt = 1, max_units = 1, max_src = 1;
    for (int ib = 0; ib < n_batches; ++ib) {
      if (batches[ib].device != d) continue;
      ++n_mine;
      if (batches[ib].p_last - batches[ib].p_first + 1 > max_pairs) max_pairs = batches[ib].p_last - batches[ib].p_first + 1;
      if (batches[ib].n_pt > max_pt) max_pt = batches[ib].n_pt;
      if (batches[ib].n_units > max_units) max_units = batches[ib].n_units;
      if (batches[ib].n_src > max_src) max_src = batches[ib].n_src;
    }
    enum { DEPTH = SONG_LOS_MAX_IN_FLIGHT };
    los_staging stg[DEPTH];
    memset(stg, 0, sizeof stg);
    const int depth = n_mine < DEPTH ? (n_mine > 0 ? n_mine : 1) : DEPTH;
    for (int i = 0; i < depth && ok; ++i)
      if (staging_alloc(&stg[i], max_pairs, max_pt, max_units, max_src) == _FAILURE_) {
        snprintf(local, _ERRORMSGSIZE_, "could not allocate the page-locked staging of a GPU batch (%ld k3 values): %.1000s",
                 (long)max_units, song_los_last_error(NULL));
        ok = _FALSE_;
      }
    if (d == 0) LAP("page-locked staging (GPU 0)");
    /* no allocation (= device-wide synchronisation) inside the pipeline */
    if (ok && song_los_reserve(ctx, mode, depth, max_pairs, max_units, max_pt, max_src) != SONG_LOS_SUCCESS) {
      snprintf(local, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(ctx)); ok = _FALSE_;
    }
    if (d == 0) LAP("device buffers reserved (GPU 0)");
    int cur = 0, first = _TRUE_;
    for (int ib = 0; ok && ib < n_batches; ++ib) {
      if (batches[ib].device != d) continue;
      if (abort_flag == _TRUE_) break;
      if (ptr2->transfer2_verbose > 1)
        printf("     * GPU %d: transfer functions of index_k1=%d..%d (%d pairs, %ld k3 values)\n", d, R.pair_k1[batches[ib].p_first],
               R.pair_k1[batches[ib].p_last], batches[ib].p_last - batches[ib].p_first + 1, (long)batches[ib].n_units);
      /* the staging about to be reused must have been drained: at most `depth` batches are in flight */
      if (stg[cur].pb != NULL && finish_batch(&R, ctx, &stg[cur], local) == _FAILURE_) { ok = _FALSE_; break; }
      if (submit_batch(&R, ctx, &batches[ib], &stg[cur], local) == _FAILURE_) { ok = _FALSE_; break; }
      cur = (cur + 1) % depth;
      if (d == 0 && first) LAP("first batch packed and submitted (GPU 0)");
      first = _FALSE_;
    }
    /* drain in submission order: stg[cur] holds the oldest batch */
    for (int i = 0; i < depth; ++i) {
      los_staging *s = &stg[(cur + i) % depth];
      if (s->pb == NULL) continue;
      if (ok) { if (finish_batch(&R, ctx, s, local) == _FAILURE_) ok = _FALSE_; }
      else { song_los_wait(ctx, NULL); s->pb = NULL; }
    }
    for (int i = 0; i < DEPTH; ++i) if (stg[i].block) staging_free(&stg[i]);
    if (!ok) {
      #pragma omp critical(los_first_error)
      { if (abort_flag == _FALSE_) { abort_flag = _TRUE_; snprintf(first_error, _ERRORMSGSIZE_, "%s", local); } }
    }
  }
  LAP("all batches computed and stored in ptr2->transfer");
  if (R.writer != NULL) {
    ErrorMsg werr = "";
    if (writer_finish(R.writer, werr) == _FAILURE_ && abort_flag == _FALSE_) { abort_flag = _TRUE_; snprintf(first_error, _ERRORMSGSIZE_, "%s", werr); }
    LAP("transfers_%03d.dat written");
  }
  RELEASE_ENGINES();
  free(batches); free(cost); free(device_of_pair); free(R.pair_k1); free(R.pair_k2); free(R.pairs_to_pack);
  class_test(abort_flag == _TRUE_, ptr2->error_message, "%.1900s", first_error);

  if (ppr2->store_transfers_to_disk == _TRUE_)                                                 /* :598-600 */
    for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt) fclose(ptr2->transfers_files[index_tt]);
  if (ptr2->transfer2_verbose > 1)
    printf(" -> filled ptr2->transfer with %ld values (%g MB)\n", ptr2->count_memorised_transfers, ptr2->count_memorised_transfers / 1e6 * 8);
  class_test(ptr2->count_allocated_transfers != ptr2->count_memorised_transfers, ptr2->error_message,        /* :607-610 */
             "there is a mismatch between allocated (%ld) and used (%ld) space!", ptr2->count_allocated_transfers, ptr2->count_memorised_transfers);

  if (ptr2->stop_at_transfers2 == _TRUE_) {                                                    /* :613-618 */
    ppt->has_cls = _FALSE_; ppt->has_cmb_bispectra = _FALSE_;
    ppt2->has_cmb_spectra = _FALSE_; ppt2->has_cmb_bispectra = _FALSE_;
  }
  return _SUCCESS_;
}
