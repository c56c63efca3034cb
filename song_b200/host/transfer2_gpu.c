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
 * Batching: the reference handles one (k1,k2) pair at a time; here consecutive k1 levels are packed into one GPU
 * batch (SONG_LOS_BATCH_UNITS k3 values; default: an even share of the run per GPU, at most what
 * song_los_max_batch_units says fits in device memory) because the tensor-core kernel wants many k3 values
 * per table row.  Batches of a GPU run as a three-deep pipeline (song_los_submit / song_los_wait) through page-locked
 * staging, so that the copies of neighbouring batches overlap the kernels.  Side effects keep the reference's per-k1 order: sources of a k1 level are loaded before and freed
 * after use (transfer2.c:332-344, 553), transfers are stored and freed per k1 (transfer2.c:559-567).
 * With several visible GPUs and nothing to write to disk, batches are distributed over one host thread per GPU
 * (static longest-processing-time schedule on the number of k3 values); the results land directly in
 * ptr2->transfer, which is the "gather to the host".
 *
 * Environment: SONG_LOS_MODE=fast|exact (default fast; cubic Bessel interpolation forces exact),
 * SONG_LOS_DEVICES=<n> (default: all visible), SONG_LOS_BATCH_UNITS=<n>.
 */
#include "transfer2.h"
#include "song_los.h"
#include "los_host.h"

#ifdef _OPENMP
#include <omp.h>
#endif

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
int transfer2_allocate_k1_level(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_k1)
{
  class_test(ptr2->has_allocated_transfers[index_k1] == _TRUE_, ptr2->error_message,
             "the index_k1=%d level of ptr2->transfers is already allocated, stop to prevent error", index_k1);
  long int count = 0;
  for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt) {
    class_alloc(ptr2->transfer[index_tt][index_k1], (index_k1 + 1) * sizeof(double *), ptr2->error_message);
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) {
      class_alloc(ptr2->transfer[index_tt][index_k1][index_k2],
                  ptr2->k_size_k1k2[index_k1][index_k2] * sizeof(double), ptr2->error_message);
      count += ptr2->k_size_k1k2[index_k1][index_k2];
    }
  }
  #pragma omp atomic
  ptr2->count_allocated_transfers += count;
  ptr2->has_allocated_transfers[index_k1] = _TRUE_;
  return _SUCCESS_;
}

int transfer2_free_k1_level(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_k1)
{
  class_test(ptr2->has_allocated_transfers[index_k1] == _FALSE_, ptr2->error_message,
             "the index_k1=%d level of ptr2->transfers is already free, stop to prevent error", index_k1);
  for (int index_tt = 0; index_tt < ptr2->tt2_size; ++index_tt) {
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) free(ptr2->transfer[index_tt][index_k1][index_k2]);
    free(ptr2->transfer[index_tt][index_k1]);
  }
  ptr2->has_allocated_transfers[index_k1] = _FALSE_;
  return _SUCCESS_;
}

int transfer2_allocate_type_level(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_tt)
{
  int k1_size = ppt2->k_size;
  class_alloc(ptr2->transfer[index_tt], k1_size * sizeof(double **), ptr2->error_message);
  for (int index_k1 = 0; index_k1 < k1_size; ++index_k1) {
    class_alloc(ptr2->transfer[index_tt][index_k1], (index_k1 + 1) * sizeof(double *), ptr2->error_message);
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) {
      class_alloc(ptr2->transfer[index_tt][index_k1][index_k2],
                  ptr2->k_size_k1k2[index_k1][index_k2] * sizeof(double), ptr2->error_message);
      #pragma omp atomic
      ptr2->count_allocated_transfers += ptr2->k_size_k1k2[index_k1][index_k2];
    }
  }
  return _SUCCESS_;
}

int transfer2_free_type_level(struct perturbs2 *ppt2, struct transfers2 *ptr2, int index_tt)
{
  for (int index_k1 = 0; index_k1 < ppt2->k_size; ++index_k1) {
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) free(ptr2->transfer[index_tt][index_k1][index_k2]);
    free(ptr2->transfer[index_tt][index_k1]);
  }
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
  for (int index_k1 = 0; index_k1 < ppt2->k_size; ++index_k1)
    for (int index_k2 = 0; index_k2 <= index_k1; ++index_k2) {
      int n_to_read = ptr2->k_size_k1k2[index_k1][index_k2];
      int n_read = (int)fread(ptr2->transfer[index_tt][index_k1][index_k2], sizeof(double), n_to_read,
                              ptr2->transfers_files[index_tt]);
      class_test(n_read != n_to_read, ptr2->error_message,
                 "Could not read in '%s' file, read %d entries but expected %d (index_tt=%d,index_k1=%d,index_k2=%d)",
                 ptr2->transfers_paths[index_tt], n_read, n_to_read, index_tt, index_k1, index_k2);
      #pragma omp atomic
      ptr2->count_memorised_transfers += n_to_read;
    }
  fclose(ptr2->transfers_files[index_tt]);
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
/*  engine set-up from the SONG structures                                                           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
  song_los_ctx *ctx;
  int device;
} los_device;

static int create_engine(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels2 *pbs2,
                         struct transfers2 *ptr2, int device, song_los_ctx **out)
{
  song_los_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.device = device;
  cfg.tau0 = ptr2->tau0;
  cfg.tau_size = ppt2->tau_size; cfg.tau_sampling = ppt2->tau_sampling;
  cfg.tau_sampling_mode = (ptr2->tau_sampling == sources_tau_sampling) ? SONG_LOS_TAU_SOURCES
                        : (ptr2->tau_sampling == bessel_tau_sampling) ? SONG_LOS_TAU_BESSEL : SONG_LOS_TAU_CUSTOM;
  cfg.tau_linstep_song = ppr2->tau_linstep_song;
  cfg.sources_k3_interpolation = (ppr2->sources_k3_interpolation == cubic_interpolation) ? SONG_LOS_INTERP_CUBIC : SONG_LOS_INTERP_LINEAR;
  cfg.sources_time_interpolation = (ppr2->sources_time_interpolation == cubic_interpolation) ? SONG_LOS_INTERP_CUBIC : SONG_LOS_INTERP_LINEAR;
  cfg.bessels_interpolation = (ppr->bessels_interpolation == cubic_interpolation) ? SONG_LOS_INTERP_CUBIC : SONG_LOS_INTERP_LINEAR;
  cfg.k3_extrapolation = (ppr->bispectra_k3_extrapolation == no_k3_extrapolation) ? SONG_LOS_NO_K3_EXTRAPOLATION : SONG_LOS_FLAT_K3_EXTRAPOLATION;
  cfg.xx_size = pbs2->xx_size; cfg.xx = pbs2->xx; cfg.xx_step = pbs2->xx_step;
  cfg.l_size = ptr2->l_size; cfg.l = ptr2->l;
  cfg.m_size = ptr2->m_size; cfg.m = ptr2->m;
  cfg.L_size = pbs2->L_size;
  cfg.n_transfers = ptr2->n_transfers;

  /* lm_array as [l_size][m_size] and lm(L,m) of the sources as [L_size][m_size], -1 where not defined */
  int *lm_array, *lm_src;
  class_alloc(lm_array, (size_t)ptr2->l_size * ptr2->m_size * sizeof(int), ptr2->error_message);
  class_alloc(lm_src, (size_t)pbs2->L_size * ptr2->m_size * sizeof(int), ptr2->error_message);
  for (int index_l = 0; index_l < ptr2->l_size; ++index_l)
    for (int index_m = 0; index_m < ptr2->m_size; ++index_m)
      lm_array[index_l * ptr2->m_size + index_m] =
          (index_m <= ppr2->index_m_max[ptr2->l[index_l]]) ? ptr2->lm_array[index_l][index_m] : -1;
  for (int L = 0; L < pbs2->L_size; ++L)
    for (int index_m = 0; index_m < ptr2->m_size; ++index_m)
      lm_src[L * ptr2->m_size + index_m] = (ptr2->m[index_m] <= L) ? ppt2->lm_array[L][index_m] : -1;   /* lm(L,m), perturbations2_macros.h:53 */
  cfg.lm_array = lm_array; cfg.lm_src = lm_src;

  cfg.has_source_T = ppt2->has_source_T; cfg.has_source_E = ppt2->has_source_E; cfg.has_source_B = ppt2->has_source_B;
  cfg.index_tp2_T = ppt2->index_tp2_T; cfg.index_tp2_E = ppt2->index_tp2_E; cfg.index_tp2_B = ppt2->index_tp2_B;
  cfg.tp2_size = ppt2->tp2_size;
  cfg.l_max_los_t = ppr2->l_max_los_t; cfg.l_max_los_p = ppr2->l_max_los_p;
  cfg.index_tt2_T = ptr2->index_tt2_T; cfg.index_tt2_E = ptr2->index_tt2_E; cfg.index_tt2_B = ptr2->index_tt2_B;
  cfg.tt2_size = ptr2->tt2_size;
  cfg.tt_index_l = ptr2->corresponding_index_l; cfg.tt_index_m = ptr2->corresponding_index_m;
  cfg.J_size = pbs2->J_size;
  cfg.index_J_TT = pbs2->has_J_TT ? pbs2->index_J_TT : -1;
  cfg.index_J_EE = pbs2->has_J_EE ? pbs2->index_J_EE : -1;
  cfg.index_J_EB = pbs2->has_J_EB ? pbs2->index_J_EB : -1;

  int rc = song_los_create(&cfg, out);
  free(lm_array); free(lm_src);
  class_test(rc != SONG_LOS_SUCCESS, ptr2->error_message, "%.1500s", song_los_last_error(NULL));

  /* projection functions of bessel2_init: ragged [J][L][l][m] rows -> flat list (bessel2.h:64-91) */
  size_t n_j = (size_t)pbs2->J_size * pbs2->L_size * ptr2->l_size * ptr2->m_size;
  const double **rows, **ddrows; int *xmin, *xsize;
  class_calloc(rows, n_j, sizeof(double *), ptr2->error_message);
  class_calloc(ddrows, n_j, sizeof(double *), ptr2->error_message);
  class_calloc(xmin, n_j, sizeof(int), ptr2->error_message);
  class_calloc(xsize, n_j, sizeof(int), ptr2->error_message);
  int cubic_J = (ppr->bessels_interpolation == cubic_interpolation);
  for (int J = 0; J < pbs2->J_size; ++J)
    for (int L = 0; L < pbs2->L_size; ++L)
      for (int index_l = 0; index_l < ptr2->l_size; ++index_l)
        for (int index_m = 0; index_m < ptr2->m_size; ++index_m) {
          if (ptr2->m[index_m] > MIN(pbs2->L[L], ptr2->l[index_l])) continue;       /* not allocated by bessel2_init */
          size_t q = (((size_t)J * pbs2->L_size + L) * ptr2->l_size + index_l) * ptr2->m_size + index_m;
          rows[q] = pbs2->J_Llm_x[J][L][index_l][index_m];
          if (cubic_J) ddrows[q] = pbs2->ddJ_Llm_x[J][L][index_l][index_m];
          xmin[q] = pbs2->index_xmin_J[J][L][index_l][index_m];
          xsize[q] = pbs2->x_size_J[J][L][index_l][index_m];
        }
  rc = song_los_set_projection(*out, rows, cubic_J ? ddrows : NULL, xmin, xsize);
  free(rows); free(ddrows); free(xmin); free(xsize);
  if (rc != SONG_LOS_SUCCESS) {
    snprintf(ptr2->error_message, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(*out));
    song_los_destroy(*out); *out = NULL;
    return _FAILURE_;
  }
  return _SUCCESS_;
}

/* one GPU batch = the k1 levels [k1_first, k1_last] */
typedef struct { int k1_first, k1_last; long n_units; int device; } los_k1_batch;

/* Page-locked staging of one batch in flight (song_los_host_alloc): the packed inputs the C-ABI reads
   and the [tt2_size][n_units] result it writes.  SONG_LOS_MAX_IN_FLIGHT of these per GPU, used in turn. */
typedef struct {
  int *np, *nt, *ps, *pz; int64_t *pt_off, *tr_off, *src_off;
  double *k3_pt, *k3_tr, *src, *out;
  int n_pairs; int64_t n_pt, n_units, n_src;
  const los_k1_batch *kb;          /* NULL when the staging is free */
  int pinned;                      /* page-locked (song_los_host_alloc) or plain malloc */
} los_staging;

static int stg_alloc(los_staging *s, void **ptr, size_t bytes)
{
  if (s->pinned) return song_los_host_alloc(ptr, bytes);
  *ptr = malloc(bytes > 0 ? bytes : 1);
  return *ptr ? SONG_LOS_SUCCESS : SONG_LOS_FAILURE;
}
static void stg_free(los_staging *s, void *ptr) { if (s->pinned) song_los_host_free(ptr); else free(ptr); }

static void staging_free(los_staging *s)
{
  stg_free(s, s->np); stg_free(s, s->nt); stg_free(s, s->ps); stg_free(s, s->pz);
  stg_free(s, s->pt_off); stg_free(s, s->tr_off); stg_free(s, s->src_off);
  stg_free(s, s->k3_pt); stg_free(s, s->k3_tr); stg_free(s, s->src); stg_free(s, s->out);
  memset(s, 0, sizeof *s);
}

/* capacity of the largest batch of a device.  Page-locking costs about half a second per GB: it pays only when
   copies have something to overlap with, i.e. when the device runs more than one batch. */
static int staging_alloc(los_staging *s, int pinned, int max_pairs, int64_t max_pt, int64_t max_units, int64_t max_src, int tt2_size)
{
  memset(s, 0, sizeof *s);
  s->pinned = pinned;
  int rc = 0;
  rc |= stg_alloc(s, (void **)&s->np, sizeof(int) * (size_t)max_pairs);
  rc |= stg_alloc(s, (void **)&s->nt, sizeof(int) * (size_t)max_pairs);
  rc |= stg_alloc(s, (void **)&s->ps, sizeof(int) * (size_t)max_pairs);
  rc |= stg_alloc(s, (void **)&s->pz, sizeof(int) * (size_t)max_pairs);
  rc |= stg_alloc(s, (void **)&s->pt_off, sizeof(int64_t) * (size_t)max_pairs);
  rc |= stg_alloc(s, (void **)&s->tr_off, sizeof(int64_t) * (size_t)max_pairs);
  rc |= stg_alloc(s, (void **)&s->src_off, sizeof(int64_t) * (size_t)max_pairs);
  rc |= stg_alloc(s, (void **)&s->k3_pt, sizeof(double) * (size_t)max_pt);
  rc |= stg_alloc(s, (void **)&s->k3_tr, sizeof(double) * (size_t)max_units);
  rc |= stg_alloc(s, (void **)&s->src, sizeof(double) * (size_t)max_src);
  rc |= stg_alloc(s, (void **)&s->out, sizeof(double) * (size_t)max_units * tt2_size);
  if (rc != SONG_LOS_SUCCESS) { staging_free(s); return _FAILURE_; }
  return _SUCCESS_;
}

static void batch_sizes(struct perturbs2 *ppt2, struct transfers2 *ptr2, const los_k1_batch *kb,
                        int *n_pairs, int64_t *n_pt, int64_t *n_units, int64_t *n_src)
{
  *n_pairs = 0; *n_pt = 0; *n_units = 0; *n_src = 0;
  for (int k1 = kb->k1_first; k1 <= kb->k1_last; ++k1)
    for (int k2 = 0; k2 <= k1; ++k2) {
      ++*n_pairs; *n_pt += ppt2->k3_size[k1][k2]; *n_units += ptr2->k_size_k1k2[k1][k2];
      *n_src += (int64_t)ppt2->tp2_size * ppt2->tau_size * ppt2->k3_size[k1][k2];
    }
}

/* packs the k1 levels of `kb` into the staging and enqueues the batch on the GPU (song_los_submit) */
static int submit_batch(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels *pbs,
                        struct bessels2 *pbs2, struct transfers2 *ptr2, song_los_ctx *ctx, int mode, const los_k1_batch *kb,
                        los_staging *s, int host_threads, char *errmsg)
{
  const int tp2 = ppt2->tp2_size, ntau = ppt2->tau_size;
  batch_sizes(ppt2, ptr2, kb, &s->n_pairs, &s->n_pt, &s->n_units, &s->n_src);
  /* sizes and offsets first (serial), then every pair is packed independently by the host threads of this GPU */
  int p = 0;
  int64_t o_pt = 0, o_tr = 0, o_src = 0;
  for (int k1 = kb->k1_first; k1 <= kb->k1_last; ++k1)
    for (int k2 = 0; k2 <= k1; ++k2, ++p) {
      s->np[p] = ppt2->k3_size[k1][k2]; s->nt[p] = ptr2->k_size_k1k2[k1][k2];
      s->ps[p] = ptr2->k_physical_start_k1k2[k1][k2]; s->pz[p] = ptr2->k_physical_size_k1k2[k1][k2];
      s->pt_off[p] = o_pt; s->tr_off[p] = o_tr; s->src_off[p] = o_src;
      o_pt += s->np[p]; o_tr += s->nt[p]; o_src += (int64_t)tp2 * ntau * s->np[p];
    }
  int failed = 0;
  #pragma omp parallel for schedule(dynamic, 8) num_threads(host_threads) if (host_threads > 1)
  for (int q = 0; q < s->n_pairs; ++q) {
    /* pair q of the batch -> (k1,k2): k1 levels are consecutive, level k1 holds k1+1 pairs */
    int k1 = kb->k1_first, r = q;
    while (r > k1) { r -= k1 + 1; ++k1; }
    const int k2 = r;
    memcpy(s->k3_pt + s->pt_off[q], ppt2->k3[k1][k2], sizeof(double) * s->np[q]);
    if (transfer2_get_k3_list(ppr, ppr2, ppt2, pbs, pbs2, ptr2, k1, k2, s->k3_tr + s->tr_off[q], NULL) == _FAILURE_) {
      #pragma omp atomic write
      failed = 1;
      continue;
    }
    /* ppt2->sources[tp][k1][k2][tau*k3_size + k3] (transfer2.h:47-51) -> packed [pair][tp][tau][k3] */
    for (int tp = 0; tp < tp2; ++tp)
      memcpy(s->src + s->src_off[q] + (int64_t)tp * ntau * s->np[q], ppt2->sources[tp][k1][k2], sizeof(double) * (size_t)ntau * s->np[q]);
  }
  if (failed) { snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", ptr2->error_message); return _FAILURE_; }
  song_los_batch b;
  memset(&b, 0, sizeof b);
  b.n_pairs = s->n_pairs; b.k3_pt_size = s->np; b.k3_tr_size = s->nt; b.phys_start = s->ps; b.phys_size = s->pz;
  b.pt_off = s->pt_off; b.tr_off = s->tr_off; b.src_off = s->src_off;
  b.n_pt_total = s->n_pt; b.n_units = s->n_units; b.n_src_total = s->n_src;
  b.k3_pt = s->k3_pt; b.k3_tr = s->k3_tr; b.src = s->src;
  if (song_los_submit(ctx, &b, mode, s->out) != SONG_LOS_SUCCESS) {
    snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(ctx)); return _FAILURE_;
  }
  s->kb = kb;
  return _SUCCESS_;
}

/* waits for the oldest batch in flight and scatters its result into ptr2->transfer[tt][k1][k2][k3]
   (transfer2.c:2107), then applies the per-k1 side effects of transfer2.c:559-567 */
static int finish_batch(struct precision2 *ppr2, struct perturbs2 *ppt2, struct transfers2 *ptr2, song_los_ctx *ctx,
                        los_staging *s, int host_threads, char *errmsg)
{
  const los_k1_batch *kb = s->kb;
  s->kb = NULL;
  if (song_los_wait(ctx, NULL) != SONG_LOS_SUCCESS) {
    snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(ctx)); return _FAILURE_;
  }
  long stored = 0;
  #pragma omp parallel for schedule(dynamic, 8) reduction(+ : stored) num_threads(host_threads) if (host_threads > 1)
  for (int q = 0; q < s->n_pairs; ++q) {
    int k1 = kb->k1_first, r = q;
    while (r > k1) { r -= k1 + 1; ++k1; }
    const int k2 = r;
    for (int tt = 0; tt < ptr2->tt2_size; ++tt) {
      memcpy(ptr2->transfer[tt][k1][k2], s->out + (int64_t)tt * s->n_units + s->tr_off[q], sizeof(double) * s->nt[q]);
      stored += s->nt[q];
    }
  }
  #pragma omp atomic
  ptr2->count_memorised_transfers += stored;                                    /* transfer2.c:2110-2111 */
  if (ppr2->store_transfers_to_disk == _TRUE_)                                  /* :559-567, in k1 order */
    for (int k1 = kb->k1_first; k1 <= kb->k1_last; ++k1)
      if (transfer2_store_transfers_to_disk(ppt2, ptr2, k1) == _FAILURE_ || transfer2_free_k1_level(ppt2, ptr2, k1) == _FAILURE_) {
        snprintf(errmsg, _ERRORMSGSIZE_, "%.1500s", ptr2->error_message); return _FAILURE_;
      }
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
#define LAP(what) do { if (ptr2->transfer2_verbose > 1) printf("     * %-44s %8.1f ms since the start of transfer2_init\n", what, (omp_get_wtime() - t_start) * 1e3); } while (0)
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

  /* devices: one host thread per GPU when nothing has to be written in k1 order */
  const int sources_on_disk = (ppr2->load_sources_from_disk == _TRUE_) || (ppr2->store_sources_to_disk == _TRUE_);
  const int ordered = sources_on_disk || (ppr2->store_transfers_to_disk == _TRUE_);
  int n_dev = env_int("SONG_LOS_DEVICES", 0);
  if (n_dev <= 0) n_dev = 64;          /* probe: engines are created on devices 0,1,... until creation fails */
  if (ordered) n_dev = 1;
  los_device *dev;
  class_calloc(dev, n_dev, sizeof(los_device), ptr2->error_message);
  if (create_engine(ppr, ppr2, ppt2, pbs2, ptr2, 0, &dev[0].ctx) == _FAILURE_) { free(dev); return _FAILURE_; }   /* no usable GPU: there is no CPU fallback */

  LAP("engine on GPU 0 (context, projection tables)");
  /* k1 batches: consecutive k1 levels up to `batch_units` k3 values.  Large batches make the tensor-core kernel
     efficient (many k3 per table row); several batches per GPU let the copies overlap the kernels; the upper
     bound is what fits in device memory twice (two batches in flight). */
  long total_units = 0, total_pt = 0;
  for (int k1 = 0; k1 < ppt2->k_size; ++k1)
    for (int k2 = 0; k2 <= k1; ++k2) { total_units += ptr2->k_size_k1k2[k1][k2]; total_pt += ppt2->k3_size[k1][k2]; }
  int64_t cap_units = 0;
  if (song_los_max_batch_units(dev[0].ctx, mode, (double)ppt2->tp2_size * ppt2->tau_size * (double)total_pt / (double)(total_units > 0 ? total_units : 1),
                               &cap_units) != SONG_LOS_SUCCESS) {
    snprintf(ptr2->error_message, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(dev[0].ctx));
    song_los_destroy(dev[0].ctx); free(dev); return _FAILURE_;
  }
  long batch_units = (long)env_int("SONG_LOS_BATCH_UNITS", 0);
  if (batch_units <= 0) {
    /* as large as device memory allows (the kernel's efficiency grows with the number of k3 values per table
       row), but not larger than an even share of the run per GPU */
    int n_dev_guess = ordered ? 1 : env_int("SONG_LOS_DEVICES", 1);
    if (n_dev_guess < 1) n_dev_guess = 1;
    batch_units = (total_units + n_dev_guess - 1) / n_dev_guess;
  }
  if (batch_units > (long)cap_units) batch_units = (long)cap_units;
  los_k1_batch *batches;
  class_alloc(batches, ppt2->k_size * sizeof(los_k1_batch), ptr2->error_message);
  int n_batches = 0;
  for (int k1 = 0; k1 < ppt2->k_size;) {
    los_k1_batch kb = {k1, k1, 0, 0};
    for (; k1 < ppt2->k_size; ++k1) {
      long level = 0;
      for (int k2 = 0; k2 <= k1; ++k2) level += ptr2->k_size_k1k2[k1][k2];
      if (kb.n_units > 0 && kb.n_units + level > batch_units) break;
      kb.n_units += level; kb.k1_last = k1;
    }
    batches[n_batches++] = kb;
  }

  if (n_dev > n_batches) n_dev = n_batches > 0 ? n_batches : 1;
  int n_created = 1;
  for (int d = 1; d < n_dev; ++d) {
    if (create_engine(ppr, ppr2, ppt2, pbs2, ptr2, d, &dev[d].ctx) == _FAILURE_) break;
    dev[d].device = d; ++n_created;
  }
  n_dev = n_created;
  if (ptr2->transfer2_verbose > 0)
    printf(" -> %d k1 batches (up to %ld k3 values each) on %d GPU(s), %s arithmetic\n", n_batches, batch_units, n_dev,
           mode == SONG_LOS_MODE_FAST ? "tensor-core" : "reference-order");

  /* static longest-processing-time schedule: batches by decreasing size to the least loaded device */
  {
    int64_t *cost = malloc(sizeof(int64_t) * (n_batches > 0 ? n_batches : 1));
    int *where = malloc(sizeof(int) * (n_batches > 0 ? n_batches : 1));
    for (int i = 0; i < n_batches; ++i) cost[i] = batches[i].n_units;
    song_host_lpt_schedule(cost, n_batches, n_dev, where, NULL);
    for (int i = 0; i < n_batches; ++i) batches[i].device = where[i];
    free(cost); free(where);
  }

  /* Per GPU: a pipeline of up to SONG_LOS_MAX_IN_FLIGHT batches (in k1 order).  While the kernels of batch n run, the host
     packs batch n+1 into the other staging and submits it (its host->device copy overlaps), and the result of
     batch n-1 travels back and is scattered into ptr2->transfer. */
  /* host threads per GPU for packing the sources and scattering the results (SONG_LOS_HOST_THREADS, default: an even
     share of the processors) */
  int host_threads = env_int("SONG_LOS_HOST_THREADS", 0);
#ifdef _OPENMP
  if (host_threads <= 0) host_threads = omp_get_num_procs() / (n_dev > 0 ? n_dev : 1);
  if (n_dev > 1) omp_set_max_active_levels(2);
#endif
  if (host_threads < 1) host_threads = 1;
  if (host_threads > 32) host_threads = 32;
  int abort_flag = _FALSE_;
  ErrorMsg first_error = "";
  #pragma omp parallel for num_threads(n_dev) schedule(static, 1) if (n_dev > 1)
  for (int d = 0; d < n_dev; ++d) {
    ErrorMsg local = "";
    int ok = _TRUE_;
    /* staging sized for the largest batch of this device */
    int max_pairs = 1; int64_t max_pt = 1, max_units = 1, max_src = 1;
    for (int ib = 0; ib < n_batches; ++ib) {
      if (batches[ib].device != d) continue;
      int np_; int64_t a, b, c;
      batch_sizes(ppt2, ptr2, &batches[ib], &np_, &a, &b, &c);
      if (np_ > max_pairs) max_pairs = np_;
      if (a > max_pt) max_pt = a;
      if (b > max_units) max_units = b;
      if (c > max_src) max_src = c;
    }
    enum { DEPTH = SONG_LOS_MAX_IN_FLIGHT };
    los_staging stg[DEPTH];
    memset(stg, 0, sizeof stg);
    int n_mine = 0;
    for (int ib = 0; ib < n_batches; ++ib) if (batches[ib].device == d) ++n_mine;
    const int depth = n_mine < DEPTH ? (n_mine > 0 ? n_mine : 1) : DEPTH;    /* page-locking memory is slow: no more sets than batches */
    for (int i = 0; i < depth && ok; ++i)
      if (staging_alloc(&stg[i], n_mine > 1, max_pairs, max_pt, max_units, max_src, ptr2->tt2_size) == _FAILURE_) {
        snprintf(local, _ERRORMSGSIZE_, "could not allocate the page-locked staging of a GPU batch (%ld k3 values): %.1000s",
                 (long)max_units, song_los_last_error(NULL));
        ok = _FALSE_;
      }
    if (d == 0) LAP("page-locked staging allocated (GPU 0)");
    /* device buffers for the largest batch: no allocation (= device-wide synchronisation) inside the pipeline */
    if (ok && song_los_reserve(dev[d].ctx, mode, depth, max_pairs, max_units, max_pt, max_src) != SONG_LOS_SUCCESS) {
      snprintf(local, _ERRORMSGSIZE_, "%.1500s", song_los_last_error(dev[d].ctx)); ok = _FALSE_;
    }
    if (d == 0) LAP("device buffers reserved (GPU 0)");
    int cur = 0, in_flight = 0;
    for (int ib = 0; ok && ib < n_batches; ++ib) {
      if (batches[ib].device != d) continue;
      if (abort_flag == _TRUE_) break;
      const los_k1_batch *kb = &batches[ib];
      for (int k1 = kb->k1_first; ok && k1 <= kb->k1_last; ++k1) {
        if (ptr2->transfer2_verbose > 1)
          printf("     * computing transfer functions today for index_k1=%d of %d, k1=%g\n", k1, ppt2->k_size, ppt2->k[k1]);
        if (transfer2_allocate_k1_level(ppt2, ptr2, k1) == _FAILURE_) { snprintf(local, _ERRORMSGSIZE_, "%.1500s", ptr2->error_message); ok = _FALSE_; break; }
        if (sources_on_disk) {                                                                 /* :332-344 */
          if (perturb2_allocate_k1_level(ppt2, k1) == _FAILURE_ || perturb2_load_sources_from_disk(ppt2, k1) == _FAILURE_) {
            snprintf(local, _ERRORMSGSIZE_, "%.1500s", ppt2->error_message); ok = _FALSE_; break;
          }
        }
      }
      /* the staging about to be reused must have been drained: at most DEPTH batches are in flight */
      if (ok && stg[cur].kb != NULL) {
        if (finish_batch(ppr2, ppt2, ptr2, dev[d].ctx, &stg[cur], host_threads, local) == _FAILURE_) ok = _FALSE_;
        --in_flight;
      }
      if (d == 0 && ib == 0) LAP("ptr2->transfer levels of the first batch allocated");
      if (ok && submit_batch(ppr, ppr2, ppt2, pbs, pbs2, ptr2, dev[d].ctx, mode, kb, &stg[cur], host_threads, local) == _FAILURE_) ok = _FALSE_;
      if (ok) { ++in_flight; cur = (cur + 1) % depth; }
      if (d == 0 && ib == 0) LAP("first batch packed and submitted");
      /* the sources of these k1 levels now live in the staging: free them as the reference does (:553) */
      for (int k1 = kb->k1_first; ok && k1 <= kb->k1_last; ++k1)
        if (perturb2_free_k1_level(ppt2, k1) == _FAILURE_) { snprintf(local, _ERRORMSGSIZE_, "%.1500s", ppt2->error_message); ok = _FALSE_; break; }
    }
    /* drain in submission order: stg[cur] holds the older batch */
    for (int i = 0; i < depth; ++i) {
      los_staging *s = &stg[(cur + i) % depth];
      if (s->kb == NULL) continue;
      if (ok) { if (finish_batch(ppr2, ppt2, ptr2, dev[d].ctx, s, host_threads, local) == _FAILURE_) ok = _FALSE_; }
      else { song_los_wait(dev[d].ctx, NULL); s->kb = NULL; }
    }
    for (int i = 0; i < DEPTH; ++i) staging_free(&stg[i]);
    if (!ok) {
      #pragma omp critical
      { if (abort_flag == _FALSE_) { abort_flag = _TRUE_; snprintf(first_error, _ERRORMSGSIZE_, "%s", local); } }
    }
  }
  LAP("all batches computed and stored");
  for (int d = 0; d < n_dev; ++d) song_los_destroy(dev[d].ctx);
  free(dev); free(batches);
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
