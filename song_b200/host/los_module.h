/*
 * los_module.h -- state shared by the two drop-in modules (transfer2_gpu.c, bessel2_gpu.c) that SONG's structures have no
 * room for (their layout is the reference's, include/transfer2.h:76-238 and include/bessel2.h:32-112):
 *
 *   - the GPU engines bessel2_init leaves behind, keyed by the struct bessels2 they were built for, so that
 *     transfer2_init finds the projection tables already on the devices (main/song.c:100 -> :106);
 *   - a pool of page-locked host blocks.  Locking memory costs ~0.4 s per GB, which is more than the whole GPU
 *     computation of a 10 % run; blocks are therefore kept when ptr2->transfer levels or staging buffers are freed and
 *     handed out again (SONG_LOS_PINNED_POOL=0 returns them to the system at once).
 */
#ifndef SONG_LOS_MODULE_H
#define SONG_LOS_MODULE_H
#include <stddef.h>
#include "song_los.h"
#ifdef __cplusplus
extern "C" {
#endif

#define LOS_MAX_DEVICES 16

typedef struct los_engine_set {
  const void *owner;                       /* the struct bessels2 (or NULL for a set private to one transfer2_init) */
  int n_dev;
  song_los_ctx *ctx[LOS_MAX_DEVICES];      /* ctx[d] lives on CUDA device d */
} los_engine_set;

los_engine_set *los_engines_find(const void *owner);
los_engine_set *los_engines_register(const void *owner);   /* empty set; fails (NULL) if the owner already has one */
void los_engines_drop(const void *owner);                  /* destroys the contexts and forgets the owner */

/* page-locked blocks (cudaHostAlloc behind song_los_host_alloc); NULL + song_los_last_error(NULL) on failure */
void *los_pinned_get(size_t bytes);
void  los_pinned_put(void *ptr);
void  los_pinned_trim(void);                               /* frees every pooled block */
size_t los_pinned_pooled_bytes(void);

/* ------------------------------------------------------------------------------------------------
 * Backing store of ptr2->transfer (struct transfers2 has no room for it: its layout is the reference's).
 * The reference allocates one array per (transfer type, k1, k2) -- tt2_size * (k1+1) mallocs per k1 level,
 * transfer2.c:2976-3032.  Here the pointers of ALL levels live in one block ([tt2_size][n_pairs], pair = k1(k1+1)/2 + k2
 * in the reference loop order) and the values in a few large page-locked slabs, one per GPU batch, laid out
 * [tt2_size][k3 values of the batch] -- exactly what a batch's device->host copy writes, so results land in
 * ptr2->transfer without a staging copy or a scatter.  A slab is returned to the pool when every k1 level that
 * points into it has been freed (transfer2_free_k1_level keeps its per-level meaning).
 * ------------------------------------------------------------------------------------------------ */
typedef struct los_transfer_store los_transfer_store;

los_transfer_store *los_store_create(const void *owner, int k_size, int tt2_size);   /* NULL: out of memory / owner known */
los_transfer_store *los_store_find(const void *owner);
void los_store_drop(const void *owner);                        /* releases every slab still referenced and the pointer block */
/* &ptr2->transfer[tt][k1][0]: the (k1+1) pointers of one (type, k1) row */
double **los_store_row(los_transfer_store *s, int tt, int k1);
/* a slab of `doubles` values (page-locked when `pinned` and the allocation succeeds, plain memory otherwise); returns its
   id (>= 0) and base address, -1 when out of memory */
int  los_store_new_slab(los_transfer_store *s, size_t doubles, int pinned, double **base);
void los_store_link(los_transfer_store *s, int k1, int slab);   /* level k1 points into `slab` (idempotent per pair of ids) */
void los_store_release_level(los_transfer_store *s, int k1);    /* drops the level's references; frees slabs nobody points into */
size_t los_store_live_bytes(const los_transfer_store *s);

#ifdef __cplusplus
}
#endif
#endif
