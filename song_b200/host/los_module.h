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

#ifdef __cplusplus
}
#endif
#endif
