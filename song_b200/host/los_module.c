/* los_module.c -- see los_module.h */
#include <stdlib.h>
#include <string.h>
#include "los_module.h"

#include <pthread.h>

static pthread_mutex_t g_pool_lock = PTHREAD_MUTEX_INITIALIZER;   /* callers are OpenMP threads and the disk writer thread */
#define POOL_LOCK()   pthread_mutex_lock(&g_pool_lock)
#define POOL_UNLOCK() pthread_mutex_unlock(&g_pool_lock)

/* ------------------------------------------------------------------------------------------------ */
/*  engines left on the devices by bessel2_init                                                      */
/* ------------------------------------------------------------------------------------------------ */
enum { LOS_MAX_OWNERS = 8 };
static los_engine_set g_sets[LOS_MAX_OWNERS];
static int g_set_used[LOS_MAX_OWNERS];

los_engine_set *los_engines_find(const void *owner)
{
  if (!owner) return NULL;
  for (int i = 0; i < LOS_MAX_OWNERS; ++i)
    if (g_set_used[i] && g_sets[i].owner == owner) return &g_sets[i];
  return NULL;
}

los_engine_set *los_engines_register(const void *owner)
{
  if (!owner || los_engines_find(owner)) return NULL;
  for (int i = 0; i < LOS_MAX_OWNERS; ++i)
    if (!g_set_used[i]) {
      memset(&g_sets[i], 0, sizeof g_sets[i]);
      g_sets[i].owner = owner;
      g_set_used[i] = 1;
      return &g_sets[i];
    }
  return NULL;
}

void los_engines_drop(const void *owner)
{
  los_engine_set *s = los_engines_find(owner);
  if (!s) return;
  for (int d = 0; d < s->n_dev; ++d) if (s->ctx[d]) song_los_destroy(s->ctx[d]);
  g_set_used[s - g_sets] = 0;
}

/* ------------------------------------------------------------------------------------------------ */
/*  page-locked block pool                                                                           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { void *ptr; size_t bytes; int in_use; } los_block;
static los_block *g_blocks;
static int g_n_blocks, g_cap_blocks;

static int pool_enabled(void)
{
  const char *v = getenv("SONG_LOS_PINNED_POOL");
  return !(v && v[0] == '0');
}

void *los_pinned_get(size_t bytes)
{
  void *out = NULL;
  if (bytes == 0) bytes = 1;
  POOL_LOCK();
  {
    /* best fit among the free blocks, as long as less than half of the block would be wasted */
    int best = -1;
    for (int i = 0; i < g_n_blocks; ++i)
      if (!g_blocks[i].in_use && g_blocks[i].bytes >= bytes && g_blocks[i].bytes / 2 <= bytes &&
          (best < 0 || g_blocks[i].bytes < g_blocks[best].bytes)) best = i;
    if (best >= 0) { g_blocks[best].in_use = 1; out = g_blocks[best].ptr; }
  }
  POOL_UNLOCK();
  if (out) return out;
  if (song_los_host_alloc(&out, bytes) != SONG_LOS_SUCCESS) {
    /* page-locked memory is a limited resource: give the pooled blocks back and try once more */
    los_pinned_trim();
    if (song_los_host_alloc(&out, bytes) != SONG_LOS_SUCCESS) return NULL;
  }
  POOL_LOCK();
  {
    if (g_n_blocks == g_cap_blocks) {
      int cap = g_cap_blocks ? 2 * g_cap_blocks : 64;
      los_block *nb = (los_block *)realloc(g_blocks, sizeof(los_block) * (size_t)cap);
      if (nb) { g_blocks = nb; g_cap_blocks = cap; }
    }
    if (g_n_blocks < g_cap_blocks) { g_blocks[g_n_blocks].ptr = out; g_blocks[g_n_blocks].bytes = bytes; g_blocks[g_n_blocks].in_use = 1; ++g_n_blocks; }
  }
  POOL_UNLOCK();
  return out;
}

void los_pinned_put(void *ptr)
{
  if (!ptr) return;
  int release = !pool_enabled(), known = 0;
  POOL_LOCK();
  for (int i = 0; i < g_n_blocks; ++i)
    if (g_blocks[i].ptr == ptr) {
      known = 1;
      if (release) g_blocks[i] = g_blocks[--g_n_blocks]; else g_blocks[i].in_use = 0;
      break;
    }
  POOL_UNLOCK();
  if (release || !known) song_los_host_free(ptr);
}

void los_pinned_trim(void)
{
  POOL_LOCK();
  int w = 0;
  for (int i = 0; i < g_n_blocks; ++i) {
    if (g_blocks[i].in_use) g_blocks[w++] = g_blocks[i];
    else song_los_host_free(g_blocks[i].ptr);
  }
  g_n_blocks = w;
  POOL_UNLOCK();
}

size_t los_pinned_pooled_bytes(void)
{
  size_t n = 0;
  POOL_LOCK();
  for (int i = 0; i < g_n_blocks; ++i) if (!g_blocks[i].in_use) n += g_blocks[i].bytes;
  POOL_UNLOCK();
  return n;
}
