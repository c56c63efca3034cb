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

/* ------------------------------------------------------------------------------------------------ */
/*  backing store of ptr2->transfer                                                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double *base; size_t doubles; int refs; int pinned; } los_slab;
typedef struct { int n, cap; int *ids; } los_level;

struct los_transfer_store {
  const void *owner;
  int k_size, tt2_size;
  long n_pairs;
  double **ptr_block;            /* [tt2_size][n_pairs] */
  los_slab *slabs; int n_slabs, cap_slabs;
  los_level *levels;             /* [k_size] */
  pthread_mutex_t lock;
};

enum { LOS_MAX_STORES = 8 };
static los_transfer_store *g_stores[LOS_MAX_STORES];

los_transfer_store *los_store_find(const void *owner)
{
  if (!owner) return NULL;
  for (int i = 0; i < LOS_MAX_STORES; ++i) if (g_stores[i] && g_stores[i]->owner == owner) return g_stores[i];
  return NULL;
}

los_transfer_store *los_store_create(const void *owner, int k_size, int tt2_size)
{
  if (!owner || k_size < 1 || tt2_size < 1 || los_store_find(owner)) return NULL;
  int at = -1;
  for (int i = 0; i < LOS_MAX_STORES; ++i) if (!g_stores[i]) { at = i; break; }
  if (at < 0) return NULL;
  los_transfer_store *s = (los_transfer_store *)calloc(1, sizeof *s);
  if (!s) return NULL;
  s->owner = owner; s->k_size = k_size; s->tt2_size = tt2_size;
  s->n_pairs = (long)k_size * (k_size + 1) / 2;
  s->ptr_block = (double **)calloc((size_t)tt2_size * (size_t)s->n_pairs, sizeof(double *));   /* untouched pages cost nothing */
  s->levels = (los_level *)calloc((size_t)k_size, sizeof(los_level));
  if (!s->ptr_block || !s->levels) { free(s->ptr_block); free(s->levels); free(s); return NULL; }
  pthread_mutex_init(&s->lock, NULL);
  g_stores[at] = s;
  return s;
}

double **los_store_row(los_transfer_store *s, int tt, int k1)
{
  return s->ptr_block + (size_t)tt * (size_t)s->n_pairs + (size_t)k1 * (k1 + 1) / 2;
}

int los_store_new_slab(los_transfer_store *s, size_t doubles, int pinned, double **base)
{
  double *p = NULL;
  int is_pinned = 0;
  const size_t bytes = (doubles > 0 ? doubles : 1) * sizeof(double);
  if (pinned) { p = (double *)los_pinned_get(bytes); is_pinned = (p != NULL); }
  if (!p) p = (double *)malloc(bytes);
  if (!p) return -1;
  pthread_mutex_lock(&s->lock);
  int id = -1;
  for (int i = 0; i < s->n_slabs; ++i) if (!s->slabs[i].base) { id = i; break; }     /* reuse a vacated entry */
  if (id < 0) {
    if (s->n_slabs == s->cap_slabs) {
      int cap = s->cap_slabs ? 2 * s->cap_slabs : 32;
      los_slab *ns = (los_slab *)realloc(s->slabs, sizeof(los_slab) * (size_t)cap);
      if (ns) { s->slabs = ns; s->cap_slabs = cap; }
    }
    if (s->n_slabs < s->cap_slabs) id = s->n_slabs++;
  }
  if (id >= 0) { s->slabs[id].base = p; s->slabs[id].doubles = doubles; s->slabs[id].refs = 0; s->slabs[id].pinned = is_pinned; }
  pthread_mutex_unlock(&s->lock);
  if (id < 0) { if (is_pinned) los_pinned_put(p); else free(p); return -1; }
  *base = p;
  return id;
}

void los_store_link(los_transfer_store *s, int k1, int slab)
{
  pthread_mutex_lock(&s->lock);
  los_level *L = &s->levels[k1];
  int known = 0;
  for (int i = 0; i < L->n; ++i) if (L->ids[i] == slab) known = 1;
  if (!known) {
    if (L->n == L->cap) {
      int cap = L->cap ? 2 * L->cap : 4;
      int *ni = (int *)realloc(L->ids, sizeof(int) * (size_t)cap);
      if (ni) { L->ids = ni; L->cap = cap; }
    }
    if (L->n < L->cap) { L->ids[L->n++] = slab; s->slabs[slab].refs++; }
  }
  pthread_mutex_unlock(&s->lock);
}

static void slab_unref(los_transfer_store *s, int id)
{
  los_slab *b = &s->slabs[id];
  if (--b->refs > 0) return;
  if (b->pinned) los_pinned_put(b->base); else free(b->base);
  b->base = NULL; b->doubles = 0; b->refs = 0;
}

void los_store_release_level(los_transfer_store *s, int k1)
{
  pthread_mutex_lock(&s->lock);
  los_level *L = &s->levels[k1];
  for (int i = 0; i < L->n; ++i) slab_unref(s, L->ids[i]);
  L->n = 0;
  pthread_mutex_unlock(&s->lock);
}

size_t los_store_live_bytes(const los_transfer_store *s)
{
  size_t n = 0;
  for (int i = 0; i < s->n_slabs; ++i) if (s->slabs[i].base) n += s->slabs[i].doubles * sizeof(double);
  return n;
}

void los_store_drop(const void *owner)
{
  los_transfer_store *s = los_store_find(owner);
  if (!s) return;
  for (int k1 = 0; k1 < s->k_size; ++k1) { los_store_release_level(s, k1); free(s->levels[k1].ids); }
  for (int i = 0; i < s->n_slabs; ++i)          /* slabs never linked to a level */
    if (s->slabs[i].base) { if (s->slabs[i].pinned) los_pinned_put(s->slabs[i].base); else free(s->slabs[i].base); }
  for (int i = 0; i < LOS_MAX_STORES; ++i) if (g_stores[i] == s) g_stores[i] = NULL;
  pthread_mutex_destroy(&s->lock);
  free(s->slabs); free(s->levels); free(s->ptr_block); free(s);
}
