/*
 * los_glue.h -- the one place where SONG's structures are flattened into the C-ABI's song_los_config
 * (include/song_los.h).  Used by both drop-in modules: bessel2_init creates the GPU engines with it and transfer2_init
 * checks that those engines describe the run it is about to compute (or creates its own when bessel2_init was the
 * reference's).  Compiled against SONG's own headers, like transfer2_gpu.c and bessel2_gpu.c.
 */
#ifndef SONG_LOS_GLUE_H
#define SONG_LOS_GLUE_H
#include "transfer2.h"
#include "song_los.h"

typedef struct los_config_buf {
  song_los_config cfg;                    /* its pointers point into the arrays below */
  int *l, *m, *lm_array, *lm_src, *tt_index_l, *tt_index_m;
} los_config_buf;

/* Everything song_los_create needs, derived from the structures bessel2_init receives (main/song.c:100): the (l,m) and
   transfer-type tables are those of transfer2_indices_of_transfers / transfer2_get_lm_lists (transfer2.c:866-1285).
   tau0 and the time-sampling mode only exist in transfer2_init (pba->conformal_age, ptr2->tau_sampling); bessel2_init
   passes tau0 = 0 and SONG_LOS_TAU_SOURCES, transfer2_init sets the real ones with song_los_set_time. */
int los_config_from_song(struct precision *ppr, struct precision2 *ppr2, struct perturbs2 *ppt2, struct bessels *pbs,
                         struct bessels2 *pbs2, double tau0, int tau_sampling_mode, int device, los_config_buf *out,
                         ErrorMsg errmsg);
void los_config_release(los_config_buf *buf);

/* pbs2->J_Llm_x & co. (bessel2.h:64-91) flattened for song_los_set_projection and handed to the engine */
int los_upload_host_projection(struct precision *ppr, struct precision2 *ppr2, struct bessels *pbs, struct bessels2 *pbs2,
                               song_los_ctx *ctx, ErrorMsg errmsg);
#endif
