/*
 * bessel2_host.h -- host half of the bessel2 drop-in: everything bessel2_init (reference
 * source/bessel2.c:75-468) computes on the host before the J_Llm assembly, which runs on the GPU
 * through song_los_build_projection.
 */
#ifndef SONG_BESSEL2_HOST_H
#define SONG_BESSEL2_HOST_H
#include "../../include/song_los.h"
#ifdef __cplusplus
extern "C" {
#endif

/* the j_l1 members of struct bessels2 (include/bessel2.h:96-106) plus the 3j symbols of every
   (J,L,l,m) table, flattened with jidx = ((J*L_size + L)*l_size + il)*m_size + im */
typedef struct song_bessel2_host {
  int l1_size; int *l1; int *index_l1; int index_l1_len;
  double **j_l1; int *index_xmin_l1; int *x_size_l1; double *x_min_l1;
  int n_jidx, ndim;
  double *first_3j, *second_3j;            /* [n_jidx][ndim] */
  int *l1_min, *l1_count, *index_l1_of_l1min;
  long count_allocated_Js;
} song_bessel2_host;

/* bessel2_get_l1_list (:832-943), bessel2_j_for_l1 (:1669-1815) for every l1, and the drc3jj
   calls of bessel2_J_for_Llm (:1131,:1154).  J_types[index_J] is SONG_LOS_J_TT/EE/EB. */
int song_host_bessel2_prepare(const int *l, int l_size, const int *m, int m_size, int L_max,
                              const double *xx, int xx_size, double xx_max,
                              int J_size, const int *J_types,
                              double j_l1_cut, double bessel_tol_x_min, double smallest_allowed_variation,
                              song_bessel2_host **out, char *errmsg, int errlen);
void song_host_bessel2_free(song_bessel2_host *t);
/* The j_l(x) provider of song_host_bessel2_prepare.  NULL (default): the recurrence of class_compat.c, where ONE downward
   recurrence per x serves every l1 of the table.  In a SONG tree bessel2_gpu.c registers CLASS's own bessel_j, which is
   then called point by point exactly as the reference's bessel2_j_for_l1 does (bessel2.c:1706,1724,1804), so that
   pbs2->j_l1 holds the reference's numbers bit for bit (a recurrence started above the largest l1 and one started above
   l1 itself agree only to the last 1-3 bits).  SONG_LOS_JL1_PER_POINT=1 selects the point-by-point pattern with the
   default provider (tests/test_reference_pin.py compares it with the reference's own table). */
typedef double (*song_host_bessel_fn)(int l, double x, void *user);
void song_host_bessel2_set_provider(song_host_bessel_fn fn, void *user);
/* Likewise the Wigner 3j provider (bessel2.c:1131,1154): NULL = the recursion of class_compat.c; a SONG tree registers
   the fork's drc3jj (tools/slatec_3j_C.c).  Returns 0 on success; thrcof has ndim entries. */
typedef int (*song_host_3j_fn)(double l2, double l3, double m2, double m3, double *l1min, double *l1max, double *thrcof, int ndim, void *user);
void song_host_bessel2_set_3j_provider(song_host_3j_fn fn, void *user);
/* hands the tables to the engine: J_Llm(x) is assembled on the device */
int song_host_bessel2_upload(song_los_ctx *ctx, const song_bessel2_host *t, const int *J_types,
                             double J_Llm_cut, double xx_max, char *errmsg, int errlen);
#ifdef __cplusplus
}
#endif
#endif
