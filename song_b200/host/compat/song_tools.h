/* compat/song_tools.h -- stand-in for the fork's tools/song_tools.h: the (l,m) offset helpers the hot path calls
   (source/transfer2.c:900,1154,1207; source/perturbations2.c:854,1159).  Implemented in ../class_compat.c. */
#ifndef __SONG_TOOLS__
#define __SONG_TOOLS__
#include "common.h"
#ifdef __cplusplus
extern "C" {
#endif
int size_l_indexm(int l_max, int *m_vec, int m_size);
int multipole2offset_l_indexm(int l, int m, int *m_vec, int m_size);
int size_indexl_indexm(int *l_vec, int l_size, int *m_vec, int m_size);
int multipole2offset_indexl_indexm(int l, int m, int *l_vec, int l_size, int *m_vec, int m_size);
int offset2multipole_indexl_indexm(int offset, int *l_vec, int l_size, int *m_vec, int m_size,
                                   int *index_l, int *index_m);
int lin_space(double *x, double x_min, double x_max, int n_points);
#ifdef __cplusplus
}
#endif
#endif
