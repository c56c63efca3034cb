/*
 * compat/class_symbols.c -- providers of the CLASS-fork symbols the transfer2 / bessel2 drop-in calls, for builds
 * WITHOUT the fork (this repository's tests).  In a real SONG tree these come from class.git (tools/song_tools.c,
 * tools/slatec_3j_C.c, tools/arrays.c, source/bessel.c) and this file is not compiled.
 * Ordering of the (l,m) offsets: l-major, m-minor, only m <= l -- inferred from the loop nests at
 * source/transfer2.c:1141-1166 and source/perturbations2.c:1151-1161 (unambiguous for m = {0}).
 */
#include "common.h"
#include "song_tools.h"
#include "slatec_3j_C.h"
#include "bessel.h"
#include "arrays.h"
#include "../class_compat.h"

int size_l_indexm(int l_max, int *m_vec, int m_size)
{
  int n = 0;
  for (int l = 0; l <= l_max; ++l)
    for (int im = 0; im < m_size; ++im) if (m_vec[im] <= l) ++n;
  return n;
}
int multipole2offset_l_indexm(int L, int M, int *m_vec, int m_size)
{
  int off = 0;
  for (int l = 0; l <= L; ++l)
    for (int im = 0; im < m_size; ++im) {
      if (m_vec[im] > l) continue;
      if (l == L && m_vec[im] == M) return off;
      ++off;
    }
  return -1;
}
int size_indexl_indexm(int *l_vec, int l_size, int *m_vec, int m_size)
{
  int n = 0;
  for (int il = 0; il < l_size; ++il)
    for (int im = 0; im < m_size; ++im) if (m_vec[im] <= l_vec[il]) ++n;
  return n;
}
int multipole2offset_indexl_indexm(int L, int M, int *l_vec, int l_size, int *m_vec, int m_size)
{
  int off = 0;
  for (int il = 0; il < l_size; ++il)
    for (int im = 0; im < m_size; ++im) {
      if (m_vec[im] > l_vec[il]) continue;
      if (l_vec[il] == L && m_vec[im] == M) return off;
      ++off;
    }
  return -1;
}
int offset2multipole_indexl_indexm(int offset, int *l_vec, int l_size, int *m_vec, int m_size, int *index_l, int *index_m)
{
  int off = 0;
  for (int il = 0; il < l_size; ++il)
    for (int im = 0; im < m_size; ++im) {
      if (m_vec[im] > l_vec[il]) continue;
      if (off == offset) { *index_l = il; *index_m = im; return _SUCCESS_; }
      ++off;
    }
  *index_l = -1; *index_m = -1;
  return _FAILURE_;
}
int lin_space(double *x, double x_min, double x_max, int n_points)
{
  compat_lin_space(x, x_min, x_max, n_points);
  return _SUCCESS_;
}
int drc3jj(double l2, double l3, double m2, double m3, double *l1min, double *l1max, double *thrcof, int ndim, ErrorMsg errmsg)
{
  if (compat_drc3jj(l2, l3, m2, m3, l1min, l1max, thrcof, ndim)) {
    snprintf(errmsg, _ERRORMSGSIZE_, "drc3jj: invalid arguments or ndim=%d too small", ndim);
    return _FAILURE_;
  }
  return _SUCCESS_;
}
int bessel_j(struct bessels *pbs, int l, double x, double *j)
{
  (void)pbs;
  *j = compat_bessel_j(l, x);
  return _SUCCESS_;
}
int array_spline_table_one_column(double *x, int x_size, double *y_array, int y_size, int index_y, double *ddy_array,
                                  short spline_mode, ErrorMsg errmsg)
{
  (void)y_size;
  if (x_size < 3) {          /* upstream has no such test (it would read x[2]); tables this short are never interpolated in */
    for (int i = 0; i < x_size; ++i) ddy_array[(size_t)index_y * x_size + i] = 0.;
    return _SUCCESS_;
  }
  if (compat_spline_second_derivatives(x, x_size, y_array + (size_t)index_y * x_size, ddy_array + (size_t)index_y * x_size,
                                       spline_mode == _SPLINE_EST_DERIV_)) {
    snprintf(errmsg, _ERRORMSGSIZE_, "array_spline_table_one_column: x_size=%d too small or repeated abscissa", x_size);
    return _FAILURE_;
  }
  return _SUCCESS_;
}
