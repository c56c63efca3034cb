/* compat/arrays.h -- stand-in for the fork's include/arrays.h: the one spline routine bessel2.c calls on single tables
   (source/bessel2.c:417,445).  Implemented in class_symbols.c on top of ../class_compat.c. */
#ifndef __ARRAYS__
#define __ARRAYS__
#include "common.h"
#ifdef __cplusplus
extern "C" {
#endif
int array_spline_table_one_column(double *x, int x_size, double *y_array, int y_size, int index_y, double *ddy_array,
                                  short spline_mode, ErrorMsg errmsg);
#ifdef __cplusplus
}
#endif
#endif
