/*
 * ref_glue.c -- the one CLASS-fork helper source/transfer2.c calls that the drop-in does not need:
 * array_spline_table_columns (class.git tools/arrays.c; used at transfer2.c:2639).  Restated in upstream CLASS's own
 * structure -- all columns advance together, x outer / column inner, workspaces u[(x_size-1)*y_size], p, qn, un[y_size];
 * layout y_array[index_y * x_size + index_x] -- which is NOT how the oracle's restatement (oracle_tools.c, one column at
 * a time) is organised, so that reference-vs-oracle comparisons do not compare a routine with itself; the arithmetic per
 * column is the published one and both must agree bit for bit.  TEST INFRASTRUCTURE ONLY.
 */
#include "common.h"
#include "arrays.h"

int array_spline_table_columns(double *x, int x_size, double *y_array, int y_size, double *ddy_array, short spline_mode,
                               ErrorMsg errmsg)
{
  double *p, *qn, *un, *u;
  double sig, dy_first, dy_last;
  int index_x, index_y;

  class_test(x_size < 3, errmsg, "array_spline_table_columns: x_size=%d too small", x_size);
  u = malloc((size_t)(x_size - 1) * y_size * sizeof(double));
  p = malloc((size_t)y_size * sizeof(double));
  qn = malloc((size_t)y_size * sizeof(double));
  un = malloc((size_t)y_size * sizeof(double));
  if (u == NULL || p == NULL || qn == NULL || un == NULL) {
    free(u); free(p); free(qn); free(un);
    snprintf(errmsg, _ERRORMSGSIZE_, "array_spline_table_columns: cannot allocate the workspaces");
    return _FAILURE_;
  }

  index_x = 0;
  if (spline_mode == _SPLINE_NATURAL_) {
    for (index_y = 0; index_y < y_size; index_y++) {
      ddy_array[index_y * x_size + index_x] = 0.0;
      u[index_x * y_size + index_y] = 0.0;
    }
  } else {
    for (index_y = 0; index_y < y_size; index_y++) {
      dy_first = ((x[2] - x[0]) * (x[2] - x[0]) * (y_array[index_y * x_size + 1] - y_array[index_y * x_size + 0])
                  - (x[1] - x[0]) * (x[1] - x[0]) * (y_array[index_y * x_size + 2] - y_array[index_y * x_size + 0]))
                 / ((x[2] - x[0]) * (x[1] - x[0]) * (x[2] - x[1]));
      ddy_array[index_y * x_size + index_x] = -0.5;
      u[index_x * y_size + index_y] = (3. / (x[1] - x[0]))
        * ((y_array[index_y * x_size + 1] - y_array[index_y * x_size + 0]) / (x[1] - x[0]) - dy_first);
    }
  }

  for (index_x = 1; index_x < x_size - 1; index_x++) {
    sig = (x[index_x] - x[index_x - 1]) / (x[index_x + 1] - x[index_x - 1]);
    for (index_y = 0; index_y < y_size; index_y++) {
      p[index_y] = sig * ddy_array[index_y * x_size + (index_x - 1)] + 2.0;
      ddy_array[index_y * x_size + index_x] = (sig - 1.0) / p[index_y];
      u[index_x * y_size + index_y] =
        (y_array[index_y * x_size + (index_x + 1)] - y_array[index_y * x_size + index_x]) / (x[index_x + 1] - x[index_x])
        - (y_array[index_y * x_size + index_x] - y_array[index_y * x_size + (index_x - 1)]) / (x[index_x] - x[index_x - 1]);
      u[index_x * y_size + index_y] = (6.0 * u[index_x * y_size + index_y] / (x[index_x + 1] - x[index_x - 1])
                                       - sig * u[(index_x - 1) * y_size + index_y]) / p[index_y];
    }
  }

  if (spline_mode == _SPLINE_NATURAL_) {
    for (index_y = 0; index_y < y_size; index_y++) { qn[index_y] = un[index_y] = 0.0; }
  } else {
    for (index_y = 0; index_y < y_size; index_y++) {
      dy_last = ((x[x_size - 3] - x[x_size - 1]) * (x[x_size - 3] - x[x_size - 1])
                   * (y_array[index_y * x_size + (x_size - 2)] - y_array[index_y * x_size + (x_size - 1)])
                 - (x[x_size - 2] - x[x_size - 1]) * (x[x_size - 2] - x[x_size - 1])
                   * (y_array[index_y * x_size + (x_size - 3)] - y_array[index_y * x_size + (x_size - 1)]))
                / ((x[x_size - 3] - x[x_size - 1]) * (x[x_size - 2] - x[x_size - 1]) * (x[x_size - 3] - x[x_size - 2]));
      qn[index_y] = 0.5;
      un[index_y] = (3. / (x[x_size - 1] - x[x_size - 2]))
        * (dy_last - (y_array[index_y * x_size + (x_size - 1)] - y_array[index_y * x_size + (x_size - 2)]) / (x[x_size - 1] - x[x_size - 2]));
    }
  }

  index_x = x_size - 1;
  for (index_y = 0; index_y < y_size; index_y++)
    ddy_array[index_y * x_size + index_x] = (un[index_y] - qn[index_y] * u[(index_x - 1) * y_size + index_y])
                                            / (qn[index_y] * ddy_array[index_y * x_size + (index_x - 1)] + 1.0);

  for (index_x = x_size - 2; index_x >= 0; index_x--)
    for (index_y = 0; index_y < y_size; index_y++)
      ddy_array[index_y * x_size + index_x] = ddy_array[index_y * x_size + index_x] * ddy_array[index_y * x_size + (index_x + 1)]
                                              + u[index_x * y_size + index_y];

  free(qn); free(un); free(p); free(u);
  return _SUCCESS_;
}
