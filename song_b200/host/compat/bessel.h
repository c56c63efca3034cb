/* compat/bessel.h -- stand-in for the fork's include/bessel.h: struct bessels members used by transfer2.c
   (:1092-1098) and bessel2.c (:89, :277-288, :976, :1710), and bessel_j (bessel2.c:1633,1706,1724,1804). */
#ifndef __BESSEL__
#define __BESSEL__
#include "common.h"
struct bessels {
  int l_max;
  int l_size;
  int *l;
  double x_max;
  ErrorMsg error_message;
};
#ifdef __cplusplus
extern "C" {
#endif
int bessel_j(struct bessels *pbs, int l, double x, double *j);
#ifdef __cplusplus
}
#endif
#endif
