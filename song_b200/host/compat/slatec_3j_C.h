/* compat/slatec_3j_C.h -- stand-in for the fork's tools/slatec_3j_C.h (SLATEC DRC3JJ). */
#ifndef __SLATEC_3J_C__
#define __SLATEC_3J_C__
#include "common.h"
#ifdef __cplusplus
extern "C" {
#endif
int drc3jj(double l2, double l3, double m2, double m3, double *l1min, double *l1max,
           double *thrcof, int ndim, ErrorMsg errmsg);
#ifdef __cplusplus
}
#endif
#endif
