/* stand-in for the fork's primordial.h */
#ifndef __PRIMORDIAL__
#define __PRIMORDIAL__
#include "common.h"
struct primordial { double A_s; ErrorMsg error_message; };
int primordial_spectrum_at_k(struct primordial *ppm, int index_md, int mode, double k, double *pk);   /* the fork's primordial.c */
#endif
