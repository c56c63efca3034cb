/* stand-in for the fork's transfer.h: what transfer2.c and bispectra2.c read of struct transfers */
#ifndef __TRANSFER__
#define __TRANSFER__
#include "common.h"
struct transfers { int dummy; int q_size; double *q; int *tt_size; int *l_size; double **transfer; ErrorMsg error_message; };
#endif
