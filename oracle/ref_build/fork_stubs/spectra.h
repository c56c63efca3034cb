/* stand-in for the fork's spectra.h: the members of struct spectra that source/spectra2.c touches anywhere */
#ifndef __SPECTRA__
#define __SPECTRA__
#include "common.h"
#include "perturbations.h"
#include "primordial.h"
enum cl_types { first_order, second_order };      /* spectra2.c:299 compares psp->cl_type[index_ct] with second_order */
enum linear_or_logarithmic { linear, logarithmic }; /* third argument of primordial_spectrum_at_k */
#ifndef _MAX_NUM_CT_
#define _MAX_NUM_CT_ 32
#endif
struct spectra {
  int spectra_verbose;
  int ct_size; enum cl_types cl_type[_MAX_NUM_CT_];
  char ct_labels[_MAX_NUM_CT_][_MAX_LENGTH_LABEL_];
  short has_tt2, has_te2, has_ee2, has_bb2, compute_cl_derivative;
  int index_ct_tt2, index_ct_te2, index_ct_ee2, index_ct_bb2;
  int k_size; double *k;
  int l_size_song; int *l_song;
  int *l_size; double *l; int **l_max_ct;
  double **cl, **ddcl, **d_lsq_cl, **dd_lsq_cl, **spline_d_lsq_cl;
  ErrorMsg error_message;
};
#endif
