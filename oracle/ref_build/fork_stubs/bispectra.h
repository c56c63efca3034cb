/* stand-in for the fork's bispectra.h: the members of struct bispectra that source/bispectra2.c touches anywhere */
#ifndef __BISPECTRA__
#define __BISPECTRA__
#include "common.h"
#include "perturbations.h"
#include "transfer.h"
#include "bessel.h"
#include "primordial.h"
enum bispectra_types { separable_bispectrum, non_separable_bispectrum, analytical_bispectrum, intrinsic_bispectrum };

struct bispectra {
  short has_reduced_bispectrum[_MAX_NUM_BISPECTRA_];
  short has_bispectra, has_intrinsic, has_bispectra_t, has_bispectra_e, has_bispectra_b, add_quadratic_correction;
  int bispectra_verbose;
  int *l; int l_size; int l_max;
  int bf_size, bt_size, n_probes;
  int index_bf_t, index_bf_e, index_bf_b, index_bt_intrinsic, index_bt_quadratic;
  char bf_labels[_MAX_NUM_FIELDS_][_MAX_LENGTH_LABEL_];
  char bfff_labels[_MAX_NUM_FIELDS_][_MAX_NUM_FIELDS_][_MAX_NUM_FIELDS_][_MAX_LENGTH_LABEL_];
  char bt_labels[_MAX_NUM_BISPECTRA_][_MAX_LENGTH_LABEL_];
  int field_parity[_MAX_NUM_FIELDS_];
  int index_tt_of_bf[_MAX_NUM_FIELDS_];
  enum bispectra_types bispectrum_type[_MAX_NUM_BISPECTRA_];
  int **index_l_triangular_min, **index_l_triangular_max, **l_triangular_size;
  long int ***index_l1_l2_l3;
  long int n[_MAX_NUM_BISPECTRA_];
  double *****bispectra;
  double *pk, *pk_pt, *delta_k;
  int k_smooth_size;
  long int count_allocated_for_bispectra, count_memorised_for_bispectra;
  ErrorMsg error_message;
};
#endif
