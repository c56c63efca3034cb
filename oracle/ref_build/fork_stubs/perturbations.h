/* stand-in for the fork's perturbations.h (superset of song_b200/host/compat/perturbations.h) */
#ifndef __PERTURBATIONS__
#define __PERTURBATIONS__
#include "common.h"
struct background { double conformal_age; ErrorMsg error_message; };
struct thermo { double tau_rec; double rs_rec; ErrorMsg error_message; };
struct perturbs {
  short has_cls;
  short has_cmb_bispectra;
  int index_md_scalars; int index_ic_ad;
  short has_scalars; int md_size; int *ic_size;
  ErrorMsg error_message;
};
struct perturb_vector { int pt_size; double *y; double *dy; int *used_in_sources; };
struct perturb_workspace { int dummy; };
enum tca_flags { tca_on, tca_off };
enum rsa_flags { rsa_off, rsa_on };
enum possible_gauges { newtonian, synchronous };
enum phi_equation { poisson, longitudinal, huang };
enum sources_tau_samplings { lin_tau_sampling, log_tau_sampling, class_tau_sampling };
#endif
