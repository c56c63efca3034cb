/* compat/perturbations.h -- stand-in for CLASS include/perturbations.h: the members of struct background, thermo
   and perturbs that source/transfer2.c touches (transfer2.c:128-130, 614-615), plus what SONG's perturbations2.h
   needs to parse. */
#ifndef __PERTURBATIONS__
#define __PERTURBATIONS__
#include "common.h"
struct background { double conformal_age; ErrorMsg error_message; };
struct thermo { double tau_rec; double rs_rec; ErrorMsg error_message; };
struct perturbs {
  short has_cls;                 /* cleared by transfer2_init when stop_at_transfers2 (transfer2.c:614) */
  short has_cmb_bispectra;       /* transfer2.c:615 */
  ErrorMsg error_message;
};
struct perturb_vector { int pt_size; double *y; double *dy; int *used_in_sources; };   /* opaque to the hot path */
struct perturb_workspace { int dummy; };
enum tca_flags { tca_on, tca_off };
enum rsa_flags { rsa_off, rsa_on };
enum possible_gauges { newtonian, synchronous };
enum phi_equation { poisson, longitudinal, huang };                      /* input2.c:882-896 */
enum sources_tau_samplings { lin_tau_sampling, log_tau_sampling, class_tau_sampling };   /* input2.c:350-353 */
#endif
