/*
 * ref_prelude_bisp.h -- prelude of the variant oracle/_ref/libsong_ref_bisp.so, which adds the reference's own
 * source/bispectra2.c to transfer2.c and bessel2.c so that bispectra2_intrinsic_integrate_over_k3 (bispectra2.c:1107-1439)
 * can be run on the harness structures.  bispectra2.c needs more of the absent CLASS fork than the two path files:
 * the members of struct precision / transfers / perturbs / bispectra it reads anywhere in the file (stand-ins in this
 * directory, shadowing song_b200/host/compat/) and a few constants.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef SONG_REF_PRELUDE_BISP_H
#define SONG_REF_PRELUDE_BISP_H
#define SONG_COMPAT_PRECISION_EXTRA \
  enum interpolation_methods transfers_k1_interpolation; enum interpolation_methods transfers_k2_interpolation; \
  short store_bispectra_to_disk; short compute_only_odd_ls; short compute_only_even_ls;
#include "../ref_prelude.h"
#define _EVEN_ 0
#define _ODD_ 1
#define _SMALL_ 1e-7
#define _MINUSCULE_ 1e-75
#define _PI_CUBE_ (_PI_ * _PI_ * _PI_)
#define _PI_FOURTH_ (_PI_ * _PI_ * _PI_ * _PI_)
#define sqrt_2 1.4142135623730950488016887242097
/* the fork's class_test_permissive prints and goes on */
#define class_test_permissive(condition, error_message_output, ...) { if (condition) { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } }
/* the fork's verbosity-gated printf */
#define printf_log_if(verbose_level, threshold, ...) { if ((verbose_level) > (threshold)) printf(__VA_ARGS__); }
struct spectra; struct lensing; struct primordial; struct bispectra; struct transfers;
#endif
