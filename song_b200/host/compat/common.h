/*
 * compat/common.h -- stand-in for the CLASS fork's include/common.h (coccoinomane/class_public @ 2643d443,
 * branch song_support), which is an empty submodule in the reference checkout.
 *
 * It declares ONLY what SONG's own headers (include/transfer2.h, bessel2.h, perturbations2.h, common2.h) and the
 * drop-in transfer2/bessel2 modules touch: error macros, constants, and the fields of `struct precision` that
 * source/transfer2.c and source/bessel2.c read.  It sits LAST on the include path: in a real
 * SONG tree the fork's own common.h is found first and this file is never seen.  Values marked "recalled" come
 * from upstream CLASS and must be checked against the fork when it is available.
 */
#ifndef __COMMON__
#define __COMMON__
#define SONG_B200_COMPAT_HEADERS 1

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define _TRUE_ 1
#define _FALSE_ 0
#define _SUCCESS_ 0          /* recalled */
#define _FAILURE_ 1          /* recalled */
#define _ERRORMSGSIZE_ 2048  /* recalled */
typedef char ErrorMsg[_ERRORMSGSIZE_];
#define _FILENAMESIZE_ 256   /* recalled */
typedef char FileName[_FILENAMESIZE_];
#define _MAX_LENGTH_LABEL_ 64
#define _MAX_INFO_SIZE_ 8192
#define _MAX_NUM_FIELDS_ 8
#define _MAX_NUM_BISPECTRA_ 32
#define _PI_ 3.1415926535897932384626433832795e0
#define _EPS_ 1.e-10         /* recalled */
#define _SPLINE_NATURAL_ 0   /* recalled */
#define _SPLINE_EST_DERIV_ 1 /* recalled */
#define _LINE_LENGTH_MAX_ 1024
#define _ARGUMENT_LENGTH_MAX_ 1024

#ifndef MIN
#define MIN(a,b) (((a)<(b)) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a,b) (((a)<(b)) ? (b) : (a))
#endif
#define ALTERNATING_SIGN(m) ((m)%2 == 0 ? 1 : -1)

/* error macros: same observable behaviour as CLASS (message "Error in f; => inner message", return _FAILURE_) */
#define class_call(function, error_message_from_function, error_message_output) do {                        \
  if ((function) == _FAILURE_) {                                                                            \
    ErrorMsg class_tmp_;                                                                                    \
    snprintf(class_tmp_, _ERRORMSGSIZE_, "%s(L:%d) :error in %s;\n=>%s", __func__, __LINE__, #function,     \
             error_message_from_function);                                                                  \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s", class_tmp_);                                       \
    return _FAILURE_; } } while (0)
#define class_test(condition, error_message_output, ...) do {                                               \
  if (condition) {                                                                                          \
    ErrorMsg class_tmp1_, class_tmp2_;                                                                      \
    snprintf(class_tmp1_, _ERRORMSGSIZE_, __VA_ARGS__);                                                     \
    snprintf(class_tmp2_, _ERRORMSGSIZE_, "%s(L:%d) :condition (%s) is true; %.1500s", __func__, __LINE__,  \
             #condition, class_tmp1_);                                                                      \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s", class_tmp2_);                                      \
    return _FAILURE_; } } while (0)
#define class_stop(error_message_output, ...) class_test(_TRUE_, error_message_output, __VA_ARGS__)
#define class_alloc(pointer, size, error_message_output) do {                                               \
  pointer = malloc(size);                                                                                   \
  if (pointer == NULL) {                                                                                    \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s(L:%d) :could not allocate %s with size %zu",         \
             __func__, __LINE__, #pointer, (size_t)(size));                                                 \
    return _FAILURE_; } } while (0)
#define class_calloc(pointer, number, size, error_message_output) do {                                      \
  pointer = calloc(number, size);                                                                           \
  if (pointer == NULL) {                                                                                    \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s(L:%d) :could not allocate %s with size %zu",         \
             __func__, __LINE__, #pointer, (size_t)(number) * (size_t)(size));                              \
    return _FAILURE_; } } while (0)
#define class_open(pointer, filename, mode, error_output) do {                                              \
  pointer = fopen(filename, mode);                                                                          \
  if (pointer == NULL) {                                                                                    \
    snprintf(error_output, _ERRORMSGSIZE_, "%s(L:%d) :could not open %s with name %s and mode %s",          \
             __func__, __LINE__, #pointer, filename, #mode);                                                \
    return _FAILURE_; } } while (0)

/* enums of the fork's common.h used by the hot path (names from source/transfer2.c, source/bessel2.c) */
enum interpolation_methods { linear_interpolation, cubic_interpolation,
                             sum_linear_interpolation, sum_cubic_interpolation, smart_interpolation };
enum bispectra_k3_extrapolation_methods { no_k3_extrapolation, flat_k3_extrapolation };  /* tag name assumed */
enum file_format { class_format, camb_format };
enum evolver_type { rk, ndf15 };
enum pk_def { delta_m_squared, delta_tot_squared, delta_bc_squared, delta_tot_from_poisson_squared };

/* struct precision: only the members the transfer2 / bessel2 modules read */
struct precision {
  enum bispectra_k3_extrapolation_methods bispectra_k3_extrapolation; /* transfer2.c:1365,1471,1539,1624,2740 */
  double extra_k3_oscillations_left;      /* transfer2.c:1372-1393 (<0: extend to the global k range) */
  double extra_k3_oscillations_right;
  double smallest_allowed_variation;      /* transfer2.c:1482,1634; bessel2.c:1719 */
  short load_run;                         /* transfer2.c:1434,1602 */
  enum interpolation_methods bessels_interpolation;   /* transfer2.c:2259-2262; bessel2.c:194,261 */
  double k_max_tau0_over_l_max;           /* transfer2.c:2175 (debug message only) */
  double bessel_tol_x_min;                /* bessel2.c:1717 */
  short load_bispectra_from_disk;         /* bessel2.c:90 */
#ifdef SONG_COMPAT_PRECISION_EXTRA
  SONG_COMPAT_PRECISION_EXTRA             /* test builds that compile further files of the reference add the members those read */
#endif
  ErrorMsg error_message;
};
#endif
