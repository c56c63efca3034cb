/*
 * ref_prelude.h -- force-included (gcc -include) when the REFERENCE's own source/transfer2.c and source/bessel2.c are
 * compiled, where they lie under /root/reference, into oracle/_ref/libsong_ref.so (recipe: oracle/Makefile, target `ref`).
 * TEST INFRASTRUCTURE ONLY: the library pins the oracle (tests/test_reference_pin.py) and is the CPU arm of bench.py.
 *
 * The reference needs its CLASS fork (class.git, an empty submodule in the checkout).  What stands in for it:
 *   - the declarations of song_b200/host/compat/*.h (struct members and enums the two modules read);
 *   - the macros below, written as plain blocks like CLASS's own (the reference omits the semicolon after some
 *     class_test(...) calls, transfer2.c:1538) plus the *_parallel variants used inside OpenMP regions (transfer2.c:219, 362);
 *   - restatements of the fork's helpers the two files call: array_spline_table_columns / _one_column, lin_space, drc3jj,
 *     bessel_j, *_indexl_indexm (song_b200/host/compat/class_symbols.c, song_b200/host/class_compat.c, ref_glue.c).
 * Everything between transfer2_init and transfer2_integrate, and between bessel2_init and bessel2_J_Llm, is the
 * reference's own code, unmodified.
 */
#ifndef SONG_REF_PRELUDE_H
#define SONG_REF_PRELUDE_H
#include "common.h"
#undef class_call
#undef class_test
#undef class_stop
#undef class_alloc
#undef class_calloc
#undef class_open
#define class_call(function, error_message_from_function, error_message_output) {                        \
  if ((function) == _FAILURE_) {                                                                            \
    ErrorMsg class_tmp_;                                                                                    \
    snprintf(class_tmp_, _ERRORMSGSIZE_, "%s(L:%d) :error in %s;\n=>%s", __func__, __LINE__, #function,     \
             error_message_from_function);                                                                  \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s", class_tmp_);                                       \
    return _FAILURE_; } }
#define class_test(condition, error_message_output, ...) {                                               \
  if (condition) {                                                                                          \
    ErrorMsg class_tmp1_, class_tmp2_;                                                                      \
    snprintf(class_tmp1_, _ERRORMSGSIZE_, __VA_ARGS__);                                                     \
    snprintf(class_tmp2_, _ERRORMSGSIZE_, "%s(L:%d) :condition (%s) is true; %.1500s", __func__, __LINE__,  \
             #condition, class_tmp1_);                                                                      \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s", class_tmp2_);                                      \
    return _FAILURE_; } }
#define class_stop(error_message_output, ...) class_test(_TRUE_, error_message_output, __VA_ARGS__)
#define class_alloc(pointer, size, error_message_output) {                                               \
  pointer = malloc(size);                                                                                   \
  if (pointer == NULL) {                                                                                    \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s(L:%d) :could not allocate %s", __func__, __LINE__, #pointer); \
    return _FAILURE_; } }
#define class_calloc(pointer, number, size, error_message_output) {                                      \
  pointer = calloc(number, size);                                                                           \
  if (pointer == NULL) {                                                                                    \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s(L:%d) :could not allocate %s", __func__, __LINE__, #pointer); \
    return _FAILURE_; } }
#define class_open(pointer, filename, mode, error_output) {                                              \
  pointer = fopen(filename, mode);                                                                          \
  if (pointer == NULL) {                                                                                    \
    snprintf(error_output, _ERRORMSGSIZE_, "%s(L:%d) :could not open %s", __func__, __LINE__, #pointer);    \
    return _FAILURE_; } }
/* parallel-region variants of CLASS: set abort instead of returning */
#define class_call_parallel(function, error_message_from_function, error_message_output) {               \
  if (abort == _FALSE_) { if ((function) == _FAILURE_) {                                                    \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s(L:%d) :error in %s;\n=>%s", __func__, __LINE__, #function, error_message_from_function); \
    abort = _TRUE_; } } }
#define class_test_parallel(condition, error_message_output, ...) {                                      \
  if (abort == _FALSE_) { if (condition) {                                                                  \
    ErrorMsg class_tmp1_; snprintf(class_tmp1_, _ERRORMSGSIZE_, __VA_ARGS__);                               \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s(L:%d) :condition (%s) is true; %.1500s", __func__, __LINE__, #condition, class_tmp1_); \
    abort = _TRUE_; } } }
#define class_alloc_parallel(pointer, size, error_message_output) {                                      \
  pointer = NULL; if (abort == _FALSE_) { pointer = malloc(size); if (pointer == NULL) {                     \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s(L:%d) :could not allocate %s", __func__, __LINE__, #pointer); \
    abort = _TRUE_; } } }
#define class_calloc_parallel(pointer, number, size, error_message_output) {                             \
  pointer = NULL; if (abort == _FALSE_) { pointer = calloc(number, size); if (pointer == NULL) {             \
    snprintf(error_message_output, _ERRORMSGSIZE_, "%s(L:%d) :could not allocate %s", __func__, __LINE__, #pointer); \
    abort = _TRUE_; } } }

#endif
