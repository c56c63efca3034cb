/*
 * ref_oracle_providers.c -- variant _ref/libsong_ref_op.so: the two INPUT providers of bessel2.c that live in the absent
 * CLASS fork, bessel_j (class.git source/bessel.c) and drc3jj (tools/slatec_3j_C.c), answered by the oracle's own
 * providers.  Parity is defined "on the same inputs": with the same j_l(x) and 3j numbers on both sides, everything the
 * reference's bessel2.c derives from them (l1 list, lower cuts by bisection, j_l1 tables, J_Llm(x), second derivatives)
 * can be compared with the oracle's restatement bit for bit.  TEST INFRASTRUCTURE ONLY.
 */
#include "common.h"
#include "bessel.h"
#include "slatec_3j_C.h"
#include "../los_oracle.h"

/* The oracle fills its j_l1 table with one downward recurrence per x, started above the LARGEST l1 of the list, and reads
   every l1 off it (oracle_bessel2.c); a recurrence started above l1 itself rounds differently in the last bits.  The test
   announces that largest l1 here so that bessel_j(l, x) returns the very number the oracle's table holds. */
static int g_table_lmax = -1;
void ref_set_bessel_table_lmax(int lmax) { g_table_lmax = lmax; }
void oracle_sph_bessel_all(int lmax, double x, double *out);

int bessel_j(struct bessels *pbs, int l, double x, double *j)
{
  (void)pbs;
  if (g_table_lmax >= l) {
    double *col = (double *)malloc((size_t)(g_table_lmax + 1) * sizeof(double));
    if (col == NULL) return _FAILURE_;
    oracle_sph_bessel_all(g_table_lmax, x, col);
    *j = col[l];
    free(col);
  } else {
    *j = oracle_sph_bessel_j(l, x);
  }
  return _SUCCESS_;
}

int drc3jj(double l2, double l3, double m2, double m3, double *l1min, double *l1max, double *thrcof, int ndim, ErrorMsg errmsg)
{
  if (oracle_drc3jj(l2, l3, m2, m3, l1min, l1max, thrcof, ndim)) {
    snprintf(errmsg, _ERRORMSGSIZE_, "drc3jj: %s", oracle_last_error());
    return _FAILURE_;
  }
  return _SUCCESS_;
}
