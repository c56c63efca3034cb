/*
 * los_oracle.h -- CPU oracle for SONG's second-order line-of-sight (LOS) transfer stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may build,
 * load or call it, and only as the checker / the CPU baseline.
 *
 * PARITY STATUS.  The reference (coccoinomane/song @ 2ea44ecf) ships no golden vectors or known-answer tests for
 * this path (test/print_transfers2.c only prints), and as a whole it cannot be compiled here because its CLASS fork
 * (class.git @ 2643d443) is an empty submodule.  Its two source files of THIS path, source/transfer2.c and
 * source/bessel2.c, do compile where they lie against stand-ins for the fork's declarations and the handful of fork
 * helpers they call (oracle/Makefile target `ref`, oracle/ref_build/): oracle/_ref/libsong_ref.so runs the reference's
 * own transfer2_init, bessel2_init and bessel2_convolution on the test harness's structures, and
 * tests/test_reference_pin.py finds this oracle's restatement of transfer2.c and bessel2.c BIT-IDENTICAL to them over
 * the options of the path and on BASELINE configs[0] at full size.  A third build adds the reference's own
 * source/bispectra2.c and source/spectra2.c (more stand-ins: oracle/ref_build/fork_stubs/): their
 * bispectra2_intrinsic_integrate_over_k3 and spectra2_cls are bit-identical to oracle_bispectra2.c / oracle_spectra2.c.
 * Still unpinned by reference execution: the helpers that live in the missing fork themselves
 * (array_spline_table_columns, drc3jj, lin_space, *_indexl_indexm, bessel_j: restated from the published upstream
 * CLASS / SLATEC algorithms, flagged "UPSTREAM-RECALLED"; bessel_j is a Miller recurrence, not CLASS's approximation --
 * parity is defined on the same j_l(x) inputs).  Those are pinned by analytic known-answer tests only
 * (tests/test_oracle_kat.py).
 *
 * Floating-point contract: compile with -ffp-contract=off so every statement is evaluated
 * in the C left-to-right order written in the reference (no FMA contraction).
 */
#ifndef LOS_ORACLE_H
#define LOS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_OK = 0, ORACLE_FAIL = 1 };

/* enum mirrors (values are the oracle's own; the reference uses C enums in
   include/transfer2.h:18-32 and the CLASS fork's common.h) */
enum { ORACLE_K3_BESSEL_SAMPLING = 0, ORACLE_K3_CLASS_SAMPLING = 1 };       /* transfer2.h:18-21 */
enum { ORACLE_TAU_SOURCES = 0, ORACLE_TAU_CUSTOM = 1, ORACLE_TAU_BESSEL = 2 }; /* transfer2.h:27-32 */
enum { ORACLE_INTERP_LINEAR = 0, ORACLE_INTERP_CUBIC = 1 };
enum { ORACLE_NO_K3_EXTRAPOLATION = 0, ORACLE_FLAT_K3_EXTRAPOLATION = 1 };
enum { ORACLE_SPLINE_NATURAL = 0, ORACLE_SPLINE_EST_DERIV = 1 };
enum { ORACLE_J_TT = 0, ORACLE_J_EE = 1, ORACLE_J_EB = 2 };                 /* bessel2.h:12-16 */

/*
 * Everything transfer2.c reads from ppr, ppr2, pba, pth, ppt2, pbs, pbs2 and everything
 * transfer2_indices_of_transfers() derives, as one flat read-only record.
 * Flattening rule for the ragged [J][L][l][m] tables of bessel2.h:64-91:
 *     jidx = ((index_J*L_size + index_L)*l_size + index_l)*m_size + index_m
 * with a NULL row for (L,l,m) combinations the reference does not allocate (|m| > min(L,l)).
 */
typedef struct oracle_setup {
  /* cosmology (transfer2.c:128-130) */
  double tau0, tau_rec, rs_rec;

  /* k1,k2 sampling (ppt2->k) and source time sampling (ppt2->tau_sampling) */
  int k_size;            const double *k;
  int tau_size;          const double *tau_sampling;

  /* precision knobs */
  int    k3_extrapolation;              /* ppr->bispectra_k3_extrapolation            */
  double extra_k3_oscillations_left;    /* ppr->extra_k3_oscillations_left            */
  double extra_k3_oscillations_right;   /* ppr->extra_k3_oscillations_right           */
  int    k_sampling;                    /* ptr2->k_sampling                           */
  double q_linstep_song;                /* ppr2->q_linstep_song                       */
  int    old_run_step;                  /* (ppr->load_run && ppr2->old_run)           */
  double smallest_allowed_variation;    /* ppr->smallest_allowed_variation            */
  int    tau_sampling_mode;             /* ptr2->tau_sampling                         */
  double tau_linstep_song;              /* ppr2->tau_linstep_song                     */
  int    sources_k3_interpolation;      /* ppr2->sources_k3_interpolation             */
  int    sources_time_interpolation;    /* ppr2->sources_time_interpolation           */
  int    bessels_interpolation;         /* ppr->bessels_interpolation                 */

  /* projection functions (struct bessels2) */
  int xx_size;           const double *xx;
  double xx_step, xx_max;
  int J_size, index_J_TT, index_J_EE, index_J_EB;   /* -1 when absent */
  int L_size;                                       /* pbs2->L[i] == i  (bessel2.c:123-126) */
  int l_size;            const int *l;              /* ptr2->l == pbs->l (transfer2.c:1092-1098) */
  int m_size;            const int *m;
  const double *const *J_rows;                      /* pbs2->J_Llm_x, flattened, see above */
  const double *const *ddJ_rows;                    /* pbs2->ddJ_Llm_x or NULL             */
  const int *index_xmin_J;                          /* pbs2->index_xmin_J flattened        */
  const int *x_size_J;                              /* pbs2->x_size_J flattened            */

  /* source types (struct perturbs2) */
  int has_source_T, has_source_E, has_source_B;
  int index_tp2_T, index_tp2_E, index_tp2_B;
  int tp2_size;
  int l_max_los_t, l_max_los_p;
  const int *lm_src;        /* lm(L,m) of perturbations2_macros.h:53 as [L_size][m_size], -1 if m>L */

  /* transfer types (struct transfers2, filled by oracle_indices_of_transfers) */
  int n_transfers, tt2_size;
  int index_tt2_T, index_tt2_E, index_tt2_B;
  const int *tt_index_l;    /* ptr2->corresponding_index_l [tt2_size] */
  const int *tt_index_m;    /* ptr2->corresponding_index_m [tt2_size] */
} oracle_setup;

/* ---------------------------------------------------------------------------------------
 * UPSTREAM-RECALLED helpers (CLASS fork tools/arrays.c, tools/song_tools.c, slatec_3j_C.c)
 * ------------------------------------------------------------------------------------- */
int  oracle_spline_table_columns(const double *x, int x_size, const double *y, int y_size,
                                 double *ddy, int mode);
int  oracle_spline_table_one_column(const double *x, int x_size, const double *y, int y_size,
                                    int index_y, double *ddy, int mode);
void oracle_lin_space(double *x, double x_min, double x_max, int n);
int  oracle_size_l_indexm(int l_max, const int *m, int m_size);
int  oracle_multipole2offset_l_indexm(int l, int mm, const int *m, int m_size);
int  oracle_size_indexl_indexm(const int *l, int l_size, const int *m, int m_size);
int  oracle_multipole2offset_indexl_indexm(int ll, int mm, const int *l, int l_size,
                                           const int *m, int m_size);
int  oracle_offset2multipole_indexl_indexm(int offset, const int *l, int l_size,
                                           const int *m, int m_size, int *index_l, int *index_m);
int  oracle_drc3jj(double l2, double l3, double m2, double m3,
                   double *l1min, double *l1max, double *thrcof, int ndim);
double oracle_sph_bessel_j(int l, double x);   /* accurate j_l(x); NOT CLASS's approximate bessel_j */

/* ---------------------------------------------------------------------------------------
 * bessel2.c restatement
 * ------------------------------------------------------------------------------------- */
/* bessel2.c:960-985 (+ input2.c:1982-2006 for xx_max); returns xx_size, fills xx if non-NULL */
int oracle_bessel2_xx_size(double xx_max, double xx_step);
/* bessel2.c:832-943; l1 must hold l[l_size-1]+L_max_plus_m+1 ints; returns l1_size */
int oracle_bessel2_l1_list(const int *l, int l_size, int L_max_plus_m, int *l1);

/* j_l1 table record (struct bessels2, bessel2.h:96-106); rows malloc'd by the oracle */
typedef struct oracle_jl1 {
  int l1_size; int *l1; int *index_l1; int index_l1_len;
  double **j_l1; int *index_xmin_l1; int *x_size_l1; double *x_min_l1;
} oracle_jl1;

/* bessel2.c:1669-1815 for every l1 (the omp loop of bessel2.c:205-214) */
int  oracle_bessel2_j_for_all_l1(const int *l, int l_size, int L_max, int m_max,
                                 const double *xx, int xx_size, double xx_max,
                                 double j_l1_cut, double bessel_tol_x_min,
                                 double smallest_allowed_variation, oracle_jl1 *out);
void oracle_jl1_free(oracle_jl1 *t);

/* One (J,L,l,m) projection function: bessel2.c:1041-1437 (+ bessel2_J_Llm :1460-1588).
   On success *row is malloc'd with *x_size doubles.  first_3j/second_3j (size 2*L_max+1)
   may be passed in (non-NULL) to reuse externally computed Wigner symbols; l1_min is then
   read from *l1_min_io.  Otherwise they are computed with oracle_drc3jj. */
int oracle_bessel2_J_for_Llm(int projection_function, int L, int l, int m, int L_max,
                             const oracle_jl1 *jl, const double *xx, int xx_size, double xx_max,
                             double J_Llm_cut,
                             double **row, int *index_xmin, int *x_size, double *x_min);

/* bessel2.c:499-605 (bessel2_convolution): trapezoidal k integral against j_l1(k r) of one l1 */
int oracle_bessel2_convolution(const double *kk, const double *delta_kk, int k_size, const double *f, const double *g,
                               const double *j, const double *ddj, double x_min_l1, double xx_step, int cubic, double r,
                               double *integral);

/* bispectra2.c:1107-1439 (bispectra2_intrinsic_integrate_over_k3), one (l3, M3): out[index_r * n_pairs + pair] */
int oracle_integral_over_k3(int n_pairs, const int *k3_size, const long *tr_off, const double *k3_tr, const double *transfer,
                            const double *j, const double *ddj, double x_min_l1, double xx_step, int cubic,
                            int r_size, const double *r, int abs_M3, double M3_coefficient, double *out);

/* spectra2.c:241-570 (spectra2_cls), one (index_ct, index_l, index_M): the (k1,k2,k3) sum over the given pairs */
double oracle_spectra2_cl(int n_pairs, const int *pair_k1, const int *pair_k2, int k_size, const double *k,
                          const double *primordial, const int *k3_size, const long *tr_off, const double *k3_tr,
                          const double *transfer_1, const double *transfer_2, int m, int rescale_cmb_sources);

/* ---------------------------------------------------------------------------------------
 * transfer2.c restatement
 * ------------------------------------------------------------------------------------- */
typedef struct oracle_k3_info {
  double k_min, k_max;           /* ptr2->k_min_k1k2 / k_max_k1k2           */
  int k_size;                    /* ptr2->k_size_k1k2                       */
  int k_physical_start;          /* ptr2->k_physical_start_k1k2             */
  int k_physical_size;           /* ptr2->k_physical_size_k1k2              */
} oracle_k3_info;

/* transfer2.c:1329-1551 */
int oracle_get_k3_size(const oracle_setup *s, const double *k3_pt, int k3_pt_size,
                       oracle_k3_info *info);
/* transfer2.c:1566-1737 */
int oracle_get_k3_list(const oracle_setup *s, const double *k3_pt, int k3_pt_size,
                       const oracle_k3_info *info, double *k3, int *last_used_index_pt);
/* transfer2.c:2366-2585; arrays sized by oracle_time_grid_size_max() */
int oracle_time_grid_size_max(const oracle_setup *s, double k_max_global);
int oracle_get_time_grid(const oracle_setup *s, double k, double *tau_grid, int *index_tau_left,
                         double *tau0_minus_tau, double *delta_tau, int *tau_grid_size);
/* transfer2.c:2612-2763; sources = ppt2->sources[tp][k1][k2], layout [tau*k3_pt_size + j] */
int oracle_interpolate_sources_in_k(const oracle_setup *s, const double *k3_pt, int k3_pt_size,
                                    const double *sources, const oracle_k3_info *info,
                                    const double *k_grid, double *sources_k_spline,
                                    double *interpolated_sources_in_k);
/* transfer2.c:2783-2908 */
int oracle_interpolate_sources_in_time(const oracle_setup *s, int index_k,
                                       const double *interpolated_sources_in_k,
                                       const double *tau_grid, const int *index_tau_left,
                                       int tau_grid_size, double *sources_time_spline,
                                       double *interpolated_sources_in_time);
/* transfer2.c:2140-2341 */
int oracle_integrate(const oracle_setup *s, double k, int index_l, int index_m,
                     const double *const *interpolated_sources_in_time,
                     int index_J, int index_source_monopole, int L_max,
                     const double *tau0_minus_tau, const double *delta_tau, int tau_grid_size,
                     double *integral, long *n_terms);
/* transfer2.c:1861-2116 (returns the stored value) */
int oracle_compute(const oracle_setup *s, double k, int index_tt,
                   const double *const *interpolated_sources_in_time,
                   const double *tau0_minus_tau, const double *delta_tau, int tau_grid_size,
                   double *transfer, long *n_terms);

/* One (k1,k2) pair: the body of the index_k2 loop, transfer2.c:348-547.
   sources[tp] points at ppt2->sources[tp][k1][k2]; out[tt] receives k_size doubles
   (ptr2->transfer[tt][k1][k2]).  OpenMP `for schedule(static)` over k3 as transfer2.c:450. */
int oracle_transfer2_pair(const oracle_setup *s, const double *k3_pt, int k3_pt_size,
                          const double *const *sources, const oracle_k3_info *info,
                          const double *k3_grid, double *const *out, long *n_terms);

/* Many pairs, serial pair loop as transfer2.c:315,348.  Packed inputs:
   pair p has k3_pt at k3_pt[pt_off[p]..], sources at src[src_off[p] + tp*tau_size*np ...],
   fine grid at k3_tr[tr_off[p]..]; output out[tt*n_units_total + tr_off[p] + i]. */
int oracle_transfer2_run(const oracle_setup *s, int n_pairs, const int *k3_pt_size,
                         const long *pt_off, const double *k3_pt, const long *src_off,
                         const double *src, const oracle_k3_info *info, const long *tr_off,
                         const double *k3_tr, long n_units_total, double *out, long *n_terms);

int oracle_num_threads(void);
const char *oracle_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
