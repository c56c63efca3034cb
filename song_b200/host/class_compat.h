/* class_compat.h -- stand-ins for CLASS-fork helpers (see class_compat.c). */
#ifndef SONG_CLASS_COMPAT_H
#define SONG_CLASS_COMPAT_H
#ifdef __cplusplus
extern "C" {
#endif
void   compat_lin_space(double *x, double x_min, double x_max, int n);
void   compat_sph_bessel_table(int lmax, double x, double *out /* lmax+1 */);
double compat_bessel_j(int l, double x);
/* second derivatives of the cubic spline through (x[i], y[i]); est_deriv = 0: natural (ddy = 0 at both ends),
   1: end slopes estimated from the first / last three points.  Returns 1 when n < 3 or two abscissae coincide. */
int    compat_spline_second_derivatives(const double *x, int n, const double *y, double *ddy, int est_deriv);
int    compat_drc3jj(double l2, double l3, double m2, double m3, double *l1min, double *l1max,
                     double *thrcof, int ndim);
#ifdef __cplusplus
}
#endif
#endif
