/* class_compat.h -- stand-ins for CLASS-fork helpers (see class_compat.c). */
#ifndef SONG_CLASS_COMPAT_H
#define SONG_CLASS_COMPAT_H
#ifdef __cplusplus
extern "C" {
#endif
void   compat_lin_space(double *x, double x_min, double x_max, int n);
void   compat_sph_bessel_table(int lmax, double x, double *out /* lmax+1 */);
double compat_bessel_j(int l, double x);
int    compat_drc3jj(double l2, double l3, double m2, double m3, double *l1min, double *l1max,
                     double *thrcof, int ndim);
#ifdef __cplusplus
}
#endif
#endif
