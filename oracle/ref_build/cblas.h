/* TEST INFRASTRUCTURE (oracle/): stand-in for the system <cblas.h>, which this image does not ship.
 *
 * The reference's preprocessing extension (/root/reference/src/geoconv/preprocessing/c_extension/c_extension.c:7)
 * includes "cblas.h" and calls five level-1 routines.  The image has no BLAS development package, but scipy's
 * wheel bundles a complete OpenBLAS whose C interface is exported with a `scipy_` prefix
 * (scipy.libs/libscipy_openblas-*.so: scipy_cblas_ddot, ...).  This header declares exactly those five entry
 * points and maps the unprefixed names onto them, so the reference source compiles UNMODIFIED, where it lies,
 * against a real OpenBLAS.  Built by oracle/build_ref.sh into oracle/_ref/ (git-ignored); used only to
 * generate and check fixtures - never by the product.
 */
#ifndef GCB_ORACLE_CBLAS_SHIM_H
#define GCB_ORACLE_CBLAS_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
double scipy_cblas_dnrm2(int n, const double *x, int incx);
double scipy_cblas_ddot(int n, const double *x, int incx, const double *y, int incy);
void scipy_cblas_dcopy(int n, const double *x, int incx, double *y, int incy);
void scipy_cblas_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy);
void scipy_cblas_dscal(int n, double alpha, double *x, int incx);
#define cblas_dnrm2 scipy_cblas_dnrm2
#define cblas_ddot scipy_cblas_ddot
#define cblas_dcopy scipy_cblas_dcopy
#define cblas_daxpy scipy_cblas_daxpy
#define cblas_dscal scipy_cblas_dscal
#ifdef __cplusplus
}
#endif
#endif
