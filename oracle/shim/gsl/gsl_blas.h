/* Test-infrastructure shim (oracle build only): the reference asks for <gsl/gsl_blas.h>
 * purely to get CBLAS level-1 prototypes (local/include/pcgsolver/blas_wrapper.h:9,
 * src/math/blas_wrapper.h:9). GSL is not in this image, so the textbook definitions are
 * supplied inline here. Strict left-to-right summation (reference BLAS order). */
#ifndef SHKZ_ORACLE_SHIM_GSL_BLAS_H
#define SHKZ_ORACLE_SHIM_GSL_BLAS_H
#include <cmath>
#include <cstddef>
#define SHKZ_SHIM_DOT(name, R, A, X)                                              \
	static inline R name(int n, const X *x, int incx, const X *y, int incy) {      \
		A acc = 0;                                                                 \
		for (int i = 0; i < n; ++i) acc += (A)x[(size_t)i * incx] * (A)y[(size_t)i * incy]; \
		return (R)acc;                                                             \
	}
SHKZ_SHIM_DOT(cblas_sdot, float, float, float)
SHKZ_SHIM_DOT(cblas_dsdot, double, double, float)
SHKZ_SHIM_DOT(cblas_ddot, double, double, double)
#undef SHKZ_SHIM_DOT
template <class X> static inline X shkz_shim_nrm2(int n, const X *x, int inc) {
	X acc = 0; for (int i = 0; i < n; ++i) acc += x[(size_t)i * inc] * x[(size_t)i * inc]; return std::sqrt(acc);
}
template <class X> static inline X shkz_shim_asum(int n, const X *x, int inc) {
	X acc = 0; for (int i = 0; i < n; ++i) acc += std::fabs(x[(size_t)i * inc]); return acc;
}
template <class X> static inline size_t shkz_shim_iamax(int n, const X *x, int inc) {
	size_t best = 0; X bv = n > 0 ? std::fabs(x[0]) : 0;
	for (int i = 1; i < n; ++i) { X v = std::fabs(x[(size_t)i * inc]); if (v > bv) { bv = v; best = i; } }
	return best;
}
static inline float cblas_snrm2(int n, const float *x, int inc) { return shkz_shim_nrm2(n, x, inc); }
static inline double cblas_dnrm2(int n, const double *x, int inc) { return shkz_shim_nrm2(n, x, inc); }
static inline float cblas_sasum(int n, const float *x, int inc) { return shkz_shim_asum(n, x, inc); }
static inline double cblas_dasum(int n, const double *x, int inc) { return shkz_shim_asum(n, x, inc); }
static inline size_t cblas_isamax(int n, const float *x, int inc) { return shkz_shim_iamax(n, x, inc); }
static inline size_t cblas_idamax(int n, const double *x, int inc) { return shkz_shim_iamax(n, x, inc); }
static inline void cblas_saxpy(int n, float a, const float *x, int incx, float *y, int incy) {
	for (int i = 0; i < n; ++i) y[(size_t)i * incy] += a * x[(size_t)i * incx];
}
static inline void cblas_daxpy(int n, double a, const double *x, int incx, double *y, int incy) {
	for (int i = 0; i < n; ++i) y[(size_t)i * incy] += a * x[(size_t)i * incx];
}
static inline void cblas_sscal(int n, float a, float *x, int inc) { for (int i = 0; i < n; ++i) x[(size_t)i * inc] *= a; }
static inline void cblas_dscal(int n, double a, double *x, int inc) { for (int i = 0; i < n; ++i) x[(size_t)i * inc] *= a; }
#endif
