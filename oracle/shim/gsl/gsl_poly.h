/*
 * oracle/shim/gsl/gsl_poly.h -- stand-in for the one GSL entry point curveFittingClass.cpp:936-1043 uses
 * (gsl_poly_complex_solve: all complex roots of a real polynomial).  GSL is not vendored by the reference and
 * not installed here; cv_shim.cpp implements the call with GSL's documented method (eigenvalues of the balanced
 * companion matrix by shifted QR).  TEST INFRASTRUCTURE ONLY.
 */
#ifndef CPS_ORACLE_SHIM_GSL_POLY_H
#define CPS_ORACLE_SHIM_GSL_POLY_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t nc; double *matrix; } gsl_poly_complex_workspace;
typedef double *gsl_complex_packed_ptr;
gsl_poly_complex_workspace *gsl_poly_complex_workspace_alloc(size_t n);
void gsl_poly_complex_workspace_free(gsl_poly_complex_workspace *w);
/* a[0..n-1]: coefficients, a[n-1] != 0; z: 2(n-1) doubles (re, im).  Returns 0 on success. */
int gsl_poly_complex_solve(const double *a, size_t n, gsl_poly_complex_workspace *w, gsl_complex_packed_ptr z);
#ifdef __cplusplus
}
#endif
#endif
