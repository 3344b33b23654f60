/*
 * orc_refshim.c -- adapters that run the REFERENCE's own levmar-2.5 (compiled unchanged into
 * oracle/_ref/liblevmar_ref.so by `make ref`) under the oracle's batched driver and restated curve
 * model.  TEST INFRASTRUCTURE ONLY.  Built into oracle/_ref/liborc_refshim.so (git-ignored).
 *
 * Declarations follow the reference ABI, levmar-2.5/levmar.h:98-107.
 */
#include "orc_model.h"

extern int dlevmar_der(void (*func)(double *p, double *hx, int m, int n, void *adata),
                       void (*jacf)(double *p, double *j, int m, int n, void *adata), double *p,
                       double *x, int m, int n, int itmax, double *opts, double *info, double *work,
                       double *covar, void *adata);
extern int dlevmar_dif(void (*func)(double *p, double *hx, int m, int n, void *adata), double *p,
                       double *x, int m, int n, int itmax, double *opts, double *info, double *work,
                       double *covar, void *adata);

static int ref_der(orc_func_t func, orc_func_t jacf, double *p, const double *x, int m, int n, int itmax,
                   const double *opts, double *info, void *adata, orc_lm_extra *extra)
{
  if (extra) { extra->n_jtj = -1; extra->n_broyden = -1; }
  return dlevmar_der(func, jacf, p, (double *)x, m, n, itmax, (double *)opts, info, 0, 0, adata);
}

static int ref_dif(orc_func_t func, double *p, const double *x, int m, int n, int itmax,
                   const double *opts, double *info, void *adata, orc_lm_extra *extra)
{
  if (extra) { extra->n_jtj = -1; extra->n_broyden = -1; }
  return dlevmar_dif(func, p, (double *)x, m, n, itmax, (double *)opts, info, 0, 0, adata);
}

int ref_fit_batch(int model, int variant, int math_mode, int B, double *p, const double *meas,
                  const double *t, const long long *off, const int *counts, int itmax,
                  const double *opts, double *info, int *ret, int *extra, int nthreads)
{
  return orc_fit_batch_impl(ref_der, ref_dif, model, variant, math_mode, B, p, meas, t, off, counts,
                            itmax, opts, info, ret, extra, nthreads);
}
