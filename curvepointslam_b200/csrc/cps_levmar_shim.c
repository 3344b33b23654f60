/*
 * cps_levmar_shim.c -> libcps_levmar.so: the two levmar-2.5 entry points the reference calls, under their OWN names
 * and signatures (levmar-2.5/levmar.h:98-107), forwarding to libcps.so.  A program built against levmar -- the
 * reference's curveFittingClass.cpp:681 `dlevmar_dif(modelfunc, params, final_meas, m, num_meas, 1000, opts, info,
 * NULL, NULL, &funcdata)` -- relinks against this library instead of liblevmar.a without a source change; the one
 * addition is a call to cps_register_model(modelfunc, CPS_MODEL_STEREO19) at start-up so that the function pointer
 * is recognised (a host callback cannot run on the device: an unregistered pointer is refused, there is no CPU path).
 */
#include "../../include/cps_lm.h"

int dlevmar_der(void (*func)(double *p, double *hx, int m, int n, void *adata),
                void (*jacf)(double *p, double *j, int m, int n, void *adata), double *p, double *x, int m, int n,
                int itmax, double *opts, double *info, double *work, double *covar, void *adata)
{
  return cps_dlevmar_der(func, jacf, p, x, m, n, itmax, opts, info, work, covar, adata);
}

int dlevmar_dif(void (*func)(double *p, double *hx, int m, int n, void *adata), double *p, double *x, int m, int n,
                int itmax, double *opts, double *info, double *work, double *covar, void *adata)
{
  return cps_dlevmar_dif(func, p, x, m, n, itmax, opts, info, work, covar, adata);
}
