/*
 * orc_detmath.h -- deterministic sin/cos for the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * Why this exists: the reference evaluates sin/cos of the two camera angles with glibc
 * (common.h:174-186 via curveFittingClass.cpp:831-847).  glibc's and CUDA's double-precision
 * sin/cos are both accurate to about one ulp but are NOT bit-identical to each other, and the
 * secant LM path amplifies a one-ulp difference into a different iteration count.
 * To make "GPU == oracle" a bit-exact statement, the oracle's default math mode and the CUDA
 * kernels both evaluate sin/cos with the SAME sequence of IEEE-754 +,-,* operations written
 * below (no FMA, no table lookups).  The oracle's "libm" math mode (orc_model.h) keeps glibc's
 * sin/cos/pow exactly as the reference calls them; tests report the distance between the modes.
 *
 * Algorithm: classical Cody-Waite reduction by pi/2 in two stages followed by the fdlibm kernel
 * polynomials on [-pi/4, pi/4] (coefficients are the published fdlibm minimax coefficients,
 * k_sin.c / k_cos.c; "Copyright (C) 1993 by Sun Microsystems, Inc. ... Permission to use, copy,
 * modify, and distribute this software is freely granted, provided that this notice is
 * preserved").  Accurate to < 1 ulp for |x| < 2^20 * pi/2; outside that range (or non-finite)
 * the caller-provided fallback is used and determinism across platforms is not claimed.
 *
 * The independent device-side statement of the same operation sequence lives in
 * curvepointslam_b200/csrc/cps_detmath.cuh; tests/test_detmath.py checks the two agree bit for
 * bit on the GPU and that both stay within 1 ulp of glibc.
 */
#ifndef ORC_DETMATH_H
#define ORC_DETMATH_H

#include <math.h>

static inline double orc_ksin(double x, double y, int have_tail)
{
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x;
  const double v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (!have_tail) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

static inline double orc_kcos(double x, double y)
{
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = x * x;
  const double w = z * z;
  const double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
  const double hz = 0.5 * z;
  const double one_m = 1.0 - hz;
  return one_m + (((1.0 - one_m) - hz) + (z * r - x * y));
}

/* returns 1 if the deterministic path was taken, 0 if it fell back to libm */
static inline int orc_det_sincos(double x, double *s, double *c)
{
  const double PIO4 = 7.85398163397448278999e-01;
  const double INVPIO2 = 6.36619772367581382433e-01;
  const double PIO2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
  const double PIO2_2 = 6.07710050630396597660e-11;  /* second 33 bits        */
  const double PIO2_2T = 2.02226624879595063154e-21; /* pi/2 - (PIO2_1+PIO2_2) */
  const double ax = (x >= 0.0) ? x : -x;

  if (ax <= PIO4) {
    *s = orc_ksin(x, 0.0, 0);
    *c = orc_kcos(x, 0.0);
    return 1;
  }
  if (!(ax < 1647099.0)) { /* 2^20 * pi/2 ~= 1647099.3 ; also catches NaN/Inf */
    *s = sin(x);
    *c = cos(x);
    return 0;
  }
  {
    const double fn = floor(x * INVPIO2 + 0.5);
    const double t = x - fn * PIO2_1;
    const double w2 = fn * PIO2_2;
    const double r = t - w2;
    const double w = fn * PIO2_2T - ((t - r) - w2);
    const double y0 = r - w;
    const double y1 = (r - y0) - w;
    const long long q = (long long)fn;
    const double ks = orc_ksin(y0, y1, 1);
    const double kc = orc_kcos(y0, y1);
    switch ((int)(q & 3)) {
      case 0: *s = ks; *c = kc; break;
      case 1: *s = kc; *c = -ks; break;
      case 2: *s = -ks; *c = -kc; break;
      default: *s = -kc; *c = ks; break;
    }
    return 1;
  }
}

/* deterministic arcsine for |x| <= 0.9: Newton on sin(t) = x with the deterministic sin/cos above, six steps
 * from t0 = x (quadratic convergence: exact to the last bits well before the sixth step, and a fixed step count
 * keeps the operation sequence identical everywhere).  Outside that range, or for non-finite x, falls back to
 * libm.  Returns 1 if the deterministic path was taken. */
static inline int orc_det_asin(double x, double *out)
{
  const double ax = (x >= 0.0) ? x : -x;
  double t = x, s, c;
  int k;
  if (!(ax <= 0.9)) { *out = asin(x); return 0; }
  for (k = 0; k < 6; ++k) {
    orc_det_sincos(t, &s, &c);
    t = t - (s - x) / c;
  }
  *out = t;
  return 1;
}

#endif
