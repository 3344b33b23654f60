// cps_detmath.cuh -- deterministic double-precision sin/cos for the LM kernels.
//
// The reference evaluates sin/cos of the two camera angles with the host libm (common.h:174-186 via
// curveFittingClass.cpp:831-847).  CUDA's sin/cos are not bit-identical to any host libm, and the
// secant LM path amplifies a one-ulp difference into a different iteration count, so the kernels use
// a fixed sequence of IEEE-754 +,-,* operations instead: Cody-Waite reduction by pi/2 in two stages
// and the classical fdlibm kernel polynomials on [-pi/4, pi/4] (coefficients: the published fdlibm
// k_sin.c / k_cos.c minimax sets, Copyright (C) 1993 by Sun Microsystems, Inc., "Permission to use,
// copy, modify, and distribute this software is freely granted, provided that this notice is
// preserved").  < 1 ulp for |x| < 2^20*pi/2.  Compiled with -fmad=false so no operation is fused; the
// CPU oracle states the same sequence independently and tests/test_lm_gpu.py compares bits.
#pragma once

namespace cps {

__device__ __forceinline__ double det_ksin(double x, double y, bool tail) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x;
  const double v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (!tail) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

__device__ __forceinline__ double det_kcos(double x, double y) {
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

__device__ __forceinline__ void det_sincos(double x, double* s, double* c) {
  const double PIO4 = 7.85398163397448278999e-01;
  const double INVPIO2 = 6.36619772367581382433e-01;
  const double PIO2_1 = 1.57079632673412561417e+00;
  const double PIO2_2 = 6.07710050630396597660e-11;
  const double PIO2_2T = 2.02226624879595063154e-21;
  const double ax = (x >= 0.0) ? x : -x;
  if (ax <= PIO4) {
    *s = det_ksin(x, 0.0, false);
    *c = det_kcos(x, 0.0);
    return;
  }
  if (!(ax < 1647099.0)) {  // out of the deterministic domain (or NaN/Inf): library fallback
    sincos(x, s, c);
    return;
  }
  const double fn = floor(x * INVPIO2 + 0.5);
  const double t = x - fn * PIO2_1;
  const double w2 = fn * PIO2_2;
  const double r = t - w2;
  const double w = fn * PIO2_2T - ((t - r) - w2);
  const double y0 = r - w;
  const double y1 = (r - y0) - w;
  const long long q = (long long)fn;
  const double ks = det_ksin(y0, y1, true);
  const double kc = det_kcos(y0, y1);
  switch ((int)(q & 3)) {
    case 0: *s = ks; *c = kc; break;
    case 1: *s = kc; *c = -ks; break;
    case 2: *s = -ks; *c = -kc; break;
    default: *s = -kc; *c = ks; break;
  }
}

// deterministic arcsine for |x| <= 0.9: six Newton steps on sin(t) = x from t0 = x (same sequence as the oracle)
__device__ __forceinline__ double det_asin(double x) {
  const double ax = (x >= 0.0) ? x : -x;
  if (!(ax <= 0.9)) return asin(x);
  double t = x, s, c;
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
    det_sincos(t, &s, &c);
    t = t - (s - x) / c;
  }
  return t;
}

}  // namespace cps
