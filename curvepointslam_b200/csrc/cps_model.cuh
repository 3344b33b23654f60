// cps_model.cuh -- device statement of the two curve models the LM kernels fit.
//
//   Stereo19 (m=19): the reference's live model, modelfunc curveFittingClass.cpp:82-119:
//     p = [curve1 (xe,ye)x4 | curve2 (xe,ye)x4 | theta, phi, z]; each ground control point goes
//     earth->body (earth2body :831-847 with generate_Rbe common.h:174-186, psi = 0) -> pinhole into both
//     images (cp_earth2image :804-817, constants common.h:70-102); the prediction of a sample is the
//     Bernstein combination of the projected control points (closest_bezier_pt(cp,t,..) :1057-1073).
//     Measurement order [L c1 | L c2 | R c1 | R c2] (:99-117).
//   Image8 (m=8): the Bernstein evaluation alone, one image-space contour per fit.
//
// The analytic Jacobian of Stereo19 does not exist in the reference (its jacmodelfunc :122-405
// differentiates an older model); it is derived ("Jacobian").
//
// Arithmetic contract: every expression is evaluated exactly as parenthesised, one rounding per
// operator (-fmad=false).  pow(x,k), k in 0..3, of the reference becomes 1, x, x*x, (x*x)*x.
#pragma once
#include "cps_detmath.cuh"

namespace cps {

// camera constants written with the arithmetic of common.h:70-87 so they fold to the same doubles
#define CPS_FX ((521.22 / 2.0 + 527.03 / 2.0) / 2.0)
#define CPS_FY ((518.32 / 2.0 + 524.73 / 2.0) / 2.0)
#define CPS_CX ((296.08 / 2.0 - 1.0 + 320.95 / 2.0 + 1.0) / 2.0)
#define CPS_CY ((236.17 / 2.0 + 228.81 / 2.0) / 2.0)
#define CPS_BASELINE 0.535

// sample-index boundaries of the four contour segments of one fit
struct FitGeom {
  int n;      // measurements (doubles) = 2 * nsamp
  int nsamp;  // sample points
  int b1, b2, b3;  // segment s covers samples [b_s, b_{s+1}), b0 = 0, b4 = nsamp
};

__device__ __forceinline__ void bernstein3(double t, double w[4]) {
  const double u = 1 - t;
  const double u2 = u * u, t2 = t * t;
  w[0] = u2 * u;
  w[1] = (3.0 * u2) * t;
  w[2] = (3.0 * u) * t2;
  w[3] = t2 * t;
}

struct Trig {
  double st, ct, sp, cp, stsp, stcp, ctsp, ctcp;
};

__device__ __noinline__ void make_trig(double theta, double phi, Trig& g) {
  det_sincos(theta, &g.st, &g.ct);
  det_sincos(phi, &g.sp, &g.cp);
  g.stsp = g.st * g.sp;
  g.stcp = g.st * g.cp;
  g.ctsp = g.ct * g.sp;
  g.ctcp = g.ct * g.cp;
}

__device__ __forceinline__ void body_point(const Trig& g, double xe, double ye, double z, double xb[3]) {
  xb[0] = g.ct * xe + g.st * z;
  xb[1] = (g.stsp * xe + g.cp * ye) - g.ctsp * z;
  xb[2] = (g.stcp * xe - g.sp * ye) - g.ctcp * z;
}

template <int MODEL>
struct Model;

// ------------------------------------------------------------------------------------------------
template <>
struct Model<19> {
  static constexpr int M = 19;
  static constexpr int NSEG = 4;
  static constexpr int SCRATCH = 8 * 3 + 8 * 3 * 5;  // projected control points + their derivatives

  // segment of sample s -> curve (0/1) and image side (0 = left, 1 = right)
  __device__ static __forceinline__ void locate(const FitGeom& g, int s, int& curve, int& side) {
    const int seg = (s >= g.b1) + (s >= g.b2) + (s >= g.b3);
    curve = seg & 1;
    side = seg >> 1;
  }

  // lanes 0..7 project one control point each: scratch[3*cp + {uL,uR,v}]
  __device__ static __forceinline__ void prep_func(const double* p, double* scratch, int lane) {
    if (lane < 8) {
      Trig g;
      make_trig(p[16], p[17], g);
      double xb[3];
      body_point(g, p[2 * lane], p[2 * lane + 1], p[18], xb);
      scratch[3 * lane + 0] = CPS_FX * (xb[1] + 0.5 * CPS_BASELINE) / xb[0] + CPS_CX;
      scratch[3 * lane + 1] = CPS_FX * (xb[1] - 0.5 * CPS_BASELINE) / xb[0] + CPS_CX;
      scratch[3 * lane + 2] = CPS_FY * (xb[2] / xb[0]) + CPS_CY;
    }
  }

  __device__ static __forceinline__ void eval_sample(const double* scratch, const FitGeom& g, int s, double t,
                                                     double& hu, double& hv) {
    int curve, side;
    locate(g, s, curve, side);
    double w[4];
    bernstein3(t, w);
    const double* c = scratch + 12 * curve;
    double au = 0.0, av = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      au += w[k] * c[3 * k + side];
      av += w[k] * c[3 * k + 2];
    }
    hu = au;
    hv = av;
  }

  // lanes 0..7: derivatives of the three image coordinates of one control point with respect to
  // (xe, ye, theta, phi, z): scratch[24 + 15*cp + 5*coord + q]
  __device__ static __forceinline__ void prep_jac(const double* p, double* scratch, int lane) {
    if (lane < 8) {
      Trig g;
      make_trig(p[16], p[17], g);
      const double xe = p[2 * lane], ye = p[2 * lane + 1], z = p[18];
      double xb[3];
      body_point(g, xe, ye, z, xb);
      double dx[5][3];
      dx[0][0] = g.ct;  dx[0][1] = g.stsp;  dx[0][2] = g.stcp;
      dx[1][0] = 0.0;   dx[1][1] = g.cp;    dx[1][2] = -g.sp;
      dx[2][0] = g.ct * z - g.st * xe;
      dx[2][1] = g.ctsp * xe + g.stsp * z;
      dx[2][2] = g.ctcp * xe + g.stcp * z;
      dx[3][0] = 0.0;   dx[3][1] = xb[2];   dx[3][2] = -xb[1];
      dx[4][0] = g.st;  dx[4][1] = -g.ctsp; dx[4][2] = -g.ctcp;
      const double yl = xb[1] + 0.5 * CPS_BASELINE;
      const double yr = xb[1] - 0.5 * CPS_BASELINE;
      const double ix2 = 1.0 / (xb[0] * xb[0]);
      double* D = scratch + 24 + 15 * lane;
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        D[0 + q] = (CPS_FX * ((dx[q][1] * xb[0]) - (yl * dx[q][0]))) * ix2;
        D[5 + q] = (CPS_FX * ((dx[q][1] * xb[0]) - (yr * dx[q][0]))) * ix2;
        D[10 + q] = (CPS_FY * ((dx[q][2] * xb[0]) - (xb[2] * dx[q][0]))) * ix2;
      }
    }
  }

  // one Jacobian row (measurement r = 2*s + uv) written densely (zeros included) to row[0..18]
  __device__ static __forceinline__ void jac_row(const double* scratch, const FitGeom& g, int r, double t,
                                                 double* row) {
    const int s = r >> 1, uv = r & 1;
    int curve, side;
    locate(g, s, curve, side);
    const int coord = uv ? 2 : side;
    double w[4];
    bernstein3(t, w);
    const double* D = scratch + 24 + 60 * curve + 5 * coord;
#pragma unroll
    for (int q = 0; q < 8; ++q) row[8 * (1 - curve) + q] = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      row[8 * curve + 2 * k] = w[k] * D[15 * k + 0];
      row[8 * curve + 2 * k + 1] = w[k] * D[15 * k + 1];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) a += w[k] * D[15 * k + 2 + q];
      row[16 + q] = a;
    }
  }

  // the same row in compact form: [0..7] the 8 entries of the row's own curve, [8..10] theta/phi/z, [11] = 0;
  // returns the curve (0/1) the row belongs to
  __device__ static __forceinline__ int jac_row_compact(const double* scratch, const FitGeom& g, int r, double t,
                                                        double* row) {
    const int s = r >> 1, uv = r & 1;
    int curve, side;
    locate(g, s, curve, side);
    const int coord = uv ? 2 : side;
    double w[4];
    bernstein3(t, w);
    const double* D = scratch + 24 + 60 * curve + 5 * coord;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      row[2 * k] = w[k] * D[15 * k + 0];
      row[2 * k + 1] = w[k] * D[15 * k + 1];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) a += w[k] * D[15 * k + 2 + q];
      row[8 + q] = a;
    }
    row[11] = 0.0;
    return curve;
  }
};

// ------------------------------------------------------------------------------------------------
template <>
struct Model<8> {
  static constexpr int M = 8;
  static constexpr int NSEG = 1;
  static constexpr int SCRATCH = 8;

  __device__ static __forceinline__ void prep_func(const double* p, double* scratch, int lane) {
    if (lane < 8) scratch[lane] = p[lane];
  }
  __device__ static __forceinline__ void eval_sample(const double* scratch, const FitGeom&, int, double t,
                                                     double& hu, double& hv) {
    double w[4];
    bernstein3(t, w);
    double au = 0.0, av = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      au += w[k] * scratch[2 * k];
      av += w[k] * scratch[2 * k + 1];
    }
    hu = au;
    hv = av;
  }
  __device__ static __forceinline__ void prep_jac(const double*, double*, int) {}
  __device__ static __forceinline__ void jac_row(const double*, const FitGeom&, int r, double t, double* row) {
    const int uv = r & 1;
    double w[4];
    bernstein3(t, w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      row[2 * k + uv] = w[k];
      row[2 * k + 1 - uv] = 0.0;
    }
  }
};

}  // namespace cps
