/*
 * orc_model.c -- CPU oracle for the two curve models and the fit_curve driver pieces.
 * TEST INFRASTRUCTURE ONLY (see orc_model.h for the reference lines restated here).
 *
 * Arithmetic contract shared with the CUDA kernels (csrc/cps_model.cuh): every expression below is
 * evaluated left to right exactly as parenthesised, one IEEE-754 rounding per operator, no FMA
 * (-ffp-contract=off here, -fmad=false there).  The CUDA side is an independent statement of the
 * same formulas; tests compare the two bit for bit.
 */
#include "orc_model.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "orc_detmath.h"

/* ---- Bernstein weights: binom[3][k] * pow(1-t, 3-k) * pow(t, k), curveFittingClass.cpp:1067-1068.
 * LIBM mode calls pow() like the reference does: inside a loop over k, i.e. with RUN-TIME exponents.  The
 * exponent goes through a volatile so that the compiler cannot fold pow(x, 2.0) into x*x: glibc's pow is
 * accurate to < 1 ulp but not correctly rounded, and pow(t, 2.0) != t*t for about one t in 10^4.  With the
 * genuine calls this mode is bit-identical to the reference's own modelfunc compiled from
 * /root/reference/curveFittingClass.cpp (tests/test_ref_pinning.py).  DET mode uses pow(x,0)=1, pow(x,1)=x,
 * pow(x,2)=x*x and pow(x,3)=(x*x)*x (each may differ from glibc by 1 ulp). */
static double orc_pow_rt(double x, int k)
{
  volatile double e = (double)k;
  return pow(x, e);
}

void orc_bernstein3(double t, int math_mode, double w[4])
{
  const double u = 1 - t;
  if (math_mode == ORC_MATH_LIBM) {
    w[0] = 1.0 * orc_pow_rt(u, 3) * orc_pow_rt(t, 0);
    w[1] = 3.0 * orc_pow_rt(u, 2) * orc_pow_rt(t, 1);
    w[2] = 3.0 * orc_pow_rt(u, 1) * orc_pow_rt(t, 2);
    w[3] = 1.0 * orc_pow_rt(u, 0) * orc_pow_rt(t, 3);
  } else {
    const double u2 = u * u, t2 = t * t;
    w[0] = (1.0 * (u2 * u)) * 1.0;
    w[1] = (3.0 * u2) * t;
    w[2] = (3.0 * u) * t2;
    w[3] = (1.0 * 1.0) * (t2 * t);
  }
}

static void orc_sincos(double a, int math_mode, double *s, double *c)
{
  if (math_mode == ORC_MATH_LIBM) { *s = sin(a); *c = cos(a); }
  else orc_det_sincos(a, s, c);
}

/* One ground control point (xe, ye, 0) -> body frame -> both images.
 * earth2body with psi = 0 (curveFittingClass.cpp:831-847, common.h:174-186): with cos(0)=1 and
 * sin(0)=0 exactly the rotation reduces to
 *   R = [ ct, 0, -st ; st*sp, cp, ct*sp ; st*cp, -sp, ct*cp ],  d = (xe, ye, -z),
 * and each row of R*d is summed left to right.  Pinhole: cp_earth2image :804-817. */
typedef struct { double st, ct, sp, cp, stsp, stcp, ctsp, ctcp; } orc_trig;

static void orc_trig_make(double theta, double phi, int math_mode, orc_trig *g)
{
  orc_sincos(theta, math_mode, &g->st, &g->ct);
  orc_sincos(phi, math_mode, &g->sp, &g->cp);
  g->stsp = g->st * g->sp;
  g->stcp = g->st * g->cp;
  g->ctsp = g->ct * g->sp;
  g->ctcp = g->ct * g->cp;
}

static void orc_body(const orc_trig *g, double xe, double ye, double z, double xb[3])
{
  xb[0] = g->ct * xe + g->st * z;
  xb[1] = (g->stsp * xe + g->cp * ye) - g->ctsp * z;
  xb[2] = (g->stcp * xe - g->sp * ye) - g->ctcp * z;
}

/* img[0] = u_left, img[1] = u_right, img[2] = v (same in both images) */
static void orc_project(const double xb[3], double img[3])
{
  img[0] = ORC_FX * (xb[1] + 0.5 * ORC_BASELINE) / xb[0] + ORC_CX;
  img[1] = ORC_FX * (xb[1] - 0.5 * ORC_BASELINE) / xb[0] + ORC_CX;
  img[2] = ORC_FY * (xb[2] / xb[0]) + ORC_CY;
}

/* segment -> (curve, image side): measurement order is [L c1 | L c2 | R c1 | R c2],
 * curveFittingClass.cpp:99-117 */
static const int orc_seg_curve[4] = {0, 1, 0, 1};
static const int orc_seg_side[4] = {0, 0, 1, 1};

void orc_stereo19_func(double *p, double *hx, int m, int n, void *adata)
{
  const orc_fit_data *fd = (const orc_fit_data *)adata;
  orc_trig g;
  double cpimg[8][3]; /* [curve*4+k][uL,uR,v] */
  int c, k, seg, i, s = 0;
  (void)m; (void)n;

  orc_trig_make(p[16], p[17], fd->math_mode, &g);
  for (c = 0; c < 2; ++c)
    for (k = 0; k < 4; ++k) {
      double xb[3];
      orc_body(&g, p[8 * c + 2 * k], p[8 * c + 2 * k + 1], p[18], xb);
      orc_project(xb, cpimg[4 * c + k]);
    }
  for (seg = 0; seg < 4; ++seg) {
    const int cu = orc_seg_curve[seg], side = orc_seg_side[seg];
    for (i = 0; i < fd->count[seg]; ++i, ++s) {
      double w[4], ax = 0.0, ay = 0.0;
      orc_bernstein3(fd->t[s], fd->math_mode, w);
      for (k = 0; k < 4; ++k) {
        ax += w[k] * cpimg[4 * cu + k][side];
        ay += w[k] * cpimg[4 * cu + k][2];
      }
      hx[2 * s] = ax;
      hx[2 * s + 1] = ay;
    }
  }
}

/* Analytic Jacobian of the live m=19 model (derived; the reference has none for
 * this model).  For control point (c,k) and parameter q in {xe, ye, theta, phi, z}:
 *   d(xb)/d(xe) = R[:,0]  d(xb)/d(ye) = R[:,1]  d(xb)/d(z) = -R[:,2]
 *   d(xb)/d(theta) = (-st*xe + ct*z, ctsp*xe + stsp*z, ctcp*xe + stcp*z)
 *   d(xb)/d(phi)   = (0, xb2, -xb1)
 *   du = (FX*((dyb*xb0) - (y*dxb0)))*(1/(xb0*xb0)),  y = xb1 +- B/2
 *   dv = (FY*((dzb*xb0) - (xb2*dxb0)))*(1/(xb0*xb0))
 * Row of sample s: w_k * d(.)/d(xe_k, ye_k) at columns 8c+2k, 8c+2k+1 and the left-to-right sum
 * over k of w_k * d(.)_k/d(theta, phi, z) at columns 16..18. */
static void orc_cp_derivs(const orc_trig *g, double xe, double ye, double z, double D[3][5])
{
  double xb[3], dxb[5][3], yl, yr, ix2;
  int q;
  orc_body(g, xe, ye, z, xb);
  dxb[0][0] = g->ct;   dxb[0][1] = g->stsp; dxb[0][2] = g->stcp;   /* d/d xe    */
  dxb[1][0] = 0.0;     dxb[1][1] = g->cp;   dxb[1][2] = -g->sp;    /* d/d ye    */
  dxb[2][0] = g->ct * z - g->st * xe;                              /* d/d theta */
  dxb[2][1] = g->ctsp * xe + g->stsp * z;
  dxb[2][2] = g->ctcp * xe + g->stcp * z;
  dxb[3][0] = 0.0;     dxb[3][1] = xb[2];   dxb[3][2] = -xb[1];    /* d/d phi   */
  dxb[4][0] = g->st;   dxb[4][1] = -g->ctsp; dxb[4][2] = -g->ctcp; /* d/d z     */
  yl = xb[1] + 0.5 * ORC_BASELINE;
  yr = xb[1] - 0.5 * ORC_BASELINE;
  ix2 = 1.0 / (xb[0] * xb[0]);
  for (q = 0; q < 5; ++q) {
    D[0][q] = (ORC_FX * ((dxb[q][1] * xb[0]) - (yl * dxb[q][0]))) * ix2;
    D[1][q] = (ORC_FX * ((dxb[q][1] * xb[0]) - (yr * dxb[q][0]))) * ix2;
    D[2][q] = (ORC_FY * ((dxb[q][2] * xb[0]) - (xb[2] * dxb[q][0]))) * ix2;
  }
}

void orc_stereo19_jac(double *p, double *jac, int m, int n, void *adata)
{
  const orc_fit_data *fd = (const orc_fit_data *)adata;
  orc_trig g;
  double D[8][3][5];
  int c, k, seg, i, q, s = 0;
  (void)n;

  orc_trig_make(p[16], p[17], fd->math_mode, &g);
  for (c = 0; c < 2; ++c)
    for (k = 0; k < 4; ++k)
      orc_cp_derivs(&g, p[8 * c + 2 * k], p[8 * c + 2 * k + 1], p[18], D[4 * c + k]);

  for (seg = 0; seg < 4; ++seg) {
    const int cu = orc_seg_curve[seg], side = orc_seg_side[seg];
    for (i = 0; i < fd->count[seg]; ++i, ++s) {
      double w[4];
      double *ru = jac + (size_t)(2 * s) * m, *rv = ru + m;
      orc_bernstein3(fd->t[s], fd->math_mode, w);
      for (q = 0; q < m; ++q) { ru[q] = 0.0; rv[q] = 0.0; }
      for (k = 0; k < 4; ++k) {
        ru[8 * cu + 2 * k] = w[k] * D[4 * cu + k][side][0];
        ru[8 * cu + 2 * k + 1] = w[k] * D[4 * cu + k][side][1];
        rv[8 * cu + 2 * k] = w[k] * D[4 * cu + k][2][0];
        rv[8 * cu + 2 * k + 1] = w[k] * D[4 * cu + k][2][1];
      }
      for (q = 0; q < 3; ++q) {
        double au = 0.0, av = 0.0;
        for (k = 0; k < 4; ++k) {
          au += w[k] * D[4 * cu + k][side][2 + q];
          av += w[k] * D[4 * cu + k][2][2 + q];
        }
        ru[16 + q] = au;
        rv[16 + q] = av;
      }
    }
  }
}

/* m=8 image-space model: the Bernstein evaluation alone (curveFittingClass.cpp:1057-1073) */
void orc_image8_func(double *p, double *hx, int m, int n, void *adata)
{
  const orc_fit_data *fd = (const orc_fit_data *)adata;
  int s, k;
  (void)m;
  for (s = 0; s < n / 2; ++s) {
    double w[4], ax = 0.0, ay = 0.0;
    orc_bernstein3(fd->t[s], fd->math_mode, w);
    for (k = 0; k < 4; ++k) {
      ax += w[k] * p[2 * k];
      ay += w[k] * p[2 * k + 1];
    }
    hx[2 * s] = ax;
    hx[2 * s + 1] = ay;
  }
}

void orc_image8_jac(double *p, double *jac, int m, int n, void *adata)
{
  const orc_fit_data *fd = (const orc_fit_data *)adata;
  int s, k, q;
  (void)p;
  for (s = 0; s < n / 2; ++s) {
    double w[4];
    double *ru = jac + (size_t)(2 * s) * m, *rv = ru + m;
    orc_bernstein3(fd->t[s], fd->math_mode, w);
    for (q = 0; q < m; ++q) { ru[q] = 0.0; rv[q] = 0.0; }
    for (k = 0; k < 4; ++k) {
      ru[2 * k] = w[k];
      rv[2 * k + 1] = w[k];
    }
  }
}

/* t_i = clamp((y_i - y_first)/(y_last - y_first), 0, 1), curveFittingClass.cpp:451-511.  A segment
 * whose first and last rows coincide yields NaN/Inf exactly like the reference (the clamps are
 * written so that NaN passes through). */
void orc_segment_t(const double *xy, int npts, double *t_out)
{
  int i;
  if (npts <= 0) return;
  {
    const double firsty = xy[1], lasty = xy[2 * npts - 1];
    for (i = 0; i < npts; ++i) {
      double v = (xy[2 * i + 1] - firsty) / (lasty - firsty);
      v = (v < 0.0) ? 0.0 : v;
      v = (v > 1.0) ? 1.0 : v;
      t_out[i] = v;
    }
  }
}

/* ---- fit_curve's initial guess, curveFittingClass.cpp:516-651 ------------------------------------------- */
static void orc_triangulate(const double *m, int a, int b, double pt[3])
{ /* :522-547: left sample a, right sample b (point indices into the packed measurement vector) */
  const double uL = m[2 * a], vL = m[2 * a + 1], uR = m[2 * b], vR = m[2 * b + 1];
  pt[0] = ORC_FX * ORC_BASELINE / (uL - uR);
  pt[1] = 0.5 * (pt[0] / ORC_FX * (uL - ORC_CX) - 0.5 * ORC_BASELINE) + 0.5 * (pt[0] / ORC_FX * (uR - ORC_CX) + 0.5 * ORC_BASELINE);
  pt[2] = 0.5 * pt[0] / ORC_FY * (vL - ORC_CY) + 0.5 * pt[0] / ORC_FY * (vR - ORC_CY);
}

/* eigenvector of the smallest eigenvalue of the symmetric 3x3 S = M^T M by cyclic Jacobi rotations (a fixed
 * 8 sweeps, every operation written out: the CUDA kernel states the same sequence) */
static void orc_smallest_eigvec3(const double S[9], double v[3])
{
  double a[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  int sweep, p, q, k, best;
  for (p = 0; p < 3; ++p)
    for (q = 0; q < 3; ++q) a[p][q] = S[3 * p + q];
  for (sweep = 0; sweep < 8; ++sweep)
    for (p = 0; p < 2; ++p)
      for (q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        double theta, t, c, s;
        if (apq == 0.0) continue;
        theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        t = 1.0 / (((theta >= 0.0) ? theta : -theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        c = 1.0 / sqrt(t * t + 1.0);
        s = t * c;
        for (k = 0; k < 3; ++k) { /* A <- A J */
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (k = 0; k < 3; ++k) { /* A <- J^T A */
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  best = 0;
  if (a[1][1] < a[best][best]) best = 1;
  if (a[2][2] < a[best][best]) best = 2;
  for (k = 0; k < 3; ++k) v[k] = V[k][best];
}

void orc_stereo19_init_guess(const double *meas, const int *counts, double *p19, int math_mode, int sign)
{
  const int nl1 = counts[0], nl2 = counts[1], nr1 = counts[2], nr2 = counts[3];
  const int npl = nl1 + nl2, npts = npl + nr1 + nr2;
  double s1[3], e1[3], s2[3], e2[3], c[3], M[4][3], S[9], n[3], d, theta, phi, origin[3], Rot[9], st, ct, sp, cp;
  double *pts[4];
  double cp1[8], cp2[8];
  int i, j, k;
  orc_triangulate(meas, 0, npl, s1);                      /* :523-525 */
  orc_triangulate(meas, nl1 - 1, npl + nr1 - 1, e1);      /* :528-530 */
  orc_triangulate(meas, nl1, npl + nr1, s2);              /* :537-539 */
  orc_triangulate(meas, npl - 1, npts - 1, e2);           /* :542-544 */
  for (j = 0; j < 3; ++j) c[j] = 0.25 * (s1[j] + s2[j] + e1[j] + e2[j]);
  for (j = 0; j < 3; ++j) {
    M[0][j] = s1[j] - c[j];
    M[1][j] = s2[j] - c[j];
    M[2][j] = e1[j] - c[j];
    M[3][j] = e2[j] - c[j];
  }
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (k = 0; k < 4; ++k) acc += M[k][i] * M[k][j];
      S[3 * i + j] = acc;
    }
  orc_smallest_eigvec3(S, n);
  d = (n[0] * c[0] + n[1] * c[1]) + n[2] * c[2];
  if ((sign >= 0 && d < 0.0) || (sign < 0 && d > 0.0)) {
    n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2];
    d = -d;
  }
  if (math_mode == ORC_MATH_LIBM) { theta = -asin(n[0]); phi = asin(n[1]); }
  else { orc_det_asin(n[0], &theta); theta = -theta; orc_det_asin(n[1], &phi); }
  for (j = 0; j < 3; ++j) origin[j] = d * n[j];
  orc_sincos(theta, math_mode, &st, &ct);
  orc_sincos(phi, math_mode, &sp, &cp);
  Rot[0] = ct;  Rot[1] = st * sp; Rot[2] = st * cp;  /* :581-589, row-major */
  Rot[3] = 0.0; Rot[4] = cp;      Rot[5] = -sp;
  Rot[6] = -st; Rot[7] = ct * sp; Rot[8] = ct * cp;
  pts[0] = s1; pts[1] = e1; pts[2] = s2; pts[3] = e2;
  for (i = 0; i < 4; ++i) {
    double t[3], r[3];
    for (j = 0; j < 3; ++j) t[j] = pts[i][j] - origin[j];
    for (j = 0; j < 3; ++j) r[j] = (Rot[3 * j] * t[0] + Rot[3 * j + 1] * t[1]) + Rot[3 * j + 2] * t[2];
    for (j = 0; j < 3; ++j) pts[i][j] = r[j];
  }
  cp1[0] = s1[0]; cp1[1] = s1[1]; cp1[6] = e1[0]; cp1[7] = e1[1];
  cp2[0] = s2[0]; cp2[1] = s2[1]; cp2[6] = e2[0]; cp2[7] = e2[1];
  for (i = 1; i < 3; ++i) { /* :626-632 */
    cp1[2 * i] = (3 - i) * cp1[0] / 3 + i * cp1[6] / 3;
    cp1[2 * i + 1] = (3 - i) * cp1[1] / 3 + i * cp1[7] / 3;
    cp2[2 * i] = (3 - i) * cp2[0] / 3 + i * cp2[6] / 3;
    cp2[2 * i + 1] = (3 - i) * cp2[1] / 3 + i * cp2[7] / 3;
  }
  for (i = 0; i < 8; ++i) { p19[i] = cp1[i]; p19[8 + i] = cp2[i]; }
  p19[16] = theta;
  p19[17] = phi;
  p19[18] = -d;
}

/* curveFittingClass.cpp:684-711, including the "+= PI" on the height in the second branch */
void orc_stereo19_postfit(double *p)
{
  int i;
  if (p[18] > 0.0) {
    for (i = 0; i < 8; ++i) p[2 * i + 1] *= -1.0;
    p[16] *= -1.0;
    p[17] += ORC_PI;
    p[18] *= -1.0;
  }
  if (p[16] < -ORC_PI / 2.0) {
    for (i = 0; i < 8; ++i) p[2 * i + 1] *= -1.0;
    p[17] = (-ORC_PI - p[17]);
    p[18] += ORC_PI;
  }
  while (p[16] > ORC_PI) p[16] -= 2 * ORC_PI;
  while (p[16] < -ORC_PI) p[16] += 2 * ORC_PI;
  while (p[17] > ORC_PI) p[17] -= 2 * ORC_PI;
  while (p[17] < -ORC_PI) p[17] += 2 * ORC_PI;
}

/* ---- batched driver (pthread work queue over fits; no OpenMP dependency) ---- */
#include <pthread.h>

typedef struct {
  orc_lm_der_fn der;
  orc_lm_dif_fn dif;
  int model, variant, math_mode, B, itmax;
  double *p;
  const double *meas, *t, *opts;
  const long long *off;
  const int *counts;
  double *info;
  int *ret, *extra;
  int next; /* atomically advanced chunk cursor */
} orc_batch_job;

static void orc_fit_one(const orc_batch_job *J, int f)
{
  const int m = (J->model == ORC_MODEL_STEREO19) ? 19 : 8;
  const int nseg = (J->model == ORC_MODEL_STEREO19) ? 4 : 1;
  const orc_func_t func = (J->model == ORC_MODEL_STEREO19) ? orc_stereo19_func : orc_image8_func;
  const orc_func_t jacf = (J->model == ORC_MODEL_STEREO19) ? orc_stereo19_jac : orc_image8_jac;
  orc_fit_data fd;
  orc_lm_extra ex = {0, 0};
  const int n = (int)(J->off[f + 1] - J->off[f]);
  double *tloc = NULL;
  double linfo[ORC_LM_INFO_SZ];
  int s, r, k;
  memset(linfo, 0, sizeof linfo);
  for (s = 0; s < 4; ++s) fd.count[s] = (s < nseg) ? J->counts[(size_t)f * nseg + s] : 0;
  fd.math_mode = J->math_mode;
  if (J->t) {
    fd.t = J->t + J->off[f] / 2;
  } else { /* derive t from the image rows like fit_curve does */
    const double *xy = J->meas + J->off[f];
    tloc = (double *)malloc(sizeof(double) * (size_t)(n / 2 + 1));
    for (s = 0, k = 0; s < nseg; ++s) {
      orc_segment_t(xy + 2 * k, fd.count[s], tloc + k);
      k += fd.count[s];
    }
    fd.t = tloc;
  }
  if (J->variant == ORC_LM_DER)
    r = J->der(func, jacf, J->p + (size_t)f * m, J->meas + J->off[f], m, n, J->itmax, J->opts, linfo,
               &fd, &ex);
  else
    r = J->dif(func, J->p + (size_t)f * m, J->meas + J->off[f], m, n, J->itmax, J->opts, linfo, &fd,
               &ex);
  if (J->info) memcpy(J->info + (size_t)f * ORC_LM_INFO_SZ, linfo, sizeof linfo);
  if (J->ret) J->ret[f] = r;
  if (J->extra) { J->extra[2 * f] = ex.n_jtj; J->extra[2 * f + 1] = ex.n_broyden; }
  free(tloc);
}

static void *orc_batch_worker(void *arg)
{
  orc_batch_job *J = (orc_batch_job *)arg;
  for (;;) {
    const int lo = __atomic_fetch_add(&J->next, 8, __ATOMIC_RELAXED);
    int f;
    if (lo >= J->B) break;
    for (f = lo; f < lo + 8 && f < J->B; ++f) orc_fit_one(J, f);
  }
  return NULL;
}

int orc_fit_batch_impl(orc_lm_der_fn der, orc_lm_dif_fn dif, int model, int variant, int math_mode,
                       int B, double *p, const double *meas, const double *t, const long long *off,
                       const int *counts, int itmax, const double *opts, double *info, int *ret,
                       int *extra, int nthreads)
{
  orc_batch_job J;
  pthread_t th[256];
  int i, started = 0;
  if (model != ORC_MODEL_STEREO19 && model != ORC_MODEL_IMAGE8) return -1;
  if (variant != ORC_LM_DER && variant != ORC_LM_DIF) return -1;
  if (B < 0 || !p || !meas || !off || !counts) return -1;
  J.der = der; J.dif = dif; J.model = model; J.variant = variant; J.math_mode = math_mode; J.B = B;
  J.itmax = itmax; J.p = p; J.meas = meas; J.t = t; J.opts = opts; J.off = off; J.counts = counts;
  J.info = info; J.ret = ret; J.extra = extra; J.next = 0;
  if (nthreads > 256) nthreads = 256;
  for (i = 1; i < nthreads; ++i)
    if (pthread_create(&th[started], NULL, orc_batch_worker, &J) == 0) ++started;
  orc_batch_worker(&J);
  for (i = 0; i < started; ++i) pthread_join(th[i], NULL);
  return 0;
}

int orc_fit_batch(int model, int variant, int math_mode, int B, double *p, const double *meas,
                  const double *t, const long long *off, const int *counts, int itmax,
                  const double *opts, double *info, int *ret, int *extra, int nthreads)
{
  return orc_fit_batch_impl(orc_lm_der, orc_lm_dif, model, variant, math_mode, B, p, meas, t, off,
                            counts, itmax, opts, info, ret, extra, nthreads);
}

/* ---- cleanup_and_group_edges, curveFittingClass.cpp:1207-1418 ("next" row 4) ----------------------------------
 * Integer trimming of the four edge sequences of a frame (left image curves 0/1, right image curves 0/1, each a list
 * of integer pixel (x,y) ordered bottom-to-top) so that the left/right views of a curve cover the same rows, no curve
 * is longer than LENGTH_EACH_CURVE rows, nothing extends above the tracked end point, and both curves of an image
 * end on the same row.  pts: concatenated (x,y) int pairs; seg_off[0..4]: point offsets of [L0, L1, R0, R1].
 * Writes the kept index ranges range[4][2] (inclusive, same order) and returns 1, or returns 0 where the
 * reference returns false.  LEFT_RIGHT_OFFSET = 0, LENGTH_EACH_CURVE = 100 (common.h:108-112). */
#define ORC_LR_OFFSET 0
#define ORC_LEN_CURVE 100

int orc_cleanup_edges(const int *pts, const long long *seg_off, int reset_data_assoc, const double *tracked_endpt_y,
                      int range[4][2])
{
  const int *seg[4];
  int size[4], ystart[4], yend[4], i, s;
  for (s = 0; s < 4; ++s) {
    seg[s] = pts + 2 * seg_off[s];
    size[s] = (int)(seg_off[s + 1] - seg_off[s]);
    range[s][0] = 0;
    range[s][1] = size[s] - 1;
  }
  if (size[0] < 10 || size[1] < 10 || size[2] < 10 || size[3] < 10) return 0;
#define YAT(s, k) (seg[s][2 * (k) + 1])
  for (s = 0; s < 4; ++s) { ystart[s] = YAT(s, 0); yend[s] = YAT(s, size[s] - 1); }
  /* L = segment i, R = segment i + 2 */
  for (i = 0; i < 2; ++i) { /* :1224-1254 align the starts */
    const int L = i, R = i + 2;
    if (ystart[L] > ystart[R] - ORC_LR_OFFSET) {
      while (ystart[L] > ystart[R] - ORC_LR_OFFSET) {
        range[L][0]++;
        if (range[L][0] >= size[L]) return 0;
        ystart[L] = YAT(L, range[L][0]);
      }
    } else if (ystart[L] < ystart[R] - ORC_LR_OFFSET) {
      while (ystart[L] < ystart[R] - ORC_LR_OFFSET) {
        range[R][0]++;
        if (range[R][0] >= size[R]) return 0;
        ystart[R] = YAT(R, range[R][0]);
      }
    }
  }
  for (i = 0; i < 2; ++i) { /* :1256-1334 */
    const int L = i, R = i + 2;
    if (yend[L] < yend[R] - ORC_LR_OFFSET) {
      while (yend[L] < yend[R] - ORC_LR_OFFSET) {
        range[L][1]--;
        if (range[L][1] < 0) return 0;
        yend[L] = YAT(L, range[L][1]);
      }
    } else if (yend[L] > yend[R] - ORC_LR_OFFSET) {
      while (yend[L] > yend[R] - ORC_LR_OFFSET) {
        range[R][1]--;
        if (range[R][1] < 0) return 0;
        yend[R] = YAT(R, range[R][1]);
      }
    }
    while (yend[L] < ystart[L] - ORC_LEN_CURVE) {
      range[L][1]--;
      if (range[L][1] < 0) return 0;
      yend[L] = YAT(L, range[L][1]);
    }
    while (yend[R] < ystart[R] - ORC_LEN_CURVE) {
      range[R][1]--;
      if (range[R][1] < 0) return 0;
      yend[R] = YAT(R, range[R][1]);
    }
    while (yend[L] > tracked_endpt_y[i] && !reset_data_assoc) {
      range[L][0]++;
      range[L][1]++;
      if (range[L][0] >= size[L] || range[L][1] >= size[L]) return 0;
      yend[L] = YAT(L, range[L][1]);
    }
    while (yend[R] > tracked_endpt_y[i] + ORC_LR_OFFSET && !reset_data_assoc) {
      range[R][1]++;
      range[R][0]++;
      if (range[R][0] >= size[R] || range[R][1] >= size[R]) return 0;
      yend[R] = YAT(R, range[R][1]);
    }
  }
  for (i = 0; i < 2; ++i) { /* :1335-1392 both curves of an image end on the same row (i = 0 left, 1 right image) */
    const int A = 2 * i, B = 2 * i + 1;
    if (yend[A] > yend[B]) {
      while (yend[A] > yend[B]) {
        range[A][0]++;
        range[A][1]++;
        if (range[A][0] >= size[A] || range[A][1] >= size[A]) return 0;
        yend[A] = YAT(A, range[A][1]);
      }
    } else {
      while (yend[A] < yend[B]) {
        range[B][0]++;
        range[B][1]++;
        if (range[B][0] >= size[B] || range[B][1] >= size[B]) return 0;
        yend[B] = YAT(B, range[B][1]);
      }
    }
  }
#undef YAT
  for (s = 0; s < 4; ++s)
    if (range[s][1] - range[s][0] + 1 < 1) return 0; /* cvCreateMat of a non-positive size would throw */
  return 1;
}
