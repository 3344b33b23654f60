// cps_closest.cu -- closest point on a planar cubic for a batch of (curve, point) queries (sm_100a, FP64,
// -fmad=false), "next" row 3 of the scope table.
//
// Computes what closest_bezier_pt(p, pt, nearest_pt, out) computes (curveFittingClass.cpp:936-1043, with
// control2coeffs :737-749): the quintic whose roots are the stationary points of |c(t) - pt|^2, its roots, and the
// reference's selection among real roots in [0,1] and the end points.  The reference calls GSL's
// gsl_poly_complex_solve (not vendored, not installed here); its documented algorithm -- eigenvalues of the balanced
// companion matrix by the Francis double-shift QR iteration -- is restated.  One thread per query: a 5x5 problem has
// no internal parallelism worth a warp, and the callers (dataAssocClass.cpp:272,366) ask for many points at once.
// The CPU oracle (oracle/orc_closest.c) states the same operation sequence; results are compared bit for bit.
// Deviation from the reference: its selection loop reads one root more than the solver returns (:990 vs :975).
#include <cmath>

#include "../../include/cps_lm.h"
#include "cps_common.h"

namespace {
#define NMAX 5
__device__ __forceinline__ double dsgn(double a, double b) { return (b >= 0.0) ? fabs(a) : -fabs(a); }

/* roots of a[0] + a[1] x + ... + a[deg] x^deg, a[deg] != 0, 1 <= deg <= 5; returns 0 on success */
__device__ int poly_roots(const double *a, int deg, double *zr, double *zi)
{
  double m[NMAX][NMAX];
  int n = deg, i, j, k, l, mm, its, nn, last;
  double p = 0, q = 0, r = 0, s, t, u, v, w, x, y, z, anorm, c, f, g;
  for (i = 0; i < n; ++i)
    for (j = 0; j < n; ++j) m[i][j] = 0.0;
  for (i = 1; i < n; ++i) m[i][i - 1] = 1.0;
  for (i = 0; i < n; ++i) m[i][n - 1] = -a[i] / a[n];
  /* balancing (EISPACK balanc, radix 2) */
  last = 0;
  while (!last) {
    last = 1;
    for (i = 0; i < n; ++i) {
      r = c = 0.0;
      for (j = 0; j < n; ++j)
        if (j != i) { c += fabs(m[j][i]); r += fabs(m[i][j]); }
      if (c != 0.0 && r != 0.0) {
        g = r / 2.0; f = 1.0; s = c + r;
        while (c < g) { f *= 2.0; c *= 4.0; }
        g = r * 2.0;
        while (c > g) { f /= 2.0; c /= 4.0; }
        if ((c + r) / f < 0.95 * s) {
          last = 0;
          g = 1.0 / f;
          for (j = 0; j < n; ++j) m[i][j] *= g;
          for (j = 0; j < n; ++j) m[j][i] *= f;
        }
      }
    }
  }
  /* Francis double-shift QR on the Hessenberg matrix (EISPACK hqr) */
  anorm = 0.0;
  for (i = 0; i < n; ++i)
    for (j = (i - 1 > 0 ? i - 1 : 0); j < n; ++j) anorm += fabs(m[i][j]);
  nn = n - 1;
  t = 0.0;
  while (nn >= 0) {
    its = 0;
    do {
      for (l = nn; l >= 1; --l) {
        s = fabs(m[l - 1][l - 1]) + fabs(m[l][l]);
        if (s == 0.0) s = anorm;
        if (fabs(m[l][l - 1]) + s == s) { m[l][l - 1] = 0.0; break; }
      }
      x = m[nn][nn];
      if (l == nn) {
        zr[nn] = x + t;
        zi[nn--] = 0.0;
      } else {
        y = m[nn - 1][nn - 1];
        w = m[nn][nn - 1] * m[nn - 1][nn];
        if (l == nn - 1) {
          p = 0.5 * (y - x);
          q = p * p + w;
          z = sqrt(fabs(q));
          x += t;
          if (q >= 0.0) {
            z = p + dsgn(z, p);
            zr[nn - 1] = zr[nn] = x + z;
            if (z != 0.0) zr[nn] = x - w / z;
            zi[nn - 1] = zi[nn] = 0.0;
          } else {
            zr[nn - 1] = zr[nn] = x + p;
            zi[nn] = z;
            zi[nn - 1] = -z;
          }
          nn -= 2;
        } else {
          if (its == 60) return -1;
          if (its == 10 || its == 20) {
            t += x;
            for (i = 0; i <= nn; ++i) m[i][i] -= x;
            s = fabs(m[nn][nn - 1]) + fabs(m[nn - 1][nn - 2]);
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          for (mm = nn - 2; mm >= l; --mm) {
            z = m[mm][mm];
            r = x - z;
            s = y - z;
            p = (r * s - w) / m[mm + 1][mm] + m[mm][mm + 1];
            q = m[mm + 1][mm + 1] - z - r - s;
            r = m[mm + 2][mm + 1];
            s = fabs(p) + fabs(q) + fabs(r);
            p /= s; q /= s; r /= s;
            if (mm == l) break;
            u = fabs(m[mm][mm - 1]) * (fabs(q) + fabs(r));
            v = fabs(p) * (fabs(m[mm - 1][mm - 1]) + fabs(z) + fabs(m[mm + 1][mm + 1]));
            if (u + v == v) break;
          }
          for (i = mm + 2; i <= nn; ++i) {
            m[i][i - 2] = 0.0;
            if (i != mm + 2) m[i][i - 3] = 0.0;
          }
          for (k = mm; k <= nn - 1; ++k) {
            if (k != mm) {
              p = m[k][k - 1];
              q = m[k + 1][k - 1];
              r = 0.0;
              if (k != nn - 1) r = m[k + 2][k - 1];
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
            }
            s = dsgn(sqrt(p * p + q * q + r * r), p);
            if (s != 0.0) {
              if (k == mm) {
                if (l != mm) m[k][k - 1] = -m[k][k - 1];
              } else {
                m[k][k - 1] = -s * x;
              }
              p += s;
              x = p / s; y = q / s; z = r / s;
              q /= p; r /= p;
              for (j = k; j <= nn; ++j) {
                p = m[k][j] + q * m[k + 1][j];
                if (k != nn - 1) { p += r * m[k + 2][j]; m[k + 2][j] -= p * z; }
                m[k + 1][j] -= p * y;
                m[k][j] -= p * x;
              }
              { const int mmin = (nn < k + 3) ? nn : k + 3;
                for (i = l; i <= mmin; ++i) {
                  p = x * m[i][k] + y * m[i][k + 1];
                  if (k != nn - 1) { p += z * m[i][k + 2]; m[i][k + 2] -= p * r; }
                  m[i][k + 1] -= p * q;
                  m[i][k] -= p;
                } }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  return 0;
}

__device__ void control2coeffs(const double *cp, double *p)
{ /* curveFittingClass.cpp:737-749 */
  p[3] = cp[0];
  p[2] = 3 * cp[2] - 3 * cp[0];
  p[1] = 3 * cp[4] - 6 * cp[2] + 3 * cp[0];
  p[0] = cp[6] - 3 * cp[4] + 3 * cp[2] - cp[0];
  p[7] = cp[1];
  p[6] = 3 * cp[3] - 3 * cp[1];
  p[5] = 3 * cp[5] - 6 * cp[3] + 3 * cp[1];
  p[4] = cp[7] - 3 * cp[5] + 3 * cp[3] - cp[1];
}

// returns the number of roots found (>= 0), or -1 if the QR iteration did not converge
__device__ int closest_bezier_pt(const double *p, const double *pt, double *nearest_pt, double *out)
{
  double poly[6], zr[NMAX], zi[NMAX];
  double min_dist_sq = 10000000.0, t_best = 0.0;
  int n = 6, i, nroots;
  poly[5] = 3.0 * (p[0] * p[0] + p[4] * p[4]);
  poly[4] = 5.0 * (p[1] * p[0] + p[5] * p[4]);
  poly[3] = 2 * (2 * p[2] * p[0] + p[1] * p[1] + 2 * p[6] * p[4] + p[5] * p[5]);
  poly[2] = 3.0 * (p[2] * p[1] + p[0] * (p[3] - pt[0]) + p[6] * p[5] + p[4] * (p[7] - pt[1]));
  poly[1] = p[2] * p[2] + 2 * p[1] * (p[3] - pt[0]) + p[6] * p[6] + 2 * p[5] * (p[7] - pt[1]);
  poly[0] = p[2] * (p[3] - pt[0]) + p[6] * (p[7] - pt[1]);
  for (i = 5; i >= 0; --i) {
    if (poly[i] == 0) --n;
    else break;
  }
  nroots = n - 1;
  if (nroots >= 1 && poly_roots(poly, nroots, zr, zi) != 0) nroots = -1;
  for (i = 0; i < nroots; ++i) {
    if (zi[i] == 0.0) {
      const double t = zr[i];
      if (t >= 0.0 && t <= 1.0) {
        const double t2 = t * t, t3 = t2 * t;
        const double xval = p[0] * t3 + p[1] * t2 + p[2] * t + p[3];
        const double yval = p[4] * t3 + p[5] * t2 + p[6] * t + p[7];
        const double dist_sq = (xval - pt[0]) * (xval - pt[0]) + (yval - pt[1]) * (yval - pt[1]);
        if (dist_sq < min_dist_sq) {
          t_best = t; min_dist_sq = dist_sq; nearest_pt[0] = xval; nearest_pt[1] = yval;
        }
      } else {
        const double xval0 = p[3], yval0 = p[7];
        const double xval1 = p[0] + p[1] + p[2] + p[3], yval1 = p[4] + p[5] + p[6] + p[7];
        const double d0 = (xval0 - pt[0]) * (xval0 - pt[0]) + (yval0 - pt[1]) * (yval0 - pt[1]);
        const double d1 = (xval1 - pt[0]) * (xval1 - pt[0]) + (yval1 - pt[1]) * (yval1 - pt[1]);
        if (d0 > d1 && d1 < min_dist_sq) {
          t_best = 1.0; min_dist_sq = d1; nearest_pt[0] = xval1; nearest_pt[1] = yval1;
        } else if (d1 > d0 && d0 < min_dist_sq) {
          t_best = 0.0; min_dist_sq = d0; nearest_pt[0] = xval0; nearest_pt[1] = yval0;
        }
      }
    }
  }
  out[0] = sqrt(min_dist_sq);
  out[1] = t_best;
  return nroots;
}

__global__ void __launch_bounds__(128) closest_kernel(int Q, const double* __restrict__ cps_or_coeffs, int is_control_points,
                                                      int curve_stride, const double* __restrict__ pts,
                                                      double* __restrict__ nearest, double* __restrict__ out,
                                                      int* __restrict__ nroots) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  double p[8], np_[2] = {0.0, 0.0}, o[2];
  const double* src = cps_or_coeffs + (size_t)q * curve_stride;
  if (is_control_points) control2coeffs(src, p);
  else
    for (int i = 0; i < 8; ++i) p[i] = src[i];
  const double pt[2] = {pts[2 * (size_t)q], pts[2 * (size_t)q + 1]};
  const int nr = closest_bezier_pt(p, pt, np_, o);
  nearest[2 * (size_t)q] = np_[0];
  nearest[2 * (size_t)q + 1] = np_[1];
  out[2 * (size_t)q] = o[0];
  out[2 * (size_t)q + 1] = o[1];
  if (nroots) nroots[q] = nr;
}
}  // namespace

// Q queries.  curves: Q x 8 doubles (or 1 x 8 shared by all queries when shared_curve != 0): cubic control points
// (x0,y0,...,x3,y3) when is_control_points != 0, else the polynomial coefficients p[0..7] the reference passes.
// pts: Q x 2.  nearest: Q x 2 (left at 0 when no root qualifies; the reference leaves it untouched), out: Q x 2 =
// {distance, t}, nroots: Q (may be NULL).  Host buffers.
extern "C" int cps_closest_bezier_pt_batched(int device, int Q, const double* curves, int is_control_points,
                                             int shared_curve, const double* pts, double* nearest, double* out,
                                             int* nroots) {
  if (Q < 0 || (Q && (!curves || !pts || !nearest || !out))) return CPS_ERR_INVALID;
  if (Q == 0) return CPS_OK;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cps::set_last_error("no CUDA device: libcps has no CPU path");
    return CPS_ERR_CUDA;
  }
  CPS_CUDA(cudaSetDevice(device));
  const size_t ncur = shared_curve ? 1 : (size_t)Q;
  double *d_c = nullptr, *d_p = nullptr, *d_n = nullptr, *d_o = nullptr;
  int* d_r = nullptr;
  auto cleanup = [&]() { cudaFree(d_c); cudaFree(d_p); cudaFree(d_n); cudaFree(d_o); cudaFree(d_r); };
  auto run = [&]() -> int {
    CPS_CUDA(cudaMalloc(&d_c, sizeof(double) * 8 * ncur));
    CPS_CUDA(cudaMalloc(&d_p, sizeof(double) * 2 * (size_t)Q));
    CPS_CUDA(cudaMalloc(&d_n, sizeof(double) * 2 * (size_t)Q));
    CPS_CUDA(cudaMalloc(&d_o, sizeof(double) * 2 * (size_t)Q));
    CPS_CUDA(cudaMalloc(&d_r, sizeof(int) * (size_t)Q));
    CPS_CUDA(cudaMemcpy(d_c, curves, sizeof(double) * 8 * ncur, cudaMemcpyHostToDevice));
    CPS_CUDA(cudaMemcpy(d_p, pts, sizeof(double) * 2 * (size_t)Q, cudaMemcpyHostToDevice));
    closest_kernel<<<(Q + 127) / 128, 128>>>(Q, d_c, is_control_points, shared_curve ? 0 : 8, d_p, d_n, d_o, d_r);
    CPS_CUDA(cudaGetLastError());
    CPS_CUDA(cudaMemcpy(nearest, d_n, sizeof(double) * 2 * (size_t)Q, cudaMemcpyDeviceToHost));
    CPS_CUDA(cudaMemcpy(out, d_o, sizeof(double) * 2 * (size_t)Q, cudaMemcpyDeviceToHost));
    if (nroots) CPS_CUDA(cudaMemcpy(nroots, d_r, sizeof(int) * (size_t)Q, cudaMemcpyDeviceToHost));
    return CPS_OK;
  };
  const int rc = run();
  cleanup();
  return rc;
}
