/*
 * orc_closest.c -- CPU oracle of the closest point on a planar cubic ("next" row 3).  TEST INFRASTRUCTURE ONLY.
 *
 * Restates closest_bezier_pt(p, pt, nearest_pt, out), curveFittingClass.cpp:936-1043, and control2coeffs :737-749:
 * the stationary points of |c(t) - pt|^2 are the roots of a quintic (:945-950); every real root in [0,1] is a
 * candidate, a real root outside [0,1] brings the nearer end point in (with the reference's strict comparisons), and
 * the closest candidate wins; out = {distance, t}.
 * The reference finds the roots with GSL's gsl_poly_complex_solve (GNU Scientific Library, version unpinned -- it is
 * included from /usr/include/gsl/gsl_poly.h, curveFittingClass.cpp:11, and is neither vendored under /root/reference
 * nor installed here).  Its documented algorithm is restated: eigenvalues of the balanced companion matrix by the
 * Francis double-shift QR iteration (EISPACK balanc + hqr).  PARITY UNPINNED: no reference vector exists for this
 * routine and GSL itself cannot be run here; the tests check the roots against numpy.roots and the selected point
 * against a dense scan of t, and the CUDA kernel against this file bit for bit.
 * Deviation: the reference's selection loop runs over n entries of a roots array that holds n-1 roots (:990 vs :975),
 * i.e. it reads one uninitialised root; here exactly the n-1 roots are examined.
 */
#include <math.h>

#define NMAX 5

static double sgn(double a, double b) { return (b >= 0.0) ? fabs(a) : -fabs(a); }

/* roots of a[0] + a[1] x + ... + a[deg] x^deg, a[deg] != 0, 1 <= deg <= 5; returns 0 on success */
int orc_poly_roots(const double *a, int deg, double *zr, double *zi)
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
            z = p + sgn(z, p);
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
            s = sgn(sqrt(p * p + q * q + r * r), p);
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

void orc_control2coeffs(const double *cp, double *p)
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

/* returns the number of roots found (>= 0), or -1 if the QR iteration did not converge */
int orc_closest_bezier_pt(const double *p, const double *pt, double *nearest_pt, double *out)
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
  if (nroots >= 1 && orc_poly_roots(poly, nroots, zr, zi) != 0) nroots = -1;
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
