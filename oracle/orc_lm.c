/*
 * orc_lm.c -- CPU oracle: Levenberg-Marquardt (analytic-Jacobian and secant variants) with the
 * reference's LU normal-equation solve.  TEST INFRASTRUCTURE ONLY (see orc_lm.h).
 *
 * Written from the algorithm description in SURVEY.md (A.1, A.2) while reading the reference;
 * every function names the reference lines whose floating-point operation ORDER it reproduces,
 * because the secant variant is chaotic enough that a different summation order changes the
 * iteration count.  Compile with -ffp-contract=off (no FMA) so that each '*' and '+' rounds
 * separately, exactly like the reference built with plain gcc on x86-64.
 */
#include "orc_lm.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_BLK 32                 /* reference: levmar-2.5/misc.h:48  (__BLOCKSZ__)       */
#define ORC_BLK_SQ (ORC_BLK * ORC_BLK)
#define ORC_EPSILON 1E-12          /* reference: levmar-2.5/lm.c:35                        */
#define ORC_ONE_THIRD 0.3333333334 /* reference: levmar-2.5/lm.c:36 (sic)                  */
#define ORC_INIT_MU 1E-03          /* reference: levmar-2.5/levmar.h:87-93                 */
#define ORC_STOP_THRESH 1E-17
#define ORC_DIFF_DELTA 1E-06

static double orc_abs(double v) { return (v >= 0.0) ? v : -v; } /* levmar.h FABS macro */

/* e = x - y (x may be NULL meaning zero) and the squared 2-norm of e.
 * Order follows levmar-2.5/misc_core.c:712-798: four running sums; full groups of eight are
 * visited from the top index downwards, element (top-k) feeding sum[k mod 4]; the ragged tail is
 * visited upwards, element number c-from-the-end feeding sum[(7-c) mod 4]; result
 * ((s0+s1)+s2)+s3.  e may alias y. */
double orc_residual_sqnorm(double *e, const double *x, const double *y, int n)
{
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  const int full = (n >> 3) << 3;
  int top, k;

  for (top = full - 1; top > 0; top -= 8) {
    for (k = 0; k < 8; ++k) {
      const int idx = top - k;
      const double d = x ? (x[idx] - y[idx]) : (-y[idx]);
      e[idx] = d;
      s[k & 3] += d * d;
    }
  }
  {
    const int rem = n - full;
    for (k = 0; k < rem; ++k) {
      const int idx = full + k;
      const int c = rem - k; /* the reference's switch label for this element */
      const double d = x ? (x[idx] - y[idx]) : (-y[idx]);
      e[idx] = d;
      s[(7 - c) & 3] += d * d;
    }
  }
  return s[0] + s[1] + s[2] + s[3];
}

/* b = a^T a for row-major n x m a.  Order follows levmar-2.5/misc_core.c:101-126: the upper
 * triangle in column panels of 32; inside a panel every entry is the running sum, over row
 * blocks of 32 taken in ascending order, of a per-block dot product that itself starts at 0 and
 * runs over ascending rows; then the triangle is mirrored. */
void orc_ata_blocked(const double *a, double *b, int n, int m)
{
  int jj, kk, i, j, k;
  for (jj = 0; jj < m; jj += ORC_BLK) {
    const int jend = (jj + ORC_BLK <= m) ? jj + ORC_BLK : m;
    for (i = 0; i < m; ++i)
      for (j = (jj >= i ? jj : i); j < jend; ++j) b[i * m + j] = 0.0;
    for (kk = 0; kk < n; kk += ORC_BLK) {
      const int kend = (kk + ORC_BLK <= n) ? kk + ORC_BLK : n;
      for (i = 0; i < m; ++i)
        for (j = (jj >= i ? jj : i); j < jend; ++j) {
          double part = 0.0;
          for (k = kk; k < kend; ++k) part += a[k * m + i] * a[k * m + j];
          b[i * m + j] += part;
        }
    }
  }
  for (i = 0; i < m; ++i)
    for (j = 0; j < i; ++j) b[i * m + j] = b[j * m + i];
}

/* Solve A x = B (A m x m row-major; A and B untouched).  Returns 1 on success, 0 if some row of
 * A is entirely zero.  Follows levmar-2.5/Axb_core.c:1044-1132 (dAx_eq_b_LU_noLapack): Crout
 * LU with implicit row scaling; the pivot search uses ">=" so the LAST row among equal scaled
 * magnitudes wins; an exactly-zero pivot is replaced by DBL_EPSILON; forward substitution skips
 * leading zeros of the permuted right-hand side.  Unlike the reference (static buffer under
 * LINSOLVERS_RETAIN_MEMORY, levmar.h:37-39) this version is re-entrant. */
int orc_lu_solve(const double *A, const double *B, double *x, int m)
{
  double *a = (double *)malloc(sizeof(double) * (size_t)(m * m + m));
  int *perm = (int *)malloc(sizeof(int) * (size_t)m);
  double *scale = a + m * m;
  int i, j, k, piv = -1, ok = 1;
  if (!a || !perm) { free(a); free(perm); return 0; }

  memcpy(a, A, sizeof(double) * (size_t)(m * m));
  memcpy(x, B, sizeof(double) * (size_t)m);

  for (i = 0; i < m && ok; ++i) {
    double big = 0.0;
    for (j = 0; j < m; ++j) {
      const double v = orc_abs(a[i * m + j]);
      if (v > big) big = v;
    }
    if (big == 0.0) ok = 0;
    else scale[i] = 1.0 / big;
  }
  if (!ok) { free(a); free(perm); return 0; }

  for (j = 0; j < m; ++j) {
    double big = 0.0;
    for (i = 0; i < j; ++i) {
      double acc = a[i * m + j];
      for (k = 0; k < i; ++k) acc -= a[i * m + k] * a[k * m + j];
      a[i * m + j] = acc;
    }
    for (i = j; i < m; ++i) {
      double acc = a[i * m + j], merit;
      for (k = 0; k < j; ++k) acc -= a[i * m + k] * a[k * m + j];
      a[i * m + j] = acc;
      merit = scale[i] * orc_abs(acc);
      if (merit >= big) { big = merit; piv = i; }
    }
    if (j != piv) {
      for (k = 0; k < m; ++k) {
        const double tmp = a[piv * m + k];
        a[piv * m + k] = a[j * m + k];
        a[j * m + k] = tmp;
      }
      scale[piv] = scale[j];
    }
    perm[j] = piv;
    if (a[j * m + j] == 0.0) a[j * m + j] = DBL_EPSILON;
    if (j != m - 1) {
      const double inv = 1.0 / a[j * m + j];
      for (i = j + 1; i < m; ++i) a[i * m + j] *= inv;
    }
  }

  {
    int first_nz = 0; /* 1-based index of the first non-zero rhs entry seen, 0 = none yet */
    for (i = 0; i < m; ++i) {
      double acc;
      const int src = perm[i];
      acc = x[src];
      x[src] = x[i];
      if (first_nz != 0) {
        for (j = first_nz - 1; j < i; ++j) acc -= a[i * m + j] * x[j];
      } else if (acc != 0.0) {
        first_nz = i + 1;
      }
      x[i] = acc;
    }
  }
  for (i = m - 1; i >= 0; --i) {
    double acc = x[i];
    for (j = i + 1; j < m; ++j) acc -= a[i * m + j] * x[j];
    x[i] = acc / a[i * m + i];
  }
  free(a);
  free(perm);
  return 1;
}

/* Forward-difference Jacobian, levmar-2.5/misc_core.c:135-170: step max(|1e-4 p_j|, delta),
 * column j = (f(p + d e_j) - f(p)) * (1/d). */
void orc_fdjac_forward(orc_func_t func, double *p, const double *hx, double *hxx, double delta,
                       double *jac, int m, int n, void *adata)
{
  int i, j;
  for (j = 0; j < m; ++j) {
    double d = 1E-04 * p[j], keep, inv;
    d = orc_abs(d);
    if (d < delta) d = delta;
    keep = p[j];
    p[j] += d;
    func(p, hxx, m, n, adata);
    p[j] = keep;
    inv = 1.0 / d;
    for (i = 0; i < n; ++i) jac[i * m + j] = (hxx[i] - hx[i]) * inv;
  }
}

/* Central-difference Jacobian, levmar-2.5/misc_core.c:173-209. */
void orc_fdjac_central(orc_func_t func, double *p, double *hxm, double *hxp, double delta,
                       double *jac, int m, int n, void *adata)
{
  int i, j;
  for (j = 0; j < m; ++j) {
    double d = 1E-04 * p[j], keep, half_inv;
    d = orc_abs(d);
    if (d < delta) d = delta;
    keep = p[j];
    p[j] -= d;
    func(p, hxm, m, n, adata);
    p[j] = keep + d;
    func(p, hxp, m, n, adata);
    p[j] = keep;
    half_inv = 0.5 / d;
    for (i = 0; i < n; ++i) jac[i * m + j] = (hxp[i] - hxm[i]) * half_inv;
  }
}

/* J^T J and J^T e.  "small" selects the reference's unblocked path (descending rows, lower
 * triangle then mirrored; lm_core.c:213-237 / :604-628), otherwise the blocked product plus an
 * ascending-row J^T e (lm_core.c:243-254 / :633-644). */
static void orc_normal_eqs(const double *jac, const double *e, double *jtj, double *jte, int m,
                           int n, int small)
{
  int i, j, l;
  if (small) {
    for (i = 0; i < m * m; ++i) jtj[i] = 0.0;
    for (i = 0; i < m; ++i) jte[i] = 0.0;
    for (l = n - 1; l >= 0; --l) {
      const double *row = jac + l * m;
      for (i = m - 1; i >= 0; --i) {
        const double alpha = row[i];
        for (j = i; j >= 0; --j) jtj[i * m + j] += row[j] * alpha;
        jte[i] += alpha * e[l];
      }
    }
    for (i = 0; i < m; ++i)
      for (j = i + 1; j < m; ++j) jtj[i * m + j] = jtj[j * m + i];
  } else {
    orc_ata_blocked(jac, jtj, n, m);
    for (i = 0; i < m; ++i) jte[i] = 0.0;
    for (l = 0; l < n; ++l) {
      const double *row = jac + l * m;
      const double el = e[l];
      for (i = 0; i < m; ++i) jte[i] += row[i] * el;
    }
  }
}

static void orc_fill_info(double *info, double e0, double e_now, double g_inf, double dp2,
                          double mu, const double *jtj, int m, int k, int stop, int nfev, int njev,
                          int nlss)
{
  int i;
  double big = -DBL_MAX;
  if (!info) return;
  for (i = 0; i < m; ++i)
    if (big < jtj[i * m + i]) big = jtj[i * m + i];
  info[0] = e0;
  info[1] = e_now;
  info[2] = g_inf;
  info[3] = dp2;
  info[4] = mu / big;
  info[5] = (double)k;
  info[6] = (double)stop;
  info[7] = (double)nfev;
  info[8] = (double)njev;
  info[9] = (double)nlss;
}

/* LM with analytic Jacobian.  Control flow and arithmetic follow levmar-2.5/lm_core.c:60-423
 * (dlevmar_der): Jacobian at the top of each outer iteration, inner retry loop on rejected
 * steps, Nielsen damping, stop reasons 1..7, return value = iterations or -1 for reasons 4/7. */
int orc_lm_der(orc_func_t func, orc_func_t jacf, double *p, const double *x, int m, int n,
               int itmax, const double *opts, double *info, void *adata, orc_lm_extra *extra)
{
  double tau, eps1, eps2, eps2_sq, eps3;
  double mu = 0.0, g_inf = 0.0, p_sq = 0.0, dp_sq = DBL_MAX, e_sq, e_sq_new, e_sq0, dF, dL, tmp;
  double *buf, *e, *hx, *jte, *jac, *jtj, *dp, *diag, *pdp;
  int k, i, nu = 2, nu2, stop = 0, nfev, njev = 0, nlss = 0, solved;
  const int nm = n * m;

  if (extra) { extra->n_jtj = 0; extra->n_broyden = 0; }
  if (n < m || !jacf) return ORC_LM_ERROR; /* lm_core.c:117-126 */

  if (opts) {
    tau = opts[0]; eps1 = opts[1]; eps2 = opts[2]; eps2_sq = opts[2] * opts[2]; eps3 = opts[3];
  } else {
    tau = ORC_INIT_MU; eps1 = ORC_STOP_THRESH; eps2 = ORC_STOP_THRESH;
    eps2_sq = ORC_STOP_THRESH * ORC_STOP_THRESH; eps3 = ORC_STOP_THRESH;
  }

  buf = (double *)malloc(sizeof(double) * (size_t)(2 * n + 4 * m + nm + m * m));
  if (!buf) return ORC_LM_ERROR;
  e = buf; hx = e + n; jte = hx + n; jac = jte + m; jtj = jac + nm; dp = jtj + m * m;
  diag = dp + m; pdp = diag + m;

  func(p, hx, m, n, adata); nfev = 1;
  e_sq = orc_residual_sqnorm(e, x, hx, n);
  e_sq0 = e_sq;
  if (!isfinite(e_sq)) stop = 7;

  for (k = 0; k < itmax && !stop; ++k) {
    if (e_sq <= eps3) { stop = 6; break; }

    jacf(p, jac, m, n, adata); ++njev;
    orc_normal_eqs(jac, e, jtj, jte, m, n, nm < ORC_BLK_SQ); /* strict "<" here, lm_core.c:193 */
    if (extra) extra->n_jtj++;

    for (i = 0, p_sq = g_inf = 0.0; i < m; ++i) {
      tmp = orc_abs(jte[i]);
      if (g_inf < tmp) g_inf = tmp;
      diag[i] = jtj[i * m + i];
      p_sq += p[i] * p[i];
    }
    if (g_inf <= eps1) { dp_sq = 0.0; stop = 1; break; }

    if (k == 0) {
      for (i = 0, tmp = -DBL_MAX; i < m; ++i)
        if (diag[i] > tmp) tmp = diag[i];
      mu = tau * tmp;
    }

    for (;;) {
      for (i = 0; i < m; ++i) jtj[i * m + i] += mu;
      solved = orc_lu_solve(jtj, jte, dp, m); ++nlss;
      if (solved) {
        for (i = 0, dp_sq = 0.0; i < m; ++i) {
          tmp = dp[i];
          pdp[i] = p[i] + tmp;
          dp_sq += tmp * tmp;
        }
        if (dp_sq <= eps2_sq * p_sq) { stop = 2; break; }
        if (dp_sq >= (p_sq + eps2) / (ORC_EPSILON * ORC_EPSILON)) { stop = 4; break; }

        func(pdp, hx, m, n, adata); ++nfev;
        e_sq_new = orc_residual_sqnorm(hx, x, hx, n);
        if (!isfinite(e_sq_new)) { stop = 7; break; }

        for (i = 0, dL = 0.0; i < m; ++i) dL += dp[i] * (mu * dp[i] + jte[i]);
        dF = e_sq - e_sq_new;

        if (dL > 0.0 && dF > 0.0) {
          tmp = (2.0 * dF / dL - 1.0);
          tmp = 1.0 - tmp * tmp * tmp;
          mu = mu * ((tmp >= ORC_ONE_THIRD) ? tmp : ORC_ONE_THIRD);
          nu = 2;
          for (i = 0; i < m; ++i) p[i] = pdp[i];
          for (i = 0; i < n; ++i) e[i] = hx[i];
          e_sq = e_sq_new;
          break;
        }
      }
      mu *= nu;
      nu2 = nu << 1;
      if (nu2 <= nu) { stop = 5; break; }
      nu = nu2;
      for (i = 0; i < m; ++i) jtj[i * m + i] = diag[i];
    }
  }
  if (k >= itmax) stop = 3;
  for (i = 0; i < m; ++i) jtj[i * m + i] = diag[i];

  orc_fill_info(info, e_sq0, e_sq, g_inf, dp_sq, mu, jtj, m, k, stop, nfev, njev, nlss);
  free(buf);
  return (stop != 4 && stop != 7) ? k : ORC_LM_ERROR;
}

/* Secant LM (finite-difference Jacobian refreshed on demand, Broyden rank-1 updates in
 * between).  Follows levmar-2.5/lm_core.c:429-828 (dlevmar_dif): no inner retry loop -- every
 * linear solve is one outer iteration; J is re-differenced when (an accepted step happened and
 * nu > 16) or after K = max(m,10) secant updates; the secant update runs when (updp || dF > 0)
 * BEFORE the accept test; J^T J is rebuilt only when J changed. */
int orc_lm_dif(orc_func_t func, double *p, const double *x, int m, int n, int itmax,
               const double *opts, double *info, void *adata, orc_lm_extra *extra)
{
  double tau, eps1, eps2, eps2_sq, eps3, delta;
  double mu = 0.0, g_inf = 0.0, p_sq = 0.0, dp_sq = DBL_MAX, e_sq, e_sq_new, e_sq0, dF, dL, tmp;
  double *buf, *e, *hx, *jte, *jac, *jtj, *dp, *diag, *pdp, *wrk, *wrk2;
  int k, i, j, l, nu, nu2, stop = 0, nfev, njap = 0, nlss = 0, solved;
  int updjac = 0, updp = 1, newjac = 0, forward = 1;
  const int K = (m >= 10) ? m : 10;
  const int nm = n * m;

  if (extra) { extra->n_jtj = 0; extra->n_broyden = 0; }
  if (n < m) return ORC_LM_ERROR; /* lm_core.c:493-496 */

  if (opts) {
    tau = opts[0]; eps1 = opts[1]; eps2 = opts[2]; eps2_sq = opts[2] * opts[2]; eps3 = opts[3];
    delta = opts[4];
    if (delta < 0.0) { delta = -delta; forward = 0; }
  } else {
    tau = ORC_INIT_MU; eps1 = ORC_STOP_THRESH; eps2 = ORC_STOP_THRESH;
    eps2_sq = ORC_STOP_THRESH * ORC_STOP_THRESH; eps3 = ORC_STOP_THRESH; delta = ORC_DIFF_DELTA;
  }

  buf = (double *)malloc(sizeof(double) * (size_t)(4 * n + 4 * m + nm + m * m));
  if (!buf) return ORC_LM_ERROR;
  e = buf; hx = e + n; jte = hx + n; jac = jte + m; jtj = jac + nm; dp = jtj + m * m;
  diag = dp + m; pdp = diag + m; wrk = pdp + m; wrk2 = wrk + n;

  func(p, hx, m, n, adata); nfev = 1;
  e_sq = orc_residual_sqnorm(e, x, hx, n);
  e_sq0 = e_sq;
  if (!isfinite(e_sq)) stop = 7;

  nu = 20; /* forces a Jacobian on the first pass, lm_core.c:555 */

  for (k = 0; k < itmax && !stop; ++k) {
    if (e_sq <= eps3) { stop = 6; break; }

    if ((updp && nu > 16) || updjac == K) {
      if (forward) {
        orc_fdjac_forward(func, p, hx, wrk, delta, jac, m, n, adata);
        ++njap; nfev += m;
      } else {
        orc_fdjac_central(func, p, wrk, wrk2, delta, jac, m, n, adata);
        ++njap; nfev += 2 * m;
      }
      nu = 2; updjac = 0; updp = 0; newjac = 1;
    }

    if (newjac) {
      newjac = 0;
      orc_normal_eqs(jac, e, jtj, jte, m, n, nm <= ORC_BLK_SQ); /* "<=" here, lm_core.c:585 */
      if (extra) extra->n_jtj++;
      for (i = 0, p_sq = g_inf = 0.0; i < m; ++i) {
        tmp = orc_abs(jte[i]);
        if (g_inf < tmp) g_inf = tmp;
        diag[i] = jtj[i * m + i];
        p_sq += p[i] * p[i];
      }
    }

    if (g_inf <= eps1) { dp_sq = 0.0; stop = 1; break; }

    if (k == 0) {
      for (i = 0, tmp = -DBL_MAX; i < m; ++i)
        if (diag[i] > tmp) tmp = diag[i];
      mu = tau * tmp;
    }

    for (i = 0; i < m; ++i) jtj[i * m + i] += mu;
    solved = orc_lu_solve(jtj, jte, dp, m); ++nlss;

    if (solved) {
      for (i = 0, dp_sq = 0.0; i < m; ++i) {
        tmp = dp[i];
        pdp[i] = p[i] + tmp;
        dp_sq += tmp * tmp;
      }
      if (dp_sq <= eps2_sq * p_sq) { stop = 2; break; }
      if (dp_sq >= (p_sq + eps2) / (ORC_EPSILON * ORC_EPSILON)) { stop = 4; break; }

      func(pdp, wrk, m, n, adata); ++nfev;
      e_sq_new = orc_residual_sqnorm(wrk2, x, wrk, n);
      if (!isfinite(e_sq_new)) { stop = 7; break; }

      dF = e_sq - e_sq_new;
      if (updp || dF > 0) { /* rank-1 secant update of J, lm_core.c:745-755 */
        for (i = 0; i < n; ++i) {
          for (l = 0, tmp = 0.0; l < m; ++l) tmp += jac[i * m + l] * dp[l];
          tmp = (wrk[i] - hx[i] - tmp) / dp_sq;
          for (j = 0; j < m; ++j) jac[i * m + j] += tmp * dp[j];
        }
        ++updjac;
        newjac = 1;
        if (extra) extra->n_broyden++;
      }

      for (i = 0, dL = 0.0; i < m; ++i) dL += dp[i] * (mu * dp[i] + jte[i]);

      if (dL > 0.0 && dF > 0.0) {
        tmp = (2.0 * dF / dL - 1.0);
        tmp = 1.0 - tmp * tmp * tmp;
        mu = mu * ((tmp >= ORC_ONE_THIRD) ? tmp : ORC_ONE_THIRD);
        nu = 2;
        for (i = 0; i < m; ++i) p[i] = pdp[i];
        for (i = 0; i < n; ++i) { e[i] = wrk2[i]; hx[i] = wrk[i]; }
        e_sq = e_sq_new;
        updp = 1;
        continue;
      }
    }

    mu *= nu;
    nu2 = nu << 1;
    if (nu2 <= nu) { stop = 5; break; }
    nu = nu2;
    for (i = 0; i < m; ++i) jtj[i * m + i] = diag[i];
  }
  if (k >= itmax) stop = 3;
  for (i = 0; i < m; ++i) jtj[i * m + i] = diag[i];

  orc_fill_info(info, e_sq0, e_sq, g_inf, dp_sq, mu, jtj, m, k, stop, nfev, njap, nlss);
  free(buf);
  return (stop != 4 && stop != 7) ? k : ORC_LM_ERROR;
}
