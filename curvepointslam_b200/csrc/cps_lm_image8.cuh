// cps_lm_image8.cuh -- dlevmar_der for large batches of per-contour image-space fits (m = 8: the Bernstein
// combination of four (x, y) control points, curveFittingClass.cpp:1057-1073), a thread per fit.
//
// The m = 8 problem is small enough for one thread to carry a whole fit from the first residual to info[]: eight
// parameters, a Jacobian that depends only on the sample parameters t (row u of a sample is the four Bernstein
// weights in the even columns, row v the same weights in the odd columns), an 8 x 8 solve.  So there are no rounds
// and no lists as in cps_lm_staged.cuh: one kernel, every thread runs dlevmar_der (lm_core.c:60-423) for its fit and
// takes the next one.  The contour samples come through the same per-lane cp.async staging; the 8 x 8 work matrix
// lives in shared memory [element][thread] and goes through the same LU as the m = 19 solve (lu_factor_solve).
//
// Same operation sequence as the monolithic kernel, hence bit-identical results:
//   * J^T J: per entry a partial sum per 32-row block, blocks added in ascending order (misc_core.c:101-126), or the
//     unblocked descending-row loop when n*m < 1024 (lm_core.c:213-237).  An entry (2k, 2k') only receives the u rows
//     and (2k+1, 2k'+1) only the v rows; both see the same products w_k * w_k' in the same order, so the two 4 x 4
//     blocks are bit-identical and only one is accumulated; the (even, odd) entries are structural zeros.  J does not
//     depend on p: the sums are formed once per fit (every later Jacobian evaluation of the reference reproduces the
//     same numbers) while njev still counts every evaluation.
//   * J^T e: one running sum per column over ascending rows (lm_core.c:246-254), descending in the unblocked case.
//   * ||e||^2: the four chains of dlevmar_L2nrmxmy (misc_core.c:712-798).
#pragma once
#include "cps_lm_staged.cuh"

namespace cps {

constexpr int kI8Threads = 128;
constexpr int kI8LuDoubles = 8 * 8 + 8 + 8;  // per thread: work matrix, row scales, right-hand side

// staging tile of a warp: 8 samples (16 values of x, 8 of t) per lane and stage
struct EvalTiles {
  double x[32][18];
  double t[32][10];
};

struct I8Fit {
  const double* x;   // samples of the fit
  const double* tp;  // t of every sample
  int n, nsamp;
  int xoff, toff;
};

// ||x - f(p)||^2 over the fit; `derive` additionally computes t from the image rows (fit_curve:451-511) into tile and
// work area
template <bool DERIVE>
__device__ __forceinline__ double i8_eval_pass(const I8Fit& F, const double (&p)[8], EvalTiles* wt, int lane,
                                               double firsty, double spany) {
  double* xrow = &wt->x[lane][0];
  double* trow = &wt->t[lane][0];
  const double* xt = xrow + F.xoff;
  double* tt = trow + F.toff;
  auto resid = [&](double tval, double xu, double xv, double& eu, double& ev) {
    double w[4];
    bernstein3(tval, w);
    double au = 0.0, av = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      au += w[k] * p[2 * k];
      av += w[k] * p[2 * k + 1];
    }
    eu = xu - au;
    ev = xv - av;
  };
  const int full = (F.n >> 3) << 3;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const int nstage = (F.n + 15) >> 4;  // stages of 8 samples, walked downwards
#pragma unroll 1
  for (int b = nstage - 1; b >= 0; --b) {
    row_fill(xrow, F.xoff, F.x + 16 * b, min(16, F.n - 16 * b));
    if (!DERIVE) row_fill(trow, F.toff, F.tp + 8 * b, min(8, F.nsamp - 8 * b));
    if (b > 0) {
      prefetch_l2(F.x + 16 * (b - 1));
      if (!DERIVE) prefetch_l2(F.tp + 8 * (b - 1));
    }
    cp_async_wait_all();
    if (DERIVE) {
      const int s_hi = min(8 * b + 8, F.nsamp);
      for (int s = 8 * b; s < s_hi; ++s) {
        double v = (xt[2 * (s & 7) + 1] - firsty) / spany;
        v = (v < 0.0) ? 0.0 : v;
        v = (v > 1.0) ? 1.0 : v;
        tt[s & 7] = v;
      }
    }
    const int g_hi = min(2 * b + 1, (full >> 3) - 1);
#pragma unroll 1
    for (int gq = g_hi; gq >= 2 * b; --gq) {
      const int l = (4 * gq) & 7;
      double eu0, ev0, eu1, ev1, eu2, ev2, eu3, ev3;
      resid(tt[l + 3], xt[2 * l + 6], xt[2 * l + 7], eu3, ev3);
      resid(tt[l + 2], xt[2 * l + 4], xt[2 * l + 5], eu2, ev2);
      resid(tt[l + 1], xt[2 * l + 2], xt[2 * l + 3], eu1, ev1);
      resid(tt[l], xt[2 * l], xt[2 * l + 1], eu0, ev0);
      s0 += ev3 * ev3;
      s1 += eu3 * eu3;
      s2 += ev2 * ev2;
      s3 += eu2 * eu2;
      s0 += ev1 * ev1;
      s1 += eu1 * eu1;
      s2 += ev0 * ev0;
      s3 += eu0 * eu0;
    }
    if (DERIVE) {  // publish the stage's t (the work area's rows start on even doubles: toff = 0)
      double* tw = const_cast<double*>(F.tp) + 8 * b;
      const int cnt = min(8, F.nsamp - 8 * b);
      int i = 0;
      for (; i + 2 <= cnt; i += 2) *reinterpret_cast<double2*>(tw + i) = make_double2(tt[i], tt[i + 1]);
      if (i < cnt) tw[i] = tt[i];
    }
  }
  {
    const int rem = F.n - full;  // 0, 2, 4 or 6 trailing values: at most three samples, read directly
    for (int k = 0; k < rem; k += 2) {
      const int s = (full + k) >> 1;
      double eu, ev;
      resid(F.tp[s], F.x[2 * s], F.x[2 * s + 1], eu, ev);
      const int cu = (7 - (rem - k)) & 3, cv = (7 - (rem - k - 1)) & 3;
      const double du = eu * eu, dv = ev * ev;
      if (cu == 0) s0 += du; else if (cu == 1) s1 += du; else if (cu == 2) s2 += du; else s3 += du;
      if (cv == 0) s0 += dv; else if (cv == 1) s1 += dv; else if (cv == 2) s2 += dv; else s3 += dv;
    }
  }
  return ((s0 + s1) + s2) + s3;
}

// J^T e at p (and, the first time, the 10 distinct entries of J^T J): blocked form, ascending rows
template <bool WITH_JTJ>
__device__ __forceinline__ void i8_jac_pass(const I8Fit& F, const double (&p)[8], EvalTiles* wt, int lane,
                                            double (&uu)[10], double (&jte)[8]) {
  double* xrow = &wt->x[lane][0];
  double* trow = &wt->t[lane][0];
  const double* xt = xrow + F.xoff;
  const double* tt = trow + F.toff;
#pragma unroll
  for (int q = 0; q < 8; ++q) jte[q] = 0.0;
  double part[10];
  if (WITH_JTJ) {
#pragma unroll
    for (int q = 0; q < 10; ++q) { uu[q] = 0.0; part[q] = 0.0; }
  }
  const int nstage = (F.n + 15) >> 4;
#pragma unroll 1
  for (int b = 0; b < nstage; ++b) {
    row_fill(xrow, F.xoff, F.x + 16 * b, min(16, F.n - 16 * b));
    row_fill(trow, F.toff, F.tp + 8 * b, min(8, F.nsamp - 8 * b));
    if (b + 1 < nstage) {
      prefetch_l2(F.x + 16 * (b + 1));
      prefetch_l2(F.tp + 8 * (b + 1));
    }
    cp_async_wait_all();
    const int s_hi = min(8 * b + 8, F.nsamp);
#pragma unroll 1
    for (int s = 8 * b; s < s_hi; ++s) {
      const int l = s & 7;
      double w[4];
      bernstein3(tt[l], w);
      double au = 0.0, av = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        au += w[k] * p[2 * k];
        av += w[k] * p[2 * k + 1];
      }
      const double eu = xt[2 * l] - au, ev = xt[2 * l + 1] - av;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        jte[2 * k] += w[k] * eu;
        jte[2 * k + 1] += w[k] * ev;
      }
      if (WITH_JTJ) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c <= a; ++c) part[a * (a + 1) / 2 + c] += w[c] * w[a];
      }
    }
    if (WITH_JTJ && ((b & 1) || b + 1 == nstage)) {  // a 32-row block = two stages: its partial sums join the entries
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        uu[q] += part[q];
        part[q] = 0.0;
      }
    }
  }
}

// the reference's unblocked loop for small problems (n*m < 1024: fewer than 64 points): descending rows, direct sums
template <bool WITH_JTJ>
__device__ __noinline__ void i8_jac_unblocked(const I8Fit& F, const double (&p)[8], double (&uu)[10], double (&jte)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) jte[q] = 0.0;
  if (WITH_JTJ) {
#pragma unroll
    for (int q = 0; q < 10; ++q) uu[q] = 0.0;
  }
  for (int s = F.nsamp - 1; s >= 0; --s) {
    double w[4];
    bernstein3(F.tp[s], w);
    double au = 0.0, av = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      au += w[k] * p[2 * k];
      av += w[k] * p[2 * k + 1];
    }
    const double eu = F.x[2 * s] - au, ev = F.x[2 * s + 1] - av;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      jte[2 * k] += w[k] * eu;
      jte[2 * k + 1] += w[k] * ev;
    }
    if (WITH_JTJ) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c <= a; ++c) uu[a * (a + 1) / 2 + c] += w[c] * w[a];
    }
  }
}

__global__ void __launch_bounds__(kI8Threads, 2) image8_der_kernel(const LmLaunch L, double* tws) {
  constexpr int M = 8, T = kI8Threads;
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double* a = smem + tid;
  double* sc = smem + M * M * T + tid;
  double* bx = smem + (M * M + M) * T + tid;
  EvalTiles* wt = reinterpret_cast<EvalTiles*>(smem + kI8LuDoubles * T) + wid;
  for (unsigned int f = blockIdx.x * T + tid; f < (unsigned int)L.B; f += gridDim.x * T) {
    const long long o0 = L.off[f], nn = L.off[f + 1] - o0;
    int ret_code = 0;
    if (nn < M) ret_code = -1;  // lm_core.c:117
    else if (nn > L.n_max || (nn & 1)) ret_code = -2;
    else if (L.counts[f] != (int)(nn >> 1)) ret_code = -3;
    if (ret_code != 0) {
      for (int q = 0; q < 10; ++q) L.info[(size_t)f * 10 + q] = 0.0;
      L.ret[f] = ret_code;
      if (L.extra) { L.extra[2 * f] = 0; L.extra[2 * f + 1] = 0; }
      continue;
    }
    I8Fit F;
    F.x = L.meas + o0;
    F.n = (int)nn;
    F.nsamp = F.n >> 1;
    F.tp = L.t ? L.t + (o0 >> 1) : tws + (size_t)f * (L.n_max >> 1);
    F.xoff = ptr_parity(F.x);
    F.toff = ptr_parity(F.tp);
    double p[M];
#pragma unroll
    for (int q = 0; q < M; ++q) p[q] = L.p[(size_t)f * M + q];
    const bool blocked = !(F.n * M < kBlockSzSq);  // strict "<", lm_core.c:193

    double mu = 0.0, g_inf = 0.0, p_sq = 0.0, dp_sq = DBL_MAX, maxdiag = 0.0;
    int nu = 2, stop = 0, nfev = 1, njev = 0, nlss = 0, k;
    double e_sq;
    if (L.t) {
      e_sq = i8_eval_pass<false>(F, p, wt, lane, 0.0, 1.0);
    } else {
      const double firsty = F.x[1];
      e_sq = i8_eval_pass<true>(F, p, wt, lane, firsty, F.x[2 * F.nsamp - 1] - firsty);
    }
    const double e0 = e_sq;
    if (!isfinite(e_sq)) stop = 7;
    double uu[10], jte[M];
#pragma unroll
    for (int q = 0; q < 10; ++q) uu[q] = 0.0;
#pragma unroll
    for (int q = 0; q < M; ++q) jte[q] = 0.0;
    bool have_jtj = false;
    for (k = 0; k < L.itmax && !stop; ++k) {
      if (e_sq <= L.eps3) { stop = 6; break; }
      if (blocked) {
        if (have_jtj) {
          double dummy[10];
          i8_jac_pass<false>(F, p, wt, lane, dummy, jte);
        } else {
          i8_jac_pass<true>(F, p, wt, lane, uu, jte);
        }
      } else {
        if (have_jtj) {
          double dummy[10];
          i8_jac_unblocked<false>(F, p, dummy, jte);
        } else {
          i8_jac_unblocked<true>(F, p, uu, jte);
        }
      }
      have_jtj = true;
      ++njev;
      // gradient statistics, lm_core.c:256-287; the diagonal of J^T J is uu[tri(k,k)] for columns 2k and 2k+1
      p_sq = 0.0;
      g_inf = 0.0;
#pragma unroll
      for (int q = 0; q < M; ++q) {
        const double tmp = fabs(jte[q]);
        if (g_inf < tmp) g_inf = tmp;
        p_sq += p[q] * p[q];
      }
      maxdiag = -DBL_MAX;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (uu[q * (q + 1) / 2 + q] > maxdiag) maxdiag = uu[q * (q + 1) / 2 + q];
      if (g_inf <= L.eps1) { dp_sq = 0.0; stop = 1; break; }
      if (k == 0) mu = L.tau * maxdiag;

      for (;;) {
        // (J^T J + mu I) dp = J^T e: the checkerboard matrix from the 10 distinct entries
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int c = 0; c < M; ++c) {
            double v = 0.0;
            if ((i & 1) == (c & 1)) {
              const int hi = (i >> 1) > (c >> 1) ? (i >> 1) : (c >> 1), lo = (i >> 1) > (c >> 1) ? (c >> 1) : (i >> 1);
              v = uu[hi * (hi + 1) / 2 + lo];
              if (i == c) v += mu;
            }
            a[(i * M + c) * T] = v;
          }
#pragma unroll
        for (int i = 0; i < M; ++i) bx[i * T] = jte[i];
        double dp[M];
        const int solved = lu_factor_solve<M, T>(a, sc, bx, dp);
        ++nlss;
        if (solved) {
          double pdp[M];
          dp_sq = 0.0;
#pragma unroll
          for (int q = 0; q < M; ++q) {
            pdp[q] = p[q] + dp[q];
            dp_sq += dp[q] * dp[q];
          }
          if (dp_sq <= L.eps2_sq * p_sq) { stop = 2; break; }
          if (dp_sq >= (p_sq + L.eps2) / (kEpsilon * kEpsilon)) { stop = 4; break; }
          const double e_new = i8_eval_pass<false>(F, pdp, wt, lane, 0.0, 1.0);
          ++nfev;
          if (!isfinite(e_new)) { stop = 7; break; }
          double dL = 0.0;
#pragma unroll
          for (int q = 0; q < M; ++q) dL += dp[q] * (mu * dp[q] + jte[q]);
          const double dF = e_sq - e_new;
          if (dL > 0.0 && dF > 0.0) {
            mu = nielsen(mu, dF, dL);
            nu = 2;
#pragma unroll
            for (int q = 0; q < M; ++q) p[q] = pdp[q];
            e_sq = e_new;
            break;
          }
        }
        mu *= nu;
        const int nu2 = (int)((unsigned)nu << 1);
        if (nu2 <= nu) { stop = 5; break; }
        nu = nu2;
      }
    }
    if (k >= L.itmax) stop = 3;
#pragma unroll
    for (int q = 0; q < M; ++q) L.p[(size_t)f * M + q] = p[q];
    double* info = L.info + (size_t)f * 10;
    info[0] = e0;
    info[1] = e_sq;
    info[2] = g_inf;
    info[3] = dp_sq;
    info[4] = mu / maxdiag;
    info[5] = (double)k;
    info[6] = (double)stop;
    info[7] = (double)nfev;
    info[8] = (double)njev;
    info[9] = (double)nlss;
    L.ret[f] = (stop != 4 && stop != 7) ? k : -1;
    if (L.extra) { L.extra[2 * f] = njev; L.extra[2 * f + 1] = 0; }
  }
}

// ---- dlevmar_dif (lm_core.c:429-828) for the m = 8 model, a thread per fit ----------------------------------------
// The secant Jacobian of a fit (n x 8) lives in a global work area, the rows of a CTA's 128 threads interleaved in
// 16-byte pieces so that a warp's loads and stores coalesce; everything else is
// recomputed from the parameters on the fly: f(p) and f(p + dp) are four-term Bernstein sums, so the stored copies the
// reference keeps (hx, wrk) are never needed -- a value recomputed from the same inputs by the same operations is the
// same value.  The rank-one secant update (:745-755) is applied to a row on its way into the next J^T J build, like in
// the warp-per-fit kernel; a finite-difference Jacobian (misc_core.c:135-209) is formed row by row inside the build that
// follows it (one pass over the work area instead of two) and simply overwrites the rows.
// A pending secant update: p at the time of the trial, the trial point and the step sit in the (then idle) LU work
// matrix in shared memory, [element][thread]: elements 0-7, 8-15, 16-23.
struct I8Pending {
  bool on;
  double dp_sq;
  const double* sh;
};

// a Jacobian row of this thread: four 16-byte pieces, [row][piece][thread] within the CTA's slice (coalesced)
__device__ __forceinline__ double2* i8_jrow(double* J, int row) {
  return reinterpret_cast<double2*>(J) + (size_t)row * 4 * kI8Threads;
}

__device__ __forceinline__ double i8_sum(const double (&w)[4], const double (&p)[8], int uv) {
  double a = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) a += w[k] * p[2 * k + uv];
  return a;
}

// Finite-difference step sizes of a Jacobian evaluation (misc_core.c:135-209): d_j = max(|1e-4 p_j|, delta) and the
// factor the difference is multiplied with, 1/d (forward) or 0.5/d (central).
struct I8Fd {
  double dlt[8], scale[8];
  int forward;
};

// one row of the normal-equation build.  FROM_FD: the row is formed by finite differences -- forward
// (f(p + d e_j) - f(p)) / d or central (f(p + d e_j) - f(p - d e_j)) / (2d) per column, where the perturbed evaluations
// differ from the base one only in the sums that contain the perturbed parameter (the other columns reproduce the base
// value exactly and give +0.0) -- and stored; otherwise it is loaded, and updated and written back if a secant update
// is pending.  Then: residual, accumulation into acc (36 lower entries, index i(i+1)/2 + j) and the J^T e chains.
template <bool FROM_FD>
__device__ __forceinline__ void i8_dif_row(double2* q, const double (&w)[4], int uv, double xval, const double (&p)[8],
                                           const I8Pending& P, const I8Fd& D, double (&acc)[36], double (&jte)[8]) {
  constexpr int T = kI8Threads;
  double jr[8];
  const double base = i8_sum(w, p, uv);
  if (FROM_FD) {
#pragma unroll
    for (int j = 0; j < 8; ++j) jr[j] = 0.0;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      const int j = 2 * kq + uv;
      double hi = 0.0, lo = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double pk = p[2 * k + uv];
        hi += w[k] * ((k == kq) ? pk + D.dlt[j] : pk);
        lo += w[k] * ((k == kq) ? pk - D.dlt[j] : pk);
      }
      jr[j] = D.forward ? (hi - base) * D.scale[j] : (hi - lo) * D.scale[j];
    }
  } else {
    const double2 a0 = q[0], a1 = q[T], a2 = q[2 * T], a3 = q[3 * T];
    jr[0] = a0.x; jr[1] = a0.y; jr[2] = a1.x; jr[3] = a1.y; jr[4] = a2.x; jr[5] = a2.y; jr[6] = a3.x; jr[7] = a3.y;
  }
  if (!FROM_FD && P.on) {
    double f0 = 0.0, f1 = 0.0, dp[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f0 += w[k] * P.sh[(2 * k + uv) * T];
      f1 += w[k] * P.sh[(8 + 2 * k + uv) * T];
    }
    const double d = f1 - f0;
    double tmp = 0.0;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      dp[l] = P.sh[(16 + l) * T];
      tmp += jr[l] * dp[l];
    }
    tmp = (d - tmp) / P.dp_sq;
#pragma unroll
    for (int j = 0; j < 8; ++j) jr[j] += tmp * dp[j];
  }
  if (FROM_FD || P.on) {
    q[0] = make_double2(jr[0], jr[1]);
    q[T] = make_double2(jr[2], jr[3]);
    q[2 * T] = make_double2(jr[4], jr[5]);
    q[3 * T] = make_double2(jr[6], jr[7]);
  }
  const double e = xval - base;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) acc[i * (i + 1) / 2 + j] += jr[j] * jr[i];
    jte[i] += jr[i] * e;
  }
}

// J^T J and J^T e from the stored Jacobian, or (FROM_FD) from a finite-difference Jacobian formed and stored row by
// row on the way.  BLOCKED: per-32-row partial sums, ascending rows (misc_core.c:101-126, lm_core.c:636-644); otherwise
// the unblocked loop over descending rows (lm_core.c:604-628, n*m <= 1024).
template <bool BLOCKED, bool FROM_FD>
__device__ __forceinline__ void i8_dif_normal_eqs(const I8Fit& F, const double (&p)[8], double* J, EvalTiles* wt, int lane,
                                                  const I8Pending& P, const I8Fd& D, double (&jtj)[36], double (&jte)[8]) {
  double* xrow = &wt->x[lane][0];
  double* trow = &wt->t[lane][0];
  const double* xt = xrow + F.xoff;
  const double* tt = trow + F.toff;
#pragma unroll
  for (int q = 0; q < 36; ++q) jtj[q] = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) jte[q] = 0.0;
  double part[36];
  if (BLOCKED) {
#pragma unroll
    for (int q = 0; q < 36; ++q) part[q] = 0.0;
  }
  const int nstage = (F.n + 15) >> 4;
#pragma unroll 1
  for (int it = 0; it < nstage; ++it) {
    const int b = BLOCKED ? it : nstage - 1 - it;
    row_fill(xrow, F.xoff, F.x + 16 * b, min(16, F.n - 16 * b));
    row_fill(trow, F.toff, F.tp + 8 * b, min(8, F.nsamp - 8 * b));
    cp_async_wait_all();
    const int s_lo = 8 * b, s_hi = min(8 * b + 8, F.nsamp);
#pragma unroll 1
    for (int is = 0; is < s_hi - s_lo; ++is) {
      const int s = BLOCKED ? s_lo + is : s_hi - 1 - is;
      const int l = s & 7;
      double w[4];
      bernstein3(tt[l], w);
      if (BLOCKED) {
        i8_dif_row<FROM_FD>(i8_jrow(J, 2 * s), w, 0, xt[2 * l], p, P, D, part, jte);
        i8_dif_row<FROM_FD>(i8_jrow(J, 2 * s + 1), w, 1, xt[2 * l + 1], p, P, D, part, jte);
      } else {
        i8_dif_row<FROM_FD>(i8_jrow(J, 2 * s + 1), w, 1, xt[2 * l + 1], p, P, D, jtj, jte);
        i8_dif_row<FROM_FD>(i8_jrow(J, 2 * s), w, 0, xt[2 * l], p, P, D, jtj, jte);
      }
    }
    if (BLOCKED && ((b & 1) || b + 1 == nstage)) {
#pragma unroll
      for (int q = 0; q < 36; ++q) {
        jtj[q] += part[q];
        part[q] = 0.0;
      }
    }
  }
}

__global__ void __launch_bounds__(kI8Threads, 2) image8_dif_kernel(const LmLaunch L, double* tws, double* jws) {
  constexpr int M = 8, T = kI8Threads;
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double* a = smem + tid;
  double* sc = smem + M * M * T + tid;
  double* bx = smem + (M * M + M) * T + tid;
  EvalTiles* wt = reinterpret_cast<EvalTiles*>(smem + kI8LuDoubles * T) + wid;
  double* J = jws + (size_t)blockIdx.x * T * (size_t)L.n_max * M + 2 * tid;  // see i8_jrow
  for (unsigned int f = blockIdx.x * T + tid; f < (unsigned int)L.B; f += gridDim.x * T) {
    const long long o0 = L.off[f], nn = L.off[f + 1] - o0;
    int ret_code = 0;
    if (nn < M) ret_code = -1;  // lm_core.c:493
    else if (nn > L.n_max || (nn & 1)) ret_code = -2;
    else if (L.counts[f] != (int)(nn >> 1)) ret_code = -3;
    if (ret_code != 0) {
      for (int q = 0; q < 10; ++q) L.info[(size_t)f * 10 + q] = 0.0;
      L.ret[f] = ret_code;
      if (L.extra) { L.extra[2 * f] = 0; L.extra[2 * f + 1] = 0; }
      continue;
    }
    I8Fit F;
    F.x = L.meas + o0;
    F.n = (int)nn;
    F.nsamp = F.n >> 1;
    F.tp = L.t ? L.t + (o0 >> 1) : tws + (size_t)f * (L.n_max >> 1);
    F.xoff = ptr_parity(F.x);
    F.toff = ptr_parity(F.tp);
    double p[M];
#pragma unroll
    for (int q = 0; q < M; ++q) p[q] = L.p[(size_t)f * M + q];
    const bool blocked = !(F.n * M <= kBlockSzSq);  // "<=" here, lm_core.c:585
    const int K = 10;                                // max(m, 10)

    double mu = 0.0, g_inf = 0.0, p_sq = 0.0, dp_sq = DBL_MAX, maxdiag = 0.0;
    int nu = 20, stop = 0, nfev = 1, njap = 0, nlss = 0, n_jtj = 0, n_broyden = 0, k;
    int updjac = 0, updp = 1, newjac = 0;
    double e_sq;
    if (L.t) {
      e_sq = i8_eval_pass<false>(F, p, wt, lane, 0.0, 1.0);
    } else {
      const double firsty = F.x[1];
      e_sq = i8_eval_pass<true>(F, p, wt, lane, firsty, F.x[2 * F.nsamp - 1] - firsty);
    }
    const double e0 = e_sq;
    if (!isfinite(e_sq)) stop = 7;
    double jtj[36], jte[M];
#pragma unroll
    for (int q = 0; q < 36; ++q) jtj[q] = 0.0;
#pragma unroll
    for (int q = 0; q < M; ++q) jte[q] = 0.0;
    I8Pending P;
    P.on = false;
    P.dp_sq = 1.0;
    P.sh = a;
    I8Fd D;
    D.forward = L.forward;
#pragma unroll
    for (int j = 0; j < M; ++j) { D.dlt[j] = 0.0; D.scale[j] = 0.0; }
    bool fresh_fd = false;

    for (k = 0; k < L.itmax && !stop; ++k) {
      if (e_sq <= L.eps3) { stop = 6; break; }
      if ((updp && nu > 16) || updjac == K) {
        // finite-difference Jacobian at p: formed inside the normal-equation build that always follows
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double d = 1E-04 * p[j];
          d = fabs(d);
          if (d < L.delta) d = L.delta;
          D.dlt[j] = d;
          D.scale[j] = L.forward ? 1.0 / d : 0.5 / d;
        }
        fresh_fd = true;
        ++njap;
        nfev += L.forward ? M : 2 * M;
        nu = 2; updjac = 0; updp = 0; newjac = 1;
        P.on = false;  // the fresh Jacobian overwrites every row a deferred secant update would have touched
      }
      if (newjac) {
        newjac = 0;
        if (fresh_fd) {
          if (blocked) i8_dif_normal_eqs<true, true>(F, p, J, wt, lane, P, D, jtj, jte);
          else i8_dif_normal_eqs<false, true>(F, p, J, wt, lane, P, D, jtj, jte);
        } else {
          if (blocked) i8_dif_normal_eqs<true, false>(F, p, J, wt, lane, P, D, jtj, jte);
          else i8_dif_normal_eqs<false, false>(F, p, J, wt, lane, P, D, jtj, jte);
        }
        fresh_fd = false;
        P.on = false;
        ++n_jtj;
        p_sq = 0.0;
        g_inf = 0.0;
#pragma unroll
        for (int q = 0; q < M; ++q) {
          const double tmp = fabs(jte[q]);
          if (g_inf < tmp) g_inf = tmp;
          p_sq += p[q] * p[q];
        }
        maxdiag = -DBL_MAX;
#pragma unroll
        for (int q = 0; q < M; ++q)
          if (jtj[q * (q + 1) / 2 + q] > maxdiag) maxdiag = jtj[q * (q + 1) / 2 + q];
      }
      if (g_inf <= L.eps1) { dp_sq = 0.0; stop = 1; break; }
      if (k == 0) mu = L.tau * maxdiag;

#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int c = 0; c < M; ++c) {
          double v = (i >= c) ? jtj[i * (i + 1) / 2 + c] : jtj[c * (c + 1) / 2 + i];
          if (i == c) v += mu;
          a[(i * M + c) * T] = v;
        }
#pragma unroll
      for (int i = 0; i < M; ++i) bx[i * T] = jte[i];
      double dp[M];
      const int solved = lu_factor_solve<M, T>(a, sc, bx, dp);
      ++nlss;
      if (solved) {
        double pdp[M];
        dp_sq = 0.0;
#pragma unroll
        for (int q = 0; q < M; ++q) {
          pdp[q] = p[q] + dp[q];
          dp_sq += dp[q] * dp[q];
        }
        if (dp_sq <= L.eps2_sq * p_sq) { stop = 2; break; }
        if (dp_sq >= (p_sq + L.eps2) / (kEpsilon * kEpsilon)) { stop = 4; break; }
        const double e_new = i8_eval_pass<false>(F, pdp, wt, lane, 0.0, 1.0);
        ++nfev;
        if (!isfinite(e_new)) { stop = 7; break; }
        const double dF = e_sq - e_new;
        if (updp || dF > 0.0) {
          // rank-one secant update, deferred to the next normal-equation build (nothing reads J before it)
          P.on = true;
          P.dp_sq = dp_sq;
#pragma unroll
          for (int q = 0; q < M; ++q) {
            a[q * T] = p[q];
            a[(8 + q) * T] = pdp[q];
            a[(16 + q) * T] = dp[q];
          }
          ++updjac;
          newjac = 1;
          ++n_broyden;
        }
        double dL = 0.0;
#pragma unroll
        for (int q = 0; q < M; ++q) dL += dp[q] * (mu * dp[q] + jte[q]);
        if (dL > 0.0 && dF > 0.0) {
          mu = nielsen(mu, dF, dL);
          nu = 2;
#pragma unroll
          for (int q = 0; q < M; ++q) p[q] = pdp[q];
          e_sq = e_new;
          updp = 1;
          continue;
        }
      }
      mu *= nu;
      const int nu2 = (int)((unsigned)nu << 1);
      if (nu2 <= nu) { stop = 5; break; }
      nu = nu2;
    }
    if (k >= L.itmax) stop = 3;
#pragma unroll
    for (int q = 0; q < M; ++q) L.p[(size_t)f * M + q] = p[q];
    double* info = L.info + (size_t)f * 10;
    info[0] = e0;
    info[1] = e_sq;
    info[2] = g_inf;
    info[3] = dp_sq;
    info[4] = mu / maxdiag;
    info[5] = (double)k;
    info[6] = (double)stop;
    info[7] = (double)nfev;
    info[8] = (double)njap;
    info[9] = (double)nlss;
    L.ret[f] = (stop != 4 && stop != 7) ? k : -1;
    if (L.extra) { L.extra[2 * f] = n_jtj; L.extra[2 * f + 1] = n_broyden; }
  }
}

}  // namespace cps
