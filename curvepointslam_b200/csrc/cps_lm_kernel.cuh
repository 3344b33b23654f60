// cps_lm_kernel.cuh -- batched Levenberg-Marquardt, one warp per curve fit (sm_100a, FP64).
//
// What it computes: exactly what levmar-2.5 computes for one fit -- dlevmar_der (lm_core.c:60-423) or
// dlevmar_dif (lm_core.c:429-828) with the LU normal-equation solve dAx_eq_b_LU_noLapack
// (Axb_core.c:992-1139), e = x - hx and its norm as dlevmar_L2nrmxmy (misc_core.c:712-798), J^T J as
// dlevmar_trans_mat_mat_mult (misc_core.c:80-132) or the unblocked small-problem loop
// (lm_core.c:213-237 / :604-628), forward/central difference Jacobians (misc_core.c:135-209) -- for B
// independent fits at once.  It is not a translation: the reference is a scalar, single-threaded C
// loop nest over a malloc'ed work array; here a warp owns a fit, the contour samples, residuals and a
// 32-row Jacobian tile live in shared memory, the lower triangle of J^T J is spread over the lanes'
// registers, the LU elimination runs right-looking with one matrix row per lane, and warps pull fits
// from a global queue so that fits with different iteration counts do not hold each other up.
//
// Parity contract (DESIGN.md "Bit-exactness"): every floating-point value is produced by the same
// sequence of IEEE-754 operations the reference performs on that value (same summation order, same
// pivot rule, no FMA: this translation unit is compiled with -fmad=false), so parameters, info[] and
// return codes are bit-identical to the CPU restatement in oracle/ (which is itself pinned bit-for-bit
// to the reference's levmar built from /root/reference).  Parallelism comes only from operations the
// reference leaves independent: different entries of J^T J, different rows of J, different rows of the
// elimination step, different samples of the model.
#pragma once
#include <cfloat>
#include <cstdint>

#include "cps_model.cuh"

namespace cps {

constexpr int kWarpsPerCta = 4;
constexpr int kTileRows = 32;  // the reference's __BLOCKSZ__ (misc.h:48): row blocks of J^T J partial sums
constexpr double kEpsilon = 1E-12;          // lm.c:35
constexpr double kOneThird = 0.3333333334;  // lm.c:36 (sic)
constexpr int kBlockSzSq = 32 * 32;         // lm_core.c:193,585 switch between the two J^T J loops

struct LmLaunch {
  int B;
  int n_max;  // capacity (doubles) of the per-warp measurement buffers; fits with n > n_max are refused
  int itmax;
  double tau, eps1, eps2, eps2_sq, eps3, delta;
  int forward;  // dif: 1 = forward differences, 0 = central (opts[4] < 0)
  double* p;                // [B][M] in/out
  const double* meas;       // concatenated samples
  const double* t;          // concatenated t per sample, or nullptr: derive from image rows (fit_curve:451-511)
  const long long* off;     // [B+1] offsets into meas (doubles)
  const int* counts;        // [B][NSEG]
  double* info;             // [B][10]
  int* ret;                 // [B]
  int* extra;               // [B][2] (n_jtj, n_broyden) or nullptr
  double* jstore;           // dif: [resident warps][M][n_max] column-major Jacobians
  unsigned int* queue;      // next fit index
};

__host__ __device__ constexpr int lm_ldt(int M) { return M | 1; }

// doubles of shared memory one warp needs
__host__ __device__ constexpr int lm_warp_smem_doubles(int M, int scratch, int n_max, bool dif) {
  return n_max /*x*/ + n_max / 2 /*t*/ + n_max /*e*/ + n_max /*hx*/ + (dif ? 2 * n_max : 0) +
         kTileRows * lm_ldt(M) /*tile, reused as LU matrix*/ + M * M /*jtj*/ + 8 * 32 /*small vectors*/ + scratch;
}

template <int M>
struct WarpMem {
  double *x, *t, *e, *hx, *wrk, *wrk2, *tile, *jtj;
  double *p, *dp, *pdp, *jte, *diag, *pw, *xs, *scale;
  double* scratch;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ---- model evaluation over all samples of the fit ----------------------------------------------------
template <class MD>
__device__ __forceinline__ void eval_func(const WarpMem<MD::M>& w, const FitGeom& g, const double* p, double* out,
                                          int lane) {
  MD::prep_func(p, w.scratch, lane);
  __syncwarp();
  for (int s = lane; s < g.nsamp; s += 32) {
    double hu, hv;
    MD::eval_sample(w.scratch, g, s, w.t[s], hu, hv);
    out[2 * s] = hu;
    out[2 * s + 1] = hv;
  }
  __syncwarp();
}

// e = x - y, returns ||e||^2 in the summation order of dlevmar_L2nrmxmy (misc_core.c:712-798): four
// running sums fed from the top index downwards in groups of eight, ragged tail upwards, then
// ((s0+s1)+s2)+s3.  The four chains run on lanes 0..3.  e may alias y.
__device__ __forceinline__ double residual_sqnorm(double* e, const double* x, const double* y, int n, int lane) {
  for (int i = lane; i < n; i += 32) e[i] = x[i] - y[i];
  __syncwarp();
  double s = 0.0;
  if (lane < 4) {
    const int full = (n >> 3) << 3;
    for (int top = full - 1; top > 0; top -= 8) {
      double d = e[top - lane];
      s += d * d;
      d = e[top - lane - 4];
      s += d * d;
    }
    const int rem = n - full;
    for (int k = 0; k < rem; ++k) {
      const int c = rem - k;
      if (((7 - c) & 3) == lane) {
        const double d = e[full + k];
        s += d * d;
      }
    }
  }
  const double s0 = shfl_d(s, 0), s1 = shfl_d(s, 1), s2 = shfl_d(s, 2), s3 = shfl_d(s, 3);
  return ((s0 + s1) + s2) + s3;
}

// ---- J^T J and J^T e ----------------------------------------------------------------------------------
template <int M>
struct Accum {
  static constexpr int NE = M * (M + 1) / 2;
  static constexpr int E = (NE + 31) / 32;
  double acc[E];
  int ei[E], ej[E];
  double jte;

  __device__ __forceinline__ void init(int lane) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      int idx = e * 32 + lane;
      if (idx >= NE) idx = 0;
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= idx) ++i;
      ei[e] = i;
      ej[e] = idx - i * (i + 1) / 2;
    }
  }
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.0;
    jte = 0.0;
  }
  // one tile of `rows` Jacobian rows (global rows kk..kk+rows-1).  BLOCKED: the reference's blocked
  // product (per-block partial sums, ascending rows); otherwise the unblocked descending-row loop.
  template <bool BLOCKED>
  __device__ __forceinline__ void add_tile(const double* tile, int rows, const double* e_kk, int lane) {
    constexpr int LDT = lm_ldt(M);
    const int col = (lane < M) ? lane : 0;
    if (BLOCKED) {
      double part[E];
#pragma unroll
      for (int e = 0; e < E; ++e) part[e] = 0.0;
      for (int k = 0; k < rows; ++k) {
        const double* row = tile + k * LDT;
#pragma unroll
        for (int e = 0; e < E; ++e) part[e] += row[ei[e]] * row[ej[e]];
        jte += row[col] * e_kk[k];
      }
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] += part[e];
    } else {
      for (int k = rows - 1; k >= 0; --k) {
        const double* row = tile + k * LDT;
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] += row[ej[e]] * row[ei[e]];
        jte += row[col] * e_kk[k];
      }
    }
  }
  __device__ __forceinline__ void store(double* jtj, double* jte_out, int lane) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (e * 32 + lane < NE) {
        jtj[ei[e] * M + ej[e]] = acc[e];
        jtj[ej[e] * M + ei[e]] = acc[e];
      }
    }
    if (lane < M) jte_out[lane] = jte;
    __syncwarp();
  }
};

// analytic Jacobian evaluated tile by tile, never stored (der path)
template <class MD>
__device__ __forceinline__ void normal_eqs_analytic(const WarpMem<MD::M>& w, const FitGeom& g, Accum<MD::M>& A,
                                                    bool blocked, int lane) {
  constexpr int M = MD::M;
  constexpr int LDT = lm_ldt(M);
  MD::prep_jac(w.p, w.scratch, lane);
  __syncwarp();
  A.clear();
  const int ntiles = (g.n + kTileRows - 1) / kTileRows;
  for (int it = 0; it < ntiles; ++it) {
    const int tix = blocked ? it : (ntiles - 1 - it);
    const int kk = tix * kTileRows;
    const int rows = min(kTileRows, g.n - kk);
    if (lane < rows) MD::jac_row(w.scratch, g, kk + lane, w.t[(kk + lane) >> 1], w.tile + lane * LDT);
    __syncwarp();
    if (blocked) A.template add_tile<true>(w.tile, rows, w.e + kk, lane);
    else A.template add_tile<false>(w.tile, rows, w.e + kk, lane);
    __syncwarp();
  }
  A.store(w.jtj, w.jte, lane);
}

// stored Jacobian (dif path): column-major in global memory, J[j*ld + r]
template <int M>
__device__ __forceinline__ void normal_eqs_stored(const WarpMem<M>& w, const FitGeom& g, const double* J, int ld,
                                                  Accum<M>& A, bool blocked, int lane) {
  constexpr int LDT = lm_ldt(M);
  A.clear();
  const int ntiles = (g.n + kTileRows - 1) / kTileRows;
  for (int it = 0; it < ntiles; ++it) {
    const int tix = blocked ? it : (ntiles - 1 - it);
    const int kk = tix * kTileRows;
    const int rows = min(kTileRows, g.n - kk);
    if (lane < rows) {
#pragma unroll
      for (int j = 0; j < M; ++j) w.tile[lane * LDT + j] = J[j * ld + kk + lane];
    }
    __syncwarp();
    if (blocked) A.template add_tile<true>(w.tile, rows, w.e + kk, lane);
    else A.template add_tile<false>(w.tile, rows, w.e + kk, lane);
    __syncwarp();
  }
  A.store(w.jtj, w.jte, lane);
}

// ---- (J^T J + mu I) dp = J^T e : the reference's LU (Axb_core.c:1044-1132) ---------------------------
// Crout's column sweep is re-associated into a right-looking elimination: element (i,j) still receives
// a_ij - l_i0*u_0j - l_i1*u_1j - ... in ascending k, so every stored value is bit-identical; lane i owns
// row i.  Pivot: argmax of scale_i*|a_ij| over rows >= j, LAST maximum wins (">=", :1088); a zero pivot
// becomes DBL_EPSILON (:1102); forward substitution skips leading zeros of the permuted right-hand side
// (:1112-1121); back substitution subtracts in ascending column order (:1124-1129).
// Returns 1 on success, 0 if a row of the matrix is entirely zero.
template <int M>
__device__ __forceinline__ int lu_solve(const WarpMem<M>& w, double mu, int lane) {
  constexpr int LD = M;
  double* a = w.tile;
  // augmented copy
  for (int idx = lane; idx < M * M; idx += 32) {
    const int i = idx / M, j = idx - i * M;
    double v = w.jtj[idx];
    if (i == j) v += mu;
    a[i * LD + j] = v;
  }
  double b = (lane < M) ? w.jte[lane] : 0.0;
  __syncwarp();
  double scale = 0.0;
  int singular = 0;
  if (lane < M) {
    double big = 0.0;
    for (int j = 0; j < M; ++j) {
      const double v = fabs(a[lane * LD + j]);
      if (v > big) big = v;
    }
    if (big == 0.0) singular = 1;
    else scale = 1.0 / big;
  }
  if (__any_sync(0xffffffffu, singular)) return 0;

  int piv = -1;  // the reference's maxi persists across columns
  for (int j = 0; j < M; ++j) {
    // pivot search over rows j..M-1 in ascending order with ">=" against a running maximum from 0
    double merit = -1.0;
    if (lane >= j && lane < M) merit = scale * fabs(a[lane * LD + j]);
    double best = 0.0;
    int besti = -1;
    // sequential semantics: last index whose merit >= all earlier maxima and >= 0; NaN never wins
    {
      double mx = (merit >= 0.0) ? merit : -1.0;  // NaN -> -1
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double other = shfl_d(mx, lane ^ o);
        mx = (other > mx) ? other : mx;
      }
      best = mx;
      const unsigned ballot = __ballot_sync(0xffffffffu, merit >= 0.0 && merit == best);
      besti = ballot ? (31 - __clz(ballot)) : -1;
    }
    if (besti >= 0) piv = besti;
    if (piv != j && piv >= 0) {
      if (lane < M) {
        const double tmp = a[piv * LD + lane];
        a[piv * LD + lane] = a[j * LD + lane];
        a[j * LD + lane] = tmp;
      }
      // scale[piv] = scale[j]; right-hand side follows its row
      const double sj = shfl_d(scale, j);
      const double bj = shfl_d(b, j), bp = shfl_d(b, piv);
      if (lane == piv) { scale = sj; b = bj; }
      if (lane == j) b = bp;
      __syncwarp();
    }
    if (lane == 0 && a[j * LD + j] == 0.0) a[j * LD + j] = DBL_EPSILON;
    __syncwarp();
    if (j != M - 1) {
      const double inv = 1.0 / a[j * LD + j];
      if (lane > j && lane < M) {
        const double l = a[lane * LD + j] * inv;
        a[lane * LD + j] = l;
        for (int k = j + 1; k < M; ++k) a[lane * LD + k] -= l * a[j * LD + k];
      }
      __syncwarp();
    }
  }
  // forward substitution, column oriented (same ascending-j order per row), with the leading-zero skip
  bool started = false;
  for (int j = 0; j < M - 1; ++j) {
    const double xj = shfl_d(b, j);
    if (!started) {
      if (xj != 0.0) started = true;
      else continue;
    }
    if (lane > j && lane < M) b -= a[lane * LD + j] * xj;
  }
  // back substitution: row i subtracts a_ij*x_j for j = i+1..M-1 in ascending j, then divides
  for (int i = M - 1; i >= 0; --i) {
    if (lane == i) {
      double sum = b;
      for (int jj = i + 1; jj < M; ++jj) sum -= a[i * LD + jj] * w.xs[jj];
      w.xs[i] = sum / a[i * LD + i];
    }
    __syncwarp();
  }
  if (lane < M) w.dp[lane] = w.xs[lane];
  __syncwarp();
  return 1;
}

// ---- the two LM drivers ---------------------------------------------------------------------------
struct LmResult {
  double e0, e_sq, g_inf, dp_sq, mu;
  int k, stop, nfev, njev, nlss, n_jtj, n_broyden;
};

template <int M>
__device__ __forceinline__ void grad_stats(const WarpMem<M>& w, double& p_sq, double& g_inf, int lane) {
  p_sq = 0.0;
  g_inf = 0.0;
  for (int i = 0; i < M; ++i) {
    const double tmp = fabs(w.jte[i]);
    if (g_inf < tmp) g_inf = tmp;
    p_sq += w.p[i] * w.p[i];
  }
  if (lane < M) w.diag[lane] = w.jtj[lane * M + lane];
  __syncwarp();
}

template <int M>
__device__ __forceinline__ double max_diag(const WarpMem<M>& w) {
  double tmp = -DBL_MAX;
  for (int i = 0; i < M; ++i)
    if (w.diag[i] > tmp) tmp = w.diag[i];
  return tmp;
}

// returns 0 to continue, 2 or 4 to stop; fills pdp and dp_sq
template <int M>
__device__ __forceinline__ int step_checks(const WarpMem<M>& w, double p_sq, double eps2, double eps2_sq,
                                           double& dp_sq, int lane) {
  if (lane < M) w.pdp[lane] = w.p[lane] + w.dp[lane];
  dp_sq = 0.0;
  for (int i = 0; i < M; ++i) {
    const double tmp = w.dp[i];
    dp_sq += tmp * tmp;
  }
  __syncwarp();
  if (dp_sq <= eps2_sq * p_sq) return 2;
  if (dp_sq >= (p_sq + eps2) / (kEpsilon * kEpsilon)) return 4;
  return 0;
}

template <int M>
__device__ __forceinline__ double predicted_reduction(const WarpMem<M>& w, double mu) {
  double dL = 0.0;
  for (int i = 0; i < M; ++i) dL += w.dp[i] * (mu * w.dp[i] + w.jte[i]);
  return dL;
}

__device__ __forceinline__ double nielsen(double mu, double dF, double dL) {
  double tmp = (2.0 * dF / dL - 1.0);
  tmp = 1.0 - tmp * tmp * tmp;
  return mu * ((tmp >= kOneThird) ? tmp : kOneThird);
}

// dlevmar_der, lm_core.c:60-423
template <class MD>
__device__ void lm_der(const WarpMem<MD::M>& w, const FitGeom& g, const LmLaunch& L, Accum<MD::M>& A, LmResult& R,
                       int lane) {
  constexpr int M = MD::M;
  const bool blocked = !(g.n * M < kBlockSzSq);  // strict "<" selects the unblocked loop, lm_core.c:193
  double mu = 0.0, g_inf = 0.0, p_sq = 0.0, dp_sq = DBL_MAX;
  int nu = 2, stop = 0, nfev = 0, njev = 0, nlss = 0, k;

  eval_func<MD>(w, g, w.p, w.hx, lane);
  nfev = 1;
  double e_sq = residual_sqnorm(w.e, w.x, w.hx, g.n, lane);
  const double e0 = e_sq;
  if (!isfinite(e_sq)) stop = 7;

  for (k = 0; k < L.itmax && !stop; ++k) {
    if (e_sq <= L.eps3) { stop = 6; break; }
    normal_eqs_analytic<MD>(w, g, A, blocked, lane);
    ++njev;
    grad_stats<M>(w, p_sq, g_inf, lane);
    if (g_inf <= L.eps1) { dp_sq = 0.0; stop = 1; break; }
    if (k == 0) mu = L.tau * max_diag<M>(w);

    for (;;) {
      const int solved = lu_solve<M>(w, mu, lane);
      ++nlss;
      if (solved) {
        const int chk = step_checks<M>(w, p_sq, L.eps2, L.eps2_sq, dp_sq, lane);
        if (chk) { stop = chk; break; }
        eval_func<MD>(w, g, w.pdp, w.hx, lane);
        ++nfev;
        const double e_new = residual_sqnorm(w.hx, w.x, w.hx, g.n, lane);
        if (!isfinite(e_new)) { stop = 7; break; }
        const double dL = predicted_reduction<M>(w, mu);
        const double dF = e_sq - e_new;
        if (dL > 0.0 && dF > 0.0) {
          mu = nielsen(mu, dF, dL);
          nu = 2;
          __syncwarp();
          if (lane < M) w.p[lane] = w.pdp[lane];
          for (int i = lane; i < g.n; i += 32) w.e[i] = w.hx[i];
          __syncwarp();
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
  R.e0 = e0; R.e_sq = e_sq; R.g_inf = g_inf; R.dp_sq = dp_sq; R.mu = mu;
  R.k = k; R.stop = stop; R.nfev = nfev; R.njev = njev; R.nlss = nlss; R.n_jtj = njev; R.n_broyden = 0;
}

// dlevmar_dif, lm_core.c:429-828 (secant LM; J lives column-major in `J`, leading dimension ld)
template <class MD>
__device__ void lm_dif(const WarpMem<MD::M>& w, const FitGeom& g, const LmLaunch& L, Accum<MD::M>& A, double* J,
                       int ld, LmResult& R, int lane) {
  constexpr int M = MD::M;
  const bool blocked = !(g.n * M <= kBlockSzSq);  // "<=" here, lm_core.c:585
  const int K = (M >= 10) ? M : 10;
  double mu = 0.0, g_inf = 0.0, p_sq = 0.0, dp_sq = DBL_MAX;
  int nu, stop = 0, nfev = 0, njap = 0, nlss = 0, k, n_jtj = 0, n_broyden = 0;
  int updjac = 0, updp = 1, newjac = 0;

  eval_func<MD>(w, g, w.p, w.hx, lane);
  nfev = 1;
  double e_sq = residual_sqnorm(w.e, w.x, w.hx, g.n, lane);
  const double e0 = e_sq;
  if (!isfinite(e_sq)) stop = 7;
  nu = 20;  // forces a Jacobian on the first pass, lm_core.c:555

  for (k = 0; k < L.itmax && !stop; ++k) {
    if (e_sq <= L.eps3) { stop = 6; break; }

    if ((updp && nu > 16) || updjac == K) {
      // finite-difference Jacobian, misc_core.c:135-209
      if (lane < M) w.pw[lane] = w.p[lane];
      __syncwarp();
      for (int j = 0; j < M; ++j) {
        double d = 1E-04 * w.p[j];
        d = fabs(d);
        if (d < L.delta) d = L.delta;
        const double keep = w.p[j];
        if (L.forward) {
          if (lane == 0) w.pw[j] = keep + d;
          __syncwarp();
          eval_func<MD>(w, g, w.pw, w.wrk, lane);
          const double inv = 1.0 / d;
          for (int i = lane; i < g.n; i += 32) J[j * ld + i] = (w.wrk[i] - w.hx[i]) * inv;
        } else {
          if (lane == 0) w.pw[j] = keep - d;
          __syncwarp();
          eval_func<MD>(w, g, w.pw, w.wrk, lane);
          if (lane == 0) w.pw[j] = keep + d;
          __syncwarp();
          eval_func<MD>(w, g, w.pw, w.wrk2, lane);
          const double half_inv = 0.5 / d;
          for (int i = lane; i < g.n; i += 32) J[j * ld + i] = (w.wrk2[i] - w.wrk[i]) * half_inv;
        }
        if (lane == 0) w.pw[j] = keep;
        __syncwarp();
      }
      ++njap;
      nfev += L.forward ? M : 2 * M;
      nu = 2; updjac = 0; updp = 0; newjac = 1;
    }

    if (newjac) {
      newjac = 0;
      normal_eqs_stored<M>(w, g, J, ld, A, blocked, lane);
      ++n_jtj;
      grad_stats<M>(w, p_sq, g_inf, lane);
    }
    if (g_inf <= L.eps1) { dp_sq = 0.0; stop = 1; break; }
    if (k == 0) mu = L.tau * max_diag<M>(w);

    const int solved = lu_solve<M>(w, mu, lane);
    ++nlss;
    if (solved) {
      const int chk = step_checks<M>(w, p_sq, L.eps2, L.eps2_sq, dp_sq, lane);
      if (chk) { stop = chk; break; }
      eval_func<MD>(w, g, w.pdp, w.wrk, lane);
      ++nfev;
      const double e_new = residual_sqnorm(w.wrk2, w.x, w.wrk, g.n, lane);
      if (!isfinite(e_new)) { stop = 7; break; }
      const double dF = e_sq - e_new;
      if (updp || dF > 0.0) {
        // rank-one secant update, lm_core.c:745-755: one Jacobian row per lane
        for (int i = lane; i < g.n; i += 32) {
          double tmp = 0.0;
#pragma unroll
          for (int l = 0; l < M; ++l) tmp += J[l * ld + i] * w.dp[l];
          tmp = (w.wrk[i] - w.hx[i] - tmp) / dp_sq;
#pragma unroll
          for (int j = 0; j < M; ++j) J[j * ld + i] += tmp * w.dp[j];
        }
        __syncwarp();
        ++updjac;
        newjac = 1;
        ++n_broyden;
      }
      const double dL = predicted_reduction<M>(w, mu);
      if (dL > 0.0 && dF > 0.0) {
        mu = nielsen(mu, dF, dL);
        nu = 2;
        __syncwarp();
        if (lane < M) w.p[lane] = w.pdp[lane];
        for (int i = lane; i < g.n; i += 32) {
          w.e[i] = w.wrk2[i];
          w.hx[i] = w.wrk[i];
        }
        __syncwarp();
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
  R.e0 = e0; R.e_sq = e_sq; R.g_inf = g_inf; R.dp_sq = dp_sq; R.mu = mu;
  R.k = k; R.stop = stop; R.nfev = nfev; R.njev = njap; R.nlss = nlss; R.n_jtj = n_jtj; R.n_broyden = n_broyden;
}

// ---- kernel: persistent warps pulling fits from a queue ----------------------------------------------
template <int MODEL, bool DIF>
__global__ void __launch_bounds__(kWarpsPerCta * 32) lm_fit_kernel(const LmLaunch L) {
  using MD = Model<MODEL>;
  constexpr int M = MD::M;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int per_warp = lm_warp_smem_doubles(M, MD::SCRATCH, L.n_max, DIF);
  double* base = smem + (size_t)wid * per_warp;

  WarpMem<M> w;
  w.x = base; base += L.n_max;
  w.t = base; base += L.n_max / 2;
  w.e = base; base += L.n_max;
  w.hx = base; base += L.n_max;
  if (DIF) { w.wrk = base; base += L.n_max; w.wrk2 = base; base += L.n_max; }
  else { w.wrk = nullptr; w.wrk2 = nullptr; }
  w.tile = base; base += kTileRows * lm_ldt(M);
  w.jtj = base; base += M * M;
  w.p = base; w.dp = base + 32; w.pdp = base + 64; w.jte = base + 96; w.diag = base + 128; w.pw = base + 160;
  w.xs = base + 192; w.scale = base + 224; base += 256;
  w.scratch = base;

  const int gwarp = blockIdx.x * (blockDim.x >> 5) + wid;
  double* J = DIF ? (L.jstore + (size_t)gwarp * M * L.n_max) : nullptr;

  Accum<M> A;
  A.init(lane);

  for (;;) {
    unsigned int f = 0;
    if (lane == 0) f = atomicAdd(L.queue, 1u);
    f = __shfl_sync(0xffffffffu, f, 0);
    if (f >= (unsigned)L.B) break;

    const long long o0 = L.off[f], o1 = L.off[f + 1];
    const long long nn = o1 - o0;
    FitGeom g;
    g.n = (int)nn;
    g.nsamp = g.n >> 1;
    int ret_code = 0;
    if (nn < M) ret_code = -1;               // lm_core.c:117 / :493: fewer measurements than unknowns
    else if (nn > L.n_max || (nn & 1)) ret_code = -2;  // does not fit the staged buffers / odd length
    {
      int c[4] = {0, 0, 0, 0};
#pragma unroll
      for (int s = 0; s < MD::NSEG; ++s) c[s] = L.counts[(size_t)f * MD::NSEG + s];
      g.b1 = c[0];
      g.b2 = c[0] + c[1];
      g.b3 = c[0] + c[1] + c[2];
      if (MD::NSEG == 1) { g.b1 = g.b2 = g.b3 = g.nsamp; }
      else if (ret_code == 0 && (c[0] < 0 || c[1] < 0 || c[2] < 0 || c[3] < 0 || g.b3 + c[3] != g.nsamp))
        ret_code = -3;  // segment sizes inconsistent with the measurement count
      if (MD::NSEG == 1 && ret_code == 0 && c[0] != g.nsamp) ret_code = -3;
    }
    if (ret_code != 0) {
      if (lane < 10) L.info[(size_t)f * 10 + lane] = 0.0;
      if (lane == 0) {
        L.ret[f] = ret_code;
        if (L.extra) { L.extra[2 * f] = 0; L.extra[2 * f + 1] = 0; }
      }
      continue;
    }

    // stage the contour samples (coalesced), the t parameter of every sample and the parameters
    const double* xg = L.meas + o0;
    for (int i = lane; i < g.n; i += 32) w.x[i] = xg[i];
    if (lane < M) w.p[lane] = L.p[(size_t)f * M + lane];
    __syncwarp();
    if (L.t) {
      const double* tg = L.t + (o0 >> 1);
      for (int s = lane; s < g.nsamp; s += 32) w.t[s] = tg[s];
    } else {
      // fit_curve:451-511: t = clamp((y - y_first)/(y_last - y_first), 0, 1) per segment; NaN passes through
      const int bnd[5] = {0, g.b1, g.b2, g.b3, g.nsamp};
#pragma unroll
      for (int sgi = 0; sgi < 4; ++sgi) {
        const int lo = bnd[sgi], hi = bnd[sgi + 1];
        if (hi > lo) {
          const double firsty = w.x[2 * lo + 1], lasty = w.x[2 * hi - 1];
          for (int s = lo + lane; s < hi; s += 32) {
            double v = (w.x[2 * s + 1] - firsty) / (lasty - firsty);
            v = (v < 0.0) ? 0.0 : v;
            v = (v > 1.0) ? 1.0 : v;
            w.t[s] = v;
          }
        }
      }
    }
    if (lane < 32) w.diag[lane] = 0.0;
    __syncwarp();

    LmResult R;
    if (DIF) lm_dif<MD>(w, g, L, A, J, L.n_max, R, lane);
    else lm_der<MD>(w, g, L, A, R, lane);

    // info[] exactly as lm_core.c:396-409 / :800-813 (the diagonal of J^T J is the un-augmented one)
    if (lane < M) L.p[(size_t)f * M + lane] = w.p[lane];
    if (lane == 0) {
      double big = -DBL_MAX;
      for (int i = 0; i < M; ++i)
        if (big < w.diag[i]) big = w.diag[i];
      double* info = L.info + (size_t)f * 10;
      info[0] = R.e0;
      info[1] = R.e_sq;
      info[2] = R.g_inf;
      info[3] = R.dp_sq;
      info[4] = R.mu / big;
      info[5] = (double)R.k;
      info[6] = (double)R.stop;
      info[7] = (double)R.nfev;
      info[8] = (double)R.njev;
      info[9] = (double)R.nlss;
      L.ret[f] = (R.stop != 4 && R.stop != 7) ? R.k : -1;
      if (L.extra) { L.extra[2 * f] = R.n_jtj; L.extra[2 * f + 1] = R.n_broyden; }
    }
    __syncwarp();
  }
}

}  // namespace cps
