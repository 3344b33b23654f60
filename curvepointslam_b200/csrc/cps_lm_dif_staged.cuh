// cps_lm_dif_staged.cuh -- dlevmar_dif (lm_core.c:429-828, the reference's one live call, curveFittingClass.cpp:681)
// for large batches of Stereo19 fits as rounds of phase kernels, like cps_lm_staged.cuh does for dlevmar_der.
//
// The secant algorithm keeps an n x 19 Jacobian per fit (78 KB at n = 512) that every iteration reads, updates by a
// rank-one correction (Broyden, :745-755) and contracts into J^T J / J^T e (:581-655).  The warp-per-fit kernel
// (cps_lm_kernel.cuh lm_dif) spreads one fit's tiles over 32 lanes and runs every phase at the occupancy of the
// hungriest; here every phase is its own kernel with the fits themselves as the parallelism (lane = fit):
//
//   solve  (J^T J + mu I) dp = J^T e by the reference's LU, step checks, dL                       lm_core.c:683-722
//   eval   ||x - f(p+dp)||^2 in the reference's four chains (f is never stored: it is recomputed from the
//          parameters wherever the reference reads hx / wrk -- same operations, same bits), accept / reject,
//          damping, the loop-top checks and the Jacobian policy of the next iteration         :724-792, 557-579
//   build  the J^T J / J^T e rebuild with the Jacobian produced on the way in: either the forward-difference
//          Jacobian (misc_core.c:135-170) evaluated row by row, or the stored rows with the pending rank-one
//          update applied; the rows are written back once                                      :569-655, 745-755
//
// build kernel shape: a CTA of eight warps owns 32 fits (lane = fit).  Rows are walked eight at a time: warp w
// produces row 8b + w of each of its 32 fits (model values, the 19-term dot product of the update, the corrected
// row) into a shared-memory stage, then every warp accumulates ITS EIGHTH OF THE 190 + 19 SUMS over the eight rows
// (the lower triangle cut into 5 triangles and 10 rectangles along column groups of four, two per warp), statically
// unrolled into registers.  The reference's summation orders are loop orders
// here: per-entry partial sums over 32-row blocks joined in ascending block order (misc_core.c:101-126; joined by
// red.global.add.f64 at the L2 like the der path), J^T e one chain per column over ascending rows, the unblocked
// descending loop when n*m <= 1024 (lm_core.c:585, "<=" in dif).  Bit-identical to the warp-per-fit kernel and to the
// oracle / the reference levmar (tests/test_lm_dif_staged_gpu.py).
#pragma once
#include "cps_lm_staged.cuh"

namespace cps {

struct DifScal {  // per-fit scalars of a dlevmar_dif iteration in flight
  double e0, e_sq, g_inf, dp_sq, mu, p_sq, maxdiag, dL, pend_dp_sq;
  int nu, k, nfev, njap, nlss, n_jtj, n_broyden, updjac, updp, pend, accepted, pad;
};

constexpr int kDifK = 19;            // K = max(m, 10), lm_core.c:497
constexpr int kDifRow = 20;          // doubles per stored Jacobian row (19 + 1: rows stay 16-byte aligned)
constexpr int kDifStageRow = 22;     // shared-memory pitch of a staged row: 11 x 16 B (odd: LDS.128 of a quarter-warp
                                     // hit distinct 16-byte slots); [0..18] J, [20] residual
constexpr int kDifFits = 32;         // fits per CTA of the build kernels
constexpr int kDifThreads = 256;     // eight warps
constexpr int kDifSums = kLower19 + 19;  // running sums of a fit in the L2 work area: 190 J^T J entries, 19 J^T e

struct DifArgs {
  LmLaunch L;
  DifScal* scal;        // [B]
  double* jtj;          // [B][190] lower triangle, row-major
  double* jte;          // [B][19]
  double* pdp;          // [B][19] trial point p + dp
  double* dp;           // [B][19]
  double* pold;         // [B][19] p at the time of the pending secant update
  double* tws;          // [B][n_max/2] t when the caller passes none
  double* J;            // [B][j_rows][20] stored secant Jacobians
  size_t j_stride;      // doubles per fit in J
  double* ws;           // [kDifSums][ws_stride] running sums of the resident fits: first half of a row the update
                        // kernel's CTAs, second half the finite-difference kernel's (they may run concurrently)
  int ws_stride;
  int* fd_list;         // fits whose next build evaluates a finite-difference Jacobian
  int* upd_list;        // fits whose next build applies a pending secant update
  int* eval_list;
  int* solve_list[2];
  unsigned int* cnt;    // [0] fd list, [1] upd list, [2], [3] solve lists
  int init_first, init_count;
  const int* init_order;  // nullptr, or the fits of that range ordered by length (cps_lm_staged.cuh, staged_bucket_*)
};

__device__ __forceinline__ void dif_finish(const DifArgs& A, int f, const DifScal& s, int stop) {
  if (s.k >= A.L.itmax) stop = 3;  // lm_core.c:796
  double* info = A.L.info + (size_t)f * 10;
  info[0] = s.e0;
  info[1] = s.e_sq;
  info[2] = s.g_inf;
  info[3] = s.dp_sq;
  info[4] = s.mu / s.maxdiag;
  info[5] = (double)s.k;
  info[6] = (double)stop;
  info[7] = (double)s.nfev;
  info[8] = (double)s.njap;
  info[9] = (double)s.nlss;
  A.L.ret[f] = (stop != 4 && stop != 7) ? s.k : -1;
  if (A.L.extra) { A.L.extra[2 * f] = s.n_jtj; A.L.extra[2 * f + 1] = s.n_broyden; }
}

__device__ __forceinline__ void dif_geom(const DifArgs& A, int f, FitGeom& g, long long& o0) {
  o0 = A.L.off[f];
  g.n = (int)(A.L.off[f + 1] - o0);
  g.nsamp = g.n >> 1;
  const int c0 = A.L.counts[4 * (size_t)f], c1 = A.L.counts[4 * (size_t)f + 1], c2 = A.L.counts[4 * (size_t)f + 2];
  g.b1 = c0;
  g.b2 = c0 + c1;
  g.b3 = c0 + c1 + c2;
}

__device__ __forceinline__ const double* dif_t(const DifArgs& A, int f, long long o0) {
  return A.L.t ? A.L.t + (o0 >> 1) : A.tws + (size_t)f * (A.L.n_max >> 1);
}

// Top of an outer iteration (lm_core.c:557-579) for a fit whose k has just been advanced: the loop bound, the
// ||e||^2 <= eps3 stop, then the Jacobian policy.  Returns 0: stopped (info written), 1: needs a finite-difference
// Jacobian, 2: needs a rebuild with the pending secant update, 3: straight to the solve with the J^T J it has.
__device__ __forceinline__ int dif_loop_top(const DifArgs& A, int f, DifScal& s) {
  if (s.k >= A.L.itmax) { dif_finish(A, f, s, 0); return 0; }
  if (s.e_sq <= A.L.eps3) { dif_finish(A, f, s, 6); return 0; }
  if ((s.updp && s.nu > 16) || s.updjac == kDifK) {
    s.njap += 1;
    s.nfev += A.L.forward ? 19 : 38;
    s.nu = 2; s.updjac = 0; s.updp = 0;
    s.pend = 0;  // the fresh Jacobian overwrites every entry a deferred secant update would have touched
    return 1;
  }
  return s.pend ? 2 : 3;
}

// ---- eval: thread per fit ---------------------------------------------------------------------------------------
// Same sample staging and the same four-chain ||e||^2 as staged_eval_kernel (cps_lm_staged.cuh); the iteration logic
// is dlevmar_dif's.
template <bool INIT>
__global__ void __launch_bounds__(kEvalThreads, 3) dif_eval_kernel(const DifArgs A, int nxt) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double* Csm = smem + tid;
  WarpTiles* wt = reinterpret_cast<WarpTiles*>(smem + kEvalTableDoubles * kEvalThreads) + wid;
  const unsigned int total = INIT ? (unsigned int)A.init_count : A.cnt[2 + (nxt ^ 1)];
  for (unsigned int base = (blockIdx.x * (kEvalThreads / 32) + wid) * 32; base < total;
       base += gridDim.x * kEvalThreads) {
    const unsigned int item = base + lane;
    bool have = item < total;
    int f = have ? (INIT ? (A.init_order ? A.init_order[item] : A.init_first + (int)item) : A.eval_list[item]) : 0;
    if (f < 0) { have = false; f = 0; }
    FitGeom g;
    g.n = 0; g.nsamp = 0; g.b1 = g.b2 = g.b3 = 0;
    long long o0 = 0;
    if (have) dif_geom(A, f, g, o0);
    if (INIT && have) {
      const long long nn = A.L.off[f + 1] - o0;
      int ret_code = 0;
      if (nn < 19) ret_code = -1;  // lm_core.c:493
      else if (nn > A.L.n_max || (nn & 1)) ret_code = -2;
      const int c3 = A.L.counts[4 * (size_t)f + 3];
      if (ret_code == 0 && (g.b1 < 0 || g.b2 < g.b1 || g.b3 < g.b2 || c3 < 0 || g.b3 + c3 != g.nsamp)) ret_code = -3;
      if (ret_code != 0) {
        for (int q = 0; q < 10; ++q) A.L.info[(size_t)f * 10 + q] = 0.0;
        A.L.ret[f] = ret_code;
        if (A.L.extra) { A.L.extra[2 * f] = 0; A.L.extra[2 * f + 1] = 0; }
        have = false;
      }
    }
    if (!have) { g.n = 0; g.nsamp = 0; }
    const double* x = A.L.meas + o0;
    const bool derive_t = INIT && !A.L.t;
    const double* tp = have ? dif_t(A, f, o0) : nullptr;
    const double* pp = INIT ? A.L.p + (size_t)f * 19 : A.pdp + (size_t)f * 19;
    const int xoff = ptr_parity(x), toff = tp ? ptr_parity(tp) : 0;
    double firsty[4] = {0, 0, 0, 0}, spany[4] = {1, 1, 1, 1};
    if (have) {
      Trig tg;
      make_trig(pp[16], pp[17], tg);
      const double z = pp[18];
#pragma unroll 1
      for (int cp = 0; cp < 8; ++cp) {
        double xb[3];
        body_point(tg, pp[2 * cp], pp[2 * cp + 1], z, xb);
        Csm[(3 * cp + 0) * kEvalThreads] = CPS_FX * (xb[1] + 0.5 * CPS_BASELINE) / xb[0] + CPS_CX;
        Csm[(3 * cp + 1) * kEvalThreads] = CPS_FX * (xb[1] - 0.5 * CPS_BASELINE) / xb[0] + CPS_CX;
        Csm[(3 * cp + 2) * kEvalThreads] = CPS_FY * (xb[2] / xb[0]) + CPS_CY;
      }
      if (derive_t) {
        const int bnd[5] = {0, g.b1, g.b2, g.b3, g.nsamp};
#pragma unroll
        for (int sgi = 0; sgi < 4; ++sgi)
          if (bnd[sgi + 1] > bnd[sgi]) {
            firsty[sgi] = x[2 * bnd[sgi] + 1];
            spany[sgi] = x[2 * bnd[sgi + 1] - 1] - firsty[sgi];
          }
      }
    }
    const int nblk = (g.n + 31) >> 5;
    const int full = (g.n >> 3) << 3;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const double* xt = &wt->x[lane][xoff];
    double* tt = &wt->t[lane][toff];
    // the eight projected control-point coordinates of the current contour segment stay in registers: they change
    // three times per fit, and reading them from the table for every sample was most of this kernel's shared-memory
    // traffic (same values, same operations)
    int cur_seg = -1;
    double cu[4] = {0, 0, 0, 0}, cv[4] = {0, 0, 0, 0};
    auto resid = [&](int s, double tval, double xu, double xv, double& eu, double& ev) {
      const int seg = (s >= g.b1) + (s >= g.b2) + (s >= g.b3);
      if (seg != cur_seg) {
        cur_seg = seg;
        const int cb = 12 * (seg & 1), side = seg >> 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          cu[k] = Csm[(cb + 3 * k + side) * kEvalThreads];
          cv[k] = Csm[(cb + 3 * k + 2) * kEvalThreads];
        }
      }
      double w[4];
      bernstein3(tval, w);
      double au = 0.0, av = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        au += w[k] * cu[k];
        av += w[k] * cv[k];
      }
      eu = xu - au;
      ev = xv - av;
    };
#pragma unroll 1
    for (int b = nblk - 1; b >= 0; --b) {
      row_fill(&wt->x[lane][0], xoff, x + 32 * b, min(32, g.n - 32 * b));
      if (!derive_t) row_fill(&wt->t[lane][0], toff, tp + 16 * b, min(16, g.nsamp - 16 * b));
      if (b > 0) {
        prefetch_l2(x + 32 * (b - 1));
        prefetch_l2(x + 32 * (b - 1) + 16);
        if (!derive_t) prefetch_l2(tp + 16 * (b - 1));
      }
      cp_async_wait_all();
      {
        if (derive_t) {  // fit_curve:451-511
          const int s_hi = min(16 * b + 16, g.nsamp);
          for (int s = 16 * b; s < s_hi; ++s) {
            const int seg = (s >= g.b1) + (s >= g.b2) + (s >= g.b3);
            const double fy = seg == 0 ? firsty[0] : (seg == 1 ? firsty[1] : (seg == 2 ? firsty[2] : firsty[3]));
            const double sy = seg == 0 ? spany[0] : (seg == 1 ? spany[1] : (seg == 2 ? spany[2] : spany[3]));
            double v = (xt[2 * (s & 15) + 1] - fy) / sy;
            v = (v < 0.0) ? 0.0 : v;
            v = (v > 1.0) ? 1.0 : v;
            tt[s & 15] = v;
          }
        }
        const int g_hi = min(4 * b + 3, (full >> 3) - 1);
#pragma unroll 1
        for (int gq = g_hi; gq >= 4 * b; --gq) {
          const int sa = 4 * gq, l = sa & 15;
          double eu0, ev0, eu1, ev1, eu2, ev2, eu3, ev3;
          resid(sa + 3, tt[l + 3], xt[2 * l + 6], xt[2 * l + 7], eu3, ev3);
          resid(sa + 2, tt[l + 2], xt[2 * l + 4], xt[2 * l + 5], eu2, ev2);
          resid(sa + 1, tt[l + 1], xt[2 * l + 2], xt[2 * l + 3], eu1, ev1);
          resid(sa, tt[l], xt[2 * l], xt[2 * l + 1], eu0, ev0);
          s0 += ev3 * ev3;
          s1 += eu3 * eu3;
          s2 += ev2 * ev2;
          s3 += eu2 * eu2;
          s0 += ev1 * ev1;
          s1 += eu1 * eu1;
          s2 += ev0 * ev0;
          s3 += eu0 * eu0;
        }
      }
      if (derive_t) {
        double* tw = const_cast<double*>(tp) + 16 * b;
        const int cnt = min(16, g.nsamp - 16 * b);
        int i = 0;
        for (; i + 2 <= cnt; i += 2) *reinterpret_cast<double2*>(tw + i) = make_double2(tt[i], tt[i + 1]);
        if (i < cnt) tw[i] = tt[i];
      }
    }
    int route = 0;  // 1: fd list, 2: upd list, 3: solve list
    if (have) {
      const int rem = g.n - full;
      for (int k = 0; k < rem; k += 2) {
        const int s = (full + k) >> 1;
        double eu, ev;
        resid(s, tp[s], x[2 * s], x[2 * s + 1], eu, ev);
        const int cu = (7 - (rem - k)) & 3, cv = (7 - (rem - k - 1)) & 3;
        const double du = eu * eu, dv = ev * ev;
        if (cu == 0) s0 += du; else if (cu == 1) s1 += du; else if (cu == 2) s2 += du; else s3 += du;
        if (cv == 0) s0 += dv; else if (cv == 1) s1 += dv; else if (cv == 2) s2 += dv; else s3 += dv;
      }
      const double e_new = ((s0 + s1) + s2) + s3;
      DifScal* sp = A.scal + f;
      if (INIT) {
        DifScal s;
        s.e0 = e_new; s.e_sq = e_new; s.g_inf = 0.0; s.dp_sq = DBL_MAX; s.mu = 0.0; s.p_sq = 0.0; s.maxdiag = 0.0;
        s.dL = 0.0; s.pend_dp_sq = 1.0;
        s.nu = 20;  // forces a Jacobian on the first pass, lm_core.c:555
        s.k = 0; s.nfev = 1; s.njap = 0; s.nlss = 0; s.n_jtj = 0; s.n_broyden = 0;
        s.updjac = 0; s.updp = 1; s.pend = 0; s.accepted = 0; s.pad = 0;
        if (!isfinite(e_new)) { dif_finish(A, f, s, 7); }
        else route = dif_loop_top(A, f, s);
        *sp = s;
      } else {
        DifScal s = *sp;
        s.nfev += 1;
        if (!isfinite(e_new)) {  // lm_core.c:739-742
          *sp = s;
          dif_finish(A, f, s, 7);
        } else {
          const double dF = s.e_sq - e_new;
          if (s.updp || dF > 0.0) {
            // rank-one secant update (lm_core.c:745-755), deferred to the rebuild that always precedes the next solve:
            // it needs p (f(p) = hx), p + dp (f = wrk), dp and ||dp||^2 of this iteration
            double* po = A.pold + (size_t)f * 19;
            const double* pc = A.L.p + (size_t)f * 19;
#pragma unroll
            for (int q = 0; q < 19; ++q) po[q] = pc[q];
            s.pend = 1;
            s.pend_dp_sq = s.dp_sq;
            s.updjac += 1;
            s.n_broyden += 1;
          }
          bool go_on = true;
          if (s.dL > 0.0 && dF > 0.0) {
            s.mu = nielsen(s.mu, dF, s.dL);
            s.nu = 2;
            double* pout = A.L.p + (size_t)f * 19;
#pragma unroll
            for (int q = 0; q < 19; ++q) pout[q] = pp[q];
            s.e_sq = e_new;
            s.updp = 1;
            s.accepted = 1;
          } else {
            s.accepted = 0;
            s.mu *= s.nu;
            const int nu2 = (int)((unsigned)s.nu << 1);
            if (nu2 <= s.nu) {
              *sp = s;
              dif_finish(A, f, s, 5);
              go_on = false;
            } else {
              s.nu = nu2;
            }
          }
          if (go_on) {
            s.k += 1;
            route = dif_loop_top(A, f, s);
            *sp = s;
          }
        }
      }
    }
    __syncwarp();
    staged_append(A.fd_list, A.cnt + 0, f, route == 1);
    staged_append(A.upd_list, A.cnt + 1, f, route == 2);
    staged_append(A.solve_list[nxt], A.cnt + 2 + nxt, f, route == 3);
  }
}

// ---- solve: thread per fit ---------------------------------------------------------------------------------------
// dlevmar_dif has no inner loop: a solve that fails (singular matrix) is a rejected step -- mu grows, k advances,
// the loop top runs -- and unless that asks for a fresh Jacobian the same J^T J is solved again, here in place.
__global__ void __launch_bounds__(kSolveThreads, 1) dif_solve_kernel(const DifArgs A, int cur) {
  constexpr int M = 19, T = kSolveThreads;
  extern __shared__ double smem[];
  double* a = smem + threadIdx.x;
  double* sc = smem + M * M * T + threadIdx.x;
  double* bx = smem + (M * M + M) * T + threadIdx.x;
  const unsigned int total = A.cnt[2 + cur];
  for (unsigned int i = blockIdx.x * T + threadIdx.x; i < total; i += gridDim.x * T) {
    const int f = A.solve_list[cur][i];
    const double* lower = A.jtj + (size_t)f * kLower19;
    const double* jte = A.jte + (size_t)f * M;
    const double* pp = A.L.p + (size_t)f * M;
    DifScal s = A.scal[f];
    int out = -1;  // eval list entry
    bool to_fd = false;
    double dp[M];
    for (;;) {
      lu_fill_global19(a, bx, lower, jte, s.mu);
      const int solved = lu_factor_solve<M, T>(a, sc, bx, dp);
      s.nlss += 1;
      if (solved) {
        double dp_sq = 0.0;
#pragma unroll
        for (int q = 0; q < M; ++q) dp_sq += dp[q] * dp[q];
        s.dp_sq = dp_sq;
        if (dp_sq <= A.L.eps2_sq * s.p_sq) { dif_finish(A, f, s, 2); break; }
        if (dp_sq >= (s.p_sq + A.L.eps2) / (kEpsilon * kEpsilon)) { dif_finish(A, f, s, 4); break; }
        double pv[M], gv[M];
#pragma unroll
        for (int q = 0; q < M; ++q) { pv[q] = pp[q]; gv[q] = jte[q]; }
        double dL = 0.0;
        double* pdp = A.pdp + (size_t)f * M;
        double* dpo = A.dp + (size_t)f * M;
#pragma unroll
        for (int q = 0; q < M; ++q) {
          pdp[q] = pv[q] + dp[q];
          dpo[q] = dp[q];
          dL += dp[q] * (s.mu * dp[q] + gv[q]);
        }
        s.dL = dL;
        out = f;
        break;
      }
      // rejected without a step, lm_core.c:783-792
      s.mu *= s.nu;
      const int nu2 = (int)((unsigned)s.nu << 1);
      if (nu2 <= s.nu) { dif_finish(A, f, s, 5); break; }
      s.nu = nu2;
      s.k += 1;
      const int route = dif_loop_top(A, f, s);
      if (route == 0) break;
      if (route == 1) { to_fd = true; break; }
      // route 3 (nothing pending can exist here): solve the same matrix with the larger mu
    }
    A.scal[f] = s;
    A.eval_list[i] = out;
    if (to_fd) A.fd_list[atomicAdd(A.cnt + 0, 1u)] = f;
  }
}

__global__ void dif_round_end_kernel(unsigned int* cnt, int nxt, volatile unsigned int* host) {
  host[0] = cnt[2 + nxt];
  host[1] = cnt[0];
  host[2] = cnt[1];
  __threadfence_system();
}

// ---- build: eight warps x 32 fits ----------------------------------------------------------------------------------
// The lower triangle of the 19 x 19 matrix is cut along column groups of four (g0..g3: two control points each; g4:
// theta, phi, z) and dealt to the eight warps together with the J^T e chains (table below).  The producing phase and
// the copies are ONE instruction stream for all warps; only the accumulate / fold code is instantiated per role
// (role-specialised copies of the whole sweep overflowed the instruction cache: "no instruction" was the top stall
// reason).
constexpr int kDifWarps = 8;
__device__ __forceinline__ void dif_load4(const double* p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// ---- the 190 + 19 sums dealt to the eight warps in their exact shapes: no padding column, no upper halves of the
// diagonal blocks.  Column groups g0..g3 = four control-point columns each, g4 = theta, phi, z (three columns).
//   warp 0: (1,0) + diag 0        26 sums     warp 4: (3,1) + (3,2)              32
//   warp 1: (2,0) + diag 1        26          warp 5: (4,0) + (4,1) + J^T e g4   27
//   warp 2: (2,1) + diag 2        26          warp 6: (4,2) + (4,3) + diag 4     30
//   warp 3: (3,0) + diag 3        26          warp 7: J^T e g0..g3               16
// (a uniform two-4x4-blocks-per-warp scheme formed 276 products per row for these 209 sums: build 200 -> 170 ms per
// 65 536-fit call).  The eight roles are eight instantiations of the accumulate / fold code only -- a few hundred
// instructions, not eight copies of the sweep.
constexpr int kDifExactSums = 32;
__host__ __device__ constexpr int dif_gn(int G) { return G == 4 ? 3 : 4; }

// visitors over the items of a role: FULL(R, C) = rows of group R x columns of group C, DIAG(G) = lower triangle of
// group G's square, JTE(G) = J^T e of group G's columns; `off` = first sum of the item in the thread's array
#define DIF_ROLE_ITEMS(ROLE, X)                                                      \
  if constexpr (ROLE == 0) { X(FULL, 1, 0, 0) X(DIAG, 0, 0, 16) }                    \
  if constexpr (ROLE == 1) { X(FULL, 2, 0, 0) X(DIAG, 1, 1, 16) }                    \
  if constexpr (ROLE == 2) { X(FULL, 2, 1, 0) X(DIAG, 2, 2, 16) }                    \
  if constexpr (ROLE == 3) { X(FULL, 3, 0, 0) X(DIAG, 3, 3, 16) }                    \
  if constexpr (ROLE == 4) { X(FULL, 3, 1, 0) X(FULL, 3, 2, 16) }                    \
  if constexpr (ROLE == 5) { X(FULL, 4, 0, 0) X(FULL, 4, 1, 12) X(JTE, 4, 4, 24) }   \
  if constexpr (ROLE == 6) { X(FULL, 4, 2, 0) X(FULL, 4, 3, 12) X(DIAG, 4, 4, 24) }  \
  if constexpr (ROLE == 7) { X(JTE, 0, 0, 0) X(JTE, 1, 1, 4) X(JTE, 2, 2, 8) X(JTE, 3, 3, 12) }

enum { DIF_FULL = 0, DIF_DIAG = 1, DIF_JTE = 2 };

template <int ROLE>
__device__ __forceinline__ void dif_acc_exact(const double* row, const double* eptr, double (&s)[kDifExactSums]) {
  // the column groups this role reads (each loaded once), then its items
  double g[5][4];
  constexpr bool need[8][5] = {{1, 1, 0, 0, 0}, {1, 1, 1, 0, 0}, {0, 1, 1, 0, 0}, {1, 0, 0, 1, 0},
                               {0, 1, 1, 1, 0}, {1, 1, 0, 0, 1}, {0, 0, 1, 1, 1}, {1, 1, 1, 1, 0}};
#pragma unroll
  for (int q = 0; q < 5; ++q)
    if (need[ROLE][q]) dif_load4(row + 4 * q, g[q]);
  double e = 0.0;
  if (ROLE == 5 || ROLE == 7) e = eptr[0];
#define DIF_ACC_ITEM(KIND, R, C, OFF)                                                         \
  if (DIF_##KIND == DIF_FULL) {                                                               \
    _Pragma("unroll") for (int i = 0; i < dif_gn(R); ++i)                                     \
      _Pragma("unroll") for (int j = 0; j < 4; ++j) s[OFF + 4 * i + j] += g[R][i] * g[C][j];  \
  } else if (DIF_##KIND == DIF_DIAG) {                                                        \
    _Pragma("unroll") for (int i = 0; i < dif_gn(R); ++i)                                     \
      _Pragma("unroll") for (int j = 0; j <= i; ++j) s[OFF + i * (i + 1) / 2 + j] += g[R][i] * g[R][j]; \
  } else {                                                                                    \
    _Pragma("unroll") for (int i = 0; i < dif_gn(R); ++i) s[OFF + i] += g[R][i] * e;          \
  }
  DIF_ROLE_ITEMS(ROLE, DIF_ACC_ITEM)
#undef DIF_ACC_ITEM
}

// the partial sums of a 32-row block join the running entries (ZERO: the running entries are cleared instead);
// the J^T e chains are not block sums and stay in the registers until the end
template <int ROLE, bool ZERO>
__device__ __forceinline__ void dif_fold_exact(double* ws, size_t wsn, double (&s)[kDifExactSums]) {
  // (the slot addresses are formed here, at every fold: hoisted out of the step loop -- for all eight roles -- they
  // would occupy a few hundred registers' worth of stack)
  asm volatile("" : "+l"(ws), "+l"(wsn));
#define DIF_FOLD_ITEM(KIND, R, C, OFF)                                                        \
  if (DIF_##KIND == DIF_FULL) {                                                               \
    _Pragma("unroll") for (int i = 0; i < dif_gn(R); ++i)                                     \
      _Pragma("unroll") for (int j = 0; j < 4; ++j) {                                         \
        double* slot = ws + (size_t)tri(4 * R + i, 4 * C + j) * wsn;                          \
        if (ZERO) __stcg(slot, 0.0);                                                          \
        else { red_add_f64(slot, s[OFF + 4 * i + j]); s[OFF + 4 * i + j] = 0.0; }             \
      }                                                                                       \
  } else if (DIF_##KIND == DIF_DIAG) {                                                        \
    _Pragma("unroll") for (int i = 0; i < dif_gn(R); ++i)                                     \
      _Pragma("unroll") for (int j = 0; j <= i; ++j) {                                        \
        double* slot = ws + (size_t)tri(4 * R + i, 4 * R + j) * wsn;                          \
        if (ZERO) __stcg(slot, 0.0);                                                          \
        else { red_add_f64(slot, s[OFF + i * (i + 1) / 2 + j]); s[OFF + i * (i + 1) / 2 + j] = 0.0; } \
      }                                                                                       \
  }
  DIF_ROLE_ITEMS(ROLE, DIF_FOLD_ITEM)
#undef DIF_FOLD_ITEM
}

template <int ROLE>
__device__ __forceinline__ void dif_store_jte_exact(double* ws, size_t wsn, const double (&s)[kDifExactSums]) {
  asm volatile("" : "+l"(ws), "+l"(wsn));
#define DIF_JTE_ITEM(KIND, R, C, OFF)                                                         \
  if (DIF_##KIND == DIF_JTE) {                                                                \
    _Pragma("unroll") for (int i = 0; i < dif_gn(R); ++i) __stcg(ws + (size_t)(kLower19 + 4 * R + i) * wsn, s[OFF + i]); \
  }
  DIF_ROLE_ITEMS(ROLE, DIF_JTE_ITEM)
#undef DIF_JTE_ITEM
}

// run-time warp index -> compile-time role
#define DIF_DISPATCH(wid, CALL)        \
  switch (wid) {                       \
    case 0: { CALL(0) } break;         \
    case 1: { CALL(1) } break;         \
    case 2: { CALL(2) } break;         \
    case 3: { CALL(3) } break;         \
    case 4: { CALL(4) } break;         \
    case 5: { CALL(5) } break;         \
    case 6: { CALL(6) } break;         \
    default: { CALL(7) } break;        \
  }

// projected control points of a parameter vector: tab[(3*cp + {uL, uR, v}) * 32] (this fit's column of a [.][32] table)
__device__ __forceinline__ void dif_project_all(const double* pp, double* tab) {
  Trig tg;
  make_trig(pp[16], pp[17], tg);
  const double z = pp[18];
#pragma unroll 1
  for (int cp = 0; cp < 8; ++cp) {
    double xb[3];
    body_point(tg, pp[2 * cp], pp[2 * cp + 1], z, xb);
    tab[(3 * cp + 0) * kDifFits] = CPS_FX * (xb[1] + 0.5 * CPS_BASELINE) / xb[0] + CPS_CX;
    tab[(3 * cp + 1) * kDifFits] = CPS_FX * (xb[1] - 0.5 * CPS_BASELINE) / xb[0] + CPS_CX;
    tab[(3 * cp + 2) * kDifFits] = CPS_FY * (xb[2] / xb[0]) + CPS_CY;
  }
}

__device__ __forceinline__ void dif_project_one(const Trig& tg, double xe, double ye, double z, double* out3) {
  double xb[3];
  body_point(tg, xe, ye, z, xb);
  out3[0 * kDifFits] = CPS_FX * (xb[1] + 0.5 * CPS_BASELINE) / xb[0] + CPS_CX;
  out3[1 * kDifFits] = CPS_FX * (xb[1] - 0.5 * CPS_BASELINE) / xb[0] + CPS_CX;
  out3[2 * kDifFits] = CPS_FY * (xb[2] / xb[0]) + CPS_CY;
}

// Shared-memory tables per fit, [entry][32 fits]:
//   update build: 0..23 control points of p_old (f = hx), 24..47 of p + dp (f = wrk), 48..66 dp
//   FD build:     0..23 control points of p, 24..71 the perturbed control point of columns 0..15 (3 values each),
//                 72..143 all control points for theta+d, phi+d, z+d, 144..162 1/d
constexpr int kDifTabUpd = 67;
constexpr int kDifTabFd = 163;

constexpr int kDifXtPitch = 13;  // per fit and step: 8 samples + 4 t (+1: an odd pitch spreads the banks)
// The stage of a step, [fit][8 rows][20] with a pitch of 162 doubles per fit: the eight rows of a fit are 1 280 contiguous
// bytes in global memory AND in the stage, so they come and go as ONE bulk copy per fit (cp.async.bulk, the TMA unit)
// instead of ten 16-byte cp.async and ten load + store pairs per thread; 162 = 8 * 20 + 2 makes a warp's 16-byte
// accesses (lane = fit) fall into distinct banks.  The residuals of the step follow as [8][32].
constexpr int kDifFitPitch = kDifWarps * kDifRow + 2;
constexpr int kDifStepDoubles = kDifFits * kDifFitPitch + kDifWarps * kDifFits;
template <bool FD>
constexpr int dif_build_smem_doubles() {
  return (FD ? kDifTabFd : kDifTabUpd) * kDifFits + (FD ? 1 : 2) * kDifStepDoubles + kDifFits + 2 * kDifFits * kDifXtPitch + 2;
}

__device__ __forceinline__ unsigned int dif_smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dif_mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dif_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dif_mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dif_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dif_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dif_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dif_mbar_wait(unsigned long long* bar, unsigned int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DIF_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DIF_WAIT_DONE;\n"
      "bra DIF_WAIT_LOOP;\n"
      "DIF_WAIT_DONE:\n"
      "}\n" ::"r"(dif_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void dif_bulk_load(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dif_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(dif_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dif_bulk_store(void* dst, const void* src, unsigned int bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(dif_smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void dif_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void dif_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one model value: sum_k w_k * tab[(12*curve + 3k + sel) * 32], ascending k from 0.0 like closest_bezier_pt(cp, t, ..)
__device__ __forceinline__ double dif_bez(const double (&w)[4], const double* tab, int cb_sel) {
  double a = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) a += w[k] * tab[(cb_sel + 3 * k) * kDifFits];
  return a;
}

// per-thread view of the group of 32 fits a CTA is working on
struct DifCtx {
  const DifArgs* A;
  double* tab;     // this lane's column of the [entry][32] tables
  double* stage;   // [NBUF][8][32][22]
  double* xt;      // [2][32][13] the step's 8 samples and 4 t of every fit
  unsigned long long* mbar;  // [2] arrival of a buffer's bulk loads
  unsigned int phase0, phase1;
  const int* nrows;
  const double* x;
  const double* tp;
  double* ws;
  size_t wsn;
  FitGeom g;
  int lane, cfi, csub;
  bool have, blocked, accepted;
  double pend_dp_sq;
};

// The row sweep of one group: the same code for every warp, its blocks selected by the role.
template <bool FD>
__device__ __forceinline__ void dif_group_sweep(DifCtx& C, const int wid) {
  constexpr int STEP_DOUBLES = kDifStepDoubles;
  const DifArgs& A = *C.A;
  double* const tab = C.tab;
  double* const stage = C.stage;
  double* const xt = C.xt;
  unsigned long long* const mbar = C.mbar;
  const int* const nrows = C.nrows;
  double* const ws = C.ws;
  const size_t wsn = C.wsn;
  const FitGeom g = C.g;
  const int lane = C.lane, cfi = C.cfi, csub = C.csub;
  const bool have = C.have, blocked = C.blocked, accepted = C.accepted;
  const double pend_dp_sq = C.pend_dp_sq;
    // running sums start at zero (every warp its own entries)
    double S[kDifExactSums];
#pragma unroll
    for (int q = 0; q < kDifExactSums; ++q) S[q] = 0.0;
#define DIF_CALL_ZERO(ROLE) dif_fold_exact<ROLE, true>(ws, wsn, S);
    DIF_DISPATCH(wid, DIF_CALL_ZERO)
#undef DIF_CALL_ZERO
    __syncthreads();
    int nmax = 0;
    bool any_blocked = false, any_small = false;
    for (int q = 0; q < kDifFits; ++q) {
      const int nq = nrows[q];
      nmax = max(nmax, nq);
      if (nq > 0) {
        if (nq * 19 <= kBlockSzSq) any_small = true; else any_blocked = true;
      }
    }
    const int nsteps = (nmax + kDifWarps - 1) / kDifWarps;
    // this thread's share of the cooperative copies: fit cfi of the group
    const int cn = nrows[cfi];
    // samples and t of the copied fit (update build): the eight samples of a step are 64 contiguous bytes, its four
    // t values 32; read per lane they cost a 32-byte sector for every 8 bytes and a global latency in the producing phase
    const double* cx = nullptr;
    const double* ct = nullptr;
    if (cn > 0) {
      const int fc = nrows[kDifFits + cfi];
      const long long oc = A.L.off[fc];
      cx = A.L.meas + oc;
      ct = dif_t(A, fc, oc);
    }

    auto issue_xt = [&](int st, int buf) {
      double* xb = xt + ((size_t)buf * kDifFits + cfi) * kDifXtPitch;
      const int rx = kDifWarps * st + csub;
      const bool okx = cx != nullptr && rx < cn;
      const unsigned int dx = (unsigned int)__cvta_generic_to_shared(xb + csub);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dx), "l"(okx ? cx + rx : A.L.meas), "r"(okx ? 8 : 0) : "memory");
      if (csub < 4) {
        const int sx = 4 * st + csub;
        const bool okt = cx != nullptr && 2 * sx < cn;
        const unsigned int dt = (unsigned int)__cvta_generic_to_shared(xb + 8 + csub);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dt), "l"(okt ? ct + sx : A.L.meas), "r"(okt ? 8 : 0) : "memory");
      }
    };
    // the first warp moves the rows: thread `tid` < 32 owns fit `tid` of the group
    const int tid = threadIdx.x;
    const int bn = tid < kDifFits ? nrows[tid] : 0;
    double* const bJ = A.J + (size_t)(tid < kDifFits ? nrows[kDifFits + tid] : 0) * A.j_stride;
    unsigned int ph0 = C.phase0, ph1 = C.phase1;  // parities of the two buffers' barriers (they persist across groups)
    auto issue_rows = [&](int st, int buf, int bn_eff) {
      if (tid < kDifFits) {
        dif_bulk_wait_read();  // the store that last read this buffer has let go of it
        const int rows = min(kDifWarps, bn_eff - kDifWarps * st);
        if (rows > 0) {
          dif_mbar_expect_tx(mbar + buf, (unsigned int)rows * kDifRow * 8);
          dif_bulk_load(stage + (size_t)buf * STEP_DOUBLES + tid * kDifFitPitch, bJ + (size_t)kDifWarps * st * kDifRow,
                        (unsigned int)rows * kDifRow * 8, mbar + buf);
        } else {
          dif_mbar_arrive(mbar + buf);
        }
      }
      issue_xt(st, buf);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto write_back = [&](int st, const double* sb, int bn_eff) {
      if (tid < kDifFits) {
        const int rows = min(kDifWarps, bn_eff - kDifWarps * st);
        if (rows > 0)
          dif_bulk_store(bJ + (size_t)kDifWarps * st * kDifRow, sb + tid * kDifFitPitch, (unsigned int)rows * kDifRow * 8);
      }
    };

    // The unblocked small-problem loop walks the rows downwards (lm_core.c:604-628), the blocked one upwards.  A
    // group that holds fits of both kinds (rare: fits with fewer than 27 samples) runs both passes, each fit taking
    // part in its own.
    for (int pass = 0; pass < 2; ++pass) {
      const bool down = pass == 1;
      if (down ? !any_small : !any_blocked) continue;
      const bool mine = have && (blocked != down);
      const int n_eff = mine ? g.n : 0;                     // rows of this lane's fit that take part in this pass
      const bool bmine = bn > 0 && ((bn * 19 <= kBlockSzSq) == down);
      const int bn_eff = bmine ? bn : 0;                    // the same for the fit whose rows this thread moves
      if (!FD) issue_rows(down ? nsteps - 1 : 0, 0, bn_eff);
      else { issue_xt(down ? nsteps - 1 : 0, 0); asm volatile("cp.async.commit_group;" ::: "memory"); }
      // sample and t of this lane's row of the first step

      for (int it = 0; it < nsteps; ++it) {
        const int st = down ? nsteps - 1 - it : it;
        const int buf = FD ? 0 : (it & 1);
        double* sb = stage + (size_t)buf * STEP_DOUBLES;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (FD && tid < kDifFits) dif_bulk_wait_read();  // the one stage buffer: last step's store has read it
        __syncthreads();  // [A] this step's samples have landed; everybody is done with the other buffer
        if (!FD) {
          if (buf == 0) { dif_mbar_wait(mbar + 0, ph0); ph0 ^= 1u; } else { dif_mbar_wait(mbar + 1, ph1); ph1 ^= 1u; }
        }
        if (!FD && it + 1 < nsteps) issue_rows(down ? st - 1 : st + 1, buf ^ 1, bn_eff);
        if (FD && it + 1 < nsteps) { issue_xt(down ? st - 1 : st + 1, (it & 1) ^ 1); asm volatile("cp.async.commit_group;" ::: "memory"); }
        // ---- produce row r = 8 st + wid of this lane's fit
        {
          const int r = kDifWarps * st + wid;
          double* row = sb + lane * kDifFitPitch + wid * kDifRow;
          double* erow = sb + kDifFits * kDifFitPitch + wid * kDifFits + lane;
          const double* xb = xt + ((size_t)(it & 1) * kDifFits + lane) * kDifXtPitch;
          const double xr = xb[wid], tval = xb[8 + (wid >> 1)];
          if (r < n_eff) {
            const int sidx = r >> 1, uv = r & 1;
            const int seg = (sidx >= g.b1) + (sidx >= g.b2) + (sidx >= g.b3);
            const int curve = seg & 1, side = seg >> 1;
            const int cb_sel = 12 * curve + (uv ? 2 : side);
            double w[4];
            bernstein3(tval, w);
            if (FD) {
              const double hx = dif_bez(w, tab, cb_sel);
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const int cp = j >> 1;
                double v0 = 0.0, v1 = 0.0;  // (hx - hx) * (1/d) = +0.0 for the other curve's control points
                if ((cp >> 2) == curve) {
                  const int kk = cp & 3;
                  double a0 = 0.0, a1 = 0.0;
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const double cbase = tab[(cb_sel + 3 * k) * kDifFits];
                    const double c0 = (k == kk) ? tab[(24 + 3 * j + (uv ? 2 : side)) * kDifFits] : cbase;
                    const double c1 = (k == kk) ? tab[(24 + 3 * (j + 1) + (uv ? 2 : side)) * kDifFits] : cbase;
                    a0 += w[k] * c0;
                    a1 += w[k] * c1;
                  }
                  v0 = (a0 - hx) * tab[(144 + j) * kDifFits];
                  v1 = (a1 - hx) * tab[(144 + j + 1) * kDifFits];
                }
                *reinterpret_cast<double2*>(row + j) = make_double2(v0, v1);
              }
              double vq[3];
#pragma unroll
              for (int q = 0; q < 3; ++q) {
                const double hxx = dif_bez(w, tab + (72 + 24 * q) * kDifFits, cb_sel);
                vq[q] = (hxx - hx) * tab[(160 + q) * kDifFits];
              }
              *reinterpret_cast<double2*>(row + 16) = make_double2(vq[0], vq[1]);
              *reinterpret_cast<double2*>(row + 18) = make_double2(vq[2], 0.0);
              erow[0] = xr - hx;
            } else {
              const double hx = dif_bez(w, tab, cb_sel);
              const double wrk = dif_bez(w, tab + 24 * kDifFits, cb_sel);
              // rank-one secant update of the row, lm_core.c:745-755: tmp = (wrk_r - hx_r - J_r . dp) / ||dp||^2
              double tmp = 0.0;
#pragma unroll 3
              for (int c = 0; c < 18; c += 2) {
                const double2 t2 = *reinterpret_cast<const double2*>(row + c);
                tmp += t2.x * tab[(48 + c) * kDifFits];
                tmp += t2.y * tab[(48 + c + 1) * kDifFits];
              }
              tmp += row[18] * tab[(48 + 18) * kDifFits];
              tmp = ((wrk - hx) - tmp) / pend_dp_sq;
#pragma unroll 3
              for (int c = 0; c < 18; c += 2) {
                double2 t2 = *reinterpret_cast<const double2*>(row + c);
                t2.x += tmp * tab[(48 + c) * kDifFits];
                t2.y += tmp * tab[(48 + c + 1) * kDifFits];
                *reinterpret_cast<double2*>(row + c) = t2;
              }
              const double last = row[18] + tmp * tab[(48 + 18) * kDifFits];
              *reinterpret_cast<double2*>(row + 18) = make_double2(last, 0.0);
              erow[0] = xr - (accepted ? wrk : hx);
            }
          } else {
#pragma unroll
            for (int c = 0; c < kDifRow; c += 2) *reinterpret_cast<double2*>(row + c) = make_double2(0.0, 0.0);
            erow[0] = 0.0;
          }
        }
        dif_fence_async_smem();  // the corrected rows are about to be read by the bulk stores
        __syncthreads();  // [B] the eight rows of every fit are complete
        write_back(st, sb, bn_eff);
        // ---- every warp: its share of the sums over the eight rows, in the pass's row order
        // end of a 32-row block (blocked pass) or of the fit: the partial sums join the running entries
        const bool fold = down ? (it == nsteps - 1) : ((st & 3) == 3 || it == nsteps - 1);
        {
          const double* row0 = sb + lane * kDifFitPitch + (down ? kDifWarps - 1 : 0) * kDifRow;
          const double* e0 = sb + kDifFits * kDifFitPitch + (down ? kDifWarps - 1 : 0) * kDifFits + lane;
          const int rstep = (down ? -1 : 1) * kDifRow, estep = (down ? -1 : 1) * kDifFits;
#define DIF_CALL_ACC(ROLE)                                                                      \
  _Pragma("unroll 2") for (int q8 = 0; q8 < kDifWarps; ++q8) dif_acc_exact<ROLE>(row0 + q8 * rstep, e0 + q8 * estep, S); \
  if (fold) dif_fold_exact<ROLE, false>(ws, wsn, S);
          DIF_DISPATCH(wid, DIF_CALL_ACC)
#undef DIF_CALL_ACC
        }
      }
      __syncthreads();  // the last step's stage is no longer read (the next pass refills it)
    }
#define DIF_CALL_JTE(ROLE) dif_store_jte_exact<ROLE>(ws, wsn, S);
    DIF_DISPATCH(wid, DIF_CALL_JTE)
#undef DIF_CALL_JTE
    if (tid < kDifFits) dif_bulk_wait_read();
    C.phase0 = ph0;
    C.phase1 = ph1;
}

template <bool FD>
__global__ void __launch_bounds__(kDifThreads, 2) dif_build_kernel(const DifArgs A, int nxt) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NTAB = FD ? kDifTabFd : kDifTabUpd;
  constexpr int NBUF = FD ? 1 : 2;
  constexpr int STEP_DOUBLES = kDifStepDoubles;
  double* tab = smem + lane;                                   // tab[entry * 32]
  double* stage = smem + NTAB * kDifFits;                       // [NBUF][8][32][22]
  int* nrows = reinterpret_cast<int*>(stage + NBUF * STEP_DOUBLES);  // [32] n of each fit (0: no fit), [32] fit index
  double* xt = stage + NBUF * STEP_DOUBLES + kDifFits;                // samples / t of the step, two buffers
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(xt + 2 * kDifFits * kDifXtPitch);
  if (tid == 0) {
    dif_mbar_init(mbar + 0, kDifFits);
    dif_mbar_init(mbar + 1, kDifFits);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned int phase0 = 0, phase1 = 0;
  const size_t wsn = (size_t)A.ws_stride;
  const unsigned int total = A.cnt[FD ? 0 : 1];
  const int* list = FD ? A.fd_list : A.upd_list;
  // cooperative copies between the stored Jacobians and the stage: the eight rows of a step are 1280 contiguous bytes
  // of a fit; eight threads move a fit's 80 16-byte pieces, ten each (128 contiguous bytes per octet and instruction)
  const int cfi = tid >> 3, csub = tid & 7;
  for (unsigned int base = blockIdx.x * kDifFits; base < total; base += gridDim.x * kDifFits) {
    const unsigned int item = base + lane;
    const bool have = item < total;
    const int f = have ? list[item] : 0;
    FitGeom g;
    g.n = 0; g.nsamp = 0; g.b1 = g.b2 = g.b3 = 0;
    long long o0 = 0;
    if (have) dif_geom(A, f, g, o0);
    const double* x = A.L.meas + o0;
    const double* tp = have ? dif_t(A, f, o0) : A.L.meas;
    const double* pc = A.L.p + (size_t)f * 19;
    double* ws = A.ws + (FD ? (size_t)(A.ws_stride / 2) : 0) + (size_t)blockIdx.x * kDifFits + lane;
    const bool blocked = !(g.n * 19 <= kBlockSzSq);  // "<=" here, lm_core.c:585
    const bool accepted = have ? (A.scal[f].accepted != 0) : false;
    const double pend_dp_sq = have ? A.scal[f].pend_dp_sq : 1.0;
    __syncthreads();  // the previous group's tables and stage are no longer read
    if (wid == 0) {
      nrows[lane] = g.n;
      nrows[kDifFits + lane] = f;
    }
    // ---- per-fit tables, spread over the four warps
    if (have) {
      if (!FD) {
        if (wid == 0) dif_project_all(A.pold + (size_t)f * 19, tab);
        if (wid == 1) dif_project_all(A.pdp + (size_t)f * 19, tab + 24 * kDifFits);
        if (wid == 2) {
          const double* dpp = A.dp + (size_t)f * 19;
#pragma unroll 1
          for (int q = 0; q < 19; ++q) tab[(48 + q) * kDifFits] = dpp[q];
        }
      } else {
        // forward differences, misc_core.c:135-170: d = max(1e-4 |p_j|, delta); column j evaluated at p + d e_j
        auto delta_of = [&](int j) {
          double d = 1E-04 * pc[j];
          d = fabs(d);
          if (d < A.L.delta) d = A.L.delta;
          return d;
        };
        if (wid == 0) {
          dif_project_all(pc, tab);
#pragma unroll 1
          for (int j = 0; j < 19; ++j) tab[(144 + j) * kDifFits] = 1.0 / delta_of(j);
        } else if (wid == 1) {  // control-point columns: only that control point moves; theta, phi (the trig) do not
          Trig tg;
          make_trig(pc[16], pc[17], tg);
          const double z = pc[18];
#pragma unroll 1
          for (int j = 0; j < 16; ++j) {
            const int cp = j >> 1;
            double xe = pc[2 * cp], ye = pc[2 * cp + 1];
            const double d = delta_of(j);
            if (j & 1) ye = ye + d; else xe = xe + d;
            dif_project_one(tg, xe, ye, z, tab + (24 + 3 * j) * kDifFits);
          }
        } else if (wid <= 4) {  // theta + d (warp 2), phi + d (warp 3), z + d (warp 4)
          {
            const int q = wid - 2;
            double pw[19];
            const double dq = delta_of(16 + q);
#pragma unroll
            for (int i = 0; i < 19; ++i) pw[i] = (i == 16 + q) ? pc[i] + dq : pc[i];
            dif_project_all(pw, tab + (72 + 24 * q) * kDifFits);
          }
        }
      }
    }
    DifCtx C;
    C.A = &A; C.tab = tab; C.stage = stage; C.xt = xt; C.mbar = mbar; C.phase0 = phase0; C.phase1 = phase1; C.nrows = nrows; C.x = x; C.tp = tp; C.ws = ws; C.wsn = wsn; C.g = g;
    C.lane = lane; C.cfi = cfi; C.csub = csub; C.have = have; C.blocked = blocked; C.accepted = accepted;
    C.pend_dp_sq = pend_dp_sq;
    dif_group_sweep<FD>(C, wid);
    phase0 = C.phase0;
    phase1 = C.phase1;
    __threadfence();
    __syncthreads();
    // ---- publish J^T J / J^T e, gradient statistics, stop 1, initial mu (lm_core.c:657-678)
    bool to_solve = false;
    if (have) {
      double* Jout = A.jtj + (size_t)f * kLower19;
      double* Jte = A.jte + (size_t)f * 19;
      // the lower triangle in eight slices of about 24 entries, one per warp
#pragma unroll 1
      for (int q = wid * 24; q < (wid == kDifWarps - 1 ? kLower19 : wid * 24 + 24); ++q) Jout[q] = __ldcg(ws + (size_t)q * wsn);
      if (wid == 0) {
        // every warp's J^T e chains are in the work area by now
        double p_sq = 0.0, g_inf = 0.0, big = -DBL_MAX;
#pragma unroll 1
        for (int q = 0; q < 19; ++q) {
          const double v = __ldcg(ws + (size_t)(kLower19 + q) * wsn);
          Jte[q] = v;
          const double tmp = fabs(v);
          if (g_inf < tmp) g_inf = tmp;
          p_sq += pc[q] * pc[q];
          const double d = __ldcg(ws + (size_t)tri(q, q) * wsn);
          if (d > big) big = d;
        }
        DifScal sc = A.scal[f];
        sc.n_jtj += 1;
        sc.pend = 0;
        sc.g_inf = g_inf;
        sc.p_sq = p_sq;
        sc.maxdiag = big;
        if (g_inf <= A.L.eps1) {
          sc.dp_sq = 0.0;
          A.scal[f] = sc;
          dif_finish(A, f, sc, 1);
        } else {
          if (sc.k == 0) sc.mu = A.L.tau * big;
          A.scal[f] = sc;
          to_solve = true;
        }
      }
    }
    if (wid == 0) {
      __syncwarp();
      staged_append(A.solve_list[nxt], A.cnt + 2 + nxt, f, to_solve);
    }
  }
}

}  // namespace cps
