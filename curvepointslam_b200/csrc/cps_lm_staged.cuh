// cps_lm_staged.cuh -- dlevmar_der for large batches of Stereo19 fits as a pipeline of three phase kernels.
//
// The monolithic kernel (cps_lm_kernel.cuh) gives a warp to a fit and keeps it for the whole fit; every
// phase then runs at the occupancy of the hungriest one and the dominant phase, J^T J, pays for spreading
// one fit's 32-row blocks over lanes (tiles through shared memory, partial sums folded through shuffles).
// With >= tens of thousands of fits in a batch there is a better source of parallelism: the fits
// themselves.  Here every fit is a small state machine kept in HBM and a round of three kernels advances
// all active fits by one linear solve:
//
//   solve  (warp per fit)    (J^T J + mu I) dp = J^T e by the reference's LU (lu_solve of the monolithic
//                            kernel, unchanged), step checks, predicted reduction       lm_core.c:292-345
//   eval   (thread per fit)  f(p+dp), ||x - f||^2 in the reference's four chains, accept / reject,
//                            Nielsen damping, the loop-top checks of the next iteration  lm_core.c:347-388,198
//   jac    (thread per fit)  analytic Jacobian, J^T J as per-32-row-block partial sums, J^T e, gradient
//                            statistics, initial mu                                      lm_core.c:205-287
//
// A thread that owns a whole fit walks the rows in the reference's own order, so the summation orders
// of the parity contract are simply the loop orders: no shuffles, no tiles, 66
// independent multiply-add chains per row in registers.  Compaction lists carry the fits from kernel to
// kernel; a rejected step goes straight back to `solve` with the old J^T J (the inner loop of
// dlevmar_der).  The arithmetic -- every operation and its order -- is the same as in the monolithic
// kernel, so both paths are bit-identical to the oracle and to each other (tests/test_lm_gpu.py).
#pragma once
#include "cps_lm_kernel.cuh"

namespace cps {

struct StagedScal {  // per-fit scalars of the LM iteration, 96 bytes
  double e0, e_sq, g_inf, dp_sq, mu, p_sq, maxdiag, dL;
  int nu, k, nfev, njev, nlss, pad0, pad1, pad2;
};

constexpr int kLower19 = 19 * 20 / 2;  // J^T J is kept as its lower triangle, row-major: (i,j) -> i(i+1)/2 + j
constexpr int kJacThreads = 128;       // jac kernel: threads per CTA (2 CTAs per SM: 256 fits resident per SM)
constexpr int kJacTableDoubles = 40 + 8;  // per thread: derivative table + projected control points of a segment
constexpr int kEvalThreads = 128;
constexpr int kEvalTableDoubles = 24;  // per thread: the 8 projected control points
// per-thread slots of the jac kernel's global (L2-resident, [slot][thread] = coalesced) work area
constexpr int kWsAcc = 0;     // 60 + 60 running J^T J entries of curve 1 / curve 2, then 6 theta/phi/z x theta/phi/z
constexpr int kWsJte = 126;   // 16 J^T e chains of the control-point columns
constexpr int kWsPend = 142;  // 2 x 60 parked block partial sums (mixed blocks only)
constexpr int kWsSlots = 262;
constexpr int kSolveWarps = 4;
constexpr int kSolveWarpDoubles = 19 * 19 * 2 + 8 * 32;  // LU work matrix, J^T J, small vectors

struct StagedArgs {
  LmLaunch L;
  StagedScal* scal;    // [B]
  double* jtj;         // [B][190]
  double* jte;         // [B][19]
  double* pdp;         // [B][19] trial point p + dp
  double* tws;         // [B][n_max/2] t of every sample when the caller passes none
  double* ws;          // [kWsSlots][jac threads] work area of the jac kernel
  int ws_stride;       // number of jac threads
  int* jac_list;
  int* eval_list;
  int* solve_list[2];
  unsigned int* cnt;   // [0] jac, [1] eval, [2], [3] solve lists
};

__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// info[] / ret exactly as lm_core.c:396-409,422; `k >= itmax` overrides the reason like the statement after the
// reference's loop does (also after a stop inside the inner loop, whose break is followed by the for's ++k)
__device__ __forceinline__ void staged_finish(const StagedArgs& A, int f, const StagedScal& s, int stop) {
  if (s.k >= A.L.itmax) stop = 3;
  double* info = A.L.info + (size_t)f * 10;
  info[0] = s.e0;
  info[1] = s.e_sq;
  info[2] = s.g_inf;
  info[3] = s.dp_sq;
  info[4] = s.mu / s.maxdiag;
  info[5] = (double)s.k;
  info[6] = (double)stop;
  info[7] = (double)s.nfev;
  info[8] = (double)s.njev;
  info[9] = (double)s.nlss;
  A.L.ret[f] = (stop != 4 && stop != 7) ? s.k : -1;
  if (A.L.extra) { A.L.extra[2 * f] = s.njev; A.L.extra[2 * f + 1] = 0; }
}

__device__ __forceinline__ void staged_append(int* list, unsigned int* counter, int f) {
  const unsigned int at = atomicAdd(counter, 1u);
  list[at] = f;
}

__device__ __forceinline__ void staged_geom(const StagedArgs& A, int f, FitGeom& g, long long& o0) {
  o0 = A.L.off[f];
  g.n = (int)(A.L.off[f + 1] - o0);
  g.nsamp = g.n >> 1;
  const int c0 = A.L.counts[4 * (size_t)f], c1 = A.L.counts[4 * (size_t)f + 1], c2 = A.L.counts[4 * (size_t)f + 2];
  g.b1 = c0;
  g.b2 = c0 + c1;
  g.b3 = c0 + c1 + c2;
}

__device__ __forceinline__ const double* staged_t(const StagedArgs& A, int f, long long o0) {
  return A.L.t ? A.L.t + (o0 >> 1) : A.tws + (size_t)f * (A.L.n_max >> 1);
}


// ---- warp-transposed staging ---------------------------------------------------------------------------
// A thread that owns a fit streams that fit's samples, so a warp reads 32 different streams: read directly,
// every 8-byte load would touch its own 32-byte sector.  Instead the warp moves one 32-row block (32 doubles of
// x, 16 of t) of each of its 32 fits into shared memory with fully coalesced 256-byte loads -- lane = element,
// loop over fits -- and every lane then reads its own row.  Row pitches 33 / 17 keep both phases conflict-free.
struct WarpTiles {
  double x[32][33];
  double t[32][17];
  const double* xp[32];
  const double* tp[32];
  int n[32];
  int pad[32];
};

// The copies are cp.async (global -> shared without passing through registers): all 48 of a block are in flight
// at once, where plain loads followed by shared stores would have to run one after the other (the compiler cannot
// prove that the global pointers do not alias the tile).
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void tiles_load_x(WarpTiles* wt, int b, int lane) {
  const int idx = 32 * b + lane;
#pragma unroll 8
  for (int j = 0; j < 32; ++j)
    if (idx < wt->n[j]) cp_async8(&wt->x[j][lane], wt->xp[j] + idx);
}
__device__ __forceinline__ void tiles_load_t(WarpTiles* wt, int b, int lane) {
  const int half = lane >> 4, el = lane & 15;
  const int idx = 16 * b + el;
#pragma unroll 8
  for (int j2 = 0; j2 < 16; ++j2) {
    const int fit = 2 * j2 + half;
    if (idx < (wt->n[fit] >> 1)) cp_async8(&wt->t[fit][el], wt->tp[fit] + idx);
  }
}
__device__ __forceinline__ int seg_end(const FitGeom& g, int seg) {
  return seg == 0 ? g.b1 : (seg == 1 ? g.b2 : (seg == 2 ? g.b3 : g.nsamp));
}

// ---- eval: thread per fit -----------------------------------------------------------------------------
// INIT: validates the fit, derives t (fit_curve:451-511), e0 = ||x - f(p)||^2 and runs the loop-top checks
// of iteration 0.  Otherwise: the trial point of the last solve.
// ||e||^2 as dlevmar_L2nrmxmy (misc_core.c:712-798): four running sums fed from the top index downwards in
// groups of eight (= four samples), the ragged tail upwards, then ((s0+s1)+s2)+s3 -- the blocks are walked
// downwards and the (at most three) tail samples are visited last.
template <bool INIT>
__global__ void __launch_bounds__(kEvalThreads, 3) staged_eval_kernel(const StagedArgs A, int nxt) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double* Csm = smem + tid;  // Csm[(3*cp + {uL, uR, v}) * kEvalThreads]
  WarpTiles* wt = reinterpret_cast<WarpTiles*>(smem + kEvalTableDoubles * kEvalThreads) + wid;
  const unsigned int total = INIT ? (unsigned int)A.L.B : A.cnt[1];
  for (unsigned int base = (blockIdx.x * (kEvalThreads / 32) + wid) * 32; base < total;
       base += gridDim.x * kEvalThreads) {
    const unsigned int item = base + lane;
    bool have = item < total;
    const int f = have ? (INIT ? (int)item : A.eval_list[item]) : 0;
    FitGeom g;
    g.n = 0; g.nsamp = 0; g.b1 = g.b2 = g.b3 = 0;
    long long o0 = 0;
    if (have) staged_geom(A, f, g, o0);
    if (INIT && have) {
      const long long nn = A.L.off[f + 1] - o0;
      int ret_code = 0;
      if (nn < 19) ret_code = -1;  // lm_core.c:117
      else if (nn > A.L.n_max || (nn & 1)) ret_code = -2;
      else if (nn > 40 * kTileRows) ret_code = -2;  // same limit as the monolithic kernel's compact path
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
    const double* tp = have ? staged_t(A, f, o0) : nullptr;
    const double* pp = INIT ? A.L.p + (size_t)f * 19 : A.pdp + (size_t)f * 19;
    __syncwarp();
    wt->xp[lane] = x;
    wt->tp[lane] = tp;
    wt->n[lane] = g.n;
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
    const int nblk_max = __reduce_max_sync(0xffffffffu, nblk);
    const int full = (g.n >> 3) << 3;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const double* xt = &wt->x[lane][0];
    double* tt = &wt->t[lane][0];
    auto resid = [&](int s, double tval, double xu, double xv, double& eu, double& ev) {
      const int seg = (s >= g.b1) + (s >= g.b2) + (s >= g.b3);
      const int cb = 12 * (seg & 1), side = seg >> 1;
      double w[4];
      bernstein3(tval, w);
      double au = 0.0, av = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        au += w[k] * Csm[(cb + 3 * k + side) * kEvalThreads];
        av += w[k] * Csm[(cb + 3 * k + 2) * kEvalThreads];
      }
      eu = xu - au;
      ev = xv - av;
    };
    __syncwarp();
#pragma unroll 1
    for (int b = nblk_max - 1; b >= 0; --b) {
      tiles_load_x(wt, b, lane);
      if (!derive_t) tiles_load_t(wt, b, lane);
      cp_async_wait_all();
      __syncwarp();
      if (b < nblk) {
        if (derive_t) {  // fit_curve:451-511: t = clamp((y - y_first)/(y_last - y_first), 0, 1) per segment
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
      __syncwarp();
      if (derive_t) {  // publish the block's t, coalesced (lane = element, two fits per trip)
        const int half = lane >> 4, el = lane & 15;
        for (int j2 = 0; j2 < 16; ++j2) {
          const int fit = 2 * j2 + half;
          const int idx = 16 * b + el;
          if (idx < (wt->n[fit] >> 1)) const_cast<double*>(wt->tp[fit])[idx] = wt->t[fit][el];
        }
        __syncwarp();
      }
    }
    if (!have) continue;
    {
      const int rem = g.n - full;  // 0, 2, 4 or 6 trailing values: at most three samples, read directly
      for (int k = 0; k < rem; k += 2) {
        const int s = (full + k) >> 1;
        double eu, ev;
        resid(s, tp[s], x[2 * s], x[2 * s + 1], eu, ev);
        const int cu = (7 - (rem - k)) & 3, cv = (7 - (rem - k - 1)) & 3;
        const double du = eu * eu, dv = ev * ev;
        if (cu == 0) s0 += du; else if (cu == 1) s1 += du; else if (cu == 2) s2 += du; else s3 += du;
        if (cv == 0) s0 += dv; else if (cv == 1) s1 += dv; else if (cv == 2) s2 += dv; else s3 += dv;
      }
    }
    const double e_new = ((s0 + s1) + s2) + s3;

    StagedScal* sp = A.scal + f;
    if (INIT) {
      StagedScal s;
      s.e0 = e_new; s.e_sq = e_new; s.g_inf = 0.0; s.dp_sq = DBL_MAX; s.mu = 0.0; s.p_sq = 0.0; s.maxdiag = 0.0;
      s.dL = 0.0; s.nu = 2; s.k = 0; s.nfev = 1; s.njev = 0; s.nlss = 0; s.pad0 = s.pad1 = s.pad2 = 0;
      *sp = s;
      if (!isfinite(e_new)) staged_finish(A, f, s, 7);
      else if (!(0 < A.L.itmax)) staged_finish(A, f, s, 0);
      else if (e_new <= A.L.eps3) staged_finish(A, f, s, 6);
      else staged_append(A.jac_list, A.cnt + 0, f);
      continue;
    }
    StagedScal s = *sp;
    s.nfev += 1;
    if (!isfinite(e_new)) {  // lm_core.c:352-355, inside the inner loop: the for's ++k still runs
      s.k += 1;
      *sp = s;
      staged_finish(A, f, s, 7);
      continue;
    }
    const double dF = s.e_sq - e_new;
    if (s.dL > 0.0 && dF > 0.0) {
      s.mu = nielsen(s.mu, dF, s.dL);
      s.nu = 2;
      double* pout = A.L.p + (size_t)f * 19;
#pragma unroll
      for (int q = 0; q < 19; ++q) pout[q] = pp[q];
      s.e_sq = e_new;
      s.k += 1;
      *sp = s;
      if (s.k >= A.L.itmax) staged_finish(A, f, s, 0);
      else if (s.e_sq <= A.L.eps3) staged_finish(A, f, s, 6);
      else staged_append(A.jac_list, A.cnt + 0, f);
      continue;
    }
    s.mu *= s.nu;
    const int nu2 = (int)((unsigned)s.nu << 1);
    if (nu2 <= s.nu) {
      s.k += 1;
      *sp = s;
      staged_finish(A, f, s, 5);
      continue;
    }
    s.nu = nu2;
    *sp = s;
    staged_append(A.solve_list[nxt], A.cnt + 2 + nxt, f);
  }
}

// ---- solve: warp per fit ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSolveWarps * 32) staged_solve_kernel(const StagedArgs A, int cur) {
  constexpr int M = 19;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* base = smem + (size_t)wid * kSolveWarpDoubles;
  WarpMem<M> w;
  w.x = nullptr; w.t = nullptr; w.e = nullptr; w.hx = nullptr; w.wrk = nullptr; w.wrk2 = nullptr; w.dbuf = nullptr;
  w.tile = base;
  w.jtj = base + M * M;
  double* v = base + 2 * M * M;
  w.p = v; w.dp = v + 32; w.pdp = v + 64; w.jte = v + 96; w.diag = v + 128; w.pw = v + 160; w.xs = v + 192;
  w.scale = v + 224;
  w.scratch = nullptr;
  const unsigned int total = A.cnt[2 + cur];
  const unsigned int nwarps = gridDim.x * kSolveWarps;
  for (unsigned int i = blockIdx.x * kSolveWarps + wid; i < total; i += nwarps) {
    const int f = A.solve_list[cur][i];
    const double* lower = A.jtj + (size_t)f * kLower19;
    for (int idx = lane; idx < M * M; idx += 32) {
      const int r = idx / M, c = idx - r * M;
      w.jtj[idx] = (c <= r) ? lower[tri(r, c)] : lower[tri(c, r)];
    }
    if (lane < M) {
      w.jte[lane] = A.jte[(size_t)f * M + lane];
      w.p[lane] = A.L.p[(size_t)f * M + lane];
    }
    __syncwarp();
    StagedScal* sp = A.scal + f;
    double mu = sp->mu;
    const double p_sq = sp->p_sq;
    int nu = sp->nu, nlss = sp->nlss;
    double dp_sq = sp->dp_sq;
    int stop = 0;
    bool to_eval = false;
    for (;;) {
      const int solved = lu_solve<M>(w, mu, lane);
      ++nlss;
      if (solved) {
        const int chk = step_checks<M>(w, p_sq, A.L.eps2, A.L.eps2_sq, dp_sq, lane);
        if (chk) { stop = chk; break; }
        to_eval = true;
        break;
      }
      mu *= nu;
      const int nu2 = (int)((unsigned)nu << 1);
      if (nu2 <= nu) { stop = 5; break; }
      nu = nu2;
    }
    if (to_eval) {
      const double dL = predicted_reduction<M>(w, mu);
      if (lane < M) A.pdp[(size_t)f * M + lane] = w.pdp[lane];
      if (lane == 0) {
        sp->mu = mu; sp->nu = nu; sp->nlss = nlss; sp->dp_sq = dp_sq; sp->dL = dL;
        staged_append(A.eval_list, A.cnt + 1, f);
      }
    } else if (lane == 0) {
      StagedScal s = *sp;
      s.mu = mu; s.nu = nu; s.nlss = nlss; s.dp_sq = dp_sq;
      s.k += 1;  // the stop happened inside the inner loop
      *sp = s;
      staged_finish(A, f, s, stop);
    }
    __syncwarp();
  }
}

// ---- jac: thread per fit ---------------------------------------------------------------------------------
// Shared memory per thread (strided by the CTA size: a warp's accesses are conflict-free):
//   D[40]  derivative table of the current contour segment: [u-row | v-row][control point k][xe, ye, theta, phi, z]
//   C[8]   projected control points of the segment: u of cp 0..3, v of cp 0..3
// plus the warp's staging tiles.  Registers: the 66 partial sums of the current 32-row block and the 11 J^T e
// chains of the current curve.  The running J^T J entries are touched once per block and live in the global work
// area (A.ws, [slot][thread]: coalesced, L2-resident).
struct JacRegs {
  double part[66];
  double jte_cp[8];
  double jte_q[3];
};

__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// The segment tables are re-read from shared memory for every row.  A plain load would be hoisted out of the row
// loop (the tables are loop-invariant) into 96 registers, which the 66 partial sums need; `volatile` pins it.
__device__ __forceinline__ double lds_pinned(unsigned int saddr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
  return v;
}

// one Jacobian row (R = 0: the u row of the sample, 1: the v row): residual, compact row, 66 + 11 multiply-adds
template <int R>
__device__ __forceinline__ void jac_row_mac(unsigned int Dsa, unsigned int Csa, const double (&w)[4], double xval,
                                            JacRegs& G) {
  constexpr unsigned int P = kJacThreads * 8;  // byte pitch between table entries
  double au = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) au += w[k] * lds_pinned(Csa + (4 * R + k) * P);
  const double e = xval - au;
  double v[11];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = w[k] * lds_pinned(Dsa + (20 * R + 5 * k + 0) * P);
    v[2 * k + 1] = w[k] * lds_pinned(Dsa + (20 * R + 5 * k + 1) * P);
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) a += w[k] * lds_pinned(Dsa + (20 * R + 5 * k + 2 + q) * P);
    v[8 + q] = a;
  }
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) G.part[a * (a + 1) / 2 + b] += v[a] * v[b];
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int b = 0; b < 8; ++b) G.part[36 + 8 * q + b] += v[8 + q] * v[b];
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int r = 0; r <= q; ++r) G.part[60 + q * (q + 1) / 2 + r] += v[8 + q] * v[8 + r];
#pragma unroll
  for (int a = 0; a < 8; ++a) G.jte_cp[a] += v[a] * e;
#pragma unroll
  for (int q = 0; q < 3; ++q) G.jte_q[q] += v[8 + q] * e;
}

// derivative table and projected control points of one contour segment (curve, side): the four control points of
// prep_jac / prep_func of the monolithic kernel, same expressions
__device__ __noinline__ void jac_segment_tables(const double* pp, const Trig& tg, int curve, int side, double* Dsm,
                                                double* Csm) {
  const double z = pp[18];
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const double xe = pp[8 * curve + 2 * k], ye = pp[8 * curve + 2 * k + 1];
    double xb[3];
    body_point(tg, xe, ye, z, xb);
    double dx[5][3];
    dx[0][0] = tg.ct;  dx[0][1] = tg.stsp;  dx[0][2] = tg.stcp;
    dx[1][0] = 0.0;    dx[1][1] = tg.cp;    dx[1][2] = -tg.sp;
    dx[2][0] = tg.ct * z - tg.st * xe;
    dx[2][1] = tg.ctsp * xe + tg.stsp * z;
    dx[2][2] = tg.ctcp * xe + tg.stcp * z;
    dx[3][0] = 0.0;    dx[3][1] = xb[2];    dx[3][2] = -xb[1];
    dx[4][0] = tg.st;  dx[4][1] = -tg.ctsp; dx[4][2] = -tg.ctcp;
    const double ys = side ? (xb[1] - 0.5 * CPS_BASELINE) : (xb[1] + 0.5 * CPS_BASELINE);
    const double ix2 = 1.0 / (xb[0] * xb[0]);
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      Dsm[(5 * k + q) * kJacThreads] = (CPS_FX * ((dx[q][1] * xb[0]) - (ys * dx[q][0]))) * ix2;
      Dsm[(20 + 5 * k + q) * kJacThreads] = (CPS_FY * ((dx[q][2] * xb[0]) - (xb[2] * dx[q][0]))) * ix2;
    }
    Csm[k * kJacThreads] = CPS_FX * ys / xb[0] + CPS_CX;
    Csm[(4 + k) * kJacThreads] = CPS_FY * (xb[2] / xb[0]) + CPS_CY;
  }
}

// the reference's unblocked loop for small problems (n*m < 1024, lm_core.c:213-237): descending rows, every entry
// accumulated directly.  Rare (fewer than 27 samples in a fit); dense rows through the output array.
__device__ __noinline__ void jac_unblocked(const double* pp, const Trig& tg, const FitGeom& g, const double* x,
                                           const double* tp, double* Dsm, double* Csm, double* Jout, double* Jte) {
  for (int q = 0; q < kLower19; ++q) Jout[q] = 0.0;
  for (int q = 0; q < 19; ++q) Jte[q] = 0.0;
  int cur_seg = -1;
  for (int l = g.n - 1; l >= 0; --l) {
    const int s = l >> 1, uv = l & 1;
    const int seg = (s >= g.b1) + (s >= g.b2) + (s >= g.b3);
    const int curve = seg & 1, side = seg >> 1;
    if (seg != cur_seg) {
      jac_segment_tables(pp, tg, curve, side, Dsm, Csm);
      cur_seg = seg;
    }
    double w[4];
    bernstein3(tp[s], w);
    double au = 0.0;
    for (int k = 0; k < 4; ++k) au += w[k] * Csm[(4 * uv + k) * kJacThreads];
    const double e = x[l] - au;
    double row[19];
    for (int q = 0; q < 19; ++q) row[q] = 0.0;
    for (int k = 0; k < 4; ++k) {
      row[8 * curve + 2 * k] = w[k] * Dsm[(20 * uv + 5 * k + 0) * kJacThreads];
      row[8 * curve + 2 * k + 1] = w[k] * Dsm[(20 * uv + 5 * k + 1) * kJacThreads];
    }
    for (int q = 0; q < 3; ++q) {
      double a = 0.0;
      for (int k = 0; k < 4; ++k) a += w[k] * Dsm[(20 * uv + 5 * k + 2 + q) * kJacThreads];
      row[16 + q] = a;
    }
    for (int i = 18; i >= 0; --i) {
      for (int j = i; j >= 0; --j) Jout[tri(i, j)] += row[j] * row[i];
      Jte[i] += row[i] * e;
    }
  }
}

__global__ void __launch_bounds__(kJacThreads, 2) staged_jac_kernel(const StagedArgs A, int nxt) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double* Dsm = smem + tid;
  double* Csm = smem + 40 * kJacThreads + tid;
  const unsigned int Dsa = (unsigned int)__cvta_generic_to_shared(Dsm);
  const unsigned int Csa = (unsigned int)__cvta_generic_to_shared(Csm);
  WarpTiles* wt = reinterpret_cast<WarpTiles*>(smem + kJacTableDoubles * kJacThreads) + wid;
  const int gtid = blockIdx.x * kJacThreads + tid;
  const size_t wsn = (size_t)A.ws_stride;
  double* ws = A.ws + gtid;
  const unsigned int total = A.cnt[0];
  for (unsigned int base = (blockIdx.x * (kJacThreads / 32) + wid) * 32; base < total; base += gridDim.x * kJacThreads) {
    const unsigned int item = base + lane;
    const bool have = item < total;
    const int f = have ? A.jac_list[item] : 0;
    FitGeom g;
    g.n = 0; g.nsamp = 0; g.b1 = g.b2 = g.b3 = 0;
    long long o0 = 0;
    if (have) staged_geom(A, f, g, o0);
    const double* x = A.L.meas + o0;
    const double* tp = have ? staged_t(A, f, o0) : nullptr;
    const double* pp = A.L.p + (size_t)f * 19;
    double* Jout = A.jtj + (size_t)f * kLower19;
    double* Jte = A.jte + (size_t)f * 19;
    const bool blocked = have && !(g.n * 19 < kBlockSzSq);  // strict "<", lm_core.c:193
    __syncwarp();
    wt->xp[lane] = x;
    wt->tp[lane] = tp;
    wt->n[lane] = blocked ? g.n : 0;
    Trig tg;
    if (have) make_trig(pp[16], pp[17], tg);
    JacRegs G;
#pragma unroll
    for (int q = 0; q < 66; ++q) G.part[q] = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) G.jte_cp[q] = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) G.jte_q[q] = 0.0;
    if (blocked) {
#pragma unroll 1
      for (int q = 0; q < kWsPend; ++q) __stcg(ws + q * wsn, 0.0);
    }
    const int nblk = blocked ? (g.n + 31) >> 5 : 0;
    const int nblk_max = __reduce_max_sync(0xffffffffu, nblk);
    int cur_seg = -1, cur_curve = -1;
    unsigned parked = 0;  // bit c: curve c has a block partial parked in the work area
    const double* xt = &wt->x[lane][0];
    const double* tt = &wt->t[lane][0];
    __syncwarp();
#pragma unroll 1
    for (int b = 0; b < nblk_max; ++b) {
      tiles_load_x(wt, b, lane);
      tiles_load_t(wt, b, lane);
      cp_async_wait_all();
      __syncwarp();
      if (b < nblk) {
        int s = 16 * b;
        const int blk_end = min(16 * b + 16, g.nsamp);
#pragma unroll 1
        while (s < blk_end) {
          const int seg = (s >= g.b1) + (s >= g.b2) + (s >= g.b3);
          if (seg != cur_seg) {
            const int curve = seg & 1;
            if (curve != cur_curve) {
              if (cur_curve >= 0) {
#pragma unroll
                for (int a = 0; a < 8; ++a) ws[(kWsJte + 8 * cur_curve + a) * wsn] = G.jte_cp[a];
                if ((s & 15) != 0) {  // this block's partial of the old curve waits for the block to end
#pragma unroll
                  for (int q = 0; q < 60; ++q) ws[(kWsPend + 60 * cur_curve + q) * wsn] = G.part[q];
                  parked |= 1u << cur_curve;
                }
              }
#pragma unroll
              for (int a = 0; a < 8; ++a) G.jte_cp[a] = __ldcg(ws + (kWsJte + 8 * curve + a) * wsn);
              if (parked & (1u << curve)) {  // rows of this curve were already seen in this block
#pragma unroll
                for (int q = 0; q < 60; ++q) G.part[q] = __ldcg(ws + (kWsPend + 60 * curve + q) * wsn);
                parked &= ~(1u << curve);
              } else {
#pragma unroll
                for (int q = 0; q < 60; ++q) G.part[q] = 0.0;
              }
              cur_curve = curve;
            }
            jac_segment_tables(pp, tg, curve, seg >> 1, Dsm, Csm);
            cur_seg = seg;
          }
          const int e_end = min(blk_end, seg_end(g, seg));
#pragma unroll 1
          for (; s < e_end; ++s) {
            const int l = s & 15;
            double w[4];
            bernstein3(tt[l], w);
            jac_row_mac<0>(Dsa, Csa, w, xt[2 * l], G);
            jac_row_mac<1>(Dsa, Csa, w, xt[2 * l + 1], G);
          }
        }
        // end of a 32-row block: its partial sums join the running entries (misc_core.c:118-123).  The additions are
        // fire-and-forget reductions performed at the L2 (red.global.add.f64: the same IEEE round-to-nearest addition;
        // one owner per address, program order per address): no load latency is exposed and the partial sums' registers
        // are free again at once.
        {
          double* acc = ws + (size_t)(kWsAcc + 60 * cur_curve) * wsn;
#pragma unroll
          for (int q = 0; q < 60; ++q) {
            red_add_f64(acc + q * wsn, G.part[q]);
            G.part[q] = 0.0;
          }
          double* accq = ws + (size_t)(kWsAcc + 120) * wsn;
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            red_add_f64(accq + q * wsn, G.part[60 + q]);
            G.part[60 + q] = 0.0;
          }
          if (parked) {
#pragma unroll 1
            for (int c = 0; c < 2; ++c)
              if (parked & (1u << c))
                for (int q = 0; q < 60; ++q)
                  red_add_f64(ws + (kWsAcc + 60 * c + q) * wsn, __ldcg(ws + (kWsPend + 60 * c + q) * wsn));
            parked = 0;
          }
        }
      }
      __syncwarp();
    }
    if (!have) continue;
    if (blocked) {
      if (cur_curve >= 0) {
#pragma unroll
        for (int a = 0; a < 8; ++a) ws[(kWsJte + 8 * cur_curve + a) * wsn] = G.jte_cp[a];
      }
      // the lower triangle, row-major: curve blocks, the structurally zero cross block, theta/phi/z rows.  Every
      // group is read into registers before it is written so that the loads overlap (the work area is read past the
      // L1: its entries were updated by reductions at the L2).
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double v[36];
#pragma unroll
        for (int q = 0; q < 36; ++q) v[q] = __ldcg(ws + (kWsAcc + 60 * c + q) * wsn);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int bq = 0; bq <= a; ++bq) Jout[tri(8 * c + a, 8 * c + bq)] = v[a * (a + 1) / 2 + bq];
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int bq = 0; bq < 8; ++bq) Jout[tri(8 + a, bq)] = 0.0;
      {
        double v[54];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int q = 0; q < 24; ++q) v[24 * c + q] = __ldcg(ws + (kWsAcc + 60 * c + 36 + q) * wsn);
#pragma unroll
        for (int q = 0; q < 6; ++q) v[48 + q] = __ldcg(ws + (kWsAcc + 120 + q) * wsn);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
          for (int j = 0; j < 16; ++j) Jout[tri(16 + q, j)] = v[24 * (j >> 3) + 8 * q + (j & 7)];
#pragma unroll
          for (int r = 0; r <= q; ++r) Jout[tri(16 + q, 16 + r)] = v[48 + q * (q + 1) / 2 + r];
        }
      }
      {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __ldcg(ws + (kWsJte + q) * wsn);
#pragma unroll
        for (int q = 0; q < 16; ++q) Jte[q] = v[q];
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) Jte[16 + q] = G.jte_q[q];
    } else {
      jac_unblocked(pp, tg, g, x, tp, Dsm, Csm, Jout, Jte);
    }

    // gradient statistics, lm_core.c:256-287
    double p_sq = 0.0, g_inf = 0.0, big = -DBL_MAX;
#pragma unroll 1
    for (int q = 0; q < 19; ++q) {
      const double tmp = fabs(Jte[q]);
      if (g_inf < tmp) g_inf = tmp;
      p_sq += pp[q] * pp[q];
      const double d = Jout[tri(q, q)];
      if (d > big) big = d;
    }
    StagedScal* sp = A.scal + f;
    StagedScal sc = *sp;
    sc.njev += 1;
    sc.g_inf = g_inf;
    sc.p_sq = p_sq;
    sc.maxdiag = big;
    if (g_inf <= A.L.eps1) {
      sc.dp_sq = 0.0;
      *sp = sc;
      staged_finish(A, f, sc, 1);
      continue;
    }
    if (sc.k == 0) sc.mu = A.L.tau * big;
    *sp = sc;
    staged_append(A.solve_list[nxt], A.cnt + 2 + nxt, f);
  }
}

}  // namespace cps
