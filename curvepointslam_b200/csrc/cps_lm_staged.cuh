// cps_lm_staged.cuh -- dlevmar_der for large batches of Stereo19 fits as a pipeline of three phase kernels.
//
// The monolithic kernel (cps_lm_kernel.cuh) gives a warp to a fit and keeps it for the whole fit; every
// phase then runs at the occupancy of the hungriest one and the dominant phase, J^T J, pays for spreading
// one fit's 32-row blocks over lanes (tiles through shared memory, partial sums folded through shuffles).
// With >= tens of thousands of fits in a batch there is a better source of parallelism: the fits
// themselves.  Here every fit is a small state machine kept in HBM and a round of three kernels advances
// all active fits by one linear solve, every kernel with a thread per fit:
//
//   solve  (J^T J + mu I) dp = J^T e by the reference's LU on a work matrix in shared memory, step checks,
//          predicted reduction                                                          lm_core.c:292-345
//   eval   f(p+dp), ||x - f||^2 in the reference's four chains, accept / reject, Nielsen damping, the
//          loop-top checks of the next iteration                                         lm_core.c:347-388,198
//   jac    analytic Jacobian, J^T J as per-32-row-block partial sums, J^T e, gradient statistics, initial
//          mu (two threads per fit when few fits are left)                               lm_core.c:205-287
//
// A thread that owns a whole fit walks the rows in the reference's own order, so the summation orders
// of the parity contract are simply the loop orders: no shuffles, no tiles shared
// between lanes, no idle lanes.  Compaction lists carry the fits from kernel to kernel; a rejected step goes
// straight back to `solve` with the old J^T J (the inner loop of dlevmar_der).  The arithmetic -- every
// operation and its order -- is the same as in the monolithic kernel, so both paths are bit-identical to the
// oracle and to each other (tests/test_lm_staged_gpu.py).
#pragma once
#include "cps_lm_kernel.cuh"

namespace cps {

using StagedScal = LmResume;  // per-fit scalars of the LM iteration, 96 bytes (cps_lm_kernel.cuh)

constexpr int kLower19 = 19 * 20 / 2;  // J^T J is kept as its lower triangle, row-major: (i,j) -> i(i+1)/2 + j
constexpr int kJacThreads = 128;       // jac kernel: threads per CTA (2 CTAs per SM: 256 fits resident per SM)
constexpr int kJacTableDoubles = 40 + 8;  // per thread: derivative table + projected control points of a segment
constexpr int kEvalThreads = 128;
constexpr int kEvalTableDoubles = 24;  // per thread: the 8 projected control points
// per-thread slots of the jac kernel's global (L2-resident, [slot][thread] = coalesced) work area
constexpr int kWsAcc = 0;     // 60 + 60 running J^T J entries of curve 1 / curve 2, then 6 theta/phi/z x theta/phi/z
constexpr int kWsJte = 126;   // 16 J^T e chains of the control-point columns
constexpr int kWsPend = 142;  // 2 x 60 + 6 block partial sums of a block that straddles contour segments
constexpr int kWsPendQ = 262;
constexpr int kWsSlots = 268;

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
  unsigned int* cnt;   // [0] jac, [1] Jacobian evaluations of the call so far, [2], [3] solve lists,
                       // [4] entries of fb_list this round, [5] of the call so far
  int* fb_list;        // positions in the solve list whose system the block-arrow solve handed to the dense one
  int force_handover;  // test knob: > 0 hands every force_handover-th list position over regardless
  int init_first, init_count;  // range of the fit index an INIT launch covers
  const int* init_order;       // nullptr, or [init_count] the fits of that range in the order the INIT launch takes them
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

// Appends the fits of the lanes with `pred` to a list, one atomic per warp, lanes in order: fits that sit next to each
// other in one list sit next to each other in the next (their samples are neighbours in memory).  Call convergent.
__device__ __forceinline__ void staged_append(int* list, unsigned int* counter, int f, bool pred) {
  const unsigned int mask = __ballot_sync(0xffffffffu, pred);
  if (mask == 0u) return;
  const int lane = threadIdx.x & 31;
  unsigned int base = 0;
  if (lane == (__ffs(mask) - 1)) base = atomicAdd(counter, (unsigned int)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
  if (pred) list[base + __popc(mask & ((1u << lane) - 1u))] = f;
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


// ---- staging of a fit's samples ---------------------------------------------------------------------------
// A thread that owns a fit streams that fit's samples, so a warp reads 32 different streams; read element by
// element, every 8-byte load would touch its own 32-byte sector and expose a global-memory latency per sample.
// Instead every lane copies one 32-row block of its fit (32 doubles of x, 16 of t) into its own row of a
// shared-memory tile with 16-byte cp.async (global -> shared without passing through registers, all copies of the
// block in flight at once), prefetches the next block into the L2, and then reads the row at shared-memory
// latency.  cp.async needs source and destination equally aligned: a stream that starts on an odd double is stored
// one element into the row (`off`).
struct WarpTiles {
  double x[32][34];
  double t[32][18];
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// row[off + i] = src[i], i < cnt; off = parity of src's double index
__device__ __forceinline__ void row_fill(double* row, int off, const double* src, int cnt) {
  int i = 0;
  if (off && cnt > 0) {
    cp_async8(row + 1, src);
    i = 1;
  }
#pragma unroll 4
  for (; i + 2 <= cnt; i += 2) cp_async16(row + off + i, src + i);
  if (i < cnt) cp_async8(row + off + i, src + i);
}
__device__ __forceinline__ int ptr_parity(const double* p) { return (int)((reinterpret_cast<size_t>(p) >> 3) & 1); }

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
  const unsigned int total = INIT ? (unsigned int)A.init_count : A.cnt[2 + (nxt ^ 1)];  // eval list = this round's solve list
  for (unsigned int base = (blockIdx.x * (kEvalThreads / 32) + wid) * 32; base < total;
       base += gridDim.x * kEvalThreads) {
    const unsigned int item = base + lane;
    bool have = item < total;
    int f = have ? (INIT ? (A.init_order ? A.init_order[item] : A.init_first + (int)item) : A.eval_list[item]) : 0;
    if (f < 0) { have = false; f = 0; }  // the fit stopped in the solve
    FitGeom g;
    g.n = 0; g.nsamp = 0; g.b1 = g.b2 = g.b3 = 0;
    long long o0 = 0;
    if (have) staged_geom(A, f, g, o0);
    if (INIT && have) {
      const long long nn = A.L.off[f + 1] - o0;
      int ret_code = 0;
      if (nn < 19) ret_code = -1;  // lm_core.c:117
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
    const double* tp = have ? staged_t(A, f, o0) : nullptr;
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
      if (derive_t) {  // publish the block's t (the work area's rows start on even doubles: toff = 0)
        double* tw = const_cast<double*>(tp) + 16 * b;
        const int cnt = min(16, g.nsamp - 16 * b);
        int i = 0;
        for (; i + 2 <= cnt; i += 2) *reinterpret_cast<double2*>(tw + i) = make_double2(tt[i], tt[i + 1]);
        if (i < cnt) tw[i] = tt[i];
      }
    }
    bool to_jac = false, to_solve = false;
    if (have) {
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
        else to_jac = true;
      } else {
        StagedScal s = *sp;
        s.nfev += 1;
        const double dF = s.e_sq - e_new;
        if (!isfinite(e_new)) {  // lm_core.c:352-355, inside the inner loop: the for's ++k still runs
          s.k += 1;
          *sp = s;
          staged_finish(A, f, s, 7);
        } else if (s.dL > 0.0 && dF > 0.0) {
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
          else to_jac = true;
        } else {
          s.mu *= s.nu;
          const int nu2 = (int)((unsigned)s.nu << 1);
          if (nu2 <= s.nu) {
            s.k += 1;
            *sp = s;
            staged_finish(A, f, s, 5);
          } else {
            s.nu = nu2;
            *sp = s;
            to_solve = true;
          }
        }
      }
    }
    __syncwarp();
    staged_append(A.jac_list, A.cnt + 0, f, to_jac);
    staged_append(A.solve_list[nxt], A.cnt + 2 + nxt, f, to_solve);
  }
}

// ---- solve: thread per fit --------------------------------------------------------------------------------
// (J^T J + mu I) dp = J^T e by the reference's LU (dAx_eq_b_LU_noLapack, Axb_core.c:1044-1132), step checks
// (lm_core.c:311-331) and the predicted reduction dL (:357-360).  A warp per fit spends ten instructions of
// bookkeeping (pivot reductions, syncs, idle lanes: 19 of 32) on every useful multiply-add of a 19x19 elimination;
// a thread per fit spends two loads and a store.  The work matrix, the row scales and the right-hand side of a
// thread live in shared memory, [element][thread] so that any (row, column) a thread picks -- pivoting makes the
// row dynamic -- is conflict-free: 399 doubles per thread, 64 threads per SM.
//
// Same operation sequence as lu_solve of the monolithic kernel (right-looking form of the reference's Crout
// sweep: element (i,k) receives a_ik - l_i0*u_0k - l_i1*u_1k - ... in ascending order, so every stored value is
// bit-identical): scaled partial pivoting where the LAST maximum wins (">=", :1088) and a search that finds no
// valid merit keeps the previous pivot row, zero pivot -> DBL_EPSILON (:1102), forward substitution that skips
// the leading zeros of the permuted right-hand side (:1112-1121), back substitution in ascending column order.
constexpr int kSolveThreads = 64;
constexpr int kSolveDoubles = 19 * 19 + 19 + 19;

// Trailing update a_ik -= l_i * u_k of the elimination, restricted to what can change a value: a row whose
// multiplier l_i is exactly zero receives "- (0 * u_k)", which leaves every entry as it is, and so does a column
// whose pivot-row entry u_k is zero.  J^T J + mu I of the Stereo19 model has two 8 x 8 blocks of structural zeros
// (the control points of one curve against those of the other); while the first curve's columns are eliminated the
// second curve's rows have zero multipliers and -- unless a theta/phi/z row was chosen as pivot -- the pivot row is
// zero in the second curve's columns: 60 % of the multiply-adds of a dense sweep touch nothing.  Skipping them is
// exact for finite data (sums here are never -0.0, so x - (+-0) == x bit for bit); it is decided at run time from
// the values, not assumed from the structure.
template <int M, int T, int W, int NR>
__device__ __forceinline__ void lu_update_rows(double* a, int j, int k0, const int (&rows)[NR], const double (&u)[W]) {
  double l[NR], x[NR][W];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    l[r] = a[(rows[r] * M + j) * T];
#pragma unroll
    for (int c = 0; c < W; ++c) x[r][c] = a[(rows[r] * M + k0 + c) * T];
  }
  double p[NR][W];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < W; ++c) p[r][c] = l[r] * u[c];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < W; ++c) x[r][c] -= p[r][c];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < W; ++c) a[(rows[r] * M + k0 + c) * T] = x[r][c];
}

// W columns k0..k0+W-1, the rows named by `rowmask`; the pivot-row entries stay in registers; four rows are in
// flight per trip (all loads first, then the arithmetic, then the stores) so that the shared-memory latencies overlap
template <int M, int T, int W>
__device__ __forceinline__ void lu_update_cols(double* a, int j, int k0, unsigned int rowmask) {
  double u[W];
#pragma unroll
  for (int c = 0; c < W; ++c) u[c] = a[(j * M + k0 + c) * T];
  unsigned int m = rowmask;
#pragma unroll 1
  while (m) {
    int rows[4];
    int n = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      rows[r] = m ? (__ffs(m) - 1) : 0;
      if (m) { m &= m - 1; ++n; }
    }
    if (n == 4) {
      lu_update_rows<M, T, W, 4>(a, j, k0, rows, u);
    } else {
      if (n >= 2) {
        const int r2[2] = {rows[0], rows[1]};
        lu_update_rows<M, T, W, 2>(a, j, k0, r2, u);
      }
      if (n & 1) {
        const int r1[1] = {n == 1 ? rows[0] : rows[2]};
        lu_update_rows<M, T, W, 1>(a, j, k0, r1, u);
      }
    }
  }
}

// the columns lo..hi-1 in chunks of at most six
template <int M, int T>
__device__ __forceinline__ void lu_update_window(double* a, int j, int lo, int hi, unsigned int rowmask) {
  int k0 = lo;
  for (; k0 + 6 <= hi; k0 += 6) lu_update_cols<M, T, 6>(a, j, k0, rowmask);
  switch (hi - k0) {
    case 5: lu_update_cols<M, T, 5>(a, j, k0, rowmask); break;
    case 4: lu_update_cols<M, T, 4>(a, j, k0, rowmask); break;
    case 3: lu_update_cols<M, T, 3>(a, j, k0, rowmask); break;
    case 2: lu_update_cols<M, T, 2>(a, j, k0, rowmask); break;
    case 1: lu_update_cols<M, T, 1>(a, j, k0, rowmask); break;
    default: break;
  }
}

// With one warp per scheduler nothing hides a latency but the thread's own independent work, so every phase is
// written as "issue all loads of a batch, then the arithmetic, then the stores" over statically unrolled batches.
// fills the work matrix of the m=19 solve from the lower triangle in global memory
__device__ __forceinline__ void lu_fill_global19(double* a, double* bx, const double* lower, const double* jte, double mu) {
  constexpr int M = 19, T = kSolveThreads;
#define A_(i, k) a[((i) * M + (k)) * T]
  {  // augmented copy of the symmetric matrix and the right-hand side, global -> shared by cp.async (every element
     // of the lower triangle lands in both halves; 380 copies in flight at once, no register staging); then mu
     // joins the diagonal
    int idx = 0;
#pragma unroll 1
    for (int i = 0; i < M; ++i) {
      for (int k = 0; k < i; ++k, ++idx) {
        cp_async8(&A_(i, k), lower + idx);
        cp_async8(&A_(k, i), lower + idx);
      }
      cp_async8(&A_(i, i), lower + idx);
      ++idx;
    }
#pragma unroll
    for (int i = 0; i < M; ++i) cp_async8(bx + i * T, jte + i);
    cp_async_wait_all();
    double d[M];
#pragma unroll
    for (int i = 0; i < M; ++i) d[i] = A_(i, i);
#pragma unroll
    for (int i = 0; i < M; ++i) A_(i, i) = d[i] + mu;
  }
#undef A_
}

// factorises the M x M matrix in a[(i*M+k)*T] (shared memory, [element][thread]) and solves for the right-hand side
// in bx[i*T]; returns 0 if a row is entirely zero
template <int M, int T>
__device__ __forceinline__ int lu_factor_solve(double* a, double* sc, double* bx, double (&xout)[M]) {
#define A_(i, k) a[((i) * M + (k)) * T]
  {  // row scales: 1 / max_j |a_ij|, all running maxima advance together
    double big[M];
#pragma unroll
    for (int i = 0; i < M; ++i) big[i] = 0.0;
#pragma unroll 1
    for (int j = 0; j < M; ++j) {
      double v[M];
#pragma unroll
      for (int i = 0; i < M; ++i) v[i] = fabs(A_(i, j));
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (v[i] > big[i]) big[i] = v[i];
    }
    int singular = 0;
#pragma unroll
    for (int i = 0; i < M; ++i)
      if (big[i] == 0.0) singular = 1;
    if (singular) return 0;
#pragma unroll
    for (int i = 0; i < M; ++i) sc[i * T] = 1.0 / big[i];
  }

  int piv = -1;  // the reference's maxi persists across columns
#pragma unroll 1
  for (int j = 0; j < M; ++j) {
    {
      double merit[M];
#pragma unroll
      for (int i = 0; i < M; ++i) merit[i] = (i >= j) ? sc[i * T] * fabs(A_(i, j)) : 0.0;
      double big = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (i >= j && merit[i] >= big) { big = merit[i]; piv = i; }
    }
    if (piv != j && piv >= 0) {
      double rp[M], rj[M];
#pragma unroll
      for (int k = 0; k < M; ++k) { rp[k] = A_(piv, k); rj[k] = A_(j, k); }
#pragma unroll
      for (int k = 0; k < M; ++k) { A_(piv, k) = rj[k]; A_(j, k) = rp[k]; }
      sc[piv * T] = sc[j * T];
      const double tb = bx[piv * T], tj = bx[j * T];
      bx[piv * T] = tj;
      bx[j * T] = tb;
    }
    if (A_(j, j) == 0.0) A_(j, j) = DBL_EPSILON;
    if (j != M - 1) {
      const double inv = 1.0 / A_(j, j);
      unsigned int rowmask = 0;  // rows below j with a non-zero multiplier
      {
        double l[M];
#pragma unroll
        for (int i = 1; i < M; ++i) l[i] = (i > j) ? A_(i, j) : 0.0;
#pragma unroll
        for (int i = 1; i < M; ++i)
          if (i > j) {
            A_(i, j) = l[i] * inv;
            if (l[i] != 0.0) rowmask |= 1u << i;
          }
      }
      // columns right of j; if the pivot row is zero across the second curve's columns 8..15, those are left out
      bool gap = false;
      if (M == 19 && j < 8) {
        double u8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) u8[k] = A_(j, (8 + k) % M);
        gap = true;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (u8[k] != 0.0) gap = false;
      }
      if (gap) {
        lu_update_window<M, T>(a, j, j + 1, 8, rowmask);
        lu_update_window<M, T>(a, j, 16, M, rowmask);
      } else {
        lu_update_window<M, T>(a, j, j + 1, M, rowmask);
      }
    }
  }
  // the permutation is final: the right-hand side moves to registers and both substitutions are fully unrolled
  double b[M];
#pragma unroll
  for (int i = 0; i < M; ++i) b[i] = bx[i * T];
  // forward substitution, column oriented (same ascending-j order per row), with the leading-zero skip
  bool started = false;
#pragma unroll
  for (int j = 0; j < M - 1; ++j) {
    const double xj = b[j];
    if (!started && xj != 0.0) started = true;
    if (started) {
      double l[M];
#pragma unroll
      for (int i = j + 1; i < M; ++i) l[i] = A_(i, j);
#pragma unroll
      for (int i = j + 1; i < M; ++i) b[i] -= l[i] * xj;
    }
  }
#pragma unroll
  for (int i = M - 1; i >= 0; --i) {
    double u[M];
#pragma unroll
    for (int jj = i; jj < M; ++jj) u[jj] = A_(i, jj);
    double sum = b[i];
#pragma unroll
    for (int jj = i + 1; jj < M; ++jj) sum -= u[jj] * b[jj];
    b[i] = sum / u[i];
  }
#pragma unroll
  for (int i = 0; i < M; ++i) xout[i] = b[i];
#undef A_
  return 1;
}

// ---- the same solve on the block-arrow form of the Stereo19 matrix --------------------------------------------
// J^T J + mu I of the derived Jacobian has two 8 x 8 blocks of structural zeros (the control points of one curve
// against those of the other, written as +0.0 by the jac kernel): in the dense work matrix 128 of 361 entries are
// zeros that no elimination step touches while the pivot of a control-point column comes from that curve's own eight
// rows.  Only what can become non-zero is stored:
//   rows 0..15   11 entries each: the eight columns of the row's own curve, then columns 16..18 (theta, phi, z)
//   rows 16..18  all 19 columns
// 233 doubles instead of 361 (+ 19 scales + 19 right-hand side = 271 per thread): 96 threads per SM instead of 64,
// and the sweep never visits the zero blocks.  Every arithmetic operation that is performed is the dense
// algorithm's, in its order; the operations left out are subtractions of a product with a structural zero, which
// change nothing as long as the other factor is finite.  The two assumptions are checked, not trusted: (i) the scaled
// pivot search of a control-point column must pick one of that curve's rows (measured: always, on 1 800 matrices of
// scripts/pivot_stats.py); (ii) every stored factor and every solution entry must be finite.  A thread that finds
// either violated puts its list position on the fallback list and the dense kernel solves that system afterwards.
constexpr int kArrowThreads = 96;
constexpr int kArrowMat = 16 * 11 + 3 * 19;           // 233
constexpr int kArrowDoubles = kArrowMat + 19 + 19;    // + scales + right-hand side
constexpr int kArrowPitch = kArrowThreads + 1;        // doubles between two entries of a system ([entry][thread], padded:
                                                      // the warp-cooperative fill writes one system's entries from 32 lanes)
constexpr int kArrowSlots = kArrowDoubles + 1;        // + a flag: the jac kernel's structural zeros were not zeros

// first entry of row i's own-block columns (block b = i / 8 for i < 16; for the bottom rows the columns 8 b ..) and
// of its columns 16..18
__device__ __forceinline__ int arrow_blk(int i, int b) { return i < 16 ? i * 11 : 176 + (i - 16) * 19 + 8 * b; }
__device__ __forceinline__ int arrow_tail(int i) { return i < 16 ? i * 11 + 8 : 176 + (i - 16) * 19 + 16; }

// a_ik -= l_i u_k for the rows in `rowmask` (bit = row index 0..18), W consecutive stored entries starting at local
// column c0 of the window (`tail`: columns 16..18, otherwise the own-block columns of block b); four rows in flight
template <int T, int W, int NR>
__device__ __forceinline__ void arrow_update_rows(double* a, const int (&base)[NR], const int (&lpos)[NR], const double (&u)[W]) {
  double l[NR], x[NR][W];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    l[r] = a[lpos[r] * T];
#pragma unroll
    for (int c = 0; c < W; ++c) x[r][c] = a[(base[r] + c) * T];
  }
  double p[NR][W];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < W; ++c) p[r][c] = l[r] * u[c];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < W; ++c) x[r][c] -= p[r][c];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < W; ++c) a[(base[r] + c) * T] = x[r][c];
}

template <int T, int W>
__device__ __forceinline__ void arrow_update_cols(double* a, int j, int b, bool tail, int c0, unsigned int rowmask) {
  // pivot row j (a row of block b, or a bottom row when b == 2)
  const int jb = (tail ? arrow_tail(j) : arrow_blk(j, b)) + c0;
  double u[W];
#pragma unroll
  for (int c = 0; c < W; ++c) u[c] = a[(jb + c) * T];
  unsigned int m = rowmask;
#pragma unroll 1
  while (m) {
    int base[4], lpos[4];
    int n = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = m ? (__ffs(m) - 1) : 0;
      if (m) { m &= m - 1; ++n; }
      base[r] = (tail ? arrow_tail(i) : arrow_blk(i, b)) + c0;
      lpos[r] = (b == 2 ? arrow_tail(i) + (j - 16) : arrow_blk(i, b) + (j - 8 * b));  // the multiplier l_ij
    }
    if (n == 4) {
      arrow_update_rows<T, W, 4>(a, base, lpos, u);
    } else {
      if (n >= 2) {
        const int b2[2] = {base[0], base[1]}, l2[2] = {lpos[0], lpos[1]};
        arrow_update_rows<T, W, 2>(a, b2, l2, u);
      }
      if (n & 1) {
        const int b1[1] = {n == 1 ? base[0] : base[2]}, l1[1] = {n == 1 ? lpos[0] : lpos[2]};
        arrow_update_rows<T, W, 1>(a, b1, l1, u);
      }
    }
  }
}

// local columns lo..hi-1 of a window in chunks of at most six
template <int T>
__device__ __forceinline__ void arrow_update_window(double* a, int j, int b, bool tail, int lo, int hi, unsigned int rowmask) {
  int c0 = lo;
  for (; c0 + 6 <= hi; c0 += 6) arrow_update_cols<T, 6>(a, j, b, tail, c0, rowmask);
  switch (hi - c0) {
    case 5: arrow_update_cols<T, 5>(a, j, b, tail, c0, rowmask); break;
    case 4: arrow_update_cols<T, 4>(a, j, b, tail, c0, rowmask); break;
    case 3: arrow_update_cols<T, 3>(a, j, b, tail, c0, rowmask); break;
    case 2: arrow_update_cols<T, 2>(a, j, b, tail, c0, rowmask); break;
    case 1: arrow_update_cols<T, 1>(a, j, b, tail, c0, rowmask); break;
    default: break;
  }
}

// the work matrix from the lower triangle in global memory (cp.async, everything in flight at once), mu on the diagonal
template <int T>
__device__ __forceinline__ void arrow_fill19(double* a, double* bx, const double* lower, const double* jte, double mu) {
#pragma unroll 1
  for (int i = 0; i < 16; ++i) {
    const int b = i >> 3, r0 = 8 * b;
    for (int k = 0; k < 8; ++k) {
      const int col = r0 + k;
      cp_async8(a + (i * 11 + k) * T, lower + (col <= i ? tri(i, col) : tri(col, i)));
    }
    for (int q = 0; q < 3; ++q) cp_async8(a + (i * 11 + 8 + q) * T, lower + tri(16 + q, i));
  }
#pragma unroll 1
  for (int q = 0; q < 3; ++q)
    for (int k = 0; k < 19; ++k) cp_async8(a + (176 + q * 19 + k) * T, lower + (k <= 16 + q ? tri(16 + q, k) : tri(k, 16 + q)));
#pragma unroll
  for (int i = 0; i < 19; ++i) cp_async8(bx + i * T, jte + i);
  cp_async_wait_all();
  double d[19];
#pragma unroll
  for (int i = 0; i < 19; ++i) d[i] = a[(i < 16 ? i * 11 + (i & 7) : 176 + (i - 16) * 19 + i) * T];
#pragma unroll
  for (int i = 0; i < 19; ++i) a[(i < 16 ? i * 11 + (i & 7) : 176 + (i - 16) * 19 + i) * T] = d[i] + mu;
}

// Elimination of column j = 8 b + JL (a control-point column of curve b).  Rows that take part: JL..7 of the curve's
// block and the three bottom rows; columns: JL+1..7 of the block and 16..18.  Returns false when the reference's pivot
// search would leave the curve's rows (-> dense solve).
template <int T, int JL>
__device__ __forceinline__ bool arrow_block_step(double* a, double* sc, double* bx, int b) {
  const int r0 = 8 * b, j = r0 + JL;
  double* blk = a + (r0 * 11) * T;        // block row r, stored column c (0..7 own, 8..10 = columns 16..18): blk[(r*11 + c)*T]
  double* tlb = a + (176 + r0) * T;        // bottom row q, own-block column c: tlb[(q*19 + c)*T]
  double* tlt = a + (176 + 16) * T;        // bottom row q, column 16 + c:      tlt[(q*19 + c)*T]
  int piv = -1;
  {
    double merit[8], mt[3];
#pragma unroll
    for (int r = JL; r < 8; ++r) merit[r] = sc[(r0 + r) * T] * fabs(blk[(r * 11 + JL) * T]);
#pragma unroll
    for (int q = 0; q < 3; ++q) mt[q] = sc[(16 + q) * T] * fabs(tlb[(q * 19 + JL) * T]);
    double big = 0.0;
#pragma unroll
    for (int r = JL; r < 8; ++r)
      if (merit[r] >= big) { big = merit[r]; piv = r; }
    // the dense search goes on over the other curve's rows (merit exactly 0: they win only against big == 0) and the
    // bottom rows
    if (piv < 0 || !(big > 0.0)) return false;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (mt[q] >= big) return false;
  }
  if (piv != JL) {
    double rp[11], rj[11];
#pragma unroll
    for (int c = 0; c < 11; ++c) { rp[c] = blk[(piv * 11 + c) * T]; rj[c] = blk[(JL * 11 + c) * T]; }
#pragma unroll
    for (int c = 0; c < 11; ++c) { blk[(piv * 11 + c) * T] = rj[c]; blk[(JL * 11 + c) * T] = rp[c]; }
    sc[(r0 + piv) * T] = sc[j * T];
    const double tb = bx[(r0 + piv) * T], tj = bx[j * T];
    bx[(r0 + piv) * T] = tj;
    bx[j * T] = tb;
  }
  // (the pivot cannot be zero: its merit is positive)
  const double inv = 1.0 / blk[(JL * 11 + JL) * T];
  // multipliers, stored where the eliminated entries were
  double lm[8], lt[3];
#pragma unroll
  for (int r = JL + 1; r < 8; ++r) lm[r] = blk[(r * 11 + JL) * T];
#pragma unroll
  for (int q = 0; q < 3; ++q) lt[q] = tlb[(q * 19 + JL) * T];
#pragma unroll
  for (int r = JL + 1; r < 8; ++r) { lm[r] *= inv; blk[(r * 11 + JL) * T] = lm[r]; }
#pragma unroll
  for (int q = 0; q < 3; ++q) { lt[q] *= inv; tlb[(q * 19 + JL) * T] = lt[q]; }
  // pivot row right of the diagonal: own columns JL+1..7, then 16..18
  double u[11];
#pragma unroll
  for (int c = JL + 1; c < 11; ++c) u[c] = blk[(JL * 11 + c) * T];
  // a_ik -= l_i u_k: the block's rows below the pivot ...
#pragma unroll
  for (int r = JL + 1; r < 8; ++r) {
    double x[11];
#pragma unroll
    for (int c = JL + 1; c < 11; ++c) x[c] = blk[(r * 11 + c) * T];
#pragma unroll
    for (int c = JL + 1; c < 11; ++c) x[c] -= lm[r] * u[c];
#pragma unroll
    for (int c = JL + 1; c < 11; ++c) blk[(r * 11 + c) * T] = x[c];
  }
  // ... and the three bottom rows (their entries in the other curve's columns meet zeros of the pivot row)
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double x[11];
#pragma unroll
    for (int c = JL + 1; c < 8; ++c) x[c] = tlb[(q * 19 + c) * T];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[8 + c] = tlt[(q * 19 + c) * T];
#pragma unroll
    for (int c = JL + 1; c < 11; ++c) x[c] -= lt[q] * u[c];
#pragma unroll
    for (int c = JL + 1; c < 8; ++c) tlb[(q * 19 + c) * T] = x[c];
#pragma unroll
    for (int c = 0; c < 3; ++c) tlt[(q * 19 + c) * T] = x[8 + c];
  }
  return true;
}

// 1: solved, 0: a row is entirely zero (the dense routine's return 0), -1: the structure does not hold for this
// system -> dense kernel
template <int T>
__device__ __forceinline__ int arrow_factor_solve(double* a, double* sc, double* bx, double (&xout)[19]) {
  constexpr int M = 19;
  {  // row scales over the stored entries (the entries left out are zeros)
    int singular = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < 16; i0 += 4) {
      double v[4][11];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 11; ++c) v[r][c] = fabs(a[((i0 + r) * 11 + c) * T]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double big = 0.0;
#pragma unroll
        for (int c = 0; c < 11; ++c)
          if (v[r][c] > big) big = v[r][c];
        if (big == 0.0) singular = 1;
        sc[(i0 + r) * T] = 1.0 / big;
      }
    }
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
      double v[19];
#pragma unroll
      for (int c = 0; c < 19; ++c) v[c] = fabs(a[(176 + q * 19 + c) * T]);
      double big = 0.0;
#pragma unroll
      for (int c = 0; c < 19; ++c)
        if (v[c] > big) big = v[c];
      if (big == 0.0) singular = 1;
      sc[(16 + q) * T] = 1.0 / big;
    }
    if (singular) return 0;
  }
  // ---- columns of the two curves: one elimination step per column, the column's position inside its curve (JL) a
  // compile-time value so that every row and column of the step has a fixed offset (no row masks, no index tables)
#pragma unroll 1
  for (int b = 0; b < 2; ++b) {
    int ok = 1;
    ok = ok && arrow_block_step<T, 0>(a, sc, bx, b);
    ok = ok && arrow_block_step<T, 1>(a, sc, bx, b);
    ok = ok && arrow_block_step<T, 2>(a, sc, bx, b);
    ok = ok && arrow_block_step<T, 3>(a, sc, bx, b);
    ok = ok && arrow_block_step<T, 4>(a, sc, bx, b);
    ok = ok && arrow_block_step<T, 5>(a, sc, bx, b);
    ok = ok && arrow_block_step<T, 6>(a, sc, bx, b);
    ok = ok && arrow_block_step<T, 7>(a, sc, bx, b);
    if (!ok) return -1;
  }
  // ---- theta, phi, z: a dense 3 x 3 on the bottom rows
  int piv = 15;
#pragma unroll 1
  for (int j = 16; j < M; ++j) {
    {
      double big = 0.0;
      for (int i = j; i < M; ++i) {
        const double merit = sc[i * T] * fabs(a[(176 + (i - 16) * 19 + j) * T]);
        if (merit >= big) { big = merit; piv = i; }
      }
      if (piv < j) return -1;  // no valid merit: the dense routine would keep a pivot row from the curves' part
    }
    if (piv != j) {
      double rp[M], rj[M];
#pragma unroll
      for (int k = 0; k < M; ++k) { rp[k] = a[(176 + (piv - 16) * 19 + k) * T]; rj[k] = a[(176 + (j - 16) * 19 + k) * T]; }
#pragma unroll
      for (int k = 0; k < M; ++k) { a[(176 + (piv - 16) * 19 + k) * T] = rj[k]; a[(176 + (j - 16) * 19 + k) * T] = rp[k]; }
      sc[piv * T] = sc[j * T];
      const double tb = bx[piv * T], tj = bx[j * T];
      bx[piv * T] = tj;
      bx[j * T] = tb;
    }
    const int jj = 176 + (j - 16) * 19 + j;
    if (a[jj * T] == 0.0) a[jj * T] = DBL_EPSILON;
    if (j != M - 1) {
      const double inv = 1.0 / a[jj * T];
      unsigned int rowmask = 0;
      for (int i = j + 1; i < M; ++i) {
        const double l = a[(176 + (i - 16) * 19 + j) * T];
        a[(176 + (i - 16) * 19 + j) * T] = l * inv;
        if (l != 0.0) rowmask |= 1u << i;
      }
      arrow_update_window<T>(a, j, 2, true, j - 16 + 1, 3, rowmask);
    }
  }
  // ---- every stored factor finite?  (then the products with structural zeros that were left out are exact zeros)
  {
    bool finite = true;
#pragma unroll 1
    for (int e0 = 0; e0 < kArrowMat - 1; e0 += 8) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = a[(e0 + q) * T];
#pragma unroll
      for (int q = 0; q < 8; ++q) finite = finite && (fabs(v[q]) <= DBL_MAX);
    }
    finite = finite && (fabs(a[(kArrowMat - 1) * T]) <= DBL_MAX);
    if (!finite) return -1;
  }
  // ---- substitutions on the permuted right-hand side, in registers
  double bv[M];
#pragma unroll
  for (int i = 0; i < M; ++i) bv[i] = bx[i * T];
  bool started = false;
#pragma unroll
  for (int j = 0; j < M - 1; ++j) {
    const double xj = bv[j];
    if (!started && xj != 0.0) started = true;
    if (started) {
      if (j < 16) {
        const int b = j >> 3, jl = j & 7;
        double l[8], lt[3];
#pragma unroll
        for (int r = jl + 1; r < 8; ++r) l[r] = a[((8 * b + r) * 11 + jl) * T];
#pragma unroll
        for (int q = 0; q < 3; ++q) lt[q] = a[(176 + q * 19 + j) * T];
#pragma unroll
        for (int r = jl + 1; r < 8; ++r) bv[8 * b + r] -= l[r] * xj;
#pragma unroll
        for (int q = 0; q < 3; ++q) bv[16 + q] -= lt[q] * xj;
      } else {
#pragma unroll
        for (int i = j + 1; i < M; ++i) bv[i] -= a[(176 + (i - 16) * 19 + j) * T] * xj;
      }
    }
  }
#pragma unroll
  for (int i = M - 1; i >= 0; --i) {
    double sum = bv[i];
    if (i >= 16) {
      double u[M];
#pragma unroll
      for (int jj = i; jj < M; ++jj) u[jj] = a[(176 + (i - 16) * 19 + jj) * T];
#pragma unroll
      for (int jj = i + 1; jj < M; ++jj) sum -= u[jj] * bv[jj];
      bv[i] = sum / u[i];
    } else {
      const int b = i >> 3, il = i & 7;
      double u[11];
#pragma unroll
      for (int c = il; c < 11; ++c) u[c] = a[(i * 11 + c) * T];
#pragma unroll
      for (int c = il + 1; c < 8; ++c) sum -= u[c] * bv[8 * b + c];  // ascending columns, like the dense loop
#pragma unroll
      for (int q = 0; q < 3; ++q) sum -= u[8 + q] * bv[16 + q];
      bv[i] = sum / u[il];
    }
  }
  bool finite = true;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    xout[i] = bv[i];
    finite = finite && (fabs(bv[i]) <= DBL_MAX);
  }
  return finite ? 1 : -1;
}

// Every entry of the solve list yields an entry of the eval list at the same position (the fit, or -1 if it stopped
// here), which keeps neighbouring fits -- neighbouring memory -- together from kernel to kernel.
// Where entry e of the row-major lower triangle goes in the block-arrow matrix: (first, second) position, -1: none.
// Entries of the zero blocks (rows 8..15 x columns 0..7) have no position at all.
__device__ __forceinline__ short2 arrow_map_entry(int e) {
  int i = 0;
  while ((i + 1) * (i + 2) / 2 <= e) ++i;
  const int k = e - i * (i + 1) / 2;
  short2 m;
  if (i < 16) {
    if ((i >> 3) != (k >> 3)) { m.x = -1; m.y = -1; return m; }
    m.x = (short)(i * 11 + (k & 7));
    m.y = (short)(i != k ? k * 11 + (i & 7) : -1);
  } else {
    const int q = i - 16;
    m.x = (short)(176 + q * 19 + k);
    m.y = (short)(k < 16 ? k * 11 + 8 + q : (i != k ? 176 + (k - 16) * 19 + i : -1));
  }
  return m;
}

// Warp-cooperative fill: the 190 + 19 doubles of a system are contiguous in global memory, so the 32 lanes read them
// as full lines and scatter them into that system's column of the work area; one system after the other, eight in
// flight (the wait for these loads was 40 % of the kernel with four).  (A thread copying its own system touches 32 different lines per instruction, 8 bytes of each.)
template <int TS>
__device__ __forceinline__ void arrow_fill_coop(double* wcol, int lane, int f, const double* jtj, const double* jte, const short2* amap) {
  wcol[lane + kArrowDoubles * TS] = 0.0;  // this thread's flag
  __syncwarp();
#pragma unroll 1
  for (int s0 = 0; s0 < 32; s0 += 8) {
    int fs[8];
    double v[8][6], r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) fs[q] = __shfl_sync(0xffffffffu, f, s0 + q);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double* lo = jtj + (size_t)(fs[q] < 0 ? 0 : fs[q]) * kLower19;
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        const int e = lane + 32 * t;
        v[q][t] = (fs[q] >= 0 && e < kLower19) ? __ldcg(lo + e) : 0.0;
      }
      r[q] = (fs[q] >= 0 && lane < 19) ? __ldcg(jte + (size_t)fs[q] * 19 + lane) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (fs[q] < 0) continue;
      double* col = wcol + (s0 + q);
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        const int e = lane + 32 * t;
        if (e < kLower19) {
          const short2 m = amap[e];
          if (m.x >= 0) col[m.x * TS] = v[q][t];
          else if (v[q][t] != 0.0) col[kArrowDoubles * TS] = 1.0;
          if (m.y >= 0) col[m.y * TS] = v[q][t];
        }
      }
      if (lane < 19) col[(kArrowMat + 19 + lane) * TS] = r[q];
    }
  }
  __syncwarp();
}

// ARROW: the block-arrow work matrix, 96 threads per CTA; a system that does not keep the structure goes to fb_list.
// SUB: the dense solve over the positions listed in fb_list.
template <bool ARROW, bool SUB>
__global__ void __launch_bounds__(ARROW ? kArrowThreads : kSolveThreads, 1) staged_solve_kernel(const StagedArgs A, int cur) {
  constexpr int M = 19, T = ARROW ? kArrowThreads : kSolveThreads;
  constexpr int TS = ARROW ? kArrowPitch : T;  // doubles between two entries of a thread's system
  constexpr int MAT = ARROW ? kArrowMat : M * M;
  extern __shared__ double smem[];
  __shared__ short2 amap[ARROW ? kLower19 : 1];
  double* a = smem + threadIdx.x;
  double* sc = smem + MAT * TS + threadIdx.x;
  double* bx = smem + (MAT + M) * TS + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (ARROW) {
    for (int e = threadIdx.x; e < kLower19; e += T) amap[e] = arrow_map_entry(e);
    __syncthreads();
  }
  const unsigned int total = SUB ? A.cnt[4] : A.cnt[2 + cur];
  for (unsigned int base = blockIdx.x * T; base < total; base += gridDim.x * T) {
    const unsigned int i0 = base + threadIdx.x;
    const bool have = i0 < total;
    const unsigned int i = have ? (SUB ? (unsigned int)A.fb_list[i0] : i0) : 0u;
    const int f = have ? A.solve_list[cur][i] : -1;
    if (ARROW) {
      // the systems of this thread's next trip start their way into the L2 now
      const unsigned int in = i0 + gridDim.x * T;
      if (in < total) {
        const int fn = A.solve_list[cur][SUB ? (unsigned int)A.fb_list[in] : in];
        const char* pj = reinterpret_cast<const char*>(A.jtj + (size_t)fn * kLower19);
#pragma unroll
        for (int l = 0; l < 13; ++l) prefetch_l2(pj + 128 * l);
        prefetch_l2(A.jte + (size_t)fn * M);
        prefetch_l2(A.jte + (size_t)fn * M + 18);
      }
      arrow_fill_coop<TS>(smem + (threadIdx.x - lane), lane, f, A.jtj, A.jte, amap);
    }
    if (!have) continue;
    const double* lower = A.jtj + (size_t)f * kLower19;
    const double* jte = A.jte + (size_t)f * M;
    const double* pp = A.L.p + (size_t)f * M;
    StagedScal* sp = A.scal + f;
    double mu = sp->mu;
    const double p_sq = sp->p_sq;
    int nu = sp->nu, nlss = sp->nlss;
    double dp_sq = sp->dp_sq;
    int stop = 0;
    bool to_eval = false;
    double dp[M];
    bool handed_over = false;
    bool filled = ARROW;  // the first system is in place (without mu)
    for (;;) {
      int solved;
      if (ARROW) {
        if (filled) {
          filled = false;
          if (a[kArrowDoubles * TS] != 0.0) { handed_over = true; break; }
          double d[M];
#pragma unroll
          for (int q = 0; q < M; ++q) d[q] = a[(q < 16 ? q * 11 + (q & 7) : 176 + (q - 16) * 19 + q) * TS];
#pragma unroll
          for (int q = 0; q < M; ++q) a[(q < 16 ? q * 11 + (q & 7) : 176 + (q - 16) * 19 + q) * TS] = d[q] + mu;
        } else {
          arrow_fill19<TS>(a, bx, lower, jte, mu);  // a re-solve with a larger mu (the previous system was singular)
        }
        solved = arrow_factor_solve<TS>(a, sc, bx, dp);
        if (A.force_handover > 0 && i % (unsigned int)A.force_handover == 0u) solved = -1;
        if (solved < 0) { handed_over = true; break; }  // nothing of the fit's state has been touched
      } else {
        lu_fill_global19(a, bx, lower, jte, mu);
        solved = lu_factor_solve<M, T>(a, sc, bx, dp);
      }
      ++nlss;
      if (solved) {
        dp_sq = 0.0;
#pragma unroll
        for (int q = 0; q < M; ++q) dp_sq += dp[q] * dp[q];
        if (dp_sq <= A.L.eps2_sq * p_sq) { stop = 2; break; }
        if (dp_sq >= (p_sq + A.L.eps2) / (kEpsilon * kEpsilon)) { stop = 4; break; }
        to_eval = true;
        break;
      }
      mu *= nu;
      const int nu2 = (int)((unsigned)nu << 1);
      if (nu2 <= nu) { stop = 5; break; }
      nu = nu2;
    }
    if (handed_over) {
      A.fb_list[atomicAdd(A.cnt + 4, 1u)] = (int)i;
      continue;
    }
    if (to_eval) {
      double pv[M], gv[M];
#pragma unroll
      for (int q = 0; q < M; ++q) { pv[q] = pp[q]; gv[q] = jte[q]; }
      double dL = 0.0;
      double* pdp = A.pdp + (size_t)f * M;
#pragma unroll
      for (int q = 0; q < M; ++q) {
        pdp[q] = pv[q] + dp[q];
        dL += dp[q] * (mu * dp[q] + gv[q]);
      }
      sp->mu = mu; sp->nu = nu; sp->nlss = nlss; sp->dp_sq = dp_sq; sp->dL = dL;
      A.eval_list[i] = f;
    } else {
      StagedScal s = *sp;
      s.mu = mu; s.nu = nu; s.nlss = nlss; s.dp_sq = dp_sq;
      s.k += 1;  // the stop happened inside the inner loop
      *sp = s;
      staged_finish(A, f, s, stop);
      A.eval_list[i] = -1;
    }
  }
}

// The host reads the list counters through pinned host memory the device writes directly (a cudaMemcpy of four
// bytes would queue on a copy engine behind the megabytes of samples the host-buffer entry point is streaming in).
// ---- fits of different lengths: the order in which a batch enters the lists -------------------------------------------
// Every phase kernel gives a lane one fit and a warp walks as many 32-row blocks as its LONGEST fit has: on contours of
// real lengths (the frames pipeline: 49 points on average, 30 to 70) a warp of 32 neighbouring fits keeps about 0.65 of
// its lanes busy.  The lists the rounds work on are compacted from the order in which the INIT launch appended the
// fits, so that order is chosen here: a counting sort of the batch (or of a piece of it) by the fits' number of row
// blocks, index order kept inside a bin up to the 1 024-fit chunk a CTA scatters (neighbours in a list stay
// neighbours in memory: a random order cost the eval kernel 3x).  Three small kernels over the offsets; per-fit results
// do not depend on list order.
constexpr int kBucketBins = 64;
constexpr int kBucketChunk = 1024;
__device__ __forceinline__ int staged_bucket_bin(const long long* __restrict__ off, int f) {
  long long nb = (off[f + 1] - off[f] + 31) >> 5;
  if (nb < 0) nb = 0;
  return (int)(nb < kBucketBins - 1 ? nb : kBucketBins - 1);
}
// the range of the fit index being ordered (shared by the der and the dif launchers)
struct BucketRange {
  const long long* off;
  int init_first, init_count;
};
__global__ void __launch_bounds__(256) staged_bucket_hist_kernel(const BucketRange A, int* __restrict__ hist, int nchunks) {
  __shared__ int h[kBucketBins];
  if (threadIdx.x < kBucketBins) h[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < kBucketChunk; i += blockDim.x) {
    const int item = blockIdx.x * kBucketChunk + i;
    if (item < A.init_count) atomicAdd(&h[staged_bucket_bin(A.off, A.init_first + item)], 1);
  }
  __syncthreads();
  if (threadIdx.x < kBucketBins) hist[threadIdx.x * nchunks + blockIdx.x] = h[threadIdx.x];
}
// exclusive scan of hist[bin][chunk] in bin-major order (one CTA of 1 024 threads)
__global__ void __launch_bounds__(1024) staged_bucket_scan_kernel(int* __restrict__ hist, int total) {
  __shared__ int part[1024];
  const int t = threadIdx.x, per = (total + 1023) / 1024;
  const int lo = min(total, t * per), hi = min(total, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += hist[i];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = part[t] - s;  // exclusive prefix of this thread's run
  for (int i = lo; i < hi; ++i) {
    const int v = hist[i];
    hist[i] = run;
    run += v;
  }
}
__global__ void __launch_bounds__(256) staged_bucket_scatter_kernel(const BucketRange A, const int* __restrict__ hist,
                                                                    int nchunks, int* __restrict__ order) {
  __shared__ int pos[kBucketBins];
  if (threadIdx.x < kBucketBins) pos[threadIdx.x] = hist[threadIdx.x * nchunks + blockIdx.x];
  __syncthreads();
  for (int i = threadIdx.x; i < kBucketChunk; i += blockDim.x) {
    const int item = blockIdx.x * kBucketChunk + i;
    if (item < A.init_count) {
      const int f = A.init_first + item;
      order[atomicAdd(&pos[staged_bucket_bin(A.off, f)], 1)] = f;
    }
  }
}

// End of a round: the length of the next solve list for the host, the jac list length of the round that just ended
// (cnt[0]: the host's estimate for the next round), and the running total of Jacobian evaluations the jac kernel has
// performed in this call (cnt[1]).
__global__ void staged_round_end_kernel(unsigned int* cnt, int nxt, volatile unsigned int* host) {
  cnt[5] += cnt[4];
  cnt[4] = 0;
  host[3] = cnt[5];
  cnt[1] += cnt[0];
  host[2] = cnt[1];
  host[1] = cnt[0];
  host[0] = cnt[2 + nxt];
  __threadfence_system();
}

// ---- jac: thread per fit ---------------------------------------------------------------------------------
// Shared memory per thread (strided by the CTA size: a warp's accesses are conflict-free):
//   D[40]  derivative table of the current contour segment: [u-row | v-row][control point k][xe, ye, theta, phi, z]
//   C[8]   projected control points of the segment: u of cp 0..3, v of cp 0..3
// plus the warp's staging tiles.  Registers: the partial sums of the current 32-row block (36 or 30 at a time) and
// the 11 J^T e chains of the current curve.  The running J^T J entries are touched once per block and live in the global work
// area (A.ws, [slot][thread]: coalesced, L2-resident).
__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// The segment tables are re-read from shared memory for every row.  A plain load would be hoisted out of the row
// loop (the tables are loop-invariant) into 96 registers, which the partial sums need; `volatile` pins it.
__device__ __forceinline__ double lds_pinned(unsigned int saddr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
  return v;
}

// A block's rows are walked twice, each time for one half of the entries (pass A / pass B below), so that the partial
// sums in flight (36 or 30) leave registers for a second set holding the products of the next row.  The residual,
// the Bernstein weights and the 8 control-point entries of a row are recomputed in pass B; same expressions, same
// values.
constexpr unsigned int kJacPitch = kJacThreads * 8;  // byte pitch between table entries

// A pass is software-pipelined over rows by hand: the products of a row are formed in one loop trip and added to
// the partial sums in the next, so that no add ever waits for its own multiply (a multiply-add pair issued back to
// back stalls for the FP64 latency, and with two resident warps per scheduler nothing else hides it).  The row
// index r = 2*sample + (0: u row, 1: v row) selects the half of the tables through an address offset.
struct RowIn {
  double w[4];
  double e;
};

__device__ __forceinline__ void jac_row_in(unsigned int Csa, const double* xt, const double* tt, int r, RowIn& in) {
  bernstein3(tt[(r >> 1) & 15], in.w);
  const unsigned int c = Csa + (r & 1) * (4 * kJacPitch);
  double au = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) au += in.w[k] * lds_pinned(c + k * kJacPitch);
  in.e = xt[r & 31] - au;
}

__device__ __forceinline__ void jac_row_cp(unsigned int Dsa, int r, const double (&w)[4], double (&v)[8]) {
  const unsigned int d = Dsa + (r & 1) * (20 * kJacPitch);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = w[k] * lds_pinned(d + (5 * k + 0) * kJacPitch);
    v[2 * k + 1] = w[k] * lds_pinned(d + (5 * k + 1) * kJacPitch);
  }
}

__device__ __forceinline__ void jac_products_a(unsigned int Dsa, unsigned int Csa, const double* xt, const double* tt,
                                               int r, double (&pr)[36], double (&pe)[8]) {
  RowIn in;
  jac_row_in(Csa, xt, tt, r, in);
  double v[8];
  jac_row_cp(Dsa, r, in.w, v);
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) pr[a * (a + 1) / 2 + b] = v[a] * v[b];
#pragma unroll
  for (int a = 0; a < 8; ++a) pe[a] = v[a] * in.e;
}

__device__ __forceinline__ void jac_products_b(unsigned int Dsa, unsigned int Csa, const double* xt, const double* tt,
                                               int r, double (&pr)[30], double (&pe)[3]) {
  RowIn in;
  jac_row_in(Csa, xt, tt, r, in);
  double v[8];
  jac_row_cp(Dsa, r, in.w, v);
  const unsigned int d = Dsa + (r & 1) * (20 * kJacPitch);
  double vq[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) a += in.w[k] * lds_pinned(d + (5 * k + 2 + q) * kJacPitch);
    vq[q] = a;
  }
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int b = 0; b < 8; ++b) pr[8 * q + b] = vq[q] * v[b];
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int r2 = 0; r2 <= q; ++r2) pr[24 + q * (q + 1) / 2 + r2] = vq[q] * vq[r2];
#pragma unroll
  for (int q = 0; q < 3; ++q) pe[q] = vq[q] * in.e;
}

// pass A: control-point x control-point entries of the row's curve (36) and their 8 J^T e chains
__device__ __forceinline__ void jac_pass_a(unsigned int Dsa, unsigned int Csa, const double* xt, const double* tt,
                                           int s0, int s1, double (&pa)[36], double (&jte_cp)[8]) {
  if (s0 >= s1) return;
#ifdef CPS_FMA
  // the optional fused arithmetic (cps_lm_fast.cu: this header compiled with multiply-add contraction): one fused
  // multiply-add per sum and row, no products in flight; NOT the reference's operation sequence
#pragma unroll 2
  for (int r = 2 * s0; r < 2 * s1; ++r) {
    RowIn in;
    jac_row_in(Csa, xt, tt, r, in);
    double v[8];
    jac_row_cp(Dsa, r, in.w, v);
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) pa[a * (a + 1) / 2 + b] = fma(v[a], v[b], pa[a * (a + 1) / 2 + b]);
#pragma unroll
    for (int a = 0; a < 8; ++a) jte_cp[a] = fma(v[a], in.e, jte_cp[a]);
  }
  return;
#endif
  double pr[36], pe[8];
  jac_products_a(Dsa, Csa, xt, tt, 2 * s0, pr, pe);
#pragma unroll 2  // two rows per trip give the scheduler independent work to interleave (measured: -5 %; 4: worse)
  for (int r = 2 * s0 + 1; r < 2 * s1; ++r) {
#pragma unroll
    for (int q = 0; q < 36; ++q) pa[q] += pr[q];
#pragma unroll
    for (int a = 0; a < 8; ++a) jte_cp[a] += pe[a];
    jac_products_a(Dsa, Csa, xt, tt, r, pr, pe);
  }
#pragma unroll
  for (int q = 0; q < 36; ++q) pa[q] += pr[q];
#pragma unroll
  for (int a = 0; a < 8; ++a) jte_cp[a] += pe[a];
}

// pass B: theta/phi/z x control-point (24), theta/phi/z x theta/phi/z (6) and the 3 theta/phi/z J^T e chains
__device__ __forceinline__ void jac_pass_b(unsigned int Dsa, unsigned int Csa, const double* xt, const double* tt,
                                           int s0, int s1, double (&pb)[30], double (&jte_q)[3]) {
  if (s0 >= s1) return;
#ifdef CPS_FMA
#pragma unroll 2
  for (int r = 2 * s0; r < 2 * s1; ++r) {
    RowIn in;
    jac_row_in(Csa, xt, tt, r, in);
    double v[8];
    jac_row_cp(Dsa, r, in.w, v);
    const unsigned int d = Dsa + (r & 1) * (20 * kJacPitch);
    double vq[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) a = fma(in.w[k], lds_pinned(d + (5 * k + 2 + q) * kJacPitch), a);
      vq[q] = a;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int b = 0; b < 8; ++b) pb[8 * q + b] = fma(vq[q], v[b], pb[8 * q + b]);
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int r2 = 0; r2 <= q; ++r2) pb[24 + q * (q + 1) / 2 + r2] = fma(vq[q], vq[r2], pb[24 + q * (q + 1) / 2 + r2]);
#pragma unroll
    for (int q = 0; q < 3; ++q) jte_q[q] = fma(vq[q], in.e, jte_q[q]);
  }
  return;
#endif
  double pr[30], pe[3];
  jac_products_b(Dsa, Csa, xt, tt, 2 * s0, pr, pe);
#pragma unroll 2  // two rows per trip give the scheduler independent work to interleave (measured: -5 %; 4: worse)
  for (int r = 2 * s0 + 1; r < 2 * s1; ++r) {
#pragma unroll
    for (int q = 0; q < 30; ++q) pb[q] += pr[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) jte_q[q] += pe[q];
    jac_products_b(Dsa, Csa, xt, tt, r, pr, pe);
  }
#pragma unroll
  for (int q = 0; q < 30; ++q) pb[q] += pr[q];
#pragma unroll
  for (int q = 0; q < 3; ++q) jte_q[q] += pe[q];
}

#ifdef CPS_FMA
// fused arithmetic, one thread per fit: both halves of the entries in ONE walk over the rows (no products in flight, so
// all 66 sums fit in registers and the row inputs are formed once)
__device__ __forceinline__ void jac_pass_ab(unsigned int Dsa, unsigned int Csa, const double* xt, const double* tt, int s0,
                                            int s1, double (&pa)[36], double (&pb)[30], double (&jte_cp)[8],
                                            double (&jte_q)[3]) {
#pragma unroll 2
  for (int r = 2 * s0; r < 2 * s1; ++r) {
    RowIn in;
    jac_row_in(Csa, xt, tt, r, in);
    double v[8];
    jac_row_cp(Dsa, r, in.w, v);
    const unsigned int d = Dsa + (r & 1) * (20 * kJacPitch);
    double vq[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) a = fma(in.w[k], lds_pinned(d + (5 * k + 2 + q) * kJacPitch), a);
      vq[q] = a;
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) pa[a * (a + 1) / 2 + b] = fma(v[a], v[b], pa[a * (a + 1) / 2 + b]);
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int b = 0; b < 8; ++b) pb[8 * q + b] = fma(vq[q], v[b], pb[8 * q + b]);
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int r2 = 0; r2 <= q; ++r2) pb[24 + q * (q + 1) / 2 + r2] = fma(vq[q], vq[r2], pb[24 + q * (q + 1) / 2 + r2]);
#pragma unroll
    for (int a = 0; a < 8; ++a) jte_cp[a] = fma(v[a], in.e, jte_cp[a]);
#pragma unroll
    for (int q = 0; q < 3; ++q) jte_q[q] = fma(vq[q], in.e, jte_q[q]);
  }
}
#endif

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

// SPLIT: when fewer fits than half the resident threads need a Jacobian (the late rounds of a batch, mid-size
// batches), a fit gets two threads -- warps 0,1 of a CTA run pass A for 64 fits, warps 2,3 pass B for the same
// fits -- which halves the latency floor of a round (one fit per thread takes 250 us however few fits there are).
// The two halves write disjoint entries; whichever finishes second (an atomic counter in the fit's state) computes
// the gradient statistics from the finished arrays and hands the fit on.
template <bool SPLIT>
__global__ void __launch_bounds__(kJacThreads, 2) staged_jac_kernel(const StagedArgs A, int nxt) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool doA = !SPLIT || wid < 2, doB = !SPLIT || wid >= 2;
  double* Dsm = smem + tid;
  double* Csm = smem + 40 * kJacThreads + tid;
  const unsigned int Dsa = (unsigned int)__cvta_generic_to_shared(Dsm);
  const unsigned int Csa = (unsigned int)__cvta_generic_to_shared(Csm);
  WarpTiles* wt = reinterpret_cast<WarpTiles*>(smem + kJacTableDoubles * kJacThreads) + wid;
  const int gtid = blockIdx.x * kJacThreads + tid;
  const size_t wsn = (size_t)A.ws_stride;
  double* ws = A.ws + gtid;
  const unsigned int total = A.cnt[0];
  const unsigned int warp0 = SPLIT ? (blockIdx.x * 2 + (wid & 1)) * 32 : (blockIdx.x * (kJacThreads / 32) + wid) * 32;
  const unsigned int stride = SPLIT ? gridDim.x * (kJacThreads / 2) : gridDim.x * kJacThreads;
  for (unsigned int base = warp0; base < total; base += stride) {
    const unsigned int item = base + lane;
    bool have = item < total;
    const int f = have ? A.jac_list[item] : 0;
    FitGeom g;
    g.n = 0; g.nsamp = 0; g.b1 = g.b2 = g.b3 = 0;
    long long o0 = 0;
    if (have) staged_geom(A, f, g, o0);
    if (SPLIT && !doA && g.n * 19 < kBlockSzSq) have = false;  // a small (unblocked) fit is left to its pass-A thread
    const double* x = A.L.meas + o0;
    const double* tp = have ? staged_t(A, f, o0) : nullptr;
    const double* pp = A.L.p + (size_t)f * 19;
    double* Jout = A.jtj + (size_t)f * kLower19;
    double* Jte = A.jte + (size_t)f * 19;
    const bool blocked = have && !(g.n * 19 < kBlockSzSq);  // strict "<", lm_core.c:193
    const int xoff = ptr_parity(x), toff = tp ? ptr_parity(tp) : 0;
    Trig tg;
    if (have) make_trig(pp[16], pp[17], tg);
    double jte_cp[8], jte_q[3];
#pragma unroll
    for (int q = 0; q < 8; ++q) jte_cp[q] = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) jte_q[q] = 0.0;
    if (blocked) {
#pragma unroll 1
      for (int q = 0; q < kWsPend; ++q) __stcg(ws + q * wsn, 0.0);
    }
    const int nblk = blocked ? (g.n + 31) >> 5 : 0;
    int cur_seg = -1, cur_curve = -1;
    const double* xt = &wt->x[lane][xoff];
    const double* tt = &wt->t[lane][toff];
    // entering contour segment `seg`: the J^T e chains of the control-point columns follow the curve, the tables the
    // (curve, image side)
    auto enter_segment = [&](int seg) {
      const int curve = seg & 1;
      if (curve != cur_curve) {
        if (cur_curve >= 0) {
#pragma unroll
          for (int a = 0; a < 8; ++a) __stcg(ws + (kWsJte + 8 * cur_curve + a) * wsn, jte_cp[a]);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) jte_cp[a] = __ldcg(ws + (kWsJte + 8 * curve + a) * wsn);
        cur_curve = curve;
      }
      jac_segment_tables(pp, tg, curve, seg >> 1, Dsm, Csm);
      cur_seg = seg;
    };
#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
      row_fill(const_cast<double*>(xt) - xoff, xoff, x + 32 * b, min(32, g.n - 32 * b));
      row_fill(const_cast<double*>(tt) - toff, toff, tp + 16 * b, min(16, g.nsamp - 16 * b));
      if (b + 1 < nblk) {
        prefetch_l2(x + 32 * (b + 1));
        if (32 * (b + 1) + 16 < g.n) prefetch_l2(x + 32 * (b + 1) + 16);
        prefetch_l2(tp + 16 * (b + 1));
      }
      cp_async_wait_all();
      {
        const int s_lo = 16 * b, blk_end = min(16 * b + 16, g.nsamp);
        const int seg_lo = (s_lo >= g.b1) + (s_lo >= g.b2) + (s_lo >= g.b3);
        const int seg_hi = (blk_end - 1 >= g.b1) + (blk_end - 1 >= g.b2) + (blk_end - 1 >= g.b3);
        if (__all_sync(__activemask(), seg_lo == seg_hi)) {
          // every lane's block lies inside one contour segment (always, for batches of equal-length contours): partial
          // sums start at zero in registers and join the running entries at the end of the pass (misc_core.c:118-123)
          if (seg_lo != cur_seg) enter_segment(seg_lo);
          double* acc = ws + (size_t)(kWsAcc + 60 * cur_curve) * wsn;
#ifdef CPS_FMA
          if (doA && doB) {
            double pa[36], pb[30];
#pragma unroll
            for (int q = 0; q < 36; ++q) pa[q] = 0.0;
#pragma unroll
            for (int q = 0; q < 30; ++q) pb[q] = 0.0;
            jac_pass_ab(Dsa, Csa, xt, tt, s_lo, blk_end, pa, pb, jte_cp, jte_q);
#pragma unroll
            for (int q = 0; q < 36; ++q) red_add_f64(acc + q * wsn, pa[q]);
#pragma unroll
            for (int q = 0; q < 24; ++q) red_add_f64(acc + (36 + q) * wsn, pb[q]);
            double* accq = ws + (size_t)(kWsAcc + 120) * wsn;
#pragma unroll
            for (int q = 0; q < 6; ++q) red_add_f64(accq + q * wsn, pb[24 + q]);
          } else
#endif
          {
          if (doA) {
            double pa[36];
#pragma unroll
            for (int q = 0; q < 36; ++q) pa[q] = 0.0;
            jac_pass_a(Dsa, Csa, xt, tt, s_lo, blk_end, pa, jte_cp);
#pragma unroll
            for (int q = 0; q < 36; ++q) red_add_f64(acc + q * wsn, pa[q]);
          }
          if (doB) {
            double pb[30];
#pragma unroll
            for (int q = 0; q < 30; ++q) pb[q] = 0.0;
            jac_pass_b(Dsa, Csa, xt, tt, s_lo, blk_end, pb, jte_q);
#pragma unroll
            for (int q = 0; q < 24; ++q) red_add_f64(acc + (36 + q) * wsn, pb[q]);
            double* accq = ws + (size_t)(kWsAcc + 120) * wsn;
#pragma unroll
            for (int q = 0; q < 6; ++q) red_add_f64(accq + q * wsn, pb[24 + q]);
          }
          }
        } else if (seg_hi <= seg_lo + 1) {
          // The block lies inside one contour segment, or holds the end of one and the start of the next (with
          // contours of real lengths: one block in four, at different rows in every lane of a warp).  ONE code path
          // for both, so that the lanes of a warp stay together: two pieces, the second one empty for a lane whose
          // block does not cross a boundary.  Partial sums start at zero in registers and join the running entries
          // when their piece ends (misc_core.c:118-123: the two pieces of a crossing block belong to different curves,
          // so the first curve's partial sums are complete there); only the six theta/phi/z x theta/phi/z sums run on
          // over both pieces.  The additions are fire-and-forget reductions performed at the L2 (red.global.add.f64:
          // the same IEEE round-to-nearest addition; one owner per address, program order per address): no load
          // latency is exposed.
          const int mid = (seg_hi == seg_lo) ? blk_end : seg_end(g, seg_lo);
          double pq[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
          for (int piece = 0; piece < 2; ++piece) {
            // (no lane of the warp has a second piece: batches of equal-length contours never do)
            if (piece == 1 && !__any_sync(__activemask(), mid < blk_end)) break;
            const int p0 = piece ? mid : s_lo, p1 = piece ? blk_end : mid;
            if (p0 < p1 && seg_lo + piece != cur_seg) enter_segment(seg_lo + piece);
            double* acc = ws + (size_t)(kWsAcc + 60 * (cur_curve < 0 ? 0 : cur_curve)) * wsn;
#ifdef CPS_FMA
            if (doA && doB) {
              double pa[36], pb[30];
#pragma unroll
              for (int q = 0; q < 36; ++q) pa[q] = 0.0;
#pragma unroll
              for (int q = 0; q < 24; ++q) pb[q] = 0.0;
#pragma unroll
              for (int q = 0; q < 6; ++q) pb[24 + q] = pq[q];
              jac_pass_ab(Dsa, Csa, xt, tt, p0, p1, pa, pb, jte_cp, jte_q);
              if (p0 < p1) {
#pragma unroll
                for (int q = 0; q < 36; ++q) red_add_f64(acc + q * wsn, pa[q]);
#pragma unroll
                for (int q = 0; q < 24; ++q) red_add_f64(acc + (36 + q) * wsn, pb[q]);
              }
#pragma unroll
              for (int q = 0; q < 6; ++q) pq[q] = pb[24 + q];
              continue;
            }
#endif
            if (doA) {
              double pa[36];
#pragma unroll
              for (int q = 0; q < 36; ++q) pa[q] = 0.0;
              jac_pass_a(Dsa, Csa, xt, tt, p0, p1, pa, jte_cp);
              if (p0 < p1) {
#pragma unroll
                for (int q = 0; q < 36; ++q) red_add_f64(acc + q * wsn, pa[q]);
              }
            }
            if (doB) {
              double pb[30];
#pragma unroll
              for (int q = 0; q < 24; ++q) pb[q] = 0.0;
#pragma unroll
              for (int q = 0; q < 6; ++q) pb[24 + q] = pq[q];
              jac_pass_b(Dsa, Csa, xt, tt, p0, p1, pb, jte_q);
              if (p0 < p1) {
#pragma unroll
                for (int q = 0; q < 24; ++q) red_add_f64(acc + (36 + q) * wsn, pb[q]);
              }
#pragma unroll
              for (int q = 0; q < 6; ++q) pq[q] = pb[24 + q];
            }
          }
          if (doB) {
            double* accq = ws + (size_t)(kWsAcc + 120) * wsn;
#pragma unroll
            for (int q = 0; q < 6; ++q) red_add_f64(accq + q * wsn, pq[q]);
          }
        } else {
          // the block straddles three or more segments: its partial sums are kept per curve in the work area while
          // the pieces are walked (a curve's rows come in two pieces: curve 1, a short curve 2, curve 1 again), and
          // join the running entries once, at the end of the block
#pragma unroll 1
          for (int q = kWsPend; q < kWsSlots; ++q) __stcg(ws + q * wsn, 0.0);
          unsigned seen = 0;
          int s = s_lo;
#pragma unroll 1
          while (s < blk_end) {
            const int seg = (s >= g.b1) + (s >= g.b2) + (s >= g.b3);
            if (seg != cur_seg) enter_segment(seg);
            const int e_end = min(blk_end, seg_end(g, seg));
            double* pend = ws + (size_t)(kWsPend + 60 * cur_curve) * wsn;
            double* pendq = ws + (size_t)kWsPendQ * wsn;
            if (doA) {
              double pa[36];
#pragma unroll
              for (int q = 0; q < 36; ++q) pa[q] = __ldcg(pend + q * wsn);
              jac_pass_a(Dsa, Csa, xt, tt, s, e_end, pa, jte_cp);
#pragma unroll
              for (int q = 0; q < 36; ++q) __stcg(pend + q * wsn, pa[q]);
            }
            if (doB) {
              double pb[30];
#pragma unroll
              for (int q = 0; q < 24; ++q) pb[q] = __ldcg(pend + (36 + q) * wsn);
#pragma unroll
              for (int q = 0; q < 6; ++q) pb[24 + q] = __ldcg(pendq + q * wsn);
              jac_pass_b(Dsa, Csa, xt, tt, s, e_end, pb, jte_q);
#pragma unroll
              for (int q = 0; q < 24; ++q) __stcg(pend + (36 + q) * wsn, pb[q]);
#pragma unroll
              for (int q = 0; q < 6; ++q) __stcg(pendq + q * wsn, pb[24 + q]);
            }
            seen |= 1u << cur_curve;
            s = e_end;
          }
#pragma unroll 1
          for (int c = 0; c < 2; ++c)
            if (seen & (1u << c))
              for (int q = 0; q < 60; ++q)
                red_add_f64(ws + (kWsAcc + 60 * c + q) * wsn, __ldcg(ws + (kWsPend + 60 * c + q) * wsn));
#pragma unroll 1
          for (int q = 0; q < 6; ++q) red_add_f64(ws + (kWsAcc + 120 + q) * wsn, __ldcg(ws + (kWsPendQ + q) * wsn));
        }
      }
    }
    bool to_solve = false;
    if (have) {
    // gradient statistics, lm_core.c:256-287 (maxima do not depend on the order of the comparisons)
    double p_sq = 0.0, g_inf = 0.0, big = -DBL_MAX;
    if (blocked) {
      if (cur_curve >= 0) {
#pragma unroll
        for (int a = 0; a < 8; ++a) __stcg(ws + (kWsJte + 8 * cur_curve + a) * wsn, jte_cp[a]);
      }
      __threadfence();  // the block folds were reductions at the L2: order this thread's reads of the sums after them
      // the lower triangle, row-major: curve blocks, the structurally zero cross block, theta/phi/z rows.  Every
      // group is read into registers before it is written so that the loads overlap (the work area is read past the
      // L1: its entries were updated by reductions at the L2).
      if (doA) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double v[36];
#pragma unroll
        for (int q = 0; q < 36; ++q) v[q] = __ldcg(ws + (kWsAcc + 60 * c + q) * wsn);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int bq = 0; bq <= a; ++bq) Jout[tri(8 * c + a, 8 * c + bq)] = v[a * (a + 1) / 2 + bq];
#pragma unroll
        for (int a = 0; a < 8; ++a)
          if (v[a * (a + 1) / 2 + a] > big) big = v[a * (a + 1) / 2 + a];
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int bq = 0; bq < 8; ++bq) Jout[tri(8 + a, bq)] = 0.0;
      }
      if (doB) {
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
          if (v[48 + q * (q + 1) / 2 + q] > big) big = v[48 + q * (q + 1) / 2 + q];
        }
      }
      if (doA) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __ldcg(ws + (kWsJte + q) * wsn);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          Jte[q] = v[q];
          const double tmp = fabs(v[q]);
          if (g_inf < tmp) g_inf = tmp;
        }
      }
      if (doB) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          Jte[16 + q] = jte_q[q];
          const double tmp = fabs(jte_q[q]);
          if (g_inf < tmp) g_inf = tmp;
        }
      }
    } else {
      jac_unblocked(pp, tg, g, x, tp, Dsm, Csm, Jout, Jte);
#pragma unroll 1
      for (int q = 0; q < 19; ++q) {
        const double tmp = fabs(Jte[q]);
        if (g_inf < tmp) g_inf = tmp;
        const double d = Jout[tri(q, q)];
        if (d > big) big = d;
      }
    }
    {
      double pv[19];
#pragma unroll
      for (int q = 0; q < 19; ++q) pv[q] = pp[q];
#pragma unroll
      for (int q = 0; q < 19; ++q) p_sq += pv[q] * pv[q];
    }
    StagedScal* sp = A.scal + f;
    bool last = true;
    if (SPLIT && blocked) {
      // publish this half, then see whether the other half is already there
      __threadfence();
      last = atomicAdd(&sp->pad0, 1) == 1;
      if (last) {
        sp->pad0 = 0;
        __threadfence();
        g_inf = 0.0;
        big = -DBL_MAX;
        double jv[19], dv[19];
#pragma unroll
        for (int q = 0; q < 19; ++q) { jv[q] = __ldcg(Jte + q); dv[q] = __ldcg(Jout + tri(q, q)); }
#pragma unroll
        for (int q = 0; q < 19; ++q) {
          const double tmp = fabs(jv[q]);
          if (g_inf < tmp) g_inf = tmp;
          if (dv[q] > big) big = dv[q];
        }
      }
    }
    if (last) {
    StagedScal sc = *sp;
    sc.njev += 1;
    sc.g_inf = g_inf;
    sc.p_sq = p_sq;
    sc.maxdiag = big;
    if (g_inf <= A.L.eps1) {
      sc.dp_sq = 0.0;
      *sp = sc;
      staged_finish(A, f, sc, 1);
    } else {
      if (sc.k == 0) sc.mu = A.L.tau * big;
      *sp = sc;
      to_solve = true;
    }
    }
    }
    __syncwarp();
    staged_append(A.solve_list[nxt], A.cnt + 2 + nxt, f, to_solve);
  }
}

}  // namespace cps
