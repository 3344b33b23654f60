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

#ifndef CPS_SP_NG
#define CPS_SP_NG 4
#endif
#ifndef CPS_LM_MIN_BLOCKS
#define CPS_LM_MIN_BLOCKS 3
#endif
constexpr int kWarpsPerCta = 4;
constexpr int kTileRows = 32;  // the reference's __BLOCKSZ__ (misc.h:48): row blocks of J^T J partial sums
constexpr double kEpsilon = 1E-12;          // lm.c:35
constexpr double kOneThird = 0.3333333334;  // lm.c:36 (sic)
constexpr int kBlockSzSq = 32 * 32;         // lm_core.c:193,585 switch between the two J^T J loops

// Per-fit scalars of a dlevmar_der iteration in flight (the staged shape's state, cps_lm_staged.cuh).  A fit parked
// here is about to solve (J^T J + mu I) dp = J^T e: every check of the current outer iteration has passed and its
// Jacobian evaluation has been counted.
struct LmResume {
  double e0, e_sq, g_inf, dp_sq, mu, p_sq, maxdiag, dL;
  int nu, k, nfev, njev, nlss, pad0, pad1, pad2;
};

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
  double* jstore;           // per resident warp: dif [M][n_max] column-major Jacobian; der [n_max] residual + [n_max/2] t
  unsigned int* queue;      // next fit index
  const int* fit_list;      // nullptr, or the B fit indices to work on (der: the staged shape's late fits)
  const LmResume* resume;   // with fit_list: [fit] iteration state to continue from,
  const double* resume_jtj; //   [fit][M(M+1)/2] lower triangle of J^T J at the fit's current p, row-major,
  const double* resume_jte; //   [fit][M] J^T e
};

__host__ __device__ constexpr int lm_ldt(int M) { return M | 1; }  // tile stride of the unblocked (small-problem) path

// Geometry of the blocked J^T J accumulation (see BlockAccum).  SP = the compact 12-column tile of the
// Stereo19 analytic Jacobian (11 of 19 entries per row are non-zero); otherwise dense tiles.
template <int M, bool SP>
struct Blk {
  static constexpr int NB = (M + 3) / 4;                         // groups of 4 columns
  static constexpr int R = SP ? 6 : NB * (NB + 1) / 2;           // roles = 4x4 blocks of J^T J entries
  static constexpr int NG = SP ? CPS_SP_NG : ((32 / R) >= 4 ? 4 : (32 / R));  // tiles processed at once
  static constexpr int TC = SP ? 12 : NB * 4;                    // tile columns
  static constexpr int LDT = SP ? TC : TC + 2;                   // even: rows stay 16-byte aligned
  // +6 doubles (48 bytes) between tiles: lanes of neighbouring groups read the same columns of different tiles,
  // the skew puts their 16-byte operand pairs on different banks
  static constexpr int TILE_DOUBLES = kTileRows * LDT + 6;
  static constexpr int DOUBLES = NG * TILE_DOUBLES + NG * kTileRows / 2 /*row tags*/ + 32 /*item list, flags*/ +
                                 NG * kTileRows /*residuals of the tiles' rows*/;
};

__host__ __device__ constexpr int lm_tile_doubles(int M, bool sparse) {
  return M == 19 ? (sparse ? Blk<19, true>::DOUBLES : Blk<19, false>::DOUBLES) : Blk<8, false>::DOUBLES;
}

// the tile region doubles as the trial-residual buffer (tiles are only alive inside the normal-equation build)
__host__ __device__ constexpr int lm_tile_region(int M, int n_max, bool dif) {
  return lm_tile_doubles(M, !dif && M == 19) < n_max ? n_max : lm_tile_doubles(M, !dif && M == 19);
}

// per-warp global (L2-resident) work area, doubles: the accepted residual e and t; dif adds f(p), f(p+dp) and the
// column-major secant Jacobian.  Everything in it is only streamed with lane-coalesced accesses.
__host__ __device__ constexpr int lm_ws_doubles(int M, int n_max, bool dif) {
  return n_max + n_max / 2 + (dif ? 3 * n_max + M * n_max : 0);
}

// doubles of shared memory one warp needs
__host__ __device__ constexpr int lm_warp_smem_doubles(int M, int scratch, int n_max, bool dif) {
  // n_max is a multiple of 4 and the total is rounded to an even count so that every warp's tiles stay
  // 16-byte aligned (the 4x4 blocks read operand pairs)
  // The samples x stay in global memory (read once per residual, coalesced).  der: the trial residual lives in
  // the tile region (tiles are only alive inside the normal-equation build); dif keeps hx, wrk, wrk2.
  // der additionally keeps the accepted residual e and t in a per-warp global (L2-resident) work area: both are
  // only streamed, lane-coalesced; e reaches the J^T e chain through the spare column of the compact tile.
  return ((lm_tile_region(M, n_max, dif) + M * M /*jtj*/ + 8 * 32 /*small vectors*/ + scratch) + 1) & ~1;
}

template <int M>
struct WarpMem {
  const double* x;  // the fit's samples in global memory
  double *t, *e, *hx, *wrk, *wrk2, *dbuf, *tile, *jtj;
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

// Trial evaluation of the secant path in one pass over the samples: f(p+dp) -> fout (global), the trial residual
// x - f(p+dp) -> eout (shared memory, fed to the norm chains without a round trip through L2), and
// f(p+dp) - f(p) -> dout (what a deferred secant update needs; an accepted step overwrites f(p)).
template <class MD>
__device__ __forceinline__ void eval_trial(const WarpMem<MD::M>& w, const FitGeom& g, const double* p, double* fout,
                                           double* eout, double* dout, const double* fold, int lane) {
  MD::prep_func(p, w.scratch, lane);
  __syncwarp();
  for (int s = lane; s < g.nsamp; s += 32) {
    double hu, hv;
    MD::eval_sample(w.scratch, g, s, w.t[s], hu, hv);
    const double2 xv = *reinterpret_cast<const double2*>(w.x + 2 * s);
    const double2 fo = *reinterpret_cast<const double2*>(fold + 2 * s);
    *reinterpret_cast<double2*>(fout + 2 * s) = make_double2(hu, hv);
    *reinterpret_cast<double2*>(eout + 2 * s) = make_double2(xv.x - hu, xv.y - hv);
    *reinterpret_cast<double2*>(dout + 2 * s) = make_double2(hu - fo.x, hv - fo.y);
  }
  __syncwarp();
}

// e = x - y, returns ||e||^2 in the summation order of dlevmar_L2nrmxmy (misc_core.c:712-798): four
// running sums fed from the top index downwards in groups of eight, ragged tail upwards, then
// ((s0+s1)+s2)+s3.  The four chains run on lanes 0..3.  e may alias y.
__device__ __forceinline__ double residual_chains(const double* e, int n, int lane);
__device__ __forceinline__ double residual_sqnorm(double* e, const double* x, const double* y, int n, int lane) {
  for (int i = lane; i < n; i += 32) e[i] = x[i] - y[i];
  __syncwarp();
  return residual_chains(e, n, lane);
}
// ||e||^2 of a residual already stored (shared memory): the four reference chains on lanes 0..3
__device__ __forceinline__ double residual_chains(const double* e, int n, int lane) {
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

// ---- blocked J^T J: 4x4 register blocks, several 32-row tiles in flight ------------------------------------
// The reference's blocked product (misc_core.c:101-126) gives every entry b_ij = sum over 32-row blocks (ascending)
// of a partial sum that starts at 0 and runs over the block's rows (ascending).  Partial sums of different
// blocks are independent, so NG tiles are processed at once by NG groups of lanes; within a group every lane
// owns one 4x4 block of entries ("role": 8 operand loads feed 16 multiply-adds instead of 2 feeding 1), and the
// partial sums are folded into the owners' accumulators in ascending block order through shuffles.
// J^T e has no block structure in the reference (one running sum over ascending rows, lm_core.c:246-254), so it
// stays a chain per column, lanes = columns, over the tiles in order.
template <int M, bool SP>
struct BlockAccum {
  using B = Blk<M, SP>;
  double jte;
  int grp, role, lbase, rbase;  // tile column of the first L / R operand of this lane's block
  int out0, na, nb;             // owner side: jtj index of entry (0,0) of the block, valid rows / columns
  bool active, is_c_role, diag;

  __device__ __forceinline__ void init(int lane) {
    grp = lane / B::R;
    role = lane - grp * B::R;
    active = grp < B::NG;
    if (!active) { grp = 0; role = 0; }
    int bi = 0, bj = 0;
    if (SP) {  // (g0,g0) (g1,g0) (g1,g1) (gq,g0) (gq,g1) (gq,gq); tile columns: g0 = 0..3, g1 = 4..7, gq = 8..11
      bi = (role >= 3) ? 2 : (role >= 1 ? 1 : 0);
      bj = (role == 2 || role == 4) ? 1 : (role == 5 ? 2 : 0);
    } else {
      int r = role;
      while (r > bi) { r -= bi + 1; ++bi; }
      bj = r;
    }
    lbase = 4 * bi;
    rbase = 4 * bj;
    is_c_role = SP && role == 5;
    diag = bi == bj;
    const int oc = grp & 1;  // an owner of curve entries sits in group 0 (curve 0) or group 1 (curve 1)
    const int ri0 = SP ? (lbase == 8 ? 16 : 8 * oc + lbase) : lbase;
    const int rj0 = SP ? (rbase == 8 ? 16 : 8 * oc + rbase) : rbase;
    na = SP ? (lbase == 8 ? 3 : 4) : min(4, M - lbase);
    nb = SP ? (rbase == 8 ? 3 : 4) : min(4, M - rbase);
    out0 = ri0 * M + rj0;
  }
  // one row of the block: 4 + 4 operands, 16 multiply-adds
  __device__ __forceinline__ void mac_row(const double* row, double (&part)[16]) const {
    const double2 l01 = *reinterpret_cast<const double2*>(row + lbase);
    const double2 l23 = *reinterpret_cast<const double2*>(row + lbase + 2);
    const double2 r01 = *reinterpret_cast<const double2*>(row + rbase);
    const double2 r23 = *reinterpret_cast<const double2*>(row + rbase + 2);
    const double l[4] = {l01.x, l01.y, l23.x, l23.y};
    const double r[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) part[4 * a + b] += l[a] * r[b];
  }
  // partial sums of this lane's block over one tile.  `want` = row tag to accept (-1: every row); `pure` = every
  // row of the tile carries the tag `tag0` (the usual case: a 32-row block lies inside one contour segment).
  __device__ __forceinline__ void mac_tile(const double* tile, const int* tags, int rows, int want, bool pure, int tag0,
                                           double (&part)[16]) const {
#pragma unroll
    for (int e = 0; e < 16; ++e) part[e] = 0.0;
    if (!SP || want < 0 || pure) {
      if (SP && want >= 0 && tag0 != want) return;
      int k = 0;
      for (; k + 2 <= rows; k += 2) {
        mac_row(tile + k * B::LDT, part);
        mac_row(tile + (k + 1) * B::LDT, part);
      }
      if (k < rows) mac_row(tile + k * B::LDT, part);
    } else {
      for (int k = 0; k < rows; ++k)
        if (tags[k] == want) mac_row(tile + k * B::LDT, part);
    }
  }
  // output index of operand slot (group base, offset) for curve c: -1 = padding
  __device__ __forceinline__ int out_index(int base, int off, int curve) const {
    if (SP) {
      if (base == 8) return off < 3 ? 16 + off : -1;
      return 8 * curve + base + off;
    }
    const int v = base + off;
    return v < M ? v : -1;
  }
};

// item list of the compact Stereo19 path: one item per (32-row block, curve present in it), ascending blocks;
// bit 8 = curve, bit 9 = primary (the item that also carries the block's theta/phi/z-only entries and J^T e).
// Built once per fit by lane 0.
__device__ __forceinline__ int build_items19(const FitGeom& g, int* items, int lane) {
  int n_items = 0;
  if (lane == 0) {
    const int nblocks = (g.n + kTileRows - 1) / kTileRows;
    for (int b = 0; b < nblocks; ++b) {
      const int s0 = b * (kTileRows / 2), s1 = min(s0 + kTileRows / 2, g.nsamp) - 1;
      const int seg0 = (s0 >= g.b1) + (s0 >= g.b2) + (s0 >= g.b3);
      const int seg1 = (s1 >= g.b1) + (s1 >= g.b2) + (s1 >= g.b3);
      const int c0 = seg0 & 1;
      items[n_items++] = b | (c0 << 8) | (1 << 9);
      if (seg1 != seg0) items[n_items++] = b | ((c0 ^ 1) << 8);
    }
  }
  return __shfl_sync(0xffffffffu, n_items, 0);
}

// J^T J and J^T e by the blocked scheme.  STORED: rows come from the column-major secant Jacobian in global
// memory (dif); otherwise they are evaluated from the analytic Jacobian (der), compact when SP.
template <class MD, bool SP, bool STORED>
__device__ __forceinline__ void normal_eqs_blocked(const WarpMem<MD::M>& w, const FitGeom& g,
                                                   BlockAccum<MD::M, SP>& A, double* J, int ld, int n_items,
                                                   int lane, bool pending = false, double pend_dp_sq = 1.0) {
  constexpr int M = MD::M;
  using B = Blk<M, SP>;
  double* tiles = w.tile;
  int* tags = reinterpret_cast<int*>(w.tile + B::NG * B::TILE_DOUBLES);
  int* items_rw = reinterpret_cast<int*>(w.tile + B::NG * B::TILE_DOUBLES + B::NG * kTileRows / 2);
  const int* items = items_rw;
  int* pure_flags = items_rw + 48;  // the item list holds at most 32 + 3 entries
  double* etile = w.tile + B::NG * B::TILE_DOUBLES + B::NG * kTileRows / 2 + 32;  // [NG][32] residuals, contiguous
  (void)pure_flags;
  (void)etile;
  if (!STORED) {
    MD::prep_jac(w.p, w.scratch, lane);
    __syncwarp();
  }
  A.jte = 0.0;
  for (int i = lane; i < M * M; i += 32) w.jtj[i] = 0.0;  // accumulators of the owners live in the jtj array
  const int nblocks = (g.n + kTileRows - 1) / kTileRows;
  const int total = SP ? n_items : nblocks;
  for (int step = 0; step * B::NG < total; ++step) {
    // ---- produce up to NG tiles, lane = row.  The t and e of all NG rows are requested first so that the
    // L2 latencies overlap.
    double t_pre[B::NG], e_pre[B::NG];
    if constexpr (SP) {
#pragma unroll
      for (int gi = 0; gi < B::NG; ++gi) {
        const int it = step * B::NG + gi;
        t_pre[gi] = 0.0;
        e_pre[gi] = 0.0;
        if (it < total) {
          const int r = (items[it] & 0xff) * kTileRows + lane;
          if (r < g.n) { t_pre[gi] = w.t[r >> 1]; e_pre[gi] = w.e[r]; }
        }
      }
    }
#pragma unroll
    for (int gi = 0; gi < B::NG; ++gi) {
      const int it = step * B::NG + gi;
      if (it < total) {
        const int blk = SP ? (items[it] & 0xff) : it;
        const int kk = blk * kTileRows;
        const int r = kk + lane;
        double* row = tiles + gi * B::TILE_DOUBLES + lane * B::LDT;
        int tag = -1;
        if (r < g.n) {
          if constexpr (STORED) {
            double jr[M];
#pragma unroll
            for (int j = 0; j < M; ++j) jr[j] = __ldcg(J + j * ld + r);  // streamed: keep L1 for t, x, f(p)
            if (pending) {
              // the deferred rank-one secant update of this row (lm_core.c:745-755), applied on its way into the
              // tile: tmp = (f(p+dp)_r - f(p)_r - J_r.dp)/||dp||^2, J_r += tmp*dp, written back to the stored Jacobian
              double tmp = 0.0;
#pragma unroll
              for (int l = 0; l < M; ++l) tmp += jr[l] * w.dp[l];
              tmp = (w.dbuf[r] - tmp) / pend_dp_sq;
#pragma unroll
              for (int j = 0; j < M; ++j) {
                jr[j] += tmp * w.dp[j];
                __stcg(J + j * ld + r, jr[j]);
              }
            }
#pragma unroll
            for (int j = 0; j < M; ++j) row[j] = jr[j];
          } else if constexpr (SP) {
            double v[12];
            tag = MD::jac_row_compact(w.scratch, g, r, t_pre[gi], v);
#pragma unroll
            for (int q = 0; q < 12; q += 2)  // 16-byte stores: 2-way instead of 8-way bank conflicts at this pitch
              *reinterpret_cast<double2*>(row + q) = make_double2(v[q], v[q + 1]);
            etile[gi * kTileRows + lane] = e_pre[gi];  // residual of this row, for the J^T e chain
            tags[gi * kTileRows + lane] = tag;
          } else {
            MD::jac_row(w.scratch, g, r, w.t[r >> 1], row);
          }
          if constexpr (!SP) {
#pragma unroll
            for (int j = M; j < B::TC; ++j) row[j] = 0.0;
            etile[gi * kTileRows + lane] = w.e[r];
          }
        }
        if constexpr (SP) {  // "pure" tile: every row belongs to the same curve (the usual case)
          const int t0 = __shfl_sync(0xffffffffu, tag, 0);
          const unsigned same = __ballot_sync(0xffffffffu, tag < 0 || tag == t0);
          if (lane == 0) pure_flags[gi] = (same == 0xffffffffu) ? 1 : 0;
        }
      }
    }
    __syncwarp();
    // ---- every lane: partial sums of its 4x4 block over its group's tile
    double part[16];
    {
      const int it = step * B::NG + A.grp;
      int rows = 0, want = -1;
      if (A.active && it < total) {
        const int item = SP ? items[it] : it;
        const int blk = item & 0xff;
        rows = min(kTileRows, g.n - blk * kTileRows);
        if (SP) {
          want = A.is_c_role ? -1 : ((item >> 8) & 1);
          if (A.is_c_role && !((item >> 9) & 1)) rows = 0;
        }
      }
      bool pure = true;
      int tag0 = 0;
      if constexpr (SP) {
        tag0 = tags[A.grp * kTileRows];
        pure = pure_flags[A.grp] != 0;
      }
      A.mac_tile(tiles + A.grp * B::TILE_DOUBLES, tags + A.grp * kTileRows, rows, want, pure, tag0, part);
    }
    // ---- fold the partial sums into the owners in ascending block order.  An owner keeps its 16 running
    // entries in the jtj array: one load, up to NG ordered additions, one store per entry and step.
    {
      const int ngrp = min(B::NG, total - step * B::NG);
      // which of this step's tiles feed this lane's entries (bit gi), and the curve they belong to
      unsigned feed = 0;
      if (A.active) {
        for (int gi = 0; gi < ngrp; ++gi) {
          bool mine;
          if (SP) {
            const int item = items[step * B::NG + gi];
            mine = A.is_c_role ? (A.grp == 0 && ((item >> 9) & 1)) : (A.grp == ((item >> 8) & 1));
          } else {
            mine = A.grp == 0;
          }
          feed |= mine ? (1u << gi) : 0u;
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const bool keep = feed != 0u && a < A.na && b < A.nb && (!A.diag || b <= a);
          const int idx = A.out0 + a * M + b;
          double acc = keep ? w.jtj[idx] : 0.0;
#pragma unroll 1
          for (int gi = 0; gi < ngrp; ++gi) {
            const double v = shfl_d(part[4 * a + b], gi * B::R + A.role);
            if ((feed >> gi) & 1u) acc += v;
          }
          if (keep) w.jtj[idx] = acc;
        }
    }
    // ---- J^T e: one running sum per column over ascending rows (lanes = columns)
#pragma unroll 1
    for (int gi = 0; gi < B::NG; ++gi) {
      const int it = step * B::NG + gi;
      if (it < total) {
        const int item = SP ? items[it] : it;
        if (!SP || ((item >> 9) & 1)) {
          const int blk = item & 0xff;
          const int kk = blk * kTileRows;
          const int rows = min(kTileRows, g.n - kk);
          if (lane < M) {
            const double* tile = tiles + gi * B::TILE_DOUBLES;
            const int* tg = tags + gi * kTileRows;
            const double* ev = etile + gi * kTileRows;
            const int col = SP ? (lane < 16 ? (lane & 7) : lane - 8) : lane;
            const int want = (SP && lane < 16) ? (lane >> 3) : -1;
            const bool pure = !SP || pure_flags[gi] != 0;
            // a pure tile either belongs to this column's curve entirely or not at all; a mixed tile (a block
            // that straddles two contour segments) selects row by row: a row of the other curve contributes +0.0
            // (its true entry is a structural zero; adding +-0 never changes a running sum that started at +0)
            if (pure) {
              if (want < 0 || tg[0] == want) {
                int k = 0;
                for (; k + 4 <= rows; k += 4) {
                  const double t0 = tile[(k + 0) * B::LDT + col], t1 = tile[(k + 1) * B::LDT + col];
                  const double t2 = tile[(k + 2) * B::LDT + col], t3 = tile[(k + 3) * B::LDT + col];
                  const double2 e01 = *reinterpret_cast<const double2*>(ev + k);
                  const double2 e23 = *reinterpret_cast<const double2*>(ev + k + 2);
                  const double p0 = t0 * e01.x, p1 = t1 * e01.y, p2 = t2 * e23.x, p3 = t3 * e23.y;
                  A.jte += p0;
                  A.jte += p1;
                  A.jte += p2;
                  A.jte += p3;
                }
                for (; k < rows; ++k) A.jte += tile[k * B::LDT + col] * ev[k];
              }
            } else {
              for (int k = 0; k < rows; ++k) {
                if (want >= 0 && tg[k] != want) continue;
                A.jte += tile[k * B::LDT + col] * ev[k];
              }
            }
          }
        }
      }
    }
    __syncwarp();
  }
  // mirror the lower triangle (the reference mirrors its upper triangle, misc_core.c:128-131) and publish J^T e
  __syncwarp();
  for (int idx = lane; idx < M * M; idx += 32) {
    const int i = idx / M, j = idx - i * M;
    if (j > i) w.jtj[idx] = w.jtj[j * M + i];
  }
  if (lane < M) w.jte[lane] = A.jte;
  __syncwarp();
}

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
    // sequential semantics: the last row whose merit equals the running maximum (">="); NaN never wins.
    // A non-negative double orders like its bit pattern, so the warp maximum is two hardware integer
    // reductions (high word, then low word among the lanes that hold the high-word maximum).
    int besti = -1;
    {
      const bool valid = merit >= 0.0;
      const unsigned long long bits = valid ? (unsigned long long)__double_as_longlong(merit) : 0ull;
      const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
      const unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
      const unsigned lmax = __reduce_max_sync(0xffffffffu, (hi == hmax) ? lo : 0u);
      const unsigned ballot = __ballot_sync(0xffffffffu, valid && hi == hmax && lo == lmax);
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
      // multipliers l_i = a_ij * (1/pivot) (one row per lane), then the trailing update a_ik -= l_i * u_k spread
      // over a 4 x 8 grid of lanes: every element still receives exactly this one subtraction at step j
      const double inv = 1.0 / a[j * LD + j];
      if (lane > j && lane < M) a[lane * LD + j] = a[lane * LD + j] * inv;
      __syncwarp();
      const int li = lane >> 3, lk = lane & 7;
      for (int k = j + 1 + lk; k < M; k += 8) {
        const double u = a[j * LD + k];
        for (int i = j + 1 + li; i < M; i += 4) a[i * LD + k] -= a[i * LD + j] * u;
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
  // back substitution: row i subtracts a_ij*x_j for j = i+1..M-1 in ascending j, then divides.  As soon as
  // x_j is known every row above it turns its a_ij into the product a_ij*x_j (all lanes at once), so the
  // serial chain of row i is plain subtractions of ready values.
  for (int i = M - 1; i >= 0; --i) {
    double xi = 0.0;
    if (lane == i) {
      double sum = b;
      int jj = i + 1;
      for (; jj + 2 <= M; jj += 2) {
        const double p0 = a[i * LD + jj], p1 = a[i * LD + jj + 1];
        sum -= p0;
        sum -= p1;
      }
      if (jj < M) sum -= a[i * LD + jj];
      xi = sum / a[i * LD + i];
      w.xs[i] = xi;
    }
    xi = shfl_d(xi, i);
    if (lane < i) a[lane * LD + i] *= xi;
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
__device__ void lm_der(const WarpMem<MD::M>& w, const FitGeom& g, const LmLaunch& L, Accum<MD::M>& A,
                       BlockAccum<MD::M, MD::M == 19>& BA, int n_items, LmResult& R, int lane,
                       const LmResume* rs = nullptr, const double* rs_jtj = nullptr,
                       const double* rs_jte = nullptr) {
  constexpr int M = MD::M;
  const bool blocked = !(g.n * M < kBlockSzSq);  // strict "<" selects the unblocked loop, lm_core.c:193
  double mu = 0.0, g_inf = 0.0, p_sq = 0.0, dp_sq = DBL_MAX;
  int nu = 2, stop = 0, nfev = 0, njev = 0, nlss = 0, k;
  double e_sq, e0;
  // A fit taken over in mid-iteration (the staged shape's last fits) is about to solve: it brings J^T J, J^T e, the
  // gradient statistics, mu, nu and its counters, and enters the damping loop directly.  Its residual vector is not
  // needed before an accepted step replaces it.
  bool resumed = rs != nullptr;
  int k0 = 0;
  if (resumed) {
    e0 = rs->e0; e_sq = rs->e_sq; mu = rs->mu; nu = rs->nu; k0 = rs->k; nfev = rs->nfev; njev = rs->njev;
    nlss = rs->nlss; dp_sq = rs->dp_sq; p_sq = rs->p_sq; g_inf = rs->g_inf;
    for (int idx = lane; idx < M * M; idx += 32) {
      const int i = idx / M, j = idx - i * M;
      const int hi = i > j ? i : j, lo = i > j ? j : i;
      w.jtj[idx] = rs_jtj[hi * (hi + 1) / 2 + lo];
    }
    if (lane < M) {
      w.jte[lane] = rs_jte[lane];
      w.diag[lane] = rs_jtj[lane * (lane + 1) / 2 + lane];
    }
    __syncwarp();
  } else {
    eval_func<MD>(w, g, w.p, w.hx, lane);
    nfev = 1;
    e_sq = residual_sqnorm(w.hx, w.x, w.hx, g.n, lane);  // in shared memory, then published to e
    for (int i = lane; i < g.n; i += 32) w.e[i] = w.hx[i];
    __syncwarp();
    e0 = e_sq;
    if (!isfinite(e_sq)) stop = 7;
  }

  for (k = k0; k < L.itmax && !stop; ++k) {
    if (!resumed) {
      if (e_sq <= L.eps3) { stop = 6; break; }
      if (blocked) normal_eqs_blocked<MD, MD::M == 19, false>(w, g, BA, nullptr, 0, n_items, lane);
      else normal_eqs_analytic<MD>(w, g, A, false, lane);
      ++njev;
      grad_stats<M>(w, p_sq, g_inf, lane);
      if (g_inf <= L.eps1) { dp_sq = 0.0; stop = 1; break; }
      if (k == 0) mu = L.tau * max_diag<M>(w);
    }
    resumed = false;

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
__device__ void lm_dif(const WarpMem<MD::M>& w, const FitGeom& g, const LmLaunch& L, Accum<MD::M>& A,
                       BlockAccum<MD::M, false>& BA, double* J, int ld, LmResult& R, int lane) {
  constexpr int M = MD::M;
  const bool blocked = !(g.n * M <= kBlockSzSq);  // "<=" here, lm_core.c:585
  const int K = (M >= 10) ? M : 10;
  double mu = 0.0, g_inf = 0.0, p_sq = 0.0, dp_sq = DBL_MAX;
  int nu, stop = 0, nfev = 0, njap = 0, nlss = 0, k, n_jtj = 0, n_broyden = 0;
  int updjac = 0, updp = 1, newjac = 0;
  bool pending = false;  // a secant update whose application is deferred to the next J^T J build

  eval_func<MD>(w, g, w.p, w.hx, lane);
  nfev = 1;
  double e_sq = residual_sqnorm(w.wrk2, w.x, w.hx, g.n, lane);  // in shared memory, then published to e
  for (int i = lane; i < g.n; i += 32) w.e[i] = w.wrk2[i];
  __syncwarp();
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
          for (int i = lane; i < g.n; i += 32) __stcg(J + j * ld + i, (w.wrk[i] - w.hx[i]) * inv);
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
      pending = false;  // the fresh Jacobian overwrites every entry a deferred secant update would have touched
    }

    if (newjac) {
      newjac = 0;
      if (blocked) normal_eqs_blocked<MD, false, true>(w, g, BA, J, ld, 0, lane, pending, dp_sq);
      else normal_eqs_stored<M>(w, g, J, ld, A, false, lane);
      pending = false;
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
      eval_trial<MD>(w, g, w.pdp, w.wrk, w.wrk2, w.dbuf, w.hx, lane);
      ++nfev;
      const double e_new = residual_chains(w.wrk2, g.n, lane);
      if (!isfinite(e_new)) { stop = 7; break; }
      const double dF = e_sq - e_new;
      if (updp || dF > 0.0) {
        if (blocked) {
          // rank-one secant update, lm_core.c:745-755, DEFERRED: nothing reads J before the next J^T J build, which
          // streams every row through registers anyway, so the update is applied there (one pass over the 78 KB
          // Jacobian instead of two).  It needs f(p+dp) - f(p) of this iteration, saved now because an accepted step
          // overwrites f(p); dp and ||dp||^2 stay untouched until the next solve, which comes after the build.
          pending = true;  // f(p+dp) - f(p) was saved by eval_trial
        } else {
        // rank-one secant update, lm_core.c:745-755: one Jacobian row per lane, two rows in flight per trip.  The row
          // is read once into registers (all M loads of both rows are issued before anything waits on them: the
          // Jacobian lives in L2) and written back updated.
          for (int i0 = lane; i0 < g.n; i0 += 64) {
            const int i1 = i0 + 32;
            const bool two = i1 < g.n;
            double r0[M], r1[M];
#pragma unroll
            for (int l = 0; l < M; ++l) {
              r0[l] = J[l * ld + i0];
              r1[l] = two ? J[l * ld + i1] : 0.0;
            }
            const double d0 = w.wrk[i0] - w.hx[i0];
            const double d1 = two ? (w.wrk[i1] - w.hx[i1]) : 0.0;
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int l = 0; l < M; ++l) {
              const double dpl = w.dp[l];
              t0 += r0[l] * dpl;
              t1 += r1[l] * dpl;
            }
            t0 = (d0 - t0) / dp_sq;
            t1 = (d1 - t1) / dp_sq;
#pragma unroll
            for (int j = 0; j < M; ++j) {
              const double dpj = w.dp[j];
              J[j * ld + i0] = r0[j] + t0 * dpj;
              if (two) J[j * ld + i1] = r1[j] + t1 * dpj;
            }
          }
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

// ---- fit_curve's initial guess (curveFittingClass.cpp:516-651), one thread per fit ------------------------------
// Stereo triangulation of the start / end points of both curves, plane fit through them (normal = eigenvector of
// the smallest eigenvalue of the centred 3x3 scatter matrix, the reference's cvSVD of the 4x3 matrix), rotation into
// the ground frame, straight-line control points.  The sign of the reference's singular vector is
// implementation-defined; here the normal is oriented so that d = n.c >= 0 (height parameter <= 0, the branch the
// reference's own post-normalisation :684-702 maps every fit to).  Same operation sequence as the CPU oracle.
__device__ __forceinline__ void triangulate(const double* m, int a, int b, double pt[3]) {
  const double uL = m[2 * a], vL = m[2 * a + 1], uR = m[2 * b], vR = m[2 * b + 1];
  pt[0] = CPS_FX * CPS_BASELINE / (uL - uR);
  pt[1] = 0.5 * (pt[0] / CPS_FX * (uL - CPS_CX) - 0.5 * CPS_BASELINE) + 0.5 * (pt[0] / CPS_FX * (uR - CPS_CX) + 0.5 * CPS_BASELINE);
  pt[2] = 0.5 * pt[0] / CPS_FY * (vL - CPS_CY) + 0.5 * pt[0] / CPS_FY * (vR - CPS_CY);
}

__global__ void __launch_bounds__(128) lm_init_guess_kernel(int B, const double* __restrict__ meas,
                                                            const long long* __restrict__ off,
                                                            const int* __restrict__ counts, double* __restrict__ p_out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  const double* m = meas + off[f];
  const int nl1 = counts[4 * f], nl2 = counts[4 * f + 1], nr1 = counts[4 * f + 2], nr2 = counts[4 * f + 3];
  const int npl = nl1 + nl2, npts = npl + nr1 + nr2;
  double* p = p_out + (size_t)f * 19;
  if (nl1 < 1 || nl2 < 1 || nr1 < 1 || nr2 < 1 || (long long)2 * npts != off[f + 1] - off[f]) {
    for (int i = 0; i < 19; ++i) p[i] = __longlong_as_double(0x7ff8000000000000LL);  // no guess: NaN
    return;
  }
  double s1[3], e1[3], s2[3], e2[3], c[3], M[4][3], a[3][3], V[3][3], n[3];
  triangulate(m, 0, npl, s1);
  triangulate(m, nl1 - 1, npl + nr1 - 1, e1);
  triangulate(m, nl1, npl + nr1, s2);
  triangulate(m, npl - 1, npts - 1, e2);
  for (int j = 0; j < 3; ++j) c[j] = 0.25 * (s1[j] + s2[j] + e1[j] + e2[j]);
  for (int j = 0; j < 3; ++j) {
    M[0][j] = s1[j] - c[j]; M[1][j] = s2[j] - c[j]; M[2][j] = e1[j] - c[j]; M[3][j] = e2[j] - c[j];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 4; ++k) acc += M[k][i] * M[k][j];
      a[i][j] = acc;
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 8; ++sweep)
    for (int pp = 0; pp < 2; ++pp)
      for (int q = pp + 1; q < 3; ++q) {
        const double apq = a[pp][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[pp][pp]) / (2.0 * apq);
        double t = 1.0 / (((theta >= 0.0) ? theta : -theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        const double cc = 1.0 / sqrt(t * t + 1.0);
        const double ss = t * cc;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][pp], akq = a[k][q];
          a[k][pp] = cc * akp - ss * akq;
          a[k][q] = ss * akp + cc * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[pp][k], aqk = a[q][k];
          a[pp][k] = cc * apk - ss * aqk;
          a[q][k] = ss * apk + cc * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][pp], vkq = V[k][q];
          V[k][pp] = cc * vkp - ss * vkq;
          V[k][q] = ss * vkp + cc * vkq;
        }
      }
  int best = 0;
  if (a[1][1] < a[best][best]) best = 1;
  if (a[2][2] < a[best][best]) best = 2;
  for (int k = 0; k < 3; ++k) n[k] = V[k][best];
  double d = (n[0] * c[0] + n[1] * c[1]) + n[2] * c[2];
  if (d < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; d = -d; }
  const double theta = -det_asin(n[0]);
  const double phi = det_asin(n[1]);
  double origin[3], st, ct, sp, cp, Rot[9];
  for (int j = 0; j < 3; ++j) origin[j] = d * n[j];
  det_sincos(theta, &st, &ct);
  det_sincos(phi, &sp, &cp);
  Rot[0] = ct;  Rot[1] = st * sp; Rot[2] = st * cp;
  Rot[3] = 0.0; Rot[4] = cp;      Rot[5] = -sp;
  Rot[6] = -st; Rot[7] = ct * sp; Rot[8] = ct * cp;
  double* pts[4] = {s1, e1, s2, e2};
  for (int i = 0; i < 4; ++i) {
    double t[3], r[3];
    for (int j = 0; j < 3; ++j) t[j] = pts[i][j] - origin[j];
    for (int j = 0; j < 3; ++j) r[j] = (Rot[3 * j] * t[0] + Rot[3 * j + 1] * t[1]) + Rot[3 * j + 2] * t[2];
    for (int j = 0; j < 3; ++j) pts[i][j] = r[j];
  }
  double cp1[8], cp2[8];
  cp1[0] = s1[0]; cp1[1] = s1[1]; cp1[6] = e1[0]; cp1[7] = e1[1];
  cp2[0] = s2[0]; cp2[1] = s2[1]; cp2[6] = e2[0]; cp2[7] = e2[1];
  for (int i = 1; i < 3; ++i) {
    cp1[2 * i] = (3 - i) * cp1[0] / 3 + i * cp1[6] / 3;
    cp1[2 * i + 1] = (3 - i) * cp1[1] / 3 + i * cp1[7] / 3;
    cp2[2 * i] = (3 - i) * cp2[0] / 3 + i * cp2[6] / 3;
    cp2[2 * i + 1] = (3 - i) * cp2[1] / 3 + i * cp2[7] / 3;
  }
  for (int i = 0; i < 8; ++i) { p[i] = cp1[i]; p[8 + i] = cp2[i]; }
  p[16] = theta;
  p[17] = phi;
  p[18] = -d;
}

// ---- kernel: persistent warps pulling fits from a queue ----------------------------------------------
template <int MODEL, bool DIF>
__global__ void __launch_bounds__(kWarpsPerCta * 32, CPS_LM_MIN_BLOCKS) lm_fit_kernel(const LmLaunch L) {
  using MD = Model<MODEL>;
  constexpr int M = MD::M;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int per_warp = lm_warp_smem_doubles(M, MD::SCRATCH, L.n_max, DIF);
  double* base = smem + (size_t)wid * per_warp;

  WarpMem<M> w;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + wid;
  w.x = nullptr;
  double* ws = L.jstore + (size_t)gwarp * lm_ws_doubles(M, L.n_max, DIF);
  w.e = ws;
  w.t = ws + L.n_max;
  w.tile = base; base += lm_tile_region(M, L.n_max, DIF);
  double* J = nullptr;
  if (DIF) {
    w.hx = ws + L.n_max + L.n_max / 2;  // f(p)
    w.wrk = w.hx + L.n_max;             // f(p + dp), finite-difference evaluations
    w.wrk2 = w.tile;                    // trial residual / second central-difference evaluation
    w.dbuf = w.wrk + L.n_max;           // f(p+dp) - f(p) of a deferred secant update
    J = w.dbuf + L.n_max;
  } else {
    w.hx = w.tile;  // trial residual
    w.wrk = nullptr; w.wrk2 = nullptr; w.dbuf = nullptr;
  }
  w.jtj = base; base += M * M;
  w.p = base; w.dp = base + 32; w.pdp = base + 64; w.jte = base + 96; w.diag = base + 128; w.pw = base + 160;
  w.xs = base + 192; w.scale = base + 224; base += 256;
  w.scratch = base;



  constexpr bool SP = !DIF && M == 19;
  Accum<M> A;
  A.init(lane);
  BlockAccum<M, SP> BA;
  BA.init(lane);

  for (;;) {
    unsigned int f = 0;
    if (lane == 0) f = atomicAdd(L.queue, 1u);
    f = __shfl_sync(0xffffffffu, f, 0);
    if (f >= (unsigned)L.B) break;
    if (L.fit_list) f = (unsigned int)L.fit_list[f];

    const long long o0 = L.off[f], o1 = L.off[f + 1];
    const long long nn = o1 - o0;
    FitGeom g;
    g.n = (int)nn;
    g.nsamp = g.n >> 1;
    int ret_code = 0;
    if (nn < M) ret_code = -1;               // lm_core.c:117 / :493: fewer measurements than unknowns
    else if (nn > L.n_max || (nn & 1)) ret_code = -2;  // does not fit the staged buffers / odd length
    else if (!DIF && M == 19 && nn > 40 * kTileRows) ret_code = -2;
    else if (DIF && (o0 & 1)) ret_code = -3;  // the secant path reads (x,y) sample pairs as 16-byte words  // the compact path lists at most 40+ row blocks
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
    w.x = L.meas + o0;
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
    if constexpr (DIF) {
      lm_dif<MD>(w, g, L, A, BA, J, L.n_max, R, lane);
    } else {
      int n_items = 0;
      if constexpr (SP) {
        int* items = reinterpret_cast<int*>(w.tile + Blk<M, SP>::NG * Blk<M, SP>::TILE_DOUBLES +
                                            Blk<M, SP>::NG * kTileRows / 2);
        n_items = build_items19(g, items, lane);
        __syncwarp();
      }
      const bool takeover = L.fit_list && L.resume;
      lm_der<MD>(w, g, L, A, BA, n_items, R, lane, takeover ? L.resume + f : nullptr,
                 takeover ? L.resume_jtj + (size_t)f * (M * (M + 1) / 2) : nullptr,
                 takeover ? L.resume_jte + (size_t)f * M : nullptr);
    }

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
