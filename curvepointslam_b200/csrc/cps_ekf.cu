// cps_ekf.cu -- EKF predict / measurement update on a device-resident dense covariance (sm_100a, FP64).
//
// Computes what KalmanFilter::UpdateNCurvesAndPoints (kalmanClass.cpp:669-945), UpdatePoints (:948-1126)
// and PredictKF (:161-272) compute, re-designed around the fact that the update is bandwidth bound:
//   * H has at most 6 + 8n + 3n_pts non-zero columns, so P H^T touches N x (that many) entries of P, not
//     N^2 (the reference multiplies the dense M x N H three times);
//   * S^-1 is an M x M problem solved by one thread block (Cholesky; the reference's cvInvert(CV_SVD) of a
//     well-conditioned SPD matrix agrees to cond(S)*eps);
//   * P <- 1/2[(P - K S K^T) + (P - K S K^T)^T] is ONE streaming pass: every thread block owns a pair of
//     mirrored 64x64 tiles (I,J)/(J,I), reads each P entry once, writes each once (16 N^2 bytes, the
//     minimum for a dense in/out covariance), and forms the rank-M correction from K and K S^T tiles that
//     stay in L2/shared memory.  The correction K_I (S K_J^T) is computed once per pair and used for both
//     mirrored tiles: K S K^T is symmetric in exact arithmetic and the reference averages the two roundings
//     anyway (cvAddWeighted(P,0.5,P^T,0.5)), so the result differs by O(eps) per entry, inside the 1e-9
//     tolerance of the parity tests, and is exactly symmetric like the reference's.
// The reference allocates 14 temporaries (three of them N x N) per call; here nothing is allocated per call.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/cps_ekf.h"
#include "cps_common.h"
#include "cps_ekf_internal.h"

namespace {

constexpr int MAXM = CPS_EKF_MAX_MEAS;
constexpr int MAXNC = 6 + MAXM;  // compact columns of H: 6 pose columns + landmark columns
constexpr int TILE = 64;
constexpr int kFrontSmemMax = 225 * 1024;  // dynamic shared memory of the fused front end (M = 65 needs 209 KB)
constexpr double kPI = 3.1415926535;  // common.h:114

// P is symmetric to the last bit whenever the streaming covariance kernel runs (mirror_exact), and that kernel may
// leave the tiles BELOW the diagonal tiles unwritten (cps_ekf::lower_stale): every kernel that gathers single
// entries of P reads the copy with row <= column, which is always current (and equal to its mirror image when
// both are).
__device__ __forceinline__ size_t upper_index(int r, int c, int ld, int sym) {
  return (r <= c || !sym) ? (size_t)r * ld + c : (size_t)c * ld + r;
}

// constants of kalmanClass.h:17-36
constexpr double MEAS_COV = 0.5, PT_MEAS_COV = 5.0, PHI_MEAS_COV = 0.1, THETA_MEAS_COV = 0.1, Z_MEAS_COV = 0.2;
constexpr double VX_COV = 0.01, VY_COV = 0.01, VZ_COV = 0.5, WX_COV = 0.01, WY_COV = 0.01, WZ_COV = 0.01;
constexpr double DT = 0.2;

// layout of d_small (doubles)
constexpr int SM_HC = 0;                       // [MAXM][MAXNC] compact H
constexpr int SM_S = SM_HC + MAXM * MAXNC;     // [MAXM][MAXM]
constexpr int SM_SINV = SM_S + MAXM * MAXM;    // [MAXM][MAXM]
constexpr int SM_NU = SM_SINV + MAXM * MAXM;   // [MAXM] innovation z - zhat
constexpr int SM_R = SM_NU + MAXM;             // [MAXM] diagonal of R
constexpr int SM_IN = SM_R + MAXM;             // staged inputs: z[MAXM] | A[16*MAXM/8] 
constexpr int SM_TOTAL = SM_IN + MAXM + 16 * (MAXM / 8 + 1);

struct UpdateDesc {
  int N, ld;      // state dimension, leading dimension of P
  int n, n_pts;   // curves and points measured
  int M, nc;      // measurement size, compact columns
  int sym;        // P is symmetric to the last bit (cps_ekf::mirror_exact): single entries are read from the upper triangle
};

// generate_Rbe (common.h:174-186), row-major, with the reference's sign in entry [6]
__device__ void rbe(double phi, double theta, double psi, double* R) {
  R[0] = cos(psi) * cos(theta);
  R[3] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
  R[6] = cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi);
  R[1] = sin(psi) * cos(theta);
  R[4] = sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi);
  R[7] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
  R[2] = -sin(theta);
  R[5] = cos(theta) * sin(phi);
  R[8] = cos(theta) * cos(phi);
}

// get_Rbe_derivs (kalmanClass.cpp:1275-1298) including its typo: R_be_theta(1,1) is written twice and
// (2,1) is never set
__device__ void rbe_derivs(double phi, double theta, double psi, double* Rp, double* Rt, double* Rs) {
  for (int i = 0; i < 9; ++i) { Rp[i] = 0.0; Rt[i] = 0.0; Rs[i] = 0.0; }
  Rp[3] = cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi);
  Rp[6] = -cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
  Rp[4] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
  Rp[7] = -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi);
  Rp[5] = cos(theta) * cos(phi);
  Rp[8] = -cos(theta) * sin(phi);
  Rt[0] = -cos(psi) * sin(theta);
  Rt[3] = cos(psi) * cos(theta) * sin(phi);
  Rt[6] = cos(psi) * cos(theta) * cos(phi);
  Rt[1] = -sin(psi) * sin(theta);
  Rt[4] = sin(psi) * cos(theta) * cos(phi);  // second write wins
  Rt[2] = -cos(theta);
  Rt[5] = -sin(theta) * sin(phi);
  Rt[8] = -sin(theta) * cos(phi);
  Rs[0] = -sin(psi) * cos(theta);
  Rs[3] = -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi);
  Rs[6] = -sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
  Rs[1] = cos(psi) * cos(theta);
  Rs[4] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
  Rs[7] = cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi);
}

// ---- kernel 1: compact H, innovation and R from the current state (one block) ----------------------------
// Compact columns: 0..5 = pose columns 0..5; then 8 per measured curve; then 3 per measured point.
// in: small[SM_IN..] = z (M) followed by the A matrices (16 n); cols_in = landmark first-state indices.
__device__ __forceinline__ void ekf_build_device(const UpdateDesc& d, const double* __restrict__ x,
                                                 double* __restrict__ small, int* __restrict__ cols,
                                                 const int* __restrict__ land_cols) {
  const int tid = threadIdx.x, nt = blockDim.x;
  double* Hc = small + SM_HC;
  double* nu = small + SM_NU;
  double* Rd = small + SM_R;
  const double* z = small + SM_IN;
  const double* A = small + SM_IN + MAXM;
  const int cms = d.n > 0 || d.n_pts >= 0 ? 0 : 0;
  (void)cms;
  const bool curves_call = (d.M != 3 * d.n_pts);  // UpdateNCurvesAndPoints has the 3 direct pose rows
  const int curve_meas = curves_call ? 8 * d.n + 3 : 0;
  for (int i = tid; i < d.M * d.nc; i += nt) Hc[i] = 0.0;
  for (int c = tid; c < d.nc; c += nt) {
    int col;
    if (c < 6) col = c;
    else if (c < 6 + 8 * d.n) col = land_cols[(c - 6) / 8] + (c - 6) % 8;
    else col = land_cols[d.n + (c - 6 - 8 * d.n) / 3] + (c - 6 - 8 * d.n) % 3;
    cols[c] = col;
  }
  __syncthreads();
  const double Tx = x[0], Ty = x[1], psi = x[5];
  const double cpsi = cos(psi), spsi = sin(psi);
  // curve rows, kalmanClass.cpp:741-787, and their predicted measurement (GetPredictedMeasurement :1162-1199)
  for (int k = tid; k < d.n; k += nt) {
    const int c0 = land_cols[k];
    const double* Ak = A + 16 * k;
    double ax[4], ay[4];
    for (int i = 0; i < 4; ++i) {
      double sx = 0.0, sy = 0.0;
      for (int j = 0; j < 4; ++j) {
        sx += Ak[4 * i + j] * x[c0 + j];
        sy += Ak[4 * i + j] * x[c0 + 4 + j];
      }
      ax[i] = sx - Tx;
      ay[i] = sy - Ty;
    }
    for (int i = 0; i < 4; ++i) {
      double* hx = Hc + (8 * k + i) * d.nc;      // x rows:  cos*A | sin*A
      double* hy = Hc + (8 * k + 4 + i) * d.nc;  // y rows: -sin*A | cos*A
      for (int j = 0; j < 4; ++j) {
        hx[6 + 8 * k + j] = cpsi * Ak[4 * i + j];
        hx[6 + 8 * k + 4 + j] = spsi * Ak[4 * i + j];
        hy[6 + 8 * k + j] = -spsi * Ak[4 * i + j];
        hy[6 + 8 * k + 4 + j] = cpsi * Ak[4 * i + j];
      }
      hx[0] = -cpsi; hx[1] = -spsi;
      hy[0] = spsi;  hy[1] = -cpsi;
      hx[5] = -spsi * ax[i] + cpsi * ay[i];  // Rotderiv * (A x - T)
      hy[5] = -cpsi * ax[i] - spsi * ay[i];
      nu[8 * k + i] = z[8 * k + i] - (cpsi * ax[i] + spsi * ay[i]);
      nu[8 * k + 4 + i] = z[8 * k + 4 + i] - (-spsi * ax[i] + cpsi * ay[i]);
    }
  }
  if (curves_call && tid == 0) {  // direct rows: height, phi, theta (:790-792) with R of :705-707
    for (int q = 0; q < 3; ++q) {
      Hc[(8 * d.n + q) * d.nc + 2 + q] = 1.0;
      nu[8 * d.n + q] = z[8 * d.n + q] - x[2 + q];
    }
  }
  for (int i = tid; i < d.M; i += nt) {
    double r = MEAS_COV * MEAS_COV;
    if (!curves_call || i >= curve_meas) r = PT_MEAS_COV * PT_MEAS_COV;
    else if (i == 8 * d.n) r = Z_MEAS_COV * Z_MEAS_COV;
    else if (i == 8 * d.n + 1) r = PHI_MEAS_COV * PHI_MEAS_COV;
    else if (i == 8 * d.n + 2) r = THETA_MEAS_COV * THETA_MEAS_COV;
    Rd[i] = r;
  }
  // point rows, kalmanClass.cpp:815-851 / :1016-1050
  if (d.n_pts > 0) {
    double R_be[9], Rd3[3][9];
    rbe(x[3], x[4], x[5], R_be);
    rbe_derivs(x[3], x[4], x[5], Rd3[0], Rd3[1], Rd3[2]);
    for (int k = tid; k < d.n_pts; k += nt) {
      const int c0 = land_cols[d.n + k];
      const int row0 = curve_meas + 3 * k, cc = 6 + 8 * d.n + 3 * k;
      double xe[3];
      for (int i = 0; i < 3; ++i) xe[i] = x[c0 + i] - x[i];
      for (int i = 0; i < 3; ++i) {
        double* h = Hc + (row0 + i) * d.nc;
        double zh = 0.0;
        for (int j = 0; j < 3; ++j) {
          zh += R_be[3 * i + j] * xe[j];
          h[j] = -R_be[3 * i + j];
          h[cc + j] = R_be[3 * i + j];
          double dd = 0.0;
          for (int q = 0; q < 3; ++q) dd += Rd3[j][3 * i + q] * xe[q];
          h[3 + j] = dd;
        }
        nu[row0 + i] = z[row0 + i] - zh;
      }
    }
  }
}

__global__ void ekf_build_kernel(UpdateDesc d, const double* __restrict__ x, double* __restrict__ small,
                                 int* __restrict__ cols, const int* __restrict__ land_cols) {
  ekf_build_device(d, x, small, cols, land_cols);
}

// ---- kernel 2: T = P H^T through the compact columns (one warp per state row) -----------------------------
__global__ void __launch_bounds__(256) ekf_pht_kernel(UpdateDesc d, const double* __restrict__ P,
                                                      const double* __restrict__ small, const int* __restrict__ cols,
                                                      double* __restrict__ T) {
  extern __shared__ double sh[];
  double* Hc = sh;                        // [M][nc]
  double* prow = sh + d.M * d.nc;         // [8 warps][nc]
  int* scols = (int*)(prow + 8 * d.nc);
  for (int i = threadIdx.x; i < d.M * d.nc; i += blockDim.x) Hc[i] = small[SM_HC + i];
  for (int i = threadIdx.x; i < d.nc; i += blockDim.x) scols[i] = cols[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* pr = prow + w * d.nc;
  for (int row = blockIdx.x * 8 + w; row < d.N; row += gridDim.x * 8) {
    for (int c = lane; c < d.nc; c += 32) pr[c] = P[upper_index(row, scols[c], d.ld, d.sym)];
    __syncwarp();
    for (int r = lane; r < d.M; r += 32) {
      const double* h = Hc + r * d.nc;
      double s = 0.0;
      for (int c = 0; c < d.nc; ++c) s += pr[c] * h[c];
      T[(size_t)row * MAXM + r] = s;
    }
    __syncwarp();
  }
}

// ---- kernel 3: S = H T + R, S^-1 by Cholesky (one block) ---------------------------------------------------
__global__ void __launch_bounds__(256) ekf_s_kernel(UpdateDesc d, const double* __restrict__ T,
                                                    double* __restrict__ small, const int* __restrict__ cols,
                                                    int* __restrict__ flag, int staged) {
  extern __shared__ double sh[];
  const int M = d.M;
  double* L = sh;           // [M][M] Cholesky factor, then L^-1
  double* S = sh + M * M;   // [M][M]
  const int tid = threadIdx.x, nt = blockDim.x;
  // S = Hc Tc + R.  When it fits, the compact H and the nc rows of T it touches are first staged in shared memory
  // (coalesced) so that the M*M dot products do not each walk global memory.
  if (staged) {
    double* Hs = sh + 2 * M * M;  // [M][nc]
    double* Ts = Hs + M * d.nc;   // [nc][M]
    for (int e = tid; e < M * d.nc; e += nt) Hs[e] = small[SM_HC + e];
    for (int e = tid; e < d.nc * M; e += nt) {
      const int c = e / M, sidx = e - c * M;
      Ts[e] = T[(size_t)cols[c] * MAXM + sidx];
    }
    __syncthreads();
    for (int e = tid; e < M * M; e += nt) {
      const int r = e / M, s = e - r * M;
      double acc = 0.0;
      for (int c = 0; c < d.nc; ++c) acc += Hs[r * d.nc + c] * Ts[c * M + s];
      if (r == s) acc += small[SM_R + r];
      S[e] = acc;
    }
  } else {
    const double* Hc = small + SM_HC;
    for (int e = tid; e < M * M; e += nt) {
      const int r = e / M, s = e - r * M;
      double acc = 0.0;
      for (int c = 0; c < d.nc; ++c) {
        const double hh = Hc[r * d.nc + c];
        if (hh != 0.0) acc += hh * T[(size_t)cols[c] * MAXM + s];
      }
      if (r == s) acc += small[SM_R + r];
      S[e] = acc;
    }
  }
  __syncthreads();
  for (int e = tid; e < M * M; e += nt) {
    const int r = e / M, s = e - r * M;
    small[SM_S + r * MAXM + s] = S[e];
    L[e] = (s <= r) ? 0.5 * (S[e] + S[s * M + r]) : 0.0;  // factor the symmetric part
  }
  __syncthreads();
  // right-looking Cholesky on the lower triangle
  // A pivot that is not positive (S singular, indefinite or not finite) fails the update: the flag is raised
  // (reported by the next cps_ekf_sync as CPS_ERR_INVALID), S^-1 is written as zero so that the gain is zero and the
  // kernels that follow leave x and P exactly as they were, and the factorisation carries on with a unit pivot so
  // that no NaN / inf is ever produced.
  __shared__ int bad;
  if (tid == 0) bad = 0;
  __syncthreads();
  for (int k = 0; k < M; ++k) {
    __shared__ double piv;
    if (tid == 0) {
      const double v = L[k * M + k];
      const bool ok = (v > 0.0) && (v <= 1.79769313486231571e308);
      if (!ok) { *flag = 1; bad = 1; }
      piv = ok ? sqrt(v) : 1.0;
      L[k * M + k] = piv;
    }
    __syncthreads();
    for (int i = k + 1 + tid; i < M; i += nt) L[i * M + k] /= piv;
    __syncthreads();
    const int rem = M - k - 1;
    for (int e = tid; e < rem * rem; e += nt) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j <= i) L[i * M + j] -= L[i * M + k] * L[j * M + k];
    }
    __syncthreads();
  }
  // invert L in place column by column (thread per column): X = L^-1, lower triangular
  double* X = S;  // reuse
  for (int c = tid; c < M; c += nt) {
    for (int i = 0; i < M; ++i) {
      if (i < c) { X[i * M + c] = 0.0; continue; }
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= L[i * M + k] * X[k * M + c];
      X[i * M + c] = s / L[i * M + i];
    }
  }
  __syncthreads();
  // S^-1 = X^T X
  for (int e = tid; e < M * M; e += nt) {
    const int r = e / M, s = e - r * M;
    double acc = 0.0;
    for (int k = (r > s ? r : s); k < M; ++k) acc += X[k * M + r] * X[k * M + s];
    small[SM_SINV + r * MAXM + s] = bad ? 0.0 : acc;
  }
}

// ---- kernel 4: K = T S^-1, Yt = K S^T, x += K nu (one warp per state row) ----------------------------------
__global__ void __launch_bounds__(256) ekf_gain_kernel(UpdateDesc d, const double* __restrict__ T,
                                                       const double* __restrict__ small, double* __restrict__ Kt,
                                                       double* __restrict__ Yt, double* __restrict__ x, int ldk) {
  extern __shared__ double sh[];
  const int M = d.M;
  double* Sinv = sh;              // [M][M]
  double* S = sh + M * M;         // [M][M]
  double* nu = S + M * M;         // [M]
  double* rowbuf = nu + M;        // [8 warps][2][M]
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int r = e / M, s = e - r * M;
    Sinv[e] = small[SM_SINV + r * MAXM + s];
    S[e] = small[SM_S + r * MAXM + s];
  }
  for (int e = threadIdx.x; e < M; e += blockDim.x) nu[e] = small[SM_NU + e];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* trow = rowbuf + w * 2 * M;
  double* krow = trow + M;
  for (int row = blockIdx.x * 8 + w; row < d.N; row += gridDim.x * 8) {
    for (int r = lane; r < M; r += 32) trow[r] = T[(size_t)row * MAXM + r];
    __syncwarp();
    for (int s = lane; s < M; s += 32) {
      double acc = 0.0;
      for (int r = 0; r < M; ++r) acc += trow[r] * Sinv[r * M + s];
      krow[s] = acc;
      Kt[(size_t)s * ldk + row] = acc;
    }
    __syncwarp();
    double dx = 0.0;
    for (int r = lane; r < M; r += 32) {
      double acc = 0.0;
      for (int s = 0; s < M; ++s) acc += krow[s] * S[r * M + s];
      Yt[(size_t)r * ldk + row] = acc;
      dx += krow[r] * nu[r];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dx += __shfl_xor_sync(0xffffffffu, dx, o);
    if (lane == 0) {
      double v = x[row] + dx;
      if (row >= 3 && row <= 5 && fabs(v) < 1e6) {  // kalmanClass.cpp:902-913 (bounded: a wild angle cannot spin)
        while (v > kPI) v -= 2 * kPI;
        while (v < -kPI) v += 2 * kPI;
      }
      x[row] = v;
    }
    __syncwarp();
  }
}

// Gauss-Jordan steps of the fused front end with a thread's entries in registers from step to step.  A thread owns
// ONE column j of the M x M matrix and every RG-th row of it (RG = NT / M row groups, at most QMAX rows), so a step
// loads the pivot, the thread's pivot-row value W[k][j] once, and one pivot-column value per owned row (the same
// address for the lanes of a row group: a broadcast) -- the first register version loaded two values per entry and
// spent a sixth of its work on an inactive sixth entry.  A step is branch-free: the three cases of the update are
// selects, so the loads of a step overlap each other and the reciprocal instead of forming one dependent load ->
// multiply -> load -> update -> store chain per entry (21 us of the front end's 43 at M = 35 in the plain loop).
// Same operations on the same values as the plain loop in the kernel.  The matrix ping-pongs between Wa (the input)
// and Wb; after M steps the result is in Wb when M is odd.  Requires ceil(M / (NT / M)) <= QMAX.
template <int QMAX, int NT>
__device__ __forceinline__ void gj_register_steps(double* Wa, double* Wb, int M, int tid, int* flag, int* bad) {
  const int RG = NT / M;
  const int g = tid / M, j = tid - g * M;
  const bool owner = g < RG;
  const int nq = owner ? (M - g + RG - 1) / RG : 0;  // rows g, g + RG, ... below M
  double v[QMAX];
#pragma unroll
  for (int q = 0; q < QMAX; ++q) v[q] = q < nq ? Wa[(g + RG * q) * M + j] : 0.0;
  double* Wsrc = Wa;
  double* Wdst = Wb;
  for (int k = 0; k < M; ++k) {
    const double piv = Wsrc[k * M + k];
    const double rk = owner ? Wsrc[k * M + j] : 0.0;
    double cik[QMAX];
#pragma unroll
    for (int q = 0; q < QMAX; ++q) cik[q] = q < nq ? Wsrc[(g + RG * q) * M + k] : 0.0;
    const bool ok = (piv > 0.0) && (piv <= 1.79769313486231571e308);
    const double ip = ok ? 1.0 / piv : 1.0;
    if (!ok && tid == 0) { *flag = 1; *bad = 1; }
    const double rkj = rk * ip;
    const double in_row = (j == k) ? ip : rkj;
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
      if (q < nq) {
        const int i = g + RG * q;
        const double off_row = v[q] - cik[q] * rkj, in_col = -cik[q] * ip;
        const double nv = (i == k) ? in_row : ((j == k) ? in_col : off_row);
        v[q] = nv;
        Wdst[i * M + j] = nv;
      }
    }
    __syncthreads();
    double* tmp = Wsrc; Wsrc = Wdst; Wdst = tmp;
  }
}

// ---- fused front end (used whenever its tables fit shared memory: every live measurement size) ---------------
// ekf_front_kernel, ONE block: kernel 1, then only the nc rows of T = P H^T that S needs (the rows of the compact
// columns: an nc x nc gather of P), S = Hc Tc + R, and S^-1 by Gauss-Jordan elimination on the symmetric part of S
// (S is SPD: no pivoting; all 256 threads update the matrix in every step -- the Cholesky + triangular inverse of
// kernel 3 kept 35 threads busy and cost 57 us at M = 35).  A pivot that is not positive fails the update as in
// kernel 3: flag raised, S^-1 = 0, x and P stay untouched.
// ekf_rows_kernel, one warp per state row: T_row = P_row[cols] Hc^T, K_row = T_row S^-1, Y_row = K_row S^T,
// x_row += K_row nu -- kernels 2 and 4 in one pass, T never written to memory.
// Same sums in the same order as kernels 1-4 for H, T, S, K, Y and x; S^-1 differs by rounding (another elimination).
// NT threads: 256 for M <= 39 (every live size: the elimination steps are latency-bound and a barrier over eight warps
// is cheaper), 512 above (M = 65: the T and S loops have 3.4x the entries with dot products twice as long).
template <int NT>
__global__ void __launch_bounds__(NT) ekf_front_kernel(UpdateDesc d, const double* __restrict__ x,
                                                        const double* __restrict__ P, double* __restrict__ small,
                                                        int* __restrict__ cols, const int* __restrict__ land_cols,
                                                        int* __restrict__ flag) {
  extern __shared__ double sh[];
  const int M = d.M, nc = d.nc;
  const int tid = threadIdx.x, nt = blockDim.x;
  ekf_build_device(d, x, small, cols, land_cols);
  __syncthreads();
  // Hs rows have an odd pitch: the T loop below reads Hs[r][b] with r across the lanes, and with the natural pitch
  // (nc = 38: 4-way, nc = 68: 8-way bank conflicts) that loop was most of this kernel at M = 65
  const int ncp = nc | 1;
  double* Hs = sh;                 // [M][ncp]
  double* Ps = Hs + M * ncp;       // [nc][nc]  P restricted to the compact columns
  double* Tc = Ps + nc * nc;       // [nc][M]   rows of T at the compact columns
  double* S = Tc + nc * M;         // [M][M]
  double* W = S + M * M;           // [M][M]    elimination work matrix -> S^-1 (ping)
  double* W2 = W + M * M;          // [M][M]    (pong)
  __shared__ int bad;
  if (tid == 0) bad = 0;
  // (row, column) of entry e = tid + q * nt of a [.][W] matrix, advanced without a division per entry (the integer
  // divisions by the run-time M and nc were a quarter of this kernel's issue slots and most of its latency)
  struct Walk {
    int i, j, di, dj, W;
    __device__ Walk(int e0, int step, int W_) : i(e0 / W_), j(e0 - (e0 / W_) * W_), di(step / W_), dj(step - (step / W_) * W_), W(W_) {}
    __device__ void next() { i += di; j += dj; if (j >= W) { j -= W; ++i; } }
  };
  {
    Walk w(tid, nt, nc);
    for (int e = tid; e < M * nc; e += nt, w.next()) Hs[w.i * ncp + w.j] = small[SM_HC + e];
  }
  {
    Walk w(tid, nt, nc);
    for (int e = tid; e < nc * nc; e += nt, w.next()) Ps[e] = P[upper_index(cols[w.i], cols[w.j], d.ld, d.sym)];
  }
  __syncthreads();
  // T and S: a thread owns 2 x 2 blocks of the result -- rows (a, a + half) x columns (b, b + half), so that the lanes of
  // a warp still walk consecutive columns -- and carries the four sums through the dot-product loop together: four
  // shared-memory loads for four multiply-adds (one entry per thread took two loads per multiply-add, and the loops
  // were bound by the load unit of the one SM: 5 us each at M = 35, 21 us at M = 65).  Every sum keeps its order.
  {
    // T[cols[c]][r] = sum_b P[cols[c]][cols[b]] Hc[r][b], ascending b
    const int ch = (nc + 1) >> 1, rh = (M + 1) >> 1;
    int cb = tid / rh, rb = tid - cb * rh;
    const int dcb = nt / rh, drb = nt - dcb * rh;
    for (int bid = tid; bid < ch * rh; bid += nt) {
      const int c0 = cb, c1 = cb + ch, r0 = rb, r1 = rb + rh;
      const double* p0 = Ps + c0 * nc;
      const double* p1 = Ps + min(c1, nc - 1) * nc;
      const double* h0 = Hs + r0 * ncp;
      const double* h1 = Hs + min(r1, M - 1) * ncp;
      double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
#pragma unroll 4
      for (int b = 0; b < nc; ++b) {
        const double pv0 = p0[b], pv1 = p1[b], hv0 = h0[b], hv1 = h1[b];
        a00 += pv0 * hv0;
        a01 += pv0 * hv1;
        a10 += pv1 * hv0;
        a11 += pv1 * hv1;
      }
      Tc[c0 * M + r0] = a00;
      if (r1 < M) Tc[c0 * M + r1] = a01;
      if (c1 < nc) {
        Tc[c1 * M + r0] = a10;
        if (r1 < M) Tc[c1 * M + r1] = a11;
      }
      cb += dcb; rb += drb;
      if (rb >= rh) { rb -= rh; ++cb; }
    }
  }
  __syncthreads();
  const Walk w0(tid, nt, M);  // the start of every [M][M] walk below
  {
    // S[r][q] = sum_c Hc[r][c] T[cols[c]][q] (+ R on the diagonal), ascending c
    const int rh = (M + 1) >> 1;
    int rb = tid / rh, qb = tid - rb * rh;
    const int drb = nt / rh, dqb = nt - drb * rh;
    for (int bid = tid; bid < rh * rh; bid += nt) {
      const int r0 = rb, r1 = rb + rh, q0 = qb, q1 = qb + rh;
      const double* h0 = Hs + r0 * ncp;
      const double* h1 = Hs + min(r1, M - 1) * ncp;
      const double* t0 = Tc + q0;
      const double* t1 = Tc + min(q1, M - 1);
      double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
#pragma unroll 4
      for (int c = 0; c < nc; ++c) {
        const double hv0 = h0[c], hv1 = h1[c], tv0 = t0[c * M], tv1 = t1[c * M];
        a00 += hv0 * tv0;
        a01 += hv0 * tv1;
        a10 += hv1 * tv0;
        a11 += hv1 * tv1;
      }
      if (r0 == q0) a00 += small[SM_R + r0];
      S[r0 * M + q0] = a00;
      if (q1 < M) S[r0 * M + q1] = a01;  // r0 < rh <= q1: never on the diagonal
      if (r1 < M) {
        S[r1 * M + q0] = a10;            // q0 < rh <= r1
        if (q1 < M) {
          if (r1 == q1) a11 += small[SM_R + r1];
          S[r1 * M + q1] = a11;
        }
      }
      rb += drb; qb += dqb;
      if (qb >= rh) { qb -= rh; ++rb; }
    }
  }
  __syncthreads();
  Walk wc = w0;
  for (int e = tid; e < M * M; e += nt, wc.next()) {
    const int r = wc.i, q = wc.j;
    small[SM_S + r * MAXM + q] = S[e];
    W[e] = 0.5 * (S[e] + S[q * M + r]);  // invert the symmetric part
  }
  __syncthreads();
  // a step reads the matrix as it stands and writes the next one into the other buffer: one barrier per step
  double* Wsrc = W;
  double* Wdst = W2;
  // M <= 39 (every live size; M <= 67 with 512 threads): a thread's entries stay in registers from step to step
  // and a step is branch-free (gj_register_steps above)
  constexpr int GJQ = NT == 256 ? 7 : 10;  // rows of its column a thread owns: 256 threads: M <= 39; 512 threads: M <= 67
  const int gj_rg = NT / M;
  if (gj_rg >= 1 && (M + gj_rg - 1) / gj_rg <= GJQ && nt == NT) {
    gj_register_steps<GJQ, NT>(W, W2, M, tid, flag, &bad);
    Wsrc = (M & 1) ? W2 : W;
  } else
  for (int k = 0; k < M; ++k) {
    const double piv = Wsrc[k * M + k];
    const bool ok = (piv > 0.0) && (piv <= 1.79769313486231571e308);
    const double ip = ok ? 1.0 / piv : 1.0;
    if (!ok && tid == 0) { *flag = 1; bad = 1; }
    Walk wg = w0;
    for (int e = tid; e < M * M; e += nt, wg.next()) {
      const int i = wg.i, j = wg.j;
      const double cik = Wsrc[i * M + k], rkj = Wsrc[k * M + j] * ip;
      double v;
      if (i == k) v = (j == k) ? ip : rkj;
      else if (j == k) v = -cik * ip;
      else v = Wsrc[e] - cik * rkj;
      Wdst[e] = v;
    }
    __syncthreads();
    double* tmp = Wsrc; Wsrc = Wdst; Wdst = tmp;
  }
  W = Wsrc;
  Walk wf = w0;
  for (int e = tid; e < M * M; e += nt, wf.next()) {
    const int r = wf.i, q = wf.j;
    small[SM_SINV + r * MAXM + q] = bad ? 0.0 : 0.5 * (W[e] + W[q * M + r]);
  }
}

// ROWS_R consecutive state rows of ekf_rows_kernel at once: T, K, Y rows and the rows' shares of K nu.  A lane owns
// the entries lane, lane + 32, ... (NP = ceil(M / 32) of them) of every row and carries all NP x ROWS_R sums TOGETHER
// through each loop: independent chains instead of passes (M = 35 spent a whole second pass on three entries), and a
// table entry (H, S^-1, S) is read from shared memory ONCE for the ROWS_R rows -- with one row per pass the kernel ran at
// 76 % of the shared-memory bandwidth at M = 65 (three table reads for three multiply-adds).  Every sum runs over the
// same terms in the same order as before.  Rows past N are computed on row N - 1's data and not stored.
constexpr int ROWS_R = 4;
template <int NP>
__device__ __forceinline__ void rows_multi(int M, int nc, int ncp, int Mo, int lane, const double* pr, double* trow,
                                           double* krow, const double* Hc, const double* Sinv, const double* S,
                                           const double* nu, double* __restrict__ Kt, double* __restrict__ Yt,
                                           size_t ldk, int row0, int nrows, double (&dx)[ROWS_R]) {
  // pr [ROWS_R][nc], trow [ROWS_R][M], krow [ROWS_R][M]
  {
    double a[NP][ROWS_R];
    const double* h[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      h[p] = Hc + min(lane + 32 * p, M - 1) * ncp;
#pragma unroll
      for (int r = 0; r < ROWS_R; ++r) a[p][r] = 0.0;
    }
    for (int c = 0; c < nc; ++c) {
      double pv[ROWS_R], hv[NP];
#pragma unroll
      for (int r = 0; r < ROWS_R; ++r) pv[r] = pr[r * nc + c];
#pragma unroll
      for (int p = 0; p < NP; ++p) hv[p] = h[p][c];
#pragma unroll
      for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int r = 0; r < ROWS_R; ++r) a[p][r] += pv[r] * hv[p];
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
      if (lane + 32 * p < M) {
#pragma unroll
        for (int r = 0; r < ROWS_R; ++r) trow[r * M + lane + 32 * p] = a[p][r];
      }
  }
  __syncwarp();
  {
    double a[NP][ROWS_R];
    int q[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      q[p] = min(lane + 32 * p, M - 1);
#pragma unroll
      for (int r = 0; r < ROWS_R; ++r) a[p][r] = 0.0;
    }
    for (int k = 0; k < M; ++k) {
      double tv[ROWS_R], sv[NP];
#pragma unroll
      for (int r = 0; r < ROWS_R; ++r) tv[r] = trow[r * M + k];
      const double* sr = Sinv + k * M;
#pragma unroll
      for (int p = 0; p < NP; ++p) sv[p] = sr[q[p]];
#pragma unroll
      for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int r = 0; r < ROWS_R; ++r) a[p][r] += tv[r] * sv[p];
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
      if (lane + 32 * p < M) {
#pragma unroll
        for (int r = 0; r < ROWS_R; ++r) {
          krow[r * M + lane + 32 * p] = a[p][r];
          if (r < nrows) Kt[(size_t)(lane + 32 * p) * ldk + row0 + r] = a[p][r];
        }
      }
  }
  __syncwarp();
  {
    double a[NP][ROWS_R];
    const double* sr[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      sr[p] = S + min(lane + 32 * p, M - 1) * Mo;
#pragma unroll
      for (int r = 0; r < ROWS_R; ++r) a[p][r] = 0.0;
    }
    for (int k = 0; k < M; ++k) {
      double kv[ROWS_R], sv[NP];
#pragma unroll
      for (int r = 0; r < ROWS_R; ++r) kv[r] = krow[r * M + k];
#pragma unroll
      for (int p = 0; p < NP; ++p) sv[p] = sr[p][k];
#pragma unroll
      for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int r = 0; r < ROWS_R; ++r) a[p][r] += kv[r] * sv[p];
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
      if (lane + 32 * p < M) {
#pragma unroll
        for (int r = 0; r < ROWS_R; ++r) {
          if (r < nrows) Yt[(size_t)(lane + 32 * p) * ldk + row0 + r] = a[p][r];
          dx[r] += krow[r * M + lane + 32 * p] * nu[lane + 32 * p];
        }
      }
  }
}

__global__ void __launch_bounds__(512) ekf_rows_kernel(UpdateDesc d, const double* __restrict__ P,
                                                       const double* __restrict__ small, const int* __restrict__ cols,
                                                       double* __restrict__ Kt, double* __restrict__ Yt,
                                                       double* __restrict__ x, int ldk) {
  extern __shared__ double sh[];
  const int M = d.M, nc = d.nc;
  // odd pitches where a matrix is read with its ROW index across the lanes (Hc in the T loop, S in the Y loop)
  const int ncp = nc | 1, Mo = M | 1;
  double* Hc = sh;                  // [M][ncp]
  double* Sinv = Hc + M * ncp;      // [M][M]
  double* S = Sinv + M * M;         // [M][Mo]
  double* nu = S + M * Mo;          // [M]
  const int nw = blockDim.x >> 5;   // 8 warps; 16 when the tables leave room for one block per SM only (M > 36)
  double* rowbuf = nu + M;          // [nw warps][ROWS_R][nc + 2 M]
  int* scols = reinterpret_cast<int*>(rowbuf + nw * ROWS_R * (nc + 2 * M));
  for (int e = threadIdx.x; e < M * nc; e += blockDim.x) {
    const int r = e / nc, b = e - r * nc;
    Hc[r * ncp + b] = small[SM_HC + e];
  }
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int r = e / M, q = e - r * M;
    Sinv[e] = small[SM_SINV + r * MAXM + q];
    S[r * Mo + q] = small[SM_S + r * MAXM + q];
  }
  for (int e = threadIdx.x; e < M; e += blockDim.x) nu[e] = small[SM_NU + e];
  for (int e = threadIdx.x; e < nc; e += blockDim.x) scols[e] = cols[e];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* pr = rowbuf + w * ROWS_R * (nc + 2 * M);
  double* trow = pr + ROWS_R * nc;
  double* krow = trow + ROWS_R * M;
  for (int row0 = (blockIdx.x * nw + w) * ROWS_R; row0 < d.N; row0 += gridDim.x * nw * ROWS_R) {
    const int nrows = min(ROWS_R, d.N - row0);
    for (int e = lane; e < ROWS_R * nc; e += 32) {
      const int r = e / nc, c = e - r * nc;
      pr[e] = P[upper_index(min(row0 + r, d.N - 1), scols[c], d.ld, d.sym)];
    }
    __syncwarp();
    double dx[ROWS_R];
#pragma unroll
    for (int r = 0; r < ROWS_R; ++r) dx[r] = 0.0;
    switch ((M + 31) >> 5) {
      case 1: rows_multi<1>(M, nc, ncp, Mo, lane, pr, trow, krow, Hc, Sinv, S, nu, Kt, Yt, (size_t)ldk, row0, nrows, dx); break;
      case 2: rows_multi<2>(M, nc, ncp, Mo, lane, pr, trow, krow, Hc, Sinv, S, nu, Kt, Yt, (size_t)ldk, row0, nrows, dx); break;
      case 3: rows_multi<3>(M, nc, ncp, Mo, lane, pr, trow, krow, Hc, Sinv, S, nu, Kt, Yt, (size_t)ldk, row0, nrows, dx); break;
      default: rows_multi<4>(M, nc, ncp, Mo, lane, pr, trow, krow, Hc, Sinv, S, nu, Kt, Yt, (size_t)ldk, row0, nrows, dx); break;
    }
#pragma unroll
    for (int r = 0; r < ROWS_R; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dx[r] += __shfl_xor_sync(0xffffffffu, dx[r], o);
    }
    if (lane < nrows) {
      const int row = row0 + lane;
      double dxr = dx[0];
#pragma unroll
      for (int r = 1; r < ROWS_R; ++r)
        if (lane == r) dxr = dx[r];
      double v = x[row] + dxr;
      if (row >= 3 && row <= 5 && fabs(v) < 1e6) {  // kalmanClass.cpp:902-913 (bounded: a wild angle cannot spin)
        while (v > kPI) v -= 2 * kPI;
        while (v < -kPI) v += 2 * kPI;
      }
      x[row] = v;
    }
    __syncwarp();
  }
}

// ---- kernel 5: P <- sym(P - K S K^T), one block per mirrored tile pair --------------------------------------
// 256 threads, thread (tx,ty) owns the 4x4 micro-tile rows I*64+4ty.., cols J*64+4tx.. of tile (I,J) and its
// mirror image in tile (J,I).  Warps are 8 (tx) x 4 (ty) threads so that the per-step shared-memory reads are
// 1 + 2 wavefronts (K is read by 4 distinct ty, Y by 8 distinct tx).
__global__ void __launch_bounds__(256, 2) ekf_cov_kernel(double* __restrict__ P, int ld, int M,
                                                          const double* __restrict__ Kt, const double* __restrict__ Yt,
                                                          int ldk, int ntiles) {
  extern __shared__ double sh[];
  double* Ks = sh;             // [M][64]  K rows of tile I
  double* Ys = sh + M * TILE;  // [M][64]  (K S^T) rows of tile J
  // pair index -> (I <= J)
  const long long pidx = blockIdx.x;
  int J = (int)((sqrt(8.0 * (double)pidx + 1.0) - 1.0) * 0.5);
  while ((long long)(J + 1) * (J + 2) / 2 <= pidx) ++J;
  while ((long long)J * (J + 1) / 2 > pidx) --J;
  const int I = (int)(pidx - (long long)J * (J + 1) / 2);
  (void)ntiles;

  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int tx = (w & 1) * 8 + (l & 7);
  const int ty = (w >> 1) * 4 + (l >> 3);
  const int i0 = I * TILE + 4 * ty, j0 = J * TILE + 4 * tx;
  const bool diag = (I == J);
  const bool active = !diag || tx >= ty;

  // issue the P loads first so they overlap the K/Y staging
  double pij[4][4], pji[4][4];
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double2* s0 = reinterpret_cast<const double2*>(P + (size_t)(i0 + a) * ld + j0);
      const double2 v0 = s0[0], v1 = s0[1];
      pij[a][0] = v0.x; pij[a][1] = v0.y; pij[a][2] = v1.x; pij[a][3] = v1.y;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const double2* s0 = reinterpret_cast<const double2*>(P + (size_t)(j0 + b) * ld + i0);
      const double2 v0 = s0[0], v1 = s0[1];
      pji[b][0] = v0.x; pji[b][1] = v0.y; pji[b][2] = v1.x; pji[b][3] = v1.y;
    }
  }
  for (int e = t; e < M * (TILE / 2); e += 256) {
    const int r = e / (TILE / 2), c = (e % (TILE / 2)) * 2;
    *reinterpret_cast<double2*>(Ks + r * TILE + c) =
        *reinterpret_cast<const double2*>(Kt + (size_t)r * ldk + I * TILE + c);
    *reinterpret_cast<double2*>(Ys + r * TILE + c) =
        *reinterpret_cast<const double2*>(Yt + (size_t)r * ldk + J * TILE + c);
  }
  __syncthreads();
  if (!active) return;

  double wacc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) wacc[a][b] = 0.0;
#pragma unroll 5
  for (int r = 0; r < M; ++r) {
    const double2 k0 = *reinterpret_cast<const double2*>(Ks + r * TILE + 4 * ty);
    const double2 k1 = *reinterpret_cast<const double2*>(Ks + r * TILE + 4 * ty + 2);
    const double2 y0 = *reinterpret_cast<const double2*>(Ys + r * TILE + 4 * tx);
    const double2 y1 = *reinterpret_cast<const double2*>(Ys + r * TILE + 4 * tx + 2);
    const double kk[4] = {k0.x, k0.y, k1.x, k1.y};
    const double yy[4] = {y0.x, y0.y, y1.x, y1.y};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) wacc[a][b] = fma(kk[a], yy[b], wacc[a][b]);
  }
  double out[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) out[a][b] = 0.5 * (pij[a][b] + pji[b][a]) - wacc[a][b];
  if (diag && tx == ty) {  // the micro-tile straddles the diagonal: make it exactly symmetric
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < a; ++b) out[a][b] = out[b][a];
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double2* dst = reinterpret_cast<double2*>(P + (size_t)(i0 + a) * ld + j0);
    dst[0] = make_double2(out[a][0], out[a][1]);
    dst[1] = make_double2(out[a][2], out[a][3]);
  }
  if (!(diag && tx == ty)) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      double2* dst = reinterpret_cast<double2*>(P + (size_t)(j0 + b) * ld + i0);
      dst[0] = make_double2(out[0][b], out[1][b]);
      dst[1] = make_double2(out[2][b], out[3][b]);
    }
  }
}

// ---- kernel 5b: the same tile-pair update on the FP64 tensor cores (mma.sync.m8n8k4.f64, "DMMA") ----------
// The CUDA-core version above is limited by shared-memory operand bandwidth (8 operand doubles per 16 FMA);
// a DMMA instruction consumes 2 operand doubles per thread for 256 FMA per warp.  8 warps as 4 (rows) x 2
// (cols); a warp owns 16 x 32 outputs = 2 x 4 DMMA tiles.  Fragment layout of m8n8k4 (T = lane):
//   A (8x4, row): a = A[T/4][T%4]   B (4x8, col): b = B[T%4][T/4]   C (8x8): c0,c1 = C[T/4][2*(T%4) + {0,1}]
// K and K S^T tiles sit in shared memory as [r][i] with a row stride of 68 doubles (68 mod 16 = 4 keeps the
// half-warp's 16 eight-byte reads on distinct banks); rows r >= M are zero so M is padded to a multiple of 4.
constexpr int DSTRIDE = TILE + 4;

__global__ void __launch_bounds__(256, 2) ekf_cov_dmma_kernel(double* __restrict__ P, int ld, int M,
                                                               const double* __restrict__ Kt,
                                                               const double* __restrict__ Yt, int ldk) {
  extern __shared__ double sh[];
  const int Mp = (M + 3) & ~3;
  double* Ks = sh;
  double* Ys = sh + Mp * DSTRIDE;
  const long long pidx = blockIdx.x;
  int J = (int)((sqrt(8.0 * (double)pidx + 1.0) - 1.0) * 0.5);
  while ((long long)(J + 1) * (J + 2) / 2 <= pidx) ++J;
  while ((long long)J * (J + 1) / 2 > pidx) --J;
  const int I = (int)(pidx - (long long)J * (J + 1) / 2);
  const bool diag = (I == J);

  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int lr = l >> 2, lc = l & 3;
  const int wr0 = (w >> 1) * 16, wc0 = (w & 1) * 32;  // warp tile origin inside the 64x64 tile
  const size_t gi0 = (size_t)I * TILE + wr0 + lr;      // global row of this lane's C rows (plus 8*mi)
  const size_t gj0 = (size_t)J * TILE + wc0 + 2 * lc;  // global col of c0 (plus 8*ni)

  // prefetch the P entries: tile (I,J) as 16-byte pairs, tile (J,I) transposed as scalars
  double2 pij[2][4];
  double pji[2][4][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      pij[mi][ni] = *reinterpret_cast<const double2*>(P + (gi0 + 8 * mi) * ld + gj0 + 8 * ni);
      pji[mi][ni][0] = P[(gj0 + 8 * ni) * ld + gi0 + 8 * mi];
      pji[mi][ni][1] = P[(gj0 + 8 * ni + 1) * ld + gi0 + 8 * mi];
    }
  for (int e = t; e < Mp * (TILE / 2); e += 256) {
    const int r = e / (TILE / 2), c = (e % (TILE / 2)) * 2;
    double2 kv = make_double2(0.0, 0.0), yv = make_double2(0.0, 0.0);
    if (r < M) {
      kv = *reinterpret_cast<const double2*>(Kt + (size_t)r * ldk + (size_t)I * TILE + c);
      yv = *reinterpret_cast<const double2*>(Yt + (size_t)r * ldk + (size_t)J * TILE + c);
    }
    *reinterpret_cast<double2*>(Ks + r * DSTRIDE + c) = kv;
    *reinterpret_cast<double2*>(Ys + r * DSTRIDE + c) = yv;
  }
  __syncthreads();

  double acc[2][4][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
  for (int ks = 0; ks < Mp; ks += 4) {
    double a[2], b[4];
    const double* kr = Ks + (ks + lc) * DSTRIDE + wr0 + lr;
    const double* yr = Ys + (ks + lc) * DSTRIDE + wc0 + lr;
    a[0] = kr[0];
    a[1] = kr[8];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) b[ni] = yr[8 * ni];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                     : "d"(a[mi]), "d"(b[ni]));
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const size_t gi = gi0 + 8 * mi, gj = gj0 + 8 * ni;
      const double o0 = 0.5 * (pij[mi][ni].x + pji[mi][ni][0]) - acc[mi][ni][0];
      const double o1 = 0.5 * (pij[mi][ni].y + pji[mi][ni][1]) - acc[mi][ni][1];
      if (!diag) {
        *reinterpret_cast<double2*>(P + gi * ld + gj) = make_double2(o0, o1);
        P[gj * ld + gi] = o0;
        P[(gj + 1) * ld + gi] = o1;
      } else {  // diagonal tile: the upper triangle (col >= row) is authoritative and mirrored
        if (gj >= gi) { P[gi * ld + gj] = o0; if (gj != gi) P[gj * ld + gi] = o0; }
        if (gj + 1 >= gi) { P[gi * ld + gj + 1] = o1; if (gj + 1 != gi) P[(gj + 1) * ld + gi] = o1; }
      }
    }
}

// ---- kernel 5c: persistent, double-buffered version of 5b (M <= 36) -------------------------------------------
// One block per SM walks a contiguous range of tile pairs (I-major order, so K_I stays resident while J advances).
// The two P tiles and the Y_J tile of pair p+1 are copied global -> shared with cp.async (LDGSTS: no registers
// held, 64 KB in flight per SM) while pair p is multiplied on the tensor cores and written back: loads, DMMA and
// stores overlap instead of alternating.  Shared layout per stage: P_IJ [64][72] (pitch 72: the 16-byte fragment
// reads of a quarter-warp hit distinct banks), P_JI [64][66] (pitch 66: the transposed 8-byte reads of a half-warp
// hit distinct banks), Y_J [Mp][68]; K_I [Mp][68] single-buffered, reloaded when I changes.
constexpr int PITCH_A = 72, PITCH_B = 66;
constexpr int PIPE_MAX_M = 36;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void pair_from_index(long long p, int nt, int& I, int& J) {
  // pairs in I-major order: row I holds nt - I pairs (J = I..nt-1)
  int i = 0;
  long long rem = p;
  // closed form guess, then fix up
  const double b = 2.0 * nt + 1.0;
  i = (int)((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
  if (i < 0) i = 0;
  if (i > nt - 1) i = nt - 1;
  while (i > 0 && (long long)i * nt - (long long)i * (i - 1) / 2 > p) --i;
  while ((long long)(i + 1) * nt - (long long)(i + 1) * i / 2 <= p) ++i;
  rem = p - ((long long)i * nt - (long long)i * (i - 1) / 2);
  I = i;
  J = i + (int)rem;
}

__global__ void __launch_bounds__(256, 1) ekf_cov_pipe_kernel(double* __restrict__ P, int ld, int M,
                                                               const double* __restrict__ Kt,
                                                               const double* __restrict__ Yt, int ldk, int nt,
                                                               long long total_pairs) {
  extern __shared__ double sh[];
  const int Mp = (M + 3) & ~3;
  const int stage_doubles = TILE * PITCH_A + TILE * PITCH_B + Mp * DSTRIDE;
  double* Ks = sh + 2 * stage_doubles;
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int lr = l >> 2, lc = l & 3;
  const int wr0 = (w >> 1) * 16, wc0 = (w & 1) * 32;

  const long long per = (total_pairs + gridDim.x - 1) / gridDim.x;
  const long long p_begin = (long long)blockIdx.x * per;
  const long long p_end = min(total_pairs, p_begin + per);
  if (p_begin >= p_end) return;

  // zero the padding rows (r >= M) of K and both Y buffers once
  for (int e = t; e < (Mp - M) * DSTRIDE; e += 256) {
    Ks[M * DSTRIDE + e] = 0.0;
    sh[0 * stage_doubles + TILE * PITCH_A + TILE * PITCH_B + M * DSTRIDE + e] = 0.0;
    sh[1 * stage_doubles + TILE * PITCH_A + TILE * PITCH_B + M * DSTRIDE + e] = 0.0;
  }

  auto issue_pair = [&](int I, int J, int stage) {
    double* pa = sh + stage * stage_doubles;
    double* pb = pa + TILE * PITCH_A;
    double* ys = pb + TILE * PITCH_B;
    for (int c = t; c < TILE * 32; c += 256) {
      const int row = c >> 5, cc = (c & 31) * 2;
      cp_async16(pa + row * PITCH_A + cc, P + ((size_t)I * TILE + row) * ld + (size_t)J * TILE + cc);
      if (I != J) cp_async16(pb + row * PITCH_B + cc, P + ((size_t)J * TILE + row) * ld + (size_t)I * TILE + cc);
    }
    for (int c = t; c < M * 32; c += 256) {
      const int r = c >> 5, cc = (c & 31) * 2;
      cp_async16(ys + r * DSTRIDE + cc, Yt + (size_t)r * ldk + (size_t)J * TILE + cc);
    }
  };
  auto issue_k = [&](int I) {
    for (int c = t; c < M * 32; c += 256) {
      const int r = c >> 5, cc = (c & 31) * 2;
      cp_async16(Ks + r * DSTRIDE + cc, Kt + (size_t)r * ldk + (size_t)I * TILE + cc);
    }
  };

  int I, J;
  pair_from_index(p_begin, nt, I, J);
  int k_loaded = I;
  issue_k(I);
  issue_pair(I, J, 0);
  cp_async_commit();

  for (long long p = p_begin; p < p_end; ++p) {
    const int stage = (int)((p - p_begin) & 1);
    int In = I, Jn = J + 1;
    if (Jn >= nt) { In = I + 1; Jn = In; }
    const bool has_next = (p + 1 < p_end);
    if (has_next) issue_pair(In, Jn, stage ^ 1);
    cp_async_commit();
    if (has_next) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    if (k_loaded != I) {  // first pair of a new tile row: bring K_I in (rare: once per row)
      issue_k(I);
      cp_async_commit();
      cp_async_wait<0>();
      k_loaded = I;
      __syncthreads();
    }
    const double* pa = sh + stage * stage_doubles;
    const double* pb = pa + TILE * PITCH_A;
    const double* ys = pb + TILE * PITCH_B;
    const bool diag = (I == J);

    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
    for (int ks = 0; ks < Mp; ks += 4) {
      double a[2], b[4];
      const double* kr = Ks + (ks + lc) * DSTRIDE + wr0 + lr;
      const double* yr = ys + (ks + lc) * DSTRIDE + wc0 + lr;
      a[0] = kr[0];
      a[1] = kr[8];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = yr[8 * ni];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                       : "d"(a[mi]), "d"(b[ni]));
    }
    const size_t gi0 = (size_t)I * TILE + wr0 + lr, gj0 = (size_t)J * TILE + wc0 + 2 * lc;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int ri = wr0 + lr + 8 * mi, cj = wc0 + 2 * lc + 8 * ni;  // position inside the 64x64 tile
        const double2 pij = *reinterpret_cast<const double2*>(pa + ri * PITCH_A + cj);
        double q0, q1;
        if (!diag) { q0 = pb[cj * PITCH_B + ri]; q1 = pb[(cj + 1) * PITCH_B + ri]; }
        else { q0 = pa[cj * PITCH_A + ri]; q1 = pa[(cj + 1) * PITCH_A + ri]; }
        const double o0 = 0.5 * (pij.x + q0) - acc[mi][ni][0];
        const double o1 = 0.5 * (pij.y + q1) - acc[mi][ni][1];
        const size_t gi = gi0 + 8 * mi, gj = gj0 + 8 * ni;
        if (!diag) {
          *reinterpret_cast<double2*>(P + gi * ld + gj) = make_double2(o0, o1);
          P[gj * ld + gi] = o0;
          P[(gj + 1) * ld + gi] = o1;
        } else {
          if (gj >= gi) { P[gi * ld + gj] = o0; if (gj != gi) P[gj * ld + gi] = o0; }
          if (gj + 1 >= gi) { P[gi * ld + gj + 1] = o1; if (gj + 1 != gi) P[(gj + 1) * ld + gi] = o1; }
        }
      }
    __syncthreads();  // everyone is done with this stage before the prefetch of pair p+2 overwrites it
    I = In;
    J = Jn;
  }
}

// ---- kernel 5d: the tile-pair update streamed with TMA, reading HALF of P ---------------------------------------
// Whenever the off-diagonal tiles of P are exact mirror images of each other -- true after every update (this
// kernel and the ones above write both tiles from one value), after cps_ekf_set_cov_lowrank, and preserved by the
// predict step (it rewrites 12 rows and the mirrored 12 columns with the same values; the 12 x 12 pose block lies
// inside the diagonal tile (0,0)) -- 0.5*(P_ij + P_ji) is P_ij exactly, so the mirrored tile need not be read: the
// pass moves 12 N^2 bytes instead of 16 N^2 and its result is bit-identical to kernel 5c's.  Diagonal tiles are
// symmetrised from their own entries as before.
//   * P tiles arrive by cp.async.bulk.tensor (TMA, SASS UTMALDG) into a 3-deep ring, completion on mbarriers;
//   * the corrected tile is written in place in its stage and leaves by one TMA store (UTMASTG); its transpose is
//     written into a second buffer in the 128-byte-swizzled layout of four 16-column boxes (conflict-free 8-byte
//     stores from the DMMA accumulator layout) and leaves by four TMA stores into tile (J,I): both stores are
//     full lines;
//   * the Y_J rows of a pair (M x 64, L2-resident) arrive as M bulk copies of 512 bytes issued by the copy warp's
//     32 lanes into the padded layout the DMMA fragment reads need, on the same mbarrier as the pair's P tile: the
//     compute warps issue no copies and wait on one barrier per pair (their share of the cp.async issue loop was
//     12 % of the kernel's instructions); K_I comes by cp.async once per tile row and, for M <= 36, its fragments
//     stay in registers while J advances;
//   * by default only tile (I,J), I <= J, of a pair is written (write_lower = 0): every reader on the predict /
//     update / gate / augmentation path takes single entries from the upper copy (upper_index), and the tiles below
//     the diagonal tiles are rebuilt by ekf_mirror_lower_kernel when an entry point needs them (cps_ekf::lower_stale):
//     8 N^2 bytes per update instead of 12 N^2.
constexpr int TMA_STAGES = 3;
constexpr int TMA_TILE_BYTES = TILE * TILE * 8;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"(smem_u32(src))
               : "memory");
}
// one contiguous run (a multiple of 16 bytes, both ends 16-byte aligned) global -> shared, completing on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the copy warp's wait: it polls one barrier for most of a pair's duration, so it sleeps between polls instead of
// competing with two compute warps for its scheduler's issue slots
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity) {
  for (;;) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(64);
  }
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// element (row, col) of a 64 x 64 tile stored as four 16-column boxes in the TMA 128-byte swizzle
__device__ __forceinline__ int swz_off(int row, int col) {
  return (col >> 4) * (TILE * 16) + row * 16 + ((((col & 15) >> 1) ^ (row & 7)) << 1) + (col & 1);
}

struct alignas(64) TmaMaps {
  CUtensorMap full;  // P as [cap][cap], box 64 x 64, no swizzle: tile loads and in-place stores
  CUtensorMap swz;   // P as [cap][cap], box 16 columns x 64 rows, 128-byte swizzle: transposed stores
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void compute_warps_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 8 compute warps + 1 copy warp.  The copy warp's lane 0 owns every TMA operation: it fills the 3-deep ring of P
// tiles, and for every pair waits until the compute warps have written the corrected tile and its transpose
// (mbarrier `done`), stores both, waits until the stores have READ shared memory, then hands the transpose buffer
// back (`tfree`) and refills the freed stage with the pair three ahead.  The compute warps therefore never wait
// for a store: they only wait for tiles that were requested two to three pairs earlier.
// STAGES = depth of the P-tile ring and of the Y ring: 3 for M <= 36 (the reference's live sizes), 2 for M <= 80 (a
// stage's Y tile grows with M and the four-warp K / Y fragments keep their padded layout, so the deeper ring no
// longer fits the 227 KB of an SM; at those sizes a pair carries 2-3x the DMMA work and one tile in flight covers it).
constexpr int TMA_THREADS = 288;
constexpr int TMA2_MAX_M = 80;
template <int STAGES>
__global__ void __launch_bounds__(TMA_THREADS, 1) ekf_cov_tma_kernel(const __grid_constant__ TmaMaps maps, int M,
                                                                      const double* __restrict__ Kt,
                                                                      const double* __restrict__ Yt, int ldk, int nt,
                                                                      long long total_pairs, int write_lower) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<size_t>(smem_raw) + 1023) & ~(size_t)1023);
  const int Mp = (M + 3) & ~3;
  double* Tt = reinterpret_cast<double*>(base);                                     // 32 KB, swizzled transpose
  double* Pst = reinterpret_cast<double*>(base + TMA_TILE_BYTES);                   // [3][64][64]
  double* Yst = reinterpret_cast<double*>(base + (1 + STAGES) * TMA_TILE_BYTES);  // [3][Mp][68]
  double* Ks = Yst + STAGES * Mp * DSTRIDE;                                     // [Mp][68]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(Ks + Mp * DSTRIDE);  // [3]
  unsigned long long* done = full + STAGES;                                     // compute -> copy warp
  unsigned long long* tfree = done + 1;                                             // copy warp -> compute
  const int t = threadIdx.x, w = t >> 5, l = t & 31;

  const long long per = (total_pairs + gridDim.x - 1) / gridDim.x;
  const long long p_begin = (long long)blockIdx.x * per;
  const long long p_end = min(total_pairs, p_begin + per);
  if (p_begin >= p_end) return;
  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
    mbar_init(done, 1);
    mbar_init(tfree, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 256)
    for (int e = t; e < (Mp - M) * DSTRIDE; e += 256) {  // padding rows of K and Y stay zero
      Ks[M * DSTRIDE + e] = 0.0;
      for (int s = 0; s < STAGES; ++s) Yst[s * Mp * DSTRIDE + M * DSTRIDE + e] = 0.0;
    }
  __syncthreads();
  auto next_pair = [&](int& I, int& J) {
    if (++J >= nt) { ++I; J = I; }
  };
  int I, J;
  pair_from_index(p_begin, nt, I, J);

  if (w == 8) {
    // ---------------------------------------------------------------------------------------- copy warp
    // lane 0 owns the TMA operations on P; all 32 lanes bring the Y_J rows of a pair (M rows of 512 bytes into the
    // padded layout the fragment reads need) as bulk copies that complete on the same mbarrier as the P tile, so the
    // compute warps issue no copies at all and wait on one barrier per pair
    auto issue_pair = [&](int Ia, int Ja, int stage) {
      if (l == 0) {
        mbar_expect_tx(full + stage, (unsigned)(TMA_TILE_BYTES + M * TILE * 8));
        tma_load_2d(Pst + stage * TILE * TILE, &maps.full, Ja * TILE, Ia * TILE, full + stage);
      }
      __syncwarp();
      double* ys = Yst + stage * Mp * DSTRIDE;
      for (int r = l; r < M; r += 32)
        bulk_load_1d(ys + r * DSTRIDE, Yt + (size_t)r * ldk + (size_t)Ja * TILE, TILE * 8, full + stage);
    };
    int Il = I, Jl = J;  // next pair to load
    long long pl = p_begin;
    for (int s = 0; s < STAGES && pl < p_end; ++s, ++pl) {
      issue_pair(Il, Jl, s);
      next_pair(Il, Jl);
    }
    for (long long p = p_begin; p < p_end; ++p) {
      const int it = (int)(p - p_begin);
      const int stage = it % STAGES;
      mbar_wait_backoff(done, (unsigned)(it & 1));  // tile and transpose of pair p are in shared memory, fenced
      if (l == 0) {
        const double* pa = Pst + stage * TILE * TILE;
        if (I != J) {
          tma_store_2d(&maps.full, J * TILE, I * TILE, pa);
          if (write_lower) {
#pragma unroll
            for (int q = 0; q < 4; ++q) tma_store_2d(&maps.swz, I * TILE + 16 * q, J * TILE, Tt + q * TILE * 16);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) tma_store_2d(&maps.swz, I * TILE + 16 * q, I * TILE, Tt + q * TILE * 16);
        }
        tma_commit();
        tma_wait_read_all();  // the stores have read the stage and the transpose buffer
        mbar_arrive(tfree);
      }
      __syncwarp();
      if (pl < p_end) {  // the stage (P tile and Y rows: the compute warps are past the pair's products) takes the pair S ahead
        issue_pair(Il, Jl, stage);
        next_pair(Il, Jl);
        ++pl;
      }
      next_pair(I, J);
    }
    if (l == 0) tma_wait_all();
    return;
  }

  // ------------------------------------------------------------------------------------------ compute warps
  const int lr = l >> 2, lc = l & 3;
  const int wr0 = (w >> 1) * 16, wc0 = (w & 1) * 32;
  auto issue_k = [&](int Ia) {
    for (int c = t; c < M * 32; c += 256) {
      const int r = c >> 5, cc = (c & 31) * 2;
      cp_async16(Ks + r * DSTRIDE + cc, Kt + (size_t)r * ldk + (size_t)Ia * TILE + cc);
    }
  };
  // K_I stays the same while J advances: with at most nine k-steps (the three-stage variant, M <= 36) a thread keeps its
  // K fragments in registers from pair to pair (two of the six shared-memory loads of a k-step)
  constexpr int KREG = STAGES == 3 ? PIPE_MAX_M / 4 : 1;
  double ka[KREG][2];
  auto load_k = [&](int Ia) {
    issue_k(Ia);
    cp_async_commit();
    cp_async_wait<0>();
    compute_warps_sync();
    if (STAGES == 3) {
#pragma unroll
      for (int s = 0; s < KREG; ++s) {
        const double* kr = Ks + (4 * s + lc) * DSTRIDE + wr0 + lr;
        ka[s][0] = (4 * s < Mp) ? kr[0] : 0.0;
        ka[s][1] = (4 * s < Mp) ? kr[8] : 0.0;
      }
    }
  };
  int k_loaded = I;
  load_k(I);
  for (long long p = p_begin; p < p_end; ++p) {
    const int it = (int)(p - p_begin);
    const int stage = it % STAGES;
    if (k_loaded != I) {  // every warp is past the products of the pair before (the barrier that ends a pair)
      load_k(I);
      k_loaded = I;
    }
    mbar_wait(full + stage, (unsigned)((it / STAGES) & 1));  // the P tile and the Y rows of pair p
    double* pa = Pst + stage * TILE * TILE;
    const double* ys = Yst + stage * Mp * DSTRIDE;
    const bool diag = (I == J);

    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
    auto mma_step = [&](int ks, double a0, double a1) {
      double b[4];
      const double* yr = ys + (ks + lc) * DSTRIDE + wc0 + lr;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = yr[8 * ni];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[0][ni][0]), "+d"(acc[0][ni][1])
                     : "d"(a0), "d"(b[ni]));
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[1][ni][0]), "+d"(acc[1][ni][1])
                     : "d"(a1), "d"(b[ni]));
      }
    };
    if (STAGES == 3) {
#pragma unroll
      for (int s = 0; s < KREG; ++s)
        if (4 * s < Mp) mma_step(4 * s, ka[s][0], ka[s][1]);
    } else {
      for (int ks = 0; ks < Mp; ks += 4) {
        const double* kr = Ks + (ks + lc) * DSTRIDE + wr0 + lr;
        mma_step(ks, kr[0], kr[8]);
      }
    }
    if (it > 0) mbar_wait(tfree, (unsigned)((it - 1) & 1));  // the stores of pair p - 1 have read the transpose buffer
    if (!diag) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int ri = wr0 + lr + 8 * mi, cj = wc0 + 2 * lc + 8 * ni;
          double2* pp = reinterpret_cast<double2*>(pa + ri * TILE + cj);
          const double2 pij = *pp;
          // the mirrored tile holds the same values: 0.5*(p + p) - acc == p - acc bit for bit
          const double o0 = pij.x - acc[mi][ni][0];
          const double o1 = pij.y - acc[mi][ni][1];
          *pp = make_double2(o0, o1);
          if (write_lower) {
            Tt[swz_off(cj, ri)] = o0;
            Tt[swz_off(cj + 1, ri)] = o1;
          }
        }
    } else {
      // diagonal tile: entries on and above the diagonal are authoritative and mirrored, as in kernel 5c
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int ri = wr0 + lr + 8 * mi, cj = wc0 + 2 * lc + 8 * ni;
          const double2 pij = *reinterpret_cast<const double2*>(pa + ri * TILE + cj);
          const double q0 = pa[cj * TILE + ri], q1 = pa[(cj + 1) * TILE + ri];
          const double o0 = 0.5 * (pij.x + q0) - acc[mi][ni][0];
          const double o1 = 0.5 * (pij.y + q1) - acc[mi][ni][1];
          if (cj >= ri) { Tt[swz_off(ri, cj)] = o0; if (cj != ri) Tt[swz_off(cj, ri)] = o0; }
          if (cj + 1 >= ri) { Tt[swz_off(ri, cj + 1)] = o1; if (cj + 1 != ri) Tt[swz_off(cj + 1, ri)] = o1; }
        }
    }
    fence_async_smem();
    compute_warps_sync();
    if (t == 0) mbar_arrive(done);
    next_pair(I, J);
  }
}

// ---- kernel 5e: the half-written streaming pass with TWO compute groups on alternating pairs (M <= 36) -------------
// Kernel 5d's eight compute warps walk a pair in lockstep -- products, then the wait for the accumulators, then the
// epilogue, then a barrier -- so the DMMA pipe idles during the epilogue and the load / store unit during the products
// (DMMA 58 % busy, 28 % of the stall samples on the first use of the accumulators), and interleaving the two inside
// a warp lost twice.  Here warps 0-3 take the even pairs of the CTA's range and warps 4-7 the odd ones,
// each group a whole 64 x 64 tile (a warp: 16 rows x 64 columns, 16 DMMAs per k-step), with its own named barrier and
// its own pairs' `done` mbarriers: while one group corrects and hands over its tile the other multiplies.  Only the
// half-written mode (write_lower = 0 of kernel 5d): no transpose buffer -- a diagonal tile is symmetrised in place
// in two phases (all reads, group barrier, all writes) and leaves by the same full-tile store -- which pays for a
// 4-deep ring of P tiles and Y rows (two pairs in work, two in flight).  K_I changes once per tile row: a thread
// keeps its K fragments in registers (nine k-steps) and reloads them from Kt (L2) when its pair's row changes, so the
// groups may be on different tile rows.  Same sums in the same order as kernel 5d: bit-identical.
// KS = k-steps a thread keeps K fragments for, NST = ring depth: <9, 4> for M <= 36, <17, 3> for M <= 68 (a stage's Y rows
// grow with M: three stages of 32 + 37 KB; two pairs in work, one in flight -- a pair carries twice the DMMA work there).
__device__ __forceinline__ void group_sync(int g) {
  if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
template <int KS, int NST>
__global__ void __launch_bounds__(TMA_THREADS, 1) ekf_cov_pp_kernel(const __grid_constant__ TmaMaps maps, int M,
                                                                     const double* __restrict__ Kt,
                                                                     const double* __restrict__ Yt, int ldk, int nt,
                                                                     long long total_pairs) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<size_t>(smem_raw) + 1023) & ~(size_t)1023);
  const int Mp = (M + 3) & ~3;
  double* Pst = reinterpret_cast<double*>(base);                                  // [4][64][64]
  double* Yst = reinterpret_cast<double*>(base + NST * TMA_TILE_BYTES);     // [4][Mp][68]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(Yst + NST * Mp * DSTRIDE);  // [4]
  // [4] one per stage: a stage's `full` and `done` phases alternate strictly (a group can only run ahead as far as the
  // ring has been refilled), so a parity wait can never be lapped
  unsigned long long* done = full + NST;
  const int t = threadIdx.x, w = t >> 5, l = t & 31;

  const long long per = (total_pairs + gridDim.x - 1) / gridDim.x;
  const long long p_begin = (long long)blockIdx.x * per;
  const long long p_end = min(total_pairs, p_begin + per);
  if (p_begin >= p_end) return;
  const int n_pairs = (int)(p_end - p_begin);
  if (t == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full + s, 1); mbar_init(done + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 256)
    for (int e = t; e < (Mp - M) * DSTRIDE; e += 256)  // padding rows of Y stay zero
      for (int s = 0; s < NST; ++s) Yst[s * Mp * DSTRIDE + M * DSTRIDE + e] = 0.0;
  __syncthreads();
  auto next_pair = [&](int& I, int& J) {
    if (++J >= nt) { ++I; J = I; }
  };
  int I, J;
  pair_from_index(p_begin, nt, I, J);

  if (w == 8) {
    // ---------------------------------------------------------------------------------------- copy warp
    auto issue_pair = [&](int Ia, int Ja, int stage) {
      if (l == 0) {
        mbar_expect_tx(full + stage, (unsigned)(TMA_TILE_BYTES + M * TILE * 8));
        tma_load_2d(Pst + stage * TILE * TILE, &maps.full, Ja * TILE, Ia * TILE, full + stage);
      }
      __syncwarp();
      double* ys = Yst + stage * Mp * DSTRIDE;
      for (int r = l; r < M; r += 32)
        bulk_load_1d(ys + r * DSTRIDE, Yt + (size_t)r * ldk + (size_t)Ja * TILE, TILE * 8, full + stage);
    };
    int Il = I, Jl = J, np = 0;
    for (; np < NST && np < n_pairs; ++np) {
      issue_pair(Il, Jl, np);
      next_pair(Il, Jl);
    }
    for (int it = 0; it < n_pairs; ++it) {
      const int stage = it % NST;
      mbar_wait_backoff(done + stage, (unsigned)((it / NST) & 1));  // the pair's tile is corrected and fenced
      if (l == 0) {
        tma_store_2d(&maps.full, J * TILE, I * TILE, Pst + stage * TILE * TILE);
        tma_commit();
        tma_wait_read_all();  // the store has read the stage
      }
      __syncwarp();
      if (np < n_pairs) {
        issue_pair(Il, Jl, stage);
        next_pair(Il, Jl);
        ++np;
      }
      next_pair(I, J);
    }
    if (l == 0) tma_wait_all();
    return;
  }

  // ------------------------------------------------------------------------------------------ compute groups
  const int g = w >> 2, wr0 = (w & 3) * 16;
  const int lr = l >> 2, lc = l & 3;
  if (g == 1) next_pair(I, J);  // the odd pairs
  double ka[KS][2];
  int k_loaded = -1;
  for (int it = g; it < n_pairs; it += 2) {
    if (k_loaded != I) {
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int r = 4 * s + lc;
        const double* kr = Kt + (size_t)min(r, M - 1) * ldk + (size_t)I * TILE + wr0 + lr;
        ka[s][0] = r < M ? kr[0] : 0.0;
        ka[s][1] = r < M ? kr[8] : 0.0;
      }
      k_loaded = I;
    }
    const int stage = it % NST;
    mbar_wait(full + stage, (unsigned)((it / NST) & 1));  // the P tile and the Y rows of this pair
    double* pa = Pst + stage * TILE * TILE;
    const double* ys = Yst + stage * Mp * DSTRIDE;
    double acc[2][8][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      if (4 * s < Mp) {
        double b[8];
        const double* yr = ys + (4 * s + lc) * DSTRIDE + lr;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) b[ni] = yr[8 * ni];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 8; ++ni)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                         : "d"(ka[s][mi]), "d"(b[ni]));
      }
    }
    if (I != J) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          const int ri = wr0 + lr + 8 * mi, cj = 2 * lc + 8 * ni;
          double2* pp = reinterpret_cast<double2*>(pa + ri * TILE + cj);
          const double2 pij = *pp;
          // the mirrored tile holds the same values: 0.5*(p + p) - acc == p - acc bit for bit
          *pp = make_double2(pij.x - acc[mi][ni][0], pij.y - acc[mi][ni][1]);
        }
    } else {
      // diagonal tile: entries on and above the diagonal are authoritative and mirrored, as in kernels 5c / 5d; every
      // thread reads what it needs before anybody writes
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          const int ri = wr0 + lr + 8 * mi, cj = 2 * lc + 8 * ni;
          const double2 pij = *reinterpret_cast<const double2*>(pa + ri * TILE + cj);
          const double q0 = pa[cj * TILE + ri], q1 = pa[(cj + 1) * TILE + ri];
          acc[mi][ni][0] = 0.5 * (pij.x + q0) - acc[mi][ni][0];
          acc[mi][ni][1] = 0.5 * (pij.y + q1) - acc[mi][ni][1];
        }
      group_sync(g);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          const int ri = wr0 + lr + 8 * mi, cj = 2 * lc + 8 * ni;
          if (cj >= ri) { pa[ri * TILE + cj] = acc[mi][ni][0]; if (cj != ri) pa[cj * TILE + ri] = acc[mi][ni][0]; }
          if (cj + 1 >= ri) { pa[ri * TILE + cj + 1] = acc[mi][ni][1]; if (cj + 1 != ri) pa[(cj + 1) * TILE + ri] = acc[mi][ni][1]; }
        }
    }
    fence_async_smem();
    group_sync(g);
    if ((t & 127) == 0) mbar_arrive(done + stage);
    next_pair(I, J);
    next_pair(I, J);
  }
}

// ---- tiles below the diagonal tiles, on demand ----------------------------------------------------------------------
// The streaming kernel above normally writes only tile (I,J), I <= J, of every mirrored pair (8 N^2 bytes per update
// instead of 12 N^2: cps_ekf::lower_stale).  Everything on the predict / update path reads the upper copy; an entry
// point that hands out or walks full rows (cps_ekf_get_cov*, the two-tile fallback kernels) first rebuilds the lower
// tiles as exact transposes with this kernel.
__global__ void __launch_bounds__(256) ekf_mirror_lower_kernel(double* __restrict__ P, int ld) {
  const int I = blockIdx.y, J = blockIdx.x;
  if (I >= J) return;
  __shared__ double tile[TILE][TILE + 1];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int r = ty; r < TILE; r += 4) tile[r][tx] = P[((size_t)I * TILE + r) * ld + (size_t)J * TILE + tx];
  __syncthreads();
  for (int r = ty; r < TILE; r += 4) P[((size_t)J * TILE + r) * ld + (size_t)I * TILE + tx] = tile[tx][r];
}

// A block of P for cps_ekf_get_cov_block while the lower tiles are stale: every entry from the copy with row <= column.
__global__ void __launch_bounds__(256) ekf_gather_block_kernel(const double* __restrict__ P, int ld, int r0, int c0,
                                                               int nr, int nc, double* __restrict__ out) {
  const long long total = (long long)nr * nc;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / nc), j = (int)(e - (long long)i * nc);
    out[e] = P[upper_index(r0 + i, c0 + j, ld, 1)];
  }
}

// ---- PredictKF (kalmanClass.cpp:161-272): only the 12 pose rows/columns of P change ----------------------
// Pri <- F Pri, mirrored; Prr <- F Prr F^T + Q; x[0..5] += x[6..11] DT.  F = I + DT*[0 I6; 0 0].
// One launch: block 0 advances the 12 x 12 pose block and x, the other blocks the 12 cross rows / columns.  The two
// parts touch disjoint entries of P (the cross part does not read x), so they need no order between them.
__device__ __forceinline__ void ekf_predict_cross(double* __restrict__ P, int ld, int N, int block) {
  const int j = 12 + block * blockDim.x + threadIdx.x;
  if (j >= N) return;
  double col[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) col[i] = P[(size_t)i * ld + j];
#pragma unroll
  for (int i = 0; i < 6; ++i) col[i] = col[i] + DT * col[i + 6];
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    P[(size_t)i * ld + j] = col[i];
    P[(size_t)j * ld + i] = col[i];
  }
}

__device__ __forceinline__ void ekf_predict_pose(double* __restrict__ P, int ld, double* __restrict__ x) {
  __shared__ double Prr[144], Tm[144], Q[144], F[144];
  const int t = threadIdx.x;  // the first 144 threads of the block work, all of them take the barriers
  const bool act = t < 144;
  const int i = act ? t / 12 : 0, j = act ? t % 12 : 0;
  if (act) {
    F[t] = (i == j) ? 1.0 : ((j == i + 6 && i < 6) ? DT : 0.0);
    Prr[t] = P[(size_t)i * ld + j];
    Q[t] = 0.0;
  }
  __syncthreads();
  if (t == 0) {
    const double phi = x[3], theta = x[4], psi = x[5];
    double Reb[9], t33[9], t33b[9];
    Reb[0] = cos(psi) * cos(theta);
    Reb[1] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
    Reb[2] = cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi);
    Reb[3] = sin(psi) * cos(theta);
    Reb[4] = sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi);
    Reb[5] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
    Reb[6] = -sin(theta);
    Reb[7] = cos(theta) * sin(phi);
    Reb[8] = cos(theta) * cos(phi);
    const double v2[3] = {VX_COV * VX_COV, VY_COV * VY_COV, VZ_COV * VZ_COV};
    const double w2[3] = {WX_COV * WX_COV, WY_COV * WY_COV, WZ_COV * WZ_COV};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) t33b[3 * a + b] = Reb[3 * a + b] * v2[b];  // Reb diag(v^2)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += t33b[3 * a + k] * Reb[3 * b + k];  // * Reb^T
        t33[3 * a + b] = s;
      }
    const double dt3 = DT * DT * DT, dt2 = DT * DT;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        Q[12 * a + b] = dt3 * t33[3 * a + b];
        Q[12 * a + b + 6] = dt2 * t33[3 * a + b];
        Q[12 * (b + 6) + a] = dt2 * t33[3 * a + b];
        Q[12 * (a + 6) + b + 6] = DT * t33[3 * a + b];
        const double wv = (a == b) ? w2[a] : 0.0;
        Q[12 * (a + 3) + b + 3] = dt3 * wv;
        Q[12 * (a + 3) + b + 9] = dt2 * wv;
        Q[12 * (b + 9) + a + 3] = dt2 * wv;
        Q[12 * (a + 9) + b + 9] = DT * wv;
      }
  }
  __syncthreads();
  double s = 0.0;
  for (int k = 0; k < 12; ++k) s += F[12 * i + k] * Prr[12 * k + j];
  if (act) Tm[t] = s;
  __syncthreads();
  s = 0.0;
  for (int k = 0; k < 12; ++k) s += Tm[12 * i + k] * F[12 * j + k];
  if (act) P[(size_t)i * ld + j] = s + Q[t];
  if (t < 6) x[t] += x[t + 6] * DT;
}

__global__ void __launch_bounds__(256) ekf_predict_kernel(double* __restrict__ P, int ld, int N, double* __restrict__ x) {
  if (blockIdx.x == 0) ekf_predict_pose(P, ld, x);
  else ekf_predict_cross(P, ld, N, (int)blockIdx.x - 1);
}

// P = scale * G G^T + diag * I (benchmark-size synthetic covariances)
__global__ void __launch_bounds__(256) ekf_lowrank_kernel(double* __restrict__ P, int ld, int N, const double* __restrict__ G,
                                                          int r, double scale, double dg) {
  __shared__ double Ga[16][65], Gb[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int k0 = 0; k0 < r; k0 += 64) {
    const int kn = min(64, r - k0);
    for (int e = threadIdx.x; e < 16 * kn; e += 256) {
      const int rr = e / kn, kk = e % kn;
      const int gi = blockIdx.y * 16 + rr, gj = blockIdx.x * 16 + rr;
      Ga[rr][kk] = (gi < N) ? G[(size_t)gi * r + k0 + kk] : 0.0;
      Gb[rr][kk] = (gj < N) ? G[(size_t)gj * r + k0 + kk] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < kn; ++k) acc += Ga[ty][k] * Gb[tx][k];
    __syncthreads();
  }
  if (i < N && j < N) {
    // symmetric by construction: both (i,j) and (j,i) sum the same products in the same order
    P[(size_t)i * ld + j] = scale * acc + ((i == j) ? dg : 0.0);
  }
}

// ---- state augmentation ("next" rows): in place on the device-resident P ------------------------------------
// The reference re-allocates and copies all of P (O(N^2)) for every added landmark (kalmanClass.cpp:362-389,
// :517-535); here the new rows/columns are appended inside the capacity: O(N) work.
constexpr double VREAL = 1.0;

__global__ void ekf_first_states_kernel(double* __restrict__ x, double* __restrict__ P, int ld, const double* z19) {
  // AddFirstStates, kalmanClass.cpp:275-311 (the 28 x 28 block was zeroed by the caller)
  const int i = threadIdx.x;
  if (i >= 28) return;
  const double pose_diag[12] = {0.0000001, 0.0000001, 0.1, 0.1, 0.1, 0.0000001, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1};
  double xv = 0.0;
  if (i >= 12) { xv = z19[i - 12]; P[(size_t)i * ld + i] = MEAS_COV * MEAS_COV; }
  else P[(size_t)i * ld + i] = pose_diag[i];
  if (i == 2) xv = z19[16];
  if (i == 3) xv = z19[17];
  if (i == 4) xv = z19[18];
  if (i == 6) xv = VREAL;
  x[i] = xv;
}

// inverse of a 4x4 by Gauss-Jordan with partial pivoting (the reference's cvInvert(A, Ainv, CV_SVD) of a
// well-conditioned split matrix)
__device__ void inv4(const double* A, double* Ai) {
  double m[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { m[i][j] = A[4 * i + j]; m[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 4; ++c) {
    int p = c;
    for (int r = c + 1; r < 4; ++r)
      if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
    for (int j = 0; j < 8; ++j) { const double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
    const double inv = 1.0 / m[c][c];
    for (int j = 0; j < 8; ++j) m[c][j] *= inv;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        const double f = m[r][c];
        for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
      }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) Ai[4 * i + j] = m[i][4 + j];
}

// AddNewCurve, kalmanClass.cpp:314-499.  in = z (8) | A (16).  One block: thread 0 builds the small matrices
// in shared memory, then all threads fill the new rows/columns: P[N+i][j] = (Gx P[0:12, j]) mirrored.
__global__ void __launch_bounds__(256) ekf_add_curve_kernel(double* __restrict__ x, double* __restrict__ P, int ld,
                                                            int N, const double* __restrict__ in) {
  __shared__ double Gx[8][12], Gz[8][8], xnew[8], Prr[12][12];
  const int t = threadIdx.x;
  {
    if (t < 144) Prr[t / 12][t % 12] = P[(size_t)(t / 12) * ld + t % 12];
    if (t == 0) {
      const double* z = in;
      const double* A = in + 8;
      const double Tx = x[0], Ty = x[1], psi = x[5];
      const double c = cos(psi), s_ = sin(psi);
      double Ainv[16], rz[8], rdz[8];
      inv4(A, Ainv);
      // Rot = [[c I, -s I], [s I, c I]] (the transpose of the update's Rot, :333-343); Rotderiv likewise
      for (int i = 0; i < 4; ++i) {
        rz[i] = (c * z[i] - s_ * z[i + 4]) + Tx;
        rz[i + 4] = (s_ * z[i] + c * z[i + 4]) + Ty;
        rdz[i] = -s_ * z[i] - c * z[i + 4];
        rdz[i + 4] = c * z[i] - s_ * z[i + 4];
      }
      for (int i = 0; i < 4; ++i) {
        double sx = 0.0, sy = 0.0, dx = 0.0, dy = 0.0, so = 0.0;
        for (int j = 0; j < 4; ++j) {
          sx += Ainv[4 * i + j] * rz[j];
          sy += Ainv[4 * i + j] * rz[j + 4];
          dx += Ainv[4 * i + j] * rdz[j];
          dy += Ainv[4 * i + j] * rdz[j + 4];
          so += Ainv[4 * i + j];
        }
        xnew[i] = sx; xnew[i + 4] = sy;
        for (int j = 0; j < 12; ++j) { Gx[i][j] = 0.0; Gx[i + 4][j] = 0.0; }
        Gx[i][0] = so; Gx[i + 4][1] = so; Gx[i][5] = dx; Gx[i + 4][5] = dy;
        for (int j = 0; j < 4; ++j) {  // Gz = Hinv Rot
          Gz[i][j] = Ainv[4 * i + j] * c;      Gz[i][j + 4] = -Ainv[4 * i + j] * s_;
          Gz[i + 4][j] = Ainv[4 * i + j] * s_; Gz[i + 4][j + 4] = Ainv[4 * i + j] * c;
        }
      }
    }
    __syncthreads();
  }
  // cross blocks: new rows x all existing columns j in [0, N): Gx * P[0:12, j]  (for j < 12 the reference uses
  // Gx * Prr^T, i.e. column j of Prr^T = row j of Prr)
  for (int j = blockIdx.x * blockDim.x + t; j < N; j += gridDim.x * blockDim.x) {
    double col[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) col[k] = (j < 12) ? Prr[j][k] : P[(size_t)k * ld + j];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double s_ = 0.0;
#pragma unroll
      for (int k = 0; k < 12; ++k) s_ += Gx[i][k] * col[k];
      P[(size_t)(N + i) * ld + j] = s_;
      P[(size_t)j * ld + N + i] = s_;
    }
  }
  if (blockIdx.x == 0 && t < 64) {  // PN1N1 = Gx Prr Gx^T + Gz (MEAS_COV I) Gz^T   (R1 is not squared, :95-98)
    const int i = t / 8, j = t % 8;
    double acc = 0.0;
    for (int a = 0; a < 12; ++a) {
      double inner = 0.0;
      for (int b = 0; b < 12; ++b) inner += Prr[a][b] * Gx[j][b];
      acc += Gx[i][a] * inner;
    }
    double gz = 0.0;
    for (int a = 0; a < 8; ++a) gz += Gz[i][a] * (MEAS_COV * Gz[j][a]);
    P[(size_t)(N + i) * ld + N + j] = acc + gz;
    if (t < 8) x[N + t] = xnew[t];
  }
}

// AddNewPoints, kalmanClass.cpp:502-666: grid.y = point, blocks along x cover the existing columns
__global__ void __launch_bounds__(256) ekf_add_points_kernel(double* __restrict__ x, double* __restrict__ P, int ld,
                                                             int N, const double* __restrict__ meas, int n_pts) {
  __shared__ double Gx[3][12], Reb[9], Prr[12][12], xnew[3];
  const int t = threadIdx.x, n = blockIdx.y;
  const int o = N + 3 * n;
  if (t < 144) Prr[t / 12][t % 12] = P[(size_t)(t / 12) * ld + t % 12];
  if (t == 0) {
    const double phi = x[3], theta = x[4], psi = x[5];
    double Rd[3][9];
    // generate_Reb common.h:189-201 and get_Reb_derivs kalmanClass.cpp:1230-1262
    Reb[0] = cos(psi) * cos(theta);
    Reb[1] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
    Reb[2] = cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi);
    Reb[3] = sin(psi) * cos(theta);
    Reb[4] = sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi);
    Reb[5] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
    Reb[6] = -sin(theta);
    Reb[7] = cos(theta) * sin(phi);
    Reb[8] = cos(theta) * cos(phi);
    double* Rp = Rd[0]; double* Rt = Rd[1]; double* Rs = Rd[2];
    Rp[0] = 0.0; Rp[1] = cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi); Rp[2] = -cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
    Rp[3] = 0.0; Rp[4] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi); Rp[5] = -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi);
    Rp[6] = 0.0; Rp[7] = cos(theta) * cos(phi); Rp[8] = -cos(theta) * sin(phi);
    Rt[0] = -cos(psi) * sin(theta); Rt[1] = cos(psi) * cos(theta) * sin(phi); Rt[2] = cos(psi) * cos(theta) * cos(phi);
    Rt[3] = -sin(psi) * sin(theta); Rt[4] = sin(psi) * cos(theta) * sin(phi); Rt[5] = sin(psi) * cos(theta) * cos(phi);
    Rt[6] = -cos(theta); Rt[7] = -sin(theta) * sin(phi); Rt[8] = -sin(theta) * cos(phi);
    Rs[0] = -sin(psi) * cos(theta); Rs[1] = -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi); Rs[2] = -sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
    Rs[3] = cos(psi) * cos(theta); Rs[4] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi); Rs[5] = cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi);
    Rs[6] = 0.0; Rs[7] = 0.0; Rs[8] = 0.0;
    const double* xb = meas + 3 * n;
    for (int i = 0; i < 3; ++i) {
      double s_ = 0.0;
      for (int k = 0; k < 3; ++k) s_ += Reb[3 * i + k] * xb[k];
      xnew[i] = s_ + x[i];
      for (int j = 0; j < 12; ++j) Gx[i][j] = 0.0;
      Gx[i][i] = 1.0;
      for (int j = 0; j < 3; ++j) {
        double d = 0.0;
        for (int k = 0; k < 3; ++k) d += Rd[j][3 * i + k] * xb[k];
        Gx[i][3 + j] = d;
      }
    }
  }
  __syncthreads();
  // cross block with the states that existed before the call: j < 12 uses Gx * Prr (column j of Prr)
  for (int j = blockIdx.x * blockDim.x + t; j < N; j += gridDim.x * blockDim.x) {
    double col[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) col[k] = (j < 12) ? Prr[k][j] : P[(size_t)k * ld + j];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double s_ = 0.0;
#pragma unroll
      for (int k = 0; k < 12; ++k) s_ += Gx[i][k] * col[k];
      P[(size_t)(o + i) * ld + j] = s_;
      P[(size_t)j * ld + o + i] = s_;
    }
  }
  if (blockIdx.x == 0 && t < 9) {  // PN1N1 = Gx Prr Gx^T + Reb (PT_MEAS_COV I) Reb^T  (R_pts is not squared, :99-102)
    const int i = t / 3, j = t % 3;
    double acc = 0.0;
    for (int a = 0; a < 12; ++a) {
      double inner = 0.0;
      for (int b = 0; b < 12; ++b) inner += Prr[a][b] * Gx[j][b];
      acc += Gx[i][a] * inner;
    }
    double gz = 0.0;
    for (int a = 0; a < 3; ++a) gz += Reb[3 * i + a] * (PT_MEAS_COV * Reb[3 * j + a]);
    P[(size_t)(o + i) * ld + o + j] = acc + gz;
    if (t < 3) x[o + t] = xnew[t];
  }
}

int round_up(int v, int q) { return (v + q - 1) / q * q; }

}  // namespace

// =================================== C ABI ===================================
extern "C" void cps_ekf_destroy(cps_ekf* h);
namespace {
// grow-only scratch: nothing is allocated per call once the handle has seen its largest request
int ensure_aux(cps_ekf* h, size_t bytes) {
  if (bytes <= h->aux_bytes) return CPS_OK;
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_aux);
  h->d_aux = nullptr;
  h->aux_bytes = 0;
  CPS_CUDA(cudaMalloc(&h->d_aux, bytes));
  h->aux_bytes = bytes;
  return CPS_OK;
}

// Rebuilds the tiles below the diagonal tiles from their mirror images if the last updates left them unwritten.
int ensure_lower(cps_ekf* h) {
  if (!h->lower_stale) return CPS_OK;
  const int nt = (h->N + TILE - 1) / TILE;
  if (nt > 1) {
    ekf_mirror_lower_kernel<<<dim3(nt, nt), 256, 0, h->stream>>>(h->d_P, h->cap);
    CPS_CUDA(cudaGetLastError());
    h->launches++;
  }
  h->lower_stale = false;
  return CPS_OK;
}

constexpr int PP_MAX_M = 68;  // seventeen k-steps of K fragments in registers
size_t pp_smem_bytes(int Mp) {
  const int nst = Mp <= PIPE_MAX_M ? 4 : 3;
  return 1024 + (size_t)nst * TMA_TILE_BYTES + (size_t)nst * Mp * DSTRIDE * sizeof(double) + 128;
}

size_t tma_smem_bytes(int Mp, int stages) {
  return 1024 + (size_t)(1 + stages) * TMA_TILE_BYTES + (size_t)(stages + 1) * Mp * DSTRIDE * sizeof(double) + 128;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
bool encode_maps(cps_ekf* h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  EncodeFn encode = (EncodeFn)fn;
  const cuuint64_t dims[2] = {(cuuint64_t)h->cap, (cuuint64_t)h->cap};           // innermost first
  const cuuint64_t strides[1] = {(cuuint64_t)h->cap * sizeof(double)};           // bytes between rows
  const cuuint32_t estr[2] = {1, 1};
  const cuuint32_t box_full[2] = {TILE, TILE}, box_swz[2] = {16, TILE};
  if (encode(&h->map_full, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, h->d_P, dims, strides, box_full, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (encode(&h->map_swz, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, h->d_P, dims, strides, box_swz, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  return true;
}

int ekf_create_impl(cps_ekf* h, int device, int capacity) {
  h->device = device;
  h->cap = round_up(capacity, TILE);
  CPS_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
  CPS_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const size_t cap = (size_t)h->cap;
  if (cudaMalloc(&h->d_P, cap * cap * sizeof(double)) != cudaSuccess) {
    cps::set_last_error("cudaMalloc of the covariance failed");
    cudaGetLastError();
    return CPS_ERR_NOMEM;
  }
  CPS_CUDA(cudaMalloc(&h->d_x, cap * sizeof(double)));
  CPS_CUDA(cudaMalloc(&h->d_T, cap * MAXM * sizeof(double)));
  CPS_CUDA(cudaMalloc(&h->d_Kt, cap * MAXM * sizeof(double)));
  CPS_CUDA(cudaMalloc(&h->d_Yt, cap * MAXM * sizeof(double)));
  CPS_CUDA(cudaMalloc(&h->d_small, SM_TOTAL * sizeof(double)));
  CPS_CUDA(cudaMalloc(&h->d_cols, (MAXNC + MAXM) * sizeof(int)));
  CPS_CUDA(cudaMalloc(&h->d_flag, sizeof(int)));
  h->pinned_bytes = (MAXM + 16 * (MAXM / 8 + 1)) * sizeof(double) + MAXM * sizeof(int);
  CPS_CUDA(cudaMallocHost(&h->h_pinned, h->pinned_bytes));
  CPS_CUDA(cudaMemsetAsync(h->d_P, 0, cap * cap * sizeof(double), h->stream));
  CPS_CUDA(cudaMemsetAsync(h->d_x, 0, cap * sizeof(double), h->stream));
  CPS_CUDA(cudaMemsetAsync(h->d_Kt, 0, cap * MAXM * sizeof(double), h->stream));
  CPS_CUDA(cudaMemsetAsync(h->d_Yt, 0, cap * MAXM * sizeof(double), h->stream));
  CPS_CUDA(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
  for (auto& e : h->ev) CPS_CUDA(cudaEventCreate(&e));
  for (auto& e : h->gev) CPS_CUDA(cudaEventCreate(&e));
  CPS_CUDA(cudaFuncSetAttribute(ekf_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                200 * 1024));
  CPS_CUDA(cudaFuncSetAttribute(ekf_gain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (2 * MAXM * MAXM + MAXM + 16 * MAXM) * (int)sizeof(double)));
  CPS_CUDA(cudaFuncSetAttribute(ekf_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                2 * MAXM * TILE * (int)sizeof(double)));
  CPS_CUDA(cudaFuncSetAttribute(ekf_cov_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (2 * (TILE * PITCH_A + TILE * PITCH_B + PIPE_MAX_M * DSTRIDE) + PIPE_MAX_M * DSTRIDE) *
                                    (int)sizeof(double)));
  CPS_CUDA(cudaFuncSetAttribute(ekf_cov_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                2 * MAXM * DSTRIDE * (int)sizeof(double)));
  CPS_CUDA(cudaFuncSetAttribute(ekf_pht_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (MAXM * MAXNC + 8 * MAXNC) * (int)sizeof(double) + MAXNC * (int)sizeof(int)));
  CPS_CUDA(cudaFuncSetAttribute(ekf_front_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrontSmemMax));
  CPS_CUDA(cudaFuncSetAttribute(ekf_front_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrontSmemMax));
  CPS_CUDA(cudaFuncSetAttribute(ekf_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrontSmemMax));
  CPS_CUDA(cudaFuncSetAttribute(ekf_cov_tma_kernel<TMA_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                tma_smem_bytes(PIPE_MAX_M, TMA_STAGES)));
  CPS_CUDA(cudaFuncSetAttribute(ekf_cov_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                tma_smem_bytes(TMA2_MAX_M, 2)));
  CPS_CUDA(cudaFuncSetAttribute(ekf_cov_pp_kernel<PIPE_MAX_M / 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                pp_smem_bytes(PIPE_MAX_M)));
  CPS_CUDA(cudaFuncSetAttribute(ekf_cov_pp_kernel<PP_MAX_M / 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                pp_smem_bytes(PP_MAX_M)));
  h->maps_ok = encode_maps(h);
  h->mirror_exact = true;  // P = 0
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  return CPS_OK;
}
}  // namespace

extern "C" int cps_ekf_create(cps_ekf** out, int device, int capacity) {
  if (!out || capacity < CPS_ROBOT_STATE_SIZE) return CPS_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cps::set_last_error("no CUDA device: libcps has no CPU path");
    return CPS_ERR_CUDA;
  }
  if (device < 0 || device >= count) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(device));
  cps_ekf* h = new cps_ekf();
  const int rc = ekf_create_impl(h, device, capacity);
  if (rc != CPS_OK) {  // release whatever was allocated before the failure
    cps_ekf_destroy(h);
    return rc;
  }
  *out = h;
  return CPS_OK;
}


extern "C" void cps_ekf_destroy(cps_ekf* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_P); cudaFree(h->d_x); cudaFree(h->d_T); cudaFree(h->d_Kt); cudaFree(h->d_Yt);
  cudaFree(h->d_small); cudaFree(h->d_cols); cudaFree(h->d_flag); cudaFree(h->d_aux);
  cudaFree(h->d_pack); cudaFree(h->d_pcols); cudaFree(h->d_gz); cudaFree(h->d_gidx); cudaFree(h->d_gd2);
  cudaFree(h->d_part_d2); cudaFree(h->d_part_idx);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->gev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int cps_ekf_sync(cps_ekf* h) {
  if (!h) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  int flag = 0;
  CPS_CUDA(cudaMemcpy(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    cps::set_last_error("innovation covariance S is not positive definite");
    CPS_CUDA(cudaMemset(h->d_flag, 0, sizeof(int)));
    return CPS_ERR_INVALID;
  }
  return CPS_OK;
}

extern "C" int cps_ekf_set_state(cps_ekf* h, const double* x, const double* P, int N, const int* curve_inds,
                                 int num_curves, const int* point_inds, int num_points) {
  if (!h || !x || N < CPS_ROBOT_STATE_SIZE || N > h->cap || num_curves < 0 || num_points < 0) return CPS_ERR_INVALID;
  if ((num_curves && !curve_inds) || (num_points && !point_inds)) return CPS_ERR_INVALID;
  for (int k = 0; k < num_curves; ++k)
    if (curve_inds[k] < 0 || curve_inds[k] + 8 > N) return CPS_ERR_INVALID;
  for (int k = 0; k < num_points; ++k)
    if (point_inds[k] < 0 || point_inds[k] + 3 > N) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  const size_t cap = (size_t)h->cap;
  CPS_CUDA(cudaMemsetAsync(h->d_P, 0, cap * cap * sizeof(double), h->stream));
  CPS_CUDA(cudaMemsetAsync(h->d_x, 0, cap * sizeof(double), h->stream));
  CPS_CUDA(cudaMemcpyAsync(h->d_x, x, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
  if (P)
    CPS_CUDA(cudaMemcpy2DAsync(h->d_P, cap * sizeof(double), P, (size_t)N * sizeof(double), (size_t)N * sizeof(double),
                               N, cudaMemcpyHostToDevice, h->stream));
  h->N = N;
  h->mirror_exact = (P == nullptr);  // a caller-supplied covariance need not be symmetric to the last bit
  h->lower_stale = false;
  h->curve_inds.assign(curve_inds, curve_inds + num_curves);
  h->point_inds.assign(point_inds, point_inds + num_points);
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  return CPS_OK;
}

extern "C" int cps_ekf_set_cov_lowrank(cps_ekf* h, const double* G, int r, double scale, double diag) {
  if (!h || !G || r <= 0 || h->N <= 0) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  { const int rc = ensure_aux(h, sizeof(double) * (size_t)h->N * r); if (rc != CPS_OK) return rc; }
  double* dG = h->d_aux;
  CPS_CUDA(cudaMemcpyAsync(dG, G, sizeof(double) * (size_t)h->N * r, cudaMemcpyHostToDevice, h->stream));
  dim3 grid((h->N + 15) / 16, (h->N + 15) / 16);
  ekf_lowrank_kernel<<<grid, 256, 0, h->stream>>>(h->d_P, h->cap, h->N, dG, r, scale, diag);
  CPS_CUDA(cudaGetLastError());
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  h->mirror_exact = true;  // (i,j) and (j,i) sum the same products in the same order
  h->lower_stale = false;
  return CPS_OK;
}

extern "C" int cps_ekf_dims(cps_ekf* h, int* N, int* num_curves, int* num_points) {
  if (!h) return CPS_ERR_INVALID;
  if (N) *N = h->N;
  if (num_curves) *num_curves = (int)h->curve_inds.size();
  if (num_points) *num_points = (int)h->point_inds.size();
  return CPS_OK;
}

extern "C" int cps_ekf_get_state(cps_ekf* h, double* x_out) {
  if (!h || !x_out) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  CPS_CUDA(cudaMemcpyAsync(x_out, h->d_x, sizeof(double) * h->N, cudaMemcpyDeviceToHost, h->stream));
  return cps_ekf_sync(h);
}

extern "C" int cps_ekf_get_cov_block(cps_ekf* h, int r0, int c0, int nr, int nc, double* out) {
  if (!h || !out || r0 < 0 || c0 < 0 || nr < 0 || nc < 0 || r0 + nr > h->N || c0 + nc > h->N) return CPS_ERR_INVALID;
  if (nr == 0 || nc == 0) return CPS_OK;
  CPS_CUDA(cudaSetDevice(h->device));
  // a block that reaches below the diagonal tiles while they are stale: a small one is gathered from the upper copies
  // (a per-frame read of a landmark's 3 x 3 block must not cost a pass over P), a large one has them rebuilt first
  if (h->lower_stale && (r0 + nr - 1) / TILE > c0 / TILE) {
    if ((long long)nr * nc <= (1LL << 20)) {
      { const int rc = ensure_aux(h, sizeof(double) * (size_t)nr * nc); if (rc != CPS_OK) return rc; }
      const int grid = (int)std::min<long long>(((long long)nr * nc + 255) / 256, 1024);
      ekf_gather_block_kernel<<<grid, 256, 0, h->stream>>>(h->d_P, h->cap, r0, c0, nr, nc, h->d_aux);
      CPS_CUDA(cudaGetLastError());
      h->launches++;
      CPS_CUDA(cudaMemcpyAsync(out, h->d_aux, sizeof(double) * (size_t)nr * nc, cudaMemcpyDeviceToHost, h->stream));
      return cps_ekf_sync(h);
    }
    const int rc = ensure_lower(h);
    if (rc != CPS_OK) return rc;
  }
  CPS_CUDA(cudaMemcpy2DAsync(out, (size_t)nc * sizeof(double), h->d_P + (size_t)r0 * h->cap + c0,
                             (size_t)h->cap * sizeof(double), (size_t)nc * sizeof(double), nr, cudaMemcpyDeviceToHost,
                             h->stream));
  return cps_ekf_sync(h);
}

extern "C" int cps_ekf_get_cov(cps_ekf* h, double* P_out) {
  if (!h) return CPS_ERR_INVALID;
  return cps_ekf_get_cov_block(h, 0, 0, h->N, h->N, P_out);
}

extern "C" int cps_ekf_predict(cps_ekf* h) {
  if (!h || h->N < CPS_ROBOT_STATE_SIZE) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  CPS_CUDA(cudaEventRecord(h->ev[0], h->stream));
  // the pose block reads x before it is advanced; the cross block does not read x
  ekf_predict_kernel<<<1 + (h->N - 12 + 255) / 256, 256, 0, h->stream>>>(h->d_P, h->cap, h->N, h->d_x);
  h->launches++;
  CPS_CUDA(cudaGetLastError());
  CPS_CUDA(cudaEventRecord(h->ev[1], h->stream));
  CPS_CUDA(cudaEventRecord(h->ev[2], h->stream));
  CPS_CUDA(cudaEventRecord(h->ev[3], h->stream));
  h->timing_valid = true;
  return CPS_OK;
}

namespace {
int run_update(cps_ekf* h, const double* z, int n, const double* A, const int* curve_num, const double* point_meas,
               const int* point_nums, int n_pts, bool curves_call) {
  const int M = (curves_call ? 8 * n + 3 : 0) + 3 * n_pts;
  if (M <= 0 || M > MAXM) {
    cps::set_last_error("measurement vector empty or larger than CPS_EKF_MAX_MEAS");
    return CPS_ERR_INVALID;
  }
  for (int k = 0; k < n; ++k)
    if (curve_num[k] < 0 || curve_num[k] >= (int)h->curve_inds.size()) return CPS_ERR_INVALID;
  for (int k = 0; k < n_pts; ++k)
    if (point_nums[k] < 0 || point_nums[k] >= (int)h->point_inds.size()) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  // the previous update's staging must have been consumed before the pinned block is rewritten
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  double* hz = (double*)h->h_pinned;
  double* hA = hz + MAXM;
  int* hcols = (int*)(hA + 16 * (MAXM / 8 + 1));
  std::memset(hz, 0, sizeof(double) * MAXM);
  if (curves_call) {
    std::memcpy(hz, z, sizeof(double) * (8 * n + 3));
    std::memcpy(hA, A, sizeof(double) * 16 * n);
  }
  if (n_pts) std::memcpy(hz + (curves_call ? 8 * n + 3 : 0), point_meas, sizeof(double) * 3 * n_pts);
  for (int k = 0; k < n; ++k) hcols[k] = h->curve_inds[curve_num[k]];
  for (int k = 0; k < n_pts; ++k) hcols[n + k] = h->point_inds[point_nums[k]];
  UpdateDesc d;
  d.N = h->N; d.ld = h->cap; d.n = n; d.n_pts = n_pts; d.M = M; d.nc = 6 + 8 * n + 3 * n_pts;
  d.sym = h->mirror_exact ? 1 : 0;
  cudaStream_t st = h->stream;
  CPS_CUDA(cudaMemcpyAsync(h->d_small + SM_IN, hz, sizeof(double) * (MAXM + 16 * (size_t)n), cudaMemcpyHostToDevice, st));
  int* d_land = h->d_cols + MAXNC;
  if (n + n_pts)
    CPS_CUDA(cudaMemcpyAsync(d_land, hcols, sizeof(int) * (n + n_pts), cudaMemcpyHostToDevice, st));
  CPS_CUDA(cudaEventRecord(h->ev[0], st));
  const int row_blocks = std::min((d.N + 7) / 8, h->sm_count * 8);
  const int ncp = d.nc | 1, Mo = d.M | 1;  // odd shared-memory pitches of the kernels' H and S tables
  const size_t sh_front = (size_t)(d.M * ncp + d.M * d.nc + d.nc * d.nc + 3 * d.M * d.M) * sizeof(double);
  // rows kernel: eight warps per block while two blocks fit an SM, sixteen in one block above that (M > 36)
  auto rows_smem = [&](int nw) {
    return (size_t)(d.M * ncp + d.M * d.M + d.M * Mo + d.M + nw * ROWS_R * (d.nc + 2 * d.M)) * sizeof(double) + d.nc * sizeof(int);
  };
  const int rows_warps = 2 * (rows_smem(8) + 1024) <= (size_t)227 * 1024 ? 8 : 16;
  const size_t sh_rows = rows_smem(rows_warps);
  const int rows_per_sm = rows_warps == 8 ? std::max(1, std::min(8, (int)((size_t)227 * 1024 / (sh_rows + 1024)))) : 1;
  const bool no_fused = std::getenv("CPS_EKF_NO_FUSED_FRONT") != nullptr;  // A/B switch (read per call: tests flip it)
  int launches = 0;
  if (!no_fused && sh_front <= (size_t)kFrontSmemMax && sh_rows <= (size_t)kFrontSmemMax) {
    if (d.M * d.M <= 6 * 256) ekf_front_kernel<256><<<1, 256, sh_front, st>>>(d, h->d_x, h->d_P, h->d_small, h->d_cols, d_land, h->d_flag);
    else ekf_front_kernel<512><<<1, 512, sh_front, st>>>(d, h->d_x, h->d_P, h->d_small, h->d_cols, d_land, h->d_flag);
    const int rows_grid = std::min((d.N + rows_warps * ROWS_R - 1) / (rows_warps * ROWS_R), h->sm_count * rows_per_sm);
    ekf_rows_kernel<<<rows_grid, 32 * rows_warps, sh_rows, st>>>(d, h->d_P, h->d_small, h->d_cols, h->d_Kt, h->d_Yt, h->d_x, h->cap);
    launches = 2;
  } else {
    ekf_build_kernel<<<1, 128, 0, st>>>(d, h->d_x, h->d_small, h->d_cols, d_land);
    const size_t sh_pht = (size_t)(d.M * d.nc + 8 * d.nc) * sizeof(double) + d.nc * sizeof(int);
    ekf_pht_kernel<<<row_blocks, 256, sh_pht, st>>>(d, h->d_P, h->d_small, h->d_cols, h->d_T);
    {
      const size_t base_b = 2 * (size_t)M * M * sizeof(double), stage_b = 2 * (size_t)M * d.nc * sizeof(double);
      const int staged = (base_b + stage_b <= 200 * 1024) ? 1 : 0;
      ekf_s_kernel<<<1, 256, base_b + (staged ? stage_b : 0), st>>>(d, h->d_T, h->d_small, h->d_cols, h->d_flag, staged);
    }
    const size_t sh_gain = (size_t)(2 * M * M + M + 16 * M) * sizeof(double);
    ekf_gain_kernel<<<row_blocks, 256, sh_gain, st>>>(d, h->d_T, h->d_small, h->d_Kt, h->d_Yt, h->d_x, h->cap);
    launches = 4;
  }
  CPS_CUDA(cudaEventRecord(h->ev[2], st));
  const int nt = (d.N + TILE - 1) / TILE;
  const long long pairs = (long long)nt * (nt + 1) / 2;
  static const bool use_cuda_cores = std::getenv("CPS_EKF_COV_CUDA_CORES") != nullptr;  // A/B switch for profiling
  // half-written mode (default): the streaming kernel leaves the tiles below the diagonal tiles alone
  const bool lazy_lower = std::getenv("CPS_EKF_EAGER_LOWER") == nullptr;
  if (use_cuda_cores) {
    { const int rc = ensure_lower(h); if (rc != CPS_OK) return rc; }
    ekf_cov_kernel<<<(unsigned)pairs, 256, 2 * (size_t)M * TILE * sizeof(double), st>>>(h->d_P, h->cap, M, h->d_Kt,
                                                                                        h->d_Yt, h->cap, nt);
  } else {
    const int Mp = (M + 3) & ~3;
    const bool no_pipe = std::getenv("CPS_EKF_COV_NO_PIPE") != nullptr;
    const bool no_tma = std::getenv("CPS_EKF_COV_NO_TMA") != nullptr;
    if (Mp <= TMA2_MAX_M && !no_pipe && !no_tma && h->maps_ok && h->mirror_exact && pairs >= 2LL * h->sm_count) {
      TmaMaps maps;
      maps.full = h->map_full;
      maps.swz = h->map_swz;
      const int write_lower = lazy_lower ? 0 : 1;
      const bool no_pp = std::getenv("CPS_EKF_COV_NO_PINGPONG") != nullptr;  // A/B switch
      if (Mp <= PIPE_MAX_M && lazy_lower && !no_pp)
        ekf_cov_pp_kernel<PIPE_MAX_M / 4, 4><<<h->sm_count, TMA_THREADS, pp_smem_bytes(Mp), st>>>(maps, M, h->d_Kt, h->d_Yt, h->cap, nt, pairs);
      else if (Mp <= PP_MAX_M && lazy_lower && !no_pp)
        ekf_cov_pp_kernel<PP_MAX_M / 4, 3><<<h->sm_count, TMA_THREADS, pp_smem_bytes(Mp), st>>>(maps, M, h->d_Kt, h->d_Yt, h->cap, nt, pairs);
      else if (Mp <= PIPE_MAX_M)
        ekf_cov_tma_kernel<TMA_STAGES><<<h->sm_count, TMA_THREADS, tma_smem_bytes(Mp, TMA_STAGES), st>>>(maps, M, h->d_Kt, h->d_Yt, h->cap, nt, pairs, write_lower);
      else
        ekf_cov_tma_kernel<2><<<h->sm_count, TMA_THREADS, tma_smem_bytes(Mp, 2), st>>>(maps, M, h->d_Kt, h->d_Yt, h->cap, nt, pairs, write_lower);
      if (!write_lower) h->lower_stale = true;
    } else if (Mp <= PIPE_MAX_M && !no_pipe && pairs >= 2LL * h->sm_count) {
      { const int rc = ensure_lower(h); if (rc != CPS_OK) return rc; }
      const size_t shb = (size_t)(2 * (TILE * PITCH_A + TILE * PITCH_B + Mp * DSTRIDE) + Mp * DSTRIDE) * sizeof(double);
      ekf_cov_pipe_kernel<<<h->sm_count, 256, shb, st>>>(h->d_P, h->cap, M, h->d_Kt, h->d_Yt, h->cap, nt, pairs);
    } else {
      { const int rc = ensure_lower(h); if (rc != CPS_OK) return rc; }
      ekf_cov_dmma_kernel<<<(unsigned)pairs, 256, 2 * (size_t)Mp * DSTRIDE * sizeof(double), st>>>(
          h->d_P, h->cap, M, h->d_Kt, h->d_Yt, h->cap);
    }
  }
  CPS_CUDA(cudaEventRecord(h->ev[3], st));
  CPS_CUDA(cudaEventRecord(h->ev[1], st));
  CPS_CUDA(cudaGetLastError());
  h->launches += launches + 1;
  h->last_update_launches = launches + 1;
  h->mirror_exact = true;  // every tile pair was written from one value
  h->timing_valid = true;
  return CPS_OK;
}
}  // namespace

extern "C" int cps_ekf_update_curves_points(cps_ekf* h, const double* z, int n, const double* A, const int* curve_num,
                                            const double* point_meas, const int* point_nums, int n_pts) {
  if (!h || !z || n < 0 || n_pts < 0 || (n && (!A || !curve_num)) || (n_pts && (!point_meas || !point_nums)))
    return CPS_ERR_INVALID;
  return run_update(h, z, n, A, curve_num, point_meas, point_nums, n_pts, true);
}

extern "C" int cps_ekf_update_points(cps_ekf* h, const double* point_meas, const int* point_nums, int n_pts) {
  if (!h || n_pts <= 0 || !point_meas || !point_nums) return CPS_ERR_INVALID;
  return run_update(h, nullptr, 0, nullptr, nullptr, point_meas, point_nums, n_pts, false);
}

extern "C" int cps_ekf_last_timing(cps_ekf* h, float* total_ms, float* cov_kernel_ms, long long* launches) {
  if (!h) return CPS_ERR_INVALID;
  if (launches) *launches = h->launches;
  if (!h->timing_valid) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  CPS_CUDA(cudaEventSynchronize(h->ev[1]));
  if (total_ms) CPS_CUDA(cudaEventElapsedTime(total_ms, h->ev[0], h->ev[1]));
  if (cov_kernel_ms) CPS_CUDA(cudaEventElapsedTime(cov_kernel_ms, h->ev[2], h->ev[3]));
  return CPS_OK;
}


// ---- state augmentation ------------------------------------------------------------------------------------
extern "C" int cps_ekf_add_first_states(cps_ekf* h, const double* z19) {
  if (!h || !z19 || h->cap < 28) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  std::memcpy(h->h_pinned, z19, 19 * sizeof(double));
  const size_t cap = (size_t)h->cap;
  CPS_CUDA(cudaMemsetAsync(h->d_P, 0, cap * cap * sizeof(double), h->stream));  // cvSetZero(P), cvSetZero(x)
  CPS_CUDA(cudaMemsetAsync(h->d_x, 0, cap * sizeof(double), h->stream));
  CPS_CUDA(cudaMemcpyAsync(h->d_small + SM_IN, h->h_pinned, 19 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  ekf_first_states_kernel<<<1, 32, 0, h->stream>>>(h->d_x, h->d_P, h->cap, h->d_small + SM_IN);
  CPS_CUDA(cudaGetLastError());
  h->launches++;
  h->N = 28;
  h->mirror_exact = true;  // a diagonal matrix
  h->lower_stale = false;  // all of P was rewritten
  h->curve_inds = {CPS_ROBOT_STATE_SIZE, CPS_ROBOT_STATE_SIZE + 8};  // kalmanClass.cpp:307-309
  h->point_inds.clear();
  return CPS_OK;
}

extern "C" int cps_ekf_add_new_curve(cps_ekf* h, const double* z8, const double* A) {
  if (!h || !z8 || !A || h->N < CPS_ROBOT_STATE_SIZE) return CPS_ERR_INVALID;
  if (h->N + 8 > h->cap) {
    cps::set_last_error("state capacity exhausted (cps_ekf_create capacity)");
    return CPS_ERR_NOMEM;
  }
  CPS_CUDA(cudaSetDevice(h->device));
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  double* hp = (double*)h->h_pinned;
  std::memcpy(hp, z8, 8 * sizeof(double));
  std::memcpy(hp + 8, A, 16 * sizeof(double));
  CPS_CUDA(cudaMemcpyAsync(h->d_small + SM_IN, hp, 24 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  // reads touch rows < 12 and columns < N, writes touch rows >= N or columns >= N: blocks are independent
  ekf_add_curve_kernel<<<std::max(1, std::min(64, (h->N + 255) / 256)), 256, 0, h->stream>>>(h->d_x, h->d_P, h->cap, h->N,
                                                                                     h->d_small + SM_IN);
  CPS_CUDA(cudaGetLastError());
  h->launches++;
  h->mirror_exact = false;  // the new 8 x 8 diagonal block may straddle two tiles and is symmetric only to rounding
  h->curve_inds.push_back(h->N);
  h->N += 8;
  return CPS_OK;
}

extern "C" int cps_ekf_add_new_points(cps_ekf* h, const double* meas, int n_pts) {
  if (!h || n_pts < 0 || (n_pts && !meas) || h->N < CPS_ROBOT_STATE_SIZE) return CPS_ERR_INVALID;
  if (n_pts == 0) return CPS_OK;
  if (h->N + 3 * n_pts > h->cap) {
    cps::set_last_error("state capacity exhausted (cps_ekf_create capacity)");
    return CPS_ERR_NOMEM;
  }
  CPS_CUDA(cudaSetDevice(h->device));
  { const int rc = ensure_aux(h, sizeof(double) * 3 * (size_t)n_pts); if (rc != CPS_OK) return rc; }
  double* d_meas = h->d_aux;
  CPS_CUDA(cudaMemcpyAsync(d_meas, meas, sizeof(double) * 3 * (size_t)n_pts, cudaMemcpyHostToDevice, h->stream));
  // every point reads only the pose rows/columns of the covariance as it was before the call and writes its own
  // new rows/columns (the reference gives points added together no cross-covariance), except the mirrored pose
  // columns P[j][o+i], j < 12, which no other point reads: one launch, grid.y = point
  ekf_add_points_kernel<<<dim3(std::max(1, std::min(32, (h->N + 255) / 256)), n_pts), 256, 0, h->stream>>>(
      h->d_x, h->d_P, h->cap, h->N, d_meas, n_pts);
  CPS_CUDA(cudaGetLastError());
  CPS_CUDA(cudaStreamSynchronize(h->stream));  // `meas` is the caller's (possibly pageable) buffer
  h->launches++;
  h->mirror_exact = false;  // the new 3 x 3 diagonal blocks may straddle two tiles and are symmetric only to rounding
  for (int n = 0; n < n_pts; ++n) h->point_inds.push_back(h->N + 3 * n);
  h->N += 3 * n_pts;
  return CPS_OK;
}

extern "C" int cps_ekf_get_index_tables(cps_ekf* h, int* curve_inds, int* point_inds) {
  if (!h) return CPS_ERR_INVALID;
  if (curve_inds) std::memcpy(curve_inds, h->curve_inds.data(), sizeof(int) * h->curve_inds.size());
  if (point_inds) std::memcpy(point_inds, h->point_inds.data(), sizeof(int) * h->point_inds.size());
  return CPS_OK;
}

// GetSplitMatrices, kalmanClass.cpp:1144-1159
extern "C" void cps_ekf_split_matrices(double t, double* A1, double* A2) {
  static const double binom[4][4] = {{1, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 1, 0}, {1, 3, 3, 1}};
  for (int i = 0; i < 16; ++i) { A1[i] = 0.0; A2[i] = 0.0; }
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j <= i; ++j) A1[i * 4 + j] = binom[i][j] * std::pow(t, j) * std::pow((1 - t), (i - j));
    for (int j = i; j < 4; ++j) A2[4 * i + j] = binom[3 - i][j - i] * std::pow(t, (j - i)) * std::pow((1 - t), (3 - j));
  }
}
