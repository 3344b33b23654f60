// cps_gate.cu -- Mahalanobis gate, parallel over (measurement, landmark) pairs (sm_100a, FP64, -fmad=false).
//
// The reference has no gating code; the gate is defined from its point-landmark measurement
// model (kalmanClass.cpp:815-851, R_pts :975-984) in oracle/orc_gate.c and this file states the same
// arithmetic for the device: a per-landmark pack (predicted measurement and the 6 unique entries of S_j^-1,
// one thread per landmark) followed by the pair sweep (one thread per measurement and landmark chunk, packs
// staged through shared memory and read as broadcasts), then a tiny reduction over chunks.  Index results
// are bit-exact with the CPU definition: only + - * / in a fixed order, no FMA (this file is compiled with
// -fmad=false), trigonometric values computed once on the host with the same libm expressions as the CPU
// side and passed in as kernel arguments.
#include <cmath>
#include <limits>

#include "../../include/cps_ekf.h"
#include "cps_common.h"
#include "cps_ekf_internal.h"

namespace {

constexpr double PT_MEAS_COV = 5.0;  // kalmanClass.h:18
constexpr int GATE_THREADS = 128;
constexpr int GATE_TILE = 256;  // landmarks staged per shared-memory tile

struct RotArgs {
  double R[9];
  double Rd[3][9];
};

__global__ void __launch_bounds__(256) gate_pack_kernel(const double* __restrict__ x, const double* __restrict__ P,
                                                        int ld, const int* __restrict__ pcols, int n_land,
                                                        RotArgs rot, double* __restrict__ pack, int upper_only) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_land) return;
  const int c = pcols[j];
  double xe[3], H[3][9], HP[3][9], S[3][3];
  int cols[9];
  for (int a = 0; a < 3; ++a) xe[a] = x[c + a] - x[a];
  for (int a = 0; a < 3; ++a) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += rot.R[3 * a + k] * xe[k];
    pack[9 * (size_t)j + a] = s;
    for (int b = 0; b < 3; ++b) {
      double d = 0.0;
      H[a][b] = -rot.R[3 * a + b];
      H[a][6 + b] = rot.R[3 * a + b];
      for (int k = 0; k < 3; ++k) d += rot.Rd[b][3 * a + k] * xe[k];
      H[a][3 + b] = d;
    }
  }
  for (int a = 0; a < 6; ++a) cols[a] = a;
  for (int a = 0; a < 3; ++a) cols[6 + a] = c + a;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 9; ++b) {
      double s = 0.0;
      // upper_only (cps_ekf::lower_stale): the entry with row <= column is the current one and equals its mirror image
      for (int k = 0; k < 9; ++k) {
        const int r = cols[k], q = cols[b];
        s += H[a][k] * P[(upper_only && r > q) ? (size_t)q * ld + r : (size_t)r * ld + q];
      }
      HP[a][b] = s;
    }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double s = 0.0;
      for (int k = 0; k < 9; ++k) s += HP[a][k] * H[b][k];
      S[a][b] = s;
    }
  for (int a = 0; a < 3; ++a) S[a][a] += PT_MEAS_COV * PT_MEAS_COV;
  const double s00 = S[0][0], s11 = S[1][1], s22 = S[2][2];
  const double s01 = 0.5 * (S[0][1] + S[1][0]), s02 = 0.5 * (S[0][2] + S[2][0]), s12 = 0.5 * (S[1][2] + S[2][1]);
  const double c00 = s11 * s22 - s12 * s12;
  const double c01 = s02 * s12 - s01 * s22;
  const double c02 = s01 * s12 - s02 * s11;
  const double c11 = s00 * s22 - s02 * s02;
  const double c12 = s01 * s02 - s00 * s12;
  const double c22 = s00 * s11 - s01 * s01;
  const double det = (s00 * c00 + s01 * c01) + s02 * c02;
  const double idet = 1.0 / det;
  double* q = pack + 9 * (size_t)j;
  q[3] = c00 * idet; q[4] = c01 * idet; q[5] = c02 * idet;
  q[6] = c11 * idet; q[7] = c12 * idet; q[8] = c22 * idet;
}

// grid (measurement blocks, landmark chunks); partial best per (chunk, measurement).  Every thread carries
// GATE_MPT measurements in registers, so the 9 broadcast loads of a landmark pack feed GATE_MPT * 23 flops (with one
// measurement per thread the sweep was limited by shared-memory wavefronts, not by the FP64 pipe).
// GATE_MPT = 8 for large sweeps (10k x 20k: 0.64 of the unfused FP64 peak against 0.60 with 4), 4 when the problem
// would not fill the GPU with the coarser grid.
template <int GATE_MPT>
__global__ void __launch_bounds__(GATE_THREADS) gate_sweep_kernel(const double* __restrict__ pack, int n_land,
                                                                  const double* __restrict__ z, int nz, double gate,
                                                                  int chunk, double* __restrict__ part_d2,
                                                                  int* __restrict__ part_idx) {
  __shared__ double tile[GATE_TILE * 9];
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * GATE_MPT;
  const int jlo = blockIdx.y * chunk;
  const int jhi = min(n_land, jlo + chunk);
  double z0[GATE_MPT], z1[GATE_MPT], z2[GATE_MPT], best[GATE_MPT];
  int bi[GATE_MPT];
#pragma unroll
  for (int q = 0; q < GATE_MPT; ++q) {
    const int i = i0 + q;
    z0[q] = z1[q] = z2[q] = 0.0;
    if (i < nz) { z0[q] = z[3 * (size_t)i]; z1[q] = z[3 * (size_t)i + 1]; z2[q] = z[3 * (size_t)i + 2]; }
    best[q] = gate;
    bi[q] = -1;
  }
  for (int j0 = jlo; j0 < jhi; j0 += GATE_TILE) {
    const int cnt = min(GATE_TILE, jhi - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 9; e += blockDim.x) tile[e] = pack[9 * (size_t)j0 + e];
    __syncthreads();
    for (int jj = 0; jj < cnt; ++jj) {
      const double* q9 = tile + 9 * jj;
      const double c0 = q9[0], c1 = q9[1], c2 = q9[2], s00 = q9[3], s01 = q9[4], s02 = q9[5], s11 = q9[6],
                   s12 = q9[7], s22 = q9[8];
#pragma unroll
      for (int q = 0; q < GATE_MPT; ++q) {
        const double n0 = z0[q] - c0, n1 = z1[q] - c1, n2 = z2[q] - c2;
        const double t0 = (s00 * n0 + s01 * n1) + s02 * n2;
        const double t1 = (s01 * n0 + s11 * n1) + s12 * n2;
        const double t2 = (s02 * n0 + s12 * n1) + s22 * n2;
        const double d2 = (n0 * t0 + n1 * t1) + n2 * t2;
        if (d2 < best[q]) { best[q] = d2; bi[q] = j0 + jj; }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < GATE_MPT; ++q) {
    const int i = i0 + q;
    if (i < nz) {
      part_d2[(size_t)blockIdx.y * nz + i] = best[q];
      part_idx[(size_t)blockIdx.y * nz + i] = bi[q];
    }
  }
}

__global__ void gate_reduce_kernel(const double* __restrict__ part_d2, const int* __restrict__ part_idx, int nchunks,
                                   int nz, double gate, int* __restrict__ idx, double* __restrict__ d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nz) return;
  double best = gate;
  int bi = -1;
  for (int c = 0; c < nchunks; ++c) {  // chunks ascend in j: strict "<" keeps the lowest index on ties
    const int ci = part_idx[(size_t)c * nz + i];
    const double cd = part_d2[(size_t)c * nz + i];
    if (ci >= 0 && cd < best) { best = cd; bi = ci; }
  }
  idx[i] = bi;
  if (d2) d2[i] = (bi >= 0) ? best : __longlong_as_double(0x7ff0000000000000LL);
}

// same expressions as oracle/orc_ekf.c (generate_Rbe common.h:174-186, get_Rbe_derivs kalmanClass.cpp:1275-1298)
void host_rot(double phi, double theta, double psi, RotArgs& r) {
  using std::cos;
  using std::sin;
  double* R = r.R;
  R[0] = cos(psi) * cos(theta);
  R[3] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
  R[6] = cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi);
  R[1] = sin(psi) * cos(theta);
  R[4] = sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi);
  R[7] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
  R[2] = -sin(theta);
  R[5] = cos(theta) * sin(phi);
  R[8] = cos(theta) * cos(phi);
  double *Rp = r.Rd[0], *Rt = r.Rd[1], *Rs = r.Rd[2];
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
  Rt[4] = sin(psi) * cos(theta) * cos(phi);  // the reference writes (1,1) twice; (2,1) stays 0
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

}  // namespace

extern "C" int cps_gate_mahalanobis_dev(cps_ekf* h, const double* d_z, int nz, double gate, int* d_idx, double* d_d2) {
  if (!h || nz < 0 || (nz && (!d_z || !d_idx))) return CPS_ERR_INVALID;
  if (nz == 0) return CPS_OK;
  const int n_land = (int)h->point_inds.size();
  CPS_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // pose angles -> rotation tables on the host (libm), exactly as the CPU definition does
  double pose[6];
  CPS_CUDA(cudaMemcpyAsync(pose, h->d_x, sizeof pose, cudaMemcpyDeviceToHost, st));
  CPS_CUDA(cudaStreamSynchronize(st));
  RotArgs rot;
  host_rot(pose[3], pose[4], pose[5], rot);
  if ((size_t)n_land > h->pack_cap) {
    cudaFree(h->d_pack);
    cudaFree(h->d_pcols);
    h->d_pack = nullptr; h->d_pcols = nullptr; h->pack_cap = 0;
    CPS_CUDA(cudaMalloc(&h->d_pack, sizeof(double) * 9 * (size_t)std::max(n_land, 1)));
    CPS_CUDA(cudaMalloc(&h->d_pcols, sizeof(int) * (size_t)std::max(n_land, 1)));
    h->pack_cap = (size_t)n_land;
  }
  const int mpt = ((long long)nz * n_land >= 100000000LL) ? 8 : 4;
  const int row_blocks = (nz + GATE_THREADS * mpt - 1) / (GATE_THREADS * mpt);
  const int reduce_blocks = (nz + GATE_THREADS - 1) / GATE_THREADS;
  int nchunks = std::max(1, (4 * h->sm_count + row_blocks - 1) / row_blocks);
  nchunks = std::min(nchunks, std::max(1, (n_land + GATE_TILE - 1) / GATE_TILE));
  const int chunk = std::max(1, (n_land + nchunks - 1) / nchunks);
  nchunks = std::max(1, (n_land + chunk - 1) / chunk);
  const size_t need = (size_t)nchunks * nz;
  if (need > h->part_cap) {
    cudaFree(h->d_part_d2);
    cudaFree(h->d_part_idx);
    h->d_part_d2 = nullptr; h->d_part_idx = nullptr; h->part_cap = 0;
    CPS_CUDA(cudaMalloc(&h->d_part_d2, sizeof(double) * need));
    CPS_CUDA(cudaMalloc(&h->d_part_idx, sizeof(int) * need));
    h->part_cap = need;
  }
  if (n_land)
    CPS_CUDA(cudaMemcpyAsync(h->d_pcols, h->point_inds.data(), sizeof(int) * n_land, cudaMemcpyHostToDevice, st));
  CPS_CUDA(cudaEventRecord(h->gev[0], st));
  if (n_land) {
    gate_pack_kernel<<<(n_land + 255) / 256, 256, 0, st>>>(h->d_x, h->d_P, h->cap, h->d_pcols, n_land, rot, h->d_pack,
                                                           h->lower_stale ? 1 : 0);
    h->launches++;
  }
  CPS_CUDA(cudaEventRecord(h->gev[1], st));
  if (mpt == 8)
    gate_sweep_kernel<8><<<dim3(row_blocks, nchunks), GATE_THREADS, 0, st>>>(h->d_pack, n_land, d_z, nz, gate, chunk,
                                                                             h->d_part_d2, h->d_part_idx);
  else
    gate_sweep_kernel<4><<<dim3(row_blocks, nchunks), GATE_THREADS, 0, st>>>(h->d_pack, n_land, d_z, nz, gate, chunk,
                                                                             h->d_part_d2, h->d_part_idx);
  gate_reduce_kernel<<<reduce_blocks, GATE_THREADS, 0, st>>>(h->d_part_d2, h->d_part_idx, nchunks, nz, gate, d_idx, d_d2);
  h->launches += 2;
  CPS_CUDA(cudaEventRecord(h->gev[2], st));
  CPS_CUDA(cudaGetLastError());
  h->gate_timing_valid = true;
  return CPS_OK;
}

extern "C" int cps_gate_mahalanobis(cps_ekf* h, const double* z, int nz, double gate, int* idx_out, double* d2_out) {
  if (!h || nz < 0 || (nz && (!z || !idx_out))) return CPS_ERR_INVALID;
  if (nz == 0) return CPS_OK;
  CPS_CUDA(cudaSetDevice(h->device));
  if ((size_t)nz > h->gate_cap) {
    cudaFree(h->d_gz); cudaFree(h->d_gidx); cudaFree(h->d_gd2);
    h->d_gz = nullptr; h->d_gidx = nullptr; h->d_gd2 = nullptr; h->gate_cap = 0;
    CPS_CUDA(cudaMalloc(&h->d_gz, sizeof(double) * 3 * (size_t)nz));
    CPS_CUDA(cudaMalloc(&h->d_gidx, sizeof(int) * (size_t)nz));
    CPS_CUDA(cudaMalloc(&h->d_gd2, sizeof(double) * (size_t)nz));
    h->gate_cap = (size_t)nz;
  }
  CPS_CUDA(cudaMemcpyAsync(h->d_gz, z, sizeof(double) * 3 * (size_t)nz, cudaMemcpyHostToDevice, h->stream));
  int rc = cps_gate_mahalanobis_dev(h, h->d_gz, nz, gate, h->d_gidx, h->d_gd2);
  if (rc != CPS_OK) return rc;
  CPS_CUDA(cudaMemcpyAsync(idx_out, h->d_gidx, sizeof(int) * (size_t)nz, cudaMemcpyDeviceToHost, h->stream));
  if (d2_out) CPS_CUDA(cudaMemcpyAsync(d2_out, h->d_gd2, sizeof(double) * (size_t)nz, cudaMemcpyDeviceToHost, h->stream));
  CPS_CUDA(cudaStreamSynchronize(h->stream));
  return CPS_OK;
}

extern "C" int cps_gate_last_timing(cps_ekf* h, float* pack_ms, float* sweep_ms) {
  if (!h || !h->gate_timing_valid) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(h->device));
  CPS_CUDA(cudaEventSynchronize(h->gev[2]));
  if (pack_ms) CPS_CUDA(cudaEventElapsedTime(pack_ms, h->gev[0], h->gev[1]));
  if (sweep_ms) CPS_CUDA(cudaEventElapsedTime(sweep_ms, h->gev[1], h->gev[2]));
  return CPS_OK;
}
