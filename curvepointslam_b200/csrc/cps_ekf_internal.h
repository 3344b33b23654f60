// cps_ekf_internal.h -- the EKF handle shared by cps_ekf.cu (update kernels) and cps_gate.cu (gate kernels)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>

struct cps_ekf {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  int N = 0;    // current state dimension
  int cap = 0;  // capacity, a multiple of 64; leading dimension of P
  double* d_x = nullptr;   // [cap]
  double* d_P = nullptr;   // [cap][cap] row-major; rows/cols >= N are kept zero
  std::vector<int> curve_inds, point_inds;
  // per-update device work space
  double* d_T = nullptr;    // [cap][MAXM]   P H^T (row-major, leading dimension MAXM)
  double* d_Kt = nullptr;   // [MAXM][cap]   K transposed
  double* d_Yt = nullptr;   // [MAXM][cap]   (K S^T) transposed
  double* d_small = nullptr;  // H (compact), S, S^-1, innovation, R, measurement staging
  int* d_cols = nullptr;
  int* d_flag = nullptr;    // 0 = ok, 1 = S not positive definite
  char* h_pinned = nullptr; // host staging of one update's inputs
  double* d_aux = nullptr;  // grow-only device scratch (new point measurements, the low-rank factor)
  size_t aux_bytes = 0;
  size_t pinned_bytes = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  long long launches = 0;
  bool timing_valid = false;
  int last_update_launches = 0;
  // true while every off-diagonal 64 x 64 tile of P is the exact mirror image of its partner (cps_ekf.cu, kernel 5d)
  bool mirror_exact = true;
  // true while the tiles BELOW the diagonal tiles have not been rewritten by the last update(s): the streaming kernel
  // then wrote only tile (I,J), I <= J, of every pair.  Nothing on the predict / update / gate / augmentation path reads
  // them; ensure_lower() (cps_ekf.cu) rebuilds them for the entry points that do.
  bool lower_stale = false;
  // TMA descriptors of P (encoded once: d_P and cap never change)
  alignas(64) CUtensorMap map_full;
  alignas(64) CUtensorMap map_swz;
  bool maps_ok = false;
  // gate work space
  double* d_pack = nullptr;  // [num_points][9]
  int* d_pcols = nullptr;    // [num_points]
  size_t pack_cap = 0;
  double* d_gz = nullptr;
  int* d_gidx = nullptr;
  double* d_gd2 = nullptr;
  double* d_part_d2 = nullptr;
  int* d_part_idx = nullptr;
  size_t gate_cap = 0, part_cap = 0;
  cudaEvent_t gev[3] = {nullptr, nullptr, nullptr};
  bool gate_timing_valid = false;
};
