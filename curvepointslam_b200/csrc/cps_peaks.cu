// cps_peaks.cu -- FP64 roofline denominators measured live on the device (MEASURED_PEAKS.json has HBM
// and bf16 entries only).  Two micro-kernels: dependent-chain-free DFMA streams (the FP64 CUDA-core FMA
// peak, 2 flop per instruction) and the same with separate DMUL + DADD (what a kernel that may not fuse,
// like the bit-exact LM path, can reach at best: 1 flop per instruction).
#include "../../include/cps_lm.h"
#include "cps_common.h"

namespace {

template <bool FMA>
__global__ void __launch_bounds__(256) fp64_stream_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    if (FMA) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    } else {
      x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b);
      x2 = __dadd_rn(__dmul_rn(x2, a), b); x3 = __dadd_rn(__dmul_rn(x3, a), b);
      x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
      x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// FP64 tensor-core stream: independent mma.sync.m8n8k4.f64 accumulations (512 flop per warp instruction)
__global__ void __launch_bounds__(256) dmma_stream_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = c[i][0] + 0.5; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool FMA>
int run(double* d_out, int grid, int iters, double* tflops) {
  cudaEvent_t e0, e1;
  CPS_CUDA(cudaEventCreate(&e0));
  CPS_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CPS_CUDA(cudaEventRecord(e0));
    fp64_stream_kernel<FMA><<<grid, 256>>>(d_out, iters, 0.999999, 1e-7);
    CPS_CUDA(cudaEventRecord(e1));
    CPS_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CPS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = (double)grid * 256 * (double)iters * 8 * 2;  // mul+add (fused or not) = 2 flop
  *tflops = flops / (best * 1e-3) / 1e12;
  return CPS_OK;
}

}  // namespace

// tflops_fma: DFMA stream; tflops_muladd: DMUL+DADD stream (both count 2 flop per multiply-add pair)
extern "C" int cps_measure_fp64_peak(int device, double* tflops_fma, double* tflops_muladd) {
  if (!tflops_fma || !tflops_muladd) return CPS_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cps::set_last_error("no CUDA device");
    return CPS_ERR_CUDA;
  }
  CPS_CUDA(cudaSetDevice(device));
  int sms = 0;
  CPS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int grid = sms * 8;
  double* d_out = nullptr;
  CPS_CUDA(cudaMalloc(&d_out, sizeof(double) * (size_t)grid * 256));
  int rc = run<true>(d_out, grid, 4096, tflops_fma);
  if (rc == CPS_OK) rc = run<false>(d_out, grid, 4096, tflops_muladd);
  cudaFree(d_out);
  return rc;
}

// FP64 tensor-core (DMMA, mma.sync.m8n8k4) peak in TFLOP/s
extern "C" int cps_measure_dmma_peak(int device, double* tflops) {
  if (!tflops) return CPS_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cps::set_last_error("no CUDA device");
    return CPS_ERR_CUDA;
  }
  CPS_CUDA(cudaSetDevice(device));
  int sms = 0;
  CPS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int grid = sms * 8, iters = 2048;
  double* d_out = nullptr;
  CPS_CUDA(cudaMalloc(&d_out, sizeof(double) * (size_t)grid * 256));
  cudaEvent_t e0, e1;
  CPS_CUDA(cudaEventCreate(&e0));
  CPS_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CPS_CUDA(cudaEventRecord(e0));
    dmma_stream_kernel<<<grid, 256>>>(d_out, iters, 0.999999, 1e-7);
    CPS_CUDA(cudaEventRecord(e1));
    CPS_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CPS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = (double)grid * 8 /*warps*/ * (double)iters * 8 * 512.0 / (best * 1e-3) / 1e12;
  return CPS_OK;
}
