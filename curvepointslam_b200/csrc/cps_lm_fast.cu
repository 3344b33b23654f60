// cps_lm_fast.cu -- the OPTIONAL fused-arithmetic build of the staged dlevmar_der kernels (cps_lm_set_arith).
//
// Everything else in libcps.so is compiled with -fmad=false and performs the reference's floating-point operation
// sequence: results are the reference levmar's bit for bit, at the price of issuing every multiply and add
// separately (18.5 of the 33.5 TFLOP/s the FP64 pipe can deliver).  This translation unit compiles the SAME kernel
// sources once more, in their own namespace, with multiply-add contraction on and with CPS_FMA selecting the
// fused-accumulation form of the J^T J passes.  Its results are NOT bit-identical to the reference: they differ from
// it the way two CPU builds of the reference differ from each other (the distribution is measured by
// tests/test_lm_staged_gpu.py and reported by bench.py).  Off by default; nothing here is used unless
// cps_lm_set_arith(ctx, CPS_ARITH_FMA) was called.
#define CPS_FMA 1
#define cps cps_fast  // every `namespace cps` / `cps::` of the kernel headers below names cps_fast in this unit
#include "cps_lm_staged.cuh"
#undef cps

namespace {
bool g_attr_done = false;
}

// launchers the host code of cps_lm_api.cu calls (kernel arguments by value: StagedArgs has the same layout in both
// namespaces -- it is the same declaration)
extern "C" int cps_fast_prepare(int jac_smem, int eval_smem) {
  if (g_attr_done) return 0;
  using namespace cps_fast;
  if (cudaFuncSetAttribute(staged_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(staged_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(staged_jac_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, jac_smem) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(staged_jac_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, jac_smem) != cudaSuccess) return -1;
  g_attr_done = true;
  return 0;
}

extern "C" void cps_fast_launch_jac(int two_threads, int grid, int block, int smem, cudaStream_t st, const void* args, int nxt) {
  const cps_fast::StagedArgs& A = *reinterpret_cast<const cps_fast::StagedArgs*>(args);
  if (two_threads) cps_fast::staged_jac_kernel<true><<<grid, block, smem, st>>>(A, nxt);
  else cps_fast::staged_jac_kernel<false><<<grid, block, smem, st>>>(A, nxt);
}

extern "C" void cps_fast_launch_eval(int init, int grid, int block, int smem, cudaStream_t st, const void* args, int nxt) {
  const cps_fast::StagedArgs& A = *reinterpret_cast<const cps_fast::StagedArgs*>(args);
  if (init) cps_fast::staged_eval_kernel<true><<<grid, block, smem, st>>>(A, nxt);
  else cps_fast::staged_eval_kernel<false><<<grid, block, smem, st>>>(A, nxt);
}
