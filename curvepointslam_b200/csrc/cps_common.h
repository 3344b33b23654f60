// cps_common.h -- host-side helpers shared by the translation units of libcps.so
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

namespace cps {
void set_last_error(const std::string& s);
inline bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
  return false;
}
}  // namespace cps
#define CPS_CUDA(call)                                  \
  do {                                                  \
    if (!cps::cuda_ok((call), #call)) return CPS_ERR_CUDA; \
  } while (0)
