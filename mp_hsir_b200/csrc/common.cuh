// Shared helpers for libmphsir.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mphsir.h"

namespace mphsir {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

__device__ __forceinline__ float gelu_erf(float x) {
  // exact erf GELU: nn.GELU() default / F.gelu (net/MP_HSIR.py:67,:263,:389)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace mphsir

#define MPHSIR_REQUIRE(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      mphsir::set_error(__VA_ARGS__);      \
      return MPHSIR_ERR_INVALID;           \
    }                                      \
  } while (0)
