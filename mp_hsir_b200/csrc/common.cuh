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

// Branch-free erf GELU for the tensor-core epilogues: Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, i.e. at
// fp32 rounding level), 2 MUFU + ~12 FMA-pipe instructions instead of erff's two-branch polynomial.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float pl = fmaf(1.061405429f, t, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float erf_abs = fmaf(-pl * t, e, 1.0f);  // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
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
