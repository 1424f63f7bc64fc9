// Micro-benchmark: sustained tcgen05.mma rate of ONE CTA per SM (cta_group::1, kind::f16, bf16 operands, M = 128, K = 16)
// as a function of N, of where the A operand lives (shared memory descriptor / tensor memory) and of how many
// accumulators consecutive instructions rotate over.  No other shared-memory traffic runs beside the MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mp_hsir_b200/csrc -o tools/mma_rate tools/mma_rate.cu && tools/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace mphsir::tc;

struct Res { long long clk; };

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int a_tmem, int nacc, int iters, int bslots, Res* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  // fill the operand area with bf16 values around 1 (power draw of real data, no NaN / denormal special cases)
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u ^ ((i * 2654435761u) & 0x007F007Fu);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x < 32) {
    const uint32_t idesc = make_idesc(N);
    const uint32_t a_addr = smem_u32(smem);                 // 16 KB per A slab (128 rows x 128 B)
    const uint32_t b_addr = smem_u32(smem + 32 * 1024);     // B slabs of N rows x 128 B, `bslots` of them
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t ad = make_desc(a_addr + (it & 1) * 16384);
      const uint64_t bd = make_desc(b_addr + (it % bslots) * (N * 128));
      const uint32_t d = tmem_base + (nacc > 1 ? (it % nacc) * N : 0);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (a_tmem) umma_bf16_tmem_a(d, tmem_base + 448 + 8 * k, bd + 2 * k, idesc, 1);
          else umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].clk = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  Res* d;
  cudaMalloc(&d, sizeof(Res) * sms);
  Res* h = new Res[sms];
  const int iters = 4000;  // x 4 MMAs
  printf("%5s %6s %5s %7s | %10s %12s %10s\n", "N", "A", "nacc", "bslots", "clk/MMA", "clk/128cols", "TFLOP/s");
  for (int grid : {1, sms})
    for (int N : {64, 128, 192, 256})
      for (int a_tmem : {0, 1})
        for (int nacc : {1, 2}) {
          if (nacc * N > 384) continue;
          for (int bslots : {1, 4}) {
            if (bslots * N * 128 > 160 * 1024) continue;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            mma_rate_kernel<<<grid, 128, 200 * 1024>>>(N, a_tmem, nacc, 100, bslots, d);  // warm
            cudaEventRecord(e0);
            mma_rate_kernel<<<grid, 128, 200 * 1024>>>(N, a_tmem, nacc, iters, bslots, d);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(h, d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
            double clk = 0;
            for (int i = 0; i < grid; ++i) clk += h[i].clk;
            clk /= grid;
            const double per = clk / (iters * 4.0);
            const double tf = 2.0 * 128 * N * 16 * iters * 4.0 * grid / (ms * 1e-3) / 1e12;
            printf("%5d %6s %5d %7d | %10.1f %12.1f %10.1f   grid=%d  %.3f ms\n", N, a_tmem ? "tmem" : "smem", nacc, bslots, per,
                   per * 128.0 / N, tf, grid, ms);
          }
        }
  return 0;
}
