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

// ---- cta_group::2: a CTA pair issues M = 256 instructions (128 rows per CTA); each CTA holds its rows of A and HALF the
// rows of B; only the leader CTA issues, the commit is multicast to both.
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
               "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma2_bf16_tmem_a(uint32_t tmem_d, uint32_t a_taddr, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
               "r"(a_taddr), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc2(uint32_t n) {   // M = 256
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma_rate2_kernel(int N, int a_tmem, int iters, Res* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u ^ ((i * 2654435761u) & 0x007F007Fu);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x < 32) {
    long long t0 = clock64();
    if (rank == 0) {
      const uint32_t idesc = make_idesc2(N);
      const uint32_t a_addr = smem_u32(smem);
      const uint32_t b_addr = smem_u32(smem + 32 * 1024);
      for (int it = 0; it < iters; ++it) {
        const uint64_t ad = make_desc(a_addr + (it & 1) * 16384);
        const uint64_t bd = make_desc(b_addr + (it & 3) * (N / 2 * 128));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (a_tmem) umma2_bf16_tmem_a(tmem_base, tmem_base + 448 + 8 * k, bd + 2 * k, idesc);
            else umma2_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc);
          }
        }
        __syncwarp();
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].clk = t1 - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
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
  cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("cta_group::2 (M = 256 per instruction, a CTA pair):\n");
  for (int grid : {2, sms})
    for (int N : {64, 128, 256})
      for (int a_tmem : {0, 1}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        mma_rate2_kernel<<<grid, 128, 200 * 1024>>>(N, a_tmem, 100, d);
        cudaEventRecord(e0);
        mma_rate2_kernel<<<grid, 128, 200 * 1024>>>(N, a_tmem, iters, d);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
        double clk = 0;
        for (int i = 0; i < grid; i += 2) clk += h[i].clk;
        clk /= (grid / 2);
        const double per = clk / (iters * 4.0);
        const double tf = 2.0 * 256 * N * 16 * iters * 4.0 * (grid / 2) / (ms * 1e-3) / 1e12;
        printf("%5d %6s | %10.1f clk/MMA  %8.1f clk per 128x128x16-equivalent %10.1f TFLOP/s  grid=%d  %.3f ms\n", N, a_tmem ? "tmem" : "smem",
               per, per * 128.0 / N / 2.0, tf, grid, ms);
      }
  return 0;
}
