// Fused depthwise-3x3 conv + Gram statistics, TMA-fed variant of dwgram.cu (same math, same partial layout).
//
// dwgram.cu reads every 3x6 input patch straight from global memory, so each CTA alternates between a
// load-latency phase and a tensor-core phase (ncu: 41% long-scoreboard + 30% barrier stalls, ~2 TB/s).
// Here one producer thread streams the haloed tile of each operand -- a [10 x 10 pixels x CG channels]
// fp32 box of the 1x1-conv output, described by a 4-D tensor map whose out-of-bounds zero fill IS the
// conv padding -- into a shared-memory ring, running up to NST boxes ahead of the 8 consumer warps:
//
//   per (8x8 tile, head group):  q box -> conv -> bf16 hi/lo tile Qs   |  sum q^2 kept in registers
//                                k box -> conv -> bf16 hi/lo tile Ks   |  sum k^2 kept in registers
//                                bar; Gram q^T k on the tensor cores (mma.sync, accumulators persist over tiles)
//                                v box -> conv -> V (HBM);  bar
//
// The conv reads shared memory only (2x4-pixel register blocks for 128-channel groups: 24 LDS.128 per 8
// outputs), the squared norms come from the fp32 conv results each thread already holds (its (block, quad)
// assignment is the same for every tile, so they stay in registers until the end), and the global loads
// never stall a warp.  HBM traffic per token is unchanged: read 3C (+ halo, L2 hits), write C floats.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mphsir {
namespace dwg {

using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

__device__ __forceinline__ void t_ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void t_mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void t_split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void consumer_bar() {
  __syncwarp();  // bar.sync is .aligned: reconverge after the lane-0 mbarrier arrives
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

constexpr int TMA_CONSUMERS = 256;
constexpr int TMA_THREADS = TMA_CONSUMERS + 32;
constexpr int HALO = 10;  // 8 + 2

template <int CG, int PARTS>
struct TmaPlan {
  static constexpr int NST = CG == 128 ? 2 : 4;             // ring stages
  static constexpr int STAGE_BYTES = HALO * HALO * CG * 4;  // one operand box
  static constexpr int LD = CG + 8;
  static constexpr int ARR = 64 * LD;
  static constexpr int QK_BYTES = 2 * PARTS * ARR * 2;
  static size_t smem(int C) { return (size_t)NST * STAGE_BYTES + QK_BYTES + (size_t)9 * 3 * C * 4 + 128; }
};

// depthwise conv of a BR x 4 pixel block x 4 channels out of the haloed box in shared memory
template <int CG, int BR>
__device__ __forceinline__ void conv_block(const float* __restrict__ box, const float4 (&w)[9], int by, int bx, int quad,
                                           float4 (&acc)[BR][4]) {
#pragma unroll
  for (int r = 0; r < BR; ++r)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[r][o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int ir = 0; ir < BR + 2; ++ir) {  // input row (halo coordinates by + ir)
    float4 v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] = *reinterpret_cast<const float4*>(box + ((by + ir) * HALO + bx + j) * CG + quad * 4);
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const int dy = ir - r;  // kernel row feeding output row r
      if (dy < 0 || dy > 2) continue;
      const float4 w0 = w[3 * dy], w1 = w[3 * dy + 1], w2 = w[3 * dy + 2];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        acc[r][o].x = fmaf(v[o].x, w0.x, fmaf(v[o + 1].x, w1.x, fmaf(v[o + 2].x, w2.x, acc[r][o].x)));
        acc[r][o].y = fmaf(v[o].y, w0.y, fmaf(v[o + 1].y, w1.y, fmaf(v[o + 2].y, w2.y, acc[r][o].y)));
        acc[r][o].z = fmaf(v[o].z, w0.z, fmaf(v[o + 1].z, w1.z, fmaf(v[o + 2].z, w2.z, acc[r][o].z)));
        acc[r][o].w = fmaf(v[o].w, w0.w, fmaf(v[o + 1].w, w1.w, fmaf(v[o + 2].w, w2.w, acc[r][o].w)));
      }
    }
  }
}

// CG: channels per head group (64 or 128), CH: channels per head, NG: head groups (C = NG*CG)
template <int CG, int CH, int NG, int PARTS>
__global__ void __launch_bounds__(TMA_THREADS, 1) dwgram_tma_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                    const float* __restrict__ w9, float* __restrict__ V,
                                                                    long long ldv, float* __restrict__ partial, int H, int W,
                                                                    int tiles_x, int tiles_per_sample, int gy0, int gy1) {
  // gy0, gy1: only tiles whose first row lies in [gy0, gy1) enter the Gram statistics (see dwgram.cu)
  using P = TmaPlan<CG, PARTS>;
  constexpr int C = CG * NG;
  constexpr int NST = P::NST, LD = P::LD, ARR = P::ARR;
  constexpr int Q4 = CG / 4;                // channel quads per group
  constexpr int BLKS = TMA_CONSUMERS / Q4;  // pixel blocks per tile: 8 (2x4) or 16 (1x4)
  constexpr int BR = 64 / (BLKS * 4);       // rows per block
  constexpr int UNITS = CG / 16;            // (head, 16-row m-tile) units per group, one per warp
  constexpr int NT = CH / 8;
  constexpr int HEADS = C / CH;
  static_assert(BLKS * Q4 == TMA_CONSUMERS && (BR == 1 || BR == 2), "CG must be 64 or 128");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;                                                        // NST boxes [10][10][CG] fp32
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(ring + NST * P::STAGE_BYTES);  // [PARTS][64][LD]
  __nv_bfloat16* Ks = Qs + PARTS * ARR;
  float* w9s = reinterpret_cast<float*>(Ks + PARTS * ARR);                         // [9][3C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w9s + 9 * 3 * C);                   // full[NST], empty[NST]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const long long sample0 = (long long)b * H * W;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[NST + s]), TMA_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < 9 * 3 * C / 4; e += TMA_THREADS) reinterpret_cast<float4*>(w9s)[e] = ldg4(w9 + 4 * e);
  __syncthreads();

  if (warp == TMA_CONSUMERS / 32) {
    // =============================== producer: one thread, TMA boxes ==============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < tiles_per_sample; tile += gridDim.x) {
        const int ty0 = (tile / tiles_x) * 8, tx0 = (tile - (tile / tiles_x) * tiles_x) * 8;
#pragma unroll 1
        for (int gi = 0; gi < NG; ++gi)
#pragma unroll 1
          for (int op = 0; op < 3; ++op, ++it) {
            const int s = it % NST;
            mbar_wait(smem_u32(&bars[NST + s]), ((it / NST) & 1) ^ 1);
            const uint32_t full = smem_u32(&bars[s]);
            mbar_expect_tx(full, P::STAGE_BYTES);
            tc::tma_load_4d(smem_u32(ring + (size_t)s * P::STAGE_BYTES), &tmX, op * C + gi * CG, tx0 - 1, ty0 - 1, b, full);
          }
      }
    }
    return;
  }

  // ================================= consumers: 8 warps ===========================================
  const int quad = tid % Q4, blk = tid / Q4;
  const int by = (blk >> 1) * BR, bx = (blk & 1) * 4;  // block origin inside the tile (== halo row/col of its top-left tap)
  float acc[NG][NT][4];
#pragma unroll
  for (int gi = 0; gi < NG; ++gi)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[gi][nt][e] = 0.f;
  float4 nrm[NG][2];  // sum of squares of this thread's 4 q / k channels over its pixels of every tile
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) nrm[gi][0] = nrm[gi][1] = make_float4(0.f, 0.f, 0.f, 0.f);

  const uint32_t q_base = smem_u32(Qs), k_base = smem_u32(Ks);
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < tiles_per_sample; tile += gridDim.x) {
    const int ty0 = (tile / tiles_x) * 8, tx0 = (tile - (tile / tiles_x) * tiles_x) * 8;
    const bool in_gram = ty0 >= gy0 && ty0 < gy1;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
#pragma unroll
      for (int op = 0; op < 3; ++op, ++it) {
        if (op == 2) {
          // ---- Gram of every head of the group (q, k tiles are complete) ------------------------------
          consumer_bar();
          if (warp < UNITS && in_gram) {
            const int m0 = warp * 16;
            const int n_base = (m0 / CH) * CH;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              uint32_t ah[4], al[4];
              const uint32_t a_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 4) & 1) * 8) * LD + m0 + ((lane >> 3) & 1) * 8) * 2);
              t_ldsm_x4_trans(q_base + a_off, ah);
              if (PARTS == 2) t_ldsm_x4_trans(q_base + ARR * 2 + a_off, al);
#pragma unroll
              for (int np = 0; np < NT / 2; ++np) {
                const uint32_t b_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + n_base + 8 * (2 * np + (lane >> 4))) * 2);
                uint32_t bh[4], bl[4];
                t_ldsm_x4_trans(k_base + b_off, bh);
                t_mma_bf16(acc[gi][2 * np], ah, bh[0], bh[1]);
                t_mma_bf16(acc[gi][2 * np + 1], ah, bh[2], bh[3]);
                if (PARTS == 2) {
                  t_ldsm_x4_trans(k_base + ARR * 2 + b_off, bl);
                  t_mma_bf16(acc[gi][2 * np], ah, bl[0], bl[1]);
                  t_mma_bf16(acc[gi][2 * np + 1], ah, bl[2], bl[3]);
                  t_mma_bf16(acc[gi][2 * np], al, bh[0], bh[1]);
                  t_mma_bf16(acc[gi][2 * np + 1], al, bh[2], bh[3]);
                }
              }
            }
          }
        }
        const int s = it % NST;
        float4 w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = *reinterpret_cast<const float4*>(w9s + t * 3 * C + op * C + gi * CG + quad * 4);
        mbar_wait(smem_u32(&bars[s]), (it / NST) & 1);
        float4 r[BR][4];
        conv_block<CG, BR>(reinterpret_cast<const float*>(ring + (size_t)s * P::STAGE_BYTES), w, by, bx, quad, r);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[NST + s]));  // this warp is done reading the box
        if (op == 2) {
#pragma unroll
          for (int rr = 0; rr < BR; ++rr) {
            const long long pix0 = sample0 + (long long)(ty0 + by + rr) * W + tx0 + bx;
#pragma unroll
            for (int o = 0; o < 4; ++o) *reinterpret_cast<float4*>(V + (pix0 + o) * ldv + gi * CG + quad * 4) = r[rr][o];
          }
          consumer_bar();  // the Gram reads of Qs / Ks are done before the next tile overwrites them
        } else {
          __nv_bfloat16* dst = (op == 0 ? Qs : Ks);
          float4& n4 = nrm[gi][op];
#pragma unroll
          for (int rr = 0; rr < BR; ++rr)
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              const float4 x = r[rr][o];
              if (in_gram) {
                n4.x = fmaf(x.x, x.x, n4.x); n4.y = fmaf(x.y, x.y, n4.y);
                n4.z = fmaf(x.z, x.z, n4.z); n4.w = fmaf(x.w, x.w, n4.w);
              }
              const int p = (by + rr) * 8 + bx + o;  // pixel index inside the tile
              uint2 hi, lo;
              t_split_pair(x.x, x.y, hi.x, lo.x);
              t_split_pair(x.z, x.w, hi.y, lo.y);
              *reinterpret_cast<uint2*>(dst + p * LD + quad * 4) = hi;
              if (PARTS == 2) *reinterpret_cast<uint2*>(dst + ARR + p * LD + quad * 4) = lo;
            }
        }
      }
    }
  }

  // ---- this CTA's partial: [b*HEADS + head][blockIdx.x][CH*CH + 2*CH] ----------------------------------
  constexpr int PER = CH * CH + 2 * CH;
  const int g = lane >> 2, qd = lane & 3;
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) {
    if (warp < UNITS) {
      const int m0 = warp * 16;
      const int head = gi * (CG / CH) + m0 / CH;
      const int i0 = m0 - (m0 / CH) * CH;
      float* dst = partial + ((long long)(b * HEADS + head) * gridDim.x + blockIdx.x) * PER;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int j = 8 * nt + 2 * qd;
        *reinterpret_cast<float2*>(dst + (i0 + g) * CH + j) = make_float2(acc[gi][nt][0], acc[gi][nt][1]);
        *reinterpret_cast<float2*>(dst + (i0 + g + 8) * CH + j) = make_float2(acc[gi][nt][2], acc[gi][nt][3]);
      }
    }
  }
  // squared norms: reduce the BLKS per-thread partial sums of every channel through shared memory (the ring is idle:
  // every box has been consumed, the producer has issued its last load long ago)
  consumer_bar();
  float* red = reinterpret_cast<float*>(ring);  // [BLKS][2][C]
#pragma unroll
  for (int gi = 0; gi < NG; ++gi)
#pragma unroll
    for (int op = 0; op < 2; ++op)
      *reinterpret_cast<float4*>(red + (blk * 2 + op) * C + gi * CG + quad * 4) = nrm[gi][op];
  consumer_bar();
  for (int t = tid; t < 2 * C; t += TMA_CONSUMERS) {
    const int op = t / C, ch = t - op * C;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < BLKS; ++k) s += red[(k * 2 + op) * C + ch];
    const int head = ch / CH, i = ch - head * CH;
    partial[((long long)(b * HEADS + head) * gridDim.x + blockIdx.x) * PER + CH * CH + op * CH + i] = s;
  }
}

static PFN_cuTensorMapEncodeTiled dwg_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(ptr);
  }
  return fn;
}

template <int CG, int CH, int NG, int PARTS>
static int launch_tma_t(const float* X, int ldx, const float* w9, float* V, int ldv, float* partial, int B, int H, int W,
                        int ctas, int gy0, int gy1, cudaStream_t st) {
  constexpr int C = CG * NG;
  PFN_cuTensorMapEncodeTiled enc = dwg_encode_fn();
  if (enc == nullptr) {
    set_error("dwgram(tma): cuTensorMapEncodeTiled unavailable");
    return MPHSIR_ERR_CUDA;
  }
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)3 * C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
  cuuint32_t box[4] = {(cuuint32_t)CG, HALO, HALO, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    set_error("dwgram(tma): cuTensorMapEncodeTiled failed (C=%d H=%d W=%d ldx=%d)", C, H, W, ldx);
    return MPHSIR_ERR_CUDA;
  }
  const size_t smem = TmaPlan<CG, PARTS>::smem(C);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwgram_tma_kernel<CG, CH, NG, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("dwgram(tma): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  const int tiles_x = W / 8, tiles = tiles_x * (H / 8);
  dim3 grid(ctas, B);
  dwgram_tma_kernel<CG, CH, NG, PARTS><<<grid, TMA_THREADS, smem, st>>>(tm, w9, V, ldv, partial, H, W, tiles_x, tiles, gy0, gy1);
  return check_launch("dwgram(tma)");
}

// cfg numbering of dwgram.cu: 1 (64,32,1)  2 (128,64,1)  3 (128,32,1)  4 (128,32,2)
int launch_tma(int cfg, bool x3, const float* X, int ldx, const float* w9, float* V, int ldv, float* partial, int B, int H,
               int W, int ctas, int gy0, int gy1, cudaStream_t st) {
#define DWT(CG, CH, NG) (x3 ? launch_tma_t<CG, CH, NG, 2>(X, ldx, w9, V, ldv, partial, B, H, W, ctas, gy0, gy1, st) \
                            : launch_tma_t<CG, CH, NG, 1>(X, ldx, w9, V, ldv, partial, B, H, W, ctas, gy0, gy1, st))
  switch (cfg) {
    case 1: return DWT(64, 32, 1);
    case 2: return DWT(128, 64, 1);
    case 3: return DWT(128, 32, 1);
    case 4: return DWT(128, 32, 2);
    default: set_error("dwgram(tma): unsupported configuration %d", cfg); return MPHSIR_ERR_INVALID;
  }
#undef DWT
}

}  // namespace dwg
}  // namespace mphsir
