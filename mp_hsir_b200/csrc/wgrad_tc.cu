// Weight gradients on tcgen05:  dW[o, i] += sum_m dY[m, o] * X[src(m), i]   (same contract as wgrad.cu)
//
// The contraction runs over tokens — the slow axis of both token-major fp32 operands — so the operands are
// *transposed on the way into shared memory*: 8 converter warps read 64-token chunks with 128-bit loads (a thread owns
// 8 consecutive tokens of 4 channels), round to bf16 (hi, or hi+lo for the fp32-grade mode) and write, per channel,
// one 16-byte chunk of 8 tokens into the 128-byte-swizzled K-major image that tcgen05.mma reads (rows = channels,
// K = tokens; the same SWIZZLE_128B layout / descriptors as the forward GEMM engine, gemm_tc.cu).  One elected thread
// issues tcgen05.mma (M = 128 output channels, N <= 128 input channels, K = 16 tokens per instruction) into a TMEM
// accumulator that lives for the whole kernel; the image is double-buffered through two mbarrier rings
// (full: converters -> MMA, empty: tcgen05.commit -> converters).  At the end four warps drain TMEM with
// tcgen05.ld and scatter-add the tile through the output map (fp32 atomics into the pre-zeroed gradient).
//
// Each CTA owns one 128 x 128 tile of dW and a strided subset of the token chunks (grid.x = tile so that CTAs resident
// together walk the same chunks and share them through L2, grid.y = token split, grid.z = sample / conv tap).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mphsir {
namespace wgt {

using namespace mphsir::tc;

constexpr int KT = 64;                  // tokens per stage (= one 128-byte K row of the image)
constexpr int TO = 128, TI = 128;       // dW tile
constexpr int kThreads = 288;           // warp 0: MMA issue + TMEM owner; warps 1-8: converters; warps 1-4 also drain TMEM
constexpr int IMG_BYTES = 128 * 128;    // one part (hi or lo) of a 128-row x 64-token image
constexpr int TMEM_COLS = 128;

__device__ __forceinline__ int map_index(int o, int mode, int a, int b) {
  if (mode == MPHSIR_MAP_IDENTITY) return o < a ? o : -1;
  if (mode == MPHSIR_MAP_INTERLEAVE) {
    const int j = o >> 1;
    return j < a ? (o & 1) * a + j : -1;
  }
  if (o < b) return o < a ? o : -1;
  return (o - b) < a ? a + (o - b) : -1;
}

struct Args {
  const float* dY;
  long long lddy;
  const float* X;
  long long ldx;
  float* dW;
  int M;
  int O, I;
  int rows_per_batch;
  long long dw_batch_stride;
  int x_row_mod;
  int H, W, taps;
  long long so, si, st;
  int map_mode, map_a, map_b;
  int i_valid;
  int tiles_i;
  float* dbias;
};

struct Bars {
  uint64_t full[2], empty[2], acc_full;
  uint32_t tmem_base;
};

template <int PARTS>
__global__ void __launch_bounds__(kThreads, (PARTS == 1 ? 2 : 1)) wgrad_tc_kernel(const Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // [stage][A hi | A lo | B hi | B lo] images (1024-byte aligned), barriers behind them
  uint8_t* img = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int STAGE_BYTES = 2 * PARTS * IMG_BYTES;
  Bars* bars = reinterpret_cast<Bars*>(img + 2 * STAGE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile_o = blockIdx.x / p.tiles_i, tile_i = blockIdx.x - tile_o * p.tiles_i;
  const int o0 = tile_o * TO, i0 = tile_i * TI;
  int n_mma = min(TI, p.I - i0);
  n_mma = (n_mma + 15) & ~15;  // tcgen05 N granularity at M = 128; the padding rows are zero-filled by the converters

  int m_begin = 0, m_end = p.M;
  int tap = 0, dy = 0, dx = 0;
  float* dW = p.dW;
  if (p.taps == 9) {
    tap = blockIdx.z;
    dy = tap / 3 - 1;
    dx = tap - (tap / 3) * 3 - 1;
  } else if (p.rows_per_batch > 0) {
    m_begin = blockIdx.z * p.rows_per_batch;
    m_end = m_begin + p.rows_per_batch;
    dW += (long long)blockIdx.z * p.dw_batch_stride;
  }
  const int n_chunks = (m_end - m_begin + KT - 1) / KT;
  const int my_chunks = (int)blockIdx.y < n_chunks ? (n_chunks - 1 - (int)blockIdx.y) / (int)gridDim.y + 1 : 0;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 8);   // one arrival per converter warp
      mbar_init(smem_u32(&bars->empty[s]), 1);  // tcgen05.commit
    }
    mbar_init(smem_u32(&bars->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&bars->tmem_base), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // =============================== MMA issuer ======================================================
    const uint32_t idesc = make_idesc((uint32_t)n_mma);
    for (int it = 0; it < my_chunks; ++it) {
      const int s = it & 1;
      mbar_wait(smem_u32(&bars->full[s]), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(img + s * STAGE_BYTES);
      const uint32_t b_addr = a_addr + PARTS * IMG_BYTES;
      const uint64_t ah0 = make_desc(a_addr), bh0 = make_desc(b_addr);
      const uint64_t al0 = make_desc(a_addr + IMG_BYTES), bl0 = make_desc(b_addr + IMG_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {  // consecutive k16 steps are 32 bytes apart (+2 in the 16-byte address field)
          umma_bf16(tmem_base, ah0 + 2 * k, bh0 + 2 * k, idesc, (it | k) != 0);
          if (PARTS == 2) {
            umma_bf16(tmem_base, ah0 + 2 * k, bl0 + 2 * k, idesc, 1);
            umma_bf16(tmem_base, al0 + 2 * k, bh0 + 2 * k, idesc, 1);
          }
        }
        umma_commit(smem_u32(&bars->empty[s]));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bars->acc_full));
    __syncwarp();
  } else {
    // =============================== converters =======================================================
    const int ct = tid - 32;
    const int tg = (ct & 31) >> 2;                 // token group: tokens 8 tg .. 8 tg + 7 of the chunk
    const int cq = (ct >> 5) * 4 + (ct & 3);       // channel quad: channels 4 cq .. 4 cq + 3 of the tile
    const bool a_col_ok = o0 + 4 * cq < p.O;
    const bool b_col_ok = i0 + 4 * cq < p.I;
    const bool b_store = 4 * cq < n_mma;
    const float* a_col = p.dY + o0 + 4 * cq;
    const float* b_col = p.X + i0 + 4 * cq;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ra[8], rb[8];

    auto load_chunk = [&](int chunk) {
      const int m0 = m_begin + chunk * KT + 8 * tg;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        ra[j] = (a_col_ok && m0 + j < m_end) ? ldg4(a_col + (long long)(m0 + j) * p.lddy) : zero4;
      if (!b_store) return;
      if (p.taps == 9) {
        if ((p.W & 7) == 0) {
          // 8 consecutive tokens share an image row: one div/mod per chunk
          const int x0 = m0 % p.W, y = (m0 / p.W) % p.H;
          const bool row_ok = b_col_ok && m0 < m_end && (unsigned)(y + dy) < (unsigned)p.H;
          const float* base = b_col + (long long)(m0 + dy * p.W + dx) * p.ldx;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            rb[j] = (row_ok && (unsigned)(x0 + j + dx) < (unsigned)p.W) ? ldg4(base + (long long)j * p.ldx) : zero4;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int m = m0 + j;
            const int x = m % p.W, y = (m / p.W) % p.H;
            const bool ok = b_col_ok && m < m_end && (unsigned)(y + dy) < (unsigned)p.H && (unsigned)(x + dx) < (unsigned)p.W;
            rb[j] = ok ? ldg4(b_col + (long long)(m + dy * p.W + dx) * p.ldx) : zero4;
          }
        }
      } else if (p.x_row_mod > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          rb[j] = (b_col_ok && m0 + j < m_end) ? ldg4(b_col + (long long)((m0 + j) % p.x_row_mod) * p.ldx) : zero4;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          rb[j] = (b_col_ok && m0 + j < m_end) ? ldg4(b_col + (long long)(m0 + j) * p.ldx) : zero4;
      }
    };
    // 8 tokens x 4 channels -> per channel one 16-byte chunk (8 tokens) of image row `4 cq + e`, chunk index tg
    auto store_unit = [&](const float4 (&r)[8], uint8_t* dst) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        auto comp = [e](const float4& v) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); };
        uint4 hi, lo;
        split2(comp(r[0]), comp(r[1]), hi.x, lo.x);
        split2(comp(r[2]), comp(r[3]), hi.y, lo.y);
        split2(comp(r[4]), comp(r[5]), hi.z, lo.z);
        split2(comp(r[6]), comp(r[7]), hi.w, lo.w);
        const int row = 4 * cq + e;
        const int off = row * 128 + ((tg ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst + off) = hi;
        if (PARTS == 2) *reinterpret_cast<uint4*>(dst + IMG_BYTES + off) = lo;
      }
    };

    const bool want_bias = p.dbias != nullptr && tile_i == 0 && blockIdx.z == 0;
    float4 bsum = zero4;
    if (my_chunks > 0) load_chunk(blockIdx.y);
    for (int it = 0; it < my_chunks; ++it) {
      const int s = it & 1;
      mbar_wait(smem_u32(&bars->empty[s]), ((it >> 1) & 1) ^ 1);  // the MMAs that read this stage two chunks ago are done
      uint8_t* a_dst = img + s * STAGE_BYTES;
      if (want_bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bsum.x += ra[j].x;
          bsum.y += ra[j].y;
          bsum.z += ra[j].z;
          bsum.w += ra[j].w;
        }
      }
      store_unit(ra, a_dst);
      if (b_store) store_unit(rb, a_dst + PARTS * IMG_BYTES);
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->full[s]));
      if (it + 1 < my_chunks) load_chunk(blockIdx.y + (it + 1) * gridDim.y);
    }

    // bias gradient: the 8 token groups of a channel quad sit in lanes l, l^4, l^8, ... of one warp
    if (want_bias) {
#pragma unroll
      for (int sh = 4; sh < 32; sh <<= 1) {
        bsum.x += __shfl_xor_sync(0xffffffffu, bsum.x, sh);
        bsum.y += __shfl_xor_sync(0xffffffffu, bsum.y, sh);
        bsum.z += __shfl_xor_sync(0xffffffffu, bsum.z, sh);
        bsum.w += __shfl_xor_sync(0xffffffffu, bsum.w, sh);
      }
      if (tg == 0) {
        const float bv[4] = {bsum.x, bsum.y, bsum.z, bsum.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int o = o0 + 4 * cq + e;
          const int ro = o < p.O ? map_index(o, p.map_mode, p.map_a, p.map_b) : -1;
          if (ro >= 0) atomicAdd(p.dbias + ro, bv[e]);
        }
      }
    }

    // =============================== epilogue (warps 1-4: TMEM lane quadrant = warp % 4) ===============
    if (warp <= 4 && my_chunks > 0) {
      mbar_wait(smem_u32(&bars->acc_full), 0);
      tc_fence_after();
      const int quad = warp & 3;
      // a lane holds one dW row out of tcgen05.ld; transpose 32x32 blocks through shared memory (the image area is dead:
      // every MMA has completed) so that a warp adds 32 consecutive input channels of one row per instruction
      float* tr = reinterpret_cast<float*>(img) + quad * 32 * 33;
      for (int c0 = 0; c0 < n_mma; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + c0, r);
#pragma unroll
        for (int c = 0; c < 32; ++c) tr[lane * 33 + c] = __uint_as_float(r[c]);
        __syncwarp();
        const int i = i0 + c0 + lane;
        if (i < p.i_valid) {
          float* col = dW + (long long)i * p.si + (long long)tap * p.st;
          for (int rr = 0; rr < 32; ++rr) {
            const int o = o0 + quad * 32 + rr;
            if (o >= p.O) break;
            const int ro = map_index(o, p.map_mode, p.map_a, p.map_b);
            if (ro >= 0) atomicAdd(col + (long long)ro * p.so, tr[rr * 33 + lane]);
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

template <int PARTS>
static int launch(const Args& a, int z, cudaStream_t st) {
  const size_t smem = (size_t)2 * 2 * PARTS * IMG_BYTES + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("wgrad(tc): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  const int tiles = ((a.O + TO - 1) / TO) * a.tiles_i;
  const int rows = a.rows_per_batch > 0 && a.taps != 9 ? a.rows_per_batch : a.M;
  const int n_chunks = (rows + KT - 1) / KT;
  const int per_sm = PARTS == 1 ? 2 : 1;  // 2 x 64 KB of images fit one SM in bf16 mode
  long long splits = ((long long)per_sm * sm_count()) / ((long long)tiles * z);  // floor: a single wave
  if (splits > (n_chunks + 3) / 4) splits = (n_chunks + 3) / 4;                  // >= 4 chunks per CTA
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)tiles, (unsigned)splits, (unsigned)z);
  wgrad_tc_kernel<PARTS><<<grid, kThreads, smem, st>>>(a);
  return check_launch("wgrad(tc)");
}

int launch_wgrad_tc(const mphsir_wgrad_params* p, int z, cudaStream_t st) {
  Args a;
  a.dY = p->dY;
  a.lddy = p->lddy;
  a.X = p->X;
  a.ldx = p->ldx;
  a.dW = p->dW;
  a.M = (int)p->M;
  a.O = p->O;
  a.I = p->I;
  a.rows_per_batch = p->rows_per_batch;
  a.dw_batch_stride = p->dw_batch_stride;
  a.x_row_mod = p->x_row_mod;
  a.H = p->H;
  a.W = p->W;
  a.taps = p->taps;
  a.so = p->so;
  a.si = p->si;
  a.st = p->st;
  a.map_mode = p->map_mode;
  a.map_a = p->map_a;
  a.map_b = p->map_b;
  a.i_valid = p->i_valid > 0 ? p->i_valid : p->I;
  a.tiles_i = (p->I + TI - 1) / TI;
  a.dbias = p->dbias;
  return p->precision == MPHSIR_PREC_BF16X3 ? launch<2>(a, z, st) : launch<1>(a, z, st);
}

}  // namespace wgt
}  // namespace mphsir
