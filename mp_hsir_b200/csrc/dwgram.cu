// Fused depthwise-3x3 conv + Gram statistics of the global spectral attention
// (Spectral_Attention.forward, net/MP_HSIR.py:98-107; Attention.forward :409-419).
//
// qkv = dwconv3x3(conv1x1(x)) is never materialised: each persistent CTA walks 8x8-pixel tiles of one
// sample, computes the depthwise conv of the q, k and v channel blocks from the 1x1 output (128-bit loads,
// 3x6 patch reuse per 1x4 strip), writes only v to HBM, and keeps q and k of the tile in shared memory as
// bf16 hi/lo parts.  The Gram matrix q^T k of every head is accumulated across tiles in registers on the
// tensor cores (mma.sync m16n8k16, ldmatrix.trans reads the [pixel][channel] tile as both operands; hi*hi +
// hi*lo + lo*hi in the fp32-grade mode), together with sum q^2 / sum k^2 for the L2 normalisation, which is
// applied after the reduction (q^ k^T = (q k^T)/(|q||k|)).  One partial per CTA goes to HBM and is reduced by
// spectral_finish.  HBM traffic per token: read 3C, write C floats (the unfused path moved 8C).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mphsir {
namespace dwg {

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// depthwise conv of a 1x4 strip x 4 channels: rows y-1..y+1, columns x0-1..x0+4 of one channel quad
__device__ __forceinline__ void dw_strip(const float* __restrict__ X, long long ldx, const float* __restrict__ w9, int wld,
                                         long long pix0, int y, int x0, int H, int W, int ch, float4 (&acc)[4]) {
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    const float* rowp = X + (pix0 + (long long)dy * W) * ldx + ch;
    const float* wr = w9 + (dy + 1) * 3 * wld + ch;
    const float4 w0 = ldg4(wr), w1 = ldg4(wr + wld), w2 = ldg4(wr + 2 * wld);
    float4 v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int xx = x0 - 1 + j;
      v[j] = (xx >= 0 && xx < W) ? ldg4(rowp + (long long)(j - 1) * ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      acc[o].x = fmaf(v[o].x, w0.x, fmaf(v[o + 1].x, w1.x, fmaf(v[o + 2].x, w2.x, acc[o].x)));
      acc[o].y = fmaf(v[o].y, w0.y, fmaf(v[o + 1].y, w1.y, fmaf(v[o + 2].y, w2.y, acc[o].y)));
      acc[o].z = fmaf(v[o].z, w0.z, fmaf(v[o + 1].z, w1.z, fmaf(v[o + 2].z, w2.z, acc[o].z)));
      acc[o].w = fmaf(v[o].w, w0.w, fmaf(v[o + 1].w, w1.w, fmaf(v[o + 2].w, w2.w, acc[o].w)));
    }
  }
}

// CG: channels per head group (multiple of 16, <= 128), CH: channels per head, NG: head groups (C = NG*CG)
template <int CG, int CH, int NG, int PARTS>
__global__ void __launch_bounds__(256) dwgram_kernel(const float* __restrict__ X, long long ldx,
                                                     const float* __restrict__ w9, float* __restrict__ V, long long ldv,
                                                     float* __restrict__ partial, int H, int W, int tiles_x,
                                                     int tiles_per_sample, int gy0, int gy1) {
  // gy0, gy1: only tiles whose first row lies in [gy0, gy1) contribute to the Gram statistics (row-sharded scenes: the
  // depthwise conv and V are computed on the halo rows too, the statistics on the rank's own rows only)
  constexpr int C = CG * NG;
  constexpr int LD = CG + 8;           // bf16 elements per smem row
  constexpr int ARR = 64 * LD;
  constexpr int UNITS = CG / 16;       // (head, 16-row m-tile) units per group, one per warp
  constexpr int NT = CH / 8;           // n-tiles per unit
  constexpr int Q4 = CG / 4;           // channel quads per group
  constexpr int HEADS = C / CH;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [PARTS][64][LD]
  __nv_bfloat16* Ks = Qs + PARTS * ARR;                             // [PARTS][64][LD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const long long sample0 = (long long)b * H * W;

  float acc[NG][NT][4];
#pragma unroll
  for (int gi = 0; gi < NG; ++gi)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[gi][nt][e] = 0.f;
  float nrm[NG];  // thread t < 2*CG: sum of squares of q channel t (t < CG) or k channel t-CG
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) nrm[gi] = 0.f;

  const uint32_t q_base = (uint32_t)__cvta_generic_to_shared(Qs);
  const uint32_t k_base = (uint32_t)__cvta_generic_to_shared(Ks);

  for (int tile = blockIdx.x; tile < tiles_per_sample; tile += gridDim.x) {
    const int ty0 = (tile / tiles_x) * 8, tx0 = (tile - (tile / tiles_x) * tiles_x) * 8;
    const bool in_gram = ty0 >= gy0 && ty0 < gy1;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      // ---- phase A: depthwise conv of this group's q, k, v channels for the 8x8 tile ----------------
      for (int item = tid; item < 3 * 16 * Q4; item += 256) {
        const int op = item / (16 * Q4);            // 0 = q, 1 = k, 2 = v
        const int rem = item - op * (16 * Q4);
        const int strip = rem / Q4, quad = rem - strip * Q4;
        const int y = ty0 + (strip >> 1), x0 = tx0 + (strip & 1) * 4;
        const int ch = op * C + gi * CG + quad * 4;  // channel in the [q | k | v] layout of X / w9
        const long long pix0 = sample0 + (long long)y * W + x0;
        float4 r[4];
        dw_strip(X, ldx, w9, 3 * C, pix0, y, x0, H, W, ch, r);
        if (op == 2) {
#pragma unroll
          for (int o = 0; o < 4; ++o) *reinterpret_cast<float4*>(V + (pix0 + o) * ldv + gi * CG + quad * 4) = r[o];
        } else {
          __nv_bfloat16* dst = (op == 0 ? Qs : Ks);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int p = (strip >> 1) * 8 + (strip & 1) * 4 + o;  // pixel index inside the tile
            uint2 hi, lo;
            split_pair(r[o].x, r[o].y, hi.x, lo.x);
            split_pair(r[o].z, r[o].w, hi.y, lo.y);
            *reinterpret_cast<uint2*>(dst + p * LD + quad * 4) = hi;
            if (PARTS == 2) *reinterpret_cast<uint2*>(dst + ARR + p * LD + quad * 4) = lo;
          }
        }
      }
      __syncthreads();
      // ---- phase B: Gram of every head of the group on the tensor cores + squared norms ---------------
      if (warp < UNITS && in_gram) {
        const int m0 = warp * 16;                  // first q channel of this unit (inside the group)
        const int n_base = (m0 / CH) * CH;         // first k channel of the same head
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {           // 16 pixels per k-step
          uint32_t ah[4], al[4];
          const uint32_t a_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 4) & 1) * 8) * LD + m0 + ((lane >> 3) & 1) * 8) * 2);
          ldsm_x4_trans(q_base + a_off, ah);
          if (PARTS == 2) ldsm_x4_trans(q_base + ARR * 2 + a_off, al);
#pragma unroll
          for (int np = 0; np < NT / 2; ++np) {
            const uint32_t b_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + n_base + 8 * (2 * np + (lane >> 4))) * 2);
            uint32_t bh[4], bl[4];
            ldsm_x4_trans(k_base + b_off, bh);
            mma_bf16(acc[gi][2 * np], ah, bh[0], bh[1]);
            mma_bf16(acc[gi][2 * np + 1], ah, bh[2], bh[3]);
            if (PARTS == 2) {
              ldsm_x4_trans(k_base + ARR * 2 + b_off, bl);
              mma_bf16(acc[gi][2 * np], ah, bl[0], bl[1]);
              mma_bf16(acc[gi][2 * np + 1], ah, bl[2], bl[3]);
              mma_bf16(acc[gi][2 * np], al, bh[0], bh[1]);
              mma_bf16(acc[gi][2 * np + 1], al, bh[2], bh[3]);
            }
          }
        }
      }
      if (tid < 2 * CG && in_gram) {
        const __nv_bfloat16* src = (tid < CG ? Qs + tid : Ks + (tid - CG));
        float s = 0.f;
#pragma unroll 8
        for (int p = 0; p < 64; ++p) {
          float v = __bfloat162float(src[p * LD]);
          if (PARTS == 2) v += __bfloat162float(src[ARR + p * LD]);
          s = fmaf(v, v, s);
        }
        nrm[gi] += s;
      }
      __syncthreads();
    }
  }

  // ---- write this CTA's partial: [b*HEADS + head][blockIdx.x][CH*CH + 2*CH] -------------------------
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
    if (tid < 2 * CG) {
      const int t = tid < CG ? tid : tid - CG;
      const int head = gi * (CG / CH) + t / CH;
      const int i = t - (t / CH) * CH;
      float* dst = partial + ((long long)(b * HEADS + head) * gridDim.x + blockIdx.x) * PER;
      dst[CH * CH + (tid < CG ? 0 : CH) + i] = nrm[gi];
    }
  }
}

// TMA-fed variant (dwgram_tma.cu): 64/128-channel head groups on images of at least 16x16 pixels
int launch_tma(int cfg, bool x3, const float* X, int ldx, const float* w9, float* V, int ldv, float* partial, int B, int H,
               int W, int ctas, int gy0, int gy1, cudaStream_t st);
static bool g_use_tma = true;
static bool tma_eligible(int cfg, int H, int W) { return g_use_tma && cfg >= 1 && cfg <= 4 && H >= 16 && W >= 16; }

// CTAs per sample: the direct kernel runs 2 CTAs per SM, the TMA kernel is persistent with one CTA per SM
static int ctas_per_sample(int B, int tiles, bool tma) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  int n = (tma ? sms : 2 * sms) / (B > 0 ? B : 1);
  if (n < 1) n = 1;
  if (n > tiles) n = tiles;
  return n;
}

template <int CG, int CH, int NG, int PARTS>
static int launch(const float* X, int ldx, const float* w9, float* V, int ldv, float* partial, int B, int H, int W,
                  int gy0, int gy1, cudaStream_t st) {
  const size_t smem = (size_t)2 * PARTS * 64 * (CG + 8) * 2;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwgram_kernel<CG, CH, NG, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("dwgram: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  const int tiles_x = W / 8, tiles = tiles_x * (H / 8);
  dim3 grid(ctas_per_sample(B, tiles, false), B);
  dwgram_kernel<CG, CH, NG, PARTS><<<grid, 256, smem, st>>>(X, ldx, w9, V, ldv, partial, H, W, tiles_x, tiles, gy0, gy1);
  return check_launch("dwgram");
}

}  // namespace dwg
}  // namespace mphsir

using namespace mphsir;

// supported (C, c) combinations: C = NG*CG with CG in {64, 96, 128}
static int dwgram_config(int C, int c) {
  if (C == 64 && c == 32) return 1;
  if (C == 128 && c == 64) return 2;
  if (C == 128 && c == 32) return 3;
  if (C == 256 && c == 32) return 4;
  if (C == 96 && c == 48) return 5;
  return 0;
}

extern "C" int mphsir_dwgram_supported(int C, int c) { return dwgram_config(C, c) != 0; }

static int dwgram_config(int C, int c);

extern "C" void mphsir_debug_dwgram_tma(int enabled) { dwg::g_use_tma = enabled != 0; }

extern "C" size_t mphsir_dwgram_partial_floats(int B, int heads, int c, int H, int W, int* n_chunks) {
  const int n = dwg::ctas_per_sample(B, (H / 8) * (W / 8), dwg::tma_eligible(dwgram_config(heads * c, c), H, W));
  if (n_chunks) *n_chunks = n;
  return (size_t)B * heads * n * ((size_t)c * c + 2 * c);
}

extern "C" int mphsir_dwgram_band_fwd(const float* X, int ldx, const float* w9, float* V, int ldv, float* partial, int B,
                                      int H, int W, int C, int heads, int precision, int gram_y0, int gram_y1, void* stream) {
  MPHSIR_REQUIRE(X && w9 && V && partial, "dwgram: null operand");
  MPHSIR_REQUIRE(B > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "dwgram: H=%d W=%d must be multiples of 8", H, W);
  MPHSIR_REQUIRE(heads > 0 && C % heads == 0, "dwgram: C=%d not divisible by heads=%d", C, heads);
  MPHSIR_REQUIRE(ldx >= 3 * C && ldx % 4 == 0 && ldv >= C && ldv % 4 == 0, "dwgram: bad leading dimensions");
  MPHSIR_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(w9) |
                   reinterpret_cast<uintptr_t>(partial)) & 15) == 0, "dwgram: operands must be 16-byte aligned");
  MPHSIR_REQUIRE(precision == MPHSIR_PREC_BF16X3 || precision == MPHSIR_PREC_BF16, "dwgram: tensor-core precisions only (use dwconv3x3 + gram_partial for fp32 SIMT)");
  MPHSIR_REQUIRE(gram_y0 % 8 == 0 && gram_y1 % 8 == 0 && 0 <= gram_y0 && gram_y0 <= gram_y1, "dwgram: Gram row range [%d,%d) must be tile aligned", gram_y0, gram_y1);
  const int cfg = dwgram_config(C, C / heads);
  MPHSIR_REQUIRE(cfg != 0, "dwgram: unsupported (C=%d, c=%d); use dwconv3x3 + gram_partial", C, C / heads);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool x3 = precision == MPHSIR_PREC_BF16X3;
  if (dwg::tma_eligible(cfg, H, W))
    return dwg::launch_tma(cfg, x3, X, ldx, w9, V, ldv, partial, B, H, W, dwg::ctas_per_sample(B, (H / 8) * (W / 8), true),
                           gram_y0, gram_y1, st);
#define DWG(CG, CH, NG) (x3 ? dwg::launch<CG, CH, NG, 2>(X, ldx, w9, V, ldv, partial, B, H, W, gram_y0, gram_y1, st) \
                            : dwg::launch<CG, CH, NG, 1>(X, ldx, w9, V, ldv, partial, B, H, W, gram_y0, gram_y1, st))
  switch (cfg) {
    case 1: return DWG(64, 32, 1);
    case 2: return DWG(128, 64, 1);
    case 3: return DWG(128, 32, 1);
    case 4: return DWG(128, 32, 2);
    default: return DWG(96, 48, 1);
  }
#undef DWG
}

extern "C" int mphsir_dwgram_fwd(const float* X, int ldx, const float* w9, float* V, int ldv, float* partial, int B,
                                 int H, int W, int C, int heads, int precision, void* stream) {
  return mphsir_dwgram_band_fwd(X, ldx, w9, V, ldv, partial, B, H, W, C, heads, precision, 0, H, stream);
}
