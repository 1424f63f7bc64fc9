// Window attention core on tensor cores (warp-level mma.sync m16n8k16 bf16, fp32 accumulate).
//
// One CTA (4 warps) per (window, head): q,k,v are gathered from the token-major qkv buffer with the
// roll / window_partition addressing folded in (net/MP_HSIR.py:671-678), split into bf16 hi (+lo)
// parts in shared memory, S = q k^T and O = P v run on the tensor cores with the probabilities kept in
// registers (accumulator fragments are re-used as A fragments), bias + closed-form Swin mask + softmax
// in fp32 registers.  PARTS = 2 evaluates every product as hi*hi + hi*lo + lo*hi (fp32-grade, used by
// the "fp32" precision mode); PARTS = 1 is plain bf16.
//
// The 64x64xhd problem is far too small to feed a tcgen05 M=128 tile without packing two windows per
// MMA and wasting half of it; it is HBM-bound either way (16*C bytes/token vs 256*C MACs/token), so the
// warp-level path is used here and the TMEM engine is reserved for the projections (gemm_tc.cu).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mphsir {
namespace wa {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
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

template <int HD, int PARTS>
__global__ void __launch_bounds__(128, (HD <= 48 ? 6 : 4)) window_attn_mma_kernel(const float* __restrict__ qkv, long long ldqkv,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              long long ldo, float* __restrict__ win_mean, int H, int W,
                                                              int C, int shift, int mask_H, int mask_y0) {
  // mask_H, mask_y0: the Swin mask (net/MP_HSIR.py:643-658) is a function of the row in the shifted coordinates of the WHOLE
  // image; for a row band of a sharded scene local shifted row ys is row (ys + mask_y0) mod mask_H of the scene (the plain
  // call passes mask_H = H, mask_y0 = 0)
  constexpr int LD = HD + 8;            // bf16 elements per smem row (16-byte pad: conflict-free ldmatrix)
  constexpr int ARR = 64 * LD;          // elements per operand array
  constexpr int KS = HD / 16;           // k-steps of q k^T
  constexpr int NT_O = HD / 8;          // n-tiles of the output
  constexpr int HD4 = HD / 4;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // [Q parts][K parts][V parts] bf16, then rows[64], label[64]; the fp32 output tile later overlays Q|K.
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Ks_ = Qs + PARTS * ARR;
  __nv_bfloat16* Vs = Ks_ + PARTS * ARR;
  int* rows = reinterpret_cast<int*>(Vs + PARTS * ARR);
  int* label = rows + 64;
  float* Os = reinterpret_cast<float*>(smem_raw);  // [64][HD] fp32 (needs 2*ARR*2 >= 64*HD*4 bytes: true for HD >= 16)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;
  const int win = blockIdx.x;
  const int nWx = W >> 3, nW = (H >> 3) * nWx;
  const int b = win / nW;
  const int wrem = win - b * nW;
  const int wi = wrem / nWx, wj = wrem - wi * nWx;

  if (tid < 64) {
    const int r = tid >> 3, c = tid & 7;
    const int ys = wi * 8 + r, xs = wj * 8 + c;  // shifted coordinates
    int y = ys + shift, x = xs + shift;          // roll(-s): shifted[ys] = x[(ys+s) mod H]
    if (y >= H) y -= H;
    if (x >= W) x -= W;
    rows[tid] = (b * H + y) * W + x;
    int ysg = ys + mask_y0;
    if (ysg >= mask_H) ysg -= mask_H;
    const int rh = (ysg >= mask_H - 8) + (ysg >= mask_H - 4);
    const int rw = (xs >= W - 8) + (xs >= W - 4);
    label[tid] = shift ? 3 * rh + rw : 0;
  }
  __syncthreads();

  // ---- gather + split q (pre-scaled), k, v -----------------------------------------------------------
  const float scale = rsqrtf((float)HD);
  // all loads of a batch are issued before the first conversion, so one CTA keeps GB x 3 x 16 B per thread in
  // flight instead of paying one global round trip per 4 channels (the gather was the top stall in ncu)
  constexpr int ITERS = HD / 8;                       // (64 tokens x HD/4 quads) / 128 threads
  constexpr int GB = (ITERS % 4 == 0) ? 4 : 3;        // HD 32/64 -> 4, HD 48/96 -> 3
#pragma unroll
  for (int it0 = 0; it0 < ITERS; it0 += GB) {
    float4 q[GB], k[GB], v[GB];
#pragma unroll
    for (int u = 0; u < GB; ++u) {
      const int idx = tid + (it0 + u) * 128;
      const int t = idx / HD4, d4 = idx - t * HD4;
      const float* base = qkv + (long long)rows[t] * ldqkv + head * HD + d4 * 4;
      q[u] = ldg4(base);
      k[u] = ldg4(base + C);
      v[u] = ldg4(base + 2 * C);
    }
#pragma unroll
    for (int u = 0; u < GB; ++u) {
      const int idx = tid + (it0 + u) * 128;
      const int t = idx / HD4, d4 = idx - t * HD4;
      q[u].x *= scale; q[u].y *= scale; q[u].z *= scale; q[u].w *= scale;
      const int off = t * LD + d4 * 4;
      uint2 hi, lo;
      split_pair(q[u].x, q[u].y, hi.x, lo.x); split_pair(q[u].z, q[u].w, hi.y, lo.y);
      *reinterpret_cast<uint2*>(Qs + off) = hi;
      if (PARTS == 2) *reinterpret_cast<uint2*>(Qs + ARR + off) = lo;
      split_pair(k[u].x, k[u].y, hi.x, lo.x); split_pair(k[u].z, k[u].w, hi.y, lo.y);
      *reinterpret_cast<uint2*>(Ks_ + off) = hi;
      if (PARTS == 2) *reinterpret_cast<uint2*>(Ks_ + ARR + off) = lo;
      split_pair(v[u].x, v[u].y, hi.x, lo.x); split_pair(v[u].z, v[u].w, hi.y, lo.y);
      *reinterpret_cast<uint2*>(Vs + off) = hi;
      if (PARTS == 2) *reinterpret_cast<uint2*>(Vs + ARR + off) = lo;
    }
  }
  __syncthreads();

  const uint32_t q_base = (uint32_t)__cvta_generic_to_shared(Qs);
  const uint32_t k_base = (uint32_t)__cvta_generic_to_shared(Ks_);
  const uint32_t v_base = (uint32_t)__cvta_generic_to_shared(Vs);
  const int g = lane >> 2, qd = lane & 3;

  // ---- S = q k^T : warp owns query rows 16*warp .. +15 --------------------------------------------------
  float s[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t ah[4], al[4];
    const uint32_t a_off = (uint32_t)(((16 * warp + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + 16 * ks + (lane >> 4) * 8) * 2);
    ldsm_x4(q_base + a_off, ah);
    if (PARTS == 2) ldsm_x4(q_base + ARR * 2 + a_off, al);
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // pairs of key n-tiles
      const uint32_t b_off = (uint32_t)(((8 * (2 * np + (lane >> 4)) + (lane & 7)) * LD + 16 * ks + ((lane >> 3) & 1) * 8) * 2);
      uint32_t bh[4], bl[4];
      ldsm_x4(k_base + b_off, bh);
      mma_bf16(s[2 * np], ah, bh[0], bh[1]);
      mma_bf16(s[2 * np + 1], ah, bh[2], bh[3]);
      if (PARTS == 2) {
        ldsm_x4(k_base + ARR * 2 + b_off, bl);
        mma_bf16(s[2 * np], ah, bl[0], bl[1]);
        mma_bf16(s[2 * np + 1], ah, bl[2], bl[3]);
        mma_bf16(s[2 * np], al, bh[0], bh[1]);
        mma_bf16(s[2 * np + 1], al, bh[2], bh[3]);
      }
    }
  }

  // ---- bias + mask + softmax (fp32, rows live in a quad of lanes) ------------------------------------------
  const int r0 = 16 * warp + g, r1 = r0 + 8;
  const float* bh_ = bias + (long long)head * 64 * 64;
  const int lab0 = label[r0], lab1 = label[r1];
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = 8 * nt + 2 * qd;
    const float2 b0 = __ldg(reinterpret_cast<const float2*>(bh_ + r0 * 64 + c));
    const float2 b1 = __ldg(reinterpret_cast<const float2*>(bh_ + r1 * 64 + c));
    const int lc0 = label[c], lc1 = label[c + 1];
    s[nt][0] += b0.x + (lab0 != lc0 ? -100.f : 0.f);
    s[nt][1] += b0.y + (lab0 != lc1 ? -100.f : 0.f);
    s[nt][2] += b1.x + (lab1 != lc0 ? -100.f : 0.f);
    s[nt][3] += b1.y + (lab1 != lc1 ? -100.f : 0.f);
    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] = expf(s[nt][0] - mx0);
    s[nt][1] = expf(s[nt][1] - mx0);
    s[nt][2] = expf(s[nt][2] - mx1);
    s[nt][3] = expf(s[nt][3] - mx1);
    sum0 += s[nt][0] + s[nt][1];
    sum1 += s[nt][2] + s[nt][3];
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

  // ---- O = P v : accumulator fragments of S become A fragments ------------------------------------------------
  float o[NT_O][4];
#pragma unroll
  for (int nt = 0; nt < NT_O; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {  // 16 key tokens per step
    uint32_t ph[4], pl[4];
    split_pair(s[2 * ks][0] * inv0, s[2 * ks][1] * inv0, ph[0], pl[0]);
    split_pair(s[2 * ks][2] * inv1, s[2 * ks][3] * inv1, ph[1], pl[1]);
    split_pair(s[2 * ks + 1][0] * inv0, s[2 * ks + 1][1] * inv0, ph[2], pl[2]);
    split_pair(s[2 * ks + 1][2] * inv1, s[2 * ks + 1][3] * inv1, ph[3], pl[3]);
#pragma unroll
    for (int np = 0; np < NT_O / 2; ++np) {
      const uint32_t b_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + 8 * (2 * np + (lane >> 4))) * 2);
      uint32_t vh[4], vl[4];
      ldsm_x4_trans(v_base + b_off, vh);
      mma_bf16(o[2 * np], ph, vh[0], vh[1]);
      mma_bf16(o[2 * np + 1], ph, vh[2], vh[3]);
      if (PARTS == 2) {
        ldsm_x4_trans(v_base + ARR * 2 + b_off, vl);
        mma_bf16(o[2 * np], ph, vl[0], vl[1]);
        mma_bf16(o[2 * np + 1], ph, vl[2], vl[3]);
        mma_bf16(o[2 * np], pl, vh[0], vh[1]);
        mma_bf16(o[2 * np + 1], pl, vh[2], vh[3]);
      }
    }
  }
  __syncthreads();  // every warp is done with Q/K/V: the fp32 output tile may overlay Q|K
#pragma unroll
  for (int nt = 0; nt < NT_O; ++nt) {
    const int c = 8 * nt + 2 * qd;
    *reinterpret_cast<float2*>(Os + r0 * HD + c) = make_float2(o[nt][0], o[nt][1]);
    *reinterpret_cast<float2*>(Os + r1 * HD + c) = make_float2(o[nt][2], o[nt][3]);
  }
  __syncthreads();
  // coalesced scatter to image order + per-window token mean
  for (int idx = tid; idx < 64 * HD4; idx += 128) {
    const int t = idx / HD4, d4 = idx - t * HD4;
    const float4 v = *reinterpret_cast<const float4*>(Os + t * HD + d4 * 4);
    *reinterpret_cast<float4*>(out + (long long)rows[t] * ldo + head * HD + d4 * 4) = v;
  }
  if (tid < HD) {
    float sacc = 0.f;
#pragma unroll 8
    for (int t = 0; t < 64; ++t) sacc += Os[t * HD + tid];
    win_mean[(long long)win * C + head * HD + tid] = sacc * (1.0f / 64.0f);
  }
}

template <int HD, int PARTS>
static int launch(const float* qkv, int ldqkv, const float* bias, float* out, int ldo, float* win_mean, int B, int H,
                  int W, int C, int heads, int shift, int mask_H, int mask_y0, cudaStream_t st) {
  const size_t smem = (size_t)3 * PARTS * 64 * (HD + 8) * 2 + 2 * 64 * sizeof(int);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_mma_kernel<HD, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("window_attn(mma): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(B * (H / 8) * (W / 8), heads);
  window_attn_mma_kernel<HD, PARTS><<<grid, 128, smem, st>>>(qkv, ldqkv, bias, out, ldo, win_mean, H, W, C, shift, mask_H, mask_y0);
  return check_launch("window_attn(mma)");
}

int launch_window_attn_mma(const float* qkv, int ldqkv, const float* bias, float* out, int ldo, float* win_mean, int B,
                           int H, int W, int C, int heads, int shift, int parts, int mask_H, int mask_y0, cudaStream_t st) {
#define WA_CASE(HD)                                                                                          \
  case HD:                                                                                                   \
    return parts == 2 ? launch<HD, 2>(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, mask_H, mask_y0, st)   \
                      : launch<HD, 1>(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, mask_H, mask_y0, st);
  switch (C / heads) {
    WA_CASE(32)
    WA_CASE(48)
    WA_CASE(64)
    WA_CASE(96)
    default:
      set_error("window_attn: head_dim %d not in {32,48,64,96}", C / heads);
      return MPHSIR_ERR_INVALID;
  }
#undef WA_CASE
}

}  // namespace wa
}  // namespace mphsir
