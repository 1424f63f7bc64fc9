// Bandwidth-bound pieces of the global spectral ("transposed") attention and the GDFN:
//   * depthwise 3x3 conv on token-major data, optional GELU gate   (net/MP_HSIR.py:98, :236-237, :388-389)
//   * split-K Gram statistics over H*W: q^T k, sum q^2, sum k^2     (:104-107 with the L2 normalisation
//     applied AFTER the reduction: q^k^ = (q.k)/(|q||k|))
//   * softmax with temperature (:107-108) and folding of project_out (:113) into one C x C matrix
//     per sample:  out = W_out (A v)  ==  (W_out blockdiag(A)) v
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace mphsir {

// ---------------------------------------------------------------------------------------------
// depthwise 3x3: one thread per (1x4 pixel strip, 4 channels).  The 3x6 input patch of the strip is
// loaded once (128-bit loads along channels, neighbouring threads = neighbouring channels) and reused
// by the 4 outputs, the 9 weight vectors once per strip: 27 loads per 4 outputs instead of 72, which
// moves the kernel from L1-bandwidth-bound to HBM-bound.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dw_row(const float* __restrict__ rowp, long long ldx, int x0, int W,
                                       const float4 w0, const float4 w1, const float4 w2, float4 (&acc)[4]) {
  // rowp points at pixel (y+dy, x0) + channel offset; columns x0-1 .. x0+4
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

template <bool GATE>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const float* __restrict__ X, long long ldx,
                                                        const float* __restrict__ w9, float* __restrict__ Y,
                                                        long long ldy, int B, int H, int W, int C, int half) {
  const int c4n = (GATE ? half : C) >> 2;
  const int W4 = W >> 2;
  const long long total = (long long)B * H * W4 * c4n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4n) * 4;
    const long long strip = idx / c4n;
    const int x0 = (int)(strip % W4) * 4;
    const long long by = strip / W4;  // b*H + y
    const int y = (int)(by % H);
    const long long pix0 = by * W + x0;
    float4 a[4], g[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) a[o] = g[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      const float* rowp = X + (pix0 + (long long)dy * W) * ldx + c;
      const float* wr = w9 + (dy + 1) * 3 * C + c;
      dw_row(rowp, ldx, x0, W, ldg4(wr), ldg4(wr + C), ldg4(wr + 2 * C), a);
      if (GATE) dw_row(rowp + half, ldx, x0, W, ldg4(wr + half), ldg4(wr + C + half), ldg4(wr + 2 * C + half), g);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float4 r = a[o];
      if (GATE) {
        r.x = gelu_erf(r.x) * g[o].x; r.y = gelu_erf(r.y) * g[o].y;
        r.z = gelu_erf(r.z) * g[o].z; r.w = gelu_erf(r.w) * g[o].w;
      }
      *reinterpret_cast<float4*>(Y + (pix0 + o) * ldy + c) = r;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gram partials.  grid = (n_chunks, B*heads); CTA = (c/4)^2 threads, each a 4x4 block of q^T k.
// partial layout: [B*heads][n_chunks][c*c + 2c]  (G row-major [i][j], then sum q_i^2, sum k_j^2)
// ---------------------------------------------------------------------------------------------
constexpr int GT = 32;  // tokens staged per shared-memory tile

template <int CH>
__global__ void __launch_bounds__((CH / 4) * (CH / 4)) gram_partial_kernel(
    const float* __restrict__ q, long long ldq, int q_shared, const float* __restrict__ k, long long ldk,
    int k_shared, float* __restrict__ partial, int HW, int heads, int chunk) {
  constexpr int NB = CH / 4;
  constexpr int NTH = NB * NB;
  __shared__ __align__(16) float qs[GT][CH];
  __shared__ __align__(16) float ks[GT][CH];
  const int tid = threadIdx.x;
  const int bi = tid / NB, bj = tid - bi * NB;
  const int bh = blockIdx.y;
  const int b = bh / heads, h = bh - b * heads;
  const int t0 = blockIdx.x * chunk;
  const int t1 = min(HW, t0 + chunk);
  const float* qb = q + ((long long)(q_shared ? 0 : b) * HW) * ldq + h * CH;
  const float* kb = k + ((long long)(k_shared ? 0 : b) * HW) * ldk + h * CH;

  float g[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) g[i][j] = 0.f;
  float sq[4] = {0.f, 0.f, 0.f, 0.f}, sk[4] = {0.f, 0.f, 0.f, 0.f};

  for (int t = t0; t < t1; t += GT) {
    const int nt = min(GT, t1 - t);
    for (int idx = tid; idx < GT * NB; idx += NTH) {
      const int tt = idx / NB, c4 = idx - tt * NB;
      float4 vq = make_float4(0.f, 0.f, 0.f, 0.f), vk = vq;
      if (tt < nt) {
        vq = ldg4(qb + (long long)(t + tt) * ldq + c4 * 4);
        vk = ldg4(kb + (long long)(t + tt) * ldk + c4 * 4);
      }
      *reinterpret_cast<float4*>(&qs[tt][c4 * 4]) = vq;
      *reinterpret_cast<float4*>(&ks[tt][c4 * 4]) = vk;
    }
    __syncthreads();
#pragma unroll 8
    for (int tt = 0; tt < GT; ++tt) {
      const float4 a4 = *reinterpret_cast<const float4*>(&qs[tt][bi * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&ks[tt][bj * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) g[i][j] = fmaf(a[i], bb[j], g[i][j]);
      if (bj == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sq[i] = fmaf(a[i], a[i], sq[i]);
      }
      if (bi == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sk[j] = fmaf(bb[j], bb[j], sk[j]);
      }
    }
    __syncthreads();
  }
  float* dst = partial + ((long long)bh * gridDim.x + blockIdx.x) * (CH * CH + 2 * CH);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(dst + (bi * 4 + i) * CH + bj * 4) = make_float4(g[i][0], g[i][1], g[i][2], g[i][3]);
  if (bj == 0) *reinterpret_cast<float4*>(dst + CH * CH + bi * 4) = make_float4(sq[0], sq[1], sq[2], sq[3]);
  if (bi == 0) *reinterpret_cast<float4*>(dst + CH * CH + CH + bj * 4) = make_float4(sk[0], sk[1], sk[2], sk[3]);
}

static int gram_chunk(int B, int heads, int HW) {
  // aim for >= ~592 CTAs (4 per SM) but keep chunks in [256, 4096] tokens, multiples of GT
  long long total = (long long)B * heads * HW;
  long long chunk = total / 592;
  if (chunk < 256) chunk = 256;
  if (chunk > 4096) chunk = 4096;
  chunk = (chunk + GT - 1) / GT * GT;
  if (chunk > HW) chunk = (HW + GT - 1) / GT * GT;
  return (int)chunk;
}

// ---------------------------------------------------------------------------------------------
// stage-2 reduction of the Gram partials: grid = (ceil(per/64), B*heads), 256 threads = 64 elements x 4
// chunk groups, 4 independent accumulators per thread (the chunk count is in the hundreds at 512x512).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_reduce_kernel(const float* __restrict__ partial, int n_chunks, int per,
                                                          float* __restrict__ reduced) {
  __shared__ float red[4][64];
  const int e = blockIdx.x * 64 + (threadIdx.x & 63);
  const int g = threadIdx.x >> 6;
  const float* src = partial + (long long)blockIdx.y * n_chunks * per;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (e < per) {
    int ch = g;
    for (; ch + 12 < n_chunks; ch += 16) {
      a0 += __ldg(src + (long long)ch * per + e);
      a1 += __ldg(src + (long long)(ch + 4) * per + e);
      a2 += __ldg(src + (long long)(ch + 8) * per + e);
      a3 += __ldg(src + (long long)(ch + 12) * per + e);
    }
    for (; ch < n_chunks; ch += 4) a0 += __ldg(src + (long long)ch * per + e);
  }
  red[g][threadIdx.x & 63] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (g == 0 && e < per)
    reduced[(long long)blockIdx.y * per + e] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// reduce partials + normalise + temperature + row softmax.  grid = B*heads, 256 threads.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_softmax_kernel(const float* __restrict__ partial, int n_chunks,
                                                           const float* __restrict__ temperature,
                                                           float* __restrict__ attn, int heads, int c) {
  extern __shared__ float sm[];
  const int per = c * c + 2 * c;
  float* G = sm;  // [c*c + 2c]
  const int bh = blockIdx.x;
  const int h = bh % heads;
  const float* src = partial + (long long)bh * n_chunks * per;
  for (int e = threadIdx.x; e < per; e += blockDim.x) {
    float s = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) s += __ldg(src + (long long)ch * per + e);
    G[e] = s;
  }
  __syncthreads();
  // F.normalize: x / max(||x||, 1e-12)
  for (int e = threadIdx.x; e < 2 * c; e += blockDim.x) G[c * c + e] = fmaxf(sqrtf(G[c * c + e]), 1e-12f);
  __syncthreads();
  const float temp = __ldg(temperature + h);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < c; i += blockDim.x >> 5) {
    const float nq = G[c * c + i];
    float mx = -INFINITY;
    for (int j = lane; j < c; j += 32) {
      const float v = G[i * c + j] / (nq * G[c * c + c + j]) * temp;
      G[i * c + j] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < c; j += 32) {
      const float e = expf(G[i * c + j] - mx);
      G[i * c + j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < c; j += 32) attn[(long long)bh * c * c + i * c + j] = G[i * c + j] * inv;
  }
}

// ---------------------------------------------------------------------------------------------
// fold: Mt[b][h*c + j][o] = sum_i WoutT[h*c + i][o] * A[b,h][i][j].  grid = (B*heads, C/64), 256 thr.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spectral_fold_kernel(const float* __restrict__ attn,
                                                            const float* __restrict__ WoutT,
                                                            float* __restrict__ Mt, long long ldm,
                                                            long long m_batch_stride, int heads, int c) {
  extern __shared__ float sm[];
  const int C = heads * c;
  float* A = sm;            // [c][c]
  float* Wt = A + c * c;    // [c][64]
  const int bh = blockIdx.x;
  const int b = bh / heads, h = bh - b * heads;
  const int o0 = blockIdx.y * 64;
  for (int e = threadIdx.x; e < c * c; e += 256) A[e] = __ldg(attn + (long long)bh * c * c + e);
  for (int e = threadIdx.x; e < c * 64; e += 256) {
    const int i = e >> 6, o = e & 63;
    Wt[e] = (o0 + o < C) ? __ldg(WoutT + (long long)(h * c + i) * C + o0 + o) : 0.f;
  }
  __syncthreads();
  const int o = threadIdx.x & 63;
  for (int j = threadIdx.x >> 6; j < c; j += 4) {
    float s = 0.f;
    for (int i = 0; i < c; ++i) s = fmaf(Wt[i * 64 + o], A[i * c + j], s);
    if (o0 + o < C) Mt[(long long)b * m_batch_stride + (long long)(h * c + j) * ldm + o0 + o] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// finish: (reduced Gram, norms) -> normalise, temperature, row softmax -> fold project_out -> write the
// per-sample matrix both as fp32 "in x out" (SIMT engine) and as the bf16 hi/lo tensor-core image.
// grid = (B*heads, ceil(C/64)), 256 threads = 64 output columns x 4 groups of 8 input channels.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(256) spectral_finish_kernel(const float* __restrict__ gsum, int n_chunks,
                                                              float* __restrict__ gsum_out,
                                                              const float* __restrict__ temperature,
                                                              const float* __restrict__ WoutT, float* __restrict__ Mt,
                                                              long long ldm, long long m_batch_stride,
                                                              uint8_t* __restrict__ bimg, long long bimg_batch_bytes,
                                                              float* __restrict__ attn_out, int heads, int c) {
  extern __shared__ float sm[];
  const int C = heads * c;
  const int per = c * c + 2 * c;
  float* A = sm;             // [c][c] (+2c norms while normalising)
  float* Wt = A + per;       // [c][64]
  const int bh = blockIdx.x;
  const int b = bh / heads, h = bh - b * heads;
  const int o0 = blockIdx.y * 64;
  // n_chunks > 1: a handful of partials per (sample, head) are summed here (in chunk order: deterministic) instead of by a
  // separate reduction launch — both kernels are pure latency at these sizes
  for (int e = threadIdx.x; e < per; e += 256) {
    const float* src = gsum + (long long)bh * n_chunks * per + e;
    float a0 = 0.f, a1 = 0.f;
    int ch = 0;
    for (; ch + 1 < n_chunks; ch += 2) {
      a0 += __ldg(src + (long long)ch * per);
      a1 += __ldg(src + (long long)(ch + 1) * per);
    }
    if (ch < n_chunks) a0 += __ldg(src + (long long)ch * per);
    A[e] = a0 + a1;
    // the reduced statistics are part of the contract (scratch [B*heads*per]): the training backward reads them
    if (gsum_out != nullptr && blockIdx.y == 0) gsum_out[(long long)bh * per + e] = a0 + a1;
  }
  for (int e = threadIdx.x; e < c * 64; e += 256) {
    const int i = e >> 6, o = e & 63;
    Wt[e] = (o0 + o < C) ? __ldg(WoutT + (long long)(h * c + i) * C + o0 + o) : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * c; e += 256) A[c * c + e] = fmaxf(sqrtf(A[c * c + e]), 1e-12f);  // F.normalize eps
  __syncthreads();
  const float temp = __ldg(temperature + h);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows i = warp + 8 r of a warp run side by side, four independent reduction chains per step (c <= 128: <= 4 values per lane)
  for (int ib = warp; ib < c; ib += 32) {
    float v[4][4], mx[4], ssum[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = ib + 8 * r;
      mx[r] = -INFINITY;
      ssum[r] = 0.f;
      const float nq = i < c ? A[c * c + i] : 1.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = lane + 32 * q;
        v[r][q] = (i < c && j < c) ? A[i * c + j] / (nq * A[c * c + c + j]) * temp : -INFINITY;
        mx[r] = fmaxf(mx[r], v[r][q]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < 4; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[r][q] = (ib + 8 * r < c && lane + 32 * q < c) ? expf(v[r][q] - mx[r]) : 0.f;
        ssum[r] += v[r][q];
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < 4; ++r) ssum[r] += __shfl_xor_sync(0xffffffffu, ssum[r], o);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = ib + 8 * r;
      if (i >= c) continue;
      const float inv = 1.0f / ssum[r];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = lane + 32 * q;
        if (j >= c) continue;
        const float pv = v[r][q] * inv;
        A[i * c + j] = pv;
        if (attn_out != nullptr && blockIdx.y == 0) attn_out[(long long)bh * c * c + i * c + j] = pv;
      }
    }
  }
  __syncthreads();
  // fold: M[o][h*c + j] = sum_i Wout[o][h*c + i] * A[i][j]
  const int o = threadIdx.x & 63;
  const int Np = (C + 15) / 16 * 16, Ks = (C + 63) / 64;
  for (int j0 = 8 * (threadIdx.x >> 6); j0 < c; j0 += 32) {
    float acc[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.f;
#pragma unroll 4
    for (int i = 0; i < c; ++i) {
      const float w = Wt[i * 64 + o];
      const float4 a0 = *reinterpret_cast<const float4*>(&A[i * c + j0]);
      const float4 a1 = *reinterpret_cast<const float4*>(&A[i * c + j0 + 4]);
      acc[0] = fmaf(w, a0.x, acc[0]); acc[1] = fmaf(w, a0.y, acc[1]);
      acc[2] = fmaf(w, a0.z, acc[2]); acc[3] = fmaf(w, a0.w, acc[3]);
      acc[4] = fmaf(w, a1.x, acc[4]); acc[5] = fmaf(w, a1.y, acc[5]);
      acc[6] = fmaf(w, a1.z, acc[6]); acc[7] = fmaf(w, a1.w, acc[7]);
    }
    if (o0 + o >= C) continue;
    const int k = h * c + j0;
    if (Mt != nullptr) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) Mt[(long long)b * m_batch_stride + (long long)(k + jj) * ldm + o0 + o] = acc[jj];
    }
    if (bimg != nullptr) {
      uint4 hi, lo;
      split2_bf16(acc[0], acc[1], hi.x, lo.x);
      split2_bf16(acc[2], acc[3], hi.y, lo.y);
      split2_bf16(acc[4], acc[5], hi.z, lo.z);
      split2_bf16(acc[6], acc[7], hi.w, lo.w);
      uint8_t* img = bimg + (long long)b * bimg_batch_bytes;
      const int s_ = k >> 6, ch = (k & 63) >> 3;
      *reinterpret_cast<uint4*>(img + tc::bimg_offset(0, s_, o0 + o, ch, Np, Ks)) = hi;
      *reinterpret_cast<uint4*>(img + tc::bimg_offset(1, s_, o0 + o, ch, Np, Ks)) = lo;
    }
  }
  // zero the K tail of the image (C not a multiple of 64) once per output row
  if (bimg != nullptr && h == heads - 1 && (C & 63) != 0 && o0 + o < C) {
    uint8_t* img = bimg + (long long)b * bimg_batch_bytes;
    for (int k = C + 8 * (threadIdx.x >> 6); k < Ks * 64; k += 32) {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(img + tc::bimg_offset(0, k >> 6, o0 + o, (k & 63) >> 3, Np, Ks)) = z;
      *reinterpret_cast<uint4*>(img + tc::bimg_offset(1, k >> 6, o0 + o, (k & 63) >> 3, Np, Ks)) = z;
    }
  }
}

}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_dwconv3x3_fwd(const float* X, int ldx, const float* w9, float* Y, int ldy, int B, int H,
                                    int W, int C, int gate_half, void* stream) {
  MPHSIR_REQUIRE(X && w9 && Y, "dwconv3x3: null operand");
  MPHSIR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldx >= C, "dwconv3x3: bad shape C=%d ldx=%d ldy=%d", C, ldx, ldy);
  MPHSIR_REQUIRE(gate_half == 0 || (2 * gate_half == C && gate_half % 4 == 0), "dwconv3x3: gate_half=%d must be C/2 and a multiple of 4", gate_half);
  MPHSIR_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(w9)) & 15) == 0, "dwconv3x3: operands must be 16-byte aligned");
  MPHSIR_REQUIRE(W % 4 == 0, "dwconv3x3: W=%d must be a multiple of 4", W);
  const long long total = (long long)B * H * (W / 4) * ((gate_half ? gate_half : C) / 4);
  const int blocks = (int)((total + 255) / 256 > 148LL * 32 ? 148LL * 32 : (total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (gate_half)
    dwconv3x3_kernel<true><<<blocks, 256, 0, st>>>(X, ldx, w9, Y, ldy, B, H, W, C, gate_half);
  else
    dwconv3x3_kernel<false><<<blocks, 256, 0, st>>>(X, ldx, w9, Y, ldy, B, H, W, C, 0);
  return check_launch("dwconv3x3");
}

extern "C" size_t mphsir_gram_partial_floats(int B, int heads, int c, int HW, int* n_chunks) {
  const int chunk = gram_chunk(B, heads, HW);
  const int n = (HW + chunk - 1) / chunk;
  if (n_chunks) *n_chunks = n;
  return (size_t)B * heads * n * ((size_t)c * c + 2 * c);
}

extern "C" int mphsir_gram_partial_fwd(const float* q, int ldq, int q_shared, const float* k, int ldk,
                                       int k_shared, float* partial, int B, int HW, int heads, int c,
                                       void* stream) {
  MPHSIR_REQUIRE(q && k && partial, "gram_partial: null operand");
  MPHSIR_REQUIRE(B > 0 && HW > 0 && heads > 0 && ldq % 4 == 0 && ldk % 4 == 0, "gram_partial: bad shape");
  MPHSIR_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0, "gram_partial: operands must be 16-byte aligned");
  const int chunk = gram_chunk(B, heads, HW);
  dim3 grid((HW + chunk - 1) / chunk, B * heads);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (c) {
    case 32: gram_partial_kernel<32><<<grid, 64, 0, st>>>(q, ldq, q_shared, k, ldk, k_shared, partial, HW, heads, chunk); break;
    case 48: gram_partial_kernel<48><<<grid, 144, 0, st>>>(q, ldq, q_shared, k, ldk, k_shared, partial, HW, heads, chunk); break;
    case 64: gram_partial_kernel<64><<<grid, 256, 0, st>>>(q, ldq, q_shared, k, ldk, k_shared, partial, HW, heads, chunk); break;
    case 96: gram_partial_kernel<96><<<grid, 576, 0, st>>>(q, ldq, q_shared, k, ldk, k_shared, partial, HW, heads, chunk); break;
    default: MPHSIR_REQUIRE(false, "gram_partial: channels per head %d not in {32,48,64,96}", c);
  }
  return check_launch("gram_partial");
}

extern "C" int mphsir_gram_softmax_fwd(const float* partial, int n_chunks, const float* temperature, float* attn,
                                       float* scratch, int B, int heads, int c, void* stream) {
  MPHSIR_REQUIRE(partial && temperature && attn && n_chunks > 0 && B > 0 && heads > 0 && c > 0, "gram_softmax: bad arguments");
  const size_t smem = sizeof(float) * ((size_t)c * c + 2 * c);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int per = c * c + 2 * c;
  if (n_chunks > 8) {
    // two-stage: parallel reduction into chunk slot 0 of a scratch area placed right after the partials
    MPHSIR_REQUIRE(scratch != nullptr, "gram_softmax: scratch [B*heads*(c*c+2c)] is required when n_chunks > 8");
    dim3 grid((per + 63) / 64, B * heads);
    gram_reduce_kernel<<<grid, 256, 0, st>>>(partial, n_chunks, per, scratch);
    gram_softmax_kernel<<<B * heads, 256, smem, st>>>(scratch, 1, temperature, attn, heads, c);
  } else {
    gram_softmax_kernel<<<B * heads, 256, smem, st>>>(partial, n_chunks, temperature, attn, heads, c);
  }
  return check_launch("gram_softmax");
}

extern "C" int mphsir_spectral_fold_fwd(const float* attn, const float* WoutT, float* Mt, int ldm,
                                        long long m_batch_stride, int B, int heads, int c, void* stream) {
  MPHSIR_REQUIRE(attn && WoutT && Mt && B > 0 && heads > 0 && c > 0 && ldm >= heads * c, "spectral_fold: bad arguments");
  const int C = heads * c;
  const size_t smem = sizeof(float) * ((size_t)c * c + (size_t)c * 64);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(spectral_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("spectral_fold: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
  }
  dim3 grid(B * heads, (C + 63) / 64);
  spectral_fold_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(attn, WoutT, Mt, ldm, m_batch_stride, heads, c);
  return check_launch("spectral_fold");
}

extern "C" int mphsir_gram_reduce(const float* partial, int n_chunks, float* reduced, int B, int heads, int c, void* stream) {
  MPHSIR_REQUIRE(partial && reduced && n_chunks > 0 && B > 0 && heads > 0 && c > 0, "gram_reduce: bad arguments");
  const int per = c * c + 2 * c;
  dim3 grid((per + 63) / 64, B * heads);
  gram_reduce_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partial, n_chunks, per, reduced);
  return check_launch("gram_reduce");
}

extern "C" int mphsir_spectral_finish_fwd(const float* partial, int n_chunks, float* scratch, const float* temperature,
                                          const float* WoutT, float* Mt, int ldm, long long m_batch_stride, void* bimg,
                                          long long bimg_batch_bytes, float* attn_out, int B, int heads, int c,
                                          void* stream) {
  MPHSIR_REQUIRE(partial && temperature && WoutT && (Mt || bimg), "spectral_finish: null operand");
  MPHSIR_REQUIRE(n_chunks > 0 && B > 0 && heads > 0 && c > 0 && c % 8 == 0, "spectral_finish: bad shape (c=%d must be a multiple of 8)", c);
  MPHSIR_REQUIRE(Mt == nullptr || ldm >= heads * c, "spectral_finish: ldm too small");
  MPHSIR_REQUIRE(bimg == nullptr || (reinterpret_cast<uintptr_t>(bimg) & 127) == 0, "spectral_finish: image must be 128-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int per = c * c + 2 * c;
  const int C = heads * c;
  MPHSIR_REQUIRE(c <= 128, "spectral_finish: c=%d channels per head (at most 128)", c);
  const float* gsum = partial;
  int direct_chunks = n_chunks;
  if (n_chunks > 16) {
    direct_chunks = 1;
    MPHSIR_REQUIRE(scratch != nullptr, "spectral_finish: scratch [B*heads*(c*c+2c)] required when n_chunks > 16");
    dim3 grid((per + 63) / 64, B * heads);
    gram_reduce_kernel<<<grid, 256, 0, st>>>(partial, n_chunks, per, scratch);
    gsum = scratch;
  }
  const size_t smem = sizeof(float) * ((size_t)per + (size_t)c * 64);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(spectral_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) {
      set_error("spectral_finish: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  MPHSIR_REQUIRE(smem <= 96 * 1024, "spectral_finish: c=%d needs %zu B of shared memory", c, smem);
  dim3 grid(B * heads, (C + 63) / 64);
  spectral_finish_kernel<<<grid, 256, smem, st>>>(gsum, direct_chunks, (n_chunks > 1 && direct_chunks > 1) ? scratch : nullptr, temperature, WoutT, Mt, ldm, m_batch_stride,
                                                  reinterpret_cast<uint8_t*>(bimg), bimg_batch_bytes, attn_out, heads, c);
  return check_launch("spectral_finish");
}
