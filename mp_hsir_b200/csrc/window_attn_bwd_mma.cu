// Backward of the window attention core on tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate) — the
// tensor-core twin of window_attn_bwd.cu, mirroring the forward kernel window_attn_mma.cu.
//
// One CTA (4 warps) walks a strided set of (shifted) windows of one head.  Per window:
//   gather q*scale, k, v, dO (roll / window_partition folded into the addressing) -> bf16 hi (+lo) in shared memory
//   S = q k^T + bias + mask, P = softmax(S)            (registers; warp w owns rows 16w..16w+15)
//   dP = dO v^T,  dS = P o (dP - rowsum(P o dP))       (registers, same fragment layout as S)
//   dQ = scale * dS k                                   (dS accumulator fragments re-used as A fragments)
//   P, dS -> shared memory (bf16);  dV = P^T dO,  dK = dS^T (scale q)   (ldmatrix.trans reads [token][token'] as A)
// dS is accumulated over the CTA's windows in registers: one [64,64] relative-position-bias partial per CTA.
// PARTS = 2 evaluates every product as hi*hi + hi*lo + lo*hi (fp32-grade), PARTS = 1 is plain bf16.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mphsir {
namespace wabm {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
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

// acc[NT][4] (+)= A(rows 16*warp.., K = 16*KSTEPS) x B, A via ldmatrix (row-major [m][k]) and B stored [n][k]
template <int KSTEPS, int PARTS, int LD, int ARR>
__device__ __forceinline__ void mma_rowA_colB(float (&acc)[8][4], uint32_t a_base, uint32_t b_base, int warp, int lane) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    uint32_t ah[4], al[4];
    const uint32_t a_off = (uint32_t)(((16 * warp + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + 16 * ks + (lane >> 4) * 8) * 2);
    ldsm_x4(a_base + a_off, ah);
    if (PARTS == 2) ldsm_x4(a_base + ARR * 2 + a_off, al);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const uint32_t b_off = (uint32_t)(((8 * (2 * np + (lane >> 4)) + (lane & 7)) * LD + 16 * ks + ((lane >> 3) & 1) * 8) * 2);
      uint32_t bh[4], bl[4];
      ldsm_x4(b_base + b_off, bh);
      mma_bf16(acc[2 * np], ah, bh[0], bh[1]);
      mma_bf16(acc[2 * np + 1], ah, bh[2], bh[3]);
      if (PARTS == 2) {
        ldsm_x4(b_base + ARR * 2 + b_off, bl);
        mma_bf16(acc[2 * np], ah, bl[0], bl[1]);
        mma_bf16(acc[2 * np + 1], ah, bl[2], bl[3]);
        mma_bf16(acc[2 * np], al, bh[0], bh[1]);
        mma_bf16(acc[2 * np + 1], al, bh[2], bh[3]);
      }
    }
  }
}

template <int HD, int PARTS>
__global__ void __launch_bounds__(128) window_attn_bwd_mma_kernel(const float* __restrict__ qkv, long long ldqkv,
                                                                  const float* __restrict__ bias, const float* __restrict__ dO,
                                                                  long long ldo, float* __restrict__ dqkv, long long lddq,
                                                                  float* __restrict__ dbias_partial, int n_win, int H, int W,
                                                                  int C, int heads, int shift) {
  constexpr int LD = HD + 8;     // bf16 elements per row of the [token][channel] operand arrays
  constexpr int ARR = 64 * LD;
  constexpr int LDP = 64 + 8;    // [token][token'] arrays (P, dS)
  constexpr int PARR = 64 * LDP;
  constexpr int KS = HD / 16;
  constexpr int NT_O = HD / 8;
  constexpr int HD4 = HD / 4;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Ks_ = Qs + PARTS * ARR;
  __nv_bfloat16* Vs = Ks_ + PARTS * ARR;
  __nv_bfloat16* Gs = Vs + PARTS * ARR;
  __nv_bfloat16* Ps = Gs + PARTS * ARR;
  __nv_bfloat16* Ds = Ps + PARTS * PARR;
  int* rows = reinterpret_cast<int*>(Ds + PARTS * PARR);
  int* label = rows + 64;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;
  const int nWx = W >> 3, nW = (H >> 3) * nWx;
  const int g = lane >> 2, qd = lane & 3;
  const int r0 = 16 * warp + g, r1 = r0 + 8;
  const float scale = rsqrtf((float)HD);
  const uint32_t q_base = (uint32_t)__cvta_generic_to_shared(Qs);
  const uint32_t k_base = (uint32_t)__cvta_generic_to_shared(Ks_);
  const uint32_t v_base = (uint32_t)__cvta_generic_to_shared(Vs);
  const uint32_t g_base = (uint32_t)__cvta_generic_to_shared(Gs);
  const uint32_t p_base = (uint32_t)__cvta_generic_to_shared(Ps);
  const uint32_t d_base = (uint32_t)__cvta_generic_to_shared(Ds);
  const float* bh_ = bias + (long long)head * 64 * 64;

  float bacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) bacc[nt][e] = 0.f;

  for (int win = blockIdx.x; win < n_win; win += gridDim.x) {
    const int b = win / nW;
    const int wrem = win - b * nW;
    const int wi = wrem / nWx, wj = wrem - wi * nWx;
    __syncthreads();  // the previous window's shared-memory operands are dead
    if (tid < 64) {
      const int r = tid >> 3, c = tid & 7;
      const int ys = wi * 8 + r, xs = wj * 8 + c;
      int y = ys + shift, x = xs + shift;
      if (y >= H) y -= H;
      if (x >= W) x -= W;
      rows[tid] = (b * H + y) * W + x;
      const int rh = (ys >= H - 8) + (ys >= H - 4);
      const int rw = (xs >= W - 8) + (xs >= W - 4);
      label[tid] = shift ? 3 * rh + rw : 0;
    }
    __syncthreads();
    // ---- gather + split q (pre-scaled), k, v, dO ---------------------------------------------------
    constexpr int ITERS = HD / 8;                  // (64 tokens x HD/4 quads) / 128 threads
    constexpr int GB = (ITERS % 2 == 0) ? 2 : 3;   // HD 32/64/96 -> 2, HD 48 -> 3
#pragma unroll
    for (int it0 = 0; it0 < ITERS; it0 += GB) {
      float4 q[GB], k[GB], v[GB], go[GB];
#pragma unroll
      for (int u = 0; u < GB; ++u) {
        const int idx = tid + (it0 + u) * 128;
        const int t = idx / HD4, d4 = idx - t * HD4;
        const float* base = qkv + (long long)rows[t] * ldqkv + head * HD + d4 * 4;
        q[u] = ldg4(base);
        k[u] = ldg4(base + C);
        v[u] = ldg4(base + 2 * C);
        go[u] = ldg4(dO + (long long)rows[t] * ldo + head * HD + d4 * 4);
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
        split_pair(go[u].x, go[u].y, hi.x, lo.x); split_pair(go[u].z, go[u].w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Gs + off) = hi;
        if (PARTS == 2) *reinterpret_cast<uint2*>(Gs + ARR + off) = lo;
      }
    }
    __syncthreads();

    // ---- S = q k^T, dP = dO v^T : warp owns rows 16*warp .. +15 ----------------------------------------
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
    mma_rowA_colB<KS, PARTS, LD, ARR>(s, q_base, k_base, warp, lane);
    mma_rowA_colB<KS, PARTS, LD, ARR>(dp, g_base, v_base, warp, lane);

    // ---- P = softmax(S + bias + mask) ---------------------------------------------------------------------
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
    // ---- dS = P o (dP - rowsum(P o dP)) ; s <- P, dp <- dS ---------------------------------------------------
    float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] *= inv0;
      s[nt][1] *= inv0;
      s[nt][2] *= inv1;
      s[nt][3] *= inv1;
      dot0 = fmaf(s[nt][0], dp[nt][0], fmaf(s[nt][1], dp[nt][1], dot0));
      dot1 = fmaf(s[nt][2], dp[nt][2], fmaf(s[nt][3], dp[nt][3], dot1));
    }
    dot0 += __shfl_xor_sync(0xffffffffu, dot0, 1);
    dot0 += __shfl_xor_sync(0xffffffffu, dot0, 2);
    dot1 += __shfl_xor_sync(0xffffffffu, dot1, 1);
    dot1 += __shfl_xor_sync(0xffffffffu, dot1, 2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      dp[nt][0] = s[nt][0] * (dp[nt][0] - dot0);
      dp[nt][1] = s[nt][1] * (dp[nt][1] - dot0);
      dp[nt][2] = s[nt][2] * (dp[nt][2] - dot1);
      dp[nt][3] = s[nt][3] * (dp[nt][3] - dot1);
#pragma unroll
      for (int e = 0; e < 4; ++e) bacc[nt][e] += dp[nt][e];
    }
    // ---- P, dS -> shared memory [token][token'] for the transposed products --------------------------------------
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = 8 * nt + 2 * qd;
      uint32_t hi, lo;
      split_pair(s[nt][0], s[nt][1], hi, lo);
      *reinterpret_cast<uint32_t*>(Ps + r0 * LDP + c) = hi;
      if (PARTS == 2) *reinterpret_cast<uint32_t*>(Ps + PARR + r0 * LDP + c) = lo;
      split_pair(s[nt][2], s[nt][3], hi, lo);
      *reinterpret_cast<uint32_t*>(Ps + r1 * LDP + c) = hi;
      if (PARTS == 2) *reinterpret_cast<uint32_t*>(Ps + PARR + r1 * LDP + c) = lo;
      split_pair(dp[nt][0], dp[nt][1], hi, lo);
      *reinterpret_cast<uint32_t*>(Ds + r0 * LDP + c) = hi;
      if (PARTS == 2) *reinterpret_cast<uint32_t*>(Ds + PARR + r0 * LDP + c) = lo;
      split_pair(dp[nt][2], dp[nt][3], hi, lo);
      *reinterpret_cast<uint32_t*>(Ds + r1 * LDP + c) = hi;
      if (PARTS == 2) *reinterpret_cast<uint32_t*>(Ds + PARR + r1 * LDP + c) = lo;
    }
    // ---- dQ = scale * dS k : dS accumulator fragments are the A fragments --------------------------------------------
    {
      float o[NT_O][4];
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t ph[4], pl[4];
        split_pair(dp[2 * ks][0], dp[2 * ks][1], ph[0], pl[0]);
        split_pair(dp[2 * ks][2], dp[2 * ks][3], ph[1], pl[1]);
        split_pair(dp[2 * ks + 1][0], dp[2 * ks + 1][1], ph[2], pl[2]);
        split_pair(dp[2 * ks + 1][2], dp[2 * ks + 1][3], ph[3], pl[3]);
#pragma unroll
        for (int np = 0; np < NT_O / 2; ++np) {
          const uint32_t b_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + 8 * (2 * np + (lane >> 4))) * 2);
          uint32_t kh[4], kl[4];
          ldsm_x4_trans(k_base + b_off, kh);
          mma_bf16(o[2 * np], ph, kh[0], kh[1]);
          mma_bf16(o[2 * np + 1], ph, kh[2], kh[3]);
          if (PARTS == 2) {
            ldsm_x4_trans(k_base + ARR * 2 + b_off, kl);
            mma_bf16(o[2 * np], ph, kl[0], kl[1]);
            mma_bf16(o[2 * np + 1], ph, kl[2], kl[3]);
            mma_bf16(o[2 * np], pl, kh[0], kh[1]);
            mma_bf16(o[2 * np + 1], pl, kh[2], kh[3]);
          }
        }
      }
      float* d0 = dqkv + (long long)rows[r0] * lddq + head * HD + 2 * qd;
      float* d1 = dqkv + (long long)rows[r1] * lddq + head * HD + 2 * qd;
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        *reinterpret_cast<float2*>(d0 + 8 * nt) = make_float2(o[nt][0] * scale, o[nt][1] * scale);
        *reinterpret_cast<float2*>(d1 + 8 * nt) = make_float2(o[nt][2] * scale, o[nt][3] * scale);
      }
    }
    __syncthreads();  // P and dS of every warp are in shared memory
    // ---- dV = P^T dO, dK = dS^T (scale q): warp owns key rows 16*warp .. +15; A[m=token'][k=token] read transposed --------
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t a_src = which == 0 ? p_base : d_base;
      const uint32_t b_src = which == 0 ? g_base : q_base;
      float o[NT_O][4];
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // 16 query tokens per step
        uint32_t ah[4], al[4];
        const uint32_t a_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 4) & 1) * 8) * LDP + 16 * warp + ((lane >> 3) & 1) * 8) * 2);
        ldsm_x4_trans(a_src + a_off, ah);
        if (PARTS == 2) ldsm_x4_trans(a_src + PARR * 2 + a_off, al);
#pragma unroll
        for (int np = 0; np < NT_O / 2; ++np) {
          const uint32_t b_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + 8 * (2 * np + (lane >> 4))) * 2);
          uint32_t bh[4], bl[4];
          ldsm_x4_trans(b_src + b_off, bh);
          mma_bf16(o[2 * np], ah, bh[0], bh[1]);
          mma_bf16(o[2 * np + 1], ah, bh[2], bh[3]);
          if (PARTS == 2) {
            ldsm_x4_trans(b_src + ARR * 2 + b_off, bl);
            mma_bf16(o[2 * np], ah, bl[0], bl[1]);
            mma_bf16(o[2 * np + 1], ah, bl[2], bl[3]);
            mma_bf16(o[2 * np], al, bh[0], bh[1]);
            mma_bf16(o[2 * np + 1], al, bh[2], bh[3]);
          }
        }
      }
      // which == 0 -> dV (columns [2C, 3C)), which == 1 -> dK (columns [C, 2C))
      const int col0 = (which == 0 ? 2 * C : C) + head * HD + 2 * qd;
      float* d0 = dqkv + (long long)rows[r0] * lddq + col0;
      float* d1 = dqkv + (long long)rows[r1] * lddq + col0;
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        *reinterpret_cast<float2*>(d0 + 8 * nt) = make_float2(o[nt][0], o[nt][1]);
        *reinterpret_cast<float2*>(d1 + 8 * nt) = make_float2(o[nt][2], o[nt][3]);
      }
    }
  }
  // ---- relative-position-bias gradient partial of this CTA ---------------------------------------------
  float* dst = dbias_partial + ((long long)blockIdx.x * heads + head) * 64 * 64;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = 8 * nt + 2 * qd;
    *reinterpret_cast<float2*>(dst + r0 * 64 + c) = make_float2(bacc[nt][0], bacc[nt][1]);
    *reinterpret_cast<float2*>(dst + r1 * 64 + c) = make_float2(bacc[nt][2], bacc[nt][3]);
  }
}

template <int HD, int PARTS>
static int launch(const float* qkv, int ldqkv, const float* bias, const float* dO, int ldo, float* dqkv, int lddq,
                  float* partial, int groups, int B, int H, int W, int C, int heads, int shift, cudaStream_t st) {
  const size_t smem = (size_t)PARTS * (4 * 64 * (HD + 8) + 2 * 64 * 72) * 2 + 2 * 64 * sizeof(int);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_bwd_mma_kernel<HD, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("window_attn_bwd(mma): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(groups, heads);
  window_attn_bwd_mma_kernel<HD, PARTS><<<grid, 128, smem, st>>>(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, partial,
                                                                 B * (H / 8) * (W / 8), H, W, C, heads, shift);
  return check_launch("window_attn_bwd(mma)");
}

int launch_window_attn_bwd_mma(const float* qkv, int ldqkv, const float* bias, const float* dO, int ldo, float* dqkv, int lddq,
                               float* partial, int groups, int B, int H, int W, int C, int heads, int shift, int parts,
                               cudaStream_t st) {
  const int hd = C / heads;
#define WABM(HD_)                                                                                                          \
  case HD_:                                                                                                                \
    return parts == 2 ? launch<HD_, 2>(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, partial, groups, B, H, W, C, heads, shift, st) \
                      : launch<HD_, 1>(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, partial, groups, B, H, W, C, heads, shift, st)
  switch (hd) {
    WABM(32);
    WABM(48);
    WABM(64);
    WABM(96);
    default:
      set_error("window_attn_bwd(mma): head_dim %d not supported (32, 48, 64, 96)", hd);
      return MPHSIR_ERR_INVALID;
  }
#undef WABM
}

}  // namespace wabm
}  // namespace mphsir
