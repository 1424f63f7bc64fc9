// Local spectral branch: the low-rank spectral-prompt attention of PG_Spectral_Attention
// (net/MP_HSIR.py:132-155) collapsed to one fused kernel that turns the per-window token mean
// of the attention core into a per-window channel gate g[B_,C].
//
//   m   = projT^T core_mean + projb            (mean over tokens commutes with Spatial_Attention.proj)
//   pw  = softmax(promptT^T m)   [128]         (:136)   } m is never formed: the host folds proj into
//   dn  = downT^T m              [r]           (:137)   } promptT/downT (promptT := projT promptT, bias
//                                                          promptb := promptT^T projb; same for down)
//   sp  = pw @ param             [r]           (:139-140)
//   q   = qT^T sp ; [k;v] = kvT^T dn           (:142-144)
//   A   = softmax_j(q_i k_j r^-0.5) ; o_i = sum_j A_ij v_j     (:146-149, outer-product attention)
//   g   = upT^T (p2T^T o + p2b)                (:151-152)
//
// One CTA (160 threads) handles WPB = 8 windows so each weight element is fetched once per 8 windows;
// every weight is read "in x out" so that thread o reads W[k][o] coalesced.  The kernel is a chain of
// tiny dependent products, i.e. latency-bound: the r-sized weights are staged into shared memory in one
// coalesced burst up front (overlapping the core-mean load) and the C-long logit loop keeps 16 weight
// loads in flight, so no phase waits on a serial chain of L2 round trips.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace mphsir {

constexpr int LG_THREADS = 160;  // 128 prompt threads + 32 low-rank threads
constexpr int PLEN = 128;
constexpr int RMAX = 32;
constexpr int WPB = 8;  // windows per CTA: every weight element is fetched once per 8 windows

__global__ void __launch_bounds__(LG_THREADS) local_gate_kernel(const mphsir_local_gate_params p) {
  extern __shared__ float sm[];
  const int C = p.C, r = p.r;
  float* cm = sm;                 // [WPB][C] core mean
  float* pw = cm + WPB * C;       // [WPB][128]
  float* dn = pw + WPB * PLEN;    // [WPB][RMAX]
  float* sp = dn + WPB * RMAX;
  float* q = sp + WPB * RMAX;
  float* kv = q + WPB * RMAX;     // [WPB][2*RMAX]
  float* o = kv + WPB * 2 * RMAX;
  float* u = o + WPB * RMAX;
  float* red = u + WPB * RMAX;    // [WPB][8]
  float* s_param = red + WPB * 8;       // [128][r]
  float* s_qT = s_param + PLEN * r;     // [r][r]
  float* s_kvT = s_qT + r * r;          // [r][2r]
  float* s_p2T = s_kvT + 2 * r * r;     // [r][r]
  float* s_upT = s_p2T + r * r;         // [r][C]
  const int tid = threadIdx.x;
  const int w0 = blockIdx.x * WPB;
  const int nw = min(WPB, p.B_ - w0);

  for (int e = tid; e < WPB * C; e += LG_THREADS) {
    const int w = e / C, c = e - w * C;
    cm[e] = (w < nw) ? __ldg(p.core_mean + (long long)(w0 + w) * C + c) : 0.f;
  }
  for (int e = tid; e < PLEN * r; e += LG_THREADS) s_param[e] = __ldg(p.param + e);
  for (int e = tid; e < r * r; e += LG_THREADS) {
    s_qT[e] = __ldg(p.qT + e);
    s_p2T[e] = __ldg(p.p2T + e);
  }
  for (int e = tid; e < 2 * r * r; e += LG_THREADS) s_kvT[e] = __ldg(p.kvT + e);
  for (int e = tid; e < r * C; e += LG_THREADS) s_upT[e] = __ldg(p.upT + e);
  __syncthreads();
  // prompt logits (threads 0..127 == prompt index) and the low-rank projection (threads 128..128+r-1);
  // proj is pre-folded into promptT/downT, so both read the core mean directly.
  float logit[WPB];
  {
    const bool is_prompt = tid < PLEN;
    const int j = is_prompt ? tid : tid - PLEN;
    const bool active = is_prompt || j < r;
    const float* wsrc = is_prompt ? p.promptT : p.downT;
    const int ldw = is_prompt ? PLEN : r;
    const float b0 = active ? __ldg((is_prompt ? p.promptb : p.downb) + j) : 0.f;
#pragma unroll
    for (int w = 0; w < WPB; ++w) logit[w] = b0;
    if (active) {
      for (int k0 = 0; k0 < C; k0 += 16) {  // C % 16 == 0 (checked on the host)
        float wv[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) wv[kk] = __ldg(wsrc + (k0 + kk) * ldw + j);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
#pragma unroll
          for (int w = 0; w < WPB; ++w) logit[w] = fmaf(cm[w * C + k0 + kk], wv[kk], logit[w]);
      }
    }
    if (!is_prompt && active) {
#pragma unroll
      for (int w = 0; w < WPB; ++w) dn[w * RMAX + j] = logit[w];
    }
  }
  // softmax over the 128 logits of each window (4 warps)
  if (tid < PLEN) {
#pragma unroll
    for (int w = 0; w < WPB; ++w) {
      const float mx = warp_max(logit[w]);
      if ((tid & 31) == 0) red[w * 8 + (tid >> 5)] = mx;
    }
  }
  __syncthreads();
  float ex[WPB];
  if (tid < PLEN) {
#pragma unroll
    for (int w = 0; w < WPB; ++w) {
      const float mx = fmaxf(fmaxf(red[w * 8 + 0], red[w * 8 + 1]), fmaxf(red[w * 8 + 2], red[w * 8 + 3]));
      ex[w] = expf(logit[w] - mx);
      const float s = warp_sum(ex[w]);
      if ((tid & 31) == 0) red[w * 8 + 4 + (tid >> 5)] = s;
    }
  }
  __syncthreads();
  if (tid < PLEN) {
#pragma unroll
    for (int w = 0; w < WPB; ++w) {
      const float s = (red[w * 8 + 4] + red[w * 8 + 5]) + (red[w * 8 + 6] + red[w * 8 + 7]);
      pw[w * PLEN + tid] = ex[w] / s;
    }
  }
  __syncthreads();
  // the remaining r-sized steps: thread -> (window, index)
  for (int e = tid; e < WPB * r; e += LG_THREADS) {
    const int w = e / r, i = e - w * r;
    float a = 0.f;
    for (int k = 0; k < PLEN; ++k) a = fmaf(pw[w * PLEN + k], s_param[k * r + i], a);
    sp[w * RMAX + i] = a;
  }
  __syncthreads();
  for (int e = tid; e < WPB * 3 * r; e += LG_THREADS) {
    const int w = e / (3 * r), j = e - w * 3 * r;
    float a = 0.f;
    if (j < r) {
      for (int k = 0; k < r; ++k) a = fmaf(sp[w * RMAX + k], s_qT[k * r + j], a);
      q[w * RMAX + j] = a;
    } else {
      const int jj = j - r;
      for (int k = 0; k < r; ++k) a = fmaf(dn[w * RMAX + k], s_kvT[k * 2 * r + jj], a);
      kv[w * 2 * RMAX + jj] = a;
    }
  }
  __syncthreads();
  for (int e = tid; e < WPB * r; e += LG_THREADS) {
    const int w = e / r, i = e - w * r;
    const float* kvw = kv + w * 2 * RMAX;
    const float qi = q[w * RMAX + i] * rsqrtf((float)r);
    float mxl = -INFINITY;
    for (int j = 0; j < r; ++j) mxl = fmaxf(mxl, qi * kvw[j]);
    float den = 0.f, num = 0.f;
    for (int j = 0; j < r; ++j) {
      const float wgt = expf(qi * kvw[j] - mxl);
      den += wgt;
      num = fmaf(wgt, kvw[r + j], num);
    }
    o[w * RMAX + i] = num / den;
  }
  __syncthreads();
  for (int e = tid; e < WPB * r; e += LG_THREADS) {
    const int w = e / r, i = e - w * r;
    float a = __ldg(p.p2b + i);
    for (int k = 0; k < r; ++k) a = fmaf(o[w * RMAX + k], s_p2T[k * r + i], a);
    u[w * RMAX + i] = a;
  }
  __syncthreads();
  for (int c = tid; c < C; c += LG_THREADS) {
    float a[WPB];
#pragma unroll
    for (int w = 0; w < WPB; ++w) a[w] = 0.f;
    for (int k = 0; k < r; ++k) {
      const float wv = s_upT[k * C + c];
#pragma unroll
      for (int w = 0; w < WPB; ++w) a[w] = fmaf(u[w * RMAX + k], wv, a[w]);
    }
#pragma unroll
    for (int w = 0; w < WPB; ++w)
      if (w < nw) p.gate[(long long)(w0 + w) * C + c] = a[w];
  }
}

// ---------------------------------------------------------------------------------------------
// Two-step variant used by the engine: the C-long products (prompt logits and the low-rank projection,
// [B_, 128 + r] = core_mean [B_, C] x [promptT | downT] + [promptb | downb]) run as ONE GEMM on the tensor-core
// engine, and this kernel does the r-sized remainder with one warp per window.  The fused kernel above spends
// its time in C/16 dependent global round trips per CTA (the weights of a block are HBM-cold: they are used
// once per forward), which no amount of unrolling hides; here the small weights arrive in shared memory through
// one batch of bulk copies that overlaps the logits load, and every later step is shuffles + shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int LT_WARPS = 8;

__global__ void __launch_bounds__(LT_WARPS * 32) local_gate_tail_kernel(const float* __restrict__ logits, int ldl,
                                                                        const mphsir_local_gate_params p, int rp) {
  extern __shared__ __align__(128) float sm[];
  const int C = p.C, r = p.r;
  float* s_param = sm;                  // [128][r]
  float* s_qT = s_param + PLEN * r;     // [r][r]
  float* s_kvT = s_qT + r * r;          // [r][2r]
  float* s_p2T = s_kvT + 2 * r * r;     // [r][r]
  float* s_upT = s_p2T + r * r;         // [r][C]
  float* s_p2b = s_upT + r * C;         // [r] (padded to 4)
  float* s_pw = s_p2b + ((r + 3) & ~3); // [LT_WARPS][128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_pw + LT_WARPS * PLEN);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_a = tc::smem_u32(bar);
  if (tid == 0) {
    tc::mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t b_param = PLEN * r * 4, b_rr = r * r * 4, b_up = r * C * 4, b_b = r * 4;
    tc::mbar_expect_tx(bar_a, b_param + 4 * b_rr + b_up + b_b);
    tc::bulk_g2s(tc::smem_u32(s_param), p.param, b_param, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_qT), p.qT, b_rr, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_kvT), p.kvT, 2 * b_rr, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_p2T), p.p2T, b_rr, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_upT), p.upT, b_up, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_p2b), p.p2b, b_b, bar_a);
  }
  __syncthreads();  // the barrier is initialised before anyone polls it
  const int i = lane & (rp - 1);   // index inside the rank (rp = r rounded up to a power of two, <= 32)
  const int part = lane / rp;      // 32/rp lanes share one index and split the 128-long sum
  const int parts = 32 / rp;
  const bool iv = i < r;
  float* pw = s_pw + warp * PLEN;
  const float rs = rsqrtf((float)r);
  bool weights_ready = false;
  for (int win = blockIdx.x * LT_WARPS + warp; win < p.B_; win += gridDim.x * LT_WARPS) {
    const float* lrow = logits + (long long)win * ldl;
    const float4 lg = ldg4(lrow + 4 * lane);
    const float dn = iv ? __ldg(lrow + PLEN + i) : 0.f;
    // softmax over the 128 prompt logits (:136)
    const float mx = warp_max(fmaxf(fmaxf(lg.x, lg.y), fmaxf(lg.z, lg.w)));
    float4 e = make_float4(expf(lg.x - mx), expf(lg.y - mx), expf(lg.z - mx), expf(lg.w - mx));
    const float inv = 1.0f / warp_sum((e.x + e.y) + (e.z + e.w));
    e.x *= inv; e.y *= inv; e.z *= inv; e.w *= inv;
    __syncwarp();
    *reinterpret_cast<float4*>(pw + 4 * lane) = e;
    __syncwarp();
    if (!weights_ready) {
      tc::mbar_wait(bar_a, 0);
      weights_ready = true;
    }
    // sp = pw @ param (:139-140): lanes (i, part) sum a 128/parts slice, then fold the parts
    float sp = 0.f;
    {
      const int klen = PLEN / parts, k0 = part * klen;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (iv) {
        for (int k = k0; k < k0 + klen; k += 4) {
          a0 = fmaf(pw[k], s_param[k * r + i], a0);
          a1 = fmaf(pw[k + 1], s_param[(k + 1) * r + i], a1);
          a2 = fmaf(pw[k + 2], s_param[(k + 2) * r + i], a2);
          a3 = fmaf(pw[k + 3], s_param[(k + 3) * r + i], a3);
        }
      }
      sp = (a0 + a1) + (a2 + a3);
      for (int off = rp; off < 32; off <<= 1) sp += __shfl_xor_sync(0xffffffffu, sp, off);
    }
    // q = qT^T sp ; [k;v] = kvT^T dn (:142-144) -- every lane of an index group holds the same values
    float q = 0.f, kk = 0.f, vv = 0.f;
    for (int k = 0; k < r; ++k) {
      const float spk = __shfl_sync(0xffffffffu, sp, k);
      const float dnk = __shfl_sync(0xffffffffu, dn, k);
      if (iv) {
        q = fmaf(spk, s_qT[k * r + i], q);
        kk = fmaf(dnk, s_kvT[k * 2 * r + i], kk);
        vv = fmaf(dnk, s_kvT[k * 2 * r + r + i], vv);
      }
    }
    // outer-product attention (:146-149): A_ij = softmax_j(q_i k_j r^-0.5), o_i = sum_j A_ij v_j
    const float qi = q * rs;
    float mxl = -INFINITY;
    for (int j = 0; j < r; ++j) mxl = fmaxf(mxl, qi * __shfl_sync(0xffffffffu, kk, j));
    float den = 0.f, num = 0.f;
    for (int j = 0; j < r; ++j) {
      const float wgt = expf(qi * __shfl_sync(0xffffffffu, kk, j) - mxl);
      den += wgt;
      num = fmaf(wgt, __shfl_sync(0xffffffffu, vv, j), num);
    }
    const float o = iv ? num / den : 0.f;
    // u = p2T^T o + p2b (:151)
    float u = iv ? s_p2b[i] : 0.f;
    for (int k = 0; k < r; ++k) {
      const float ok = __shfl_sync(0xffffffffu, o, k);
      if (iv) u = fmaf(ok, s_p2T[k * r + i], u);
    }
    // g = upT^T u (:152): lane -> channels lane, lane+32, ... in groups of 4
    for (int c0 = 0; c0 < C; c0 += 128) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < r; ++k) {
        const float uk = __shfl_sync(0xffffffffu, u, k);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int c = c0 + 32 * m + lane;
          if (c < C) a[m] = fmaf(uk, s_upT[k * C + c], a[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int c = c0 + 32 * m + lane;
        if (c < C) p.gate[(long long)win * C + c] = a[m];
      }
    }
  }
  if (!weights_ready) tc::mbar_wait(bar_a, 0);  // never leave with bulk copies in flight
}

}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_local_gate_fwd(const mphsir_local_gate_params* p, void* stream) {
  MPHSIR_REQUIRE(p && p->core_mean && p->gate, "local_gate: null operand");
  MPHSIR_REQUIRE(p->promptT && p->promptb && p->downT && p->downb && p->param && p->qT && p->kvT && p->p2T && p->p2b && p->upT, "local_gate: null weight");
  MPHSIR_REQUIRE(p->B_ > 0 && p->C > 0 && p->r > 0 && p->r <= RMAX, "local_gate: bad shape B_=%d C=%d r=%d (r<=%d)", p->B_, p->C, p->r, RMAX);
  MPHSIR_REQUIRE(p->C % 16 == 0, "local_gate: C=%d must be a multiple of 16", p->C);
  const size_t smem = sizeof(float) * (WPB * (p->C + PLEN + 7 * RMAX + 8) + PLEN * p->r + 4 * p->r * p->r + p->r * p->C);
  MPHSIR_REQUIRE(smem <= 200 * 1024, "local_gate: C=%d r=%d needs %zu B of shared memory", p->C, p->r, smem);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(local_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("local_gate: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  local_gate_kernel<<<(p->B_ + WPB - 1) / WPB, LG_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(*p);
  return check_launch("local_gate");
}

extern "C" int mphsir_local_gate_tail_fwd(const float* logits, int ldl, const mphsir_local_gate_params* p, void* stream) {
  MPHSIR_REQUIRE(logits && p && p->gate, "local_gate_tail: null operand");
  MPHSIR_REQUIRE(p->param && p->qT && p->kvT && p->p2T && p->p2b && p->upT, "local_gate_tail: null weight");
  MPHSIR_REQUIRE(p->B_ > 0 && p->C > 0 && p->r >= 4 && p->r <= RMAX && p->r % 4 == 0, "local_gate_tail: bad shape B_=%d C=%d r=%d (r multiple of 4, <= %d)", p->B_, p->C, p->r, RMAX);
  MPHSIR_REQUIRE(p->C % 4 == 0 && ldl >= PLEN + p->r && ldl % 4 == 0, "local_gate_tail: C=%d, ldl=%d must be multiples of 4 with ldl >= 128 + r", p->C, ldl);
  const uintptr_t al = reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(p->param) | reinterpret_cast<uintptr_t>(p->qT) |
                       reinterpret_cast<uintptr_t>(p->kvT) | reinterpret_cast<uintptr_t>(p->p2T) | reinterpret_cast<uintptr_t>(p->p2b) |
                       reinterpret_cast<uintptr_t>(p->upT);
  MPHSIR_REQUIRE((al & 15) == 0, "local_gate_tail: logits and weights must be 16-byte aligned");
  int rp = 4;
  while (rp < p->r) rp <<= 1;
  const int r = p->r;
  const size_t smem = sizeof(float) * ((size_t)PLEN * r + 4 * r * r + (size_t)r * p->C + ((r + 3) & ~3) + LT_WARPS * PLEN) + 16;
  MPHSIR_REQUIRE(smem <= 160 * 1024, "local_gate_tail: C=%d r=%d needs %zu B of shared memory", p->C, r, smem);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(local_gate_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) {
      set_error("local_gate_tail: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  // one window per warp while the grid still fits the machine in one wave (the kernel is a latency chain;
  // re-staging ~20 KB of weights per CTA out of L2 is cheap next to a second trip through that chain)
  int grid = (p->B_ + LT_WARPS - 1) / LT_WARPS;
  if (grid > 8 * 148) grid = 8 * 148;
  local_gate_tail_kernel<<<grid, LT_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(logits, ldl, *p, rp);
  return check_launch("local_gate_tail");
}
