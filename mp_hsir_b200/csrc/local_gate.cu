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


// ---------------------------------------------------------------------------------------------
// One-launch form used by the inference engine.  The two-step form above costs two kernels of pure latency per block
// (a [B_,C] x [C,128+r] GEMM through the 608-thread tcgen05 engine: ~17 us of prologue for 0.15 GFLOP; then the tail:
// ~12 us) on the critical path between the window attention and the projection GEMM — 0.6 ms per 512x512 cube, 6 % of a
// 16-patch batch.  Here a CTA owns 8 x NW windows: the C-long products run as an fp32 register-tiled product (warp = NW
// windows, lane = 5 of the 128+r outputs; the folded weights stream through a 3-slot ring of bulk copies, all in flight at
// once: they are HBM-cold, used once per forward), the logits stay in shared memory, and the same warp finishes its windows with the r-sized chain of the tail
// kernel.  The small weights arrive by bulk copies that overlap the C-long phase.
// ---------------------------------------------------------------------------------------------
constexpr int LG2_LDL = PLEN + RMAX;
constexpr int LG2_KC = 32;        // k rows per weight chunk
constexpr int LG2_NS = 3;         // chunk ring slots

template <int NW>   // windows per warp: 4, 2 or 1 — fewer when the windows would not fill the machine otherwise (the kernel is a latency chain)
__global__ void __launch_bounds__(256) local_gate2_kernel(const mphsir_local_gate_params p, int rp) {
  constexpr int WIN = 8 * NW;   // windows per CTA
  extern __shared__ __align__(128) float sm[];
  const int C = p.C, r = p.r;
  float* s_param = sm;                  // [128][r]
  float* s_qT = s_param + PLEN * r;     // [r][r]
  float* s_kvT = s_qT + r * r;          // [r][2r]
  float* s_p2T = s_kvT + 2 * r * r;     // [r][r]
  float* s_upT = s_p2T + r * r;         // [r][C]
  float* s_p2b = s_upT + r * C;         // [r] (padded to 4)
  float* s_pw = s_p2b + ((r + 3) & ~3); // [8][128]
  float* s_lg = s_pw + 8 * PLEN;        // [32][160] logits | low-rank projection
  float* s_m = s_lg + WIN * LG2_LDL;  // [32][C] window means
  float* s_ring = s_m + WIN * C;      // LG2_NS slots of [32 k][128 + r]: K-chunks of [promptT | downT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_ring + LG2_NS * LG2_KC * (PLEN + r));
  uint64_t* full = bar + 1;               // [LG2_NS]
  uint64_t* empty = full + LG2_NS;        // [LG2_NS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_a = tc::smem_u32(bar);
  const int w0 = blockIdx.x * WIN;
  const int nw = min(WIN, p.B_ - w0);
  const int nch = C / LG2_KC;   // C % 32 == 0 (checked on the host)
  const uint32_t ch_p = LG2_KC * PLEN * 4, ch_d = LG2_KC * r * 4;
  auto issue_chunk = [&](int ch) {   // thread 0: the folded weights are HBM-cold (used once per forward) — all slots in flight at once
    const int slot = ch % LG2_NS;
    const uint32_t fb = tc::smem_u32(&full[slot]);
    float* dst = s_ring + (size_t)slot * LG2_KC * (PLEN + r);
    tc::mbar_expect_tx(fb, ch_p + ch_d);
    tc::bulk_g2s(tc::smem_u32(dst), p.promptT + (size_t)ch * LG2_KC * PLEN, ch_p, fb);
    tc::bulk_g2s(tc::smem_u32(dst + LG2_KC * PLEN), p.downT + (size_t)ch * LG2_KC * r, ch_d, fb);
  };
  if (tid == 0) {
    tc::mbar_init(bar_a, 1);
    for (int i = 0; i < LG2_NS; ++i) {
      tc::mbar_init(tc::smem_u32(&full[i]), 1);
      tc::mbar_init(tc::smem_u32(&empty[i]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int ch = 0; ch < min(nch, LG2_NS); ++ch) issue_chunk(ch);
    const uint32_t b_param = PLEN * r * 4, b_rr = r * r * 4, b_up = r * C * 4, b_b = r * 4;
    tc::mbar_expect_tx(bar_a, b_param + 4 * b_rr + b_up + b_b);
    tc::bulk_g2s(tc::smem_u32(s_param), p.param, b_param, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_qT), p.qT, b_rr, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_kvT), p.kvT, 2 * b_rr, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_p2T), p.p2T, b_rr, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_upT), p.upT, b_up, bar_a);
    tc::bulk_g2s(tc::smem_u32(s_p2b), p.p2b, b_b, bar_a);
  }
  // window means of this CTA (C % 4 == 0): coalesced 128-bit loads
  {
    const int c4n = C >> 2;
    for (int e = tid; e < WIN * c4n; e += 256) {
      const int w = e / c4n, c4 = e - w * c4n;
      reinterpret_cast<float4*>(s_m)[e] = w < nw ? ldg4(p.core_mean + (long long)(w0 + w) * C + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();  // barrier initialised, means staged
  // ---- C-long products: [4 windows of the warp] x [outputs lane + 32 j, j < 4 (prompt logits); 128 + lane (low rank)] ----
  {
    float acc[NW][5];
    const bool dv = lane < r;
    {
      float b[5];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = __ldg(p.promptb + lane + 32 * j);
      b[4] = dv ? __ldg(p.downb + lane) : 0.f;
#pragma unroll
      for (int i = 0; i < NW; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = b[j];
    }
    const float* mrow = s_m + (NW * warp) * C;
    for (int ch = 0; ch < nch; ++ch) {
      const int slot = ch % LG2_NS;
      tc::mbar_wait(tc::smem_u32(&full[slot]), (ch / LG2_NS) & 1);
      const float* wp = s_ring + (size_t)slot * LG2_KC * (PLEN + r);
      const float* wd = wp + LG2_KC * PLEN;
      // register double buffer: the weights of step s + 1 are in flight while step s is accumulated (two warps per scheduler
      // cannot hide a load-wait-compute sequence: 3.2k clk per chunk measured without it)
      float wv[2][4][5];
      auto load_w = [&](float (&dst)[4][5], int kc) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[kk][j] = wp[(kc + kk) * PLEN + lane + 32 * j];
          dst[kk][4] = dv ? wd[(kc + kk) * r + lane] : 0.f;
        }
      };
      load_w(wv[0], 0);
#pragma unroll
      for (int st = 0; st < LG2_KC / 4; ++st) {
        const int kc = 4 * st;
        if (st + 1 < LG2_KC / 4) load_w(wv[(st + 1) & 1], kc + 4);
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const float4 mv = *reinterpret_cast<const float4*>(mrow + i * C + ch * LG2_KC + kc);   // same address in every lane: broadcast
          const float m4[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int j = 0; j < 5; ++j) acc[i][j] = fmaf(m4[kk], wv[st & 1][kk][j], acc[i][j]);
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty[slot]));
      if (tid == 0 && ch + LG2_NS < nch) {   // refill the slot once all 8 warps have left it
        tc::mbar_wait(tc::smem_u32(&empty[slot]), (ch / LG2_NS) & 1);
        issue_chunk(ch + LG2_NS);
      }
      __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      float* lrow = s_lg + (NW * warp + i) * LG2_LDL;
#pragma unroll
      for (int j = 0; j < 4; ++j) lrow[lane + 32 * j] = acc[i][j];
      lrow[PLEN + lane] = acc[i][4];
    }
  }
  __syncwarp();   // a warp finishes the windows whose logits it has just written
  // ---- r-sized remainder (:136-152), as in local_gate_tail_kernel ----
  const int i = lane & (rp - 1);   // index inside the rank (rp = r rounded up to a power of two, <= 32)
  const int part = lane / rp;      // 32/rp lanes share one index and split the 128-long sum
  const int parts = 32 / rp;
  const bool iv = i < r;
  const float rs = rsqrtf((float)r);
  tc::mbar_wait(bar_a, 0);
  // The chain of one window is ~600 dependent shuffle / exp steps (6k clk measured): the warp's four windows run through it
  // side by side, four independent chains per step.
  const float* lbase = s_lg + (NW * warp) * LG2_LDL;
  float4 e4[NW];
  float dn[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    e4[w] = *reinterpret_cast<const float4*>(lbase + w * LG2_LDL + 4 * lane);
    dn[w] = iv ? lbase[w * LG2_LDL + PLEN + i] : 0.f;
  }
  {
    float mx[NW], sum[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) mx[w] = fmaxf(fmaxf(e4[w].x, e4[w].y), fmaxf(e4[w].z, e4[w].w));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int w = 0; w < NW; ++w) mx[w] = fmaxf(mx[w], __shfl_xor_sync(0xffffffffu, mx[w], o));
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      e4[w] = make_float4(expf(e4[w].x - mx[w]), expf(e4[w].y - mx[w]), expf(e4[w].z - mx[w]), expf(e4[w].w - mx[w]));
      sum[w] = (e4[w].x + e4[w].y) + (e4[w].z + e4[w].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int w = 0; w < NW; ++w) sum[w] += __shfl_xor_sync(0xffffffffu, sum[w], o);
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float inv = 1.0f / sum[w];
      e4[w].x *= inv; e4[w].y *= inv; e4[w].z *= inv; e4[w].w *= inv;
    }
  }
  // the softmax weights of the four windows overwrite their logits rows (own warp only), read back as pw[w][k]
  __syncwarp();
#pragma unroll
  for (int w = 0; w < NW; ++w) *reinterpret_cast<float4*>(s_lg + (NW * warp + w) * LG2_LDL + 4 * lane) = e4[w];
  __syncwarp();
  float sp[NW];
  {
    const int klen = PLEN / parts, k0 = part * klen;
    float a0[NW], a1[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) a0[w] = a1[w] = 0.f;
    if (iv) {
#pragma unroll 4
      for (int k = k0; k < k0 + klen; k += 2) {
        const float p0 = s_param[k * r + i], p1 = s_param[(k + 1) * r + i];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          a0[w] = fmaf(lbase[w * LG2_LDL + k], p0, a0[w]);
          a1[w] = fmaf(lbase[w * LG2_LDL + k + 1], p1, a1[w]);
        }
      }
    }
#pragma unroll
    for (int w = 0; w < NW; ++w) sp[w] = a0[w] + a1[w];
    for (int off = rp; off < 32; off <<= 1)
#pragma unroll
      for (int w = 0; w < NW; ++w) sp[w] += __shfl_xor_sync(0xffffffffu, sp[w], off);
  }
  float q[NW], kk[NW], vv[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) q[w] = kk[w] = vv[w] = 0.f;
#pragma unroll 4
  for (int k = 0; k < r; ++k) {
    const float wq = iv ? s_qT[k * r + i] : 0.f, wk = iv ? s_kvT[k * 2 * r + i] : 0.f, wv2 = iv ? s_kvT[k * 2 * r + r + i] : 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float spk = __shfl_sync(0xffffffffu, sp[w], k);
      const float dnk = __shfl_sync(0xffffffffu, dn[w], k);
      q[w] = fmaf(spk, wq, q[w]);
      kk[w] = fmaf(dnk, wk, kk[w]);
      vv[w] = fmaf(dnk, wv2, vv[w]);
    }
  }
  float o[NW];
  {
    float qi[NW], mxl[NW], den[NW], num[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) { qi[w] = q[w] * rs; mxl[w] = -INFINITY; den[w] = 0.f; num[w] = 0.f; }
#pragma unroll 4
    for (int j = 0; j < r; ++j)
#pragma unroll
      for (int w = 0; w < NW; ++w) mxl[w] = fmaxf(mxl[w], qi[w] * __shfl_sync(0xffffffffu, kk[w], j));
#pragma unroll 4
    for (int j = 0; j < r; ++j)
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const float wgt = expf(qi[w] * __shfl_sync(0xffffffffu, kk[w], j) - mxl[w]);
        den[w] += wgt;
        num[w] = fmaf(wgt, __shfl_sync(0xffffffffu, vv[w], j), num[w]);
      }
#pragma unroll
    for (int w = 0; w < NW; ++w) o[w] = iv ? num[w] / den[w] : 0.f;
  }
  float u[NW];
  {
    const float b = iv ? s_p2b[i] : 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) u[w] = b;
#pragma unroll 4
    for (int k = 0; k < r; ++k) {
      const float wp2 = iv ? s_p2T[k * r + i] : 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) u[w] = fmaf(__shfl_sync(0xffffffffu, o[w], k), wp2, u[w]);
    }
  }
  // g = upT^T u (:152): lane -> channels lane + 32 m
  for (int c0 = 0; c0 < C; c0 += 128) {
    float a[NW][4];
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int m = 0; m < 4; ++m) a[w][m] = 0.f;
#pragma unroll 4
    for (int k = 0; k < r; ++k) {
      float up[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int c = c0 + 32 * m + lane;
        up[m] = c < C ? s_upT[k * C + c] : 0.f;
      }
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const float uk = __shfl_sync(0xffffffffu, u[w], k);
#pragma unroll
        for (int m = 0; m < 4; ++m) a[w][m] = fmaf(uk, up[m], a[w][m]);
      }
    }
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int win = w0 + NW * warp + w;
      if (win >= p.B_) continue;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int c = c0 + 32 * m + lane;
        if (c < C) p.gate[(long long)win * C + c] = a[w][m];
      }
    }
  }
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

extern "C" int mphsir_local_gate2_fwd(const mphsir_local_gate_params* p, void* stream) {
  MPHSIR_REQUIRE(p && p->core_mean && p->gate, "local_gate2: null operand");
  MPHSIR_REQUIRE(p->promptT && p->promptb && p->downT && p->downb && p->param && p->qT && p->kvT && p->p2T && p->p2b && p->upT, "local_gate2: null weight");
  MPHSIR_REQUIRE(p->B_ > 0 && p->C > 0 && p->r >= 4 && p->r <= RMAX && p->r % 4 == 0, "local_gate2: bad shape B_=%d C=%d r=%d (r multiple of 4, <= %d)", p->B_, p->C, p->r, RMAX);
  MPHSIR_REQUIRE(p->C % 32 == 0, "local_gate2: C=%d must be a multiple of 32", p->C);
  const uintptr_t al = reinterpret_cast<uintptr_t>(p->core_mean) | reinterpret_cast<uintptr_t>(p->param) | reinterpret_cast<uintptr_t>(p->qT) |
                       reinterpret_cast<uintptr_t>(p->kvT) | reinterpret_cast<uintptr_t>(p->p2T) | reinterpret_cast<uintptr_t>(p->p2b) |
                       reinterpret_cast<uintptr_t>(p->upT);
  MPHSIR_REQUIRE((al & 15) == 0, "local_gate2: window means and weights must be 16-byte aligned");
  int rp = 4;
  while (rp < p->r) rp <<= 1;
  const int r = p->r;
  // windows per warp: as few as keeps the grid within ~two CTAs per SM
  // measured (tools/gate_bench.py, cold weights): 4096 windows 25.6 us at 4 per warp / 28.7 at 2; 1024 windows 16.4 us at 1 / 24.6 at 4
  const int nw = p->B_ > 8 * 296 ? 4 : 1;
  const int win = 8 * nw;
  const size_t smem = sizeof(float) * ((size_t)PLEN * r + 4 * r * r + (size_t)r * p->C + ((r + 3) & ~3) + 8 * PLEN +
                                       (size_t)win * LG2_LDL + (size_t)win * p->C + (size_t)LG2_NS * LG2_KC * (PLEN + r)) + 16 + 16 * LG2_NS;
  MPHSIR_REQUIRE(smem <= 224 * 1024, "local_gate2: C=%d r=%d needs %zu B of shared memory", p->C, r, smem);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(local_gate2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(local_gate2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(local_gate2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) {
      set_error("local_gate2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  const dim3 grid((p->B_ + win - 1) / win);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (nw == 4) local_gate2_kernel<4><<<grid, 256, smem, st>>>(*p, rp);
  else if (nw == 2) local_gate2_kernel<2><<<grid, 256, smem, st>>>(*p, rp);
  else local_gate2_kernel<1><<<grid, 256, smem, st>>>(*p, rp);
  return check_launch("local_gate2");
}
