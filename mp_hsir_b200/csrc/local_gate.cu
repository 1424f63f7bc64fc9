// Local spectral branch: the low-rank spectral-prompt attention of PG_Spectral_Attention
// (net/MP_HSIR.py:132-155) collapsed to one fused kernel that turns the per-window token mean
// of the attention core into a per-window channel gate g[B_,C].
//
//   m   = projT^T core_mean + projb            (mean over tokens commutes with Spatial_Attention.proj)
//   pw  = softmax(promptT^T m)   [128]         (:136)
//   dn  = downT^T m              [r]           (:137)
//   sp  = pw @ param             [r]           (:139-140)
//   q   = qT^T sp ; [k;v] = kvT^T dn           (:142-144)
//   A   = softmax_j(q_i k_j r^-0.5) ; o_i = sum_j A_ij v_j     (:146-149, outer-product attention)
//   g   = upT^T (p2T^T o + p2b)                (:151-152)
//
// One CTA (128 threads) handles WPB windows; every weight is read "in x out" so that thread o
// reads W[k][o] coalesced.  < 20 kFLOP per window: latency-, not throughput-, bound.
#include "common.cuh"

namespace mphsir {

constexpr int LG_THREADS = 128;
constexpr int PLEN = 128;
constexpr int RMAX = 32;

__global__ void __launch_bounds__(LG_THREADS) local_gate_kernel(const mphsir_local_gate_params p) {
  extern __shared__ float sm[];
  const int C = p.C, r = p.r;
  float* cm = sm;            // [C] core mean
  float* m = cm + C;         // [C]
  float* pw = m + C;         // [128]
  float* dn = pw + PLEN;     // [RMAX]
  float* sp = dn + RMAX;
  float* q = sp + RMAX;
  float* kv = q + RMAX;      // [2*RMAX]
  float* o = kv + 2 * RMAX;
  float* u = o + RMAX;
  float* red = u + RMAX;     // [8]
  const int tid = threadIdx.x;
  const int win = blockIdx.x;
  if (win >= p.B_) return;

  for (int c = tid; c < C; c += LG_THREADS) cm[c] = __ldg(p.core_mean + (long long)win * C + c);
  __syncthreads();
  for (int c = tid; c < C; c += LG_THREADS) {
    float a = __ldg(p.projb + c);
    for (int k = 0; k < C; ++k) a = fmaf(cm[k], __ldg(p.projT + (long long)k * C + c), a);
    m[c] = a;
  }
  __syncthreads();
  // prompt logits (thread == prompt index) and the low-rank projection
  float logit = 0.f;
  for (int k = 0; k < C; ++k) logit = fmaf(m[k], __ldg(p.promptT + k * PLEN + tid), logit);
  if (tid < r) {
    float a = 0.f;
    for (int k = 0; k < C; ++k) a = fmaf(m[k], __ldg(p.downT + k * r + tid), a);
    dn[tid] = a;
  }
  // softmax over the 128 logits (4 warps)
  float mx = warp_max(logit);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  const float e = expf(logit - mx);
  float s = warp_sum(e);
  if ((tid & 31) == 0) red[4 + (tid >> 5)] = s;
  __syncthreads();
  s = (red[4] + red[5]) + (red[6] + red[7]);
  pw[tid] = e / s;
  __syncthreads();
  if (tid < r) {
    float a = 0.f;
    for (int k = 0; k < PLEN; ++k) a = fmaf(pw[k], __ldg(p.param + k * r + tid), a);
    sp[tid] = a;
  }
  __syncthreads();
  if (tid < r) {
    float a = 0.f;
    for (int k = 0; k < r; ++k) a = fmaf(sp[k], __ldg(p.qT + k * r + tid), a);
    q[tid] = a;
  }
  if (tid >= 32 && tid < 32 + 2 * r) {
    const int j = tid - 32;
    float a = 0.f;
    for (int k = 0; k < r; ++k) a = fmaf(dn[k], __ldg(p.kvT + k * 2 * r + j), a);
    kv[j] = a;
  }
  __syncthreads();
  if (tid < r) {
    const float sc = rsqrtf((float)r);
    const float qi = q[tid] * sc;
    float mxl = -INFINITY;
    for (int j = 0; j < r; ++j) mxl = fmaxf(mxl, qi * kv[j]);
    float den = 0.f, num = 0.f;
    for (int j = 0; j < r; ++j) {
      const float w = expf(qi * kv[j] - mxl);
      den += w;
      num = fmaf(w, kv[r + j], num);
    }
    o[tid] = num / den;
  }
  __syncthreads();
  if (tid < r) {
    float a = __ldg(p.p2b + tid);
    for (int k = 0; k < r; ++k) a = fmaf(o[k], __ldg(p.p2T + k * r + tid), a);
    u[tid] = a;
  }
  __syncthreads();
  for (int c = tid; c < C; c += LG_THREADS) {
    float a = 0.f;
    for (int k = 0; k < r; ++k) a = fmaf(u[k], __ldg(p.upT + k * C + c), a);
    p.gate[(long long)win * C + c] = a;
  }
}

}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_local_gate_fwd(const mphsir_local_gate_params* p, void* stream) {
  MPHSIR_REQUIRE(p && p->core_mean && p->gate, "local_gate: null operand");
  MPHSIR_REQUIRE(p->projT && p->projb && p->promptT && p->downT && p->param && p->qT && p->kvT && p->p2T && p->p2b && p->upT, "local_gate: null weight");
  MPHSIR_REQUIRE(p->B_ > 0 && p->C > 0 && p->r > 0 && p->r <= RMAX, "local_gate: bad shape B_=%d C=%d r=%d (r<=%d)", p->B_, p->C, p->r, RMAX);
  const size_t smem = sizeof(float) * (2 * p->C + PLEN + 7 * RMAX + 8);
  local_gate_kernel<<<p->B_, LG_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(*p);
  return check_launch("local_gate");
}
