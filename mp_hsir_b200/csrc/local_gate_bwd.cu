// Backward of the local spectral branch's per-window gate (PG_Spectral_Attention.forward, net/MP_HSIR.py:135-152).
//
// Forward, per window with m = mean over its 64 tokens of the spatial-attention output:
//   w = softmax(W_prompt m) [128];  sp = w P [r];  q = W_q sp;  low = W_down m;  [k; v] = W_kv low;
//   A = softmax_j(q_i k_j r^-0.5);  o = A v;  u = W_proj o + b;  g = W_up u  [C]
// The caller supplies LL = [W_prompt m | W_down m] (one GEMM) and dG = dL/dg; one warp per window recomputes the
// chain and writes a *record* of the small per-window vectors.  All weight gradients are then token-contractions
// of record columns (mphsir_wgrad / mphsir_colsum), and dL/dm = [dlogits | dlow] [W_prompt ; W_down] is one GEMM.
//
// record layout (floats, r % 4 == 0):  dlogits[128] | dlow[r] | w[128] | dsp[r] | dq[r] | sp[r] | dkv[2r] | low[r] |
//                                      du[r] | o[r] | u[r]            (ld = 256 + 10 r)
#include "common.cuh"

namespace mphsir {
namespace lgb {

constexpr int RMAX = 32;
constexpr int WARPS = 4;

__global__ void __launch_bounds__(WARPS * 32) local_gate_bwd_kernel(const float* __restrict__ LL, int ldl,
                                                                    const float* __restrict__ dG, const float* __restrict__ param,
                                                                    const float* __restrict__ qW, const float* __restrict__ kvW,
                                                                    const float* __restrict__ p2W, const float* __restrict__ p2b,
                                                                    const float* __restrict__ upW, float* __restrict__ rec,
                                                                    int ldr, int B_, int C, int r) {
  __shared__ float s_vec[WARPS][8][RMAX];           // sp, q, k, v, o, du, do, dq
  __shared__ float s_A[WARPS][RMAX][RMAX + 1];
  __shared__ float s_dS[WARPS][RMAX][RMAX + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int win = blockIdx.x * WARPS + warp;
  if (win >= B_) return;
  float* sp = s_vec[warp][0];
  float* q = s_vec[warp][1];
  float* k = s_vec[warp][2];
  float* v = s_vec[warp][3];
  float* o = s_vec[warp][4];
  float* du = s_vec[warp][5];
  float* dO = s_vec[warp][6];
  float* dq = s_vec[warp][7];
  const float* ll = LL + (long long)win * ldl;
  float* out = rec + (long long)win * ldr;
  const int OFF_DLOW = 128, OFF_W = 128 + r, OFF_DSP = 256 + r, OFF_DQ = 256 + 2 * r, OFF_SP = 256 + 3 * r,
            OFF_DKV = 256 + 4 * r, OFF_LOW = 256 + 6 * r, OFF_DU = 256 + 7 * r, OFF_O = 256 + 8 * r, OFF_U = 256 + 9 * r;
  const float sc = rsqrtf((float)r);

  // 1. prompt weights w = softmax(logits)
  float lg[4], w[4];
  float mx = -3.0e38f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    lg[e] = __ldg(ll + lane + 32 * e);
    mx = fmaxf(mx, lg[e]);
  }
  mx = warp_max(mx);
  float se = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    w[e] = expf(lg[e] - mx);
    se += w[e];
  }
  se = 1.0f / warp_sum(se);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    w[e] *= se;
    out[OFF_W + lane + 32 * e] = w[e];
  }
  // 2. sp = w P
  for (int j = 0; j < r; ++j) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) s = fmaf(w[e], __ldg(param + (lane + 32 * e) * r + j), s);
    s = warp_sum(s);
    if (lane == 0) sp[j] = s;
  }
  __syncwarp();
  // 3. q = W_q sp ; 4. [k; v] = W_kv low
  const float low_l = lane < r ? __ldg(ll + 128 + lane) : 0.f;
  if (lane < r) {
    float s = 0.f;
    for (int j = 0; j < r; ++j) s = fmaf(__ldg(qW + lane * r + j), sp[j], s);
    q[lane] = s;
    out[OFF_SP + lane] = sp[lane];
    out[OFF_LOW + lane] = low_l;
  }
  for (int t = lane; t < 2 * r; t += 32) {
    float s = 0.f;
    for (int j = 0; j < r; ++j) s = fmaf(__ldg(kvW + t * r + j), __ldg(ll + 128 + j), s);
    if (t < r) k[t] = s; else v[t - r] = s;
  }
  __syncwarp();
  // 5. A = softmax_j(q_i k_j sc), o = A v   (lane i owns row i)
  if (lane < r) {
    float m2 = -3.0e38f;
    for (int j = 0; j < r; ++j) m2 = fmaxf(m2, q[lane] * k[j] * sc);
    float sum = 0.f;
    for (int j = 0; j < r; ++j) {
      const float e = expf(q[lane] * k[j] * sc - m2);
      s_A[warp][lane][j] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
    float acc = 0.f;
    for (int j = 0; j < r; ++j) {
      const float a = s_A[warp][lane][j] * inv;
      s_A[warp][lane][j] = a;
      acc = fmaf(a, v[j], acc);
    }
    o[lane] = acc;
    out[OFF_O + lane] = acc;
  }
  __syncwarp();
  // 6. u = W_proj o + b ; 7. du = W_up^T dG
  if (lane < r) {
    float s = __ldg(p2b + lane);
    for (int j = 0; j < r; ++j) s = fmaf(__ldg(p2W + lane * r + j), o[j], s);
    out[OFF_U + lane] = s;
    float d = 0.f;
    const float* dg = dG + (long long)win * C;
    for (int c = 0; c < C; ++c) d = fmaf(__ldg(dg + c), __ldg(upW + c * r + lane), d);
    du[lane] = d;
    out[OFF_DU + lane] = d;
  }
  __syncwarp();
  // 8. do = W_proj^T du
  if (lane < r) {
    float s = 0.f;
    for (int i = 0; i < r; ++i) s = fmaf(du[i], __ldg(p2W + i * r + lane), s);
    dO[lane] = s;
  }
  __syncwarp();
  // 9. dA_ij = do_i v_j ; dS = A o (dA - rowsum(A o dA)) ; dq_i = sc sum_j dS_ij k_j
  if (lane < r) {
    float dot = 0.f;
    for (int j = 0; j < r; ++j) dot = fmaf(s_A[warp][lane][j], dO[lane] * v[j], dot);
    float acc = 0.f;
    for (int j = 0; j < r; ++j) {
      const float ds = s_A[warp][lane][j] * (dO[lane] * v[j] - dot);
      s_dS[warp][lane][j] = ds;
      acc = fmaf(ds, k[j], acc);
    }
    dq[lane] = acc * sc;
    out[OFF_DQ + lane] = acc * sc;
  }
  __syncwarp();
  // dk_j = sc sum_i dS_ij q_i ; dv_j = sum_i A_ij do_i   (lane j owns column j); reuse k/v slots afterwards via registers
  float dk = 0.f, dv = 0.f;
  if (lane < r) {
    for (int i = 0; i < r; ++i) {
      dk = fmaf(s_dS[warp][i][lane], q[i], dk);
      dv = fmaf(s_A[warp][i][lane], dO[i], dv);
    }
    dk *= sc;
    out[OFF_DKV + lane] = dk;
    out[OFF_DKV + r + lane] = dv;
  }
  __syncwarp();
  // 10. dlow = W_kv^T [dk; dv]   (stage dkv in the k / v slots)
  if (lane < r) {
    k[lane] = dk;
    v[lane] = dv;
  }
  __syncwarp();
  if (lane < r) {
    float s = 0.f;
    for (int t = 0; t < r; ++t) s = fmaf(k[t], __ldg(kvW + t * r + lane), s);
    for (int t = 0; t < r; ++t) s = fmaf(v[t], __ldg(kvW + (r + t) * r + lane), s);
    out[OFF_DLOW + lane] = s;
    // 11. dsp = W_q^T dq
    float d = 0.f;
    for (int i = 0; i < r; ++i) d = fmaf(dq[i], __ldg(qW + i * r + lane), d);
    sp[lane] = d;  // sp is dead: reuse for dsp
    out[OFF_DSP + lane] = d;
  }
  __syncwarp();
  // 12. dw_p = sum_j dsp_j P[p,j] ; dlogit = w (dw - sum w dw)
  float dw[4], dot = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float s = 0.f;
    for (int j = 0; j < r; ++j) s = fmaf(sp[j], __ldg(param + (lane + 32 * e) * r + j), s);
    dw[e] = s;
    dot = fmaf(w[e], s, dot);
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int e = 0; e < 4; ++e) out[lane + 32 * e] = w[e] * (dw[e] - dot);
}

}  // namespace lgb
}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_local_gate_bwd_record_ld(int r) { return 256 + 10 * r; }

extern "C" int mphsir_local_gate_bwd(const float* LL, int ldl, const float* dG, const mphsir_local_gate_bwd_weights* w,
                                     float* record, int ldr, int B_, int C, int r, void* stream) {
  MPHSIR_REQUIRE(LL && dG && w && record, "local_gate_bwd: null operand");
  MPHSIR_REQUIRE(w->param && w->q && w->kv && w->proj && w->proj_bias && w->up, "local_gate_bwd: null weight");
  MPHSIR_REQUIRE(B_ > 0 && C > 0 && r > 0 && r <= lgb::RMAX && r % 4 == 0, "local_gate_bwd: rank must be a multiple of 4, <= 32");
  MPHSIR_REQUIRE(ldl >= 128 + r && ldr >= 256 + 10 * r, "local_gate_bwd: leading dimensions too small");
  const int blocks = (B_ + lgb::WARPS - 1) / lgb::WARPS;
  lgb::local_gate_bwd_kernel<<<blocks, lgb::WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      LL, ldl, dG, w->param, w->q, w->kv, w->proj, w->proj_bias, w->up, record, ldr, B_, C, r);
  return check_launch("local_gate_bwd");
}
