// Backward of the global spectral ("transposed") attention statistics (Spectral_Attention.forward,
// net/MP_HSIR.py:101-113; Attention :412-426; CrossAttention :239-248) — the small-matrix part.
//
// Forward per (sample, head):  Gh = (q^T k) / (|q_i| |k_j|),  A = softmax_j(Gh * T),  out = project_out(A v).
// With dU = dL/d(out) [tokens, C] the caller first forms  P_b = dU_b^T v_b  [C, C]  (mphsir_wgrad, per-sample mode).
// This kernel, one CTA per (sample, head), then computes
//   dA_ij   = sum_o Wout[o, hc+i] P_b[o, hc+j]
//   dS      = A o (dA - rowsum(A o dA)),   dT_h += sum dS o Gh,   dGh = dS * T_h
//   dWout[o, hc+i] += sum_j P_b[o, hc+j] A_ij
// and emits the per-sample matrix Wb [2C, 2C] ("in x out") with which ONE GEMM over the saved [q | k] gives [dq | dk]:
//   dq[n,i] = sum_j k[n,j] dGh_ij/(|q_i||k_j|) - q[n,i] s_i/|q_i|^2,     s_i  = sum_j dGh_ij Gh_ij
//   dk[n,j] = sum_i q[n,i] dGh_ij/(|q_i||k_j|) - k[n,j] s'_j/|k_j|^2,    s'_j = sum_i dGh_ij Gh_ij
// (the L2 normalisation over all tokens, F.normalize :104-105, differentiated in closed form).
// dv = dU (Wout blockdiag(A)) is the forward's folded matrix used transposed — no work here.
#include "common.cuh"

namespace mphsir {
namespace spb {

// TM = c/16: every thread owns a TM x TM micro-tile of the c x c head matrices (c in {32,48,64,96})
template <int TM>
__global__ void __launch_bounds__(256) spectral_bwd_kernel(const float* __restrict__ P, long long p_batch_stride,
                                                           const float* __restrict__ Wout, const float* __restrict__ gsum,
                                                           const float* __restrict__ temperature, float* __restrict__ Wb,
                                                           long long wb_batch_stride, int ldwb, float* __restrict__ dWout,
                                                           float* __restrict__ dTemp, int heads) {
  constexpr int c = 16 * TM;
  constexpr int LDm = c + 1;
  constexpr int KC = 32;  // rows of Wout / P staged per step
  extern __shared__ float sm[];
  const int C = heads * c;
  float* A = sm;                 // [c][LDm]
  float* D = A + c * LDm;        // dA -> dGh
  float* Gh = D + c * LDm;       // normalised Gram
  float* Ws = Gh + c * LDm;      // [KC][LDm] staged Wout[o, hc + i]
  float* Ps = Ws + KC * LDm;     // [KC][LDm] staged P_b[o, hc + j]
  float* nq = Ps + KC * LDm;     // [c] 1/|q_i|
  float* nk = nq + c;            // [c] 1/|k_j|
  float* sq = nk + c;            // [c] s_i
  float* sk = sq + c;            // [c] s'_j
  __shared__ float red[8];
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ti = tid >> 4, tj = tid & 15;
  const float* gs = gsum + (long long)blockIdx.x * (c * c + 2 * c);
  const float* Pb = P + (long long)b * p_batch_stride;
  const float T = __ldg(temperature + h);
  const int hc = h * c;

  for (int i = tid; i < c; i += 256) {
    nq[i] = 1.0f / fmaxf(sqrtf(__ldg(gs + c * c + i)), 1e-12f);
    nk[i] = 1.0f / fmaxf(sqrtf(__ldg(gs + c * c + c + i)), 1e-12f);
  }
  __syncthreads();
  for (int e = tid; e < c * c; e += 256) {
    const int i = e / c, j = e - i * c;
    Gh[i * LDm + j] = __ldg(gs + e) * nq[i] * nk[j];
  }
  // ---- dA_ij = sum_o Wout[o, hc+i] * P_b[o, hc+j]: staged mini-GEMM, micro-tile rows ti*TM.., cols tj*TM.. --------
  {
    float acc[TM][TM];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
      for (int e = 0; e < TM; ++e) acc[a][e] = 0.f;
    for (int o0 = 0; o0 < C; o0 += KC) {
      __syncthreads();
      for (int e = tid; e < KC * c; e += 256) {
        const int r = e / c, col = e - r * c;
        const bool ok = o0 + r < C;
        Ws[r * LDm + col] = ok ? __ldg(Wout + (long long)(o0 + r) * C + hc + col) : 0.f;
        Ps[r * LDm + col] = ok ? __ldg(Pb + (long long)(o0 + r) * C + hc + col) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int r = 0; r < KC; ++r) {
        float wv[TM], pv[TM];
#pragma unroll
        for (int a = 0; a < TM; ++a) {
          wv[a] = Ws[r * LDm + ti * TM + a];
          pv[a] = Ps[r * LDm + tj * TM + a];
        }
#pragma unroll
        for (int a = 0; a < TM; ++a)
#pragma unroll
          for (int e = 0; e < TM; ++e) acc[a][e] = fmaf(wv[a], pv[e], acc[a][e]);
      }
    }
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
      for (int e = 0; e < TM; ++e) D[(ti * TM + a) * LDm + tj * TM + e] = acc[a][e];
  }
  __syncthreads();
  // A = softmax_j(Gh*T); dS = A o (dA - rowsum(A o dA)); dT partial   (warp per row)
  float dT = 0.f;
  for (int i = warp; i < c; i += 8) {
    float mx = -3.0e38f;
    for (int j = lane; j < c; j += 32) mx = fmaxf(mx, Gh[i * LDm + j] * T);
    mx = warp_max(mx);
    float se = 0.f;
    for (int j = lane; j < c; j += 32) {
      const float e = expf(Gh[i * LDm + j] * T - mx);
      A[i * LDm + j] = e;
      se += e;
    }
    const float inv = 1.0f / warp_sum(se);
    float dot = 0.f;
    for (int j = lane; j < c; j += 32) {
      const float a = A[i * LDm + j] * inv;
      A[i * LDm + j] = a;
      dot = fmaf(a, D[i * LDm + j], dot);
    }
    dot = warp_sum(dot);
    float si = 0.f;
    for (int j = lane; j < c; j += 32) {
      const float ds = A[i * LDm + j] * (D[i * LDm + j] - dot);
      dT = fmaf(ds, Gh[i * LDm + j], dT);
      const float dgh = ds * T;
      si = fmaf(dgh, Gh[i * LDm + j], si);
      D[i * LDm + j] = dgh;
    }
    si = warp_sum(si);
    if (lane == 0) sq[i] = si;
  }
  dT = warp_sum(dT);
  if (lane == 0) red[warp] = dT;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += red[wv];
    atomicAdd(dTemp + h, t);
  }
  for (int j = tid; j < c; j += 256) {
    float s = 0.f;
    for (int i = 0; i < c; ++i) s = fmaf(D[i * LDm + j], Gh[i * LDm + j], s);
    sk[j] = s;
  }
  __syncthreads();
  // ---- Wb rows of this head: q-in rows hc+i, k-in rows C+hc+j; full 2C columns each (zeros outside the head) ----
  float* wb = Wb + (long long)b * wb_batch_stride;
  const int twoC = 2 * C;
  for (int e = tid; e < c * twoC; e += 256) {
    const int i = e / twoC, col = e - i * twoC;  // q-in row hc+i
    float val = 0.f;
    if (col < C) {  // dq-out column
      if (col == hc + i) val = -sq[i] * nq[i] * nq[i];
    } else {        // dk-out column C + hc + j
      const int j = col - C - hc;
      if (j >= 0 && j < c) val = D[i * LDm + j] * nq[i] * nk[j];
    }
    wb[(long long)(hc + i) * ldwb + col] = val;
  }
  for (int e = tid; e < c * twoC; e += 256) {
    const int j = e / twoC, col = e - j * twoC;  // k-in row C+hc+j
    float val = 0.f;
    if (col < C) {  // dq-out column hc + i
      const int i = col - hc;
      if (i >= 0 && i < c) val = D[i * LDm + j] * nq[i] * nk[j];
    } else {
      if (col == C + hc + j) val = -sk[j] * nk[j] * nk[j];
    }
    wb[(long long)(C + hc + j) * ldwb + col] = val;
  }
  // ---- dWout[o, hc+i] += sum_j P_b[o, hc+j] A_ij : KC rows of P staged per step, thread -> (row r, c/8 values of i) ----
  constexpr int IPT = c / 8;
  const int r = tid >> 3, ib = (tid & 7) * IPT;
  for (int o0 = 0; o0 < C; o0 += KC) {
    __syncthreads();
    for (int e = tid; e < KC * c; e += 256) {
      const int rr = e / c, col = e - rr * c;
      Ps[rr * LDm + col] = (o0 + rr < C) ? __ldg(Pb + (long long)(o0 + rr) * C + hc + col) : 0.f;
    }
    __syncthreads();
    float acc[IPT];
#pragma unroll
    for (int e = 0; e < IPT; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int j = 0; j < c; ++j) {
      const float pv = Ps[r * LDm + j];
#pragma unroll
      for (int e = 0; e < IPT; ++e) acc[e] = fmaf(pv, A[(ib + e) * LDm + j], acc[e]);
    }
    if (o0 + r < C) {
#pragma unroll
      for (int e = 0; e < IPT; ++e) atomicAdd(dWout + (long long)(o0 + r) * C + hc + ib + e, acc[e]);
    }
  }
}

template <int TM>
static int launch(const float* P, long long p_batch_stride, const float* Wout, const float* gsum, const float* temperature,
                  float* Wb, int ldwb, long long wb_batch_stride, float* dWout, float* dTemp, int B, int heads, cudaStream_t st) {
  constexpr int c = 16 * TM;
  const size_t smem = sizeof(float) * ((3 * c + 2 * 32) * (c + 1) + 4 * c);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(spectral_bwd_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("spectral_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  spectral_bwd_kernel<TM><<<B * heads, 256, smem, st>>>(P, p_batch_stride, Wout, gsum, temperature, Wb, wb_batch_stride, ldwb,
                                                        dWout, dTemp, heads);
  return check_launch("spectral_bwd");
}

}  // namespace spb
}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_spectral_bwd(const float* P, long long p_batch_stride, const float* Wout, const float* gsum,
                                   const float* temperature, float* Wb, int ldwb, long long wb_batch_stride, float* dWout,
                                   float* dTemperature, int B, int heads, int c, void* stream) {
  MPHSIR_REQUIRE(P && Wout && gsum && temperature && Wb && dWout && dTemperature, "spectral_bwd: null operand");
  MPHSIR_REQUIRE(B > 0 && heads > 0 && c > 0 && ldwb >= 2 * heads * c, "spectral_bwd: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define SPB_CASE(TM) \
  case 16 * TM: return spb::launch<TM>(P, p_batch_stride, Wout, gsum, temperature, Wb, ldwb, wb_batch_stride, dWout, dTemperature, B, heads, st)
  switch (c) {
    SPB_CASE(2);
    SPB_CASE(3);
    SPB_CASE(4);
    SPB_CASE(6);
    default:
      set_error("spectral_bwd: channels per head %d not supported (32, 48, 64, 96)", c);
      return MPHSIR_ERR_INVALID;
  }
#undef SPB_CASE
}
