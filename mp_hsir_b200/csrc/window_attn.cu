// Window attention core (fp32 SIMT): one CTA per (window, head).
//
// Folds torch.roll / window_partition / window_reverse / un-roll of PGSSTB.forward
// (net/MP_HSIR.py:671-678, 689-696) into the gather/scatter addressing, evaluates the Swin mask
// of calculate_mask (:639-660) in closed form and adds the pre-gathered relative-position bias
// (:200-203).  Also emits the per-window token mean of the result for the local spectral branch.
#include "common.cuh"

namespace mphsir {

constexpr int WT = 64;         // tokens per window
constexpr int QK_LD = WT + 4;  // padded row of the transposed q/k tiles and of S / P^T

template <int HD>
__global__ void __launch_bounds__(256) window_attn_kernel(const float* __restrict__ qkv, long long ldqkv,
                                                          const float* __restrict__ bias,
                                                          float* __restrict__ out, long long ldo,
                                                          float* __restrict__ win_mean, int H, int W, int C,
                                                          int heads, int shift) {
  constexpr int DV = HD / 16;  // output columns per thread in the PV product
  constexpr int HD4 = HD / 4;
  extern __shared__ __align__(16) float smem[];
  // layout: Qt[HD][QK_LD] | Kt[HD][QK_LD] | V[64][HD] | S[64][QK_LD] | rows[64] | label[64] | red[16][HD]
  // P^T[64][QK_LD] overlays Qt|Kt once S has been computed (needs 2*HD >= 64; HD>=32).
  float* Qt = smem;
  float* Kt = Qt + HD * QK_LD;
  float* Vs = Kt + HD * QK_LD;
  float* Ss = Vs + WT * HD;
  int* rows = reinterpret_cast<int*>(Ss + WT * QK_LD);
  int* label = rows + WT;
  float* red = reinterpret_cast<float*>(label + WT);
  float* Pt = Qt;

  const int tid = threadIdx.x;
  const int head = blockIdx.y;
  const int win = blockIdx.x;  // b*nW + wi*(W/8) + wj
  const int nWx = W >> 3, nW = (H >> 3) * nWx;
  const int b = win / nW;
  const int wrem = win - b * nW;
  const int wi = wrem / nWx, wj = wrem - wi * nWx;

  if (tid < WT) {
    const int r = tid >> 3, c = tid & 7;
    const int ys = wi * 8 + r, xs = wj * 8 + c;  // shifted coordinates
    int y = ys + shift, x = xs + shift;          // source pixel: roll(-s) => shifted[ys] = x[(ys+s) mod H]
    if (y >= H) y -= H;
    if (x >= W) x -= W;
    rows[tid] = (b * H + y) * W + x;
    const int rh = (ys >= H - 8) + (ys >= H - 4);
    const int rw = (xs >= W - 8) + (xs >= W - 4);
    label[tid] = shift ? 3 * rh + rw : 0;
  }
  __syncthreads();

  // ---- gather q,k (transposed) and v ---------------------------------------------------
  const float scale = rsqrtf((float)HD);
  for (int idx = tid; idx < WT * HD4; idx += 256) {
    const int t = idx / HD4, dq = idx - t * HD4;
    const float* base = qkv + (long long)rows[t] * ldqkv + head * HD + dq * 4;
    const float4 q = ldg4(base);
    const float4 k = ldg4(base + C);
    const float4 v = ldg4(base + 2 * C);
    const int d = dq * 4;
    Qt[(d + 0) * QK_LD + t] = q.x * scale;
    Qt[(d + 1) * QK_LD + t] = q.y * scale;
    Qt[(d + 2) * QK_LD + t] = q.z * scale;
    Qt[(d + 3) * QK_LD + t] = q.w * scale;
    Kt[(d + 0) * QK_LD + t] = k.x;
    Kt[(d + 1) * QK_LD + t] = k.y;
    Kt[(d + 2) * QK_LD + t] = k.z;
    Kt[(d + 3) * QK_LD + t] = k.w;
    *reinterpret_cast<float4*>(&Vs[t * HD + d]) = v;
  }
  __syncthreads();

  // ---- S = q k^T + bias + mask : thread -> 4x4 block ------------------------------------
  const int tx = tid & 15, ty = tid >> 4;
  {
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      const float4 a = *reinterpret_cast<const float4*>(&Qt[d * QK_LD + ty * 4]);
      const float4 k4 = *reinterpret_cast<const float4*>(&Kt[d * QK_LD + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float kv[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], kv[j], s[i][j]);
    }
    const float* bh = bias + (long long)head * WT * WT;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = ty * 4 + i;
      const float4 bb = ldg4(bh + p * WT + tx * 4);
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
      float4 o;
      float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = tx * 4 + j;
        ov[j] = s[i][j] + bv[j] + (label[p] != label[q] ? -100.0f : 0.0f);
      }
      *reinterpret_cast<float4*>(&Ss[p * QK_LD + tx * 4]) = o;
    }
  }
  __syncthreads();

  // ---- row softmax; write P transposed over the dead Qt|Kt region ------------------------
  {
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int p = warp * 8 + rr;
      const float v0 = Ss[p * QK_LD + lane], v1 = Ss[p * QK_LD + lane + 32];
      const float mx = warp_max(fmaxf(v0, v1));
      const float e0 = expf(v0 - mx), e1 = expf(v1 - mx);
      const float inv = 1.0f / warp_sum(e0 + e1);
      Pt[lane * QK_LD + p] = e0 * inv;
      Pt[(lane + 32) * QK_LD + p] = e1 * inv;
    }
  }
  __syncthreads();

  // ---- O = P V : thread -> 4 rows x DV cols ------------------------------------------------
  float o[4][DV];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < DV; ++j) o[i][j] = 0.f;
#pragma unroll 8
  for (int j = 0; j < WT; ++j) {
    const float4 pp = *reinterpret_cast<const float4*>(&Pt[j * QK_LD + ty * 4]);
    const float pv[4] = {pp.x, pp.y, pp.z, pp.w};
    float vv[DV];
#pragma unroll
    for (int e = 0; e < DV; ++e) vv[e] = Vs[j * HD + tx * DV + e];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < DV; ++e) o[i][e] = fmaf(pv[i], vv[e], o[i][e]);
  }
  float colsum[DV];
#pragma unroll
  for (int e = 0; e < DV; ++e) colsum[e] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* dst = out + (long long)rows[ty * 4 + i] * ldo + head * HD + tx * DV;
#pragma unroll
    for (int e = 0; e < DV; ++e) {
      dst[e] = o[i][e];
      colsum[e] += o[i][e];
    }
  }
#pragma unroll
  for (int e = 0; e < DV; ++e) red[ty * HD + tx * DV + e] = colsum[e];
  __syncthreads();
  if (tid < HD) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) s += red[r * HD + tid];
    win_mean[(long long)win * C + head * HD + tid] = s * (1.0f / WT);
  }
}

template <int HD>
static size_t smem_bytes() {
  return sizeof(float) * (2 * HD * QK_LD + WT * HD + WT * QK_LD + 16 * HD) + sizeof(int) * 2 * WT;
}

template <int HD>
static int launch_wa(const float* qkv, int ldqkv, const float* bias, float* out, int ldo, float* win_mean, int B,
                     int H, int W, int C, int heads, int shift, cudaStream_t st) {
  const size_t smem = smem_bytes<HD>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("window_attn: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(B * (H / 8) * (W / 8), heads);
  window_attn_kernel<HD><<<grid, 256, smem, st>>>(qkv, ldqkv, bias, out, ldo, win_mean, H, W, C, heads, shift);
  return check_launch("window_attn");
}

namespace wa {
int launch_window_attn_mma(const float* qkv, int ldqkv, const float* bias, float* out, int ldo, float* win_mean, int B,
                           int H, int W, int C, int heads, int shift, int parts, int mask_H, int mask_y0, cudaStream_t st);
}

}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_window_attn_band_fwd(const float* qkv, int ldqkv, const float* bias, float* out, int ldo,
                                           float* win_mean, int B, int H, int W, int C, int heads, int shift,
                                           int precision, int mask_H, int mask_y0, void* stream) {
  MPHSIR_REQUIRE(qkv && bias && out && win_mean, "window_attn: null operand");
  MPHSIR_REQUIRE(B > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "window_attn: H=%d W=%d must be multiples of 8", H, W);
  MPHSIR_REQUIRE(heads > 0 && C % heads == 0, "window_attn: C=%d not divisible by heads=%d", C, heads);
  MPHSIR_REQUIRE(shift == 0 || shift == 4, "window_attn: shift must be 0 or 4");
  MPHSIR_REQUIRE(ldqkv >= 3 * C && ldqkv % 4 == 0 && ldo >= C, "window_attn: bad leading dimensions");
  MPHSIR_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "window_attn: qkv/bias must be 16-byte aligned");
  MPHSIR_REQUIRE(mask_H >= 8 && mask_y0 >= 0 && mask_y0 < mask_H, "window_attn: bad mask geometry (mask_H=%d mask_y0=%d)", mask_H, mask_y0);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MPHSIR_REQUIRE(precision >= MPHSIR_PREC_FP32_SIMT && precision <= MPHSIR_PREC_BF16, "window_attn: unknown precision %d", precision);
  if (precision != MPHSIR_PREC_FP32_SIMT)
    return wa::launch_window_attn_mma(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift,
                                      precision == MPHSIR_PREC_BF16X3 ? 2 : 1, mask_H, mask_y0, st);
  MPHSIR_REQUIRE(mask_H == H && mask_y0 == 0, "window_attn: row bands of a sharded scene run on the tensor-core precisions only");
  switch (C / heads) {
    case 32: return launch_wa<32>(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, st);
    case 48: return launch_wa<48>(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, st);
    case 64: return launch_wa<64>(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, st);
    case 96: return launch_wa<96>(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, st);
    default:
      MPHSIR_REQUIRE(false, "window_attn: head_dim %d not in {32,48,64,96}", C / heads);
  }
}

extern "C" int mphsir_window_attn_fwd(const float* qkv, int ldqkv, const float* bias, float* out, int ldo,
                                      float* win_mean, int B, int H, int W, int C, int heads, int shift,
                                      int precision, void* stream) {
  return mphsir_window_attn_band_fwd(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, precision, H, 0, stream);
}
