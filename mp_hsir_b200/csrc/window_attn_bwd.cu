// Backward of the window attention core (Spatial_Attention.forward, net/MP_HSIR.py:195-215, with the roll /
// window_partition / window_reverse of PGSSTB.forward :671-696 folded into the addressing, as in the forward).
//
// One CTA walks a strided set of (shifted) windows of one head.  Per window it recomputes
//   S = (q*scale) k^T + rel_pos_bias + shift_mask,  P = softmax(S)
// from the saved qkv, then
//   dV = P^T dO,   dP = dO V^T,   dS = P o (dP - rowsum(P o dP)),   dQ = scale * dS K,   dK = dS^T (scale*Q)
// and accumulates dS over its windows in registers: the relative-position-bias gradient leaves as one
// [64,64] partial per CTA (reduced by mphsir_colsum, scattered to the [225,heads] table by
// mphsir_rpb_table_bwd).  fp32 FFMA throughout.
#include "common.cuh"

namespace mphsir {
namespace wab {

constexpr int WT = 64;
constexpr int LD = WT + 4;

template <int HD>
__global__ void __launch_bounds__(256) window_attn_bwd_kernel(const float* __restrict__ qkv, long long ldqkv,
                                                              const float* __restrict__ bias, const float* __restrict__ dO,
                                                              long long ldo, float* __restrict__ dqkv, long long lddq,
                                                              float* __restrict__ dbias_partial, int n_win, int H, int W,
                                                              int C, int heads, int shift) {
  constexpr int DV = HD / 16;
  constexpr int HD4 = HD / 4;
  extern __shared__ __align__(16) float smem[];
  // phase 1: Qt | Kt | Vt | dOt  ([HD][LD] each);  phase 2 reuses the region as Q | K | dO ([64][HD] each)
  float* Qt = smem;
  float* Kt = Qt + HD * LD;
  float* Vt = Kt + HD * LD;
  float* Gt = Vt + HD * LD;
  float* Ps = Gt + HD * LD;    // [64][LD]
  float* Ds = Ps + WT * LD;    // [64][LD]  S, then dP, then dS
  int* rows = reinterpret_cast<int*>(Ds + WT * LD);
  int* label = rows + WT;
  float* Qr = smem;
  float* Kr = Qr + WT * HD;
  float* Gr = Kr + WT * HD;

  const int tid = threadIdx.x;
  const int head = blockIdx.y;
  const int tx = tid & 15, ty = tid >> 4;
  const int warp = tid >> 5, lane = tid & 31;
  const int nWx = W >> 3, nW = (H >> 3) * nWx;
  const float scale = rsqrtf((float)HD);

  float bacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) bacc[i][j] = 0.f;

  for (int win = blockIdx.x; win < n_win; win += gridDim.x) {
    const int b = win / nW;
    const int wrem = win - b * nW;
    const int wi = wrem / nWx, wj = wrem - wi * nWx;
    __syncthreads();  // previous window fully consumed
    if (tid < WT) {
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
    // ---- gather q (scaled), k, v, dO transposed ---------------------------------------------
    for (int idx = tid; idx < WT * HD4; idx += 256) {
      const int t = idx / HD4, dq = idx - t * HD4;
      const float* base = qkv + (long long)rows[t] * ldqkv + head * HD + dq * 4;
      const float4 q = ldg4(base), k = ldg4(base + C), v = ldg4(base + 2 * C);
      const float4 g = ldg4(dO + (long long)rows[t] * ldo + head * HD + dq * 4);
      const int d = dq * 4;
      Qt[(d + 0) * LD + t] = q.x * scale;
      Qt[(d + 1) * LD + t] = q.y * scale;
      Qt[(d + 2) * LD + t] = q.z * scale;
      Qt[(d + 3) * LD + t] = q.w * scale;
      Kt[(d + 0) * LD + t] = k.x;
      Kt[(d + 1) * LD + t] = k.y;
      Kt[(d + 2) * LD + t] = k.z;
      Kt[(d + 3) * LD + t] = k.w;
      Vt[(d + 0) * LD + t] = v.x;
      Vt[(d + 1) * LD + t] = v.y;
      Vt[(d + 2) * LD + t] = v.z;
      Vt[(d + 3) * LD + t] = v.w;
      Gt[(d + 0) * LD + t] = g.x;
      Gt[(d + 1) * LD + t] = g.y;
      Gt[(d + 2) * LD + t] = g.z;
      Gt[(d + 3) * LD + t] = g.w;
    }
    __syncthreads();
    // ---- S and dP: thread -> 4x4 block --------------------------------------------------------
    float dp[4][4];
    {
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
#pragma unroll 4
      for (int d = 0; d < HD; ++d) {
        const float4 a = *reinterpret_cast<const float4*>(&Qt[d * LD + ty * 4]);
        const float4 k4 = *reinterpret_cast<const float4*>(&Kt[d * LD + tx * 4]);
        const float4 g4 = *reinterpret_cast<const float4*>(&Gt[d * LD + ty * 4]);
        const float4 v4 = *reinterpret_cast<const float4*>(&Vt[d * LD + tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, kv[4] = {k4.x, k4.y, k4.z, k4.w};
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s[i][j] = fmaf(av[i], kv[j], s[i][j]);
            dp[i][j] = fmaf(gv[i], vv[j], dp[i][j]);
          }
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
        for (int j = 0; j < 4; ++j) ov[j] = s[i][j] + bv[j] + (label[p] != label[tx * 4 + j] ? -100.0f : 0.0f);
        *reinterpret_cast<float4*>(&Ds[p * LD + tx * 4]) = o;
      }
    }
    __syncthreads();
    // ---- P = softmax(S) rows -> Ps ----------------------------------------------------------------
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int p = warp * 8 + rr;
      const float v0 = Ds[p * LD + lane], v1 = Ds[p * LD + lane + 32];
      const float mx = warp_max(fmaxf(v0, v1));
      const float e0 = expf(v0 - mx), e1 = expf(v1 - mx);
      const float inv = 1.0f / warp_sum(e0 + e1);
      Ps[p * LD + lane] = e0 * inv;
      Ps[p * LD + lane + 32] = e1 * inv;
    }
    __syncthreads();
    // ---- dP -> Ds, then dS = P o (dP - rowsum(P o dP)) in place ------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(&Ds[(ty * 4 + i) * LD + tx * 4]) = make_float4(dp[i][0], dp[i][1], dp[i][2], dp[i][3]);
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int p = warp * 8 + rr;
      const float p0 = Ps[p * LD + lane], p1 = Ps[p * LD + lane + 32];
      const float d0 = Ds[p * LD + lane], d1 = Ds[p * LD + lane + 32];
      const float dot = warp_sum(fmaf(p0, d0, p1 * d1));
      Ds[p * LD + lane] = p0 * (d0 - dot);
      Ds[p * LD + lane + 32] = p1 * (d1 - dot);
    }
    __syncthreads();
    // accumulate the bias gradient (this thread's 4x4 block, same block for every window)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(&Ds[(ty * 4 + i) * LD + tx * 4]);
      bacc[i][0] += v.x;
      bacc[i][1] += v.y;
      bacc[i][2] += v.z;
      bacc[i][3] += v.w;
    }
    // ---- phase 2 operands, row-major: Q (scaled), K, dO (the transposed copies are dead) ------------
    for (int idx = tid; idx < WT * HD4; idx += 256) {
      const int t = idx / HD4, dq = idx - t * HD4;
      const float* base = qkv + (long long)rows[t] * ldqkv + head * HD + dq * 4;
      const float4 q = ldg4(base), k = ldg4(base + C);
      const float4 g = ldg4(dO + (long long)rows[t] * ldo + head * HD + dq * 4);
      *reinterpret_cast<float4*>(&Qr[t * HD + dq * 4]) = make_float4(q.x * scale, q.y * scale, q.z * scale, q.w * scale);
      *reinterpret_cast<float4*>(&Kr[t * HD + dq * 4]) = k;
      *reinterpret_cast<float4*>(&Gr[t * HD + dq * 4]) = g;
    }
    __syncthreads();
    // ---- dV = P^T dO, dK = dS^T Qs, dQ = scale * dS K : thread -> 4 tokens x DV channels -----------
    float aq[4][DV], ak[4][DV], av[4][DV];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < DV; ++e) aq[i][e] = ak[i][e] = av[i][e] = 0.f;
#pragma unroll 4
    for (int j = 0; j < WT; ++j) {
      const float4 pc = *reinterpret_cast<const float4*>(&Ps[j * LD + ty * 4]);   // P[j][4ty..]   (contract over rows)
      const float4 sc = *reinterpret_cast<const float4*>(&Ds[j * LD + ty * 4]);   // dS[j][4ty..]
      const float pcv[4] = {pc.x, pc.y, pc.z, pc.w}, scv[4] = {sc.x, sc.y, sc.z, sc.w};
      float srv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) srv[i] = Ds[(ty * 4 + i) * LD + j];             // dS[4ty+i][j] (contract over cols)
      float gj[DV], qj[DV], kj[DV];
#pragma unroll
      for (int e = 0; e < DV; ++e) {
        gj[e] = Gr[j * HD + tx * DV + e];
        qj[e] = Qr[j * HD + tx * DV + e];
        kj[e] = Kr[j * HD + tx * DV + e];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < DV; ++e) {
          av[i][e] = fmaf(pcv[i], gj[e], av[i][e]);
          ak[i][e] = fmaf(scv[i], qj[e], ak[i][e]);
          aq[i][e] = fmaf(srv[i], kj[e], aq[i][e]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* dst = dqkv + (long long)rows[ty * 4 + i] * lddq + head * HD + tx * DV;
#pragma unroll
      for (int e = 0; e < DV; ++e) {
        dst[e] = aq[i][e] * scale;
        dst[C + e] = ak[i][e];
        dst[2 * C + e] = av[i][e];
      }
    }
  }
  // ---- bias-gradient partial of this CTA -------------------------------------------------------
  float* dst = dbias_partial + ((long long)blockIdx.x * heads + head) * WT * WT;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(dst + (ty * 4 + i) * WT + tx * 4) = make_float4(bacc[i][0], bacc[i][1], bacc[i][2], bacc[i][3]);
}

// dTable[e, h] = sum over (p,q) with relative_position_index[p,q] == e of dBias[h,p,q]   (net/MP_HSIR.py:172-181,200-202)
__global__ void rpb_table_bwd_kernel(const float* __restrict__ dbias, float* __restrict__ dtable, int heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 225 * heads) return;
  const int e = idx / heads, h = idx - e * heads;
  const int dyy = e / 15 - 7, dxx = e - (e / 15) * 15 - 7;
  float s = 0.f;
  for (int p = 0; p < 64; ++p) {
    const int yq = (p >> 3) - dyy, xq = (p & 7) - dxx;
    if ((unsigned)yq < 8u && (unsigned)xq < 8u) s += __ldg(dbias + ((long long)h * 64 + p) * 64 + yq * 8 + xq);
  }
  dtable[idx] += s;
}

template <int HD>
static size_t smem_bytes() {
  return sizeof(float) * (4 * HD * LD + 2 * WT * LD) + sizeof(int) * 2 * WT;
}

template <int HD>
static int launch(const float* qkv, int ldqkv, const float* bias, const float* dO, int ldo, float* dqkv, int lddq,
                  float* partial, int groups, int B, int H, int W, int C, int heads, int shift, cudaStream_t st) {
  const size_t smem = smem_bytes<HD>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("window_attn_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(groups, heads);
  window_attn_bwd_kernel<HD><<<grid, 256, smem, st>>>(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, partial, B * (H / 8) * (W / 8), H, W,
                                                      C, heads, shift);
  return check_launch("window_attn_bwd");
}

}  // namespace wab
namespace wabm {
int launch_window_attn_bwd_mma(const float* qkv, int ldqkv, const float* bias, const float* dO, int ldo, float* dqkv, int lddq,
                               float* partial, int groups, int B, int H, int W, int C, int heads, int shift, int parts,
                               cudaStream_t st);
}
}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_window_attn_bwd_groups(int B, int H, int W, int heads) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_win = B * (H / 8) * (W / 8);
  int g = (4 * sms) / heads;  // one wave of the 128-thread tensor-core kernel at 4 CTAs per SM (2 waves of the FFMA kernel)
  if (g > n_win) g = n_win;
  return g < 1 ? 1 : g;
}

extern "C" int mphsir_window_attn_bwd(const float* qkv, int ldqkv, const float* bias, const float* dO, int ldo, float* dqkv,
                                      int lddq, float* dbias_partial, int groups, int B, int H, int W, int C, int heads,
                                      int shift, int precision, void* stream) {
  MPHSIR_REQUIRE(qkv && bias && dO && dqkv && dbias_partial, "window_attn_bwd: null operand");
  MPHSIR_REQUIRE(B > 0 && H % 8 == 0 && W % 8 == 0 && heads > 0 && C % heads == 0 && groups > 0, "window_attn_bwd: bad shape");
  MPHSIR_REQUIRE(shift == 0 || shift == 4, "window_attn_bwd: shift must be 0 or 4");
  MPHSIR_REQUIRE(ldqkv % 4 == 0 && ldo % 4 == 0 && ldqkv >= 3 * C && lddq >= 3 * C && ldo >= C, "window_attn_bwd: bad leading dimension");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == MPHSIR_PREC_BF16X3 || precision == MPHSIR_PREC_BF16)
    return wabm::launch_window_attn_bwd_mma(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, dbias_partial, groups, B, H, W, C, heads, shift,
                                            precision == MPHSIR_PREC_BF16X3 ? 2 : 1, st);
  const int hd = C / heads;
  switch (hd) {
    case 32: return wab::launch<32>(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, dbias_partial, groups, B, H, W, C, heads, shift, st);
    case 48: return wab::launch<48>(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, dbias_partial, groups, B, H, W, C, heads, shift, st);
    case 64: return wab::launch<64>(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, dbias_partial, groups, B, H, W, C, heads, shift, st);
    case 96: return wab::launch<96>(qkv, ldqkv, bias, dO, ldo, dqkv, lddq, dbias_partial, groups, B, H, W, C, heads, shift, st);
    default:
      set_error("window_attn_bwd: head_dim %d not supported (32, 48, 64, 96)", hd);
      return MPHSIR_ERR_INVALID;
  }
}

extern "C" int mphsir_rpb_table_bwd(const float* dbias, float* dtable, int heads, void* stream) {
  MPHSIR_REQUIRE(dbias && dtable && heads > 0, "rpb_table_bwd: bad arguments");
  wab::rpb_table_bwd_kernel<<<(225 * heads + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dbias, dtable, heads);
  return check_launch("rpb_table_bwd");
}
