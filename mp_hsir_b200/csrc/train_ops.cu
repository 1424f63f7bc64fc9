// Backward / training glue kernels of the MP-HSIR hot path (fp32, HBM-bound): LayerNorm forward-with-statistics
// and backward, the gate derivatives of GatedMlp / GDFN, DropPath row scaling, per-window reductions of the
// local spectral branch, depthwise-conv weight gradients, pixel (un)shuffle on token-major data, the
// scatter form of the bilinear resize, the TVSP query gradient, clamp+L1 loss and AdamW.
//
// Everything the reference gets from autograd over net/MP_HSIR.py is restated here by hand; each kernel names
// the forward lines it differentiates.  128-bit accesses along the channel axis throughout.
#include "common.cuh"

namespace mphsir {
namespace trn {

static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

__device__ __forceinline__ float gelu_grad(float x) {
  // d/dx [x Phi(x)] = Phi(x) + x phi(x)
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  return fmaf(x, pdf, cdf);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the channel axis (nn.LayerNorm / WithBias_LayerNorm, net/MP_HSIR.py:618-619, :354-357)
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV_LIMIT = 8;  // float4 chunks per lane -> C <= 1024

// LN_MAXV = ceil(C / 128): 1..3 for every width of the network (64..384); keeps xhat / g*gamma of a row in registers
template <int LN_MAXV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ X, long long ldx,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ Y, long long ldy, float* __restrict__ stats,
                                                            long long M, int C) {
  const int lane = threadIdx.x & 31;
  const int c4n = C >> 2;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long m = warp0; m < M; m += nwarps) {
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < c4n) {
        v[i] = ldg4(X + m * ldx + 4 * c4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mu = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < c4n) {
        const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < c4n) {
        const float4 g = ldg4(gamma + 4 * c4), b = ldg4(beta + 4 * c4);
        float4 o;
        o.x = fmaf((v[i].x - mu) * rstd, g.x, b.x);
        o.y = fmaf((v[i].y - mu) * rstd, g.y, b.y);
        o.z = fmaf((v[i].z - mu) * rstd, g.z, b.z);
        o.w = fmaf((v[i].w - mu) * rstd, g.w, b.w);
        *reinterpret_cast<float4*>(Y + m * ldy + 4 * c4) = o;
      }
    }
    if (lane == 0 && stats != nullptr) {
      stats[2 * m] = mu;
      stats[2 * m + 1] = rstd;
    }
  }
}

// dX = add + rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat));  dgamma += sum g*xhat;  dbeta += sum g
template <int LN_MAXV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ X, long long ldx,
                                                            const float* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ G, long long ldg,
                                                            const float* __restrict__ add, long long lda,
                                                            float* __restrict__ dX, long long lddx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, long long M, int C) {
  extern __shared__ float red[];  // [2][C]
  const int lane = threadIdx.x & 31;
  const int c4n = C >> 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float4 ag[LN_MAXV], ab[LN_MAXV];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long m = warp0; m < M; m += nwarps) {
    const float mu = __ldg(stats + 2 * m), rstd = __ldg(stats + 2 * m + 1);
    float4 xh[LN_MAXV], gw[LN_MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < c4n) {
        const float4 x = ldg4(X + m * ldx + 4 * c4);
        const float4 g = ldg4(G + m * ldg + 4 * c4);
        const float4 w = ldg4(gamma + 4 * c4);
        xh[i] = make_float4((x.x - mu) * rstd, (x.y - mu) * rstd, (x.z - mu) * rstd, (x.w - mu) * rstd);
        gw[i] = make_float4(g.x * w.x, g.y * w.y, g.z * w.z, g.w * w.w);
        s1 += (gw[i].x + gw[i].y) + (gw[i].z + gw[i].w);
        s2 += (gw[i].x * xh[i].x + gw[i].y * xh[i].y) + (gw[i].z * xh[i].z + gw[i].w * xh[i].w);
        ag[i].x = fmaf(g.x, xh[i].x, ag[i].x);
        ag[i].y = fmaf(g.y, xh[i].y, ag[i].y);
        ag[i].z = fmaf(g.z, xh[i].z, ag[i].z);
        ag[i].w = fmaf(g.w, xh[i].w, ag[i].w);
        ab[i].x += g.x;
        ab[i].y += g.y;
        ab[i].z += g.z;
        ab[i].w += g.w;
      }
    }
    const float m1 = warp_sum(s1) / (float)C, m2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < c4n) {
        float4 o;
        o.x = rstd * (gw[i].x - m1 - xh[i].x * m2);
        o.y = rstd * (gw[i].y - m1 - xh[i].y * m2);
        o.z = rstd * (gw[i].z - m1 - xh[i].z * m2);
        o.w = rstd * (gw[i].w - m1 - xh[i].w * m2);
        if (add != nullptr) {
          const float4 a = ldg4(add + m * lda + 4 * c4);
          o.x += a.x;
          o.y += a.y;
          o.z += a.z;
          o.w += a.w;
        }
        *reinterpret_cast<float4*>(dX + m * lddx + 4 * c4) = o;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c4 = lane + 32 * i;
    if (c4 < c4n) {
      atomicAdd(&red[4 * c4 + 0], ag[i].x);
      atomicAdd(&red[4 * c4 + 1], ag[i].y);
      atomicAdd(&red[4 * c4 + 2], ag[i].z);
      atomicAdd(&red[4 * c4 + 3], ag[i].w);
      atomicAdd(&red[C + 4 * c4 + 0], ab[i].x);
      atomicAdd(&red[C + 4 * c4 + 1], ab[i].y);
      atomicAdd(&red[C + 4 * c4 + 2], ab[i].z);
      atomicAdd(&red[C + 4 * c4 + 3], ab[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
  }
}

// ------------------------------------------------------------------------------------------------
// GatedMlp gate (net/MP_HSIR.py:77-79) on the packed fc1 output h[:, (2j, 2j+1)] = (value_j, gate_j):
//   hidden_j = v * gelu(g);  given dHid:  dv = dHid * gelu(g),  dg = dHid * v * gelu'(g)
// dH overwrites H, the recomputed hidden overwrites dHid (both in place).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) glu_bwd_kernel(float* __restrict__ Hh, long long ldh, float* __restrict__ Dh,
                                                      long long ldd, long long M, int hid_pad) {
  const int j2n = hid_pad >> 1;
  const long long total = M * j2n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / j2n;
    const int j2 = (int)(idx - m * j2n);
    float4* hp = reinterpret_cast<float4*>(Hh + m * ldh + 4 * j2);
    float2* dp = reinterpret_cast<float2*>(Dh + m * ldd + 2 * j2);
    const float4 h = *hp;
    const float2 d = *dp;
    const float g0 = gelu_erf(h.y), g1 = gelu_erf(h.w);
    *hp = make_float4(d.x * g0, d.x * h.x * gelu_grad(h.y), d.y * g1, d.y * h.z * gelu_grad(h.w));
    *dp = make_float2(h.x * g0, h.z * g1);
  }
}

// GDFN gate (net/MP_HSIR.py:388-389, :262-263): halves a = T[:, 0:hp), b = T[:, hp:2hp):  y = gelu(a) * b
__global__ void __launch_bounds__(256) gdfn_gate_fwd_kernel(const float* __restrict__ T, long long ldt,
                                                            float* __restrict__ Y, long long ldy, long long M, int hp) {
  const int q4n = hp >> 2;
  const long long total = M * q4n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / q4n;
    const int c = (int)(idx - m * q4n) * 4;
    const float4 a = ldg4(T + m * ldt + c), b = ldg4(T + m * ldt + hp + c);
    *reinterpret_cast<float4*>(Y + m * ldy + c) =
        make_float4(gelu_erf(a.x) * b.x, gelu_erf(a.y) * b.y, gelu_erf(a.z) * b.z, gelu_erf(a.w) * b.w);
  }
}

// dT (may alias T):  da = dY * b * gelu'(a),  db = dY * gelu(a)
__global__ void __launch_bounds__(256) gdfn_gate_bwd_kernel(const float* T, long long ldt, const float* __restrict__ dY,
                                                            long long ldy, float* dT, long long lddt, long long M, int hp) {
  const int q4n = hp >> 2;
  const long long total = M * q4n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / q4n;
    const int c = (int)(idx - m * q4n) * 4;
    const float4 a = *reinterpret_cast<const float4*>(T + m * ldt + c);
    const float4 b = *reinterpret_cast<const float4*>(T + m * ldt + hp + c);
    const float4 d = ldg4(dY + m * ldy + c);
    *reinterpret_cast<float4*>(dT + m * lddt + c) =
        make_float4(d.x * b.x * gelu_grad(a.x), d.y * b.y * gelu_grad(a.y), d.z * b.z * gelu_grad(a.z), d.w * b.w * gelu_grad(a.w));
    *reinterpret_cast<float4*>(dT + m * lddt + hp + c) =
        make_float4(d.x * gelu_erf(a.x), d.y * gelu_erf(a.y), d.z * gelu_erf(a.z), d.w * gelu_erf(a.w));
  }
}

// ------------------------------------------------------------------------------------------------
// Y = alpha * s[row / rows_per_batch] * X + beta * Y      (DropPath scaling, gradient accumulation, copies;
// x_row_mod > 0 broadcasts a shared operand over the batch)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ Y,
                                                    long long ldy, long long M, int C, float alpha, float beta,
                                                    const float* __restrict__ row_scale, int rows_per_batch, int x_row_mod) {
  const int c4n = C >> 2;
  const long long total = M * c4n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / c4n;
    const int c = (int)(idx - m * c4n) * 4;
    const long long mx = x_row_mod > 0 ? m % x_row_mod : m;
    float a = alpha;
    if (row_scale != nullptr) a *= __ldg(row_scale + m / rows_per_batch);
    const float4 x = ldg4(X + mx * ldx + c);
    float4* yp = reinterpret_cast<float4*>(Y + m * ldy + c);
    float4 o = make_float4(a * x.x, a * x.y, a * x.z, a * x.w);
    if (beta != 0.f) {
      const float4 y = *yp;
      o.x = fmaf(beta, y.x, o.x);
      o.y = fmaf(beta, y.y, o.y);
      o.z = fmaf(beta, y.z, o.z);
      o.w = fmaf(beta, y.w, o.w);
    }
    *yp = o;
  }
}

// out[n, c] = sum_b X[b*rows + n, c]      (gradients of operands shared by the whole batch: TVSP visual prompt)
__global__ void __launch_bounds__(256) batch_sum_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ Y,
                                                        long long ldy, int B, long long rows, int C) {
  const int c4n = C >> 2;
  const long long total = rows * c4n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / c4n;
    const int c = (int)(idx - n * c4n) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4 x = ldg4(X + ((long long)b * rows + n) * ldx + c);
      s.x += x.x;
      s.y += x.y;
      s.z += x.z;
      s.w += x.w;
    }
    *reinterpret_cast<float4*>(Y + n * ldy + c) = s;
  }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[map(c)] += sum_m X[m, c].  map: see wgrad (MPHSIR_MAP_*)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int map_index(int o, int mode, int a, int b) {
  // returns the destination row for packed column o, or -1 when o is padding
  if (mode == MPHSIR_MAP_IDENTITY) return o < a ? o : -1;
  if (mode == MPHSIR_MAP_INTERLEAVE) {  // packed (2j, 2j+1) = (first-half_j, second-half_j), a = valid per half
    const int j = o >> 1;
    return j < a ? (o & 1) * a + j : -1;
  }
  // MPHSIR_MAP_HALVES: halves at [0, a) and [b, b + a)
  if (o < b) return o < a ? o : -1;
  return (o - b) < a ? a + (o - b) : -1;
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ out,
                                                     long long M, int C, int rows_per_cta, int mode, int ma, int mb) {
  // block = 32 column quads (128 columns) x 8 row lanes; grid.x = row chunks, grid.y = 128-column groups
  __shared__ float4 red[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + cx * 4;
  const long long m0 = (long long)blockIdx.x * rows_per_cta;
  const long long m1 = min(M, m0 + rows_per_cta);
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  if (c < C) {
    long long m = m0 + ry;
    for (; m + 8 < m1; m += 16) {
      const float4 a = ldg4(X + m * ldx + c), b = ldg4(X + (m + 8) * ldx + c);
      s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
      s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
    }
    if (m < m1) {
      const float4 a = ldg4(X + m * ldx + c);
      s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
    }
  }
  red[ry][cx] = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
  __syncthreads();
  if (ry == 0 && c < C) {
    float4 t = red[0][cx];
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      t.x += red[r][cx].x; t.y += red[r][cx].y; t.z += red[r][cx].z; t.w += red[r][cx].w;
    }
    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int dst = (c + e < C) ? map_index(c + e, mode, ma, mb) : -1;
      if (dst >= 0) atomicAdd(out + dst, tv[e]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Per-window helpers of the local spectral branch (PG_Spectral_Attention, net/MP_HSIR.py:132-155; the
// windows are the *shifted* 8x8 windows of PGSSTB.forward :671-678, tokens stay in image order).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long window_token(int win, int t, int H, int W, int shift) {
  const int nWx = W >> 3, nW = (H >> 3) * nWx;
  const int b = win / nW;
  const int wrem = win - b * nW;
  const int wi = wrem / nWx, wj = wrem - wi * nWx;
  int y = wi * 8 + (t >> 3) + shift, x = wj * 8 + (t & 7) + shift;
  if (y >= H) y -= H;
  if (x >= W) x -= W;
  return ((long long)b * H + y) * W + x;
}

// out[win, c] = scale * sum_{t in win} A[t, c] * (Bm ? Bm[t, c] : 1)
// CTA = (window, 128-channel group): 32 channel quads x 8 token lanes, 8 independent loads per thread, smem reduce
__global__ void __launch_bounds__(256) window_reduce_kernel(const float* __restrict__ A, long long lda,
                                                            const float* __restrict__ Bm, long long ldb,
                                                            float* __restrict__ out, int n_win, int H, int W, int C,
                                                            int shift, float scale) {
  __shared__ float4 red[8][32];
  const int q = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int win = blockIdx.x;
  const int c = blockIdx.y * 128 + q * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    float4 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ldg4(A + window_token(win, tl * 8 + i, H, W, shift) * lda + c);
    if (Bm != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = ldg4(Bm + window_token(win, tl * 8 + i, H, W, shift) * ldb + c);
        a[i].x *= b.x;
        a[i].y *= b.y;
        a[i].z *= b.z;
        a[i].w *= b.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s.x += a[i].x;
      s.y += a[i].y;
      s.z += a[i].z;
      s.w += a[i].w;
    }
  }
  red[tl][q] = s;
  __syncthreads();
  if (tl == 0 && c < C) {
    float4 t = red[0][q];
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      t.x += red[r][q].x;
      t.y += red[r][q].y;
      t.z += red[r][q].z;
      t.w += red[r][q].w;
    }
    *reinterpret_cast<float4*>(out + (long long)win * C + c) = make_float4(t.x * scale, t.y * scale, t.z * scale, t.w * scale);
  }
}

// dSA[n, c] = dU[n, c] * gate[win(n), c] + dMean[win(n), c] / 64      (x1 = sa * g and m = mean_win(sa), :135,:153)
__global__ void __launch_bounds__(256) gate_apply_bwd_kernel(const float* __restrict__ dU, long long ldu,
                                                             const float* __restrict__ gate, const float* __restrict__ dMean,
                                                             float* __restrict__ dSA, long long lds, int B, int H, int W,
                                                             int C, int shift) {
  const int c4n = C >> 2;
  const long long total = (long long)B * H * W * c4n;
  const int nWx = W >> 3, nW = (H >> 3) * nWx;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / c4n;
    const int c = (int)(idx - n * c4n) * 4;
    const int x = (int)(n % W), y = (int)((n / W) % H), b = (int)(n / ((long long)W * H));
    int ys = y - shift, xs = x - shift;
    if (ys < 0) ys += H;
    if (xs < 0) xs += W;
    const long long win = (long long)b * nW + (ys >> 3) * nWx + (xs >> 3);
    const float4 d = ldg4(dU + n * ldu + c), g = ldg4(gate + win * C + c), m = ldg4(dMean + win * C + c);
    *reinterpret_cast<float4*>(dSA + n * lds + c) =
        make_float4(fmaf(d.x, g.x, m.x * 0.015625f), fmaf(d.y, g.y, m.y * 0.015625f), fmaf(d.z, g.z, m.z * 0.015625f),
                    fmaf(d.w, g.w, m.w * 0.015625f));
  }
}

// U[n, c] = X[n, c] + s_b * SA[n, c] * gate[win(n), c]: the shortcut plus the locally gated spatial branch (:153, :715-718)
// as a stand-alone pass, so that the spectral-apply GEMM of the training forward can use its TMA-drained residual epilogue
__global__ void __launch_bounds__(256) gate_apply_fwd_kernel(const float* __restrict__ X, long long ldx,
                                                             const float* __restrict__ SA, long long lds,
                                                             const float* __restrict__ gate, const float* __restrict__ row_scale,
                                                             float* __restrict__ U, long long ldu, int B, int H, int W, int C,
                                                             int shift) {
  const int c4n = C >> 2;
  const long long total = (long long)B * H * W * c4n;
  const int nWx = W >> 3, nW = (H >> 3) * nWx;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / c4n;
    const int c = (int)(idx - n * c4n) * 4;
    const int x = (int)(n % W), y = (int)((n / W) % H), b = (int)(n / ((long long)W * H));
    int ys = y - shift, xs = x - shift;
    if (ys < 0) ys += H;
    if (xs < 0) xs += W;
    const long long win = (long long)b * nW + (ys >> 3) * nWx + (xs >> 3);
    const float s = row_scale != nullptr ? __ldg(row_scale + b) : 1.0f;
    const float4 xv = ldg4(X + n * ldx + c), a = ldg4(SA + n * lds + c), g = ldg4(gate + win * C + c);
    *reinterpret_cast<float4*>(U + n * ldu + c) =
        make_float4(fmaf(s * a.x, g.x, xv.x), fmaf(s * a.y, g.y, xv.y), fmaf(s * a.z, g.z, xv.z), fmaf(s * a.w, g.w, xv.w));
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 weight gradient: dW[map(c), tap] += sum_{b,y,x} dY[b,y,x,c] * X[b,y+dy,x+dx,c]   (zero pad)
// Thread = (channel quad, image row): it walks the row with a sliding 3x3 window of X in registers (3 new float4 of X
// + 1 of dY per pixel for 36 FMAs).  block = 64 quads x 4 row lanes, each lane takes rows_per_lane rows;
// grid.x = row groups over B*H, grid.y = 256-channel groups.  Partial sums: shared-memory reduce, then atomics.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_kernel(const float* __restrict__ X, long long ldx,
                                                              const float* __restrict__ dY, long long ldy,
                                                              float* __restrict__ dW, int B, int H, int W, int C,
                                                              int rows_per_lane, int mode, int ma, int mb) {
  __shared__ float4 red[4][64];
  const int q = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const int c = blockIdx.y * 256 + q * 4;
  const long long total_rows = (long long)B * H;
  const long long r0 = ((long long)blockIdx.x * 4 + pl) * rows_per_lane;
  float4 acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    for (long long row = r0; row < min(total_rows, r0 + rows_per_lane); ++row) {
      const int y = (int)(row % H);
      const float* xr = X + row * W * ldx + c;   // (b, y, 0)
      const float* dr = dY + row * W * ldy + c;
      const bool up = y > 0, dn = y + 1 < H;
      // window columns: l = x-1, m = x, r = x+1 for rows y-1 (0), y (1), y+1 (2)
      float4 wl[3] = {zero, zero, zero}, wm[3], wr[3], d;
      wm[0] = up ? ldg4(xr - (long long)W * ldx) : zero;
      wm[1] = ldg4(xr);
      wm[2] = dn ? ldg4(xr + (long long)W * ldx) : zero;
      // column x+1 and dY[x] are always one iteration ahead of the FMAs that consume them
      {
        const float* p = xr + ldx;
        const bool in = 1 < W;
        wr[0] = (in && up) ? ldg4(p - (long long)W * ldx) : zero;
        wr[1] = in ? ldg4(p) : zero;
        wr[2] = (in && dn) ? ldg4(p + (long long)W * ldx) : zero;
        d = ldg4(dr);
      }
      for (int x = 0; x < W; ++x) {
        float4 nr[3], nd;
        {
          const bool in = x + 2 < W;
          const float* p = xr + (long long)(x + 2) * ldx;
          nr[0] = (in && up) ? ldg4(p - (long long)W * ldx) : zero;
          nr[1] = in ? ldg4(p) : zero;
          nr[2] = (in && dn) ? ldg4(p + (long long)W * ldx) : zero;
          nd = (x + 1 < W) ? ldg4(dr + (long long)(x + 1) * ldy) : zero;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          float4& a0 = acc[ky * 3 + 0];
          float4& a1 = acc[ky * 3 + 1];
          float4& a2 = acc[ky * 3 + 2];
          a0.x = fmaf(d.x, wl[ky].x, a0.x); a0.y = fmaf(d.y, wl[ky].y, a0.y); a0.z = fmaf(d.z, wl[ky].z, a0.z); a0.w = fmaf(d.w, wl[ky].w, a0.w);
          a1.x = fmaf(d.x, wm[ky].x, a1.x); a1.y = fmaf(d.y, wm[ky].y, a1.y); a1.z = fmaf(d.z, wm[ky].z, a1.z); a1.w = fmaf(d.w, wm[ky].w, a1.w);
          a2.x = fmaf(d.x, wr[ky].x, a2.x); a2.y = fmaf(d.y, wr[ky].y, a2.y); a2.z = fmaf(d.z, wr[ky].z, a2.z); a2.w = fmaf(d.w, wr[ky].w, a2.w);
          wl[ky] = wm[ky];
          wm[ky] = wr[ky];
          wr[ky] = nr[ky];
        }
        d = nd;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    red[pl][q] = acc[t];
    __syncthreads();
    if (pl == 0 && c < C) {
      float4 s = red[0][q];
#pragma unroll
      for (int r = 1; r < 4; ++r) {
        s.x += red[r][q].x;
        s.y += red[r][q].y;
        s.z += red[r][q].z;
        s.w += red[r][q].w;
      }
      // reference layout [C,1,3,3]: dW[map(c)*9 + tap]
      const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int dst = map_index(c + e, mode, ma, mb);
        if (dst >= 0) atomicAdd(dW + (long long)dst * 9 + t, sv[e]);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// PixelUnshuffle(2) / PixelShuffle(2) on token-major data (net/MP_HSIR.py:437,:447): the backward of a
// Downsample needs shuffle(dY), of an Upsample unshuffle(dY).
//   unshuffle: out[(b,y/2,x/2), c*4 + 2(y&1) + (x&1)] = in[(b,y,x), c]      in: [B*H*W, C]  (H,W = input size)
//   shuffle  : out[(b,2y+i,2x+j), c] = in[(b,y,x), c*4 + 2i + j]            in: [B*H*W, 4C]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixel_unshuffle_kernel(const float* __restrict__ in, long long ldi,
                                                              float* __restrict__ out, long long ldo, int B, int H,
                                                              int W, int C) {
  const long long total = (long long)B * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long n = idx / C;
    const int x = (int)(n % W), y = (int)((n / W) % H), b = (int)(n / ((long long)W * H));
    const long long no = ((long long)b * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
    out[no * ldo + c * 4 + 2 * (y & 1) + (x & 1)] = __ldg(in + n * ldi + c);
  }
}

__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const float* __restrict__ in, long long ldi,
                                                            float* __restrict__ out, long long ldo, int B, int H, int W,
                                                            int C) {
  // H, W: input (low-resolution) size; C: output channels
  const long long total = (long long)B * H * W * 4 * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long no = idx / C;  // output token
    const int W2 = 2 * W, H2 = 2 * H;
    const int xo = (int)(no % W2), yo = (int)((no / W2) % H2), b = (int)(no / ((long long)W2 * H2));
    const long long n = ((long long)b * H + (yo >> 1)) * W + (xo >> 1);
    out[no * ldo + c] = __ldg(in + n * ldi + c * 4 + 2 * (yo & 1) + (xo & 1));
  }
}

// tokens [B*HW, ld] -> NCHW [B, C, HW] and back is nchw_to_tokens (misc.cu)
__global__ void __launch_bounds__(256) tokens_to_nchw_kernel(const float* __restrict__ in, long long ld,
                                                             float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (c < C && p < HW) ? __ldg(in + ((long long)b * HW + p) * ld + c) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (c < C && p < HW) out[((long long)b * C + c) * HW + p] = tile[tx][r];
  }
}

// ------------------------------------------------------------------------------------------------
// bilinear resize backward (scatter; dX pre-zeroed): the transpose of bilinear_kernel (misc.cu)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const float* __restrict__ dY, long long ldy,
                                                           float* __restrict__ dX, long long ldx, int B, int h, int w,
                                                           int H, int W, int C) {
  const long long total = (long long)B * H * W * C;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long pix = idx / C;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
    const float fy = fmaxf(((float)y + 0.5f) * sy - 0.5f, 0.f);
    const float fx = fmaxf(((float)x + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float g = __ldg(dY + pix * ldy + c);
    float* base = dX + (long long)b * h * w * ldx + c;
    atomicAdd(base + ((long long)y0 * w + x0) * ldx, g * (1.f - ly) * (1.f - lx));
    atomicAdd(base + ((long long)y0 * w + x1) * ldx, g * (1.f - ly) * lx);
    atomicAdd(base + ((long long)y1 * w + x0) * ldx, g * ly * (1.f - lx));
    atomicAdd(base + ((long long)y1 * w + x1) * ldx, g * ly * lx);
  }
}

// ------------------------------------------------------------------------------------------------
// TVSP query gradient (net/MP_HSIR.py:575-577): Q[(b,i,j), d] = tp[b,d] * src[i,j], tp = (w @ learnable)/T
//   dLearnable[t, d] += w[b,t]/T * sum_{i,j} dQ[(b,i,j), d] * src[i,j]
// grid = (pixel chunks, B), block = D threads (D <= 1024)
// ------------------------------------------------------------------------------------------------
__global__ void tvsp_query_bwd_kernel(const float* __restrict__ dQ, long long ldq, const float* __restrict__ clip_b,
                                      const float* __restrict__ w, float* __restrict__ dLearn, int B, int T, int D,
                                      int ps, int pix_per_cta) {
  const int b = blockIdx.y, d = threadIdx.x;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(ps * ps, p0 + pix_per_cta);
  float acc = 0.f;
  for (int pix = p0; pix < p1; ++pix) {
    const int i = pix / ps, j = pix - i * ps;
    const int si = min((int)floorf((float)i * ((float)B / (float)ps)), B - 1);
    const int sj = min((int)floorf((float)j * (512.0f / (float)ps)), 511);
    acc = fmaf(__ldg(dQ + ((long long)b * ps * ps + pix) * ldq + d), __ldg(clip_b + si * 512 + sj), acc);
  }
  for (int t = 0; t < T; ++t) {
    const float wt = __ldg(w + b * T + t);
    if (wt != 0.f) atomicAdd(dLearn + (long long)t * D + d, wt * acc / (float)T);
  }
}

// ------------------------------------------------------------------------------------------------
// loss = mean |clamp(out,0,1) - clean| (train.py:58-61);  dOut = sign(.) * [0 <= out <= 1] * scale / numel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l1_clamp_loss_kernel(const float* __restrict__ out, const float* __restrict__ clean,
                                                            float* __restrict__ dOut, float* __restrict__ loss,
                                                            long long numel, float grad_scale) {
  __shared__ float red[8];
  float s = 0.f;
  const float inv = 1.0f / (float)numel;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    const float o = __ldg(out + i);
    const float r = fminf(fmaxf(o, 0.f), 1.f) - __ldg(clean + i);
    s += fabsf(r);
    const float sg = (r > 0.f) ? 1.f : (r < 0.f ? -1.f : 0.f);
    dOut[i] = (o >= 0.f && o <= 1.f) ? sg * inv * grad_scale : 0.f;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4);
    t += __shfl_xor_sync(0xffu, t, 2);
    t += __shfl_xor_sync(0xffu, t, 1);
    if (threadIdx.x == 0) atomicAdd(loss, t * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// AdamW over one flat fp32 parameter range (torch.optim.AdamW semantics: decoupled decay, bias correction,
// eps added after the sqrt of the corrected second moment).  grad_scale folds the 1/world_size of the DDP mean.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float beta1, float beta2,
                                                    float eps, float wd, float bc1, float bc2, float grad_scale,
                                                    const float* __restrict__ dyn) {
  if (dyn != nullptr) {  // CUDA-graph replays: step-dependent scalars live in device memory
    lr = __ldg(dyn);
    bc1 = __ldg(dyn + 1);
    bc2 = __ldg(dyn + 2);
  }
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = reinterpret_cast<float*>(&pp);
    const float* G = reinterpret_cast<const float*>(&gg);
    float* Mm = reinterpret_cast<float*>(&mm);
    float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = G[e] * grad_scale;
      P[e] *= (1.0f - lr * wd);
      Mm[e] = beta1 * Mm[e] + (1.0f - beta1) * gr;
      V[e] = beta2 * V[e] + (1.0f - beta2) * gr * gr;
      const float denom = sqrtf(V[e]) / sqrtf(bc2) + eps;
      P[e] -= (lr / bc1) * (Mm[e] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n not a multiple of 4)
  const long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float gr = g[i] * grad_scale;
    float pv = p[i] * (1.0f - lr * wd);
    const float mv = beta1 * m[i] + (1.0f - beta1) * gr;
    const float vv2 = beta2 * v[i] + (1.0f - beta2) * gr * gr;
    pv -= (lr / bc1) * (mv / (sqrtf(vv2) / sqrtf(bc2) + eps));
    p[i] = pv;
    m[i] = mv;
    v[i] = vv2;
  }
}

static int grid_for(long long items, int per_block = 256, int max_waves = 16) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace trn
}  // namespace mphsir

using namespace mphsir;
using namespace mphsir::trn;

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int mphsir_layernorm_fwd(const float* X, int ldx, const float* gamma, const float* beta, float* Y, int ldy,
                                    float* stats, long long M, int C, void* stream) {
  MPHSIR_REQUIRE(X && gamma && beta && Y && M > 0, "layernorm_fwd: null operand");
  MPHSIR_REQUIRE(C > 0 && C % 4 == 0 && C <= 128 * LN_MAXV_LIMIT && ldx % 4 == 0 && ldy % 4 == 0, "layernorm_fwd: C must be a multiple of 4, <= 1024");
  const int grid = grid_for(M, 8, 32);
  switch ((C + 127) / 128) {
    case 1: layernorm_fwd_kernel<1><<<grid, 256, 0, ST(stream)>>>(X, ldx, gamma, beta, Y, ldy, stats, M, C); break;
    case 2: layernorm_fwd_kernel<2><<<grid, 256, 0, ST(stream)>>>(X, ldx, gamma, beta, Y, ldy, stats, M, C); break;
    case 3: layernorm_fwd_kernel<3><<<grid, 256, 0, ST(stream)>>>(X, ldx, gamma, beta, Y, ldy, stats, M, C); break;
    default: layernorm_fwd_kernel<LN_MAXV_LIMIT><<<grid, 256, 0, ST(stream)>>>(X, ldx, gamma, beta, Y, ldy, stats, M, C); break;
  }
  return check_launch("layernorm_fwd");
}

extern "C" int mphsir_layernorm_bwd(const float* X, int ldx, const float* stats, const float* gamma, const float* G, int ldg,
                                    const float* add, int lda, float* dX, int lddx, float* dgamma, float* dbeta, long long M,
                                    int C, void* stream) {
  MPHSIR_REQUIRE(X && stats && gamma && G && dX && dgamma && dbeta && M > 0, "layernorm_bwd: null operand");
  MPHSIR_REQUIRE(C > 0 && C % 4 == 0 && C <= 128 * LN_MAXV_LIMIT && ldx % 4 == 0 && ldg % 4 == 0 && lddx % 4 == 0 && lda % 4 == 0,
                 "layernorm_bwd: C must be a multiple of 4, <= 1024");
  // >= 16 rows per warp keeps the dgamma / dbeta flush (shared + global atomics per CTA) small next to the row traffic
  const int grid = grid_for(M, 8 * 16, 8);
  const size_t smem = sizeof(float) * 2 * C;
#define LNB(V) layernorm_bwd_kernel<V><<<grid, 256, smem, ST(stream)>>>(X, ldx, stats, gamma, G, ldg, add, lda, dX, lddx, dgamma, dbeta, M, C)
  switch ((C + 127) / 128) {
    case 1: LNB(1); break;
    case 2: LNB(2); break;
    case 3: LNB(3); break;
    default: LNB(LN_MAXV_LIMIT); break;
  }
#undef LNB
  return check_launch("layernorm_bwd");
}

extern "C" int mphsir_glu_bwd(float* H, int ldh, float* dHid, int ldd, long long M, int hid_pad, void* stream) {
  MPHSIR_REQUIRE(H && dHid && M > 0 && hid_pad > 0 && hid_pad % 2 == 0 && ldh % 4 == 0 && ldd % 2 == 0, "glu_bwd: bad arguments");
  glu_bwd_kernel<<<grid_for(M * (hid_pad / 2)), 256, 0, ST(stream)>>>(H, ldh, dHid, ldd, M, hid_pad);
  return check_launch("glu_bwd");
}

extern "C" int mphsir_gdfn_gate_fwd(const float* T, int ldt, float* Y, int ldy, long long M, int hid_pad, void* stream) {
  MPHSIR_REQUIRE(T && Y && M > 0 && hid_pad > 0 && hid_pad % 4 == 0 && ldt % 4 == 0 && ldy % 4 == 0, "gdfn_gate_fwd: bad arguments");
  gdfn_gate_fwd_kernel<<<grid_for(M * (hid_pad / 4)), 256, 0, ST(stream)>>>(T, ldt, Y, ldy, M, hid_pad);
  return check_launch("gdfn_gate_fwd");
}

extern "C" int mphsir_gdfn_gate_bwd(const float* T, int ldt, const float* dY, int ldy, float* dT, int lddt, long long M,
                                    int hid_pad, void* stream) {
  MPHSIR_REQUIRE(T && dY && dT && M > 0 && hid_pad > 0 && hid_pad % 4 == 0 && ldt % 4 == 0 && ldy % 4 == 0 && lddt % 4 == 0,
                 "gdfn_gate_bwd: bad arguments");
  gdfn_gate_bwd_kernel<<<grid_for(M * (hid_pad / 4)), 256, 0, ST(stream)>>>(T, ldt, dY, ldy, dT, lddt, M, hid_pad);
  return check_launch("gdfn_gate_bwd");
}

extern "C" int mphsir_axpby(const float* X, int ldx, float* Y, int ldy, long long M, int C, float alpha, float beta,
                            const float* row_scale, int rows_per_batch, int x_row_mod, void* stream) {
  MPHSIR_REQUIRE(X && Y && M > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "axpby: bad arguments");
  MPHSIR_REQUIRE(row_scale == nullptr || rows_per_batch > 0, "axpby: row_scale needs rows_per_batch");
  axpby_kernel<<<grid_for(M * (C / 4)), 256, 0, ST(stream)>>>(X, ldx, Y, ldy, M, C, alpha, beta, row_scale, rows_per_batch, x_row_mod);
  return check_launch("axpby");
}

extern "C" int mphsir_batch_sum(const float* X, int ldx, float* Y, int ldy, int B, long long rows, int C, void* stream) {
  MPHSIR_REQUIRE(X && Y && B > 0 && rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "batch_sum: bad arguments");
  batch_sum_kernel<<<grid_for(rows * (C / 4)), 256, 0, ST(stream)>>>(X, ldx, Y, ldy, B, rows, C);
  return check_launch("batch_sum");
}

extern "C" int mphsir_colsum(const float* X, int ldx, float* out, long long M, int C, int map_mode, int map_a, int map_b,
                             void* stream) {
  MPHSIR_REQUIRE(X && out && M > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0, "colsum: C and ldx must be multiples of 4");
  const int groups = (C + 127) / 128;
  long long chunks = (2LL * sm_count() + groups - 1) / groups;
  if (chunks > (M + 63) / 64) chunks = (M + 63) / 64;
  if (chunks < 1) chunks = 1;
  const int rows_per_cta = (int)((M + chunks - 1) / chunks);
  dim3 grid((unsigned)((M + rows_per_cta - 1) / rows_per_cta), groups);
  colsum_kernel<<<grid, 256, 0, ST(stream)>>>(X, ldx, out, M, C, rows_per_cta, map_mode, map_a, map_b);
  return check_launch("colsum");
}

extern "C" int mphsir_window_reduce(const float* A, int lda, const float* Bm, int ldb, float* out, int B, int H, int W, int C,
                                    int shift, float scale, void* stream) {
  MPHSIR_REQUIRE(A && out && B > 0 && H % 8 == 0 && W % 8 == 0 && C % 4 == 0 && lda % 4 == 0 && (Bm == nullptr || ldb % 4 == 0),
                 "window_reduce: bad arguments");
  const int n_win = B * (H / 8) * (W / 8);
  dim3 grid(n_win, (C + 127) / 128);
  window_reduce_kernel<<<grid, 256, 0, ST(stream)>>>(A, lda, Bm, ldb, out, n_win, H, W, C, shift, scale);
  return check_launch("window_reduce");
}

extern "C" int mphsir_gate_apply_bwd(const float* dU, int ldu, const float* gate, const float* dMean, float* dSA, int lds,
                                     int B, int H, int W, int C, int shift, void* stream) {
  MPHSIR_REQUIRE(dU && gate && dMean && dSA && B > 0 && H % 8 == 0 && W % 8 == 0 && C % 4 == 0 && ldu % 4 == 0 && lds % 4 == 0,
                 "gate_apply_bwd: bad arguments");
  gate_apply_bwd_kernel<<<grid_for((long long)B * H * W * (C / 4)), 256, 0, ST(stream)>>>(dU, ldu, gate, dMean, dSA, lds, B, H, W, C, shift);
  return check_launch("gate_apply_bwd");
}

extern "C" int mphsir_gate_apply_fwd(const float* X, int ldx, const float* SA, int lds, const float* gate, const float* row_scale,
                                     float* U, int ldu, int B, int H, int W, int C, int shift, void* stream) {
  MPHSIR_REQUIRE(X && SA && gate && U && B > 0 && H % 8 == 0 && W % 8 == 0 && C % 4 == 0 && ldx % 4 == 0 && lds % 4 == 0 && ldu % 4 == 0,
                 "gate_apply_fwd: bad arguments");
  gate_apply_fwd_kernel<<<grid_for((long long)B * H * W * (C / 4)), 256, 0, ST(stream)>>>(X, ldx, SA, lds, gate, row_scale, U, ldu, B, H, W,
                                                                                          C, shift);
  return check_launch("gate_apply_fwd");
}

extern "C" int mphsir_dwconv3x3_wgrad(const float* X, int ldx, const float* dY, int ldy, float* dW, int B, int H, int W, int C,
                                      int map_mode, int map_a, int map_b, void* stream) {
  MPHSIR_REQUIRE(X && dY && dW && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0,
                 "dwconv3x3_wgrad: bad arguments");
  const int groups = (C + 255) / 256;
  const long long rows = (long long)B * H;
  // one full wave (2 CTAs of 256 threads per SM at 122 registers), at least 2 rows per lane so the reduction +
  // atomics amortise
  long long ctas = (2LL * sm_count()) / groups;
  if (ctas < 1) ctas = 1;
  long long rpl = (rows + ctas * 4 - 1) / (ctas * 4);
  if (rpl < 2) rpl = 2;
  dim3 grid((unsigned)((rows + rpl * 4 - 1) / (rpl * 4)), groups);
  dwconv3x3_wgrad_kernel<<<grid, 256, 0, ST(stream)>>>(X, ldx, dY, ldy, dW, B, H, W, C, (int)rpl, map_mode, map_a, map_b);
  return check_launch("dwconv3x3_wgrad");
}

extern "C" int mphsir_pixel_unshuffle(const float* in, int ldi, float* out, int ldo, int B, int H, int W, int C, void* stream) {
  MPHSIR_REQUIRE(in && out && B > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && ldi >= C && ldo >= 4 * C, "pixel_unshuffle: bad arguments");
  pixel_unshuffle_kernel<<<grid_for((long long)B * H * W * C), 256, 0, ST(stream)>>>(in, ldi, out, ldo, B, H, W, C);
  return check_launch("pixel_unshuffle");
}

extern "C" int mphsir_pixel_shuffle(const float* in, int ldi, float* out, int ldo, int B, int H, int W, int C, void* stream) {
  MPHSIR_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0 && ldi >= 4 * C && ldo >= C, "pixel_shuffle: bad arguments");
  pixel_shuffle_kernel<<<grid_for((long long)B * H * W * 4 * C), 256, 0, ST(stream)>>>(in, ldi, out, ldo, B, H, W, C);
  return check_launch("pixel_shuffle");
}

extern "C" int mphsir_tokens_to_nchw(const float* in, int ld, float* out, int B, int C, int HW, void* stream) {
  MPHSIR_REQUIRE(in && out && B > 0 && C > 0 && HW > 0 && ld >= C, "tokens_to_nchw: bad arguments");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  tokens_to_nchw_kernel<<<grid, 256, 0, ST(stream)>>>(in, ld, out, C, HW);
  return check_launch("tokens_to_nchw");
}

extern "C" int mphsir_bilinear_bwd(const float* dY, int ldy, float* dX, int ldx, int B, int h, int w, int H, int W, int C,
                                   void* stream) {
  MPHSIR_REQUIRE(dY && dX && B > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0 && ldy >= C && ldx >= C, "bilinear_bwd: bad arguments");
  bilinear_bwd_kernel<<<grid_for((long long)B * H * W * C), 256, 0, ST(stream)>>>(dY, ldy, dX, ldx, B, h, w, H, W, C);
  return check_launch("bilinear_bwd");
}

extern "C" int mphsir_tvsp_query_bwd(const float* dQ, int ldq, const float* clip_b, const float* weights, float* dLearnable,
                                     int B, int T, int D, int ps, void* stream) {
  MPHSIR_REQUIRE(dQ && clip_b && weights && dLearnable && B > 0 && T > 0 && D > 0 && D <= 1024 && ps > 0 && ldq >= D,
                 "tvsp_query_bwd: bad arguments");
  const int ppc = 64;
  dim3 grid((ps * ps + ppc - 1) / ppc, B);
  tvsp_query_bwd_kernel<<<grid, D, 0, ST(stream)>>>(dQ, ldq, clip_b, weights, dLearnable, B, T, D, ps, ppc);
  return check_launch("tvsp_query_bwd");
}

extern "C" int mphsir_l1_clamp_loss(const float* out, const float* clean, float* dOut, float* loss, long long numel,
                                    float grad_scale, void* stream) {
  MPHSIR_REQUIRE(out && clean && dOut && loss && numel > 0, "l1_clamp_loss: bad arguments");
  l1_clamp_loss_kernel<<<grid_for(numel, 1024, 4), 256, 0, ST(stream)>>>(out, clean, dOut, loss, numel, grad_scale);
  return check_launch("l1_clamp_loss");
}

extern "C" int mphsir_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int step, float grad_scale, const float* dyn, void* stream) {
  MPHSIR_REQUIRE(p && g && m && v && n > 0 && step > 0, "adamw_step: bad arguments");
  MPHSIR_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                   reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adamw_step: buffers must be 16-byte aligned");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_kernel<<<grid_for(n / 4 + 4), 256, 0, ST(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale, dyn);
  return check_launch("adamw_step");
}
