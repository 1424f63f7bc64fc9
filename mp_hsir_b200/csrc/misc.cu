// Small layout / prompt kernels: NCHW->tokens, Text_Prompt mixing, TVSP query map, bilinear resize.
#include "common.cuh"

namespace mphsir {

// in [B,C,HW] -> out [B*HW, ld] (channels >= C zero-filled).  32x32 smem transpose tiles.
__global__ void __launch_bounds__(256) nchw_to_tokens_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             int C, int HW, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    tile[r][tx] = (c < C && p < HW) ? __ldg(in + ((long long)b * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    if (p < HW && c < ld) out[((long long)b * HW + p) * ld + c] = tile[tx][r];
  }
}

// clip_b[b, :] = (sum_t w[b,t] clip[t,:]) / T          (Text_Prompt.forward, net/MP_HSIR.py:528-530)
__global__ void text_prompt_kernel(const float* __restrict__ w, const float* __restrict__ clip,
                                   float* __restrict__ clip_b, int B, int T) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 512) return;
  const int b = idx >> 9, e = idx & 511;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s = fmaf(__ldg(w + b * T + t), __ldg(clip + t * 512 + e), s);
  clip_b[idx] = s / (float)T;
}

// Q[(b,i,j), d] = tp[b,d] * clip_b[floor(i*B/ps), floor(j*512/ps)],  tp = (w @ learnable)/T
// (TVSP.forward, net/MP_HSIR.py:575-577: broadcast [B,D,1,1]*[B,512] -> [B,D,B,512], nearest resize)
__global__ void __launch_bounds__(256) tvsp_query_kernel(const float* __restrict__ clip_b,
                                                         const float* __restrict__ w,
                                                         const float* __restrict__ learnable,
                                                         float* __restrict__ Q, int B, int T, int D, int ps) {
  extern __shared__ float tp[];  // [D]
  const int b = blockIdx.y;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s = fmaf(__ldg(w + b * T + t), __ldg(learnable + t * D + d), s);
    tp[d] = s / (float)T;
  }
  __syncthreads();
  const int d4n = D >> 2;
  const int per_block = ps * ps * d4n / gridDim.x;
  const int e0 = blockIdx.x * per_block;
  for (int e = e0 + threadIdx.x; e < e0 + per_block; e += blockDim.x) {
    const int d4 = e % d4n;
    const int pix = e / d4n;
    const int i = pix / ps, j = pix - i * ps;
    const int si = min((int)floorf((float)i * ((float)B / (float)ps)), B - 1);
    const int sj = min((int)floorf((float)j * (512.0f / (float)ps)), 511);
    const float s = __ldg(clip_b + si * 512 + sj);
    const float4 t4 = *reinterpret_cast<const float4*>(&tp[d4 * 4]);
    *reinterpret_cast<float4*>(Q + ((long long)b * ps * ps + pix) * D + d4 * 4) =
        make_float4(t4.x * s, t4.y * s, t4.z * s, t4.w * s);
  }
}

// F.interpolate(mode="bilinear", align_corners=False): src = max((dst+0.5)*in/out-0.5, 0)
__global__ void __launch_bounds__(256) bilinear_kernel(const float* __restrict__ X, long long ldx,
                                                       float* __restrict__ Y, long long ldy, int B, int h, int w,
                                                       int H, int W, int C) {
  const int c4n = C >> 2;
  const long long total = (long long)B * H * W * c4n;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4n) * 4;
    const long long pix = idx / c4n;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    const float fy = fmaxf(((float)y + 0.5f) * sy - 0.5f, 0.f);
    const float fx = fmaxf(((float)x + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* base = X + (long long)b * h * w * ldx + c;
    const float4 v00 = ldg4(base + ((long long)y0 * w + x0) * ldx);
    const float4 v01 = ldg4(base + ((long long)y0 * w + x1) * ldx);
    const float4 v10 = ldg4(base + ((long long)y1 * w + x0) * ldx);
    const float4 v11 = ldg4(base + ((long long)y1 * w + x1) * ldx);
    // same association as the oracle / ATen: rows first (1-ly, ly), then columns (1-lx, lx)
    float4 o;
    {
      const float r0x = v00.x * (1.f - ly) + v10.x * ly, r1x = v01.x * (1.f - ly) + v11.x * ly;
      const float r0y = v00.y * (1.f - ly) + v10.y * ly, r1y = v01.y * (1.f - ly) + v11.y * ly;
      const float r0z = v00.z * (1.f - ly) + v10.z * ly, r1z = v01.z * (1.f - ly) + v11.z * ly;
      const float r0w = v00.w * (1.f - ly) + v10.w * ly, r1w = v01.w * (1.f - ly) + v11.w * ly;
      o.x = r0x * (1.f - lx) + r1x * lx;
      o.y = r0y * (1.f - lx) + r1y * lx;
      o.z = r0z * (1.f - lx) + r1z * lx;
      o.w = r0w * (1.f - lx) + r1w * lx;
    }
    *reinterpret_cast<float4*>(Y + pix * ldy + c) = o;
  }
}

}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_nchw_to_tokens(const float* in, float* out, int B, int C, int HW, int ld_out, void* stream) {
  MPHSIR_REQUIRE(in && out && B > 0 && C > 0 && HW > 0 && ld_out >= C, "nchw_to_tokens: bad arguments");
  dim3 grid((HW + 31) / 32, (ld_out + 31) / 32, B);
  nchw_to_tokens_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, C, HW, ld_out);
  return check_launch("nchw_to_tokens");
}

extern "C" int mphsir_text_prompt_fwd(const float* weights, const float* clip, float* clip_b, int B, int T,
                                      void* stream) {
  MPHSIR_REQUIRE(weights && clip && clip_b && B > 0 && T > 0, "text_prompt: bad arguments");
  text_prompt_kernel<<<(B * 512 + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(weights, clip, clip_b, B, T);
  return check_launch("text_prompt");
}

extern "C" int mphsir_tvsp_query_fwd(const float* clip_b, const float* weights, const float* learnable, float* Q,
                                     int B, int T, int D, int ps, void* stream) {
  MPHSIR_REQUIRE(clip_b && weights && learnable && Q, "tvsp_query: null operand");
  MPHSIR_REQUIRE(B > 0 && T > 0 && D > 0 && D % 4 == 0 && ps > 0, "tvsp_query: bad shape");
  int nblk = 16;
  while ((ps * ps * (D / 4)) % nblk != 0) nblk >>= 1;
  dim3 grid(nblk, B);
  tvsp_query_kernel<<<grid, 256, sizeof(float) * D, reinterpret_cast<cudaStream_t>(stream)>>>(clip_b, weights, learnable, Q, B, T, D, ps);
  return check_launch("tvsp_query");
}

extern "C" int mphsir_bilinear_fwd(const float* X, int ldx, float* Y, int ldy, int B, int h, int w, int H, int W,
                                   int C, void* stream) {
  MPHSIR_REQUIRE(X && Y && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "bilinear: bad shape");
  MPHSIR_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldx >= C && ldy >= C, "bilinear: C/ld must be multiples of 4");
  const long long total = (long long)B * H * W * (C / 4);
  const int blocks = (int)((total + 255) / 256 > 148LL * 64 ? 148LL * 64 : (total + 255) / 256);
  bilinear_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, ldx, Y, ldy, B, h, w, H, W, C);
  return check_launch("bilinear");
}
