// Evaluation metrics and training-time degradations on the device: what wraps the hot path in
// test.py / train.py once the model itself runs at tens of cubes per second.
//
//  * psnr_ssim_kernel — utils/val_utils.py:49-69 (compute_psnr_ssim): per band, on clip(.,0,1), data_range 1:
//      PSNR  = 10 log10(1 / mean((x-y)^2))                     (skimage.metrics.peak_signal_noise_ratio)
//      SSIM  = mean over the (H-6)x(W-6) windows that fit the image of
//              ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)),   C1 = 0.01^2, C2 = 0.03^2,
//              7x7 uniform window, sample covariance (x 49/48)   (skimage.metrics.structural_similarity defaults:
//              win_size 7, use_sample_covariance, the 3-pixel border of the filtered image is cropped, so only whole
//              windows contribute and the filter's border mode never matters)
//    One pass over the two NCHW planes: a CTA stages a (32+6) x (128+6) tile of both images in shared memory, every
//    thread slides a 7-row ring of horizontal 7-sums (x, y, xx, yy, xy) down one window column.  Sums in fp64 (the
//    variance terms cancel against C2 = 9e-4).  Output: per plane { sum (x-y)^2, sum of window SSIMs }.  HBM-bound:
//    8 bytes per pixel.
//  * degrade_kernel — the array-only degradations train.py draws per sample (utils/dataset_utils.py:128-146 ->
//    utils/degradation_utils.py:25-39 Gaussian / non-iid noise, :227-233 random mask, :275-284 band loss) as ONE
//    elementwise pass:  out = clean * keep[b,c] * (u > mask_ratio[b]) + sigma[b,c] * n,  u ~ U[0,1), n ~ N(0,1) from a
//    counter-based Philox4x32-10 stream keyed by (seed, element index) — reproducible, no state, any launch shape.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace mphsir {
namespace metrics {

constexpr int TW = 128, TH = 32, WIN = 7;
constexpr int SW = TW + WIN - 1, SH = TH + WIN - 1;

__global__ void __launch_bounds__(TW) psnr_ssim_kernel(const float* __restrict__ X, const float* __restrict__ Y, int H, int W,
                                                       double* __restrict__ sums) {
  __shared__ float sx[SH][SW + 1];
  __shared__ float sy[SH][SW + 1];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const float* xp = X + (size_t)plane * H * W;
  const float* yp = Y + (size_t)plane * H * W;
  const int t = threadIdx.x;

  double sse = 0.0;
  for (int idx = t; idx < SH * SW; idx += TW) {
    const int r = idx / SW, c = idx - r * SW;
    const int gy = y0 + r, gx = x0 + c;
    float a = 0.f, b = 0.f;
    if (gy < H && gx < W) {
      a = fminf(fmaxf(__ldg(xp + (size_t)gy * W + gx), 0.f), 1.f);   // np.clip(., 0, 1), val_utils.py:51-52
      b = fminf(fmaxf(__ldg(yp + (size_t)gy * W + gx), 0.f), 1.f);
      if (r < TH && c < TW) {  // the tile's own pixels: every pixel of the image is owned by exactly one tile
        const double d = (double)a - (double)b;
        sse += d * d;
      }
    }
    sx[r][c] = a;
    sy[r][c] = b;
  }
  __syncthreads();

  // window column x0 + t: top-left corners (y0 + r, x0 + t), r = 0..TH-1, valid while the window fits the image
  double ssim = 0.0;
  if (x0 + t + WIN <= W) {
    double ring[WIN][5];
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03, NP = 49.0, COVN = 49.0 / 48.0;
#pragma unroll
    for (int r = 0; r < SH; ++r) {
      double h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
#pragma unroll
      for (int j = 0; j < WIN; ++j) {
        const double a = sx[r][t + j], b = sy[r][t + j];
        h0 += a; h1 += b; h2 += a * a; h3 += b * b; h4 += a * b;
      }
      v0 += h0; v1 += h1; v2 += h2; v3 += h3; v4 += h4;
      if (r >= WIN) {
        v0 -= ring[r % WIN][0]; v1 -= ring[r % WIN][1]; v2 -= ring[r % WIN][2]; v3 -= ring[r % WIN][3]; v4 -= ring[r % WIN][4];
      }
      ring[r % WIN][0] = h0; ring[r % WIN][1] = h1; ring[r % WIN][2] = h2; ring[r % WIN][3] = h3; ring[r % WIN][4] = h4;
      if (r >= WIN - 1) {
        const int top = y0 + r - (WIN - 1);
        if (top + WIN <= H) {
          const double ux = v0 / NP, uy = v1 / NP;
          const double vx = COVN * (v2 / NP - ux * ux), vy = COVN * (v3 / NP - uy * uy), vxy = COVN * (v4 / NP - ux * uy);
          ssim += ((2.0 * ux * uy + C1) * (2.0 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
        }
      }
    }
  }
  // CTA reduction -> one atomic per CTA and quantity
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sse += __shfl_xor_sync(0xffffffffu, sse, o);
    ssim += __shfl_xor_sync(0xffffffffu, ssim, o);
  }
  __shared__ double red[2][TW / 32];
  if ((t & 31) == 0) {
    red[0][t >> 5] = sse;
    red[1][t >> 5] = ssim;
  }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int w = 0; w < TW / 32; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    atomicAdd(sums + 2 * plane, a);
    atomicAdd(sums + 2 * plane + 1, b);
  }
}

// per plane: 1.0 if every element is exactly zero (compute_psnr_ssim2, val_utils.py:88: bands the degradation removed)
__global__ void __launch_bounds__(256) plane_all_zero_kernel(const float* __restrict__ X, long long hw, int* __restrict__ nonzero) {
  const float* p = X + (size_t)blockIdx.y * hw;
  int any = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < hw; i += (long long)gridDim.x * 256) any |= (__ldg(p + i) != 0.f);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0 && any) atomicOr(nonzero + blockIdx.y, 1);
}

// ---- Philox4x32-10 (Salmon et al., SC'11): counter = element index / 4, key = seed ---------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// element e of the [B, C, HW] batch: stream block q = e / 2 yields (u0, u1, u2, u3); element parity picks the pair:
//   mask uniform  u = (w0 >> 8) * 2^-24                      in [0, 1)
//   normal        n = sqrt(-2 ln((w1 >> 8) + 1) * 2^-24)) * cos(2 pi (w0' ...))   (Box-Muller on two further words)
// Layout per block of 4 words serving 2 elements: element 2q uses (w0: mask, w1+w2: Box-Muller cos branch),
// element 2q+1 uses (w3: mask, w1+w2: Box-Muller sin branch).
__global__ void __launch_bounds__(256) degrade_kernel(const float* __restrict__ clean, float* __restrict__ out, long long total,
                                                      long long hw, int C, const float* __restrict__ sigma /* [B*C] */,
                                                      const float* __restrict__ keep /* [B*C] */,
                                                      const float* __restrict__ mask_ratio /* [B] */, uint32_t seed_lo,
                                                      uint32_t seed_hi) {
  const long long pairs = (total + 1) / 2;
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < pairs; q += (long long)gridDim.x * 256) {
    uint32_t w[4];
    philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), 0u, 0u, seed_lo, seed_hi, w);
    const float u1 = ((float)(w[1] >> 8) + 1.0f) * (1.0f / 16777216.0f);   // (0, 1]
    const float u2 = (float)(w[2] >> 8) * (1.0f / 16777216.0f);            // [0, 1)
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long e = 2 * q + h;
      if (e >= total) break;
      const long long bc = e / hw;
      const int b = (int)(bc / C);
      const float um = (float)((h == 0 ? w[0] : w[3]) >> 8) * (1.0f / 16777216.0f);
      const float n = rad * (h == 0 ? cs : sn);
      const float m = um > __ldg(mask_ratio + b) ? 1.f : 0.f;   // np.random.rand(C,H,W) > mask_ratio (degradation_utils.py:230)
      out[e] = __ldg(clean + e) * __ldg(keep + bc) * m + __ldg(sigma + bc) * n;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// The structured half of 'complexN' (utils/degradation_utils.py:296-316: after the non-iid Gaussian noise ONE of deadline /
// impulse / stripe noise on a third of the bands), in place on the output of degrade_kernel:
//   deadline (:57-68)  columns of a band set to zero          -> colmul[b,c,x] = 0
//   stripe   (:41-55)  a per-column offset subtracted          -> coladd[b,c,x] = -stripe
//   impulse  (:70-84)  pixels flipped with probability p to 1 (salt, probability 1/2) or 0 (pepper) -> impulse[b,c] = p
// Which bands / columns / amounts is drawn by the host like the reference draws it (a few KB); the per-pixel impulse
// decisions come from a second Philox4x32-10 stream (counter word 2 = 1) so they are independent of the first pass.
// Samples with active[b] == 0 are skipped.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) degrade_structured_kernel(float* __restrict__ x, long long total, int C, int H, int W,
                                                                 const float* __restrict__ colmul, const float* __restrict__ coladd,
                                                                 const float* __restrict__ impulse, const int* __restrict__ active,
                                                                 uint32_t seed_lo, uint32_t seed_hi) {
  const long long hw = (long long)H * W;
  const long long pairs = (total + 1) / 2;
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < pairs; q += (long long)gridDim.x * 256) {
    uint32_t w[4];
    bool drawn = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long e = 2 * q + h;
      if (e >= total) break;
      const long long bc = e / hw;
      const int b = (int)(bc / C);
      if (__ldg(active + b) == 0) continue;
      const int xcol = (int)((e - bc * hw) % W);
      float v = x[e] * __ldg(colmul + bc * W + xcol) + __ldg(coladd + bc * W + xcol);
      const float pflip = __ldg(impulse + bc);
      if (pflip > 0.f) {
        if (!drawn) {
          philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), 1u, 0u, seed_lo, seed_hi, w);
          drawn = true;
        }
        const float uf = (float)((h == 0 ? w[0] : w[2]) >> 8) * (1.0f / 16777216.0f);
        const float us = (float)((h == 0 ? w[1] : w[3]) >> 8) * (1.0f / 16777216.0f);
        if (uf < pflip) v = us < 0.5f ? 1.f : 0.f;
      }
      x[e] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gaussian blur degradation (utils/degradation_utils.py:91-108): depthwise conv of every band with the outer product of a
// normalised 1-D Gaussian, sigma = 0.3 ((k - 1) / 2 - 1) + 0.8, zero padding k / 2.  One CTA per (plane, 32 x 32 tile): the
// haloed tile is staged in shared memory, a horizontal pass writes a [32 + k - 1][32] strip, a vertical pass the output —
// the separable form of the reference's k x k kernel (its kernel_2d IS the outer product).  ksize[b] == 0: sample b is not
// a blur sample and its output plane is left alone (the elementwise degradation kernel wrote it).
// ---------------------------------------------------------------------------------------------
constexpr int BLUR_T = 32;       // output tile
constexpr int BLUR_KMAX = 21;    // largest kernel of the reference's de_dict ('blur': 9, 15, 21 / 7, 11, 15)

__global__ void __launch_bounds__(256) gaussian_blur_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            const int* __restrict__ ksize, int C, int H, int W, int tiles_x,
                                                            int tiles_per_plane) {
  __shared__ float tile[BLUR_T + BLUR_KMAX - 1][BLUR_T + BLUR_KMAX - 1 + 1];
  __shared__ float hrow[BLUR_T + BLUR_KMAX - 1][BLUR_T + 1];
  // padded to whole 16-byte groups: ptxas reads the weights with LDS.128, the last group reaches past kw[k - 1]
  __shared__ __align__(16) float kw[(BLUR_KMAX + 3) / 4 * 4];
  const int plane = blockIdx.x / tiles_per_plane;
  const int t = blockIdx.x - plane * tiles_per_plane;
  const int b = plane / C;
  const int k = ksize[b];
  if (k <= 0) return;
  const int r = k >> 1;
  const int ty0 = (t / tiles_x) * BLUR_T, tx0 = (t % tiles_x) * BLUR_T;
  const float* src = in + (long long)plane * H * W;
  if (threadIdx.x < k) {
    // kernel_1d = exp(-(x - mean)^2 / (2 sigma^2)) / sum, exactly as the reference builds it (fp32)
    const float sigma = 0.3f * ((float)(k - 1) * 0.5f - 1.0f) + 0.8f;
    const float mean = (float)(k - 1) / 2.0f;
    float sum = 0.f;
    for (int i = 0; i < k; ++i) {
      const float d = (float)i - mean;
      sum += expf(-(d * d) / (2.0f * sigma * sigma));
    }
    const float d = (float)threadIdx.x - mean;
    kw[threadIdx.x] = expf(-(d * d) / (2.0f * sigma * sigma)) / sum;
  }
  const int ext = BLUR_T + 2 * r;
  for (int e = threadIdx.x; e < ext * ext; e += 256) {
    const int yy = e / ext, xx = e - yy * ext;
    const int y = ty0 + yy - r, x = tx0 + xx - r;
    tile[yy][xx] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + (long long)y * W + x) : 0.f;   // zero padding
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ext * BLUR_T; e += 256) {
    const int yy = e / BLUR_T, xx = e - yy * BLUR_T;
    float a = 0.f;
    for (int i = 0; i < k; ++i) a = fmaf(kw[i], tile[yy][xx + i], a);
    hrow[yy][xx] = a;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < BLUR_T * BLUR_T; e += 256) {
    const int yy = e / BLUR_T, xx = e - yy * BLUR_T;
    const int y = ty0 + yy, x = tx0 + xx;
    if (y >= H || x >= W) continue;
    float a = 0.f;
    for (int i = 0; i < k; ++i) a = fmaf(kw[i], hrow[yy + i][xx], a);
    out[(long long)plane * H * W + (long long)y * W + x] = a;
  }
}


// ---------------------------------------------------------------------------------------------
// Super-resolution degradation 'sr' (utils/degradation_utils.py:165-176 then :189-200 via single_degrade :431-432): bicubic
// down-sampling of every band to (H / f, W / f) as torch.nn.functional.interpolate(mode='bicubic', align_corners=True) does
// it — source coordinate i * (H - 1) / (h - 1), cubic-convolution weights with A = -0.75 on the four rows / columns
// floor(src) - 1 .. + 2, indices clamped to the image — followed by the f x f pixel replication of `_resize`.  A thread per
// OUTPUT pixel recomputes its low-resolution sample (16 taps that hit L1): coalesced stores, no intermediate tensor.
// factor[b] == 0: sample b is not an 'sr' sample and its output planes are left alone.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_taps(int i, float scale, int n, int (&idx)[4], float (&w)[4]) {
  constexpr float A = -0.75f;
  const float real = scale * (float)i;
  int i0 = (int)floorf(real);
  if (i0 > n - 1) i0 = n - 1;
  float t = real - (float)i0;
  t = fminf(fmaxf(t, 0.f), 1.f);
  const float x0 = t + 1.f, x3 = 2.f - t, x2 = 1.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = i0 + j - 1;
    idx[j] = q < 0 ? 0 : (q > n - 1 ? n - 1 : q);
  }
}

__global__ void __launch_bounds__(256) sr_degrade_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         const int* __restrict__ factor, int C, int H, int W, long long total) {
  const long long hw = (long long)H * W;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const long long plane = e / hw;
    const int f = factor[plane / C];
    if (f <= 0) continue;
    const int rem = (int)(e - plane * hw);
    const int y = rem / W, x = rem - y * W;
    const int h = H / f, w = W / f;
    // pixels beyond h * f (H not a multiple of f) replicate the last low-resolution row / column
    const int yl = min(y / f, h - 1), xl = min(x / f, w - 1);
    const float sy = h > 1 ? (float)(H - 1) / (float)(h - 1) : 0.f;
    const float sx = w > 1 ? (float)(W - 1) / (float)(w - 1) : 0.f;
    int iy[4], ix[4];
    float wy[4], wx[4];
    cubic_taps(yl, sy, H, iy, wy);
    cubic_taps(xl, sx, W, ix, wx);
    const float* src = in + plane * hw;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float* row = src + (long long)iy[a] * W;
      float r = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) r += wx[b] * __ldg(row + ix[b]);
      acc += wy[a] * r;
    }
    out[e] = acc;
  }
}


// ---------------------------------------------------------------------------------------------
// Generic k x k blur degradations: circle blur (utils/degradation_utils.py:110-128), square blur (:150-163) and motion blur
// (:130-148) all are F.conv2d(x, kernel.repeat(C,1,1,1), padding=k//2, groups=C) with ONE host-built k x k kernel — the caller
// passes its taps (row-major, device memory), this kernel does the depthwise cross-correlation with zero padding.  Same
// tiling as the Gaussian blur (a haloed 32 x 32 tile per CTA), k*k taps per output.  active[b] == 0: planes left alone.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blur2d_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                     const float* __restrict__ taps, const int* __restrict__ active, int C, int H,
                                                     int W, int k, int tiles_x, int tiles_per_plane) {
  __shared__ float tile[BLUR_T + BLUR_KMAX - 1][BLUR_T + BLUR_KMAX - 1 + 1];
  __shared__ __align__(16) float kw[(BLUR_KMAX * BLUR_KMAX + 3) / 4 * 4];
  const int plane = blockIdx.x / tiles_per_plane;
  const int t = blockIdx.x - plane * tiles_per_plane;
  if (active[plane / C] == 0) return;
  const int r = k >> 1;
  const int ty0 = (t / tiles_x) * BLUR_T, tx0 = (t % tiles_x) * BLUR_T;
  const float* src = in + (long long)plane * H * W;
  for (int e = threadIdx.x; e < k * k; e += 256) kw[e] = __ldg(taps + e);
  const int ext = BLUR_T + 2 * r;
  for (int e = threadIdx.x; e < ext * ext; e += 256) {
    const int yy = e / ext, xx = e - yy * ext;
    const int y = ty0 + yy - r, x = tx0 + xx - r;
    tile[yy][xx] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + (long long)y * W + x) : 0.f;   // zero padding
  }
  __syncthreads();
  for (int e = threadIdx.x; e < BLUR_T * BLUR_T; e += 256) {
    const int yy = e / BLUR_T, xx = e - yy * BLUR_T;
    const int y = ty0 + yy, x = tx0 + xx;
    if (y >= H || x >= W) continue;
    float a = 0.f;
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < k; ++j) a = fmaf(kw[i * k + j], tile[yy + i][xx + j], a);
    out[(long long)plane * H * W + (long long)y * W + x] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// Poisson noise 'poissonN' (utils/degradation_utils.py:86-89): out = Poisson(max(x, 0) * scale) / scale.  One uniform per
// element from the Philox4x32-10 stream with counter word 2 = 2 (element e: word e % 4 of block e / 4), inverted through the
// Poisson CDF by sequential search in double precision (lambda <= scale: the reference uses scale 10 on [0, 1] data, so a
// handful of terms).  scale[b] <= 0: sample b is not a Poisson sample and its output is left alone.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) poisson_kernel(const float* __restrict__ in, float* __restrict__ out, long long total,
                                                      long long chw, const float* __restrict__ scale, uint32_t seed_lo,
                                                      uint32_t seed_hi) {
  const long long quads = (total + 3) / 4;
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < quads; q += (long long)gridDim.x * 256) {
    uint32_t w[4];
    philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), 2u, 0u, seed_lo, seed_hi, w);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const long long e = 4 * q + h;
      if (e >= total) break;
      const float sc = __ldg(scale + e / chw);
      if (sc <= 0.f) continue;
      const double u = (double)(w[h] >> 8) * (1.0 / 16777216.0);
      const double lam = (double)(fmaxf(__ldg(in + e), 0.f) * sc);
      double p = exp(-lam), cdf = p;
      int n = 0;
      while (u > cdf && n < 4096) {
        ++n;
        p *= lam / (double)n;
        cdf += p;
      }
      out[e] = (float)n / sc;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Haze degradation (utils/degradation_utils.py:235-273), given the cirrus-band map the reference reads from its .mat files:
//   A[c]   = mean of the top_k brightest pixels of band c                 (atmospheric light, :257-261)
//   t1     = 1 - omega * cirrus,  <= 0 -> 1e-10                           (:263-264)
//   T[c]   = exp((lambda_0 / lambda_c)^gamma * log(t1))                   (:269-270; the exponent per band comes from the host)
//   out[c] = x[c] * T[c] + A[c] * (1 - T[c])                              (:271)
// topk_mean_kernel: one CTA per plane; pass j finds the largest value strictly below the previous pass's value and how often it
// occurs (a (max, count) reduction), so ties are counted like a sort would count them; k passes, sum in double.  k is
// max(int(HW * 1e-4), 1) in the reference: 1 for a 64 x 64 patch (the maximum), 26 for 512 x 512.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) topk_mean_kernel(const float* __restrict__ X, long long hw, int k, float* __restrict__ mean) {
  __shared__ float smax[8];
  __shared__ int scnt[8];
  __shared__ float bmax;
  __shared__ int bcnt;
  const float* p = X + (size_t)blockIdx.x * hw;
  float bound = INFINITY;
  double sum = 0.0;
  int remaining = k;
  bool first = true;
  while (remaining > 0) {
    float m = -INFINITY;
    int c = 0;
    for (long long i = threadIdx.x; i < hw; i += 256) {
      const float v = __ldg(p + i);
      if (!(first || v < bound)) continue;           // pass 0 takes everything (also +inf); NaN never compares true afterwards
      if (v > m) { m = v; c = 1; }
      else if (v == m) ++c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const int c2 = __shfl_xor_sync(0xffffffffu, c, o);
      if (m2 > m) { m = m2; c = c2; }
      else if (m2 == m) c += c2;
    }
    if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = m; scnt[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float mm = smax[0];
      int cc = scnt[0];
      for (int w = 1; w < 8; ++w) {
        if (smax[w] > mm) { mm = smax[w]; cc = scnt[w]; }
        else if (smax[w] == mm) cc += scnt[w];
      }
      bmax = mm; bcnt = cc;
    }
    __syncthreads();
    const float mm = bmax;
    const int cc = bcnt;
    __syncthreads();
    if (cc == 0) break;                              // fewer than k comparable values
    const int take = cc < remaining ? cc : remaining;
    sum += (double)mm * take;
    remaining -= take;
    bound = mm;
    first = false;
  }
  if (threadIdx.x == 0) mean[blockIdx.x] = (float)(sum / (double)(k - remaining > 0 ? k - remaining : 1));
}

__global__ void __launch_bounds__(256) haze_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                   const float* __restrict__ cirrus, const float* __restrict__ omega,
                                                   const float* __restrict__ expo, const float* __restrict__ light, int C,
                                                   long long hw, long long total) {
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const long long bc = e / hw;
    const int b = (int)(bc / C), c = (int)(bc - (long long)b * C);
    const float om = __ldg(omega + b);
    if (om <= 0.f) continue;
    double t1 = 1.0 - (double)om * (double)__ldg(cirrus + (long long)b * hw + (e - bc * hw));
    if (t1 <= 0.0) t1 = 1e-10;
    const double T = exp((double)__ldg(expo + c) * log(t1));
    out[e] = (float)((double)__ldg(in + e) * T + (double)__ldg(light + bc) * (1.0 - T));
  }
}

}  // namespace metrics
}  // namespace mphsir

using namespace mphsir;

extern "C" int mphsir_psnr_ssim(const float* restored, const float* clean, int planes, int H, int W, double* sums, void* stream) {
  MPHSIR_REQUIRE(restored && clean && sums && planes > 0, "psnr_ssim: null operand");
  MPHSIR_REQUIRE(H >= 7 && W >= 7, "psnr_ssim: the 7x7 SSIM window needs H, W >= 7 (got %dx%d)", H, W);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * planes, st) != cudaSuccess) {
    set_error("psnr_ssim: cudaMemsetAsync failed");
    return MPHSIR_ERR_CUDA;
  }
  for (int p0 = 0; p0 < planes; p0 += 65535) {
    const int n = planes - p0 < 65535 ? planes - p0 : 65535;
    dim3 grid((W + metrics::TW - 1) / metrics::TW, (H + metrics::TH - 1) / metrics::TH, n);
    metrics::psnr_ssim_kernel<<<grid, metrics::TW, 0, st>>>(restored + (size_t)p0 * H * W, clean + (size_t)p0 * H * W, H, W, sums + 2 * p0);
  }
  return check_launch("psnr_ssim");
}

extern "C" int mphsir_plane_nonzero(const float* X, int planes, long long hw, int* nonzero, void* stream) {
  MPHSIR_REQUIRE(X && nonzero && planes > 0 && planes <= 65535 && hw > 0, "plane_nonzero: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(nonzero, 0, sizeof(int) * planes, st) != cudaSuccess) {
    set_error("plane_nonzero: cudaMemsetAsync failed");
    return MPHSIR_ERR_CUDA;
  }
  long long gx = (hw + 255) / 256;
  if (gx > 64) gx = 64;
  metrics::plane_all_zero_kernel<<<dim3((unsigned)gx, planes), 256, 0, st>>>(X, hw, nonzero);
  return check_launch("plane_nonzero");
}

extern "C" int mphsir_degrade(const float* clean, float* out, int B, int C, long long hw, const float* sigma, const float* keep,
                              const float* mask_ratio, unsigned long long seed, void* stream) {
  MPHSIR_REQUIRE(clean && out && sigma && keep && mask_ratio && B > 0 && C > 0 && hw > 0, "degrade: bad arguments");
  const long long total = (long long)B * C * hw;
  long long blocks = ((total + 1) / 2 + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  metrics::degrade_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      clean, out, total, hw, C, sigma, keep, mask_ratio, (uint32_t)seed, (uint32_t)(seed >> 32));
  return check_launch("degrade");
}

extern "C" int mphsir_gaussian_blur(const float* in, float* out, const int* ksize, int B, int C, int H, int W, int kmax, void* stream) {
  MPHSIR_REQUIRE(in && out && ksize && B > 0 && C > 0 && H > 0 && W > 0, "gaussian_blur: bad arguments");
  MPHSIR_REQUIRE(in != out, "gaussian_blur: in place is not supported (every output pixel reads a k x k neighbourhood)");
  MPHSIR_REQUIRE(kmax >= 0 && kmax <= metrics::BLUR_KMAX, "gaussian_blur: kernel sizes up to %d (got %d)", metrics::BLUR_KMAX, kmax);
  const int tiles_x = (W + metrics::BLUR_T - 1) / metrics::BLUR_T, tiles_y = (H + metrics::BLUR_T - 1) / metrics::BLUR_T;
  const long long blocks = (long long)B * C * tiles_x * tiles_y;
  MPHSIR_REQUIRE(blocks < (1LL << 31), "gaussian_blur: too many tiles");
  metrics::gaussian_blur_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, ksize, C, H, W, tiles_x,
                                                                                                    tiles_x * tiles_y);
  return check_launch("gaussian_blur");
}

extern "C" int mphsir_sr_degrade(const float* in, float* out, const int* factor, int B, int C, int H, int W, void* stream) {
  MPHSIR_REQUIRE(in && out && factor && B > 0 && C > 0 && H > 0 && W > 0, "sr_degrade: bad arguments");
  MPHSIR_REQUIRE(in != out, "sr_degrade: in place is not supported (every output pixel reads a 4 x 4 neighbourhood of the input)");
  const long long total = (long long)B * C * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  metrics::sr_degrade_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, factor, C, H, W, total);
  return check_launch("sr_degrade");
}

extern "C" int mphsir_blur2d(const float* in, float* out, const float* taps, const int* active, int B, int C, int H, int W, int k,
                             void* stream) {
  MPHSIR_REQUIRE(in && out && taps && active && B > 0 && C > 0 && H > 0 && W > 0, "blur2d: bad arguments");
  MPHSIR_REQUIRE(in != out, "blur2d: in place is not supported (every output pixel reads a k x k neighbourhood)");
  MPHSIR_REQUIRE(k >= 1 && k <= metrics::BLUR_KMAX && (k & 1), "blur2d: odd kernel sizes up to %d (got %d)", metrics::BLUR_KMAX, k);
  const int tiles_x = (W + metrics::BLUR_T - 1) / metrics::BLUR_T, tiles_y = (H + metrics::BLUR_T - 1) / metrics::BLUR_T;
  const long long blocks = (long long)B * C * tiles_x * tiles_y;
  MPHSIR_REQUIRE(blocks < (1LL << 31), "blur2d: too many tiles");
  metrics::blur2d_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, taps, active, C, H, W, k,
                                                                                             tiles_x, tiles_x * tiles_y);
  return check_launch("blur2d");
}

extern "C" int mphsir_poisson(const float* in, float* out, const float* scale, int B, long long chw, unsigned long long seed,
                              void* stream) {
  MPHSIR_REQUIRE(in && out && scale && B > 0 && chw > 0, "poisson: bad arguments");
  const long long total = (long long)B * chw;
  long long blocks = ((total + 3) / 4 + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  metrics::poisson_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, total, chw, scale,
                                                                                              (uint32_t)seed, (uint32_t)(seed >> 32));
  return check_launch("poisson");
}

extern "C" int mphsir_topk_mean(const float* X, int planes, long long hw, int k, float* mean, void* stream) {
  MPHSIR_REQUIRE(X && mean && planes > 0 && hw > 0 && k >= 1 && k <= hw, "topk_mean: bad arguments");
  MPHSIR_REQUIRE(k <= 4096, "topk_mean: k up to 4096 (one pass over the plane per distinct value; got %d)", k);
  metrics::topk_mean_kernel<<<(unsigned)planes, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, hw, k, mean);
  return check_launch("topk_mean");
}

extern "C" int mphsir_haze(const float* in, float* out, const float* cirrus, const float* omega, const float* expo,
                           const float* light, int B, int C, long long hw, void* stream) {
  MPHSIR_REQUIRE(in && out && cirrus && omega && expo && light && B > 0 && C > 0 && hw > 0, "haze: bad arguments");
  const long long total = (long long)B * C * hw;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  metrics::haze_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, cirrus, omega, expo, light, C, hw,
                                                                                           total);
  return check_launch("haze");
}

extern "C" int mphsir_degrade_structured(float* x, int B, int C, int H, int W, const float* colmul, const float* coladd,
                                         const float* impulse, const int* active, unsigned long long seed, void* stream) {
  MPHSIR_REQUIRE(x && colmul && coladd && impulse && active && B > 0 && C > 0 && H > 0 && W > 0, "degrade_structured: bad arguments");
  const long long total = (long long)B * C * H * W;
  long long blocks = ((total + 1) / 2 + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  metrics::degrade_structured_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, total, C, H, W, colmul, coladd, impulse, active, (uint32_t)(seed & 0xFFFFFFFFull), (uint32_t)(seed >> 32));
  return check_launch("degrade_structured");
}
