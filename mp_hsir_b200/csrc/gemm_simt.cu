// fp32 SIMT GEMM / implicit-GEMM 3x3 conv with fused LayerNorm prologue and fused epilogues.
//
// This is the exact-fp32 engine of the library (FFMA, fp32 accumulate): it carries every
// nn.Linear / 1x1 conv / dense 3x3 conv of the hot path (see include/mphsir.h for the reference
// call sites) and is the numerical baseline the tensor-core engine is checked against.
//
// Tiling: CTA tile 128 x BN (BN = 64 or 128), K step 16, 256 threads, 8 x (BN/16) outputs per
// thread, double-buffered shared memory with register prefetch (one __syncthreads per K step).
// A is read token-major (row = pixel) and transposed into shared memory so the inner loop is
// 2-3 LDS.128 per 32-64 FFMA; Bt is pre-packed "in x out" so its tile loads are already in
// inner-loop order.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace mphsir {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NT = 256;
constexpr int AS_LD = BM + 4;

enum { MODE_PLAIN = 0, MODE_LN = 1, MODE_CONV = 2 };
enum { OUT_UNSHUFFLE = 10, OUT_SHUFFLE = 11, OUT_NCHW_RES = 12 };

struct GemmArgs {
  const float* A;
  long long lda;
  int a_row_mod;
  int Ka;
  const float* Bt;
  long long ldb;
  long long b_batch_stride;
  int rows_per_batch;
  int tiles_per_batch;  // >0: m-tiles never straddle samples (per-sample weights)
  float* Y;
  long long ldy;
  int M, N, nk;
  const float* ln_g;
  const float* ln_b;
  const float* bias;
  int epi;
  const float* res1;
  long long ldr1;
  const float* res2;
  long long ldr2;
  const float* gsrc;
  long long ldg;
  const float* gate;
  int H, W, shift;
  const float* row_scale;
  float* Y2;  // EPI_PROJ: second output (columns >= n_split)
  int ldy2, n_split;
  int Cin;
  const float* R;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(NT, 2) gemm_kernel(const GemmArgs p) {
  constexpr int TN = BN / 16;
  constexpr int NG = TN / 4;  // float4 column groups per thread
  constexpr int B_LD4 = BN / 4;
  constexpr int B_PER_THREAD = (BK * B_LD4) / NT;

  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ float2 stats[MODE == MODE_LN ? BM : 1];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  int m0, m_end;
  const float* Bt = p.Bt;
  if (p.tiles_per_batch > 0) {
    const int b = blockIdx.x / p.tiles_per_batch;
    const int t = blockIdx.x - b * p.tiles_per_batch;
    m0 = b * p.rows_per_batch + t * BM;
    m_end = (b + 1) * p.rows_per_batch;
    Bt += (long long)b * p.b_batch_stride;
  } else {
    m0 = blockIdx.x * BM;
    m_end = p.M;
  }
  const int n0 = blockIdx.y * BN;

  // ---- per-thread A rows -------------------------------------------------------------
  const int kq = tid & 3;
  bool valid[2];
  long long arow[2];
  int pb[2], py[2], px[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + (tid >> 2) + i * 64;
    valid[i] = m < m_end;
    const int mm = valid[i] ? m : m0;
    if (MODE == MODE_CONV) {
      const int hw = p.H * p.W;
      pb[i] = mm / hw;
      const int rem = mm - pb[i] * hw;
      py[i] = rem / p.W;
      px[i] = rem - py[i] * p.W;
      arow[i] = 0;
    } else {
      arow[i] = (long long)(p.a_row_mod > 0 ? mm % p.a_row_mod : mm) * p.lda;
      pb[i] = py[i] = px[i] = 0;
    }
  }

  // ---- LayerNorm statistics of the CTA's rows (two-pass, warp per row) -----------------
  float mean[2] = {0.f, 0.f}, rstd[2] = {1.f, 1.f};
  if (MODE == MODE_LN) {
    const int warp = tid >> 5, lane = tid & 31;
    const int k4n = p.Ka >> 2;
    for (int r = warp; r < BM; r += NT / 32) {
      const int m = m0 + r;
      float mu = 0.f, rs = 1.f;
      if (m < m_end) {
        const float* row = p.A + (long long)(p.a_row_mod > 0 ? m % p.a_row_mod : m) * p.lda;
        float s = 0.f;
        for (int k4 = lane; k4 < k4n; k4 += 32) {
          const float4 v = ldg4(row + 4 * k4);
          s += (v.x + v.y) + (v.z + v.w);
        }
        mu = warp_sum(s) / (float)p.Ka;
        float q = 0.f;
        for (int k4 = lane; k4 < k4n; k4 += 32) {
          const float4 v = ldg4(row + 4 * k4);
          const float a = v.x - mu, b = v.y - mu, c = v.z - mu, d = v.w - mu;
          q += (a * a + b * b) + (c * c + d * d);
        }
        rs = rsqrtf(warp_sum(q) / (float)p.Ka + 1e-5f);
      }
      if (lane == 0) stats[r] = make_float2(mu, rs);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 s = stats[(tid >> 2) + i * 64];
      mean[i] = s.x;
      rstd[i] = s.y;
    }
  }

  float4 ra[2];
  float4 rb[B_PER_THREAD];

  auto load_tile = [&](int kt) {
    const int k = kt * BK + kq * 4;
    if (MODE == MODE_CONV) {
      const int kk0 = kt * BK;
      const int tap = kk0 / p.Cin;
      const int c = kk0 - tap * p.Cin + kq * 4;
      const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int yy = py[i] + dy, xx = px[i] + dx;
        const bool ok = valid[i] && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        ra[i] = ok ? ldg4(p.A + ((long long)(pb[i] * p.H + yy) * p.W + xx) * p.lda + c)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const bool kin = k < p.Ka;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f), be = g;
      if (MODE == MODE_LN && kin) {
        g = ldg4(p.ln_g + k);
        be = ldg4(p.ln_b + k);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool ok = valid[i] && kin;
        float4 v = ok ? ldg4(p.A + arow[i] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == MODE_LN && ok) {
          v.x = (v.x - mean[i]) * rstd[i] * g.x + be.x;
          v.y = (v.y - mean[i]) * rstd[i] * g.y + be.y;
          v.z = (v.z - mean[i]) * rstd[i] * g.z + be.z;
          v.w = (v.w - mean[i]) * rstd[i] * g.w + be.w;
        }
        ra[i] = v;
      }
    }
#pragma unroll
    for (int j = 0; j < B_PER_THREAD; ++j) {
      const int idx = tid + j * NT;
      const int kk = idx / B_LD4, nq = idx - kk * B_LD4;
      rb[j] = ldg4(Bt + (long long)(kt * BK + kk) * p.ldb + n0 + nq * 4);
    }
  };

  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = (tid >> 2) + i * 64;
      As[buf][kq * 4 + 0][r] = ra[i].x;
      As[buf][kq * 4 + 1][r] = ra[i].y;
      As[buf][kq * 4 + 2][r] = ra[i].z;
      As[buf][kq * 4 + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int j = 0; j < B_PER_THREAD; ++j) {
      const int idx = tid + j * NT;
      const int kk = idx / B_LD4, nq = idx - kk * B_LD4;
      *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb[j];
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();

  for (int kt = 0; kt < p.nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < p.nk) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
        b[4 * g + 0] = bv.x;
        b[4 * g + 1] = bv.y;
        b[4 * g + 2] = bv.z;
        b[4 * g + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < p.nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue --------------------------------------------------------------------------
  const int hw = (p.H > 0) ? p.H * p.W : 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= m_end) continue;
    float scale = 1.f;
    if (p.row_scale != nullptr) scale = __ldg(p.row_scale + m / p.rows_per_batch);
    int ib = 0, iy = 0, ix = 0;
    if (p.epi == MPHSIR_EPI_SPECTRAL || p.epi == MPHSIR_EPI_PROJ || p.epi >= OUT_UNSHUFFLE) {
      ib = m / hw;
      const int rem = m - ib * hw;
      iy = rem / p.W;
      ix = rem - iy * p.W;
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int n = n0 + g * 64 + tx * 4;
      if (n >= p.N) continue;
      float4 v = make_float4(acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]);
      if (p.bias != nullptr) {
        const float4 bb = ldg4(p.bias + n);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      switch (p.epi) {
        case MPHSIR_EPI_BIAS:
          *reinterpret_cast<float4*>(p.Y + (long long)m * p.ldy + n) = v;
          break;
        case MPHSIR_EPI_RESIDUAL: {
          const float4 r1 = ldg4(p.res1 + (long long)m * p.ldr1 + n);
          float4 o = make_float4(r1.x + scale * v.x, r1.y + scale * v.y, r1.z + scale * v.z,
                                 r1.w + scale * v.w);
          if (p.res2 != nullptr) {
            const float4 r2 = ldg4(p.res2 + (long long)m * p.ldr2 + n);
            o.x += r2.x; o.y += r2.y; o.z += r2.z; o.w += r2.w;
          }
          *reinterpret_cast<float4*>(p.Y + (long long)m * p.ldy + n) = o;
          break;
        }
        case MPHSIR_EPI_GLU: {
          const float2 o = make_float2(v.x * gelu_erf(v.y), v.z * gelu_erf(v.w));
          *reinterpret_cast<float2*>(p.Y + (long long)m * p.ldy + (n >> 1)) = o;
          break;
        }
        case MPHSIR_EPI_SPECTRAL: {
          int ys = iy - p.shift, xs = ix - p.shift;
          if (ys < 0) ys += p.H;
          if (xs < 0) xs += p.W;
          const int win = ib * (hw >> 6) + (ys >> 3) * (p.W >> 3) + (xs >> 3);
          const float4 gt = ldg4(p.gate + (long long)win * p.N + n);
          const float4 sa = ldg4(p.gsrc + (long long)m * p.ldg + n);
          const float4 r1 = ldg4(p.res1 + (long long)m * p.ldr1 + n);
          float4 o;
          o.x = r1.x + scale * (sa.x * gt.x + v.x);
          o.y = r1.y + scale * (sa.y * gt.y + v.y);
          o.z = r1.z + scale * (sa.z * gt.z + v.z);
          o.w = r1.w + scale * (sa.w * gt.w + v.w);
          *reinterpret_cast<float4*>(p.Y + (long long)m * p.ldy + n) = o;
          break;
        }
        case MPHSIR_EPI_PROJ: {
          if (n >= p.n_split) {
            *reinterpret_cast<float4*>(p.Y2 + (long long)m * p.ldy2 + (n - p.n_split)) = v;
            break;
          }
          int ys = iy - p.shift, xs = ix - p.shift;
          if (ys < 0) ys += p.H;
          if (xs < 0) xs += p.W;
          const int win = ib * (hw >> 6) + (ys >> 3) * (p.W >> 3) + (xs >> 3);
          const float4 gt = ldg4(p.gate + (long long)win * p.n_split + n);
          const float4 r1 = ldg4(p.res1 + (long long)m * p.ldr1 + n);
          *reinterpret_cast<float4*>(p.Y + (long long)m * p.ldy + n) =
              make_float4(r1.x + scale * v.x * gt.x, r1.y + scale * v.y * gt.y, r1.z + scale * v.z * gt.z,
                          r1.w + scale * v.w * gt.w);
          break;
        }
        case OUT_UNSHUFFLE: {
          float* dst = p.Y + ((long long)(ib * (p.H >> 1) + (iy >> 1)) * (p.W >> 1) + (ix >> 1)) * p.ldy +
                       2 * (iy & 1) + (ix & 1);
          dst[(n + 0) * 4] = v.x;
          dst[(n + 1) * 4] = v.y;
          dst[(n + 2) * 4] = v.z;
          dst[(n + 3) * 4] = v.w;
          break;
        }
        case OUT_SHUFFLE: {
          const int cn_total = p.N >> 2;
          const int q = n / cn_total, cn = n - q * cn_total;
          float* dst = p.Y +
                       ((long long)(ib * 2 * p.H + 2 * iy + (q >> 1)) * (2 * p.W) + 2 * ix + (q & 1)) * p.ldy +
                       cn;
          *reinterpret_cast<float4*>(dst) = v;
          break;
        }
        case OUT_NCHW_RES: {
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = n + e;
            if (c < p.N) {
              const long long idx = ((long long)(ib * p.N + c) * p.H + iy) * p.W + ix;
              p.Y[idx] = vv[e] + __ldg(p.R + idx);
            }
          }
          break;
        }
        default:
          break;
      }
    }
  }
}

template <int MODE>
static int launch(const GemmArgs& a, cudaStream_t st, const char* what) {
  const int mt = a.tiles_per_batch > 0 ? (a.M / a.rows_per_batch) * a.tiles_per_batch : (a.M + BM - 1) / BM;
  const bool wide = a.N > 64 && (a.ldb % 128 == 0);
  if (wide) {
    dim3 grid(mt, (a.N + 127) / 128);
    gemm_kernel<128, MODE><<<grid, NT, 0, st>>>(a);
  } else {
    dim3 grid(mt, (a.N + 63) / 64);
    gemm_kernel<64, MODE><<<grid, NT, 0, st>>>(a);
  }
  return check_launch(what);
}

}  // namespace mphsir

using namespace mphsir;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int gemm_tc_dispatch(const mphsir_gemm_params* p, cudaStream_t st);

extern "C" int mphsir_gemm_fwd(const mphsir_gemm_params* p, void* stream) {
  MPHSIR_REQUIRE(p != nullptr, "gemm: null params");
  MPHSIR_REQUIRE(p->precision >= MPHSIR_PREC_FP32_SIMT && p->precision <= MPHSIR_PREC_BF16, "gemm: unknown precision %d", p->precision);
  if (p->precision != MPHSIR_PREC_FP32_SIMT) return gemm_tc_dispatch(p, reinterpret_cast<cudaStream_t>(stream));
  MPHSIR_REQUIRE(p->A && p->Bt && p->Y, "gemm: null operand");
  MPHSIR_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, "gemm: bad shape M=%d N=%d K=%d", p->M, p->N, p->K);
  MPHSIR_REQUIRE(p->K % 4 == 0 && p->lda % 4 == 0 && p->lda >= p->K, "gemm: K=%d lda=%d must be multiples of 4, lda>=K", p->K, p->lda);
  MPHSIR_REQUIRE(p->N % 4 == 0 && p->ldy % 4 == 0, "gemm: N=%d ldy=%d must be multiples of 4", p->N, p->ldy);
  MPHSIR_REQUIRE(p->ldb % 64 == 0 && p->ldb >= ((p->N + 63) / 64) * 64, "gemm: ldb=%d must be a multiple of 64 covering N=%d", p->ldb, p->N);
  MPHSIR_REQUIRE(aligned16(p->A) && aligned16(p->Bt) && aligned16(p->Y), "gemm: operands must be 16-byte aligned");
  MPHSIR_REQUIRE((p->ln_gamma == nullptr) == (p->ln_beta == nullptr), "gemm: ln_gamma/ln_beta must both be set");
  MPHSIR_REQUIRE(p->epi >= MPHSIR_EPI_BIAS && p->epi <= MPHSIR_EPI_PROJ, "gemm: unknown epilogue %d", p->epi);
  const bool per_sample = p->b_batch_stride != 0 || p->row_scale != nullptr || p->epi == MPHSIR_EPI_SPECTRAL || p->epi == MPHSIR_EPI_PROJ;
  if (per_sample)
    MPHSIR_REQUIRE(p->rows_per_batch > 0 && p->M % p->rows_per_batch == 0, "gemm: rows_per_batch=%d must divide M=%d", p->rows_per_batch, p->M);
  if (p->epi == MPHSIR_EPI_RESIDUAL || p->epi == MPHSIR_EPI_SPECTRAL || p->epi == MPHSIR_EPI_PROJ)
    MPHSIR_REQUIRE(p->res1 != nullptr && p->ldr1 % 4 == 0 && aligned16(p->res1), "gemm: residual epilogue needs aligned res1");
  if (p->res2) MPHSIR_REQUIRE(p->ldr2 % 4 == 0 && aligned16(p->res2), "gemm: res2 misaligned");
  if (p->epi == MPHSIR_EPI_PROJ) {
    MPHSIR_REQUIRE(p->gate && p->Y2 && aligned16(p->Y2) && p->ldy2 % 4 == 0, "gemm: proj epilogue needs gate and an aligned Y2");
    MPHSIR_REQUIRE(p->n_split > 0 && p->n_split % 32 == 0 && p->n_split < p->N, "gemm: proj epilogue needs 0 < n_split=%d < N, a multiple of 32", p->n_split);
  }
  if (p->epi == MPHSIR_EPI_SPECTRAL || p->epi == MPHSIR_EPI_PROJ) {
    MPHSIR_REQUIRE(p->epi == MPHSIR_EPI_PROJ || (p->gsrc && p->gate && p->ldg % 4 == 0), "gemm: spectral epilogue needs gsrc/gate");
    MPHSIR_REQUIRE(p->H > 0 && p->W > 0 && p->H % 8 == 0 && p->W % 8 == 0 && p->H * p->W == p->rows_per_batch, "gemm: spectral epilogue needs H,W multiples of 8 with H*W == rows_per_batch");
    MPHSIR_REQUIRE(p->shift == 0 || p->shift == 4, "gemm: shift must be 0 or 4");
  }
  if (p->epi == MPHSIR_EPI_GLU) MPHSIR_REQUIRE(p->ldy % 2 == 0, "gemm: GLU output ld must be even");

  GemmArgs a{};
  a.A = p->A; a.lda = p->lda; a.a_row_mod = p->a_row_mod; a.Ka = p->K;
  a.Bt = p->Bt; a.ldb = p->ldb; a.b_batch_stride = p->b_batch_stride; a.rows_per_batch = p->rows_per_batch;
  a.tiles_per_batch = p->b_batch_stride != 0 ? (p->rows_per_batch + BM - 1) / BM : 0;
  a.Y = p->Y; a.ldy = p->ldy; a.M = p->M; a.N = p->N; a.nk = (p->K + BK - 1) / BK;
  a.ln_g = p->ln_gamma; a.ln_b = p->ln_beta; a.bias = p->bias; a.epi = p->epi;
  a.res1 = p->res1; a.ldr1 = p->ldr1; a.res2 = p->res2; a.ldr2 = p->ldr2;
  a.gsrc = p->gsrc; a.ldg = p->ldg; a.gate = p->gate; a.H = p->H; a.W = p->W; a.shift = p->shift;
  a.row_scale = p->row_scale; a.Y2 = p->Y2; a.ldy2 = p->ldy2; a.n_split = p->n_split;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return p->ln_gamma ? launch<MODE_LN>(a, st, "gemm(ln)") : launch<MODE_PLAIN>(a, st, "gemm");
}

static int gemm_tc_dispatch(const mphsir_gemm_params* p, cudaStream_t st) {
  MPHSIR_REQUIRE(p->A && p->Bimg && p->Y, "gemm(tc): null operand (Bimg is required for tensor-core precisions)");
  MPHSIR_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, "gemm(tc): bad shape M=%d N=%d K=%d", p->M, p->N, p->K);
  MPHSIR_REQUIRE(p->K % 8 == 0 && p->lda % 4 == 0 && p->lda >= p->K, "gemm(tc): K=%d must be a multiple of 8, lda=%d a multiple of 4", p->K, p->lda);
  MPHSIR_REQUIRE(p->N % 4 == 0 && p->ldy % 4 == 0, "gemm(tc): N=%d ldy=%d must be multiples of 4", p->N, p->ldy);
  MPHSIR_REQUIRE(aligned16(p->A) && aligned16(p->Y) && (reinterpret_cast<uintptr_t>(p->Bimg) & 127) == 0, "gemm(tc): operands misaligned");
  MPHSIR_REQUIRE((p->ln_gamma == nullptr) == (p->ln_beta == nullptr), "gemm(tc): ln_gamma/ln_beta must both be set");
  MPHSIR_REQUIRE(p->epi >= MPHSIR_EPI_BIAS && p->epi <= MPHSIR_EPI_PROJ, "gemm(tc): unknown epilogue %d", p->epi);
  const bool per_sample = p->bimg_batch_bytes != 0 || p->row_scale != nullptr || p->epi == MPHSIR_EPI_SPECTRAL || p->epi == MPHSIR_EPI_PROJ;
  if (per_sample)
    MPHSIR_REQUIRE(p->rows_per_batch > 0 && p->M % p->rows_per_batch == 0, "gemm(tc): rows_per_batch=%d must divide M=%d", p->rows_per_batch, p->M);
  if (p->epi == MPHSIR_EPI_RESIDUAL || p->epi == MPHSIR_EPI_SPECTRAL || p->epi == MPHSIR_EPI_PROJ)
    MPHSIR_REQUIRE(p->res1 != nullptr && p->ldr1 % 4 == 0 && aligned16(p->res1), "gemm(tc): residual epilogue needs aligned res1");
  if (p->res2) MPHSIR_REQUIRE(p->ldr2 % 4 == 0 && aligned16(p->res2), "gemm(tc): res2 misaligned");
  if (p->epi == MPHSIR_EPI_PROJ) {
    MPHSIR_REQUIRE(p->gate && p->Y2 && aligned16(p->Y2) && p->ldy2 % 4 == 0, "gemm(tc): proj epilogue needs gate and an aligned Y2");
    MPHSIR_REQUIRE(p->n_split > 0 && p->n_split % 32 == 0 && p->n_split < p->N, "gemm(tc): proj epilogue needs 0 < n_split=%d < N, a multiple of 32", p->n_split);
  }
  if (p->epi == MPHSIR_EPI_SPECTRAL || p->epi == MPHSIR_EPI_PROJ) {
    MPHSIR_REQUIRE(p->epi == MPHSIR_EPI_PROJ || (p->gsrc && p->gate && p->ldg % 4 == 0), "gemm(tc): spectral epilogue needs gsrc/gate");
    MPHSIR_REQUIRE(p->H > 0 && p->W > 0 && p->H % 8 == 0 && p->W % 8 == 0 && p->H * p->W == p->rows_per_batch, "gemm(tc): spectral epilogue needs H,W multiples of 8 with H*W == rows_per_batch");
    MPHSIR_REQUIRE(p->shift == 0 || p->shift == 4, "gemm(tc): shift must be 0 or 4");
  }
  tc::TcArgs a{};
  a.A = p->A; a.lda = p->lda; a.a_row_mod = p->a_row_mod; a.Ka = p->K;
  a.Bimg = p->Bimg; a.b_batch_bytes = p->bimg_batch_bytes; a.rows_per_batch = p->rows_per_batch;
  a.tiles_per_batch = p->bimg_batch_bytes != 0 ? (p->rows_per_batch + 127) / 128 : 0;
  a.num_tiles = a.tiles_per_batch > 0 ? (p->M / p->rows_per_batch) * a.tiles_per_batch : (p->M + 127) / 128;
  a.Y = p->Y; a.ldy = p->ldy; a.M = p->M; a.N = p->N; a.Np = (p->N + 15) / 16 * 16; a.ks = (p->K + 63) / 64;
  a.parts = p->precision == MPHSIR_PREC_BF16X3 ? 2 : 1;
  a.ln_g = p->ln_gamma; a.ln_b = p->ln_beta; a.bias = p->bias; a.epi = p->epi;
  a.res1 = p->res1; a.ldr1 = p->ldr1; a.res2 = p->res2; a.ldr2 = p->ldr2;
  a.gsrc = p->gsrc; a.ldg = p->ldg; a.gate = p->gate; a.H = p->H; a.W = p->W; a.shift = p->shift;
  a.row_scale = p->row_scale; a.Y2 = p->Y2; a.ldy2 = p->ldy2; a.n_split = p->n_split;
  return tc::launch_gemm_tc(a, false, st);
}

static int conv_tc_dispatch(const mphsir_conv3x3_params* p, cudaStream_t st) {
  MPHSIR_REQUIRE(p->X && p->Bimg && p->Y, "conv3x3(tc): null operand (Bimg is required for tensor-core precisions)");
  MPHSIR_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0, "conv3x3(tc): bad image");
  MPHSIR_REQUIRE(p->Cin % 8 == 0 && p->ldx >= p->Cin && p->ldx % 4 == 0, "conv3x3(tc): Cin=%d must be a multiple of 8, ldx=%d", p->Cin, p->ldx);
  MPHSIR_REQUIRE(aligned16(p->X) && (reinterpret_cast<uintptr_t>(p->Bimg) & 127) == 0, "conv3x3(tc): operands misaligned");
  tc::TcArgs a{};
  a.A = p->X; a.lda = p->ldx; a.Ka = 9 * p->Cin; a.Cin = p->Cin;
  a.Bimg = p->Bimg;
  a.M = p->B * p->H * p->W; a.num_tiles = (a.M + 127) / 128;
  a.Y = p->Y; a.ldy = p->ldy; a.N = p->N; a.Np = (p->N + 15) / 16 * 16; a.ks = (9 * p->Cin + 63) / 64;
  a.parts = p->precision == MPHSIR_PREC_BF16X3 ? 2 : 1;
  a.H = p->H; a.W = p->W; a.rows_per_batch = p->H * p->W;
  switch (p->out_mode) {
    case MPHSIR_CONV_TOKENS:
      MPHSIR_REQUIRE(p->N % 4 == 0 && p->ldy % 4 == 0 && aligned16(p->Y), "conv3x3(tc): token output needs N,ldy multiples of 4");
      a.epi = MPHSIR_EPI_BIAS;
      break;
    case MPHSIR_CONV_UNSHUFFLE:
      MPHSIR_REQUIRE(p->N % 4 == 0 && p->H % 2 == 0 && p->W % 2 == 0, "conv3x3(tc): unshuffle needs even H,W");
      a.epi = tc::TC_OUT_UNSHUFFLE;
      break;
    case MPHSIR_CONV_SHUFFLE:
      MPHSIR_REQUIRE(p->N % 16 == 0 && p->ldy % 4 == 0 && aligned16(p->Y), "conv3x3(tc): shuffle needs N multiple of 16");
      a.epi = tc::TC_OUT_SHUFFLE;
      break;
    case MPHSIR_CONV_NCHW_RES:
      MPHSIR_REQUIRE(p->R != nullptr, "conv3x3(tc): NCHW residual output needs R");
      a.epi = tc::TC_OUT_NCHW_RES;
      a.R = p->R;
      break;
    default:
      MPHSIR_REQUIRE(false, "conv3x3(tc): unknown out_mode %d", p->out_mode);
  }
  return tc::launch_gemm_tc(a, true, st);
}

extern "C" int mphsir_conv3x3_fwd(const mphsir_conv3x3_params* p, void* stream) {
  MPHSIR_REQUIRE(p != nullptr, "conv3x3: null params");
  MPHSIR_REQUIRE(p->precision >= MPHSIR_PREC_FP32_SIMT && p->precision <= MPHSIR_PREC_BF16, "conv3x3: unknown precision %d", p->precision);
  if (p->precision != MPHSIR_PREC_FP32_SIMT) return conv_tc_dispatch(p, reinterpret_cast<cudaStream_t>(stream));
  MPHSIR_REQUIRE(p->X && p->Wt && p->Y, "conv3x3: null operand");
  MPHSIR_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0 && p->W % 8 == 0, "conv3x3: bad image %dx%dx%d (W must be a multiple of 8)", p->B, p->H, p->W);
  MPHSIR_REQUIRE(p->Cin % 16 == 0 && p->ldx >= p->Cin && p->ldx % 4 == 0, "conv3x3: Cin=%d must be a multiple of 16 (zero-pad the input), ldx=%d", p->Cin, p->ldx);
  MPHSIR_REQUIRE(p->ldb % 64 == 0 && p->ldb >= ((p->N + 63) / 64) * 64, "conv3x3: ldb=%d must be a multiple of 64 covering N=%d", p->ldb, p->N);
  MPHSIR_REQUIRE(aligned16(p->X) && aligned16(p->Wt), "conv3x3: operands must be 16-byte aligned");
  GemmArgs a{};
  a.A = p->X; a.lda = p->ldx; a.Ka = 9 * p->Cin; a.Cin = p->Cin;
  a.Bt = p->Wt; a.ldb = p->ldb;
  a.Y = p->Y; a.ldy = p->ldy; a.M = p->B * p->H * p->W; a.N = p->N; a.nk = 9 * p->Cin / BK;
  a.H = p->H; a.W = p->W; a.rows_per_batch = p->H * p->W;
  switch (p->out_mode) {
    case MPHSIR_CONV_TOKENS:
      MPHSIR_REQUIRE(p->N % 4 == 0 && p->ldy % 4 == 0 && aligned16(p->Y), "conv3x3: token output needs N,ldy multiples of 4");
      a.epi = MPHSIR_EPI_BIAS;
      break;
    case MPHSIR_CONV_UNSHUFFLE:
      MPHSIR_REQUIRE(p->N % 4 == 0 && p->H % 2 == 0 && p->W % 2 == 0, "conv3x3: unshuffle needs even H,W");
      a.epi = OUT_UNSHUFFLE;
      break;
    case MPHSIR_CONV_SHUFFLE:
      MPHSIR_REQUIRE(p->N % 16 == 0 && p->ldy % 4 == 0 && aligned16(p->Y), "conv3x3: shuffle needs N multiple of 16");
      a.epi = OUT_SHUFFLE;
      break;
    case MPHSIR_CONV_NCHW_RES:
      MPHSIR_REQUIRE(p->R != nullptr, "conv3x3: NCHW residual output needs R");
      a.epi = OUT_NCHW_RES;
      a.R = p->R;
      break;
    default:
      MPHSIR_REQUIRE(false, "conv3x3: unknown out_mode %d", p->out_mode);
  }
  return launch<MODE_CONV>(a, reinterpret_cast<cudaStream_t>(stream), "conv3x3");
}
