// Weight gradients on the tensor cores:  dW[o, i] += sum_m dY[m, o] * X[src(m), i]
//
// The contraction runs over tokens, i.e. over the *slow* axis of both token-major operands, so neither operand
// is K-major: tiles of 64 tokens are converted to bf16 (hi, or hi+lo for the fp32-grade mode) into shared
// memory as [token][channel] and read by ldmatrix.trans as the A (dY, 128 channels) and B (X, 64 channels)
// fragments of mma.sync.m16n8k16 — the same scheme as the forward Gram kernel (dwgram.cu).  Each CTA owns one
// 128x64 tile of dW and a strided subset of the token chunks; partial sums leave through fp32 atomics into
// the (pre-zeroed) gradient buffer, addressed through the output map so that packed / padded / interleaved
// activation layouts land directly in the reference parameter layout.
//
// Modes: plain; per-sample (grid.z = sample: the d(out)^T v products of the spectral attention backward);
// conv taps (grid.z = tap: X rows shifted by (dy,dx) with zero padding = dense 3x3 conv weight gradient);
// x_row_mod (X shared by all samples: TVSP visual prompt).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mphsir {
namespace wg {

constexpr int TO = 128, TI = 64, KT = 64;
constexpr int LDA = TO + 8, LDB = TI + 8;

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

__device__ __forceinline__ int map_index(int o, int mode, int a, int b) {
  if (mode == MPHSIR_MAP_IDENTITY) return o < a ? o : -1;
  if (mode == MPHSIR_MAP_INTERLEAVE) {
    const int j = o >> 1;
    return j < a ? (o & 1) * a + j : -1;
  }
  if (o < b) return o < a ? o : -1;
  return (o - b) < a ? a + (o - b) : -1;
}

struct Args {
  const float* dY;
  long long lddy;
  const float* X;
  long long ldx;
  float* dW;
  long long M;
  int O, I;
  int rows_per_batch;  // > 0: grid.z = sample
  long long dw_batch_stride;
  int x_row_mod;
  int H, W, taps;  // taps == 9: grid.z = tap
  long long so, si, st;
  int map_mode, map_a, map_b;
  int i_valid;
  int tiles_i;
  float* dbias;
};

// bx: dW tile, by / gy: token split index / count, bz: sample or conv tap
template <int PARTS>
__device__ __forceinline__ void wgrad_body(const Args& p, const int bx, const int by, const int bz, const int gy, uint8_t* smem_raw) {
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem_raw);  // [PARTS][KT][LDA]
  __nv_bfloat16* Bs = As + PARTS * KT * LDA;                       // [PARTS][KT][LDB]
  constexpr int A_ARR = KT * LDA, B_ARR = KT * LDB;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile_o = bx / p.tiles_i, tile_i = bx - tile_o * p.tiles_i;
  const int o0 = tile_o * TO, i0 = tile_i * TI;

  long long m_begin = 0, m_end = p.M;
  int tap = 0, dy = 0, dx = 0;
  float* dW = p.dW;
  if (p.taps == 9) {
    tap = bz;
    dy = tap / 3 - 1;
    dx = tap - (tap / 3) * 3 - 1;
  } else if (p.rows_per_batch > 0) {
    m_begin = (long long)bz * p.rows_per_batch;
    m_end = m_begin + p.rows_per_batch;
    dW += (long long)bz * p.dw_batch_stride;
  }
  const long long n_chunks = (m_end - m_begin + KT - 1) / KT;

  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;

  float4 ra[8], rb[4];
  // per-thread constants of the chunk loads: A unit (row ta + 8r, quad oq), B unit (row tb + 16r, quad iq)
  const int ta = tid >> 5, oq = tid & 31, tb = tid >> 4, iq = tid & 15;
  const bool a_col_ok = o0 + 4 * oq < p.O, b_col_ok = i0 + 4 * iq < p.I;
  const float* a_col = p.dY + o0 + 4 * oq;
  const float* b_col = p.X + i0 + 4 * iq;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_chunk = [&](long long chunk) {
    const long long mc = m_begin + chunk * KT;
    const float* ap = a_col + (mc + ta) * p.lddy;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      ra[r] = (a_col_ok && mc + ta + 8 * r < m_end) ? ldg4(ap + (long long)(8 * r) * p.lddy) : zero4;
    if (p.taps == 9 || p.x_row_mod > 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        long long m = mc + tb + 16 * r;
        bool ok = b_col_ok && m < m_end;
        if (ok && p.taps == 9) {
          const int x = (int)(m % p.W), y = (int)((m / p.W) % p.H);
          ok = (unsigned)(y + dy) < (unsigned)p.H && (unsigned)(x + dx) < (unsigned)p.W;
          m += (long long)dy * p.W + dx;
        }
        if (ok && p.x_row_mod > 0) m %= p.x_row_mod;
        rb[r] = ok ? ldg4(b_col + m * p.ldx) : zero4;
      }
    } else {
      const float* bp = b_col + (mc + tb) * p.ldx;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        rb[r] = (b_col_ok && mc + tb + 16 * r < m_end) ? ldg4(bp + (long long)(16 * r) * p.ldx) : zero4;
    }
  };
  const bool want_bias = p.dbias != nullptr && tile_i == 0 && bz == 0;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  auto store_chunk = [&]() {
    if (want_bias) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        bsum.x += ra[r].x;
        bsum.y += ra[r].y;
        bsum.z += ra[r].z;
        bsum.w += ra[r].w;
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      uint2 hi, lo;
      split4(ra[r], hi, lo);
      *reinterpret_cast<uint2*>(As + (ta + 8 * r) * LDA + 4 * oq) = hi;
      if (PARTS == 2) *reinterpret_cast<uint2*>(As + A_ARR + (ta + 8 * r) * LDA + 4 * oq) = lo;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint2 hi, lo;
      split4(rb[r], hi, lo);
      *reinterpret_cast<uint2*>(Bs + (tb + 16 * r) * LDB + 4 * iq) = hi;
      if (PARTS == 2) *reinterpret_cast<uint2*>(Bs + B_ARR + (tb + 16 * r) * LDB + 4 * iq) = lo;
    }
  };

  const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(As);
  const uint32_t b_base = (uint32_t)__cvta_generic_to_shared(Bs);
  const int m0 = warp * 16;

  long long chunk = by;
  if (chunk < n_chunks) load_chunk(chunk);
  for (; chunk < n_chunks; chunk += gy) {
    __syncthreads();  // the previous chunk's fragments have been consumed
    store_chunk();
    __syncthreads();
    if (chunk + gy < n_chunks) load_chunk(chunk + gy);
#pragma unroll
    for (int ks = 0; ks < KT / 16; ++ks) {
      // every fragment of this k-step first (ldmatrix latency overlaps), then 8 (24) independent MMAs
      uint32_t ah[4], al[4], bh[4][4], bl[4][4];
      const uint32_t a_off = (uint32_t)(((16 * ks + (lane & 7) + ((lane >> 4) & 1) * 8) * LDA + m0 + ((lane >> 3) & 1) * 8) * 2);
      const uint32_t b_row = (uint32_t)((16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8) * LDB + 8 * (lane >> 4));
      ldsm_x4_trans(a_base + a_off, ah);
#pragma unroll
      for (int np = 0; np < 4; ++np) ldsm_x4_trans(b_base + (b_row + 16 * np) * 2, bh[np]);
      if (PARTS == 2) {
        ldsm_x4_trans(a_base + A_ARR * 2 + a_off, al);
#pragma unroll
        for (int np = 0; np < 4; ++np) ldsm_x4_trans(b_base + B_ARR * 2 + (b_row + 16 * np) * 2, bl[np]);
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        mma_bf16(acc[2 * np], ah, bh[np][0], bh[np][1]);
        mma_bf16(acc[2 * np + 1], ah, bh[np][2], bh[np][3]);
      }
      if (PARTS == 2) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          mma_bf16(acc[2 * np], ah, bl[np][0], bl[np][1]);
          mma_bf16(acc[2 * np + 1], ah, bl[np][2], bl[np][3]);
          mma_bf16(acc[2 * np], al, bh[np][0], bh[np][1]);
          mma_bf16(acc[2 * np + 1], al, bh[np][2], bh[np][3]);
        }
      }
    }
  }

  // ---- bias gradient: column sums of this CTA's dY chunks (rows ta + 8r of every chunk, 4 columns per thread) ----
  if (want_bias) {
    __syncthreads();  // the last chunk's fragments have been read: the A tile area is free
    float* red = reinterpret_cast<float*>(smem_raw);  // [8][128]
    *reinterpret_cast<float4*>(red + ta * 128 + 4 * oq) = bsum;
    __syncthreads();
    if (tid < 128 && o0 + tid < p.O) {
      float sacc = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) sacc += red[r * 128 + tid];
      const int ro = map_index(o0 + tid, p.map_mode, p.map_a, p.map_b);
      if (ro >= 0) atomicAdd(p.dbias + ro, sacc);
    }
  }

  // ---- scatter-add the tile through the output map --------------------------------------------
  const int g = lane >> 2, qd = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int o = o0 + m0 + g + 8 * half;
    if (o >= p.O) continue;
    const int ro = map_index(o, p.map_mode, p.map_a, p.map_b);
    if (ro < 0) continue;
    float* row = dW + (long long)ro * p.so + (long long)tap * p.st;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int i = i0 + 8 * nt + 2 * qd;
      if (i < p.i_valid) atomicAdd(row + (long long)i * p.si, acc[nt][2 * half]);
      if (i + 1 < p.i_valid) atomicAdd(row + (long long)(i + 1) * p.si, acc[nt][2 * half + 1]);
    }
  }
}

template <int PARTS>
__global__ void __launch_bounds__(256, 2) wgrad_kernel(const Args p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  wgrad_body<PARTS>(p, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, smem_raw);
}

// several small plain-mode problems in one launch (grid.z = problem): the r-sized matrices of the local spectral gate
constexpr int MAX_MULTI = 8;
struct ArgsList {
  Args a[MAX_MULTI];
  int tiles[MAX_MULTI], splits[MAX_MULTI];
};

template <int PARTS>
__global__ void __launch_bounds__(256, 2) wgrad_multi_kernel(const ArgsList l) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int z = blockIdx.z;
  if ((int)blockIdx.x >= l.tiles[z] || (int)blockIdx.y >= l.splits[z]) return;
  wgrad_body<PARTS>(l.a[z], blockIdx.x, blockIdx.y, 0, l.splits[z], smem_raw);
}

static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

template <int PARTS>
static int launch(const Args& a, int z, cudaStream_t st) {
  const size_t smem = (size_t)PARTS * KT * (LDA + LDB) * 2;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("wgrad: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  const int tiles = ((a.O + TO - 1) / TO) * a.tiles_i;
  const long long rows = a.rows_per_batch > 0 && a.taps != 9 ? a.rows_per_batch : a.M;
  const long long n_chunks = (rows + KT - 1) / KT;
  long long splits = (2LL * sm_count()) / ((long long)tiles * z);  // floor: one full wave of 2 CTAs per SM, no tail wave
  if (splits > (n_chunks + 3) / 4) splits = (n_chunks + 3) / 4;  // >= 4 chunks per CTA: bounds the atomic traffic
  if (splits < 1) splits = 1;
  // blockIdx.x = tile (fastest): CTAs resident together walk the same token chunks, so each operand tile is read
  // from HBM once and re-used by the other tiles through L2
  dim3 grid((unsigned)tiles, (unsigned)splits, (unsigned)z);
  wgrad_kernel<PARTS><<<grid, 256, smem, st>>>(a);
  return check_launch("wgrad");
}

}  // namespace wg
namespace wgt {
int launch_wgrad_tc(const mphsir_wgrad_params* p, int z, cudaStream_t st);
}
static int g_wgrad_tc = -1;  // -1: per-shape choice (default), 0: always mma.sync, 1: always tcgen05
}  // namespace mphsir

using namespace mphsir;

extern "C" void mphsir_debug_wgrad_tc(int enabled) { g_wgrad_tc = enabled; }

static int fill_args(const mphsir_wgrad_params* p, wg::Args& a);

extern "C" int mphsir_wgrad_multi(const mphsir_wgrad_params* list, int count, void* stream) {
  MPHSIR_REQUIRE(list && count > 0 && count <= wg::MAX_MULTI, "wgrad_multi: 1..%d problems", wg::MAX_MULTI);
  wg::ArgsList l;
  int max_tiles = 0, max_splits = 0;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int i = 0; i < count; ++i) {
    MPHSIR_REQUIRE(list[i].taps == 0 && list[i].rows_per_batch == 0 && list[i].x_row_mod == 0, "wgrad_multi: plain-mode problems only");
    MPHSIR_REQUIRE(list[i].precision == list[0].precision, "wgrad_multi: one precision per launch");
    const int rc = fill_args(&list[i], l.a[i]);
    if (rc != MPHSIR_OK) return rc;
    l.tiles[i] = ((l.a[i].O + wg::TO - 1) / wg::TO) * l.a[i].tiles_i;
    const long long n_chunks = (l.a[i].M + wg::KT - 1) / wg::KT;
    long long splits = (2LL * sms) / ((long long)l.tiles[i] * count);
    if (splits > (n_chunks + 3) / 4) splits = (n_chunks + 3) / 4;
    if (splits < 1) splits = 1;
    l.splits[i] = (int)splits;
    max_tiles = l.tiles[i] > max_tiles ? l.tiles[i] : max_tiles;
    max_splits = l.splits[i] > max_splits ? l.splits[i] : max_splits;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(max_tiles, max_splits, count);
  if (list[0].precision == MPHSIR_PREC_BF16X3) {
    const size_t smem = (size_t)2 * wg::KT * (wg::LDA + wg::LDB) * 2;
    static bool configured = false;
    if (!configured) {
      if (cudaFuncSetAttribute(wg::wgrad_multi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        set_error("wgrad_multi: cudaFuncSetAttribute failed");
        return MPHSIR_ERR_CUDA;
      }
      configured = true;
    }
    wg::wgrad_multi_kernel<2><<<grid, 256, smem, st>>>(l);
  } else {
    const size_t smem = (size_t)wg::KT * (wg::LDA + wg::LDB) * 2;
    wg::wgrad_multi_kernel<1><<<grid, 256, smem, st>>>(l);
  }
  return check_launch("wgrad_multi");
}

static int fill_args(const mphsir_wgrad_params* p, wg::Args& a) {
  MPHSIR_REQUIRE(p && p->dY && p->X && p->dW, "wgrad: null operand");
  MPHSIR_REQUIRE(p->M > 0 && p->O > 0 && p->I > 0 && p->O % 4 == 0 && p->I % 4 == 0, "wgrad: O and I must be positive multiples of 4");
  MPHSIR_REQUIRE(p->lddy % 4 == 0 && p->ldx % 4 == 0 && p->lddy >= p->O && p->ldx >= p->I, "wgrad: leading dimensions must be multiples of 4");
  MPHSIR_REQUIRE(((reinterpret_cast<uintptr_t>(p->dY) | reinterpret_cast<uintptr_t>(p->X)) & 15) == 0, "wgrad: operands must be 16-byte aligned");
  MPHSIR_REQUIRE(p->precision == MPHSIR_PREC_BF16X3 || p->precision == MPHSIR_PREC_BF16, "wgrad: precision must be a tensor-core mode");
  MPHSIR_REQUIRE(p->taps == 0 || p->taps == 9, "wgrad: taps must be 0 or 9");
  a.dY = p->dY;
  a.lddy = p->lddy;
  a.X = p->X;
  a.ldx = p->ldx;
  a.dW = p->dW;
  a.M = p->M;
  a.O = p->O;
  a.I = p->I;
  a.rows_per_batch = p->rows_per_batch;
  a.dw_batch_stride = p->dw_batch_stride;
  a.x_row_mod = p->x_row_mod;
  a.H = p->H;
  a.W = p->W;
  a.taps = p->taps;
  a.so = p->so;
  a.si = p->si;
  a.st = p->st;
  a.map_mode = p->map_mode;
  a.map_a = p->map_a;
  a.map_b = p->map_b;
  a.i_valid = p->i_valid > 0 ? p->i_valid : p->I;
  a.tiles_i = (p->I + wg::TI - 1) / wg::TI;
  a.dbias = p->dbias;
  MPHSIR_REQUIRE(p->dbias == nullptr || (p->taps == 0 && p->rows_per_batch == 0), "wgrad: dbias is available in plain mode only");
  return MPHSIR_OK;
}

extern "C" int mphsir_wgrad(const mphsir_wgrad_params* p, void* stream) {
  wg::Args a;
  {
    const int rc = fill_args(p, a);
    if (rc != MPHSIR_OK) return rc;
  }
  int z = 1;
  if (p->taps == 9) {
    MPHSIR_REQUIRE(p->H > 0 && p->W > 0 && p->M % ((long long)p->H * p->W) == 0, "wgrad: conv mode needs H, W with M = B*H*W");
    MPHSIR_REQUIRE(p->x_row_mod == 0, "wgrad: conv mode does not take a shared X");
    z = 9;
  } else if (p->rows_per_batch > 0) {
    MPHSIR_REQUIRE(p->M % p->rows_per_batch == 0, "wgrad: M must be a multiple of rows_per_batch");
    z = (int)(p->M / p->rows_per_batch);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // Engine choice (tools/wgrad_bench.py on B200): the tcgen05 kernel (128x128 tile, transposing converters off the MMA
  // path, coalesced scatter-add epilogue) wins on the nine-tap conv gradients (2-3x) and whenever both channel counts
  // fill its tile (C >= 128 stages: 1.2-1.5x); on narrow layers (64-wide level-1 blocks, the r-sized local-gate
  // matrices) the mma.sync kernel (128x64 tile, no TMEM / barrier setup per CTA) is as fast or faster.
  bool use_tc = g_wgrad_tc == 1;
  if (g_wgrad_tc < 0) use_tc = p->taps == 9 || (p->O >= 128 && p->I >= 128 && p->rows_per_batch == 0);
  if (use_tc && p->M < (1LL << 31)) return wgt::launch_wgrad_tc(p, z, st);
  return p->precision == MPHSIR_PREC_BF16X3 ? wg::launch<2>(a, z, st) : wg::launch<1>(a, z, st);
}
