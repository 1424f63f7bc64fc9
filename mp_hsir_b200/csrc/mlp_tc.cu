// Fused gated MLP on tcgen05 (sm_100a):   Y = X + s * ( fc2( value * gelu(gate) ) + b2 ) [+ R2],
//   [value|gate] = LN(X) W1^T + b1                       (PGSSTB.forward :719 + GatedMlp.forward :76-82)
//
// Neither the 2*hidden intermediate nor the normalised input ever exists in shared or global memory: both A operands
// live in TENSOR MEMORY.  Per 128-row tile the hidden dimension is walked in chunks of 64 units (= 128 interleaved
// (value, gate) fc1 columns = one 64-k slab of fc2):
//     CVT  : 4 converter warps (lane = row) read the fp32 rows of X from the TMA landing slots (SWIZZLE_128B boxes of
//            [128 rows x 32 floats]: conflict-free row-per-lane reads), LayerNorm them, split to bf16 hi/lo and write the
//            K-major A image of fc1 with tcgen05.st (two bf16 per 32-bit column, 8 columns per k-step of 16)
//     MMA  : acc1[j&1] (TMEM, 128 cols)  = LN(X) . W1_j^T     (A from TMEM, B = weight slabs in shared memory)
//     GLU  : 8 warps drain acc1 with tcgen05.ld, add b1, value*gelu(gate), split to bf16 hi/lo and write the [128 x 64]
//            tile H_j with tcgen05.st OVER THE ACCUMULATOR COLUMNS IT CAME FROM (every warp overwrites only columns it
//            has just read; the tensor pipe executes in issue order, so fc1 of chunk j+2 cannot overtake fc2 of chunk j)
//     MMA  : acc2 (TMEM, C cols)        += H_j . W2_j^T        (A from TMEM)
//     EPI  : 4 warps drain acc2 (+ b2, residuals, DropPath scale) through a shared-memory transpose to 128-byte rows
// with fc1 of chunk j+1 issued before fc2 of chunk j.  Every role has its own warps and runs a tile ahead of its
// consumer: the landing slots are free again as soon as the converters have read them (the next tile's X travels during
// this tile's MMAs), the final epilogue of tile t overlaps the first chunks of tile t+1.
// What the measurements behind this layout say (tools/mma_rate.cu, tools/mlp_bench.py, B200):
//   * one tcgen05.mma M=128 N=128 K=16 costs 121 clk with both operands in shared memory and 100 clk with A in tensor
//     memory (N=256: 171 / 138) — the instruction rate, not shared-memory bandwidth, bounds a cta_group::1 kernel;
//   * the previous layout (converters and epilogue warps doubling as GLU warps, one tile of X in flight) ran 207 us of its
//     283 us per launch at M = 262144 with the MMAs switched off: role serialisation, not the tensor pipe, was the floor.
// TMEM columns: acc1[0] 0-127 | acc1[1] 128-255 | acc2 256-383 | X hi 384-447 | X lo 448-511.
// HBM traffic per token: read C (+C residual, an L2 hit), write C floats — the unfused pair moved 2*(C + hidden) more.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace mphsir {
namespace tc {

constexpr int kMlpThreads = 608;  // warp 0: B loader, 1: MMA, 2-9: GLU, 10-13: converters, 14-17: final epilogue, 18: X loader
constexpr int M_SLAB = 128 * 128;        // one bf16 part of a 128-row x 64-k weight slab
constexpr int M_LAND = 128 * 64 * 4;     // fp32 landing slot of one 64-k slab of X: two swizzled [128 x 32] boxes
constexpr int M_NL = 2;                  // landing slots (one tile at C = 128, two at C = 64)
constexpr int M_STG_FLOATS = 32 * 32;
constexpr uint32_t ACC2_COL = 256, XH_COL = 384, XL_COL = 448;

struct MlpArgs {
  alignas(64) CUtensorMap tmX;   // boxes [128 rows x 32 floats], SWIZZLE_128B: landing slots and the residual tile
  alignas(64) CUtensorMap tmY;   // boxes [32 rows x 32 floats], SWIZZLE_128B: the output leaves from the residual tile
  const float* X;       // [M, ldx] residual stream (also the GEMM operand behind tmX)
  long long ldx;
  const float* ln_g;
  const float* ln_b;
  const void* W1img;    // image of the interleaved fc1 weights, logical [N1, C]
  const float* b1;      // [N1] interleaved
  const void* W2img;    // image of fc2, logical [C, hid_pad]
  const float* b2;      // [C]
  const float* res2;    // optional second residual
  long long ldr2;
  const float* row_scale;
  int rows_per_batch;
  float* Y;
  long long ldy;
  int M, C, N1, Np1, ks1, nj;  // nj = chunks of 128 fc1 columns = k-slabs of fc2
  int parts, nb;
  int num_tiles;
  long long* dbg;
  int dbg_flags;  // timing experiments only (results are garbage): 16 = no weight copies, 32 = no conversion, 64 = no GLU work, 128 = no final epilogue, 256 = no X loads, 4096 = no MMAs
};

#define M_T0() (p.dbg ? clock64() : 0)
#define M_ACC(var, t0) do { if (p.dbg) var += clock64() - (t0); } while (0)

struct MlpSmem {
  uint64_t land_full[M_NL], land_empty[M_NL];
  uint64_t x_full, x_empty;
  uint64_t b_full[8], b_empty[8];
  uint64_t acc1_full[2], h_full[2];
  uint64_t acc2_full, acc2_empty;
  uint64_t res_full, res_empty;
  uint32_t tmem_base;
};

// One GLU work item: 32 fc1 columns (= 16 hidden units) x 32 rows of TMEM lane quadrant `quad`.
// acc1 (+ b1) -> value * gelu(gate) -> bf16 hi/lo -> tcgen05.st over the first 16 of the 32 columns just read:
// hidden units 16 i .. 16 i + 15 of the row = k-step i of fc2 (hi part 8 columns, lo part the next 8).
__device__ __forceinline__ void glu_item(uint32_t taddr, float bias_lane, bool cols_ok, int parts) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float hv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int cidx = 4 * e + 2 * u;
      // columns beyond N1 may hold stale TMEM bits: mask the INPUTS (a select), never branch around the gelu --
      // a branch per output serialises the 16 independent erf chains of the item
      float val = __uint_as_float(r[cidx]) + __shfl_sync(0xffffffffu, bias_lane, cidx);
      float gat = __uint_as_float(r[cidx + 1]) + __shfl_sync(0xffffffffu, bias_lane, cidx + 1);
      val = cols_ok ? val : 0.f;
      gat = cols_ok ? gat : 0.f;
      hv[u] = val * gelu_erf_fast(gat);
    }
    split2(hv[0], hv[1], hi[e], lo[e]);
  }
  tmem_st8(taddr, hi);
  if (parts == 2) tmem_st8(taddr + 8, lo);
  tmem_st_wait();
}

__global__ void __launch_bounds__(kMlpThreads, 1) mlp_tc_kernel(const __grid_constant__ MlpArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_launch_dependents();
  MlpSmem* sm = reinterpret_cast<MlpSmem*>(smem_raw);
  const int parts = p.parts;
  const int b_slot_bytes = M_SLAB * parts;
  uint8_t* land = smem_raw + 1024;
  uint8_t* resbuf = land + (size_t)M_NL * M_LAND;            // residual tile: C/32 boxes of [128 rows x 32 floats]
  uint8_t* b_ring = resbuf + (size_t)(p.C / 32) * (M_LAND / 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ks1 = p.ks1, NJ = p.nj, C = p.C;
  const int num_tiles = p.num_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < M_NL; ++i) {
      mbar_init(smem_u32(&sm->land_full[i]), 1);
      mbar_init(smem_u32(&sm->land_empty[i]), 4);
    }
    mbar_init(smem_u32(&sm->x_full), 4);
    mbar_init(smem_u32(&sm->x_empty), 1);
    for (int i = 0; i < 8; ++i) {
      mbar_init(smem_u32(&sm->b_full[i]), 1);
      mbar_init(smem_u32(&sm->b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sm->acc1_full[i]), 1);
      mbar_init(smem_u32(&sm->h_full[i]), 8);
    }
    mbar_init(smem_u32(&sm->acc2_full), 1);
    mbar_init(smem_u32(&sm->acc2_empty), 4);
    mbar_init(smem_u32(&sm->res_full), 1);
    mbar_init(smem_u32(&sm->res_empty), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&sm->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // prologue above overlaps the predecessor's tail (programmatic dependent launch)
  const uint32_t tmem_base = sm->tmem_base;

  if (warp == 0) {
    // =============================== B loader: W1 / W2 blocks =================================
    {  // warp-uniform loop, one elected lane issues the bulk copies
      uint32_t it = 0;
      auto load_block = [&](const uint8_t* img, int ks_total, int np_rows, int slab, int row0, int rows) {
        const int slot = it % p.nb;
        mbar_wait(smem_u32(&sm->b_empty[slot]), ((it / p.nb) & 1) ^ 1);
        const uint32_t bytes = rows * 128;
        const uint32_t full = smem_u32(&sm->b_full[slot]);
        if (elect_one()) {
          mbar_expect_tx(full, (p.dbg_flags & 16) ? 0 : bytes * parts);
          for (int part = 0; part < parts && !(p.dbg_flags & 16); ++part) {
            const uint8_t* src = img + ((size_t)(part * ks_total + slab) * np_rows + row0) * 128;
            bulk_g2s(smem_u32(b_ring + (size_t)slot * b_slot_bytes + part * M_SLAB), src, bytes, full);
          }
        }
        __syncwarp();
        ++it;
      };
      const uint8_t* w1 = reinterpret_cast<const uint8_t*>(p.W1img);
      const uint8_t* w2 = reinterpret_cast<const uint8_t*>(p.W2img);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int s = 0; s < Ks1; ++s) load_block(w1, Ks1, p.Np1, s, 0, min(128, p.Np1));
        for (int j = 0; j < NJ; ++j) {
          if (j + 1 < NJ)
            for (int s = 0; s < Ks1; ++s) load_block(w1, Ks1, p.Np1, s, (j + 1) * 128, min(128, p.Np1 - (j + 1) * 128));
          load_block(w2, NJ, C, j, 0, C);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ================================================
    // the whole warp walks the loop nest; one elected lane issues the MMAs / commits (see tc::elect_one)
    {
      uint32_t b_it = 0, c1_it = 0, t_it = 0;
      long long w_x = 0, w_b = 0, w_h = 0, w_acc2 = 0, t_all = M_T0();
      const bool no_mma = (p.dbg_flags & 4096) != 0;
      const uint32_t xh = tmem_base + XH_COL, xl = tmem_base + XL_COL;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_it) {
        long long tw = M_T0();
        mbar_wait(smem_u32(&sm->x_full), t_it & 1);   // LN(X) image of this tile is in tensor memory
        M_ACC(w_x, tw);
        tc_fence_after();
        auto fc1 = [&](int j) {
          const int buf = c1_it & 1;
          const int ncols = min(128, p.Np1 - j * 128);
          const uint32_t idesc = make_idesc(ncols);
          const uint32_t d_addr = tmem_base + buf * 128;
          for (int s = 0; s < Ks1; ++s) {
            const int b_slot = b_it % p.nb;
            long long tb = M_T0();
            mbar_wait(smem_u32(&sm->b_full[b_slot]), (b_it / p.nb) & 1);
            M_ACC(w_b, tb);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(b_ring + (size_t)b_slot * b_slot_bytes);
            const uint64_t bh0 = make_desc(b_addr), bl0 = make_desc(b_addr + M_SLAB);
            if (elect_one()) {
              if (!no_mma) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t kk = 8 * (s * 4 + k);
                  umma_bf16_tmem_a(d_addr, xh + kk, bh0 + 2 * k, idesc, (s | k) != 0);
                  if (parts == 2) {
                    umma_bf16_tmem_a(d_addr, xh + kk, bl0 + 2 * k, idesc, 1);
                    umma_bf16_tmem_a(d_addr, xl + kk, bh0 + 2 * k, idesc, 1);
                  }
                }
              }
              umma_commit(smem_u32(&sm->b_empty[b_slot]));
            }
            __syncwarp();
            ++b_it;
          }
          if (elect_one()) {
            umma_commit(smem_u32(&sm->acc1_full[buf]));
            if (j == NJ - 1) umma_commit(smem_u32(&sm->x_empty));  // last reader of this tile's X image
          }
          __syncwarp();
          ++c1_it;
        };
        const uint32_t c1_base = c1_it;
        fc1(0);
        for (int j = 0; j < NJ; ++j) {
          if (j + 1 < NJ) fc1(j + 1);
          const uint32_t hit = c1_base + j;
          const int hb = hit & 1;
          tw = M_T0();
          mbar_wait(smem_u32(&sm->h_full[hb]), (hit >> 1) & 1);
          M_ACC(w_h, tw);
          const int b_slot = b_it % p.nb;
          tw = M_T0();
          mbar_wait(smem_u32(&sm->b_full[b_slot]), (b_it / p.nb) & 1);
          M_ACC(w_b, tw);
          tw = M_T0();
          if (j == 0) mbar_wait(smem_u32(&sm->acc2_empty), (t_it & 1) ^ 1);
          M_ACC(w_acc2, tw);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_ring + (size_t)b_slot * b_slot_bytes);
          const uint32_t idesc = make_idesc(C);
          const uint32_t d_addr = tmem_base + ACC2_COL;
          const uint64_t bh0 = make_desc(b_addr), bl0 = make_desc(b_addr + M_SLAB);
          const uint32_t th0 = tmem_base + hb * 128;   // H_j: k-step k at columns 32 k (hi) / 32 k + 8 (lo) of acc1[hb]
          if (elect_one()) {
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_tmem_a(d_addr, th0 + 32 * k, bh0 + 2 * k, idesc, (j | k) != 0);
                if (parts == 2) {
                  umma_bf16_tmem_a(d_addr, th0 + 32 * k, bl0 + 2 * k, idesc, 1);
                  umma_bf16_tmem_a(d_addr, th0 + 32 * k + 8, bh0 + 2 * k, idesc, 1);
                }
              }
            }
            umma_commit(smem_u32(&sm->b_empty[b_slot]));
          }
          __syncwarp();
          ++b_it;
        }
        if (elect_one()) umma_commit(smem_u32(&sm->acc2_full));
        __syncwarp();
      }
      if (p.dbg && lane == 0) {
        long long* d = p.dbg + blockIdx.x * 16;
        d[0] = clock64() - t_all; d[1] = 0; d[2] = w_x; d[3] = w_b; d[4] = w_h; d[5] = w_acc2;
      }
    }
  } else if (warp < 10) {
    // =============================== GLU (warps 2..9): acc1 -> H, in place ======================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t c1_it = 0;
    long long g_wait = 0, g_all = M_T0();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int j = 0; j < NJ; ++j, ++c1_it) {
        const int buf = c1_it & 1;
        float bias_lane[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int n = j * 128 + (half + 2 * g) * 32 + lane;
          bias_lane[g] = n < p.N1 ? __ldg(p.b1 + n) : 0.f;
        }
        long long tg = M_T0();
        mbar_wait(smem_u32(&sm->acc1_full[buf]), (c1_it >> 1) & 1);
        M_ACC(g_wait, tg);
        tc_fence_after();
        if (!(p.dbg_flags & 64)) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int i = half + 2 * g;
            glu_item(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 128 + i * 32, bias_lane[g], j * 128 + i * 32 < p.N1, parts);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm->h_full[buf]));
        __syncwarp();
      }
    }
    if (p.dbg && warp == 2 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[6] = clock64() - g_all; d[7] = g_wait;
    }
  } else if (warp < 14) {
    // =============================== converters (warps 10..13): LN + split -> tensor memory =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int swz = row & 7;
    uint32_t l_it = 0, t_it = 0;
    long long c_wait = 0, c_all = M_T0();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_it) {
      // pass 1: row statistics, as soon as the slabs land
      float s1 = 0.f, s2 = 0.f;
      for (int s = 0; s < Ks1; ++s) {
        const int st = (l_it + s) % M_NL;
        mbar_wait(smem_u32(&sm->land_full[st]), ((l_it + s) / M_NL) & 1);
        if (p.dbg_flags & 32) continue;
        const uint8_t* base = land + (size_t)st * M_LAND + row * 128;
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 a = *reinterpret_cast<const float4*>(base + bx * (M_LAND / 2) + ((c ^ swz) << 4));
            s1 += (a.x + a.y) + (a.z + a.w);
            s2 += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
          }
      }
      const float mean = s1 / (float)C;
      const float rstd = rsqrtf(fmaxf(s2 / (float)C - mean * mean, 0.f) + 1e-5f);
      // pass 1b: LayerNorm + bf16 hi/lo split IN PLACE — a lane rewrites only the 64 bytes (16 channels = one k-step)
      // of its own row it has just read: [hi x 8 | lo x 8] words.  Off the critical path: the previous tile's MMAs run.
      if (!(p.dbg_flags & 32)) {
        for (int s = 0; s < Ks1; ++s) {
          uint8_t* base = land + (size_t)((l_it + s) % M_NL) * M_LAND + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k0 = s * 64 + q * 16;
            uint8_t* src = base + (q >> 1) * (M_LAND / 2);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int cc = (q & 1) * 4 + c;
              float4 v = *reinterpret_cast<const float4*>(src + ((cc ^ swz) << 4));
              const float4 g = ldg4(p.ln_g + k0 + 4 * c), e = ldg4(p.ln_b + k0 + 4 * c);
              v.x = (v.x - mean) * rstd * g.x + e.x; v.y = (v.y - mean) * rstd * g.y + e.y;
              v.z = (v.z - mean) * rstd * g.z + e.z; v.w = (v.w - mean) * rstd * g.w + e.w;
              split2(v.x, v.y, hi[2 * c], lo[2 * c]);
              split2(v.z, v.w, hi[2 * c + 1], lo[2 * c + 1]);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int cc = (q & 1) * 4 + c;
              *reinterpret_cast<uint4*>(src + ((cc ^ swz) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
              *reinterpret_cast<uint4*>(src + (((cc + 2) ^ swz) << 4)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
          }
        }
      }
      // pass 2 (the only part the tensor pipe may wait for): once the previous tile's fc1 MMAs have finished reading the
      // X image, copy the converted row into tensor memory
      long long tc0 = M_T0();
      mbar_wait(smem_u32(&sm->x_empty), (t_it & 1) ^ 1);
      M_ACC(c_wait, tc0);
      tc_fence_after();
      for (int s = 0; s < Ks1; ++s, ++l_it) {
        const int st = l_it % M_NL;
        const uint8_t* base = land + (size_t)st * M_LAND + row * 128;
        if (!(p.dbg_flags & 32)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint8_t* src = base + (q >> 1) * (M_LAND / 2);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int cc = (q & 1) * 4 + c;
              const uint4 h = *reinterpret_cast<const uint4*>(src + ((cc ^ swz) << 4));
              const uint4 l = *reinterpret_cast<const uint4*>(src + (((cc + 2) ^ swz) << 4));
              hi[4 * c] = h.x; hi[4 * c + 1] = h.y; hi[4 * c + 2] = h.z; hi[4 * c + 3] = h.w;
              lo[4 * c] = l.x; lo[4 * c + 1] = l.y; lo[4 * c + 2] = l.z; lo[4 * c + 3] = l.w;
            }
            const uint32_t col = 8 * (s * 4 + q);
            tmem_st8(lane_base + XH_COL + col, hi);
            if (parts == 2) tmem_st8(lane_base + XL_COL + col, lo);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm->land_empty[st]));   // the slot may be refilled: the next tile's X travels now
        __syncwarp();
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm->x_full));
      __syncwarp();
    }
    if (p.dbg && warp == 10 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[11] = clock64() - c_all; d[12] = c_wait;
    }
  } else if (warp < 18) {
    // =============================== final epilogue (warps 14..17) ==============================
    // The residual tile X[tile] was fetched a second time (an L2 hit) into `resbuf` by TMA while the tile's chunks ran;
    // lane = row combines  x + s * (acc2 + b2) [+ res2]  IN PLACE in the swizzled boxes (conflict-free row-per-lane
    // accesses) and the boxes leave by TMA stores: no global-memory latency anywhere in this role, and acc2 is free again
    // after four tcgen05.ld.
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int swz = row & 7;
    uint32_t t_it = 0;
    long long e_wait = 0, e_ld = 0, e_st = 0, e_pf = 0, e_all = M_T0();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_it) {
      const int m = tile * 128 + row;
      long long te = M_T0();
      mbar_wait(smem_u32(&sm->res_full), t_it & 1);
      mbar_wait(smem_u32(&sm->acc2_full), t_it & 1);
      M_ACC(e_wait, te);
      tc_fence_after();
      const float sc = (p.row_scale != nullptr && m < p.M) ? __ldg(p.row_scale + m / p.rows_per_batch) : 1.f;
      long long t1 = M_T0();
      for (int c0 = 0; c0 < C && !(p.dbg_flags & 128); c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + ACC2_COL + c0, r);
        if (c0 + 32 >= C) {   // acc2 is in registers from here on: fc2 of the next tile may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sm->acc2_empty));
          __syncwarp();
        }
        uint8_t* box = resbuf + (size_t)(c0 >> 5) * (M_LAND / 2) + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4* cell = reinterpret_cast<float4*>(box + ((c ^ swz) << 4));
          const float4 x = *cell;
          const float4 b = ldg4(p.b2 + c0 + 4 * c);
          float4 o;
          o.x = x.x + sc * (__uint_as_float(r[4 * c]) + b.x);
          o.y = x.y + sc * (__uint_as_float(r[4 * c + 1]) + b.y);
          o.z = x.z + sc * (__uint_as_float(r[4 * c + 2]) + b.z);
          o.w = x.w + sc * (__uint_as_float(r[4 * c + 3]) + b.w);
          if (p.res2 != nullptr && m < p.M) {
            const float4 x2 = ldg4(p.res2 + (size_t)m * p.ldr2 + c0 + 4 * c);
            o.x += x2.x; o.y += x2.y; o.z += x2.z; o.w += x2.w;
          }
          *cell = o;
        }
      }
      M_ACC(e_ld, t1);
      if (p.dbg_flags & 128) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm->acc2_empty));
      }
      t1 = M_T0();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        for (int c0 = 0; c0 < C; c0 += 32)
          tma_store_2d(&p.tmY, smem_u32(resbuf + (size_t)(c0 >> 5) * (M_LAND / 2) + quad * 32 * 128), c0, tile * 128 + quad * 32);
        bulk_commit_group();
        bulk_wait_group_read<0>();   // the boxes have been read: the next tile's residual may land
        mbar_arrive(smem_u32(&sm->res_empty));
      }
      __syncwarp();
      M_ACC(e_st, t1);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA exits
    (void)e_pf;
    if (p.dbg && warp == 14 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[8] = clock64() - e_all; d[9] = e_wait; d[10] = e_ld; d[13] = e_st; d[14] = e_pf;
    }
  } else {
    // =============================== X loader (TMA) ============================================
    if (lane == 0) {          // landing slots: the operand of fc1
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int s = 0; s < Ks1; ++s, ++it) {
          const int st = it % M_NL;
          mbar_wait(smem_u32(&sm->land_empty[st]), ((it / M_NL) & 1) ^ 1);
          const uint32_t full = smem_u32(&sm->land_full[st]);
          mbar_expect_tx(full, (p.dbg_flags & 256) ? 0 : M_LAND);
          if (!(p.dbg_flags & 256)) {
            const uint32_t dst = smem_u32(land + (size_t)st * M_LAND);
            tma_load_2d(dst, &p.tmX, s * 64, tile * 128, full);
            tma_load_2d(dst + M_LAND / 2, &p.tmX, s * 64 + 32, tile * 128, full);
          }
        }
      }
    } else if (lane == 1) {   // the same rows again, as the residual the final epilogue combines in place (an L2 hit)
      uint32_t t_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_it) {
        mbar_wait(smem_u32(&sm->res_empty), (t_it & 1) ^ 1);
        const uint32_t full = smem_u32(&sm->res_full);
        mbar_expect_tx(full, (p.dbg_flags & 256) ? 0 : (uint32_t)(C / 32) * (M_LAND / 2));
        if (!(p.dbg_flags & 256))
          for (int c0 = 0; c0 < C; c0 += 32)
            tma_load_2d(smem_u32(resbuf + (size_t)(c0 >> 5) * (M_LAND / 2)), &p.tmX, c0, tile * 128, full);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

long long* g_mlp_dbg = nullptr;
int g_mlp_flags = 0;

static PFN_cuTensorMapEncodeTiled mlp_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(ptr);
  }
  return fn;
}

}  // namespace tc
}  // namespace mphsir

using namespace mphsir;

extern "C" MPHSIR_API void mphsir_debug_mlp_counters(long long* buf) { tc::g_mlp_dbg = buf; }
extern "C" MPHSIR_API void mphsir_debug_mlp_flags(int f) { tc::g_mlp_flags = f; }

extern "C" int mphsir_mlp_supported(int C, int hid_pad) {
  return (C == 64 || C == 128) && hid_pad % 16 == 0 && hid_pad > 0;
}

extern "C" int mphsir_mlp_fwd(const mphsir_mlp_params* q, void* stream) {
  MPHSIR_REQUIRE(q && q->X && q->ln_gamma && q->ln_beta && q->W1img && q->b1 && q->W2img && q->b2 && q->Y, "mlp: null operand");
  MPHSIR_REQUIRE(mphsir_mlp_supported(q->C, q->hid_pad), "mlp: fused kernel supports C in {64,128} (got C=%d hid_pad=%d); use the fc1/fc2 GEMM pair", q->C, q->hid_pad);
  MPHSIR_REQUIRE(q->M > 0 && q->ldx % 4 == 0 && q->ldx >= q->C && q->ldy % 4 == 0 && q->ldy >= q->C, "mlp: bad shape");
  MPHSIR_REQUIRE(q->precision == MPHSIR_PREC_BF16X3 || q->precision == MPHSIR_PREC_BF16, "mlp: tensor-core precisions only");
  MPHSIR_REQUIRE(((reinterpret_cast<uintptr_t>(q->X) | reinterpret_cast<uintptr_t>(q->Y)) & 15) == 0 &&
                 ((reinterpret_cast<uintptr_t>(q->W1img) | reinterpret_cast<uintptr_t>(q->W2img)) & 127) == 0, "mlp: operands misaligned");
  if (q->row_scale) MPHSIR_REQUIRE(q->rows_per_batch > 0, "mlp: row_scale needs rows_per_batch");
  if (q->res2) MPHSIR_REQUIRE(q->ldr2 % 4 == 0, "mlp: res2 misaligned");
  tc::MlpArgs a{};
  a.dbg = tc::g_mlp_dbg;
  a.dbg_flags = tc::g_mlp_flags;
  a.X = q->X; a.ldx = q->ldx; a.ln_g = q->ln_gamma; a.ln_b = q->ln_beta;
  a.W1img = q->W1img; a.b1 = q->b1; a.W2img = q->W2img; a.b2 = q->b2;
  a.res2 = q->res2; a.ldr2 = q->ldr2; a.row_scale = q->row_scale; a.rows_per_batch = q->rows_per_batch;
  a.Y = q->Y; a.ldy = q->ldy; a.M = q->M; a.C = q->C;
  a.N1 = 2 * q->hid_pad; a.Np1 = (a.N1 + 15) / 16 * 16; a.ks1 = q->C / 64; a.nj = (a.N1 + 127) / 128;
  a.parts = q->precision == MPHSIR_PREC_BF16X3 ? 2 : 1;
  // shared memory (225 KB at C = 128): 1 KB barriers + 2 landing slots x 32 KB + residual tile C/32 x 16 KB + weight ring
  //   bf16x3: 3 (C = 128) / 4 slots x 32 KB (hi + lo),  bf16: 6 / 8 slots x 16 KB
  a.nb = (a.parts == 2 ? 3 : 6) + (q->C == 64 ? (a.parts == 2 ? 1 : 2) : 0);
  a.num_tiles = (q->M + 127) / 128;
  PFN_cuTensorMapEncodeTiled enc = tc::mlp_encode_fn();
  MPHSIR_REQUIRE(enc != nullptr, "mlp: cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)q->C, (cuuint64_t)q->M};
  cuuint64_t gstr[1] = {(cuuint64_t)q->ldx * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  MPHSIR_REQUIRE(enc(&a.tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(q->X), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "mlp: tensor map encode failed");
  static int sm_count = 0;
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(tc::mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("mlp: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  cuuint64_t ydim[2] = {(cuuint64_t)q->C, (cuuint64_t)q->M};
  cuuint64_t ystr[1] = {(cuuint64_t)q->ldy * 4};
  cuuint32_t ybox[2] = {32, 32};
  MPHSIR_REQUIRE(enc(&a.tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, q->Y, ydim, ystr, ybox, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "mlp: output tensor map encode failed");
  const size_t smem = 1024 + (size_t)tc::M_NL * tc::M_LAND + (size_t)(q->C / 32) * (tc::M_LAND / 2) + (size_t)a.nb * tc::M_SLAB * a.parts;
  const int grid = a.num_tiles < sm_count ? a.num_tiles : sm_count;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc::kMlpThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc::pdl_enabled() ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, tc::mlp_tc_kernel, a);
  if (le != cudaSuccess) {
    set_error("mlp(tc): cudaLaunchKernelEx failed: %s", cudaGetErrorString(le));
    return MPHSIR_ERR_CUDA;
  }
  return check_launch("mlp(tc)");
}
