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
//   * the GLU warps (2.0k clk per chunk, four per scheduler) sit on the FP32 pipe: a 3-register FFMA holds it for two cycles;
//     packed fma.rn.f32x2 arithmetic (FFMA2) and a one-MUFU erf (A&S 7.1.28) each left the 2.0k clk unchanged, so the scalar
//     7.1.26 form stays (it is the more accurate one).  MMA issue is 2.7k clk per chunk: the two are co-critical.
// TMEM columns: acc1[0] 0-127 | acc1[1] 128-255 | acc2 256-383 | X hi 384-447 | X lo 448-511.
// HBM traffic per token: read C (+C residual, an L2 hit), write C floats — the unfused pair moved 2*(C + hidden) more.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace mphsir {
namespace tc {

constexpr int kMlpThreads = 864;  // warp 0: B loader, 1: MMA (peer CTA: relay), 2-17: GLU, 18-21: converters, 22-25: final epilogue, 26: X loader
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
  uint64_t b_full[8], b_peer[8], b_empty[8];
  uint64_t acc1_full[2], h_full[2];
  uint64_t acc2_full, acc2_empty;
  uint64_t res_full, res_empty;
  uint32_t tmem_base;
};

// value * gelu_erf(gate) with the 0.5 of the Gaussian CDF folded into the Abramowitz-Stegun 7.1.26 coefficients
// (|erf error| <= 1.5e-7): 13 FP32 + 2 MUFU instructions per hidden unit — the GLU warps are issue-bound.
__device__ __forceinline__ float glu_unit(float val, float gat) {
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, fabsf(gat), 1.0f)));   // 1 / (1 + p |x| / sqrt 2)
  float pl = fmaf(0.5307027145f, t, -0.7265760135f);
  pl = fmaf(pl, t, 0.7107068705f);
  pl = fmaf(pl, t, -0.142248368f);
  pl = fmaf(pl, t, 0.127414796f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(gat * gat * -0.72134752044f));              // exp(-x^2 / 2)
  const float half_erf = fmaf(-(pl * t), e, 0.5f);                                               // erf(|x| / sqrt 2) / 2
  return (val * gat) * (0.5f + copysignf(half_erf, gat));
}

// One GLU work item: 32 fc1 columns (= 16 hidden units) x 32 rows of TMEM lane quadrant `quad`.
// acc1 (+ b1) -> value * gelu(gate) -> bf16 hi/lo -> tcgen05.st over the first 16 of the 32 columns just read:
// hidden units 16 i .. 16 i + 15 of the row = k-step i of fc2 (hi part 8 columns, lo part the next 8).
// `bias` = the item's 32 interleaved fc1 biases (the same address in every lane: broadcast loads), NULL for an item
// beyond N1 (N1 is a multiple of 32, so an item is valid or not as a whole): its columns hold stale TMEM bits and
// H must be zero there.
__device__ __forceinline__ void glu_item(uint32_t taddr, const float* __restrict__ bias, int parts, long long* prof) {
  uint32_t hi[8], lo[8];
  long long t0 = prof ? clock64() : 0;
  if (bias != nullptr) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    if (prof) { const long long t1 = clock64(); prof[0] += t1 - t0; t0 = t1; }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float4 b = ldg4(bias + 4 * e);
      const float h0 = glu_unit(__uint_as_float(r[4 * e]) + b.x, __uint_as_float(r[4 * e + 1]) + b.y);
      const float h1 = glu_unit(__uint_as_float(r[4 * e + 2]) + b.z, __uint_as_float(r[4 * e + 3]) + b.w);
      split2(h0, h1, hi[e], lo[e]);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) hi[e] = lo[e] = 0u;
  }
  if (prof) { const long long t1 = clock64(); prof[1] += t1 - t0; t0 = t1; }
  tmem_st8(taddr, hi);
  if (parts == 2) tmem_st8(taddr + 8, lo);
  tmem_st_wait();
  if (prof) prof[2] += clock64() - t0;
}

__global__ void __launch_bounds__(kMlpThreads, 1) mlp_tc_kernel(const __grid_constant__ MlpArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_launch_dependents();
  MlpSmem* sm = reinterpret_cast<MlpSmem*>(smem_raw);
  const int parts = p.parts;
  const int b_slot_bytes = (M_SLAB / 2) * parts;   // this CTA's half (64 rows) of a 128-row x 64-k weight block, hi [+ lo]
  uint8_t* land = smem_raw + 1024;
  uint8_t* resbuf = land + (size_t)M_NL * M_LAND;            // residual tile: C/32 boxes of [128 rows x 32 floats]
  uint8_t* b_ring = resbuf + (size_t)(p.C / 32) * (M_LAND / 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ks1 = p.ks1, NJ = p.nj, C = p.C;
  const int num_pairs = (p.num_tiles + 1) >> 1;   // a CTA pair walks tile pairs (2 q, 2 q + 1); an odd tail tile's twin is all out of bounds

  if (threadIdx.x == 0) {
    for (int i = 0; i < M_NL; ++i) {
      mbar_init(smem_u32(&sm->land_full[i]), 1);
      mbar_init(smem_u32(&sm->land_empty[i]), 4);
    }
    mbar_init(smem_u32(&sm->x_full), 8);     // leader: 4 converter warps of each CTA of the pair
    mbar_init(smem_u32(&sm->x_empty), 1);
    for (int i = 0; i < 8; ++i) {
      mbar_init(smem_u32(&sm->b_full[i]), 1);
      mbar_init(smem_u32(&sm->b_peer[i]), 1);   // leader: the peer CTA's half of the weight block has landed
      mbar_init(smem_u32(&sm->b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sm->acc1_full[i]), 1);
      mbar_init(smem_u32(&sm->h_full[i]), 32);  // leader: 16 GLU warps of each CTA
    }
    mbar_init(smem_u32(&sm->acc2_full), 1);
    mbar_init(smem_u32(&sm->acc2_empty), 8); // leader: 4 epilogue warps of each CTA
    mbar_init(smem_u32(&sm->res_full), 1);
    mbar_init(smem_u32(&sm->res_empty), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();   // the peer's mbarriers exist before a remote arrive / multicast commit targets them
  if (warp == 1) tmem_alloc2(smem_u32(&sm->tmem_base), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();  // prologue above overlaps the predecessor's tail (programmatic dependent launch)
  const uint32_t tmem_base = sm->tmem_base;
  const uint32_t crank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;   // this CTA's first tile pair / pairs in flight
  // barriers the MMA issuer waits on live in the leader CTA; both CTAs arrive there
  const uint32_t L_x_full = mapa_cluster(smem_u32(&sm->x_full), 0);
  const uint32_t L_acc2_empty = mapa_cluster(smem_u32(&sm->acc2_empty), 0);
  const uint32_t L_h_full[2] = {mapa_cluster(smem_u32(&sm->h_full[0]), 0), mapa_cluster(smem_u32(&sm->h_full[1]), 0)};

  if (warp == 0) {
    // =============================== B loader: W1 / W2 blocks =================================
    {  // warp-uniform loop, one elected lane issues the bulk copies
      uint32_t it = 0;
      auto load_block = [&](const uint8_t* img, int ks_total, int np_rows, int slab, int row0, int rows) {
        const int slot = it % p.nb;
        mbar_wait(smem_u32(&sm->b_empty[slot]), ((it / p.nb) & 1) ^ 1);
        const uint32_t bytes = (rows >> 1) * 128;   // this CTA's half of the block's rows (= of the MMA's N)
        row0 += (int)crank * (rows >> 1);
        const uint32_t full = smem_u32(&sm->b_full[slot]);
        if (elect_one()) {
          mbar_expect_tx(full, (p.dbg_flags & 16) ? 0 : bytes * parts);
          for (int part = 0; part < parts && !(p.dbg_flags & 16); ++part) {
            const uint8_t* src = img + ((size_t)(part * ks_total + slab) * np_rows + row0) * 128;
            bulk_g2s(smem_u32(b_ring + (size_t)slot * b_slot_bytes + part * (M_SLAB / 2)), src, bytes, full);
          }
        }
        __syncwarp();
        ++it;
      };
      const uint8_t* w1 = reinterpret_cast<const uint8_t*>(p.W1img);
      const uint8_t* w2 = reinterpret_cast<const uint8_t*>(p.W2img);
      for (int pr = pair0; pr < num_pairs; pr += npairs) {
      const int tile = 2 * pr + (int)crank; (void)tile;
        for (int s = 0; s < Ks1; ++s) load_block(w1, Ks1, p.Np1, s, 0, min(128, p.Np1));
        for (int j = 0; j < NJ; ++j) {
          if (j + 1 < NJ)
            for (int s = 0; s < Ks1; ++s) load_block(w1, Ks1, p.Np1, s, (j + 1) * 128, min(128, p.Np1 - (j + 1) * 128));
          load_block(w2, NJ, C, j, 0, C);
        }
      }
    }
  } else if (warp == 1 && crank != 0) {
    // =============================== peer CTA: weight-block relay =============================
    // only the leader issues MMAs; this warp tells it when the peer's half of a weight block has landed
    const uint32_t slots_per_pair = (uint32_t)(Ks1 * NJ + NJ);
    uint32_t b_it = 0;
    for (int pr = pair0; pr < num_pairs; pr += npairs)
      for (uint32_t i = 0; i < slots_per_pair; ++i, ++b_it) {
        const int b_slot = b_it % p.nb;
        mbar_wait(smem_u32(&sm->b_full[b_slot]), (b_it / p.nb) & 1);
        // relaxed: the block sits complete in THIS CTA's shared memory (complete_tx fired) and is read there by this SM's
        // tensor core; nothing travels with the arrive, and a release.cluster fence per block would cap the relay at one
        // block per ~1.5k clk — more than the MMAs of a block take
        if (lane == 0) mbar_arrive_remote_relaxed(mapa_cluster(smem_u32(&sm->b_peer[b_slot]), 0));
        __syncwarp();
      }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA of the pair) ========================
    // M = 256 instructions over both CTAs' tensor / shared memory (cta_group::2): 74 clk per 256 x 128 x 16 against 100 clk
    // per 128 x 128 x 16 of a single CTA (tools/mma_rate.cu).  The whole warp walks the loop nest; one elected lane issues
    // the MMAs / commits (see tc::elect_one); every commit is multicast to the same barrier in both CTAs.
    {
      uint32_t b_it = 0, c1_it = 0, t_it = 0;
      long long w_x = 0, w_b = 0, w_h = 0, w_acc2 = 0, t_all = M_T0();
      const bool no_mma = (p.dbg_flags & 4096) != 0;
      const uint32_t xh = tmem_base + XH_COL, xl = tmem_base + XL_COL;
      for (int pr = pair0; pr < num_pairs; pr += npairs, ++t_it) {
        long long tw = M_T0();
        mbar_wait_cluster(smem_u32(&sm->x_full), t_it & 1);   // LN(X) images of both tiles are in tensor memory
        M_ACC(w_x, tw);
        tc_fence_after();
        auto wait_b = [&](int b_slot) {
          long long tb = M_T0();
          mbar_wait(smem_u32(&sm->b_full[b_slot]), (b_it / p.nb) & 1);
          mbar_wait_cluster(smem_u32(&sm->b_peer[b_slot]), (b_it / p.nb) & 1);
          M_ACC(w_b, tb);
        };
        auto fc1 = [&](int j) {
          const int buf = c1_it & 1;
          const int ncols = min(128, p.Np1 - j * 128);
          const uint32_t idesc = make_idesc2(ncols);
          const uint32_t d_addr = tmem_base + buf * 128;
          for (int s = 0; s < Ks1; ++s) {
            const int b_slot = b_it % p.nb;
            wait_b(b_slot);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(b_ring + (size_t)b_slot * b_slot_bytes);
            const uint64_t bh0 = make_desc(b_addr), bl0 = make_desc(b_addr + M_SLAB / 2);
            if (elect_one()) {
              if (!no_mma) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t kk = 8 * (s * 4 + k);
                  umma2_bf16_tmem_a(d_addr, xh + kk, bh0 + 2 * k, idesc, (s | k) != 0);
                  if (parts == 2) {
                    umma2_bf16_tmem_a(d_addr, xh + kk, bl0 + 2 * k, idesc, 1);
                    umma2_bf16_tmem_a(d_addr, xl + kk, bh0 + 2 * k, idesc, 1);
                  }
                }
              }
              umma2_commit(smem_u32(&sm->b_empty[b_slot]));
            }
            __syncwarp();
            ++b_it;
          }
          if (elect_one()) {
            umma2_commit(smem_u32(&sm->acc1_full[buf]));
            if (j == NJ - 1) umma2_commit(smem_u32(&sm->x_empty));  // last reader of this pair's X images
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
          mbar_wait_cluster(smem_u32(&sm->h_full[hb]), (hit >> 1) & 1);
          M_ACC(w_h, tw);
          const int b_slot = b_it % p.nb;
          wait_b(b_slot);
          tw = M_T0();
          if (j == 0) mbar_wait_cluster(smem_u32(&sm->acc2_empty), (t_it & 1) ^ 1);
          M_ACC(w_acc2, tw);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_ring + (size_t)b_slot * b_slot_bytes);
          const uint32_t idesc = make_idesc2(C);
          const uint32_t d_addr = tmem_base + ACC2_COL;
          const uint64_t bh0 = make_desc(b_addr), bl0 = make_desc(b_addr + M_SLAB / 2);
          const uint32_t th0 = tmem_base + hb * 128;   // H_j: k-step k at columns 32 k (hi) / 32 k + 8 (lo) of acc1[hb]
          if (elect_one()) {
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma2_bf16_tmem_a(d_addr, th0 + 32 * k, bh0 + 2 * k, idesc, (j | k) != 0);
                if (parts == 2) {
                  umma2_bf16_tmem_a(d_addr, th0 + 32 * k, bl0 + 2 * k, idesc, 1);
                  umma2_bf16_tmem_a(d_addr, th0 + 32 * k + 8, bh0 + 2 * k, idesc, 1);
                }
              }
            }
            umma2_commit(smem_u32(&sm->b_empty[b_slot]));
          }
          __syncwarp();
          ++b_it;
        }
        if (elect_one()) umma2_commit(smem_u32(&sm->acc2_full));
        __syncwarp();
      }
      if (p.dbg && lane == 0) {
        long long* d = p.dbg + blockIdx.x * 16;
        d[0] = clock64() - t_all; d[1] = 0; d[2] = w_x; d[3] = w_b; d[4] = w_h; d[5] = w_acc2;
      }
    }
  } else if (warp < 18) {
    // =============================== GLU (warps 2..17): acc1 -> H, in place =====================
    // one item (TMEM lane quadrant x 32-column group) per warp and chunk: four GLU warps per scheduler hide each other's
    // tcgen05.ld / st latencies
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    uint32_t c1_it = 0;
    long long g_wait = 0, g_all = M_T0();
    long long g_prof[3] = {0, 0, 0};
    for (int pr = pair0; pr < num_pairs; pr += npairs) {
      for (int j = 0; j < NJ; ++j, ++c1_it) {
        const int buf = c1_it & 1;
        const int n0 = j * 128 + grp * 32;
        long long tg = M_T0();
        mbar_wait(smem_u32(&sm->acc1_full[buf]), (c1_it >> 1) & 1);
        M_ACC(g_wait, tg);
        tc_fence_after();
        if (!(p.dbg_flags & 64))
          glu_item(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 128 + grp * 32, n0 < p.N1 ? p.b1 + n0 : nullptr, parts, p.dbg ? g_prof : nullptr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(L_h_full[buf]);
        __syncwarp();
      }
    }
    if (p.dbg && warp == 2 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[6] = clock64() - g_all; d[7] = g_wait; d[13] = g_prof[0]; d[14] = g_prof[1]; d[15] = g_prof[2];
    }
  } else if (warp < 22) {
    // =============================== converters (warps 18..21): LN + split -> tensor memory =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int swz = row & 7;
    uint32_t l_it = 0, t_it = 0;
    long long c_wait = 0, c_all = M_T0();
    for (int pr = pair0; pr < num_pairs; pr += npairs, ++t_it) {
      const int tile = 2 * pr + (int)crank; (void)tile;
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
      if (lane == 0) mbar_arrive_remote_relaxed(L_x_full);
      __syncwarp();
    }
    if (p.dbg && warp == 18 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[11] = clock64() - c_all; d[12] = c_wait;
    }
  } else if (warp < 26) {
    // =============================== final epilogue (warps 22..25) ==============================
    // The residual tile X[tile] was fetched a second time (an L2 hit) into `resbuf` by TMA while the tile's chunks ran;
    // lane = row combines  x + s * (acc2 + b2) [+ res2]  IN PLACE in the swizzled boxes (conflict-free row-per-lane
    // accesses) and the boxes leave by TMA stores: no global-memory latency anywhere in this role, and acc2 is free again
    // after four tcgen05.ld.
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int swz = row & 7;
    uint32_t t_it = 0;
    long long e_wait = 0, e_ld = 0, e_st = 0, e_pf = 0, e_all = M_T0();
    for (int pr = pair0; pr < num_pairs; pr += npairs, ++t_it) {
      const int tile = 2 * pr + (int)crank; (void)tile;
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
          if (lane == 0) mbar_arrive_remote_relaxed(L_acc2_empty);
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
        if (lane == 0) mbar_arrive_remote_relaxed(L_acc2_empty);
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
    if (p.dbg && warp == 22 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[8] = clock64() - e_all; d[9] = e_wait; d[10] = e_ld;
    }
  } else {
    // =============================== X loader (TMA) ============================================
    if (lane == 0) {          // landing slots: the operand of fc1
      uint32_t it = 0;
      for (int pr = pair0; pr < num_pairs; pr += npairs) {
      const int tile = 2 * pr + (int)crank; (void)tile;
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
      for (int pr = pair0; pr < num_pairs; pr += npairs, ++t_it) {
      const int tile = 2 * pr + (int)crank; (void)tile;
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
  cluster_sync_all();   // the leader's MMAs read the peer's shared / tensor memory: neither CTA leaves early
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
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
  a.nb = 6;   // weight ring: 6 slots x (64 rows x 128 B) x parts — each CTA of the pair stages half of every block
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
  const size_t smem = 1024 + (size_t)tc::M_NL * tc::M_LAND + (size_t)(q->C / 32) * (tc::M_LAND / 2) + (size_t)a.nb * (tc::M_SLAB / 2) * a.parts;
  const int pairs = (a.num_tiles + 1) / 2;
  const int grid = 2 * (pairs < sm_count / 2 ? pairs : sm_count / 2);   // CTA pairs (clusters of 2 on one TPC)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc::kMlpThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc::pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, tc::mlp_tc_kernel, a);
  if (le != cudaSuccess) {
    set_error("mlp(tc): cudaLaunchKernelEx failed: %s", cudaGetErrorString(le));
    return MPHSIR_ERR_CUDA;
  }
  return check_launch("mlp(tc)");
}
