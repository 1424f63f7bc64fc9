// Fused gated MLP on tcgen05 (sm_100a):   Y = X + s * ( fc2( value * gelu(gate) ) + b2 ) [+ R2],
//   [value|gate] = LN(X) W1^T + b1                       (PGSSTB.forward :719 + GatedMlp.forward :76-82)
//
// The 2*hidden intermediate never leaves the SM.  Per 128-row tile the hidden dimension is walked in chunks of
// 64 units (= 128 interleaved (value, gate) fc1 columns = one 64-k slab of fc2):
//     MMA  : acc1[j&1] (TMEM, 128 cols)  = LN(X) . W1_j^T                (K = C, bf16 hi/lo split operands)
//     GLU  : 16 warps drain acc1 with tcgen05.ld, add b1, value*gelu(gate), split to bf16 hi/lo and write the
//            [128 x 64] tile H_j back into TENSOR MEMORY with tcgen05.st (two bf16 per 32-bit column: the row-per-lane
//            layout of the accumulator is the layout of a TMEM A operand, so no transpose and no shared memory)
//     MMA  : acc2 (TMEM, C cols)        += H_j . W2_j^T                  (K = 64, A from TMEM, B from shared memory)
// (ncu / role counters: with both operands in shared memory the N = 128 MMAs are bound by its 128 B/clk — 8 KB of operand
// reads per 64-clk instruction — and the MMA thread spent 75 % of the kernel blocked on issue; H in TMEM removes a third
// of the operand reads, the H stores, and frees 64 KB for a deeper weight ring.  debug flag 8 selects the old
// shared-memory H path.)
// with fc1 of chunk j+1 issued before fc2 of chunk j so the tensor pipe never waits for the GLU warps.
// X slabs arrive by TMA (2-D tensor map) and are LayerNorm-ed / split in place by 8 converter warps exactly as in
// gemm_tc.cu; W1 / W2 blocks stream through a cp.async.bulk ring.  HBM traffic per token: read C (+C residual),
// write C floats — the unfused pair moved 2*(C + hidden) more.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace mphsir {
namespace tc {

constexpr int kMlpThreads = 608;  // warp 0: B loader, 1: MMA, 2-9: GLU + final epilogue, 10-17: converters, 18: A loader
constexpr int M_SLAB = 128 * 128;        // one bf16 part of a 128-row x 64-k slab
constexpr int M_STAGE = 128 * 64 * 4;    // fp32 landing slot (converted in place)
constexpr int M_NA = 2;                  // A ring slots (C <= 128)
constexpr int M_STG_FLOATS = 32 * 32;

struct MlpArgs {
  alignas(64) CUtensorMap tmX;
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
  int dbg_flags;  // timing experiments only: 1 = skip gelu, 2 = skip bias shuffles, 4 = skip bf16 split
  int h_tmem;     // 1: the GLU output H lives in tensor memory (A operand of fc2 from TMEM); 0: shared-memory H ring
};

#define M_T0() (p.dbg ? clock64() : 0)
#define M_ACC(var, t0) do { if (p.dbg) var += clock64() - (t0); } while (0)

struct MlpSmem {
  uint64_t stage_full[M_NA], a_full[M_NA], a_empty[M_NA];
  uint64_t b_full[8], b_empty[8];
  uint64_t acc1_full[2], acc1_empty[2];
  uint64_t h_full[2], h_empty[2];
  uint64_t acc2_full, acc2_empty;
  uint32_t tmem_base;
};

// One GLU work item: 32 fc1 columns (= 16 hidden units) x 32 rows of TMEM lane quadrant `quad`.
// acc1 (+ b1) -> value * gelu(gate) -> bf16 hi/lo -> chunks 2i, 2i+1 of the K-major swizzled H slab row.
__device__ __forceinline__ void glu_item(const MlpArgs& p, uint32_t tmem_col_addr, float bias_lane, bool cols_ok, uint8_t* hdst,
                                         int row, int i, int parts, uint32_t h_taddr) {
  uint32_t r[32];
  tmem_ld32(tmem_col_addr, r);
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float hv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int cidx = 4 * e + 2 * u;
      // columns beyond N1 may hold stale TMEM bits: mask the INPUTS (a select), never branch around the gelu --
      // a branch per output serialises the 16 independent erf chains of the item (ncu: 4.2k clk per item)
      float val = __uint_as_float(r[cidx]) + __shfl_sync(0xffffffffu, bias_lane, cidx);
      float gat = __uint_as_float(r[cidx + 1]) + __shfl_sync(0xffffffffu, bias_lane, cidx + 1);
      val = cols_ok ? val : 0.f;
      gat = cols_ok ? gat : 0.f;
      hv[u] = val * gelu_erf_fast(gat);
    }
    split2(hv[0], hv[1], hi[e], lo[e]);
  }
  if (p.h_tmem) {
    // hidden units 16 i .. 16 i + 15 of this row = columns 8 i .. 8 i + 7 of the hi part (lo part 32 columns further)
    tmem_st8(h_taddr + 8 * i, hi);
    if (parts == 2) tmem_st8(h_taddr + 32 + 8 * i, lo);
    tmem_st_wait();
    return;
  }
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int chunk = 2 * i + cc;
    const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(hdst + off) = make_uint4(hi[4 * cc], hi[4 * cc + 1], hi[4 * cc + 2], hi[4 * cc + 3]);
    if (parts == 2)
      *reinterpret_cast<uint4*>(hdst + M_SLAB + off) = make_uint4(lo[4 * cc], lo[4 * cc + 1], lo[4 * cc + 2], lo[4 * cc + 3]);
  }
  (void)p;
}

// GLU stage of one hidden chunk for one warp: 16 warps (8 epilogue + 8 converter warps) share the 16 items
// (4 TMEM lane quadrants x 4 column groups) of a chunk; group index `i` is fixed per warp.
__device__ __forceinline__ void glu_chunk(const MlpArgs& p, MlpSmem* sm, uint8_t* h_ring, int h_slot_bytes, uint32_t tmem_base,
                                          int j, uint32_t c1_it, uint32_t h_it, int quad, int i, int lane, int parts) {
  const int buf = c1_it & 1, hs = h_it & 1;
  const int n = j * 128 + i * 32 + lane;
  const float bias_lane = n < p.N1 ? __ldg(p.b1 + n) : 0.f;
  mbar_wait(smem_u32(&sm->acc1_full[buf]), (c1_it >> 1) & 1);
  mbar_wait(smem_u32(&sm->h_empty[hs]), ((h_it >> 1) & 1) ^ 1);
  tc_fence_after();
  glu_item(p, tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 128 + i * 32, bias_lane, j * 128 + i * 32 < p.N1,
           h_ring + (size_t)hs * h_slot_bytes, quad * 32 + lane, i, parts,
           tmem_base + ((uint32_t)(quad * 32) << 16) + 384 + hs * 64);
  if (!p.h_tmem) fence_proxy_async();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    mbar_arrive(smem_u32(&sm->h_full[hs]));
    mbar_arrive(smem_u32(&sm->acc1_empty[buf]));
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kMlpThreads, 1) mlp_tc_kernel(const __grid_constant__ MlpArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_launch_dependents();
  MlpSmem* sm = reinterpret_cast<MlpSmem*>(smem_raw);
  const int parts = p.parts;
  const int h_slot_bytes = M_SLAB * parts;
  const int b_slot_bytes = M_SLAB * parts;
  uint8_t* a_ring = smem_raw + 1024;
  uint8_t* h_ring = a_ring + (size_t)M_NA * M_STAGE;
  // H in tensor memory: no H ring, the final-epilogue transpose buffers (8 x 4 KB) own their 32 KB.
  // H in shared memory: they alias H slot 0 — when acc2_full fires every MMA that read the H slabs of this tile has
  // completed, and the next tile's first GLU write waits for this warp's own epilogue (bar.sync 2 below)
  uint8_t* b_ring = h_ring + (p.h_tmem ? (size_t)8 * M_STG_FLOATS * 4 : 2 * (size_t)h_slot_bytes);
  float* staging = reinterpret_cast<float*>(h_ring);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ks1 = p.ks1, NJ = p.nj, C = p.C;
  const int num_tiles = p.num_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < M_NA; ++i) {
      mbar_init(smem_u32(&sm->stage_full[i]), 1);
      mbar_init(smem_u32(&sm->a_full[i]), 8);
      mbar_init(smem_u32(&sm->a_empty[i]), 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(smem_u32(&sm->b_full[i]), 1);
      mbar_init(smem_u32(&sm->b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sm->acc1_full[i]), 1);
      mbar_init(smem_u32(&sm->acc1_empty[i]), 16);  // 8 epilogue + 8 converter warps run the GLU stage
      mbar_init(smem_u32(&sm->h_full[i]), 16);
      mbar_init(smem_u32(&sm->h_empty[i]), 1);
    }
    mbar_init(smem_u32(&sm->acc2_full), 1);
    mbar_init(smem_u32(&sm->acc2_empty), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&sm->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // prologue above overlaps the predecessor's tail (programmatic dependent launch)
  const uint32_t tmem_base = sm->tmem_base;
  const uint32_t ACC2_COL = 256;

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
          mbar_expect_tx(full, bytes * parts);
          for (int part = 0; part < parts; ++part) {
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
      uint32_t a_it = 0, b_it = 0, c1_it = 0, h_it = 0, t_it = 0;
      long long w_acc1 = 0, w_a = 0, w_b = 0, w_h = 0, w_acc2 = 0, t_all = M_T0();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_it) {
        const uint32_t a_base = a_it;
        auto fc1 = [&](int j) {
          const int buf = c1_it & 1;
          long long tw = M_T0();
          mbar_wait(smem_u32(&sm->acc1_empty[buf]), ((c1_it >> 1) & 1) ^ 1);
          M_ACC(w_acc1, tw);
          tc_fence_after();
          const int ncols = min(128, p.Np1 - j * 128);
          const uint32_t idesc = make_idesc(ncols);
          const uint32_t d_addr = tmem_base + buf * 128;
          for (int s = 0; s < Ks1; ++s) {
            const uint32_t a_slot = (a_base + s) % M_NA;
            tw = M_T0();
            if (j == 0) mbar_wait(smem_u32(&sm->a_full[a_slot]), ((a_base + s) / M_NA) & 1);
            M_ACC(w_a, tw);
            const int b_slot = b_it % p.nb;
            tw = M_T0();
            mbar_wait(smem_u32(&sm->b_full[b_slot]), (b_it / p.nb) & 1);
            M_ACC(w_b, tw);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_ring + (size_t)a_slot * M_STAGE);
            const uint32_t b_addr = smem_u32(b_ring + (size_t)b_slot * b_slot_bytes);
            const uint64_t ah0 = make_desc(a_addr), bh0 = make_desc(b_addr);
            const uint64_t al0 = make_desc(a_addr + M_SLAB), bl0 = make_desc(b_addr + M_SLAB);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(d_addr, ah0 + 2 * k, bh0 + 2 * k, idesc, (s | k) != 0);
                if (parts == 2) {
                  umma_bf16(d_addr, ah0 + 2 * k, bl0 + 2 * k, idesc, 1);
                  umma_bf16(d_addr, al0 + 2 * k, bh0 + 2 * k, idesc, 1);
                }
              }
              umma_commit(smem_u32(&sm->b_empty[b_slot]));
              if (j == NJ - 1) umma_commit(smem_u32(&sm->a_empty[a_slot]));  // last reader of this X slab
            }
            __syncwarp();
            ++b_it;
          }
          if (elect_one()) umma_commit(smem_u32(&sm->acc1_full[buf]));
          __syncwarp();
          ++c1_it;
        };
        fc1(0);
        for (int j = 0; j < NJ; ++j) {
          if (j + 1 < NJ) fc1(j + 1);
          const int hs = h_it & 1;
          long long tw = M_T0();
          mbar_wait(smem_u32(&sm->h_full[hs]), (h_it >> 1) & 1);
          M_ACC(w_h, tw);
          const int b_slot = b_it % p.nb;
          tw = M_T0();
          mbar_wait(smem_u32(&sm->b_full[b_slot]), (b_it / p.nb) & 1);
          M_ACC(w_b, tw);
          tw = M_T0();
          if (j == 0) mbar_wait(smem_u32(&sm->acc2_empty), (t_it & 1) ^ 1);
          M_ACC(w_acc2, tw);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(h_ring + (size_t)hs * h_slot_bytes);
          const uint32_t b_addr = smem_u32(b_ring + (size_t)b_slot * b_slot_bytes);
          const uint32_t idesc = make_idesc(C);
          const uint32_t d_addr = tmem_base + ACC2_COL;
          const uint64_t ah0 = make_desc(a_addr), bh0 = make_desc(b_addr);
          const uint64_t al0 = make_desc(a_addr + M_SLAB), bl0 = make_desc(b_addr + M_SLAB);
          const uint32_t th0 = tmem_base + 384 + hs * 64, tl0 = th0 + 32;   // H hi / lo parts in tensor memory
          if (elect_one()) {
            if (p.h_tmem) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_tmem_a(d_addr, th0 + 8 * k, bh0 + 2 * k, idesc, (j | k) != 0);
                if (parts == 2) {
                  umma_bf16_tmem_a(d_addr, th0 + 8 * k, bl0 + 2 * k, idesc, 1);
                  umma_bf16_tmem_a(d_addr, tl0 + 8 * k, bh0 + 2 * k, idesc, 1);
                }
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(d_addr, ah0 + 2 * k, bh0 + 2 * k, idesc, (j | k) != 0);
                if (parts == 2) {
                  umma_bf16(d_addr, ah0 + 2 * k, bl0 + 2 * k, idesc, 1);
                  umma_bf16(d_addr, al0 + 2 * k, bh0 + 2 * k, idesc, 1);
                }
              }
            }
            umma_commit(smem_u32(&sm->b_empty[b_slot]));
            umma_commit(smem_u32(&sm->h_empty[hs]));
          }
          __syncwarp();
          ++b_it;
          ++h_it;
        }
        if (elect_one()) umma_commit(smem_u32(&sm->acc2_full));
        __syncwarp();
        a_it += Ks1;
      }
      if (p.dbg && lane == 0) {
        long long* d = p.dbg + blockIdx.x * 16;
        d[0] = clock64() - t_all; d[1] = w_acc1; d[2] = w_a; d[3] = w_b; d[4] = w_h; d[5] = w_acc2;
      }
    }
  } else if (warp < 10) {
    // =============================== GLU + final epilogue (warps 2..9) =========================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stg = staging + (warp - 2) * M_STG_FLOATS;
    const int c4 = lane & 7, rsub = lane >> 3;
    uint32_t c1_it = 0, h_it = 0, t_it = 0;
    long long g_glu = 0, g_epi = 0, g_all = M_T0();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_it) {
      const int m0 = tile * 128;
      // ---- GLU: acc1 -> H slabs (column group i = half; groups 2, 3 belong to the converter warps) ----
      long long tg = M_T0();
      for (int j = 0; j < NJ; ++j, ++c1_it, ++h_it)
        glu_chunk(p, sm, h_ring, h_slot_bytes, tmem_base, j, c1_it, h_it, quad, half, lane, parts);
      M_ACC(g_glu, tg);
      long long te = M_T0();
      // ---- final epilogue: acc2 + b2, residual(s) -> Y (smem-transposed, 128-byte coalesced) ----
      mbar_wait(smem_u32(&sm->acc2_full), t_it & 1);
      tc_fence_after();
      const int mrow0 = m0 + quad * 32;
      for (int c0 = half * 32; c0 < C; c0 += 64) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + ACC2_COL + c0, r);
        const int n = c0 + 4 * c4;
        const float4 bias4 = ldg4(p.b2 + n);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(stg + lane * 32 + 4 * (q ^ (lane & 7))) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        __syncwarp();
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          float4 acc[4], x1[4], x2[4];
          bool ok[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (hb * 4 + i) * 4 + rsub;
            const int m = mrow0 + rr;
            ok[i] = m < p.M;
            acc[i] = *reinterpret_cast<const float4*>(stg + rr * 32 + 4 * (c4 ^ (rr & 7)));
            x1[i] = x2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[i]) {
              x1[i] = ldg4(p.X + (size_t)m * p.ldx + n);
              if (p.res2 != nullptr) x2[i] = ldg4(p.res2 + (size_t)m * p.ldr2 + n);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!ok[i]) continue;
            const int m = mrow0 + (hb * 4 + i) * 4 + rsub;
            const float sc = p.row_scale != nullptr ? __ldg(p.row_scale + m / p.rows_per_batch) : 1.f;
            float4 o;
            o.x = x1[i].x + sc * (acc[i].x + bias4.x) + x2[i].x;
            o.y = x1[i].y + sc * (acc[i].y + bias4.y) + x2[i].y;
            o.z = x1[i].z + sc * (acc[i].z + bias4.z) + x2[i].z;
            o.w = x1[i].w + sc * (acc[i].w + bias4.w) + x2[i].w;
            *reinterpret_cast<float4*>(p.Y + (size_t)m * p.ldy + n) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm->acc2_empty));
      __syncwarp();  // bar.sync is .aligned: the warp must be converged again after the lane-0 branch (synccheck)
      // the transpose buffers alias the H ring: nobody (epilogue or converter warp) may start the next tile's GLU
      // writes before every epilogue warp has left its staging area
      if (!p.h_tmem && tile + (int)gridDim.x < num_tiles) asm volatile("bar.sync 2, 512;" ::: "memory");
      M_ACC(g_epi, te);
    }
    if (p.dbg && warp == 2 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[6] = clock64() - g_all; d[7] = 0; d[8] = 0; d[9] = g_glu; d[10] = g_epi;
    }
  } else if (warp == 18) {
    // =============================== A loader (TMA) ===========================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int s = 0; s < Ks1; ++s, ++it) {
          const int st = it % M_NA;
          mbar_wait(smem_u32(&sm->a_empty[st]), ((it / M_NA) & 1) ^ 1);
          const uint32_t full = smem_u32(&sm->stage_full[st]);
          mbar_expect_tx(full, M_STAGE);
          tma_load_2d(smem_u32(a_ring + (size_t)st * M_STAGE), &p.tmX, s * 64, tile * 128, full);
        }
      }
    }
  } else {
    // =============================== converters (warps 10..17): LN + split, in place ===========
    const int ct = threadIdx.x - 10 * 32;
    const int chunk = ct & 7, rbase = ct >> 3;
    const int gquad = warp & 3, ggroup = 2 + ((warp - 10) >> 2);  // this warp's GLU item (TMEM quadrant, column group)
    uint32_t a_it = 0, c1_it = 0, h_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = tile * 128;
      float sm_[4] = {0.f, 0.f, 0.f, 0.f}, sq_[4] = {0.f, 0.f, 0.f, 0.f};
      for (int s = 0; s < Ks1; ++s) {
        const int st = (a_it + s) % M_NA;
        mbar_wait(smem_u32(&sm->stage_full[st]), ((a_it + s) / M_NA) & 1);
        if (s * 64 + chunk * 8 < C) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rbase + 32 * i;
            const float* sp = reinterpret_cast<const float*>(a_ring + (size_t)st * M_STAGE + r * 256 + chunk * 32);
            const float4 a = *reinterpret_cast<const float4*>(sp);
            const float4 b = *reinterpret_cast<const float4*>(sp + 4);
            sm_[i] += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
            sq_[i] += ((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w)) + ((b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w));
          }
        }
      }
      float mean[4], rstd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          sm_[i] += __shfl_xor_sync(0xffffffffu, sm_[i], o);
          sq_[i] += __shfl_xor_sync(0xffffffffu, sq_[i], o);
        }
        mean[i] = sm_[i] / (float)C;
        rstd[i] = rsqrtf(fmaxf(sq_[i] / (float)C - mean[i] * mean[i], 0.f) + 1e-5f);
      }
      for (int s = 0; s < Ks1; ++s, ++a_it) {
        const int st = a_it % M_NA;
        const int k = s * 64 + chunk * 8;
        const bool kin = k < C;
        float4 v0[4], v1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rbase + 32 * i;
          if (kin) {
            const float* sp = reinterpret_cast<const float*>(a_ring + (size_t)st * M_STAGE + r * 256 + chunk * 32);
            v0[i] = *reinterpret_cast<const float4*>(sp);
            v1[i] = *reinterpret_cast<const float4*>(sp + 4);
          } else {
            v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            v1[i] = v0[i];
          }
        }
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, e0 = g0, e1 = g0;
        if (kin) {
          g0 = ldg4(p.ln_g + k); g1 = ldg4(p.ln_g + k + 4);
          e0 = ldg4(p.ln_b + k); e1 = ldg4(p.ln_b + k + 4);
        }
        __syncwarp();
        asm volatile("bar.sync 1, 256;" ::: "memory");  // every converter has read its cells of the slot
        uint8_t* dst = a_ring + (size_t)st * M_STAGE;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rbase + 32 * i;
          if (kin) {
            const float a = rstd[i], mu = mean[i];
            v0[i].x = (v0[i].x - mu) * a * g0.x + e0.x; v0[i].y = (v0[i].y - mu) * a * g0.y + e0.y;
            v0[i].z = (v0[i].z - mu) * a * g0.z + e0.z; v0[i].w = (v0[i].w - mu) * a * g0.w + e0.w;
            v1[i].x = (v1[i].x - mu) * a * g1.x + e1.x; v1[i].y = (v1[i].y - mu) * a * g1.y + e1.y;
            v1[i].z = (v1[i].z - mu) * a * g1.z + e1.z; v1[i].w = (v1[i].w - mu) * a * g1.w + e1.w;
          }
          uint4 hi, lo;
          split2(v0[i].x, v0[i].y, hi.x, lo.x);
          split2(v0[i].z, v0[i].w, hi.y, lo.y);
          split2(v1[i].x, v1[i].y, hi.z, lo.z);
          split2(v1[i].z, v1[i].w, hi.w, lo.w);
          const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst + off) = hi;
          if (parts == 2) *reinterpret_cast<uint4*>(dst + M_SLAB + off) = lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm->a_full[st]));
      }
      (void)m0;
      // ---- then help with the GLU stage of this tile (H slots double as the epilogue warps' staging: wait until
      //      they have finished the previous tile's final epilogue) ----
      __syncwarp();  // converged again after the lane-0 arrive of the last slab (bar.sync is .aligned)
      if (!p.h_tmem && tile != (int)blockIdx.x) asm volatile("bar.sync 2, 512;" ::: "memory");
      for (int j = 0; j < NJ; ++j, ++c1_it, ++h_it)
        glu_chunk(p, sm, h_ring, h_slot_bytes, tmem_base, j, c1_it, h_it, gquad, ggroup, lane, parts);
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
  a.h_tmem = (tc::g_mlp_flags & 8) ? 0 : 1;
  // shared memory (<= 225 KB): 1 KB barriers + A ring 64 KB + {H in TMEM: 32 KB staging | 2 H slots} + weight ring
  //   H in TMEM : bf16x3 4 x 32 KB, bf16 8 x 16 KB        H in smem : bf16x3 3 x 32 KB (H 64 KB), bf16 8 x 16 KB (H 32 KB)
  a.nb = a.parts == 2 ? (a.h_tmem ? 4 : 3) : 8;
  a.num_tiles = (q->M + 127) / 128;
  PFN_cuTensorMapEncodeTiled enc = tc::mlp_encode_fn();
  MPHSIR_REQUIRE(enc != nullptr, "mlp: cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)q->C, (cuuint64_t)q->M};
  cuuint64_t gstr[1] = {(cuuint64_t)q->ldx * 4};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  MPHSIR_REQUIRE(enc(&a.tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(q->X), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
  const size_t smem = 1024 + (size_t)tc::M_NA * tc::M_STAGE + (a.h_tmem ? (size_t)8 * tc::M_STG_FLOATS * 4 : 2 * (size_t)tc::M_SLAB * a.parts) +
                      (size_t)a.nb * tc::M_SLAB * a.parts;
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
