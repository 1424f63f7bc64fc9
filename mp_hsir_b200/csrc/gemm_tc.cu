// tcgen05 GEMM engine (sm_100a): Y = epilogue( prologue(A) @ W^T ) with fp32 activations in HBM.
//
//   * A (fp32, token-major) is read by 8 "converter" warps, LayerNorm-ed if requested, split into bf16
//     hi (+ lo) parts and written into 128-byte-swizzled K-major shared-memory slabs (128 rows x 64 k).
//   * W lives in HBM as a pre-swizzled bf16 hi/lo image ("Bimg", see bimg_offset) so that one
//     cp.async.bulk (TMA bulk copy, mbarrier complete_tx) moves a 128-row x 64-k block straight into
//     its shared-memory slot.
//   * one thread issues tcgen05.mma (kind::f16, M=128, N<=128 per instruction, fp32 accumulators in
//     TMEM).  precision 1 ("bf16x3"): hi*hi + hi*lo + lo*hi -> ~2^-16 relative product error, which keeps the
//     whole network inside the fp32 parity contract (1e-4); precision 2 ("bf16x1"): hi*hi only.
//   * 4 epilogue warps drain TMEM with tcgen05.ld and apply bias / residual / GLU / spectral-gate
//     epilogues, writing fp32 rows.
//
// Persistent CTAs (one per SM), warp-specialised, three mbarrier pipelines:
//   A slabs (converters -> MMA), B blocks (bulk copy -> MMA), accumulators (MMA -> epilogue), so the
//   conversion of tile i+1, the MMAs of tile i and the epilogue of tile i-1 overlap.
// Loop nest per 128-row tile:  pass (<=256 output columns, double-buffered in TMEM)
//                                -> k-slab (64) -> n-tile (128 columns) -> 4 x {1|3} MMAs (K=16 each)
// If all k-slabs of a tile fit the A ring (K <= 256) the tile is converted once and reused by every pass.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace mphsir {
namespace tc {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
// warp 0: B loader, warp 1: MMA + TMEM owner, warps 2-9: epilogue, warps 10-17: converters, warp 18: A loader
constexpr int kThreads = 608;
constexpr int kConvThreads = 256;
constexpr int kEpiWarps = 8;
constexpr int kFirstConvWarp = 10;
constexpr int kLoaderWarp = 18;
constexpr int STAGE_BYTES = 128 * 64 * 4; // an A ring slot: fp32 TMA landing zone (32 KB), converted IN PLACE to bf16 hi|lo
constexpr int SLAB_BYTES = 128 * 128;   // one part (hi or lo) of a 128-row x 64-k A slab
constexpr int BN = 128;                 // columns per B block / per MMA instruction (N=128: A and B smem reads balance)
constexpr int BBLK_BYTES = BN * 128;    // one part of a 128-row x 64-k B block
constexpr int PASS_COLS = 256;          // accumulator columns per pass (x2 buffers = 512 TMEM columns)
constexpr int STG_LD = 32;              // staging rows are dense; 16-byte chunks are XOR-swizzled by (row & 7)
constexpr int STG_FLOATS = 32 * STG_LD;
constexpr int MAX_RING = 8;

// optional in-kernel cycle accounting (one lane per role; p.dbg == NULL in production)
#define TC_T0() (p.dbg ? clock64() : 0)
#define TC_ACC(var, t0) do { if (p.dbg) var += clock64() - (t0); } while (0)

// p.rev: the launch walks its work items from the last to the first.  Activations at 512 x 512 are slightly larger than the
// L2 (134 MB per 128-channel tensor against 126 MB): a consumer that starts where its producer STOPPED finds the most
// recently written part still on chip, while two kernels walking in the same direction evict everything just before it is read.
__device__ __forceinline__ int tc_items(const TcArgs& p) { return p.n_full + (p.num_tiles - p.n_full) * p.psplit; }
__device__ __forceinline__ int tc_order(const TcArgs& p, int vt) {
  const int n = tc_items(p);
  return (p.rev && vt < n) ? n - 1 - vt : vt;
}

// Work item `vt` of a launch = (row tile, range of 256-column passes [p0, p1)).  The first n_full items are whole row tiles; the
// remaining tiles are cut into `psplit` pass groups each (item = n_full + (tile - n_full) * psplit + group): either ALL tiles of a
// few-tile launch (n_full = 0), or the tiles of the last, partly filled round of a many-tile launch, so that the round costs
// a fraction of a tile time instead of a whole one.
__device__ __forceinline__ void tc_decode(const TcArgs& p, int npass, int vt, int& tile, int& p0, int& p1) {
  const int w = tc_order(p, vt);
  if (w < p.n_full) {
    tile = w;
    p0 = 0;
    p1 = npass;
  } else {
    const int idx = w - p.n_full;
    const int t = idx / p.psplit;
    tile = p.n_full + t;
    p0 = (idx - t * p.psplit) * p.ppg;
    p1 = min(npass, p0 + p.ppg);
  }
}
#define TC_WORK_ITEM(vt)                 \
  int tile, p0, p1;                      \
  tc_decode(p, npass, (vt), tile, p0, p1); \
  (void)p0; (void)p1

struct Smem {
  // barriers first (8-byte aligned), rings after (1024-byte aligned, carved dynamically)
  uint64_t a_full[MAX_RING], a_empty[MAX_RING], b_full[MAX_RING], b_empty[MAX_RING];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t a_peer[MAX_RING], b_peer[MAX_RING];  // CTA pairs: the peer CTA's slab / half weight block is ready (leader side)
  uint64_t stage_full[MAX_RING];  // TMA landed the fp32 slab (a_full: converted to bf16; a_empty: MMAs done)
  uint64_t res_full[2 * 8];       // tepi: residual box landed in slot k of epilogue warp w (index 2*w + k)
  uint32_t tmem_base;
};

__device__ __forceinline__ void tile_rows(const TcArgs& p, int tile, int& m0, int& m_end) {
  if (p.tiles_per_batch > 0) {
    const int b = tile / p.tiles_per_batch;
    m0 = b * p.rows_per_batch + (tile - b * p.tiles_per_batch) * 128;
    m_end = min((b + 1) * p.rows_per_batch, p.M);  // dummy tiles of a clustered launch lie beyond M
  } else {
    m0 = tile * 128;
    m_end = p.M;
  }
}

// ------------------------------------------------------------------------------------------------
// TMA epilogue (BIAS / RESIDUAL / PROJ): the register-staged epilogue below is bound by the latency of its own
// residual loads and by store issue (ncu / role counters: 2.2k clk per 32x32 chunk for BIAS, 5.9k with a residual,
// i.e. 4 TB/s resp. 1.5 TB/s of output at best).  Here every epilogue warp owns two 4 KB boxes [32 rows x 32 cols],
// 128-byte swizzled: the residual of chunk i+1 is TMA-loaded into one box while chunk i is combined IN PLACE in the
// other (row-per-thread straight out of tcgen05.ld, conflict-free through the swizzle) and leaves by a TMA store.
// No global load ever stalls a warp, stores cost one instruction per chunk, and row/column tails are clipped by the
// tensor maps (3-D: columns, rows inside the sample, sample).
// ------------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_tma(const TcArgs& p, Smem* sm, uint8_t* eslots, uint32_t tmem_base, int num_tiles,
                                             int npass, int warp, int lane, int crank) {
  constexpr bool RES = EPI == MPHSIR_EPI_RESIDUAL, PROJ = EPI == MPHSIR_EPI_PROJ;
  const int quad = warp & 3, half = (warp - 2) >> 2, ew = warp - 2;
  // p.ebox boxes of 4 KB per warp: 2 (default), or 1 for BIAS epilogues of K > 64 GEMMs — the 32 KB saved buy a fourth
  // A-ring slot, i.e. a whole K = 128 tile of load/convert lookahead (role counters: the converters starved on TMA
  // latency with one slab of lookahead while the epilogue warps idled half of the time)
  uint8_t* const slot_base = eslots + (size_t)ew * 4096 * p.ebox;   // box k at slot_base + 4096 * k
  const uint32_t slot_a0 = smem_u32(slot_base);
  const uint32_t rbar0 = smem_u32(&sm->res_full[2 * ew]);          // barrier k at rbar0 + 8 * k
  uint32_t rph = 0u;                                                // bit k: phase of barrier k

  const int pc = p.pass_cols;   // accumulator columns per pass: 256, or 128 when a 256-column few-tile GEMM is cut in two
  auto ncols_of = [&](int pass) { return min(pc, p.Np - pass * pc); };
  // next chunk of this warp after (tile, tit, pass, c0); false when the CTA's work is finished
  auto p0_of = [&](int vt) { int t_, a_, b_; tc_decode(p, npass, vt, t_, a_, b_); return a_; };
  auto p1_of = [&](int vt) { int t_, a_, b_; tc_decode(p, npass, vt, t_, a_, b_); return b_; };
  auto tile_of = [&](int vt) { int t_, a_, b_; tc_decode(p, npass, vt, t_, a_, b_); return t_; };
  auto advance = [&](int& vt, int& tit, int& pass, int& c0) -> bool {
    c0 += 64;
    for (;;) {
      if (c0 < ncols_of(pass)) return true;
      ++pass;
      c0 = half * 32;
      if (pass >= p1_of(vt)) {
        vt += gridDim.x;
        ++tit;
        if (tit >= p.iters || vt - crank >= num_tiles) return false;
        pass = p0_of(vt);
      }
    }
  };
  auto needs_res = [&](int pass, int c0) { return RES || (PROJ && pass * pc + c0 < p.n_split); };
  auto tile_coord = [&](int tile, int& b, int& row0) {
    if (p.tiles_per_batch > 0) {
      b = tile / p.tiles_per_batch;
      row0 = (tile - b * p.tiles_per_batch) * 128 + quad * 32;
    } else {
      b = 0;
      row0 = tile * 128 + quad * 32;
    }
  };
  auto issue_res = [&](int vt, int pass, int c0, int slot) {  // lane 0 only
    int b, row0;
    tile_coord(tile_of(vt), b, row0);
    mbar_expect_tx(rbar0 + 8 * slot, 4096);
    tma_load_3d(slot_a0 + 4096 * slot, &p.tmR, pass * pc + c0, row0, b, rbar0 + 8 * slot);
  };

  uint32_t ck = 0;  // chunks processed by this warp (slot = ck & 1)
  {
    int t = blockIdx.x, ti = 0, ps = p0_of(blockIdx.x), c = half * 32 - 64;
    if (num_tiles > (int)blockIdx.x - crank && p.iters > 0 && advance(t, ti, ps, c) && needs_res(ps, c) && lane == 0) issue_res(t, ps, c, 0);
  }
  uint32_t acc_it = 0;
  long long t_wait = 0, t_tmem = 0, t_all0 = TC_T0();
  for (int vt = blockIdx.x, tit = 0; tit < p.iters && vt - crank < num_tiles; vt += gridDim.x, ++tit) {
    TC_WORK_ITEM(vt);
    int m0, m_end;
    tile_rows(p, tile, m0, m_end);
    const int mm = min(m0 + quad * 32 + lane, m_end - 1);  // this thread's row (clamped: tails are clipped by TMA)
    const float scl = p.row_scale != nullptr ? __ldg(p.row_scale + mm / p.rows_per_batch) : 1.f;
    const float* grow = nullptr;
    if (PROJ) {
      const int hw = p.H * p.W;
      const int b = mm / hw, rem = mm - b * hw;
      const int y = rem / p.W, x = rem - y * p.W;
      int ys = y - p.shift, xs = x - p.shift;
      if (ys < 0) ys += p.H;
      if (xs < 0) xs += p.W;
      grow = p.gate + (size_t)(b * (hw >> 6) + (ys >> 3) * (p.W >> 3) + (xs >> 3)) * p.n_split;
    }
    int tb, trow0;
    tile_coord(tile, tb, trow0);
    for (int pass = p0; pass < p1; ++pass, ++acc_it) {
      const int buf = acc_it & 1;
      long long tw = TC_T0();
      mbar_wait(smem_u32(&sm->acc_full[buf]), (acc_it >> 1) & 1);
      TC_ACC(t_wait, tw);
      tc_fence_after();
      const int ncols_pass = ncols_of(pass);
      for (int c0 = half * 32; c0 < ncols_pass; c0 += 64, ++ck) {
        const int s = p.ebox == 2 ? (ck & 1) : 0;
        const int n0 = pass * pc + c0;
        const bool with_res = needs_res(pass, c0);
        const bool left = PROJ && n0 < p.n_split;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * PASS_COLS + c0;
        uint32_t r[16];
        tw = TC_T0();
        tmem_ld16_nowait(taddr, r);
        // This chunk overwrites box s, last stored from two chunks ago: ONE store (the previous chunk's, from box s^1) may
        // still be reading.  Only when the next chunk needs its residual TMA-loaded into box s^1 must that store have
        // finished too — waiting for it unconditionally serialised every chunk behind the previous chunk's store.
        if (lane == 0) {
          int t2 = vt, ti2 = tit, ps2 = pass, c2 = c0;
          const bool next_res = advance(t2, ti2, ps2, c2) && needs_res(ps2, c2);
          if (ck > 0) {
            if (next_res || p.ebox == 1) bulk_wait_group_read<0>();
            else bulk_wait_group_read<1>();
          }
          if (next_res) issue_res(t2, ps2, c2, s ^ 1);
        }
        if (with_res) {
          mbar_wait(rbar0 + 8 * s, (rph >> s) & 1u);
          rph ^= 1u << s;
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 bias4[4], g4[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int n = n0 + 16 * h + 4 * qq;
            bias4[qq] = (p.bias != nullptr && n < p.N) ? ldg4(p.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (PROJ) g4[qq] = (left && n < p.n_split) ? ldg4(grow + n) : make_float4(1.f, 1.f, 1.f, 1.f);
          }
          tmem_ld_wait();
          TC_ACC(t_tmem, tw);
          float4 v[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            v[qq] = make_float4(__uint_as_float(r[4 * qq]) + bias4[qq].x, __uint_as_float(r[4 * qq + 1]) + bias4[qq].y,
                                __uint_as_float(r[4 * qq + 2]) + bias4[qq].z, __uint_as_float(r[4 * qq + 3]) + bias4[qq].w);
          if (h == 0) tmem_ld16_nowait(taddr + 16, r);  // second half in flight while the first is combined
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int q = 4 * h + qq;
            float4* cell = reinterpret_cast<float4*>(slot_base + 4096 * s + lane * 128 + ((q ^ (lane & 7)) << 4));
            float4 o = v[qq];
            if (with_res) {
              const float4 x = *cell;
              if (PROJ) {
                o.x = x.x + scl * (v[qq].x * g4[qq].x); o.y = x.y + scl * (v[qq].y * g4[qq].y);
                o.z = x.z + scl * (v[qq].z * g4[qq].z); o.w = x.w + scl * (v[qq].w * g4[qq].w);
              } else {
                o.x = x.x + scl * v[qq].x; o.y = x.y + scl * v[qq].y;
                o.z = x.z + scl * v[qq].z; o.w = x.w + scl * v[qq].w;
              }
            }
            *cell = o;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (PROJ && !left) tma_store_3d(&p.tmY2, slot_a0 + 4096 * s, n0 - p.n_split, trow0, tb);
          else tma_store_3d(&p.tmY, slot_a0 + 4096 * s, n0, trow0, tb);
          bulk_commit_group();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (p.cluster > 1) mbar_arrive_remote_relaxed(mapa_cluster(smem_u32(&sm->acc_empty[buf]), 0));   // TMEM hand-off to the leader's MMA warp
        else mbar_arrive(smem_u32(&sm->acc_empty[buf]));
      }
    }
  }
  if (lane == 0) bulk_wait_group_read<0>();  // shared memory stays allocated until the last store has read its box
  if (p.dbg && warp == 2 && lane == 0) {
    p.dbg[blockIdx.x * 16 + 6] = clock64() - t_all0;
    p.dbg[blockIdx.x * 16 + 7] = t_wait;
    p.dbg[blockIdx.x * 16 + 8] = t_tmem;
  }
}

template <int EPI, bool LN, int CG>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ TcArgs p) {
  constexpr bool CONV = EPI >= TC_OUT_TOKENS;
  // NCHW / pixel-unshuffle stores are already coalesced (or hopeless) in the row-per-thread TMEM mapping
  constexpr bool DIRECT = (EPI == TC_OUT_NCHW_RES || EPI == TC_OUT_UNSHUFFLE);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_launch_dependents();
  Smem* sm = reinterpret_cast<Smem*>(smem_raw);
  const int parts = p.parts;                       // 1 (bf16x1) or 2 (bf16x3)
  const int a_slot_bytes = STAGE_BYTES;
  const int b_slot_bytes = (BBLK_BYTES / p.cluster) * parts;
  uint8_t* a_ring = smem_raw + 1024;
  uint8_t* b_ring = a_ring + (size_t)p.na * a_slot_bytes;
  float* staging = reinterpret_cast<float*>(b_ring + (size_t)p.nb * b_slot_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ks = p.ks;
  const int NT = (p.Np + BN - 1) / BN;             // 64-column n-tiles
  const int pc = p.pass_cols;                      // accumulator columns per pass (256, or 128: see plan_work)
  const int TPP = pc / BN;                         // n-tiles per pass
  const int npass = (NT + TPP - 1) / TPP;
  const bool stationary = Ks <= p.na;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_RING; ++i) {
      mbar_init(smem_u32(&sm->a_full[i]), kConvThreads / 32);
      mbar_init(smem_u32(&sm->a_empty[i]), 1);
      mbar_init(smem_u32(&sm->b_full[i]), 1);
      mbar_init(smem_u32(&sm->b_empty[i]), 1);
      mbar_init(smem_u32(&sm->a_peer[i]), 1);
      mbar_init(smem_u32(&sm->b_peer[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sm->acc_full[i]), 1);
      mbar_init(smem_u32(&sm->acc_empty[i]), kEpiWarps * p.cluster);  // CTA pairs: both CTAs' epilogue warps arrive at the leader
    }
    for (int i = 0; i < MAX_RING; ++i) mbar_init(smem_u32(&sm->stage_full[i]), 1);
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(smem_u32(&sm->res_full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CG == 2) {
    cluster_sync_all();  // the peer's mbarriers exist before a remote arrive / multicast commit targets them
    if (warp == 1) tmem_alloc2(smem_u32(&sm->tmem_base), 512);
    tc_fence_before();
    cluster_sync_all();
  } else {
    if (warp == 1) tmem_alloc(smem_u32(&sm->tmem_base), 512);
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_after();
  pdl_wait();  // everything above overlapped the predecessor's tail; from here on its results are read / its inputs overwritten
  const uint32_t tmem_base = sm->tmem_base;
  const int num_tiles = tc_items(p);   // work items (whole row tiles + pass groups of the split tiles)
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0;
  constexpr bool pair = CG == 2;   // CTA pair: M = 256 MMAs (cta_group::2) issued by the leader (rank 0) for both tiles

  if (warp == 0) {
    // =============================== B loader ===============================================
    {  // warp-uniform loop, one elected lane issues the bulk copies
      uint32_t it = 0;
      long long t_wait = 0, t_all0 = TC_T0();
      for (int vt = blockIdx.x, tit = 0; tit < p.iters; vt += gridDim.x, ++tit) {
      if (vt - (int)crank >= num_tiles) break;  // a pair stops together; the peer may run ONE all-out-of-bounds tile (odd tile count)
      TC_WORK_ITEM(vt);
        const uint8_t* bimg = reinterpret_cast<const uint8_t*>(p.Bimg);
        if (p.tiles_per_batch > 0) bimg += (size_t)(tile / p.tiles_per_batch) * p.b_batch_bytes;
        for (int pass = p0; pass < p1; ++pass) {
          for (int s = 0; s < Ks; ++s) {
            for (int j = TPP * pass; j < min(NT, TPP * pass + TPP); ++j, ++it) {
              const int slot = it % p.nb;
              const long long tw = TC_T0();
              mbar_wait(smem_u32(&sm->b_empty[slot]), ((it / p.nb) & 1) ^ 1);
              TC_ACC(t_wait, tw);
              // a CTA pair stages HALF of the block's rows (= of the MMA's N) per CTA
              const int rows = min(BN, p.Np - j * BN) / p.cluster;
              const uint32_t bytes = rows * 128;
              const uint32_t full = smem_u32(&sm->b_full[slot]);
              if (elect_one()) {
                mbar_expect_tx(full, bytes * parts);
                for (int part = 0; part < parts; ++part) {
                  const uint8_t* src = bimg + ((size_t)(part * Ks + s) * p.Np + j * BN + crank * rows) * 128;
                  bulk_g2s(smem_u32(b_ring + (size_t)slot * b_slot_bytes + part * (BBLK_BYTES / p.cluster)), src, bytes, full);
                }
              }
              __syncwarp();
            }
          }
        }
      }
      if (p.dbg && lane == 0) {
        p.dbg[blockIdx.x * 16 + 0] = clock64() - t_all0;
        p.dbg[blockIdx.x * 16 + 1] = t_wait;
      }
    }
  } else if (CG == 2 && warp == 1 && crank != 0) {
    // =============================== peer CTA of a pair: relay ==============================
    // Only the leader issues MMAs.  This warp walks the leader's wait sequence on THIS CTA's barriers and forwards every
    // completion (converted A slab, landed half weight block) to the leader's a_peer / b_peer barriers.  Relaxed arrives:
    // the data stays in this CTA's shared memory and is read there by this SM's tensor core.
    uint32_t a_it = 0, b_it = 0;
    for (int vt = blockIdx.x, tit = 0; tit < p.iters; vt += gridDim.x, ++tit) {
      if (vt - (int)crank >= num_tiles) break;
      TC_WORK_ITEM(vt);
      const uint32_t a_base = a_it;
      for (int pass = p0; pass < p1; ++pass) {
        for (int s = 0; s < Ks; ++s) {
          if (!stationary || pass == p0) {
            const uint32_t ai = stationary ? a_base + s : a_it;
            const uint32_t a_slot = ai % p.na;
            mbar_wait(smem_u32(&sm->a_full[a_slot]), (ai / p.na) & 1);
            if (lane == 0) mbar_arrive_remote_relaxed(mapa_cluster(smem_u32(&sm->a_peer[a_slot]), 0));
            __syncwarp();
          }
          for (int j = TPP * pass; j < min(NT, TPP * pass + TPP); ++j, ++b_it) {
            const int b_slot = b_it % p.nb;
            mbar_wait(smem_u32(&sm->b_full[b_slot]), (b_it / p.nb) & 1);
            if (lane == 0) mbar_arrive_remote_relaxed(mapa_cluster(smem_u32(&sm->b_peer[b_slot]), 0));
            __syncwarp();
          }
          if (!stationary) ++a_it;
        }
      }
      if (stationary) a_it += Ks;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =============================================
    // The whole warp walks the loop nest (waits are warp-uniform); one elected lane issues the MMAs and commits.
    {
      uint32_t a_it = 0, b_it = 0, acc_it = 0;
      long long t_acc = 0, t_a = 0, t_b = 0, t_issue = 0, t_commit = 0, t_all0 = TC_T0();
      for (int vt = blockIdx.x, tit = 0; tit < p.iters; vt += gridDim.x, ++tit) {
      if (vt - (int)crank >= num_tiles) break;  // a pair stops together; the peer may run ONE all-out-of-bounds tile (odd tile count)
      TC_WORK_ITEM(vt);
        const uint32_t a_base = a_it;
        for (int pass = p0; pass < p1; ++pass, ++acc_it) {
          const int buf = acc_it & 1;
          long long tw = TC_T0();
          if (pair) mbar_wait_cluster(smem_u32(&sm->acc_empty[buf]), ((acc_it >> 1) & 1) ^ 1);
          else mbar_wait(smem_u32(&sm->acc_empty[buf]), ((acc_it >> 1) & 1) ^ 1);
          TC_ACC(t_acc, tw);
          tc_fence_after();
          for (int s = 0; s < Ks; ++s) {
            uint32_t a_slot;
            tw = TC_T0();
            if (stationary) {
              a_slot = (a_base + s) % p.na;
              if (pass == p0) {
                mbar_wait(smem_u32(&sm->a_full[a_slot]), ((a_base + s) / p.na) & 1);
                if (pair) mbar_wait_cluster(smem_u32(&sm->a_peer[a_slot]), ((a_base + s) / p.na) & 1);
              }
            } else {
              a_slot = a_it % p.na;
              mbar_wait(smem_u32(&sm->a_full[a_slot]), (a_it / p.na) & 1);
              if (pair) mbar_wait_cluster(smem_u32(&sm->a_peer[a_slot]), (a_it / p.na) & 1);
            }
            TC_ACC(t_a, tw);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_ring + (size_t)a_slot * a_slot_bytes);
            for (int j = TPP * pass; j < min(NT, TPP * pass + TPP); ++j, ++b_it) {
              const int b_slot = b_it % p.nb;
              tw = TC_T0();
              mbar_wait(smem_u32(&sm->b_full[b_slot]), (b_it / p.nb) & 1);
              if (pair) mbar_wait_cluster(smem_u32(&sm->b_peer[b_slot]), (b_it / p.nb) & 1);
              TC_ACC(t_b, tw);
              tc_fence_after();
              const uint32_t b_addr = smem_u32(b_ring + (size_t)b_slot * b_slot_bytes);
              const int ncols = min(BN, p.Np - j * BN);
              const uint32_t idesc = pair ? make_idesc2(ncols) : make_idesc(ncols);
              const uint32_t d_addr = tmem_base + buf * PASS_COLS + (j - TPP * pass) * BN;
              tw = TC_T0();
              // descriptors of consecutive k16 steps differ by 32 bytes (+2 in the 16-byte address field)
              const uint64_t ah0 = make_desc(a_addr), bh0 = make_desc(b_addr);
              const uint64_t al0 = make_desc(a_addr + SLAB_BYTES), bl0 = make_desc(b_addr + BBLK_BYTES / p.cluster);
              if (elect_one()) {
                if (pair) {   // M = 256 over both CTAs (109 clk per instruction against 121 clk per M = 128: tools/mma_rate.cu)
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    umma2_bf16(d_addr, ah0 + 2 * k, bh0 + 2 * k, idesc, (s | k) != 0);
                    if (parts == 2) {
                      umma2_bf16(d_addr, ah0 + 2 * k, bl0 + 2 * k, idesc, 1);
                      umma2_bf16(d_addr, al0 + 2 * k, bh0 + 2 * k, idesc, 1);
                    }
                  }
                  umma2_commit(smem_u32(&sm->b_empty[b_slot]));
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    umma_bf16(d_addr, ah0 + 2 * k, bh0 + 2 * k, idesc, (s | k) != 0);
                    if (parts == 2) {
                      umma_bf16(d_addr, ah0 + 2 * k, bl0 + 2 * k, idesc, 1);
                      umma_bf16(d_addr, al0 + 2 * k, bh0 + 2 * k, idesc, 1);
                    }
                  }
                  umma_commit(smem_u32(&sm->b_empty[b_slot]));
                }
              }
              __syncwarp();
              TC_ACC(t_issue, tw);
            }
            const bool last_use = stationary ? (pass == p1 - 1) : true;
            if (last_use && elect_one()) {
              if (pair) umma2_commit(smem_u32(&sm->a_empty[a_slot]));
              else umma_commit(smem_u32(&sm->a_empty[a_slot]));
            }
            __syncwarp();
            if (!stationary) ++a_it;
          }
          if (elect_one()) {
            if (pair) umma2_commit(smem_u32(&sm->acc_full[buf]));
            else umma_commit(smem_u32(&sm->acc_full[buf]));
          }
          __syncwarp();
        }
        if (stationary) a_it += Ks;
      }
      if (p.dbg && lane == 0) {
        p.dbg[blockIdx.x * 16 + 2] = clock64() - t_all0;
        p.dbg[blockIdx.x * 16 + 3] = t_acc;
        p.dbg[blockIdx.x * 16 + 4] = t_a;
        p.dbg[blockIdx.x * 16 + 5] = t_b;
        p.dbg[blockIdx.x * 16 + 12] = t_issue;
        p.dbg[blockIdx.x * 16 + 13] = t_commit;
      }
    }
  } else if (warp < kFirstConvWarp) {
    // =============================== epilogue (warps 2..9) ==================================
    // TMEM lane quadrant = warp % 4 (hardware rule); the two warps of a quadrant alternate 32-column chunks.
    // Staged path: the warp's 32x32 fp32 block goes TMEM -> registers (row per thread) -> padded smem ->
    // registers in a (4 rows x 8 float4) mapping, so every global access is a full 128-byte line segment.
    constexpr bool TEPI_OK = (EPI == MPHSIR_EPI_BIAS || EPI == MPHSIR_EPI_RESIDUAL || EPI == MPHSIR_EPI_PROJ);
    if (TEPI_OK && p.tepi) {
      epilogue_tma<EPI>(p, sm, reinterpret_cast<uint8_t*>(staging), tmem_base, num_tiles, npass, warp, lane, (int)crank);
    } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stg = staging + (warp - 2) * STG_FLOATS;
    const int c4 = lane & 7, rsub = lane >> 3;
    uint32_t acc_it = 0;
    long long t_wait = 0, t_tmem = 0, t_all0 = TC_T0();
    for (int vt = blockIdx.x, tit = 0; tit < p.iters; vt += gridDim.x, ++tit) {
      if (vt - (int)crank >= num_tiles) break;  // a pair stops together; the peer may run ONE all-out-of-bounds tile (odd tile count)
      TC_WORK_ITEM(vt);
      int m0, m_end;
      tile_rows(p, tile, m0, m_end);
      const int mrow0 = m0 + quad * 32;
      // ---- per-row state ----
      // DIRECT: one row per thread (row = lane).  Staged: 8 rows per thread (row = it*4 + rsub).
      float scale_d = 1.f;
      int ib = 0, iy = 0, ix = 0;
      const bool row_ok_d = mrow0 + lane < m_end;
      if (DIRECT) {
        const int mm = row_ok_d ? mrow0 + lane : m0;
        const int hw = p.H * p.W;
        ib = mm / hw;
        const int rem = mm - ib * hw;
        iy = rem / p.W;
        ix = rem - iy * p.W;
      }
      uint32_t win[8];
      float scl[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        win[it] = 0;
        scl[it] = 1.f;
        const int m = mrow0 + it * 4 + rsub;
        if (!DIRECT && m < m_end) {
          if (p.row_scale != nullptr) scl[it] = __ldg(p.row_scale + m / p.rows_per_batch);
          if (EPI == MPHSIR_EPI_SPECTRAL || EPI == MPHSIR_EPI_PROJ) {
            const int hw = p.H * p.W;
            const int b = m / hw;
            const int rem = m - b * hw;
            const int y = rem / p.W, x = rem - y * p.W;
            int ys = y - p.shift, xs = x - p.shift;
            if (ys < 0) ys += p.H;
            if (xs < 0) xs += p.W;
            win[it] = b * (hw >> 6) + (ys >> 3) * (p.W >> 3) + (xs >> 3);
          }
        }
      }
      (void)scale_d;

      for (int pass = p0; pass < p1; ++pass, ++acc_it) {
        const int buf = acc_it & 1;
        long long tw = TC_T0();
        mbar_wait(smem_u32(&sm->acc_full[buf]), (acc_it >> 1) & 1);
        TC_ACC(t_wait, tw);
        tc_fence_after();
        const int ncols_pass = min(pc, p.Np - pass * pc);
        // software-pipelined bias: the float4 for the NEXT chunk is requested while this one is processed
        float4 bias_nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!DIRECT && p.bias != nullptr && pass * pc + half * 32 + 4 * c4 < p.N)
          bias_nxt = ldg4(p.bias + pass * pc + half * 32 + 4 * c4);
        for (int c0 = half * 32; c0 < ncols_pass; c0 += 64) {
          const int n0 = pass * pc + c0;
          uint32_t r[32];
          tw = TC_T0();
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * PASS_COLS + c0)
              : "memory");
          if (DIRECT) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (EPI == TC_OUT_UNSHUFFLE) {
              float* dst = p.Y + ((size_t)(ib * (p.H >> 1) + (iy >> 1)) * (p.W >> 1) + (ix >> 1)) * p.ldy + 2 * (iy & 1) + (ix & 1);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (row_ok_d && n0 + j < p.N) dst[(size_t)(n0 + j) * 4] = __uint_as_float(r[j]);
            } else {  // TC_OUT_NCHW_RES: lanes are 32 consecutive pixels of one channel -> 128-byte lines
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = n0 + j;
                if (row_ok_d && c < p.N) {
                  const size_t idx = ((size_t)(ib * p.N + c) * p.H + iy) * p.W + ix;
                  p.Y[idx] = __uint_as_float(r[j]) + __ldg(p.R + idx);
                }
              }
            }
            continue;
          }
          // ---- staged path ----
          const int n = n0 + 4 * c4;          // this thread's 4 output columns (packed order)
          const bool col_ok = n < p.N;
          const bool left = n0 < p.n_split;  // EPI_PROJ: warp-uniform side of the column split (n_split % 32 == 0)
          const float4 bias4 = bias_nxt;
          if (p.bias != nullptr && c0 + 64 < ncols_pass && n + 64 < p.N) bias_nxt = ldg4(p.bias + n + 64);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          TC_ACC(t_tmem, tw);
          __syncwarp();  // previous chunk's smem reads are done
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(stg + lane * STG_LD + 4 * (q ^ (lane & 7))) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {  // two batches of 4 rows-groups: loads first, then math + stores
            float4 acc[4], x1[4], x2[4], x3[4];
            bool ok[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int it = hb * 4 + i;
              const int row = it * 4 + rsub;
              const int m = mrow0 + row;
              ok[i] = col_ok && m < m_end;
              acc[i] = *reinterpret_cast<const float4*>(stg + row * STG_LD + 4 * (c4 ^ (row & 7)));
              x1[i] = x2[i] = x3[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ok[i]) {
                if (EPI == MPHSIR_EPI_RESIDUAL || EPI == MPHSIR_EPI_SPECTRAL) x1[i] = ldg4(p.res1 + (size_t)m * p.ldr1 + n);
                if (EPI == MPHSIR_EPI_RESIDUAL && p.res2 != nullptr) x2[i] = ldg4(p.res2 + (size_t)m * p.ldr2 + n);
                if (EPI == MPHSIR_EPI_SPECTRAL) {
                  x2[i] = ldg4(p.gsrc + (size_t)m * p.ldg + n);
                  x3[i] = ldg4(p.gate + (size_t)win[it] * p.N + n);
                }
                if (EPI == MPHSIR_EPI_PROJ && left) {
                  x1[i] = ldg4(p.res1 + (size_t)m * p.ldr1 + n);
                  x3[i] = ldg4(p.gate + (size_t)win[it] * p.n_split + n);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (!ok[i]) continue;
              const int it = hb * 4 + i;
              const int m = mrow0 + it * 4 + rsub;
              float4 v = make_float4(acc[i].x + bias4.x, acc[i].y + bias4.y, acc[i].z + bias4.z, acc[i].w + bias4.w);
              if (EPI == MPHSIR_EPI_BIAS || EPI == TC_OUT_TOKENS) {
                *reinterpret_cast<float4*>(p.Y + (size_t)m * p.ldy + n) = v;
              } else if (EPI == MPHSIR_EPI_RESIDUAL) {
                const float sc = scl[it];
                float4 o;
                o.x = x1[i].x + sc * v.x + x2[i].x;
                o.y = x1[i].y + sc * v.y + x2[i].y;
                o.z = x1[i].z + sc * v.z + x2[i].z;
                o.w = x1[i].w + sc * v.w + x2[i].w;
                *reinterpret_cast<float4*>(p.Y + (size_t)m * p.ldy + n) = o;
              } else if (EPI == MPHSIR_EPI_GLU) {
                // packed columns (2j, 2j+1) = (value_j, gate_j)
                const float2 o = make_float2(v.x * gelu_erf_fast(v.y), v.z * gelu_erf_fast(v.w));
                *reinterpret_cast<float2*>(p.Y + (size_t)m * p.ldy + (n >> 1)) = o;
              } else if (EPI == MPHSIR_EPI_SPECTRAL) {
                const float sc = scl[it];
                float4 o;
                o.x = x1[i].x + sc * (x2[i].x * x3[i].x + v.x);
                o.y = x1[i].y + sc * (x2[i].y * x3[i].y + v.y);
                o.z = x1[i].z + sc * (x2[i].z * x3[i].z + v.z);
                o.w = x1[i].w + sc * (x2[i].w * x3[i].w + v.w);
                *reinterpret_cast<float4*>(p.Y + (size_t)m * p.ldy + n) = o;
              } else if (EPI == MPHSIR_EPI_PROJ) {
                if (left) {
                  const float sc = scl[it];
                  float4 o;
                  o.x = x1[i].x + sc * (v.x * x3[i].x);
                  o.y = x1[i].y + sc * (v.y * x3[i].y);
                  o.z = x1[i].z + sc * (v.z * x3[i].z);
                  o.w = x1[i].w + sc * (v.w * x3[i].w);
                  *reinterpret_cast<float4*>(p.Y + (size_t)m * p.ldy + n) = o;
                } else {
                  *reinterpret_cast<float4*>(p.Y2 + (size_t)m * p.ldy2 + (n - p.n_split)) = v;
                }
              } else if (EPI == TC_OUT_SHUFFLE) {
                const int hw = p.H * p.W;
                const int b = m / hw;
                const int rem = m - b * hw;
                const int y = rem / p.W, x = rem - y * p.W;
                const int cn_total = p.N >> 2;
                const int qq = n / cn_total, cn = n - qq * cn_total;
                float* dst = p.Y + ((size_t)(b * 2 * p.H + 2 * y + (qq >> 1)) * (2 * p.W) + 2 * x + (qq & 1)) * p.ldy + cn;
                *reinterpret_cast<float4*>(dst) = v;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (pair) mbar_arrive_remote_relaxed(mapa_cluster(smem_u32(&sm->acc_empty[buf]), 0));
          else mbar_arrive(smem_u32(&sm->acc_empty[buf]));
        }
      }
    }
    if (p.dbg && warp == 2 && lane == 0) {
      p.dbg[blockIdx.x * 16 + 6] = clock64() - t_all0;
      p.dbg[blockIdx.x * 16 + 7] = t_wait;
      p.dbg[blockIdx.x * 16 + 8] = t_tmem;
    }
    }
  } else if (warp == kLoaderWarp) {
    // =============================== A loader (TMA bulk copies, warp 18) ====================
    // Streams raw fp32 slabs of A into the ring slots (TMA tensor-map boxes; per-row bulk copies as the general
    // fallback), landing on the slot's stage_full mbarrier.  It runs as far ahead as the ring allows (4 x 32 KB),
    // so the memory-level parallelism lives in shared memory instead of registers.
    uint32_t st_it = 0;
    for (int vt = blockIdx.x, tit = 0; tit < p.iters; vt += gridDim.x, ++tit) {
      if (vt - (int)crank >= num_tiles) break;  // a pair stops together; the peer may run ONE all-out-of-bounds tile (odd tile count)
      TC_WORK_ITEM(vt);
      int m0, m_end;
      tile_rows(p, tile, m0, m_end);
      const int conv_passes = stationary ? 1 : p1 - p0;
      for (int pass = 0; pass < conv_passes; ++pass) {
        for (int s = 0; s < Ks; ++s, ++st_it) {
          const int st = st_it % p.na;
          mbar_wait(smem_u32(&sm->a_empty[st]), ((st_it / p.na) & 1) ^ 1);
          const uint32_t full = smem_u32(&sm->stage_full[st]);
          uint8_t* dst = a_ring + (size_t)st * STAGE_BYTES;
          const int k0 = s * 64;
          const int kend = min(k0 + 64, p.Ka);
          if (p.a_mode == A_TMAP2D) {
            // one TMA box [128 rows x 64 floats]; rows beyond M and columns beyond K are zero-filled by the TMA unit
            if (lane == 0) {
              mbar_expect_tx(full, STAGE_BYTES);
              const int row = p.a_row_mod > 0 ? m0 % p.a_row_mod : m0;
              tma_load_2d(smem_u32(dst), &p.tmA, k0, row, full);
            }
            __syncwarp();
            continue;
          }
          if (p.a_mode == A_TMAP4D) {
            // conv: one box [by x bx pixels x seg channels] per (tap, channel segment); the zero padding of the
            // 3x3 conv is the TMA out-of-bounds fill (coordinates -1 / W / H)
            if (lane == 0) {
              const int hw = p.H * p.W;
              const int b = m0 / hw, rem = m0 - b * hw;
              const int y0 = rem / p.W, x0 = rem - y0 * p.W;
              const int nsub = (kend - k0) / p.seg;
              const uint32_t sub_bytes = (uint32_t)p.box_rows * p.seg * 4u;
              mbar_expect_tx(full, nsub * sub_bytes);
              for (int j = 0; j < nsub; ++j) {
                const int k = k0 + j * p.seg;
                const int tap = k / p.Cin, cc = k - tap * p.Cin;
                tma_load_4d(smem_u32(dst + j * sub_bytes), &p.tmA, cc, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b, full);
              }
            }
            __syncwarp();
            continue;
          }
          // pass 1: bytes this lane will copy (rows lane, lane+32, lane+64, lane+96)
          uint32_t bytes = 0;
          if (!CONV) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (m0 + lane + 32 * i < m_end) bytes += (uint32_t)(kend - k0) * 4u;
          } else {
            const int hw = p.H * p.W;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + lane + 32 * i;
              if (m >= m_end) continue;
              const int b = m / hw, rem = m - b * hw;
              const int y = rem / p.W, x = rem - y * p.W;
              for (int k = k0; k < kend;) {
                const int tap = k / p.Cin, cc = k - tap * p.Cin;
                const int len = min(p.Cin - cc, kend - k);
                const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
                if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) bytes += (uint32_t)len * 4u;
                k += len;
              }
            }
          }
          uint32_t total = bytes;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
          if (lane == 0) {
            if (total > 0) mbar_expect_tx(full, total);
            else mbar_arrive(full);
          }
          __syncwarp();
          // pass 2: issue the copies
          if (!CONV) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = lane + 32 * i, m = m0 + r;
              if (m < m_end) {
                const float* src = p.A + (size_t)(p.a_row_mod > 0 ? m % p.a_row_mod : m) * p.lda + k0;
                bulk_g2s(smem_u32(dst + r * 256), src, (uint32_t)(kend - k0) * 4u, full);
              }
            }
          } else {
            const int hw = p.H * p.W;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = lane + 32 * i, m = m0 + r;
              if (m >= m_end) continue;
              const int b = m / hw, rem = m - b * hw;
              const int y = rem / p.W, x = rem - y * p.W;
              for (int k = k0; k < kend;) {
                const int tap = k / p.Cin, cc = k - tap * p.Cin;
                const int len = min(p.Cin - cc, kend - k);
                const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
                if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                  const float* src = p.A + ((size_t)(b * p.H + yy) * p.W + xx) * p.lda + cc;
                  bulk_g2s(smem_u32(dst + r * 256 + (k - k0) * 4), src, (uint32_t)len * 4u, full);
                }
                k += len;
              }
            }
          }
        }
      }
    }
  } else {
    // =============================== A converters (warps 10..17) ============================
    // fp32 staging slab -> (LayerNorm) -> bf16 hi/lo split -> 128B-swizzled K-major MMA slab.
    const int ct = threadIdx.x - kFirstConvWarp * 32;  // 0..255
    const int chunk = ct & 7;              // 8-element (16-byte bf16) chunk inside the 64-k slab
    const int rbase = ct >> 3;             // rows rbase + 32*i, i = 0..3
    constexpr bool has_ln = LN;
    const bool ln_from_stage = has_ln && stationary;  // the whole row is resident in the ring
    long long t_slot = 0, t_ld = 0, t_all0 = TC_T0();
    uint32_t a_it = 0;
    for (int vt = blockIdx.x, tit = 0; tit < p.iters; vt += gridDim.x, ++tit) {
      if (vt - (int)crank >= num_tiles) break;  // a pair stops together; the peer may run ONE all-out-of-bounds tile (odd tile count)
      TC_WORK_ITEM(vt);
      int m0, m_end;
      tile_rows(p, tile, m0, m_end);
      float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f};
      if (has_ln) {
        float sm_[4] = {0.f, 0.f, 0.f, 0.f}, sq_[4] = {0.f, 0.f, 0.f, 0.f};
        if (ln_from_stage) {
          // the whole row (K <= 128) is resident in the staging slabs: statistics straight from smem
          for (int s = 0; s < Ks; ++s) {
            const int st = (a_it + s) % p.na;
            const long long tws = TC_T0();
            mbar_wait(smem_u32(&sm->stage_full[st]), ((a_it + s) / p.na) & 1);
            TC_ACC(t_ld, tws);
            const int k = s * 64 + chunk * 8;
            if (k < p.Ka) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = rbase + 32 * i;
                if (m0 + r < m_end) {
                  const float* sp = reinterpret_cast<const float*>(a_ring + (size_t)st * STAGE_BYTES + r * 256 + chunk * 32);
                  const float4 a = *reinterpret_cast<const float4*>(sp);
                  const float4 b = *reinterpret_cast<const float4*>(sp + 4);
                  sm_[i] += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
                  sq_[i] += ((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w)) +
                            ((b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w));
                }
              }
            }
          }
        } else {
          for (int k = chunk * 4; k < p.Ka; k += 32) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + rbase + 32 * i;
              v[i] = (m < m_end) ? ldg4(p.A + (size_t)(p.a_row_mod > 0 ? m % p.a_row_mod : m) * p.lda + k)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              sm_[i] += (v[i].x + v[i].y) + (v[i].z + v[i].w);
              sq_[i] += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            sm_[i] += __shfl_xor_sync(0xffffffffu, sm_[i], o);
            sq_[i] += __shfl_xor_sync(0xffffffffu, sq_[i], o);
          }
          const float mu = sm_[i] / (float)p.Ka;
          mean[i] = mu;
          rstd[i] = rsqrtf(fmaxf(sq_[i] / (float)p.Ka - mu * mu, 0.f) + 1e-5f);
        }
      }
      const int conv_passes = stationary ? 1 : p1 - p0;
      for (int pass = 0; pass < conv_passes; ++pass) {
        for (int s = 0; s < Ks; ++s, ++a_it) {
          const int k = s * 64 + chunk * 8;
          const bool kin = k < p.Ka;
          long long tw = TC_T0();
          const int st = a_it % p.na;
          mbar_wait(smem_u32(&sm->stage_full[st]), (a_it / p.na) & 1);
          TC_ACC(t_ld, tw);
          // which of this thread's (row, 8-float chunk) cells were actually copied (else: zero)
          bool ok[4];
          if (p.a_mode != A_ROWCOPY) {
            // TMA boxes are complete (hardware zero fill); only whole segments beyond K (and, for images smaller
            // than a tile, rows beyond the box) are absent
#pragma unroll
            for (int i = 0; i < 4; ++i) ok[i] = kin && (p.a_mode != A_TMAP4D || rbase + 32 * i < p.box_rows);
          } else if (!CONV) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ok[i] = kin && (m0 + rbase + 32 * i < m_end);
          } else {
            const int tap = k / p.Cin;
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            const int hw = p.H * p.W;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + rbase + 32 * i;
              ok[i] = kin && m < m_end;
              if (ok[i]) {
                const int b = m / hw, rem = m - b * hw;
                const int y = rem / p.W, x = rem - y * p.W;
                const int yy = y + dy, xx = x + dx;
                ok[i] = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
              }
            }
          }
          float4 v0[4], v1[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rbase + 32 * i;
            if (ok[i]) {
              int off = r * 256 + chunk * 32;
              if (p.a_mode == A_TMAP4D) {
                const int kl = chunk * 8, sub = kl / p.seg;
                off = (sub * p.box_rows + r) * p.seg * 4 + (kl - sub * p.seg) * 4;
              }
              const float* sp = reinterpret_cast<const float*>(a_ring + (size_t)st * STAGE_BYTES + off);
              v0[i] = *reinterpret_cast<const float4*>(sp);
              v1[i] = *reinterpret_cast<const float4*>(sp + 4);
            } else {
              v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              v1[i] = v0[i];
            }
          }
          float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
          float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
          if (has_ln && kin) {
            g0 = ldg4(p.ln_g + k); g1 = ldg4(p.ln_g + k + 4);
            e0 = ldg4(p.ln_b + k); e1 = ldg4(p.ln_b + k + 4);
          }
          const int slot = st;
          // in-place conversion: every converter thread has read its fp32 cells; now the slot may be overwritten
          long long tw2 = TC_T0();
          __syncwarp();  // bar.sync is .aligned: converged again after the lane-0 arrive of the previous slab
          asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
          TC_ACC(t_slot, tw2);
          uint8_t* dst = a_ring + (size_t)slot * a_slot_bytes;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rbase + 32 * i;
            if (has_ln && ok[i]) {
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
            if (parts == 2) *reinterpret_cast<uint4*>(dst + SLAB_BYTES + off) = lo;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sm->a_full[slot]));
        }
      }
    }
    if (p.dbg && ct == 0) {
      p.dbg[blockIdx.x * 16 + 9] = clock64() - t_all0;
      p.dbg[blockIdx.x * 16 + 10] = t_slot;
      p.dbg[blockIdx.x * 16 + 11] = t_ld;
    }
  }

  // teardown: everything issued has completed once the epilogue warps are done with the last tile
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the leader's MMAs read the peer's shared / tensor memory: neither CTA leaves early
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Bimg packing kernels
// ------------------------------------------------------------------------------------------------
// W logical [N][K] fp32 (row n, ld floats per row; optionally transposed source) -> bf16 hi/lo image.
__global__ void __launch_bounds__(256) pack_bimg_kernel(const float* __restrict__ W, long long ld, int transposed,
                                                        long long w_batch_stride, uint16_t* __restrict__ img,
                                                        long long img_batch_elems, int N, int K, int Np, int Ks) {
  const int b = blockIdx.y;
  const float* Wb = W + (long long)b * w_batch_stride;
  uint16_t* out = img + (long long)b * img_batch_elems;
  const long long total = (long long)Ks * Np * 8;  // 16-byte chunks per part
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 7);
    const long long rn = idx >> 3;
    const int n = (int)(rn % Np);
    const int s = (int)(rn / Np);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = s * 64 + c * 8 + 2 * e + h;
        v[h] = (n < N && k < K) ? __ldg(Wb + (transposed ? (long long)k * ld + n : (long long)n * ld + k)) : 0.f;
      }
      split2(v[0], v[1], hi[e], lo[e]);
    }
    const size_t off = bimg_offset(0, s, n, c, Np, Ks);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    const size_t off_lo = bimg_offset(1, s, n, c, Np, Ks);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + off_lo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

static size_t smem_bytes(int na, int nb, int parts, int tepi, int ebox) {
  return 1024 + (size_t)na * STAGE_BYTES + (size_t)nb * BBLK_BYTES * parts +
         (size_t)kEpiWarps * STG_FLOATS * sizeof(float) * (tepi ? ebox : 1);
}

template <int EPI, bool LN, int CG>
static int launch_epi3(const TcArgs& a, size_t smem, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI, LN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI, LN, CG>, a);
  if (le != cudaSuccess) {
    set_error("gemm(tc): cudaLaunchKernelEx failed: %s (grid %d, cluster %d, smem %zu, tiles %d, iters %d)", cudaGetErrorString(le), grid, a.cluster, smem, a.num_tiles, a.iters);
    return MPHSIR_ERR_CUDA;
  }
  return check_launch("gemm(tc)");
}

// a kernel that contains cta_group::2 instructions can only be launched in clusters of two: one instantiation per mode
template <int EPI, bool LN>
static int launch_epi2(const TcArgs& a, size_t smem, int grid, cudaStream_t st) {
  return a.cluster == 2 ? launch_epi3<EPI, LN, 2>(a, smem, grid, st) : launch_epi3<EPI, LN, 1>(a, smem, grid, st);
}

template <int EPI>
static int launch_epi(const TcArgs& a, size_t smem, int grid, cudaStream_t st) {
  if (EPI < TC_OUT_TOKENS && a.ln_g != nullptr) return launch_epi2<EPI, (EPI < TC_OUT_TOKENS)>(a, smem, grid, st);
  return launch_epi2<EPI, false>(a, smem, grid, st);
}

static PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(ptr);
  }
  return fn;
}

// Describe the fp32 A operand to the TMA unit.  Returns false when the shape has no box decomposition
// (the loader then falls back to per-row bulk copies).
static bool make_a_tensor_map(TcArgs& a, bool conv) {
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (enc == nullptr) return false;
  if (!conv) {
    const cuuint64_t rows = a.a_row_mod > 0 ? (cuuint64_t)a.a_row_mod : (cuuint64_t)a.M;
    cuuint64_t gdim[2] = {(cuuint64_t)a.Ka, rows};
    cuuint64_t gstr[1] = {(cuuint64_t)a.lda * 4};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    a.seg = 64;
    if (enc(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.A), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
    a.a_mode = A_TMAP2D;
    return true;
  }
  const int hw = a.H * a.W;
  const int rows = hw < 128 ? hw : 128;
  int bx, by;
  if (a.W >= rows) {
    if (a.W % rows != 0) return false;
    bx = rows;
    by = 1;
  } else {
    if (rows % a.W != 0) return false;
    bx = a.W;
    by = rows / a.W;
  }
  if (hw % rows != 0) return false;
  int seg = 64;
  while (a.Cin % seg != 0) seg >>= 1;
  if (seg < 8) return false;
  const int B = a.M / hw;
  cuuint64_t gdim[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)a.lda * 4, (cuuint64_t)a.W * a.lda * 4, (cuuint64_t)hw * a.lda * 4};
  cuuint32_t box[4] = {(cuuint32_t)seg, (cuuint32_t)bx, (cuuint32_t)by, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (enc(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.A), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  a.seg = seg;
  a.box_rows = rows;
  a.a_mode = A_TMAP4D;
  return true;
}

// 3-D map [cols, rows per sample, samples] of a token-major fp32 matrix, box [32 cols x 32 rows], 128-byte swizzle
static bool make_epi_map(CUtensorMap* tm, const float* base, long long ld, int cols, int rows_per_batch, int batches) {
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (enc == nullptr || base == nullptr) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 3) != 0) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows_per_batch, (cuuint64_t)batches};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rows_per_batch * ld * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_pdl = 0;  // measured on B200: no gain (cube512 14.69 vs 14.78 ms, train64 / patch16 unchanged) — one CTA per SM at 225 KB
                       // of shared memory leaves the successor nothing to overlap but its launch latency; off by default
int pdl_enabled() { return g_pdl; }
void set_pdl_enabled(int on) { g_pdl = on; }
static long long* g_dbg = nullptr;
static int g_tepi_enabled = 1;
void set_tepi_enabled(int on) { g_tepi_enabled = on; }
static int g_ebox1_enabled = 1;
void set_ebox1_enabled(int on) { g_ebox1_enabled = on; }
static int g_tile_rev = 1;
void set_tile_rev(int on) { g_tile_rev = on; }
static int g_psplit_enabled = 1;
void set_psplit_enabled(int on) { g_psplit_enabled = on; }
static int g_cluster_enabled = 1;  // CTA pairs: cta_group::2 MMAs (M = 256), half of every weight block per CTA
void set_cluster_enabled(int on) { g_cluster_enabled = on; }
void set_debug_buffer(long long* p) { g_dbg = p; }

// How a launch is cut into work: pure host arithmetic (also exported as mphsir_gemm_plan for the CPU tests).
struct WorkPlan { int cluster, psplit, ppg, grid, iters, rev, n_full, pass_cols; };
static WorkPlan plan_work(int M, int Np, int ks, int num_tiles, int tiles_per_batch, bool per_sample_weights, int sm_count) {
  WorkPlan w{};
  // few-tile GEMMs (the 16 x 16 latent of a patch batch: 32-64 row tiles for 148 SMs): the 256-column passes of a row tile are
  // handed to several CTAs (each converts the A slabs it needs itself)
  // a few-tile GEMM of exactly two 128-column blocks (N = 256: fc2, the spectral apply, most data gradients of the latent) has a
  // single 256-column pass; with 128-column passes it has two and can be shared by two CTAs like the wider ones
  w.pass_cols = (g_psplit_enabled && num_tiles * 2 <= sm_count && Np > BN && Np <= 2 * BN) ? BN : PASS_COLS;
  const int npass = (Np + w.pass_cols - 1) / w.pass_cols;
  w.psplit = 1;
  w.ppg = npass;
  w.n_full = num_tiles;
  if (g_psplit_enabled && npass >= 2) {
    int split_tiles = 0;
    if (num_tiles * 2 <= sm_count) {
      split_tiles = num_tiles;                         // few-tile launch: every tile is split
    } else if (num_tiles > sm_count && num_tiles % sm_count != 0 && (num_tiles % sm_count) * 2 <= sm_count) {
      // many-tile launch whose last round is at most half full (512 tiles on 148 SMs: 3 rounds + 68 tiles): its tiles are split
      // so that the round costs about half a tile time (every item still converts the A slabs it needs)
      split_tiles = num_tiles % sm_count;
    }
    if (split_tiles > 0) {
      int want = sm_count / split_tiles;
      if (want > npass) want = npass;
      w.ppg = (npass + want - 1) / want;
      w.psplit = (npass + w.ppg - 1) / w.ppg;
      if (w.psplit > 1) w.n_full = num_tiles - split_tiles;
      else w.ppg = npass;
    }
  }
  // CTA pairs run tiles (2 q, 2 q + 1) on one M = 256 instruction stream: both tiles must use the same weights
  // (per-sample weights: an even number of tiles per sample).
  // Measured (tools/gemm_bench.py, tools/shape_profile.py --no-pair): pairs pay when the tensor pipe bounds the tile — many
  // weight blocks per A slab (wide 3x3 convs: -16 %, the K = 256 fc1: -3 %); the HBM-side GEMMs (K <= 128) lose 0-8 % to
  // the lock-step of the two CTAs, so they stay single.
  const int nt_blocks = (Np + BN - 1) / BN;
  const bool tensor_heavy = nt_blocks >= 2 && ks * nt_blocks >= 16;
  w.cluster = (g_cluster_enabled && w.psplit == 1 && tensor_heavy && num_tiles >= 2 && Np % 16 == 0 &&
               (!per_sample_weights || tiles_per_batch % 2 == 0)) ? 2 : 1;
  const int items = w.n_full + (num_tiles - w.n_full) * w.psplit;
  w.grid = items < sm_count ? items : sm_count;
  if (w.cluster == 2) {
    w.grid = (w.grid + 1) & ~1;
    if (w.grid > (sm_count & ~1)) w.grid = sm_count & ~1;
  }
  w.iters = (items + w.grid - 1) / w.grid;
  // measured: 512 x 512 inference -0.9 % per cube, batch-32 training +0.4 %; tensors that fit the L2 anyway gain nothing
  w.rev = (g_tile_rev && M >= 131072) ? 1 : 0;
  return w;
}

int launch_gemm_tc(TcArgs a, bool conv, cudaStream_t st) {
  a.dbg = g_dbg;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  // shared-memory plan (225 KB of 227): 1 KB barriers + A ring (32 KB slots: fp32 TMA landing zone, converted in
  // place to bf16 hi|lo) + B ring (128-row blocks) + 32 KB epilogue staging
  //   bf16x3: A 3 x 32 KB + B 3 x 32 KB        bf16x1: A 4 x 32 KB + B 4 x 16 KB
  // TMA epilogue: plain GEMMs with a BIAS / RESIDUAL (single residual) / PROJ epilogue
  a.tepi = 0;
  if (g_tepi_enabled && !conv &&
      (a.epi == MPHSIR_EPI_BIAS || (a.epi == MPHSIR_EPI_RESIDUAL && a.res2 == nullptr) || a.epi == MPHSIR_EPI_PROJ)) {
    const bool per_sample = a.tiles_per_batch > 0;
    const int rpb = per_sample ? a.rows_per_batch : a.M, nb_ = per_sample ? a.M / a.rows_per_batch : 1;
    const int ycols = a.epi == MPHSIR_EPI_PROJ ? a.n_split : a.N;
    bool ok = make_epi_map(&a.tmY, a.Y, a.ldy, ycols, rpb, nb_);
    if (ok && a.epi == MPHSIR_EPI_PROJ) ok = make_epi_map(&a.tmY2, a.Y2, a.ldy2, a.N - a.n_split, rpb, nb_);
    if (ok && a.epi != MPHSIR_EPI_BIAS) ok = make_epi_map(&a.tmR, a.res1, a.ldr1, ycols, rpb, nb_);
    a.tepi = ok ? 1 : 0;
  }
  //   with the TMA epilogue (64 KB of boxes):  bf16x3: A 3 x 32 KB + B 2 x 32 KB   bf16x1: A 3 x 32 KB + B 4 x 16 KB
  a.na = a.parts == 2 ? 3 : (a.tepi ? 3 : 4);
  a.nb = a.parts == 2 ? (a.tepi ? 2 : 3) : 4;
  a.ebox = 2;
  if (a.tepi && a.epi == MPHSIR_EPI_BIAS && a.ks == 2 && g_ebox1_enabled) {
    a.ebox = 1;   // one store box per epilogue warp -> A ring of 4 slots = two whole K = 128 tiles
    a.na = 4;
  }
  const size_t smem = smem_bytes(a.na, a.nb, a.parts, a.tepi, a.ebox);
  a.a_mode = A_ROWCOPY;
  a.seg = 64;
  if (conv) {
    // conv tiles never straddle samples so that a tile is a box of the [B,H,W,C] tensor
    a.tiles_per_batch = (a.rows_per_batch + 127) / 128;
    a.num_tiles = (a.M / a.rows_per_batch) * a.tiles_per_batch;
  }
  make_a_tensor_map(a, conv);
  const WorkPlan wp = plan_work(a.M, a.Np, a.ks, a.num_tiles, a.tiles_per_batch, a.b_batch_bytes != 0, sm_count);
  a.psplit = wp.psplit; a.ppg = wp.ppg; a.cluster = wp.cluster; a.iters = wp.iters; a.rev = wp.rev; a.n_full = wp.n_full; a.pass_cols = wp.pass_cols;
  const int grid = wp.grid;
  if (a.cluster == 2) a.nb *= 2;   // half-size weight slots: twice the ring depth in the same shared memory
  if (conv && a.epi == MPHSIR_EPI_BIAS) a.epi = TC_OUT_TOKENS;
  switch (a.epi) {
    case MPHSIR_EPI_BIAS: return launch_epi<MPHSIR_EPI_BIAS>(a, smem, grid, st);
    case MPHSIR_EPI_RESIDUAL: return launch_epi<MPHSIR_EPI_RESIDUAL>(a, smem, grid, st);
    case MPHSIR_EPI_GLU: return launch_epi<MPHSIR_EPI_GLU>(a, smem, grid, st);
    case MPHSIR_EPI_SPECTRAL: return launch_epi<MPHSIR_EPI_SPECTRAL>(a, smem, grid, st);
    case MPHSIR_EPI_PROJ: return launch_epi<MPHSIR_EPI_PROJ>(a, smem, grid, st);
    case TC_OUT_TOKENS: return launch_epi<TC_OUT_TOKENS>(a, smem, grid, st);
    case TC_OUT_UNSHUFFLE: return launch_epi<TC_OUT_UNSHUFFLE>(a, smem, grid, st);
    case TC_OUT_SHUFFLE: return launch_epi<TC_OUT_SHUFFLE>(a, smem, grid, st);
    case TC_OUT_NCHW_RES: return launch_epi<TC_OUT_NCHW_RES>(a, smem, grid, st);
    default:
      set_error("gemm_tc: unknown epilogue %d", a.epi);
      return MPHSIR_ERR_INVALID;
  }
}

GemmPlanOut gemm_plan(int M, int Np, int ks, int num_tiles, int tiles_per_batch, bool per_sample_weights, int sm_count) {
  const WorkPlan w = plan_work(M, Np, ks, num_tiles, tiles_per_batch, per_sample_weights, sm_count);
  return GemmPlanOut{w.cluster, w.psplit, w.ppg, w.grid, w.iters, w.rev, w.n_full, w.pass_cols};
}

}  // namespace tc
}  // namespace mphsir

using namespace mphsir;

extern "C" MPHSIR_API void mphsir_debug_tc_counters(long long* buf) { tc::set_debug_buffer(buf); }
extern "C" MPHSIR_API void mphsir_debug_tc_cluster(int enabled) { tc::set_cluster_enabled(enabled); }
extern "C" MPHSIR_API void mphsir_debug_tc_psplit(int enabled) { tc::set_psplit_enabled(enabled); }
extern "C" MPHSIR_API int mphsir_gemm_plan(int M, int N, int K, int rows_per_batch, int per_sample_weights, int sm_count, int* out6) {   /* out6: 8 ints */
  if (M <= 0 || N <= 0 || K <= 0 || sm_count <= 0 || out6 == nullptr) return MPHSIR_ERR_INVALID;
  const int tpb = (per_sample_weights && rows_per_batch > 0) ? (rows_per_batch + 127) / 128 : 0;
  const int tiles = tpb > 0 ? (M / rows_per_batch) * tpb : (M + 127) / 128;
  const tc::GemmPlanOut w = tc::gemm_plan(M, (N + 15) / 16 * 16, (K + 63) / 64, tiles, tpb, per_sample_weights != 0, sm_count);
  out6[0] = w.cluster; out6[1] = w.psplit; out6[2] = w.ppg; out6[3] = w.grid; out6[4] = w.iters; out6[5] = w.rev; out6[6] = w.n_full; out6[7] = w.pass_cols;
  return MPHSIR_OK;
}
extern "C" MPHSIR_API void mphsir_debug_tc_reverse(int enabled) { tc::set_tile_rev(enabled); }
extern "C" MPHSIR_API void mphsir_debug_tc_tma_epilogue(int enabled) { tc::set_tepi_enabled(enabled); }
extern "C" MPHSIR_API void mphsir_debug_tc_ebox1(int enabled) { tc::set_ebox1_enabled(enabled); }
extern "C" MPHSIR_API void mphsir_debug_pdl(int enabled) { tc::set_pdl_enabled(enabled); }

extern "C" size_t mphsir_bimg_bytes(int N, int K) {
  const int Np = (N + 15) / 16 * 16, Ks = (K + 63) / 64;
  return (size_t)2 * Ks * Np * 128;
}

// many weight images in one launch (blockIdx.y = item): the trainer re-packs ~300 matrices after every optimizer step
namespace mphsir {
namespace tc {
constexpr int PACK_MULTI_MAX = 64;
struct PackItem {
  const float* W;
  uint16_t* img;
  int ld, transposed, N, K;
};
struct PackList {
  PackItem it[PACK_MULTI_MAX];
};
__global__ void __launch_bounds__(256) pack_bimg_multi_kernel(const PackList l) {
  const PackItem& q = l.it[blockIdx.y];
  const int Np = (q.N + 15) / 16 * 16, Ks = (q.K + 63) / 64;
  const long long total = (long long)Ks * Np * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 7);
    const long long rn = idx >> 3;
    const int n = (int)(rn % Np);
    const int s = (int)(rn / Np);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = s * 64 + c * 8 + 2 * e + h;
        v[h] = (n < q.N && k < q.K) ? __ldg(q.W + (q.transposed ? (long long)k * q.ld + n : (long long)n * q.ld + k)) : 0.f;
      }
      split2(v[0], v[1], hi[e], lo[e]);
    }
    uint8_t* out = reinterpret_cast<uint8_t*>(q.img);
    *reinterpret_cast<uint4*>(out + bimg_offset(0, s, n, c, Np, Ks)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out + bimg_offset(1, s, n, c, Np, Ks)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
}  // namespace tc
}  // namespace mphsir

extern "C" int mphsir_pack_bimg_multi(const mphsir_pack_item* items, int count, void* stream) {
  MPHSIR_REQUIRE(items && count > 0, "pack_bimg_multi: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i0 = 0; i0 < count; i0 += tc::PACK_MULTI_MAX) {
    const int n = count - i0 < tc::PACK_MULTI_MAX ? count - i0 : tc::PACK_MULTI_MAX;
    tc::PackList l;
    long long max_total = 0;
    for (int j = 0; j < n; ++j) {
      const mphsir_pack_item& q = items[i0 + j];
      MPHSIR_REQUIRE(q.W && q.img && q.N > 0 && q.K > 0 && q.ld > 0, "pack_bimg_multi: bad item %d", i0 + j);
      MPHSIR_REQUIRE((reinterpret_cast<uintptr_t>(q.img) & 127) == 0, "pack_bimg_multi: image %d must be 128-byte aligned", i0 + j);
      l.it[j] = tc::PackItem{q.W, reinterpret_cast<uint16_t*>(q.img), q.ld, q.transposed, q.N, q.K};
      const long long total = (long long)((q.K + 63) / 64) * ((q.N + 15) / 16 * 16) * 8;
      max_total = total > max_total ? total : max_total;
    }
    long long gx = (max_total + 255) / 256;
    if (gx > 48) gx = 48;  // 64 items x 48 CTAs: a few waves; the grid-stride loop covers the larger matrices
    dim3 grid((unsigned)gx, (unsigned)n);
    tc::pack_bimg_multi_kernel<<<grid, 256, 0, st>>>(l);
    const int rc = check_launch("pack_bimg_multi");
    if (rc != MPHSIR_OK) return rc;
  }
  return MPHSIR_OK;
}

extern "C" int mphsir_pack_bimg(const float* W, int ld, int transposed, long long w_batch_stride, void* img,
                                int batch, int N, int K, void* stream) {
  MPHSIR_REQUIRE(W && img && batch > 0 && N > 0 && K > 0 && ld > 0, "pack_bimg: bad arguments");
  MPHSIR_REQUIRE((reinterpret_cast<uintptr_t>(img) & 127) == 0, "pack_bimg: image must be 128-byte aligned");
  const int Np = (N + 15) / 16 * 16, Ks = (K + 63) / 64;
  const long long total = (long long)Ks * Np * 8;
  dim3 grid((unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256), batch);
  tc::pack_bimg_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      W, ld, transposed, w_batch_stride, reinterpret_cast<uint16_t*>(img), (long long)(mphsir_bimg_bytes(N, K) / 2), N,
      K, Np, Ks);
  return check_launch("pack_bimg");
}
