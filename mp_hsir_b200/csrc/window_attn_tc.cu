// Shifted-window attention core on tcgen05 (sm_100a), TMA-fed, flash-style: softmax(q k^T + bias + mask) v for every
// 8x8 window and head (Spatial_Attention.forward, net/MP_HSIR.py:198-215) with the roll / partition / reverse of
// PGSSTB.forward (:671-696) folded into the TMA coordinates and the store addresses.
//
// A tile is TWO windows of one head: 128 query rows = the M of one tcgen05.mma ("pack two windows per MMA").
//   producer (1 thread) : 24 TMA boxes [4 x 4 pixels x hd] of the fp32 q | k | v columns of the LN+QKV GEMM output — a
//                         window is four such boxes, none of which straddles the cyclic wrap of the shifted grid
//   converters (8 warps): fp32 -> bf16 hi (+ lo) operand images, 128-byte-swizzled K-major: Q [128 x hd] (pre-scaled),
//                         K [128 x hd], and V TRANSPOSED, Vt [hd x 128 keys] (lane = channel, 8 keys per 16-byte store)
//   MMA (1 thread)      : S = Q K^T            (M 128, N 128, K hd; the two windows' blocks sit on the diagonal)
//   softmax (4 warps)   : tcgen05.ld of the row's OWN window's 64 logits, + relative-position bias + closed-form Swin
//                         mask, softmax in registers, P -> bf16 hi/lo -> tcgen05.st back into TENSOR MEMORY in the layout of
//                         a TMEM A operand (the other window's 64 key columns stay zero: written once at kernel start)
//   MMA                 : O = P Vt^T           (A = P from tensor memory, B = Vt from shared memory; M 128, N hd, K 128)
//   epilogue (same warps): tcgen05.ld O, store the fp32 row in image order, per-window token mean (:135) by a register
//                         butterfly over the warp + one shared-memory hand-off between the two warps of a window
// Probabilities and logits never touch shared or global memory.  All hand-offs are mbarriers.  Every role runs one tile
// ahead of the next: S / P / O are double-buffered in tensor memory (P overwrites the S columns it was computed from, as
// in FlashAttention-4), the q|k and the v landing zones are released separately, the MMA thread issues Q K^T of tile
// i+1 before P V of tile i, and the softmax warps do softmax(i+1) before the epilogue of tile i.  HBM traffic: read 3 hd,
// write hd floats per token and head (the minimum).
// Precision: bf16x3 (hi*hi + hi*lo + lo*hi) or bf16, as the mma.sync kernel it replaces (window_attn_mma.cu, kept for
// head dims 48 / 96 of the remote-sensing model).
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mphsir {
namespace watc {

using namespace tc;

// warp 0: TMA producer, 1: MMA, 2-5 / 6-9: two softmax + epilogue groups (even / odd tiles), 10-15: converters.
// The softmax path is a long dependent chain per row (tcgen05.ld -> bias -> max -> 64 ex2 -> split -> tcgen05.st -> wait for
// P V -> tcgen05.ld -> stores -> butterfly): one warp per scheduler ran it at IPC 0.17, so two groups alternate tiles.
constexpr int kThreads = 512;  // 16 warps x 128 registers = the whole register file
constexpr int kSoftWarp0 = 2, kConvWarp0 = 10, kConvThreads = 192;
// tensor memory: buffer u of S / P at columns 128 u (S fp32 [128]; then P hi [0,64) | P lo [64,128) in place), O at 256 + 64 u
constexpr int SP_COL = 0, PL_OFF = 64, O_COL = 256;

template <int HD>
struct Plan {
  static constexpr int LAND_OP = 128 * HD * 4;  // one operand (q, k or v) of both windows, fp32
  static constexpr int QK_PART = 128 * 128;     // [128 rows][128 B] (row pitch 128 B for HD = 32 as well)
  static constexpr int VT_SLAB = HD * 128;      // [HD rows][64 keys]
  static constexpr int VT_PART = 2 * VT_SLAB;
  // barriers + landing + Q, K images + TWO Vt images + psum [2 groups][4][HD]: exactly 227 KB for HD = 64 with hi/lo parts
  static size_t smem(int parts) { return 1024 + 3 * (size_t)LAND_OP + 2 * (size_t)parts * QK_PART + 2 * (size_t)parts * VT_PART + 2 * 4 * HD * 4; }
};

struct Bars {
  uint64_t lqk_full, lqk_empty, lv_full, lv_empty;   // landing zones (TMA -> converters)
  uint64_t qk_full, qk_empty, v_full[2], v_empty[2]; // operand images (converters -> MMA); Vt is double-buffered: with one
                                                     // buffer V(t+1) waited for P V of tile t, which chained every second
                                                     // tile behind a whole convert -> QK^T -> softmax -> PV round trip
  uint64_t s_full[2], sp_empty[2], p_full[2], o_full[2], o_empty[2];   // tensor-memory buffers
  uint32_t tmem_base;
};

struct Args {
  alignas(64) CUtensorMap tm;  // qkv as [B][H][W][3C] fp32, box [1][4][4][HD]
  const float* bias;           // TRANSPOSED relative-position bias [heads][key 64][query 64]: lanes = consecutive queries read one line per key
  float* out;
  long long ldo;
  float* win_mean;             // [B*nW][C]
  int B, H, W, C, heads, shift, parts, mask_H, mask_y0;
  int n_windows, n_tiles;
  long long* cnt;  // optional [grid][16] cycle counters (role totals and waits), NULL in production
  int dbg;  // timing experiments only (mphsir_debug_window_attn_tc(1 | flags << 4)): 1 skip conversion math, 2 skip softmax math, 4 skip TMA loads, 8 skip epilogue stores/butterfly
};

// token m of a tile (two windows of 64 tokens, row-major 8x8) -> float offset of its pixel inside one operand's landing zone
template <int HD>
__device__ __forceinline__ int land_pixel(int m) {
  const int win = m >> 6, t = m & 63, r = t >> 3, c = t & 7;
  const int box = win * 4 + (r >> 2) * 2 + (c >> 2);
  return (box * 16 + (r & 3) * 4 + (c & 3)) * HD;
}

#define W_T0() (p.cnt ? clock64() : 0)
#define W_ACC(var, t0) do { if (p.cnt) var += clock64() - (t0); } while (0)
#define W_WAIT(var, barp, par) do { long long _t = W_T0(); mbar_wait(smem_u32(barp), par); W_ACC(var, _t); } while (0)

template <int HD>
__global__ void __launch_bounds__(kThreads, 1) window_attn_tc_kernel(const __grid_constant__ Args p) {
  using P = Plan<HD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Bars* bar = reinterpret_cast<Bars*>(smem_raw);
  uint8_t* land = smem_raw + 1024;
  uint8_t* q_img = land + 3 * P::LAND_OP;
  uint8_t* k_img = q_img + p.parts * P::QK_PART;
  uint8_t* vt_img = k_img + p.parts * P::QK_PART;
  float* psum = reinterpret_cast<float*>(vt_img + 2 * p.parts * P::VT_PART);  // [2 groups][4 quadrants][HD]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int parts = p.parts;
  const int nWx = p.W >> 3, nW = (p.H >> 3) * nWx;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar->lqk_full), 1);
    mbar_init(smem_u32(&bar->lqk_empty), kConvThreads / 32);
    mbar_init(smem_u32(&bar->lv_full), 1);
    mbar_init(smem_u32(&bar->lv_empty), kConvThreads / 32);
    mbar_init(smem_u32(&bar->qk_full), kConvThreads / 32);
    mbar_init(smem_u32(&bar->qk_empty), 1);
    for (int u = 0; u < 2; ++u) {
      mbar_init(smem_u32(&bar->v_full[u]), kConvThreads / 32);
      mbar_init(smem_u32(&bar->v_empty[u]), 1);
      mbar_init(smem_u32(&bar->s_full[u]), 1);
      mbar_init(smem_u32(&bar->sp_empty[u]), 1);
      mbar_init(smem_u32(&bar->p_full[u]), 4);
      mbar_init(smem_u32(&bar->o_full[u]), 1);
      mbar_init(smem_u32(&bar->o_empty[u]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&bar->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;

  if (warp == 0) {
    // =============================== TMA producer ============================================================
    if (lane == 0) {
      uint32_t it = 0;
      long long w_empty = 0, t_all = W_T0();
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int pair = tile / p.heads, h = tile - pair * p.heads;
        const int nwin = (2 * pair + 1 < p.n_windows) ? 2 : 1;
        // q | k boxes first (their landing zone is released as soon as Q and K are converted), then the v boxes
#pragma unroll 1
        for (int grp = 0; grp < 2; ++grp) {
          const uint32_t full = smem_u32(grp == 0 ? &bar->lqk_full : &bar->lv_full);
          W_WAIT(w_empty, grp == 0 ? &bar->lqk_empty : &bar->lv_empty, (it & 1) ^ 1);
          if (p.dbg & 4) { mbar_arrive(full); continue; }
          mbar_expect_tx(full, (uint32_t)((grp == 0 ? 2 : 1) * nwin * 64 * HD * 4));
          for (int win = 0; win < nwin; ++win) {
            const int w = 2 * pair + win;
            const int b = w / nW, wrem = w - b * nW;
            const int wi = wrem / nWx, wj = wrem - wi * nWx;
            for (int op = (grp == 0 ? 0 : 2); op < (grp == 0 ? 2 : 3); ++op)
#pragma unroll
              for (int bx = 0; bx < 4; ++bx) {
                int y = wi * 8 + (bx >> 1) * 4 + p.shift, x = wj * 8 + (bx & 1) * 4 + p.shift;  // roll(-s): shifted[ys] = img[(ys+s) % H]
                if (y >= p.H) y -= p.H;
                if (x >= p.W) x -= p.W;
                tma_load_4d(smem_u32(land + op * P::LAND_OP + ((win * 4 + bx) * 16) * HD * 4), &p.tm, op * p.C + h * HD, x, y, b, full);
              }
          }
        }
      }
      if (p.cnt) { p.cnt[blockIdx.x * 16 + 0] = clock64() - t_all; p.cnt[blockIdx.x * 16 + 1] = w_empty; }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ==============================================================
    // ONE thread, readiness-ordered: Q K^T of the next tile and P V of the oldest pending tile are polled without
    // blocking and issued whichever is ready first (a fixed order let a late q | k landing hold back a P V that was
    // ready, and with it the epilogue, the V conversion and the next loads).  At most two tiles are in flight (two
    // tensor-memory buffers).
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(128), idesc_o = make_idesc(HD);
      int n_local = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) ++n_local;
      const uint64_t qh = make_desc(smem_u32(q_img)), kh = make_desc(smem_u32(k_img));
      const uint64_t ql = make_desc(smem_u32(q_img + P::QK_PART)), kl = make_desc(smem_u32(k_img + P::QK_PART));
      const uint32_t vt = smem_u32(vt_img);
      int n1 = 0, n2 = 0;   // tiles whose S / whose O have been issued
      long long t_all = W_T0(), t_issue = 0;
      while (n2 < n_local) {
        if (n1 < n_local && n1 - n2 < 2) {
          const int u = n1 & 1;
          if (mbar_try_wait(smem_u32(&bar->qk_full), n1 & 1) && mbar_try_wait(smem_u32(&bar->sp_empty[u]), ((n1 >> 1) & 1) ^ 1)) {
            tc_fence_after();
            const long long ti = W_T0();
            const uint32_t d = tmem_base + SP_COL + 128 * u;
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) {
              umma_bf16(d, qh + 2 * ks, kh + 2 * ks, idesc_s, ks != 0);
              if (parts == 2) {
                umma_bf16(d, qh + 2 * ks, kl + 2 * ks, idesc_s, 1);
                umma_bf16(d, ql + 2 * ks, kh + 2 * ks, idesc_s, 1);
              }
            }
            umma_commit(smem_u32(&bar->s_full[u]));
            umma_commit(smem_u32(&bar->qk_empty));
            W_ACC(t_issue, ti);
            ++n1;
          }
        }
        if (n2 < n1) {
          const int u = n2 & 1;
          if (mbar_try_wait(smem_u32(&bar->v_full[u]), (n2 >> 1) & 1) && mbar_try_wait(smem_u32(&bar->p_full[u]), (n2 >> 1) & 1) &&
              mbar_try_wait(smem_u32(&bar->o_empty[u]), ((n2 >> 1) & 1) ^ 1)) {
            tc_fence_after();
            const long long ti = W_T0();
            const uint32_t d = tmem_base + O_COL + 64 * u, ph = tmem_base + SP_COL + 128 * u, pl = ph + PL_OFF;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {  // 16 keys per step
              const uint32_t vb = vt + u * parts * P::VT_PART;
              const uint64_t vh = make_desc(vb + (kk >> 2) * P::VT_SLAB) + 2 * (kk & 3);
              const uint64_t vl = make_desc(vb + P::VT_PART + (kk >> 2) * P::VT_SLAB) + 2 * (kk & 3);
              umma_bf16_tmem_a(d, ph + 8 * kk, vh, idesc_o, kk != 0);
              if (parts == 2) {
                umma_bf16_tmem_a(d, ph + 8 * kk, vl, idesc_o, 1);
                umma_bf16_tmem_a(d, pl + 8 * kk, vh, idesc_o, 1);
              }
            }
            umma_commit(smem_u32(&bar->o_full[u]));
            umma_commit(smem_u32(&bar->v_empty[u]));
            umma_commit(smem_u32(&bar->sp_empty[u]));
            W_ACC(t_issue, ti);
            ++n2;
          }
        }
      }
      if (p.cnt) { p.cnt[blockIdx.x * 16 + 2] = clock64() - t_all; p.cnt[blockIdx.x * 16 + 3] = t_issue; }
    }
  } else if (warp < kConvWarp0) {
    // =============================== softmax + epilogue (warps 2..9, two groups) =============================
    const int grp = (warp - kSoftWarp0) >> 2;  // group g owns tiles it = g, g + 2, ... and tensor-memory buffer u = g
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
    const int m = quad * 32 + lane;            // query row of the tile
    const int win = m >> 6, t = m & 63, r = t >> 3, c = t & 7;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool rq_low = r < 4, cq_low = c < 4;
    const int et = threadIdx.x - (kSoftWarp0 + 4 * grp) * 32;   // 0..127 inside the group
    const int u = grp;
    const uint32_t sp = trow + SP_COL + 128 * u;
    float* ps_grp = psum + grp * 4 * HD;
    int n_local = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) ++n_local;
    long long w_s = 0, w_o = 0, t_soft = 0, t_epi = 0, t_all = W_T0();
    for (int it = grp; it < n_local; it += 2) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int pair = tile / p.heads, h = tile - pair * p.heads;
      const int w = 2 * pair + win;
      const bool valid = w < p.n_windows;
      const int wc = valid ? w : 2 * pair;
      const int b = wc / nW, wrem = wc - b * nW;
      const int wi = wrem / nWx, wj = wrem - wi * nWx;
      // Swin mask (:643-658) in closed form: only the last window row / column of the SCENE's shifted grid is split
      int ysg = wi * 8 + p.mask_y0;
      if (ysg >= p.mask_H) ysg -= p.mask_H;
      const bool lastrow = p.shift != 0 && ysg >= p.mask_H - 8;
      const bool lastcol = p.shift != 0 && wj == nWx - 1;
      const float* bcol = p.bias + (long long)h * 64 * 64 + t;   // + 64 * key: coalesced over the warp's 32 consecutive queries

      // bias + mask of this row first: the loads are in flight while the warp waits for S.  The table is read TRANSPOSED
      // ([key][query]): one 128-byte line per key and warp — row-per-lane float4 loads of the [query][key] table touched
      // 32 lines per instruction and, with the row-per-lane output stores, saturated the L1/LSU data path (ncu: l1tex 72 %)
      float s[64];
#pragma unroll
      for (int kj = 0; kj < 64; ++kj) {
        const bool masked = (lastrow && (rq_low != (kj < 32))) || (lastcol && (cq_low != ((kj & 7) < 4)));
        s[kj] = __ldg(bcol + 64 * kj) + (masked ? -100.f : 0.f);
      }
      W_WAIT(w_s, &bar->s_full[u], (it >> 1) & 1);
      const long long ts0 = W_T0();
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t raw[16];
        tmem_ld16_nowait(sp + 64 * win + c0, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          s[c0 + j] += __uint_as_float(raw[j]);
          mx = fmaxf(mx, s[c0 + j]);
        }
      }
      float sum = 0.f;
      if (!(p.dbg & 2)) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          s[j] = __expf(s[j] - mx);
          sum += s[j];
        }
      } else sum = 1.f;
      const float inv = 1.0f / sum;
      // P = softmax row -> bf16 hi/lo pairs -> over the S columns of this row (the A operand of O = P V lives in tensor
      // memory); the other window's 64 keys get zeros (S holds cross-window products there)
      const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split2(s[16 * g + 2 * e] * inv, s[16 * g + 2 * e + 1] * inv, hi[e], lo[e]);
        tmem_st8(sp + 32 * win + 8 * g, hi);
        tmem_st8(sp + 32 * (win ^ 1) + 8 * g, z);
        if (parts == 2) {
          tmem_st8(sp + PL_OFF + 32 * win + 8 * g, lo);
          tmem_st8(sp + PL_OFF + 32 * (win ^ 1) + 8 * g, z);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar->p_full[u]));
      __syncwarp();
      W_ACC(t_soft, ts0);

      // ---- epilogue: O row -> image order, window mean (the other group runs the next tile's softmax meanwhile) ----
      W_WAIT(w_o, &bar->o_full[u], (it >> 1) & 1);
      const long long te0 = W_T0();
      tc_fence_after();
      float o[HD];
      {
        uint32_t raw[32];
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {
          tmem_ld32(trow + O_COL + 64 * u + c0, raw);
#pragma unroll
          for (int j = 0; j < 32; ++j) o[c0 + j] = __uint_as_float(raw[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar->o_empty[u]));
      __syncwarp();
      if (valid && !(p.dbg & 8)) {
        int y = wi * 8 + r + p.shift, x = wj * 8 + c + p.shift;
        if (y >= p.H) y -= p.H;
        if (x >= p.W) x -= p.W;
        float* dst = p.out + ((long long)(b * p.H + y) * p.W + x) * p.ldo + h * HD;
#pragma unroll
        for (int j4 = 0; j4 < HD / 4; ++j4)
          *reinterpret_cast<float4*>(dst + 4 * j4) = make_float4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]);
      }
      // column sums over the warp's 32 rows: butterfly that halves the column set at every step; afterwards lane l holds
      // HD/32 consecutive columns starting at col0(l)
      int col0 = 0;
#pragma unroll
      for (int st = 0; st < 5; ++st) {
        const int off = 16 >> st, half = HD >> (st + 1);
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float send = upper ? o[i] : o[i + half];
          const float keep = upper ? o[i + half] : o[i];
          o[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        col0 += upper ? half : 0;
      }
      // the group's previous tile read ps_grp before its second barrier, so it may be overwritten now
#pragma unroll
      for (int i = 0; i < HD / 32; ++i) ps_grp[quad * HD + col0 + i] = o[i];
      if (grp == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
      else asm volatile("bar.sync 4, 128;" ::: "memory");
      if (et < 2 * HD) {
        const int ew = et / HD, col = et - ew * HD;
        const int wg = 2 * pair + ew;
        if (wg < p.n_windows)
          p.win_mean[(long long)wg * p.C + h * HD + col] = (ps_grp[2 * ew * HD + col] + ps_grp[(2 * ew + 1) * HD + col]) * (1.0f / 64.0f);
      }
      if (grp == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
      else asm volatile("bar.sync 4, 128;" ::: "memory");
      W_ACC(t_epi, te0);
    }
    if (p.cnt && et == 0) {
      long long* c = p.cnt + blockIdx.x * 16 + 4 + 5 * grp;
      c[0] = clock64() - t_all; c[1] = w_s; c[2] = w_o; c[3] = t_soft; c[4] = t_epi;
    }
  } else {
    // =============================== converters (warps 10..15) ===============================================
    const int ct = threadIdx.x - kConvWarp0 * 32;  // 0..191
    const float scale = rsqrtf((float)HD);
    const float* lq = reinterpret_cast<const float*>(land);
    const float* lk = reinterpret_cast<const float*>(land + P::LAND_OP);
    const float* lv = reinterpret_cast<const float*>(land + 2 * P::LAND_OP);
    constexpr int CH = HD / 8;
    constexpr int NQK = (128 * CH + kConvThreads - 1) / kConvThreads;  // items per thread: (row m, 16-byte chunk ch) of Q and K
    constexpr int NV = (HD * 16 + kConvThreads - 1) / kConvThreads;    // items per thread: (channel d, 8 keys of one window row)
    // the item -> address maps do not depend on the tile: hoisted out of the loop (the index arithmetic was a third of the
    // converters' instructions)
    int qk_src[NQK], qk_dst[NQK], v_src[NV], v_dst[NV];
    uint32_t win1_bits = 0;   // bit i: QK item i belongs to the second window; bit 16 + i: V item i
#pragma unroll
    for (int i = 0; i < NQK; ++i) {
      const int item = ct + kConvThreads * i;
      const int m = item / CH, ch = item - m * CH;
      qk_src[i] = item < 128 * CH ? land_pixel<HD>(m) + ch * 8 : -1;
      qk_dst[i] = m * 128 + ((ch ^ (m & 7)) << 4);
      win1_bits |= (uint32_t)(m >= 64) << i;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int item = ct + kConvThreads * i;
      const int kg = item / HD, d = item - kg * HD;   // lanes = consecutive channels: conflict-free scalar reads
      v_src[i] = item < HD * 16 ? land_pixel<HD>(kg * 8) + d : -1;   // key e of the row: + (e >> 2) * 16 * HD + (e & 3) * HD
      v_dst[i] = (kg >> 3) * P::VT_SLAB + d * 128 + (((kg & 7) ^ (d & 7)) << 4);
      win1_bits |= (uint32_t)(kg >= 8) << (16 + i);
    }
    uint32_t it = 0;
    long long w_land = 0, w_img = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int pair = tile / p.heads;
      const bool two = 2 * pair + 1 < p.n_windows;
      W_WAIT(w_land, &bar->lqk_full, it & 1);
      // ---- Q (pre-scaled), K: row m, 16-byte chunk ch of the K-major image ----
      W_WAIT(w_img, &bar->qk_empty, (it & 1) ^ 1);
      if (!(p.dbg & 1)) {
#pragma unroll
        for (int i = 0; i < NQK; ++i) {
          if (qk_src[i] < 0) continue;
          const bool ok = two || !((win1_bits >> i) & 1u);
          float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
          if (ok) {
            a0 = *reinterpret_cast<const float4*>(lq + qk_src[i]);
            a1 = *reinterpret_cast<const float4*>(lq + qk_src[i] + 4);
            b0 = *reinterpret_cast<const float4*>(lk + qk_src[i]);
            b1 = *reinterpret_cast<const float4*>(lk + qk_src[i] + 4);
          }
          uint4 hi, lo;
          split2(a0.x * scale, a0.y * scale, hi.x, lo.x);
          split2(a0.z * scale, a0.w * scale, hi.y, lo.y);
          split2(a1.x * scale, a1.y * scale, hi.z, lo.z);
          split2(a1.z * scale, a1.w * scale, hi.w, lo.w);
          *reinterpret_cast<uint4*>(q_img + qk_dst[i]) = hi;
          if (parts == 2) *reinterpret_cast<uint4*>(q_img + P::QK_PART + qk_dst[i]) = lo;
          split2(b0.x, b0.y, hi.x, lo.x);
          split2(b0.z, b0.w, hi.y, lo.y);
          split2(b1.x, b1.y, hi.z, lo.z);
          split2(b1.z, b1.w, hi.w, lo.w);
          *reinterpret_cast<uint4*>(k_img + qk_dst[i]) = hi;
          if (parts == 2) *reinterpret_cast<uint4*>(k_img + P::QK_PART + qk_dst[i]) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar->qk_full));
        mbar_arrive(smem_u32(&bar->lqk_empty));   // q | k landing zone read out: the next tile's q | k may land
      }
      __syncwarp();
      // ---- V transposed: row = channel d, K = key; one item = 8 keys (one row of a window) of one channel ----
      const int vu = it & 1;
      uint8_t* vt_buf = vt_img + vu * parts * P::VT_PART;
      W_WAIT(w_land, &bar->lv_full, it & 1);
      W_WAIT(w_img, &bar->v_empty[vu], ((it >> 1) & 1) ^ 1);
      if (!(p.dbg & 1)) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (v_src[i] < 0) continue;
          const bool ok = two || !((win1_bits >> (16 + i)) & 1u);
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = ok ? lv[v_src[i] + (e >> 2) * 16 * HD + (e & 3) * HD] : 0.f;
          uint4 hi, lo;
          split2(v[0], v[1], hi.x, lo.x);
          split2(v[2], v[3], hi.y, lo.y);
          split2(v[4], v[5], hi.z, lo.z);
          split2(v[6], v[7], hi.w, lo.w);
          *reinterpret_cast<uint4*>(vt_buf + v_dst[i]) = hi;
          if (parts == 2) *reinterpret_cast<uint4*>(vt_buf + P::VT_PART + v_dst[i]) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar->v_full[vu]));
        mbar_arrive(smem_u32(&bar->lv_empty));
      }
      __syncwarp();
    }
    if (p.cnt && ct == 0) { p.cnt[blockIdx.x * 16 + 14] = w_land; p.cnt[blockIdx.x * 16 + 15] = w_img; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static PFN_cuTensorMapEncodeTiled encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(ptr);
  }
  return fn;
}

static bool g_enabled = true;
static int g_dbg = 0;
static long long* g_cnt = nullptr;
void set_counters(long long* p) { g_cnt = p; }

template <int HD>
static int launch_t(const float* qkv, int ldqkv, const float* bias, float* out, int ldo, float* win_mean, int B, int H, int W,
                    int C, int heads, int shift, int parts, int mask_H, int mask_y0, cudaStream_t st) {
  PFN_cuTensorMapEncodeTiled enc = encode_fn();
  if (enc == nullptr) {
    set_error("window_attn(tc): cuTensorMapEncodeTiled unavailable");
    return MPHSIR_ERR_CUDA;
  }
  Args a{};
  cuuint64_t gdim[4] = {(cuuint64_t)3 * C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)ldqkv * 4, (cuuint64_t)W * ldqkv * 4, (cuuint64_t)H * W * ldqkv * 4};
  cuuint32_t box[4] = {(cuuint32_t)HD, 4, 4, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (enc(&a.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(qkv), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    set_error("window_attn(tc): cuTensorMapEncodeTiled failed (C=%d H=%d W=%d ld=%d)", C, H, W, ldqkv);
    return MPHSIR_ERR_CUDA;
  }
  a.bias = bias; a.out = out; a.ldo = ldo; a.win_mean = win_mean;
  a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads; a.shift = shift; a.parts = parts; a.mask_H = mask_H; a.mask_y0 = mask_y0;
  a.n_windows = B * (H / 8) * (W / 8);
  a.n_tiles = ((a.n_windows + 1) / 2) * heads;
  a.dbg = g_dbg;
  a.cnt = g_cnt;
  static int sm_count = 0;
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(window_attn_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Plan<HD>::smem(2));
    if (e != cudaSuccess) {
      set_error("window_attn(tc): cudaFuncSetAttribute(%zu B): %s", Plan<HD>::smem(2), cudaGetErrorString(e));
      return MPHSIR_ERR_CUDA;
    }
    configured = true;
  }
  const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
  window_attn_tc_kernel<HD><<<grid, kThreads, Plan<HD>::smem(parts), st>>>(a);
  return check_launch("window_attn(tc)");
}

void set_enabled(int on) { g_enabled = (on & 1) != 0; g_dbg = on >> 4; }

// head dims 32 / 64 (every stage of the natural-scene model), qkv rows 16-byte aligned
bool supported(int hd, int ldqkv, int ldo, const float* qkv) {
  return (hd == 32 || hd == 64) && ldqkv % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0;
}
bool enabled() { return g_enabled; }

int launch(const float* qkv, int ldqkv, const float* bias, float* out, int ldo, float* win_mean, int B, int H, int W, int C, int heads,
           int shift, int parts, int mask_H, int mask_y0, cudaStream_t st) {
  if (C / heads == 32)
    return launch_t<32>(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, parts, mask_H, mask_y0, st);
  return launch_t<64>(qkv, ldqkv, bias, out, ldo, win_mean, B, H, W, C, heads, shift, parts, mask_H, mask_y0, st);
}

}  // namespace watc
}  // namespace mphsir

extern "C" MPHSIR_API void mphsir_debug_window_attn_tc(int enabled) { mphsir::watc::set_enabled(enabled); }
extern "C" MPHSIR_API int mphsir_window_attn_tc_enabled(void) { return mphsir::watc::enabled() ? 1 : 0; }

extern "C" int mphsir_window_attn_tc_supported(int head_dim) { return head_dim == 32 || head_dim == 64; }

extern "C" int mphsir_window_attn_tc_fwd(const float* qkv, int ldqkv, const float* bias_t, float* out, int ldo, float* win_mean, int B,
                                         int H, int W, int C, int heads, int shift, int precision, int mask_H, int mask_y0,
                                         void* stream) {
  using namespace mphsir;
  MPHSIR_REQUIRE(qkv && bias_t && out && win_mean, "window_attn(tc): null operand");
  MPHSIR_REQUIRE(B > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "window_attn(tc): H=%d W=%d must be multiples of 8", H, W);
  MPHSIR_REQUIRE(heads > 0 && C % heads == 0 && (C / heads == 32 || C / heads == 64), "window_attn(tc): head_dim %d not in {32, 64}", heads > 0 ? C / heads : 0);
  MPHSIR_REQUIRE(shift == 0 || shift == 4, "window_attn(tc): shift must be 0 or 4");
  MPHSIR_REQUIRE(ldqkv >= 3 * C && ldqkv % 4 == 0 && ldo >= C && ldo % 4 == 0, "window_attn(tc): bad leading dimensions");
  MPHSIR_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "window_attn(tc): qkv / out must be 16-byte aligned");
  MPHSIR_REQUIRE(precision == MPHSIR_PREC_BF16X3 || precision == MPHSIR_PREC_BF16, "window_attn(tc): tensor-core precisions only");
  MPHSIR_REQUIRE(mask_H >= 8 && mask_y0 >= 0 && mask_y0 < mask_H, "window_attn(tc): bad mask geometry (mask_H=%d mask_y0=%d)", mask_H, mask_y0);
  return watc::launch(qkv, ldqkv, bias_t, out, ldo, win_mean, B, H, W, C, heads, shift, precision == MPHSIR_PREC_BF16X3 ? 2 : 1, mask_H,
                      mask_y0, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" MPHSIR_API void mphsir_debug_window_attn_tc_counters(long long* buf) { mphsir::watc::set_counters(buf); }
