// Shared declarations of the tcgen05 GEMM engine (gemm_tc.cu) and its callers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mphsir {
namespace tc {

enum { TC_OUT_TOKENS = 9, TC_OUT_UNSHUFFLE = 10, TC_OUT_SHUFFLE = 11, TC_OUT_NCHW_RES = 12 };  // conv epilogues

// Byte offset of the 16-byte chunk (8 bf16: k = s*64 + c*8 .. +7) of weight row n in the packed image:
//   [part (hi, lo)][k-slab s][row n] x 128 bytes, chunks XOR-swizzled by (n & 7) — i.e. exactly the
//   SWIZZLE_128B K-major shared-memory image, so a block of rows is one contiguous bulk copy.
__host__ __device__ inline size_t bimg_offset(int part, int s, int n, int c, int Np, int Ks) {
  return ((size_t)(part * Ks + s) * Np + n) * 128 + (size_t)((c ^ (n & 7)) << 4);
}

enum { A_ROWCOPY = 0, A_TMAP2D = 1, A_TMAP4D = 2 };  // how the A loader warp fetches fp32 rows

struct TcArgs {
  alignas(64) CUtensorMap tmA;  // TMA descriptor of the fp32 A operand (2-D [M,K] or 4-D [B,H,W,C])
  alignas(64) CUtensorMap tmY;  // TMA epilogue (tepi): output boxes [32 rows x 32 cols] of Y, 3-D [cols, rows/batch, batch]
  alignas(64) CUtensorMap tmY2; //   EPI_PROJ: second output (columns >= n_split)
  alignas(64) CUtensorMap tmR;  //   residual res1 (loaded into the box the result is stored from)
  int tepi;                     // 1: TMA-fed / TMA-drained epilogue (BIAS, RESIDUAL, PROJ), 0: register-staged epilogue
  int ebox;                     // tepi: 4 KB boxes per epilogue warp (2, or 1 when the smem buys a deeper A ring)
  int a_mode;                   // A_*
  int seg;                      // floats per TMA box row (64 for GEMMs, gcd(Cin,64) for convs)
  int box_rows;                 // conv: pixels per TMA box (= rows of a tile: min(128, H*W))
  const float* A;
  long long lda;
  int a_row_mod;
  int Ka;  // valid A columns (multiple of 8)
  const void* Bimg;
  long long b_batch_bytes;
  int rows_per_batch;
  int tiles_per_batch;
  int num_tiles;
  float* Y;
  long long ldy;
  int M, N, Np, ks;
  int parts;   // 1 = bf16x1, 2 = bf16x3 (hi/lo split)
  int na, nb;  // ring depths (set by the launcher)
  int cluster; // CTAs per cluster (2: weight blocks are fetched once per CTA pair and multicast)
  int iters;   // work-item iterations per CTA (identical for all CTAs so that a cluster stays in lock-step)
  int psplit;  // work items per row tile: few-tile GEMMs (M < 128 * SMs / 2) hand the 256-column passes of a tile to `psplit` CTAs
  int ppg;     // passes per work item
  int pass_cols;  // accumulator columns per pass: 256 (two 128-column blocks), or 128
  int n_full;  // the first n_full row tiles are whole work items; tiles n_full .. num_tiles-1 are cut into psplit pass groups
  int rev;     // 1: work items are walked from the last to the first (L2 reuse of the producer's most recent output)
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
  long long* dbg;  // optional [grid][16] cycle counters (tools/gemm_bench.py --profile); NULL in production
};

int launch_gemm_tc(TcArgs a, bool conv, cudaStream_t st);
struct GemmPlanOut { int cluster, psplit, ppg, grid, iters, rev, n_full, pass_cols; };
GemmPlanOut gemm_plan(int M, int Np, int ks, int num_tiles, int tiles_per_batch, bool per_sample_weights, int sm_count);
void set_debug_buffer(long long* p);
void set_cluster_enabled(int on);
void set_psplit_enabled(int on);
void set_tile_rev(int on);
void set_tepi_enabled(int on);
void set_ebox1_enabled(int on);
int pdl_enabled();           // programmatic dependent launch of the persistent tcgen05 kernels (tc_ptx.cuh: pdl_*)
void set_pdl_enabled(int on);

}  // namespace tc
}  // namespace mphsir
