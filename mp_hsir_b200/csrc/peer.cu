// Collectives of the row-sharded scene (mp_hsir_b200/sharded.py) over NVLink peer memory, one kernel each.
//
// Both collectives of the sharded forward are tiny and latency-bound: per block one halo refresh of 8 image
// rows each way (<= 2 MB) and one sum of the Gram statistics (<= 75 KB) over the ranks.  An NCCL call costs 20-40 us of launch
// + protocol per collective; ~50 of them bound a scene at 4-8 GPUs.  Here every rank owns a *window* — one device
// allocation exported through CUDA IPC and mapped by all peers — holding mailboxes, gather slots and flags:
//
//   halo exchange : push my first / last 8 rows straight into the neighbours' mailboxes (stores over NVLink), publish a
//                   flag, wait for the neighbours' flags, copy my mailboxes into my halo rows.          ONE kernel.
//   all-reduce    : push my vector into slot [rank] of every rank's gather area, publish flags, wait for all G flags, sum
//                   the G slots in rank order (bit-identical on every rank, deterministic).            ONE kernel.
//
// Protocol.  A per-rank sequence number lives in the rank's own window (device memory, so a CUDA-graph replay advances it
// like an eager launch).  All ranks issue the same collectives in the same order (the forward is SPMD), so sequence numbers
// agree.  Mailboxes / slots are double-buffered by sequence parity: a neighbour can be at most one collective ahead (its
// next collective cannot complete before mine has published), so parity p is never overwritten while it is being read.
// Data stores are made visible before the flag by __threadfence_system() + a release store; readers poll with acquire
// loads and read the payload with L1-bypassing loads.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace mphsir {
namespace peer {

constexpr int MAX_RANKS = 8;
constexpr int CTRL_WORDS = 64;  // control block at the start of a window (uint32): see offsets below
// control block layout (uint32 words)
constexpr int W_SEQ_HALO = 0;    // halo sequence number of this rank (advanced by the last CTA of every halo kernel)
constexpr int W_SEQ_AR = 1;      // all-reduce sequence number
constexpr int W_ARRIVE = 2;      // CTAs that finished pushing (reset by the last one)
constexpr int W_DEPART = 3;      // CTAs that finished the whole kernel (reset by the last one)
constexpr int W_FLAG_ABOVE = 8;  // written by the PREVIOUS rank: its rows for my top halo are in my mailbox (value = seq)
constexpr int W_FLAG_BELOW = 9;  // written by the NEXT rank
constexpr int W_FLAG_AR = 16;    // [MAX_RANKS] written by rank q: its vector is in my gather slot q

struct Window {       // what a rank knows about one window (its own or a peer's), all device pointers valid in THIS process
  uint32_t* ctrl;     // [CTRL_WORDS]
  float* mail;        // [2 parity][2 direction (0 = from above, 1 = from below)][halo_cap]
  float* slots;       // [2 parity][MAX_RANKS][ar_cap]
};

struct Peers {
  Window w[MAX_RANKS];
  int rank, world;
  long long halo_cap, ar_cap;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t seq) {
  while ((int32_t)(ld_acquire_sys(p) - seq) < 0) __nanosleep(64);
}
__device__ __forceinline__ void copy_f4(float* __restrict__ dst, const float* __restrict__ src, long long n4, int tid, int nthreads,
                                        bool bypass_l1) {
  const float4* s = reinterpret_cast<const float4*>(src);
  float4* d = reinterpret_cast<float4*>(dst);
  for (long long i = tid; i < n4; i += nthreads) d[i] = bypass_l1 ? __ldcg(s + i) : s[i];
}

// grid: a few dozen CTAs, ALL resident (they wait on remote flags); n = floats per direction (multiple of 4)
__global__ void __launch_bounds__(256) halo_exchange_kernel(const Peers P, const float* __restrict__ own_first,
                                                            const float* __restrict__ own_last, float* __restrict__ halo_top,
                                                            float* __restrict__ halo_bottom, long long n) {
  const Window me = P.w[P.rank];
  const int prev = (P.rank + P.world - 1) % P.world, next = (P.rank + 1) % P.world;
  __shared__ uint32_t s_seq;
  if (threadIdx.x == 0) s_seq = ld_acquire_sys(me.ctrl + W_SEQ_HALO) + 1u;   // stable until the last CTA departs
  __syncthreads();
  const uint32_t seq = s_seq;
  const int par = seq & 1u;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  // ---- push: my first rows are the BOTTOM halo of the previous rank, my last rows the TOP halo of the next rank ----
  copy_f4(P.w[prev].mail + ((long long)par * 2 + 1) * P.halo_cap, own_first, n4, tid, nthreads, false);
  copy_f4(P.w[next].mail + ((long long)par * 2 + 0) * P.halo_cap, own_last, n4, tid, nthreads, false);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(me.ctrl + W_ARRIVE, 1u) == gridDim.x - 1) {   // every CTA of this rank has pushed and fenced
      atomicExch(me.ctrl + W_ARRIVE, 0u);
      __threadfence_system();
      st_release_sys(P.w[prev].ctrl + W_FLAG_BELOW, seq);
      st_release_sys(P.w[next].ctrl + W_FLAG_ABOVE, seq);
    }
    // ---- wait for the neighbours' rows ----
    wait_flag(me.ctrl + W_FLAG_ABOVE, seq);
    wait_flag(me.ctrl + W_FLAG_BELOW, seq);
  }
  __syncthreads();
  // ---- pull: mailboxes -> halo rows (payload written by a peer GPU: bypass L1) ----
  copy_f4(halo_top, me.mail + ((long long)par * 2 + 0) * P.halo_cap, n4, tid, nthreads, true);
  copy_f4(halo_bottom, me.mail + ((long long)par * 2 + 1) * P.halo_cap, n4, tid, nthreads, true);
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(me.ctrl + W_DEPART, 1u) == gridDim.x - 1) {
    atomicExch(me.ctrl + W_DEPART, 0u);
    st_release_sys(me.ctrl + W_SEQ_HALO, seq);   // the next halo kernel of this rank starts from here
  }
}

// data[0..n) <- sum over ranks, identical bits on every rank
__global__ void __launch_bounds__(256) all_reduce_kernel(const Peers P, float* __restrict__ data, long long n) {
  const Window me = P.w[P.rank];
  __shared__ uint32_t s_seq;
  if (threadIdx.x == 0) s_seq = ld_acquire_sys(me.ctrl + W_SEQ_AR) + 1u;
  __syncthreads();
  const uint32_t seq = s_seq;
  const int par = seq & 1u;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  for (int q = 0; q < P.world; ++q) {
    float* dst = P.w[q].slots + ((long long)par * MAX_RANKS + P.rank) * P.ar_cap;
    for (long long i = tid; i < n; i += nthreads) dst[i] = data[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(me.ctrl + W_ARRIVE, 1u) == gridDim.x - 1) {
      atomicExch(me.ctrl + W_ARRIVE, 0u);
      __threadfence_system();
      for (int q = 0; q < P.world; ++q) st_release_sys(P.w[q].ctrl + W_FLAG_AR + P.rank, seq);
    }
    for (int q = 0; q < P.world; ++q) wait_flag(me.ctrl + W_FLAG_AR + q, seq);
  }
  __syncthreads();
  const float* src = me.slots + (long long)par * MAX_RANKS * P.ar_cap;
  for (long long i = tid; i < n; i += nthreads) {
    float s = __ldcg(src + i);
    for (int q = 1; q < P.world; ++q) s += __ldcg(src + (long long)q * P.ar_cap + i);   // rank order on every rank
    data[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(me.ctrl + W_DEPART, 1u) == gridDim.x - 1) {
    atomicExch(me.ctrl + W_DEPART, 0u);
    st_release_sys(me.ctrl + W_SEQ_AR, seq);
  }
}

static int fill_peers(Peers& P, void* const* windows, int rank, int world, long long halo_cap, long long ar_cap) {
  if (!(world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world && halo_cap > 0 && ar_cap > 0 && halo_cap % 4 == 0)) return 1;
  P.rank = rank;
  P.world = world;
  P.halo_cap = halo_cap;
  P.ar_cap = ar_cap;
  for (int q = 0; q < world; ++q) {
    if (windows[q] == nullptr || (reinterpret_cast<uintptr_t>(windows[q]) & 255) != 0) return 1;
    uint8_t* base = reinterpret_cast<uint8_t*>(windows[q]);
    P.w[q].ctrl = reinterpret_cast<uint32_t*>(base);
    P.w[q].mail = reinterpret_cast<float*>(base + 256);
    P.w[q].slots = P.w[q].mail + 4 * halo_cap;
  }
  return 0;
}

}  // namespace peer
}  // namespace mphsir

using namespace mphsir;

// ---- window life cycle: a direct cudaMalloc (an IPC handle names a whole allocation, so the window cannot be a slice of the
// host framework's caching allocator), exported / opened through CUDA IPC ----------------------------------------------
extern "C" int mphsir_peer_window_alloc(size_t bytes, void** window) {
  MPHSIR_REQUIRE(window && bytes >= 256, "peer_window_alloc: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    set_error("peer_window_alloc: %s", cudaGetErrorString(e));
    return MPHSIR_ERR_CUDA;
  }
  *window = p;
  return MPHSIR_OK;
}

extern "C" int mphsir_peer_window_free(void* window) {
  if (window && cudaFree(window) != cudaSuccess) {
    set_error("peer_window_free: cudaFree failed");
    return MPHSIR_ERR_CUDA;
  }
  return MPHSIR_OK;
}

extern "C" int mphsir_peer_export(void* window, unsigned char* handle64) {
  MPHSIR_REQUIRE(window && handle64, "peer_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, window);
  if (e != cudaSuccess) {
    set_error("peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return MPHSIR_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  return MPHSIR_OK;
}

extern "C" int mphsir_peer_open(const unsigned char* handle64, void** mapped) {
  MPHSIR_REQUIRE(handle64 && mapped, "peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);   // maps the peer GPU's memory over NVLink
  if (e != cudaSuccess) {
    set_error("peer_open: cudaIpcOpenMemHandle: %s (peers must be GPUs of one node with P2P access)", cudaGetErrorString(e));
    return MPHSIR_ERR_CUDA;
  }
  *mapped = p;
  return MPHSIR_OK;
}

extern "C" int mphsir_peer_close(void* mapped) {
  if (mapped && cudaIpcCloseMemHandle(mapped) != cudaSuccess) {
    set_error("peer_close: cudaIpcCloseMemHandle failed");
    return MPHSIR_ERR_CUDA;
  }
  return MPHSIR_OK;
}

extern "C" size_t mphsir_peer_window_bytes(long long halo_cap, long long ar_cap) {
  return 256 + sizeof(float) * (size_t)(4 * halo_cap + 2 * peer::MAX_RANKS * ar_cap);
}

extern "C" int mphsir_peer_halo_exchange(void* const* windows, int rank, int world, long long halo_cap, long long ar_cap,
                                         const float* own_first, const float* own_last, float* halo_top, float* halo_bottom,
                                         long long n, void* stream) {
  peer::Peers P;
  MPHSIR_REQUIRE(windows && peer::fill_peers(P, windows, rank, world, halo_cap, ar_cap) == 0, "peer_halo_exchange: bad window table");
  MPHSIR_REQUIRE(own_first && own_last && halo_top && halo_bottom && n > 0 && n % 4 == 0 && n <= halo_cap,
                 "peer_halo_exchange: %lld floats per direction do not fit the mailbox (%lld) or are not a multiple of 4", n, halo_cap);
  MPHSIR_REQUIRE(((reinterpret_cast<uintptr_t>(own_first) | reinterpret_cast<uintptr_t>(own_last) | reinterpret_cast<uintptr_t>(halo_top) |
                   reinterpret_cast<uintptr_t>(halo_bottom)) & 15) == 0, "peer_halo_exchange: rows must be 16-byte aligned");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 64) blocks = 64;   // all CTAs must be resident: they wait for the neighbour ranks
  peer::halo_exchange_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, own_first, own_last, halo_top,
                                                                                                  halo_bottom, n);
  return check_launch("peer_halo_exchange");
}

extern "C" int mphsir_peer_all_reduce(void* const* windows, int rank, int world, long long halo_cap, long long ar_cap, float* data,
                                      long long n, void* stream) {
  peer::Peers P;
  MPHSIR_REQUIRE(windows && peer::fill_peers(P, windows, rank, world, halo_cap, ar_cap) == 0, "peer_all_reduce: bad window table");
  MPHSIR_REQUIRE(data && n > 0 && n <= ar_cap, "peer_all_reduce: %lld floats do not fit the gather slot (%lld)", n, ar_cap);
  long long blocks = (n + 255) / 256;
  if (blocks > 32) blocks = 32;
  peer::all_reduce_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, data, n);
  return check_launch("peer_all_reduce");
}
