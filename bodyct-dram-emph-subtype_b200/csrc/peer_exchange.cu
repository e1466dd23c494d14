// SyncBatchNorm's statistics exchange over NVLink peer memory (one process per GPU, one node).
//
// The reference trains with `sync_batchnorm=True` (train.py:101): every train-mode BatchNorm3d all-reduces its
// per-channel sums across the data-parallel ranks, forward and backward — 114 latency-bound collectives of a few KB per
// ResNet-34 step, each on the critical path between K10's reduction and its apply pass.  Instead of a library
// collective this is one small kernel per exchange:
//   * every rank owns one exchange buffer (cudaMalloc + CUDA IPC, mapped by all peers once at start-up);
//   * the kernel PUSHES its 2*C fp64 sums into its row of every peer's buffer with plain stores over NVLink,
//     fences, raises a per-(slot, source rank) flag on each peer with st.release.sys, then spins with ld.acquire.sys
//     until its own flags show that every peer's row has landed, and adds the rows in rank order 0..W-1 — the same
//     order on every rank, so all ranks hold bit-identical statistics (and weights) afterwards;
//   * flags carry the call's sequence number and four slots are used in rotation, so nothing is ever reset: a rank can
//     only run one exchange ahead of the slowest peer (it needs that peer's flag for the next one), which never
//     touches the slot the peer is still reading.
// A rank that waits longer than ~60 s sets *status, stops waiting and lets every later call fall through, so a lost
// peer cannot hang the GPU.
#include <string.h>

#include "common.h"

namespace dram {

constexpr int PX_SLOTS = 4;
constexpr int PX_MAX_WORLD = 8;
constexpr int PX_MAX_N = 4096;  // doubles per exchange: 2 * C, C <= 2048
constexpr size_t PX_DATA_BYTES = (size_t)PX_SLOTS * PX_MAX_WORLD * PX_MAX_N * sizeof(double);
constexpr size_t PX_FLAG_BYTES = (size_t)PX_SLOTS * PX_MAX_WORLD * sizeof(unsigned long long);
constexpr long long PX_TIMEOUT_CYCLES = 120LL * 1000 * 1000 * 1000;  // ~60 s at 2 GHz

struct PeerPtrs {
  void *p[PX_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024) peer_allreduce_f64_kernel(const double *in, double *out, int n,
                                                                  PeerPtrs peers, int world, int rank,
                                                                  unsigned long long seq, int *__restrict__ status) {
  const int tid = threadIdx.x;
  __shared__ int dead;
  if (tid == 0) dead = *reinterpret_cast<volatile int *>(status);
  __syncthreads();
  if (dead) {  // an earlier exchange timed out: the job is lost, do not wait again
    for (int i = tid; i < n; i += blockDim.x) out[i] = in[i];
    return;
  }
  const int slot = (int)(seq % PX_SLOTS);
  const size_t row = ((size_t)slot * PX_MAX_WORLD + rank) * PX_MAX_N;
  for (int p = 0; p < world; ++p) {
    double *dst = reinterpret_cast<double *>(peers.p[p]) + row;
    for (int i = tid; i < n; i += blockDim.x) dst[i] = in[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    unsigned long long *flags = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(peers.p[tid]) + PX_DATA_BYTES);
    st_release_sys(flags + slot * PX_MAX_WORLD + rank, seq);
  }
  if (tid < world) {
    const unsigned long long *mine =
        reinterpret_cast<const unsigned long long *>(reinterpret_cast<const char *>(peers.p[rank]) + PX_DATA_BYTES) +
        slot * PX_MAX_WORLD + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < seq) {
      if (clock64() - t0 > PX_TIMEOUT_CYCLES) {
        atomicExch(status, 1 + tid);
        break;
      }
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
  const double *base = reinterpret_cast<const double *>(peers.p[rank]) + (size_t)slot * PX_MAX_WORLD * PX_MAX_N;
  for (int i = tid; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += __ldcg(base + (size_t)r * PX_MAX_N + i);  // rows were written by the peers: L2
    out[i] = s;
  }
}

}  // namespace dram

using namespace dram;

extern "C" int64_t dram_peer_exchange_bytes(void) { return (int64_t)(PX_DATA_BYTES + PX_FLAG_BYTES); }

extern "C" int dram_peer_alloc(int64_t bytes, void **ptr, void *handle64) {
  DRAM_REQUIRE(ptr && handle64 && bytes > 0, "dram_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == DRAM_PEER_HANDLE_BYTES, "CUDA IPC handle size");
  void *p = nullptr;
  int rc = check_cuda(cudaMalloc(&p, (size_t)bytes), "dram_peer_alloc: cudaMalloc");
  if (rc != DRAM_OK) return rc;
  rc = check_cuda(cudaMemset(p, 0, (size_t)bytes), "dram_peer_alloc: cudaMemset");
  if (rc == DRAM_OK) rc = check_cuda(cudaDeviceSynchronize(), "dram_peer_alloc: cudaDeviceSynchronize");
  cudaIpcMemHandle_t h;
  if (rc == DRAM_OK) rc = check_cuda(cudaIpcGetMemHandle(&h, p), "dram_peer_alloc: cudaIpcGetMemHandle");
  if (rc != DRAM_OK) {
    cudaFree(p);
    return rc;
  }
  memcpy(handle64, &h, sizeof(h));
  *ptr = p;
  return DRAM_OK;
}

extern "C" int dram_peer_open(const void *handle64, void **ptr) {
  DRAM_REQUIRE(handle64 && ptr, "dram_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return check_cuda(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "dram_peer_open: cudaIpcOpenMemHandle");
}

extern "C" int dram_peer_close(void *ptr) {
  DRAM_REQUIRE(ptr, "dram_peer_close: null pointer");
  return check_cuda(cudaIpcCloseMemHandle(ptr), "dram_peer_close: cudaIpcCloseMemHandle");
}

extern "C" int dram_peer_free(void *ptr) {
  DRAM_REQUIRE(ptr, "dram_peer_free: null pointer");
  return check_cuda(cudaFree(ptr), "dram_peer_free: cudaFree");
}

extern "C" int dram_peer_allreduce_f64(const double *in, double *out, int32_t n, void *const *peers, int32_t world,
                                       int32_t rank, uint64_t seq, int32_t *status, void *stream) {
  DRAM_REQUIRE(in && out && peers && status, "dram_peer_allreduce_f64: null pointer");
  DRAM_REQUIRE(world >= 1 && world <= PX_MAX_WORLD && rank >= 0 && rank < world,
               "dram_peer_allreduce_f64: rank %d of %d (at most %d ranks, one node)", rank, world, PX_MAX_WORLD);
  DRAM_REQUIRE(n > 0 && n <= PX_MAX_N, "dram_peer_allreduce_f64: %d values per exchange (at most %d)", n, PX_MAX_N);
  DRAM_REQUIRE(seq >= 1, "dram_peer_allreduce_f64: the sequence number counts from 1");
  PeerPtrs pp;
  for (int r = 0; r < PX_MAX_WORLD; ++r) {
    pp.p[r] = r < world ? peers[r] : nullptr;
    DRAM_REQUIRE(r >= world || pp.p[r], "dram_peer_allreduce_f64: peer %d has no buffer", r);
  }
  peer_allreduce_f64_kernel<<<1, n >= 1024 ? 1024 : ((n + 31) / 32) * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, out, n, pp, world, rank, (unsigned long long)seq, status);
  DRAM_CHECK_LAUNCH("peer_allreduce_f64_kernel launch");
  return DRAM_OK;
}
