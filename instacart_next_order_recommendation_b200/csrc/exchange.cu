// K4x: shard-candidate exchange over NVLink / NVSwitch peer memory (the all-gather of SURVEY §8e without NCCL).
//
// Every rank owns one symmetric buffer (allocated and mapped into all peers by the host: torch symmetric memory is the
// plumbing); `peer_base[p]` is rank p's buffer as seen from THIS process. Layout of a buffer:
//   [0, 256)      uint32 flag[src]: last epoch whose candidates from rank `src` have fully arrived here
//   [512, 516)    ticket counter of the local push kernel
//   [768, 772)    status: epoch of a call whose wait for a peer timed out (host-visible; PeerExchange.check raises on it)
//   [1024, ...)   two slots (epoch parity), each  scores f32 [G][n_max]  |  ids i64 [G][n_max]   (used compactly: [G][n])
// One kernel per call: the CTAs store this rank's [n] (score, id) candidates into slot `rank` of EVERY peer's buffer with
// 16-byte stores over NVLink (peers visited starting at rank+1, so the G ranks spread over the switch), fence at system
// scope and take a ticket; the last CTA publishes `epoch` into flag[rank] of every peer (st.release.sys) and then waits
// (ld.acquire.sys) until its own flags show that all G ranks' candidates for this epoch are here. The merge kernel that
// follows on the stream reads the local slot. Two slots are enough: a rank can only pass the wait of epoch e+1 after
// every peer pushed e+1, which each peer enqueues behind its own merge of epoch e.
//
// No reference counterpart (the reference is single-GPU); stands in for the ncclAllGather of §8(e).
#include <stdlib.h>

#include "common.cuh"

namespace icr {

constexpr int kPeerHeaderBytes = 1024;
constexpr int kPeerTicketOff = 512;
constexpr int kPeerStatusOff = 768;  // uint32: epoch of an exchange that gave up waiting for a peer (0 = never)

struct PeerArgs {
  const float* scores;
  const int64_t* ids;
  int64_t n;
  int rank, world;
  unsigned char* peer_base[ICR_MAX_PEERS];
  size_t scores_off, ids_off;  // byte offsets of this epoch's slot regions inside a buffer
  uint32_t epoch;
  uint64_t timeout_ns;  // how long the last CTA waits for the peers' flags
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// copy `bytes` (multiple of 4) from src to dst: 16-byte vectors where both are aligned, words otherwise
__device__ __forceinline__ void cta_copy(unsigned char* dst, const unsigned char* src, size_t bytes, int part, int parts) {
  const size_t tid = static_cast<size_t>(part) * blockDim.x + threadIdx.x, nthreads = static_cast<size_t>(parts) * blockDim.x;
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    const size_t nv = bytes / 16;
    for (size_t i = tid; i < nv; i += nthreads) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = nv * 4 + tid; i < bytes / 4; i += nthreads) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (size_t i = tid; i < bytes / 4; i += nthreads) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  }
}

__global__ void __launch_bounds__(256) peer_exchange_kernel(const PeerArgs a) {
  __shared__ bool is_last;
  for (int d = 1; d <= a.world; ++d) {
    const int p = (a.rank + d) % a.world;  // self last: the local copy needs no link
    unsigned char* base = a.peer_base[p];
    cta_copy(base + a.scores_off + static_cast<size_t>(a.rank) * a.n * 4, reinterpret_cast<const unsigned char*>(a.scores), a.n * 4, blockIdx.x, gridDim.x);
    cta_copy(base + a.ids_off + static_cast<size_t>(a.rank) * a.n * 8, reinterpret_cast<const unsigned char*>(a.ids), a.n * 8, blockIdx.x, gridDim.x);
  }
  __threadfence_system();
  __syncthreads();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(a.peer_base[a.rank] + kPeerTicketOff);
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence_system();
  if (threadIdx.x < a.world) {
    const int p = threadIdx.x;
    st_release_sys(reinterpret_cast<uint32_t*>(a.peer_base[p]) + a.rank, a.epoch);  // "rank's candidates of `epoch` are in p's buffer"
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.peer_base[a.rank]) + p;
    const uint64_t t0 = global_timer_ns();
    // epochs only grow (the host counts calls), so >= also accepts a peer that is already one call ahead
    while (static_cast<int32_t>(ld_acquire_sys(mine) - a.epoch) < 0) {
      // a peer that never arrives must not hang the GPU for ever: after the limit (ICR_PEER_TIMEOUT_S, default 600 s - a rank may
      // sit in a debugger, a page-cache miss or GC for minutes, as NCCL tolerates) the status word of the local header is set and
      // the launch ends; the host reads it at its next synchronisation point and raises (PeerExchange.check)
      if (global_timer_ns() - t0 > a.timeout_ns) {
        *reinterpret_cast<volatile uint32_t*>(a.peer_base[a.rank] + kPeerStatusOff) = a.epoch;
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *ticket = 0u;
}

void peer_layout(int64_t n_max, int world, uint32_t epoch, size_t* scores_off, size_t* ids_off, size_t* total) {
  const size_t sc = align_up(static_cast<size_t>(world) * n_max * 4, 256), id = align_up(static_cast<size_t>(world) * n_max * 8, 256);
  const size_t slot = sc + id;
  if (scores_off) *scores_off = kPeerHeaderBytes + (epoch & 1u) * slot;
  if (ids_off) *ids_off = kPeerHeaderBytes + (epoch & 1u) * slot + sc;
  if (total) *total = kPeerHeaderBytes + 2 * slot;
}

int launch_peer_exchange(const float* scores, const int64_t* ids, int64_t n, int rank, int world, const uint64_t* peer_buffers, uint32_t epoch,
                         int64_t n_max, cudaStream_t st) {
  PeerArgs a{};
  a.scores = scores;
  a.ids = ids;
  a.n = n;
  a.rank = rank;
  a.world = world;
  for (int p = 0; p < world; ++p) a.peer_base[p] = reinterpret_cast<unsigned char*>(static_cast<uintptr_t>(peer_buffers[p]));
  peer_layout(n_max, world, epoch, &a.scores_off, &a.ids_off, nullptr);
  a.epoch = epoch;
  static const uint64_t timeout_s = getenv("ICR_PEER_TIMEOUT_S") ? strtoull(getenv("ICR_PEER_TIMEOUT_S"), nullptr, 10) : 600ull;
  a.timeout_ns = (timeout_s ? timeout_s : 1ull) * 1000000000ull;
  // 12 bytes per candidate to every peer: enough CTAs to keep the links busy for large batches, one for a request
  int64_t ctas = (n * 12 * world + (64 << 10) - 1) / (64 << 10);
  ctas = ctas < 1 ? 1 : (ctas > 64 ? 64 : ctas);
  peer_exchange_kernel<<<static_cast<unsigned>(ctas), 256, 0, st>>>(a);
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

}  // namespace icr
