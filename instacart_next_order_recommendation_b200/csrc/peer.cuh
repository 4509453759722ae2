// Device side of the shard-candidate exchange over NVLink / NVSwitch peer memory, shared by the kernels that push their
// results straight to the peers: K4x (exchange.cu), the fused exchange + merge kernel (select_hist.cu) and the tail of the
// request-sized GEMV kernel (gemv_topk.cu). Buffer layout and protocol: exchange.cu.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace icr {

constexpr int kPeerHeaderBytes = 1024;
constexpr int kPeerTicketOff = 512;
constexpr int kPeerStatusOff = 768;  // uint32: epoch of an exchange that gave up waiting for a peer (0 = never)

// One rank's view of an exchange call: every rank's buffer as mapped into THIS process, the slot of this call's epoch.
struct PeerTail {
  unsigned char* peer_base[ICR_MAX_PEERS];
  size_t scores_off, ids_off;  // byte offsets of this epoch's slot regions inside a buffer
  uint64_t timeout_ns;         // how long a wait for the peers' flags may last
  uint32_t epoch;
  int rank, world;
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

// copy `bytes` (multiple of 4) from src to dst with threads tid, tid + nthreads, ...: 16-byte vectors where both are
// aligned, words otherwise
__device__ __forceinline__ void peer_copy(unsigned char* dst, const unsigned char* src, size_t bytes, size_t tid, size_t nthreads) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    const size_t nv = bytes / 16;
    for (size_t i = tid; i < nv; i += nthreads) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = nv * 4 + tid; i < bytes / 4; i += nthreads) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (size_t i = tid; i < bytes / 4; i += nthreads) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  }
}

// this rank's n (score, id) candidates -> slot `rank` of every buffer, the peers first (visited starting at rank + 1, so
// the ranks spread over the switch), the local copy last
__device__ __forceinline__ void peer_push(const PeerTail& p, const float* scores, const int64_t* ids, int64_t n, size_t tid, size_t nthreads) {
  for (int d = 1; d <= p.world; ++d) {
    unsigned char* base = p.peer_base[(p.rank + d) % p.world];
    peer_copy(base + p.scores_off + static_cast<size_t>(p.rank) * n * 4, reinterpret_cast<const unsigned char*>(scores), n * 4, tid, nthreads);
    peer_copy(base + p.ids_off + static_cast<size_t>(p.rank) * n * 8, reinterpret_cast<const unsigned char*>(ids), n * 8, tid, nthreads);
  }
}

// thread `t` < world tells rank t "my candidates of this epoch are in your buffer" (the caller made them visible at system
// scope before: __threadfence_system after the stores, a barrier, then this)
__device__ __forceinline__ void peer_publish(const PeerTail& p, int t) {
  if (t < p.world) st_release_sys(reinterpret_cast<uint32_t*>(p.peer_base[t]) + p.rank, p.epoch);
}

// thread `t` < world waits until rank t's candidates of this epoch are in the local buffer. Epochs only grow (the host counts
// calls), so >= also accepts a peer that is already one call ahead. A peer that never arrives must not hang the GPU for ever:
// after the limit (ICR_PEER_TIMEOUT_S, default 600 s - a rank may sit in a debugger, a page-cache miss or GC for minutes, as
// NCCL tolerates) the status word of the local header is set and the wait ends; the host reads it at its next synchronisation
// point and raises (PeerExchange.check).
__device__ __forceinline__ void peer_wait(const PeerTail& p, int t) {
  if (t >= p.world) return;
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(p.peer_base[p.rank]) + t;
  const uint64_t t0 = global_timer_ns();
  while (static_cast<int32_t>(ld_acquire_sys(mine) - p.epoch) < 0) {
    if (global_timer_ns() - t0 > p.timeout_ns) {
      *reinterpret_cast<volatile uint32_t*>(p.peer_base[p.rank] + kPeerStatusOff) = p.epoch;
      break;
    }
  }
}

// host: fills a PeerTail for one call (exchange.cu)
void peer_layout(int64_t n_max, int world, uint32_t epoch, size_t* scores_off, size_t* ids_off, size_t* total);
PeerTail make_peer_tail(int rank, int world, const uint64_t* peer_buffers, uint32_t epoch, int64_t n_max);

}  // namespace icr
