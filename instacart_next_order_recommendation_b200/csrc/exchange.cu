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

#include "peer.cuh"

namespace icr {

struct PeerArgs {
  const float* scores;
  const int64_t* ids;
  int64_t n;
  PeerTail peer;
};

__global__ void __launch_bounds__(256) peer_exchange_kernel(const PeerArgs a) {
  __shared__ bool is_last;
  const PeerTail& p = a.peer;
  peer_push(p, a.scores, a.ids, a.n, static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x, static_cast<size_t>(gridDim.x) * blockDim.x);
  __threadfence_system();
  __syncthreads();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(p.peer_base[p.rank] + kPeerTicketOff);
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence_system();
  peer_publish(p, threadIdx.x);
  peer_wait(p, threadIdx.x);
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

PeerTail make_peer_tail(int rank, int world, const uint64_t* peer_buffers, uint32_t epoch, int64_t n_max) {
  PeerTail p{};
  p.rank = rank;
  p.world = world;
  for (int i = 0; i < world; ++i) p.peer_base[i] = reinterpret_cast<unsigned char*>(static_cast<uintptr_t>(peer_buffers[i]));
  peer_layout(n_max, world, epoch, &p.scores_off, &p.ids_off, nullptr);
  p.epoch = epoch;
  static const uint64_t timeout_s = getenv("ICR_PEER_TIMEOUT_S") ? strtoull(getenv("ICR_PEER_TIMEOUT_S"), nullptr, 10) : 600ull;
  p.timeout_ns = (timeout_s ? timeout_s : 1ull) * 1000000000ull;
  return p;
}

int launch_peer_exchange(const float* scores, const int64_t* ids, int64_t n, int rank, int world, const uint64_t* peer_buffers, uint32_t epoch,
                         int64_t n_max, cudaStream_t st) {
  PeerArgs a{};
  a.scores = scores;
  a.ids = ids;
  a.n = n;
  a.peer = make_peer_tail(rank, world, peer_buffers, epoch, n_max);
  // 12 bytes per candidate to every peer: enough CTAs to keep the links busy for large batches, one for a request
  int64_t ctas = (n * 12 * world + (64 << 10) - 1) / (64 << 10);
  ctas = ctas < 1 ? 1 : (ctas > 64 ? 64 : ctas);
  peer_exchange_kernel<<<static_cast<unsigned>(ctas), 256, 0, st>>>(a);
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

}  // namespace icr
