// Warp-level exact selection of the k largest 64-bit candidate keys held in shared memory
// (histogram partitioning, no full sort). Shared by select_hist.cu and the in-kernel merge of gemv_topk.cu.
#pragma once

#include "common.cuh"

namespace icr {

constexpr int kHsBins = 256;

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t w = __shfl_xor_sync(kFull, v, o);
    v = w < v ? w : v;
  }
  return v;
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t w = __shfl_xor_sync(kFull, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// out[r] = the key of rank r (0 = largest) among in[0..n), for r < kout. Keys are distinct; `in` and `out` must not
// overlap. Rank counting: every lane compares its own keys (four at a time) against all n keys read as shared-memory
// broadcasts - n*n/32 comparisons but no barriers, no divergent stores and full ILP, which beats a bitonic network
// in shared memory for the list lengths met here (n <= a few hundred).
__device__ __forceinline__ void warp_rank_sort_desc(const uint64_t* in, int n, uint64_t* out, int kout, int lane) {
  __syncwarp();
  for (int base = lane; base < n; base += 128) {
    uint64_t my[4];
    int rank[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      my[u] = (base + 32 * u < n) ? in[base + 32 * u] : ~0ull;
      rank[u] = 0;
    }
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const uint64_t kj = in[j];
#pragma unroll
      for (int u = 0; u < 4; ++u) rank[u] += (kj > my[u]) ? 1 : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (base + 32 * u < n && rank[u] < kout) out[rank[u]] = my[u];
  }
  __syncwarp();
}

// Moves the k largest of buf[0..n) to sel[0..k) (unordered). n > k on entry. Destroys buf.
__device__ __forceinline__ void warp_select_topk(uint64_t* buf, int n, int k, uint64_t* sel, unsigned int* hist, int lane) {
  int need = k, nsel = 0, len = n;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (;;) {
    // ---- key range of the current list ----
    uint64_t mn = ~0ull, mx = 0ull;
#pragma unroll 4
    for (int i = lane; i < len; i += 32) {
      const uint64_t key = buf[i];
      mn = key < mn ? key : mn;
      mx = key > mx ? key : mx;
    }
    mn = warp_min_u64(mn);
    mx = warp_max_u64(mx);
    const uint64_t range = mx - mn;
    const int bits = 64 - __clzll(static_cast<long long>(range | 1ull));
    const int shift = bits > 8 ? bits - 8 : 0;  // (key - mn) >> shift  in [0, 255]
    // ---- histogram, indexed from the top: t = 255 - bin ----
    for (int i = lane; i < kHsBins; i += 32) hist[i] = 0u;
    __syncwarp();
#pragma unroll 4
    for (int i = lane; i < len; i += 32) atomicAdd(&hist[255 - static_cast<int>((buf[i] - mn) >> shift)], 1u);
    __syncwarp();
    // lane L owns t in [8L, 8L+8): cumulative counts from the top bin downwards
    unsigned int local[8];
    unsigned int lsum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      local[j] = hist[lane * 8 + j];
      lsum += local[j];
    }
    unsigned int incl = lsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned int excl = incl - lsum;
    // the boundary bin is the first t where the cumulative count reaches `need`
    int t_star = -1;
    unsigned int above = 0;
    if (excl < static_cast<unsigned int>(need) && static_cast<unsigned int>(need) <= incl) {
      unsigned int c = excl;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (t_star < 0 && c + local[j] >= static_cast<unsigned int>(need)) {
          t_star = lane * 8 + j;
          above = c;
        }
        c += local[j];
      }
    }
    const unsigned owner = __ballot_sync(kFull, t_star >= 0);
    const int src = __ffs(owner) - 1;
    t_star = __shfl_sync(kFull, t_star, src);
    above = __shfl_sync(kFull, above, src);
    const int b_star = 255 - t_star;
    // ---- classify: above the boundary bin -> selected; in it -> compacted in place; below -> dropped ----
    int nb = 0;
#pragma unroll 2
    for (int base = 0; base < len; base += 32) {
      const int i = base + lane;
      const bool valid = i < len;
      const uint64_t key = valid ? buf[i] : 0ull;
      const int bin = valid ? static_cast<int>((key - mn) >> shift) : -1;
      const bool in = bin > b_star, bnd = bin == b_star;
      const unsigned m_in = __ballot_sync(kFull, in), m_b = __ballot_sync(kFull, bnd);
      if (in) sel[nsel + __popc(m_in & lt_mask)] = key;
      // all 32 reads of this round are done (ballot is a warp barrier); writes land at positions <= base
      if (bnd) buf[nb + __popc(m_b & lt_mask)] = key;
      nsel += __popc(m_in);
      nb += __popc(m_b);
    }
    __syncwarp();
    need -= static_cast<int>(above);  // 1 <= need <= nb
    if (need == nb) {
      for (int i = lane; i < nb; i += 32) sel[nsel + i] = buf[i];
      __syncwarp();
      return;
    }
    if (nb <= 128) {
      warp_rank_sort_desc(buf, nb, sel + nsel, need, lane);
      return;
    }
    len = nb;  // refine inside the boundary bin (its key range is 256x narrower)
  }
}


// ---- the same selection with per-lane BYTE histograms instead of shared-memory atomics ---------------------------------
// Shared-memory atomics retire at ~2 cycles per lane (64 cycles per warp instruction with spread addresses), which made the
// histogram pass the limit of the warp-per-query select kernel (profiles/r02_notes.md). Here every lane counts its own keys
// in its own byte column: counter (t, lane) lives at hist8[t * 32 + lane], t = 127 - bin (top bin first), 128 bins, so a
// pass is plain LDS.U8 / STS.U8 with no contention, and the column sums are eight 128-bit loads per lane.
// Needs len <= 32 * 255 keys (the callers buffer <= 1024) and hist8 = 4096 bytes, 16-byte aligned.
constexpr int kB8Bins = 128;
constexpr int kB8Bytes = kB8Bins * 32;

__device__ __forceinline__ unsigned int b8_sum32(const unsigned char* p) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 16);
  unsigned int s = 0;
  s = __dp4a(a.x, 0x01010101u, s);
  s = __dp4a(a.y, 0x01010101u, s);
  s = __dp4a(a.z, 0x01010101u, s);
  s = __dp4a(a.w, 0x01010101u, s);
  s = __dp4a(b.x, 0x01010101u, s);
  s = __dp4a(b.y, 0x01010101u, s);
  s = __dp4a(b.z, 0x01010101u, s);
  s = __dp4a(b.w, 0x01010101u, s);
  return s;
}

// Moves the k largest of buf[0..n) to sel[0..k) (unordered). n > k on entry. Destroys buf.
__device__ __forceinline__ void warp_select_topk_b8(uint64_t* buf, int n, int k, uint64_t* sel, unsigned char* hist8, int lane) {
  int need = k, nsel = 0, len = n;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (;;) {
    uint64_t mn = ~0ull, mx = 0ull;
#pragma unroll 4
    for (int i = lane; i < len; i += 32) {
      const uint64_t key = buf[i];
      mn = key < mn ? key : mn;
      mx = key > mx ? key : mx;
    }
    mn = warp_min_u64(mn);
    mx = warp_max_u64(mx);
    const uint64_t range = mx - mn;
    if (range == 0) {  // all keys equal (duplicated candidates): any `need` of them
      for (int i = lane; i < need; i += 32) sel[nsel + i] = buf[i];
      __syncwarp();
      return;
    }
    const int bits = 64 - __clzll(static_cast<long long>(range));
    const int shift = bits > 7 ? bits - 7 : 0;  // (key - mn) >> shift  in [0, 127]
    uint4* h4 = reinterpret_cast<uint4*>(hist8);
#pragma unroll
    for (int i = 0; i < kB8Bytes / 16 / 32; ++i) h4[i * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    unsigned char* mine = hist8 + lane;
#pragma unroll 4
    for (int i = lane; i < len; i += 32) {
      const int t = (kB8Bins - 1) - static_cast<int>((buf[i] - mn) >> shift);
      mine[t * 32] += 1;
    }
    __syncwarp();
    // lane L owns t in {L, L+32, L+64, L+96}: cumulative counts from the top bin downwards
    int t_star = -1;
    unsigned int above = 0, base = 0;
#pragma unroll
    for (int j = 0; j < kB8Bins / 32; ++j) {
      const unsigned int c = b8_sum32(hist8 + (j * 32 + lane) * 32);
      unsigned int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      const unsigned int lo = base + incl - c, hi = base + incl;
      if (t_star < 0 && lo < static_cast<unsigned int>(need) && static_cast<unsigned int>(need) <= hi) {
        t_star = j * 32 + lane;
        above = lo;
      }
      base += __shfl_sync(kFull, incl, 31);
    }
    const unsigned owner = __ballot_sync(kFull, t_star >= 0);
    const int src = __ffs(owner) - 1;
    t_star = __shfl_sync(kFull, t_star, src);
    above = __shfl_sync(kFull, above, src);
    const int b_star = (kB8Bins - 1) - t_star;
    int nb = 0;
#pragma unroll 2
    for (int b0 = 0; b0 < len; b0 += 32) {
      const int i = b0 + lane;
      const bool valid = i < len;
      const uint64_t key = valid ? buf[i] : 0ull;
      const int bin = valid ? static_cast<int>((key - mn) >> shift) : -1;
      const bool in = bin > b_star, bnd = bin == b_star;
      const unsigned m_in = __ballot_sync(kFull, in), m_b = __ballot_sync(kFull, bnd);
      if (in) sel[nsel + __popc(m_in & lt_mask)] = key;
      if (bnd) buf[nb + __popc(m_b & lt_mask)] = key;  // every read of this round is done (ballot); writes land at <= b0
      nsel += __popc(m_in);
      nb += __popc(m_b);
    }
    __syncwarp();
    need -= static_cast<int>(above);  // 1 <= need <= nb
    if (need == nb) {
      for (int i = lane; i < nb; i += 32) sel[nsel + i] = buf[i];
      __syncwarp();
      return;
    }
    if (nb <= 128) {
      warp_rank_sort_desc(buf, nb, sel + nsel, need, lane);
      return;
    }
    len = nb;
  }
}

}  // namespace icr
