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
      for (int u = 0; u < 4; ++u) rank[u] += (kj > my[u] || (kj == my[u] && j < base + 32 * u)) ? 1 : 0;  // the index orders exact duplicates
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
    if (range == 0) {  // every key equal (duplicated candidates, e.g. replicated shards fed to the merge): any `need` of them
      for (int i = lane; i < need; i += 32) sel[nsel + i] = buf[i];
      __syncwarp();
      return;
    }
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


}  // namespace icr
