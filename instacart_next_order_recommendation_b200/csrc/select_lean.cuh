// Lean warp-level selection on 32-bit (score bits, ~row) arrays in shared memory: shared by the select kernels
// (select_hist.cu) and the in-kernel threshold bootstrap of the swapped GEMM kernel (gemm_topk.cu).
#pragma once

#include "common.cuh"
#include "select_warp.cuh"

namespace icr {

// =====================================================================================================================
// Warp-per-query selection, lean form. Keys live in shared memory as two 32-bit arrays - sc[] = order-preserving score
// bits, rw[] = ~row - so every pass works on 32-bit values: a 256-bin histogram over the score range finds the bin of the
// k-th best key, the few keys of that bin are ranked exactly as (score, ~row) pairs, and ONE in-place compaction keeps
// every key >= the threshold key (the k-th key itself, or `band` below its score for screened keys). ~35 thread
// instructions per key instead of ~300 for the 64-bit multi-pass version it replaces (profiles/r02_notes.md).
// =====================================================================================================================
#if defined(ICR_SELECT_TRACE) && defined(ICR_ST_OWNER)  // development builds only (ICR_NVCC_DEFS=-DICR_SELECT_TRACE): cycles per section of the warp select, summed over warps
__device__ unsigned long long g_sel_trace[16];
#define ICR_ST_BEGIN() long long st_t0_ = clock64()
#define ICR_ST_MARK(i)                                                      \
  do {                                                                      \
    const long long t_ = clock64();                                         \
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_sel_trace[i], static_cast<unsigned long long>(t_ - st_t0_)); \
    st_t0_ = clock64();                                                     \
  } while (0)
#else
#define ICR_ST_BEGIN() do {} while (0)
#define ICR_ST_MARK(i) do {} while (0)
#endif

constexpr int kLsCap = 1024;   // keys buffered per query (8 KB of shared memory as two 32-bit arrays)
constexpr int kLsList = 64;    // boundary-bin keys ranked directly, held as (sc, rw) pairs in hist[0..128)
constexpr int kLsB = 8;        // keys per lane per batch of a pass

__device__ __forceinline__ bool pair_gt(uint32_t s1, uint32_t r1, uint32_t s2, uint32_t r2) { return s1 > s2 || (s1 == s2 && r1 > r2); }
__device__ __forceinline__ bool pair_ge(uint32_t s1, uint32_t r1, uint32_t s2, uint32_t r2) { return s1 > s2 || (s1 == s2 && r1 >= r2); }

// Every pass walks the keys in batches of 8 per lane (key c + lane + 32 u, u < 8): the eight shared-memory loads of a batch
// are issued together and its eight iterations are independent instructions. One key per iteration made each pass a chain of
// dependent LDS -> ballot -> add steps (~11 cycles per issued instruction at 16 warps per SM); holding ALL keys in registers
// (32 unrolled slots) removed the chains but made the kernel 400 KB of code that thrashed the instruction cache - both
// measured in profiles/r02_notes.md. Empty slots read as 0: no finite score has order bits 0.
__device__ __forceinline__ void ls_batch(const uint32_t* a, int n, int c, int lane, uint32_t (&v)[kLsB]) {
#pragma unroll
  for (int u = 0; u < kLsB; ++u) {
    const int i = c + lane + 32 * u;
    v[u] = i < n ? a[i] : 0u;
  }
}

// cumulative scan of the 256-bin histogram from the top bin down: the bin holding the `need`-th largest key, the count
// above it (subtracted from need) and its own population
__device__ __forceinline__ int ls_scan(const uint32_t* hist, int& need, int& cnt, int lane) {
  unsigned int local[8], lsum = 0;
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
  int t_star = -1;
  unsigned int above = 0, c_star = 0;
  if (excl < static_cast<unsigned int>(need) && static_cast<unsigned int>(need) <= incl) {
    unsigned int c = excl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (t_star < 0 && c + local[j] >= static_cast<unsigned int>(need)) {
        t_star = lane * 8 + j;
        above = c;
        c_star = local[j];
      }
      c += local[j];
    }
  }
  const unsigned owner = __ballot_sync(kFull, t_star >= 0);
  const int src = __ffs(owner) - 1;
  t_star = __shfl_sync(kFull, t_star, src);
  need -= static_cast<int>(__shfl_sync(kFull, above, src));
  cnt = static_cast<int>(__shfl_sync(kFull, c_star, src));
  __syncwarp();
  return 255 - t_star;
}

// One radix level over the bit patterns of val[] (the scores, or ~row among the keys whose score equals feq) inside the
// window [lo, hi] (hi - lo < 256 << shift): returns the bin, counted from the window's lower edge, of the need-th largest.
// Cold path (crowded boundary bins only): not inlined.
static __device__ __noinline__ int ls_level(const uint32_t* val, const uint32_t* fsc, uint32_t feq, int n, uint32_t lo, uint32_t hi, int shift, int* need_io,
                                     int* cnt_out, uint32_t* hist) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kHsBins / 32; ++i) hist[i * 32 + lane] = 0u;
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const uint32_t v = val[i];
    if (v >= lo && v <= hi && (fsc == nullptr || fsc[i] == feq)) atomicAdd(&hist[255u - ((v - lo) >> shift)], 1u);
  }
  __syncwarp();
  int need = *need_io, cnt = 0;
  const int b = ls_scan(hist, need, cnt, lane);
  *need_io = need;
  *cnt_out = cnt;
  return b;
}

// the need-th largest (1-based) of the nb (sc, rw) pairs held in list[0 .. 2 nb) by rank counting (broadcast reads)
__device__ __forceinline__ void ls_rank_list(const uint32_t* list, int nb, int need, int lane, uint32_t& ks, uint32_t& kr) {
  ks = 0u;
  kr = 0u;
  bool found = false;
  for (int j0 = 0; j0 < nb; j0 += 32) {
    const int j = j0 + lane;
    uint32_t s1 = 0, r1 = 0;
    int rank = -1;
    if (j < nb) {
      s1 = list[2 * j];
      r1 = list[2 * j + 1];
      rank = 0;
      for (int t = 0; t < nb; ++t) {
        const uint32_t s2 = list[2 * t], r2 = list[2 * t + 1];
        rank += (pair_gt(s2, r2, s1, r1) || (t < j && s2 == s1 && r2 == r1)) ? 1 : 0;  // the index breaks identical pairs
      }
    }
    const unsigned hit = __ballot_sync(kFull, rank == need - 1);
    if (hit && !found) {
      const int srcl = __ffs(hit) - 1;
      ks = __shfl_sync(kFull, s1, srcl);
      kr = __shfl_sync(kFull, r1, srcl);
      found = true;
    }
  }
  __syncwarp();
}

// The k-th largest (sc, rw) pair of the n >= k >= 1 keys. hist doubles as the list of the boundary bin's keys.
//
// Level 1 bins the keys linearly in score VALUE (not in their bit patterns: float bits are logarithmic in the value, and a
// dense first phase with scores on both sides of zero would put a third of all keys into the one bin that holds the k-th).
// That leaves ~n/256 x (local density / mean density) keys in the boundary bin - a handful - which are ranked directly as
// (score, ~row) pairs. Only if that bin is crowded (> kLsList keys: near-duplicate scores) do further levels refine it,
// now linearly in the bit pattern over the bin's own narrow range, and finally over ~row among keys of ONE score.
__device__ __forceinline__ void ls_kth(const uint32_t* sc, const uint32_t* rw, int n, int k, uint32_t* hist, int lane, uint32_t& ks, uint32_t& kr) {
  const unsigned lt = (1u << lane) - 1u;
  ICR_ST_BEGIN();
  uint32_t mn = 0xFFFFFFFFu, mx = 0u;
  for (int c = 0; c < n; c += 32 * kLsB) {
    uint32_t v[kLsB];
    ls_batch(sc, n, c, lane, v);
#pragma unroll
    for (int u = 0; u < kLsB; ++u) {
      mn = min(mn, v[u] ? v[u] : 0xFFFFFFFFu);
      mx = max(mx, v[u]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(kFull, mn, o));
    mx = max(mx, __shfl_xor_sync(kFull, mx, o));
  }
  ICR_ST_MARK(0);
  int need = k, cnt = n;
  uint32_t blo = mn, bhi = mx;  // the boundary bin covers score bits [blo, bhi]
  if (mn != mx) {
    // ---- level 1: 256 bins, linear in the score value ----
    const float fmn = unorder_bits(mn);
    const float scale = 256.0f / (unorder_bits(mx) - fmn);
#pragma unroll
    for (int i = 0; i < kHsBins / 32; ++i) hist[i * 32 + lane] = 0u;
    __syncwarp();
    for (int c = 0; c < n; c += 32 * kLsB) {
      uint32_t v[kLsB];
      ls_batch(sc, n, c, lane, v);
#pragma unroll
      for (int u = 0; u < kLsB; ++u) {
        const int b = max(min(255, __float2int_rd((unorder_bits(v[u]) - fmn) * scale)), 0);
        if (v[u] != 0u) atomicAdd(&hist[255 - b], 1u);
      }
    }
    __syncwarp();
    ICR_ST_MARK(1);
    const int b_star = ls_scan(hist, need, cnt, lane);
    ICR_ST_MARK(2);
    if (cnt <= kLsList) {
      // the usual case: the bin's few keys are collected by bin number, no bit range needed
      const unsigned lt0 = (1u << lane) - 1u;
      int nb0 = 0;
      for (int c = 0; c < n; c += 32 * kLsB) {
        uint32_t v[kLsB], r[kLsB];
        ls_batch(sc, n, c, lane, v);
        ls_batch(rw, n, c, lane, r);
        unsigned m[kLsB];
#pragma unroll
        for (int u = 0; u < kLsB; ++u) {
          const int b = max(min(255, __float2int_rd((unorder_bits(v[u]) - fmn) * scale)), 0);
          m[u] = __ballot_sync(kFull, v[u] != 0u && b == b_star);
        }
#pragma unroll
        for (int u = 0; u < kLsB; ++u) {
          if ((m[u] >> lane) & 1u) {
            const int pos = nb0 + __popc(m[u] & lt0);
            hist[2 * pos] = v[u];
            hist[2 * pos + 1] = r[u];
          }
          nb0 += __popc(m[u]);
        }
      }
      __syncwarp();
      ICR_ST_MARK(5);
      ls_rank_list(hist, nb0, need, lane, ks, kr);
      ICR_ST_MARK(6);
      return;
    }
    // crowded bin: its range in score bits (the value -> bin map is monotone, so the bin is an interval of bit patterns)
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int c = 0; c < n; c += 32 * kLsB) {
      uint32_t v[kLsB];
      ls_batch(sc, n, c, lane, v);
#pragma unroll
      for (int u = 0; u < kLsB; ++u) {
        const int b = max(min(255, __float2int_rd((unorder_bits(v[u]) - fmn) * scale)), 0);
        const bool in = v[u] != 0u && b == b_star;
        lo = min(lo, in ? v[u] : 0xFFFFFFFFu);
        hi = max(hi, in ? v[u] : 0u);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(kFull, lo, o));
      hi = max(hi, __shfl_xor_sync(kFull, hi, o));
    }
    blo = lo;
    bhi = hi;
    ICR_ST_MARK(3);
  }
  // ---- crowded bin: levels over its bit range ----
  while (cnt > kLsList && blo != bhi) {
    const int bits = 32 - __clz(static_cast<int>(bhi - blo));
    const int shift = bits > 8 ? bits - 8 : 0;
    const uint32_t lo = blo;
    const int b = ls_level(sc, nullptr, 0u, n, lo, bhi, shift, &need, &cnt, hist);
    blo = lo + (static_cast<uint32_t>(b) << shift);
    bhi = min(bhi, blo + min((1u << shift) - 1u, 0xFFFFFFFFu - blo));
  }
  bool by_row = false;
  uint32_t rlo = 0, rhi = 0xFFFFFFFFu;
  if (cnt > kLsList) {
    // more than kLsList keys share ONE score (duplicated catalog rows): the same levels over ~row among them
    by_row = true;
    uint32_t rmn = 0xFFFFFFFFu, rmx = 0u;
    for (int i = lane; i < n; i += 32)
      if (sc[i] == blo) {
        rmn = min(rmn, rw[i]);
        rmx = max(rmx, rw[i]);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rmn = min(rmn, __shfl_xor_sync(kFull, rmn, o));
      rmx = max(rmx, __shfl_xor_sync(kFull, rmx, o));
    }
    rlo = rmn;
    rhi = rmx;
    while (cnt > kLsList && rlo != rhi) {
      const int bits = 32 - __clz(static_cast<int>(rhi - rlo));
      const int shift2 = bits > 8 ? bits - 8 : 0;
      const uint32_t lo2 = rlo;
      const int b = ls_level(rw, sc, blo, n, lo2, rhi, shift2, &need, &cnt, hist);
      rlo = lo2 + (static_cast<uint32_t>(b) << shift2);
      rhi = min(rhi, rlo + min((1u << shift2) - 1u, 0xFFFFFFFFu - rlo));
    }
  }
  ICR_ST_MARK(4);
  // ---- the boundary keys, as (sc, rw) pairs, into hist[]; then rank counting for the need-th largest of them ----
  int nb = 0;
  for (int c = 0; c < n; c += 32 * kLsB) {
    uint32_t v[kLsB], r[kLsB];
    ls_batch(sc, n, c, lane, v);
    ls_batch(rw, n, c, lane, r);
    unsigned m[kLsB];
#pragma unroll
    for (int u = 0; u < kLsB; ++u) {
      bool in = v[u] >= blo && v[u] <= bhi && v[u] != 0u;
      if (by_row) in = in && r[u] >= rlo && r[u] <= rhi;
      m[u] = __ballot_sync(kFull, in);
    }
#pragma unroll
    for (int u = 0; u < kLsB; ++u) {
      const int pos = nb + __popc(m[u] & lt);
      if (((m[u] >> lane) & 1u) && pos < kLsList) {
        hist[2 * pos] = v[u];
        hist[2 * pos + 1] = r[u];
      }
      nb += __popc(m[u]);
    }
  }
  nb = min(nb, kLsList);  // identical (score, row) pairs (duplicated candidates of a shard merge) can exceed the list: any of them is right
  need = min(need, nb);
  __syncwarp();
  ICR_ST_MARK(5);
  ls_rank_list(hist, nb, need, lane, ks, kr);
  ICR_ST_MARK(6);
}

// In place: keep the keys that can still belong to the result and return how many. band == 0: exactly the k best keys.
// band > 0 (screened scores): every key whose score is within `band` of the k-th best score. *tau_out = the threshold
// score (k-th best, minus band), -inf while there are fewer than k keys (all kept). Not inlined: one copy of the passes
// serves the buffer-full squeezes, the final squeeze and the whole-catalog ranking.
static __device__ __noinline__ int ls_reduce(uint32_t* sc, uint32_t* rw, int n, int k, float band, uint32_t* hist, float* tau_out) {
  const int lane = threadIdx.x & 31;
  *tau_out = -INFINITY;
  if (n < k) return n;
  uint32_t ks, kr;
  ls_kth(sc, rw, n, k, hist, lane, ks, kr);
  const float tau = unorder_bits(ks) - band;
  *tau_out = tau;
  uint32_t ts = ks, tr = kr;
  if (band > 0.f) {
    ts = order_bits(tau);
    tr = 0u;
  }
  const unsigned lt = (1u << lane) - 1u;
  ICR_ST_BEGIN();
  int m = 0;
  for (int c = 0; c < n; c += 32 * kLsB) {
    uint32_t v[kLsB], r[kLsB];
    ls_batch(sc, n, c, lane, v);
    ls_batch(rw, n, c, lane, r);
    unsigned mk[kLsB];
#pragma unroll
    for (int u = 0; u < kLsB; ++u) mk[u] = __ballot_sync(kFull, v[u] != 0u && pair_ge(v[u], r[u], ts, tr));
    // a batch's loads are issued (warp-wide, in order) before its stores, which land at positions <= the batch's own keys
#pragma unroll
    for (int u = 0; u < kLsB; ++u) {
      if ((mk[u] >> lane) & 1u) {
        const int pos = m + __popc(mk[u] & lt);
        sc[pos] = v[u];
        rw[pos] = r[u];
      }
      m += __popc(mk[u]);
    }
  }
  __syncwarp();
  ICR_ST_MARK(7);
  return m;
}

}  // namespace icr
