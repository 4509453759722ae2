// Warp-per-query exact top-k selection by histogram partitioning (no full sort).
//
// The GEMM epilogue leaves, per query and phase, a few hundred candidate keys spread over a handful of
// segments, plus the k keys carried from earlier phases. Only the k largest matter, so instead of sorting
// everything (select.cu) one warp per query
//   1. gathers the keys into shared memory (all loads independent, coalesced 8-byte lanes),
//   2. bins them linearly over [min key, max key] into 256 bins (shared-memory histogram),
//   3. scans the histogram from the top to find the bin holding the k-th largest key,
//   4. moves keys above that bin straight to the output, compacts the keys of that bin in place, and
//      either sorts those few (<= 64) or repeats 2-4 on them with 256x finer bins.
// Keys are distinct 64-bit values (score bits, ~row), so ties in score are resolved exactly by row and the
// refinement always terminates. The final phase sorts the k selected keys; earlier phases only need the set
// and its minimum (the new threshold tau).
#include "common.cuh"
#include "select_warp.cuh"

namespace icr {

#ifndef ICR_HS_WARPS
#define ICR_HS_WARPS 4
#endif
constexpr int kHsWarps = ICR_HS_WARPS;  // queries per CTA, 11 KB of shared memory each (tuning: profiles/r01_notes.md)
constexpr int kHsCap = 1024;      // keys buffered per query at once
constexpr int kHsSel = 256;       // >= ICR_MAX_K
constexpr int kHsPref = 448;      // segments per gather round: the prefix array (+1) is overlaid on sel[] (512 ints)

struct HistSelectArgs {
  const uint64_t* seg_keys;  // [Q][nseg][seg_stride]
  const int* seg_cnt;        // [Q][nseg]
  int nseg;
  int seg_stride;
  int seg_cap;
  const uint64_t* carry_in;  // [Q][k] or null
  const int* carry_cnt_in;   // [Q]
  uint64_t* carry_out;       // [Q][k] or null (unsorted set)
  int* carry_cnt_out;        // [Q]
  float* tau_out;            // [Q] or null
  float* out_scores;         // [Q][k] or null (sorted descending)
  int64_t* out_ids;          // [Q][k]
  int64_t id_offset;
  int k;
  int64_t Q;
  float out_scale;          // final score = key score * out_scale * (out_qscale ? out_qscale[q] : 1)
  const float* out_qscale;  // [Q] or null
  int raw_keys;             // segments hold (score bits, ~row) as the GEMM epilogue writes them; carry keys are ordered
  // dense front end (first phase of K2): instead of segments, row q of a [Q][dense_ld] score matrix holds the raw
  // scores of catalog rows 0 .. dense_rows-1; rows flagged in `mask` are skipped
  const float* dense;
  int64_t dense_ld;
  int dense_rows;
  const uint8_t* mask;
  // list front end (K4 shard merge): G lists of list_k (score, id) pairs per query, laid out [G][Q][list_k]; id < 0 = empty
  const float* list_scores;
  const int64_t* list_ids;
  int list_g, list_k;
};

// flat candidate t of query q from the dense matrix or the shard lists -> key; false if there is none
__device__ __forceinline__ bool front_fetch(const HistSelectArgs& a, int64_t q, int t, int count, uint64_t& key) {
  key = 0ull;
  if (t >= count) return false;
  if (a.dense) {
    if (a.mask && a.mask[t]) return false;
    key = make_key(__ldcs(a.dense + q * a.dense_ld + t), static_cast<uint32_t>(t));
    return true;
  }
  const int g = t / a.list_k, i = t - g * a.list_k;
  const int64_t off = (static_cast<int64_t>(g) * a.Q + q) * a.list_k + i;
  const int64_t id = __ldcs(a.list_ids + off);
  if (id < 0) return false;
  key = make_key(__ldcs(a.list_scores + off), static_cast<uint32_t>(id));
  return true;
}
__device__ __forceinline__ int front_count(const HistSelectArgs& a) {
  return a.dense ? a.dense_rows : (a.list_scores ? a.list_g * a.list_k : 0);
}

__device__ __forceinline__ uint64_t canonical_from_raw(uint64_t raw) {
  const float f = __uint_as_float(static_cast<uint32_t>(raw >> 32));
  return (static_cast<uint64_t>(order_bits(f)) << 32) | (raw & 0xFFFFFFFFull);
}

struct HsSmem {
  uint64_t buf[kHsWarps][kHsCap];
  uint64_t sel[kHsWarps][kHsSel];
  unsigned int hist[kHsWarps][kHsBins];
};

__global__ void __launch_bounds__(kHsWarps * 32) select_hist_kernel(HistSelectArgs a) {
  extern __shared__ __align__(16) unsigned char hs_raw[];
  HsSmem& sm = *reinterpret_cast<HsSmem*>(hs_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * kHsWarps + warp;
  if (q >= a.Q) return;  // whole warps leave; nothing below synchronises across warps
  uint64_t* buf = sm.buf[warp];
  uint64_t* sel = sm.sel[warp];
  unsigned int* hist = sm.hist[warp];
  const int k = a.k;
  const int cap = min(a.seg_cap, a.seg_stride);

  int n = 0;
  if (a.carry_in) {
    n = min(a.carry_cnt_in[q], k);
    for (int i = lane; i < n; i += 32) buf[i] = a.carry_in[q * k + i];
  }
  if (const int count = front_count(a)) {
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < count; base += 32 * 8) {
      if (kHsCap - n < 32 * 8) {  // buffer full: keep the k best and go on
        warp_select_topk(buf, n, k, sel, hist, lane);
        for (int i = lane; i < k; i += 32) buf[i] = sel[i];
        n = k;
        __syncwarp();
      }
      uint64_t key[8];
      bool ok[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) ok[u] = front_fetch(a, q, base + u * 32 + lane, count, key[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const unsigned m = __ballot_sync(kFull, ok[u]);
        if (ok[u]) buf[n + __popc(m & lt)] = key[u];
        n += __popc(m);
      }
    }
  }
  const uint64_t* qbase = a.seg_keys + q * a.nseg * static_cast<int64_t>(a.seg_stride);
  // Segment lengths of up to kHsPref segments are fetched at once (independent loads) and prefix-summed in shared
  // memory, then the keys are gathered by flat index (binary search for the segment): a small batch may have
  // hundreds of short segments per query (one per chunk and CTA), and walking them 32 at a time made the gather
  // a chain of dependent memory round trips. The prefix array lives in `sel`, which is idle until the selection.
  int* pref = reinterpret_cast<int*>(sel);
  for (int g0 = 0; g0 < a.nseg; g0 += kHsPref) {
    const int ng = min(kHsPref, a.nseg - g0);
    int top = 1;
    while (top * 2 <= ng) top *= 2;  // first step of the binary search
    int done = 0, total = 0;
    bool have_pref = false;
    do {
      if (!have_pref) {
        __syncwarp();
#pragma unroll 4
        for (int i = lane; i < ng; i += 32) pref[i] = min(__ldg(a.seg_cnt + q * a.nseg + g0 + i), cap);
        __syncwarp();
        int carry = 0;
        for (int b0 = 0; b0 < ng; b0 += 32) {
          const int v = (b0 + lane < ng) ? pref[b0 + lane] : 0;
          int incl = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += u;
          }
          if (b0 + lane < ng) pref[b0 + lane] = carry + incl - v;
          carry += __shfl_sync(kFull, incl, 31);
        }
        if (lane == 0) pref[ng] = carry;
        total = carry;
        have_pref = true;
        __syncwarp();
      }
      if (done >= total) break;
      if (kHsCap - n < 32) {  // buffer full: keep the k best and go on (the selection overwrites the prefix array)
        warp_select_topk(buf, n, k, sel, hist, lane);
        for (int i = lane; i < k; i += 32) buf[i] = sel[i];
        n = k;
        have_pref = false;
        continue;
      }
      const int take = min(kHsCap - n, total - done);
      // 8 independent global loads per lane in flight before the first shared-memory store: the gather is
      // latency-bound (cold candidates), so memory-level parallelism is what sets its speed
      for (int i0 = 0; i0 < take; i0 += 32 * 8) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 32 + lane;
          v[u] = 0ull;
          if (i < take) {
            const int idx = done + i;
            // segment holding flat index idx: largest s with pref[s] <= idx
            int sg = 0;
            for (int step = top; step > 0; step >>= 1)
              if (sg + step < ng && pref[sg + step] <= idx) sg += step;
            v[u] = __ldcs(qbase + static_cast<int64_t>(g0 + sg) * a.seg_stride + (idx - pref[sg]));
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 32 + lane;
          if (i < take) buf[n + i] = a.raw_keys ? canonical_from_raw(v[u]) : v[u];
        }
      }
      n += take;
      done += take;
      __syncwarp();
    } while (done < total);
  }
  __syncwarp();
  int kept = n;
  if (n > k) {
    warp_select_topk(buf, n, k, sel, hist, lane);
    kept = k;
  } else {
    for (int i = lane; i < n; i += 32) sel[i] = buf[i];
    __syncwarp();
  }

  if (a.carry_out) {
    for (int i = lane; i < kept; i += 32) a.carry_out[q * k + i] = sel[i];
    if (lane == 0) a.carry_cnt_out[q] = kept;
  }
  if (a.tau_out) {
    uint64_t mn = ~0ull;
    for (int i = lane; i < kept; i += 32) mn = sel[i] < mn ? sel[i] : mn;
    mn = warp_min_u64(mn);
    if (lane == 0) a.tau_out[q] = (kept >= k) ? key_score(mn) : -INFINITY;
  }
  if (a.out_scores) {
    warp_rank_sort_desc(sel, kept, buf, kept, lane);  // buf is free again: sorted output order
    const float scale = a.out_scale * (a.out_qscale ? a.out_qscale[q] : 1.0f);
    for (int i = lane; i < k; i += 32) {
      const bool ok = i < kept;
      a.out_scores[q * k + i] = ok ? key_score(buf[i]) * scale : -INFINITY;
      a.out_ids[q * k + i] = ok ? static_cast<int64_t>(key_row(buf[i])) + a.id_offset : -1;
    }
  }
}

// ---- block-per-query variant for small batches --------------------------------------------------------------
// With a few dozen queries the warp-per-query kernel leaves the GPU idle while each warp walks thousands of keys on
// its own (the swapped GEMM kernel produces one short segment per chunk and CTA: hundreds per query). Here a CTA of
// 256 threads owns a query: block-wide gather, block-wide histogram partition down to <= 1024 keys, then one warp
// finishes with warp_select_topk. Same inputs, outputs and exactness as select_hist_kernel.
constexpr int kBsThreads = 256;
constexpr int kBsCap = 6144;    // keys buffered per query at once (two buffers: 96 KB)
constexpr int kBsPref = 1024;   // segments per gather round

struct BsSmem {
  uint64_t buf[kBsCap];
  uint64_t bnd[kBsCap];
  uint64_t sel[kHsSel];
  uint64_t red[2 * (kBsThreads / 32)];
  unsigned int hist[kHsBins];
  int pref[kBsPref + 1];
  int misc[8];
};

// k largest of cur[0..n) -> sm.sel[0..k) (unordered); n > k. `cur` is sm.buf; sm.bnd is scratch.
__device__ __forceinline__ void block_select_topk(BsSmem& sm, int n, int k, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  uint64_t* cur = sm.buf;
  uint64_t* other = sm.bnd;
  int need = k, nsel = 0, len = n;
  for (;;) {
    if (len <= 1024) {
      if (warp == 0) {
        if (need == len) {
          for (int i = lane; i < len; i += 32) sm.sel[nsel + i] = cur[i];
        } else {
          warp_select_topk(cur, len, need, sm.sel + nsel, sm.hist, lane);
        }
      }
      __syncthreads();
      return;
    }
    uint64_t mn = ~0ull, mx = 0ull;
    for (int i = tid; i < len; i += kBsThreads) {
      const uint64_t key = cur[i];
      mn = key < mn ? key : mn;
      mx = key > mx ? key : mx;
    }
    mn = warp_min_u64(mn);
    mx = warp_max_u64(mx);
    if (lane == 0) {
      sm.red[warp] = mn;
      sm.red[kBsThreads / 32 + warp] = mx;
    }
    if (tid < kHsBins) sm.hist[tid] = 0u;
    __syncthreads();
    mn = sm.red[0];
    mx = sm.red[kBsThreads / 32];
#pragma unroll
    for (int w = 1; w < kBsThreads / 32; ++w) {
      mn = sm.red[w] < mn ? sm.red[w] : mn;
      mx = sm.red[kBsThreads / 32 + w] > mx ? sm.red[kBsThreads / 32 + w] : mx;
    }
    const uint64_t range = mx - mn;
    const int bits = 64 - __clzll(static_cast<long long>(range | 1ull));
    const int shift = bits > 8 ? bits - 8 : 0;
    for (int i = tid; i < len; i += kBsThreads) atomicAdd(&sm.hist[255 - static_cast<int>((cur[i] - mn) >> shift)], 1u);
    __syncthreads();
    if (warp == 0) {  // cumulative counts from the top bin: lane L owns t in [8L, 8L+8)
      unsigned int local[8], lsum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        local[j] = sm.hist[lane * 8 + j];
        lsum += local[j];
      }
      unsigned int incl = lsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      unsigned int c = incl - lsum;
      if (c < static_cast<unsigned int>(need) && static_cast<unsigned int>(need) <= incl) {
        int t_star = -1;
        unsigned int above = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (t_star < 0 && c + local[j] >= static_cast<unsigned int>(need)) {
            t_star = lane * 8 + j;
            above = c;
          }
          c += local[j];
        }
        sm.misc[0] = 255 - t_star;
        sm.misc[1] = static_cast<int>(above);
      }
      if (lane == 0) {
        sm.misc[2] = nsel;
        sm.misc[3] = 0;
      }
    }
    __syncthreads();
    const int b_star = sm.misc[0];
    for (int i = tid; i < len; i += kBsThreads) {
      const uint64_t key = cur[i];
      const int bin = static_cast<int>((key - mn) >> shift);
      if (bin > b_star) sm.sel[atomicAdd(&sm.misc[2], 1)] = key;
      else if (bin == b_star) other[atomicAdd(&sm.misc[3], 1)] = key;
    }
    __syncthreads();
    nsel = sm.misc[2];
    need -= sm.misc[1];
    const int nb = sm.misc[3];
    __syncthreads();  // misc is rewritten in the next round
    if (need == nb) {
      for (int i = tid; i < nb; i += kBsThreads) sm.sel[nsel + i] = other[i];
      __syncthreads();
      return;
    }
    uint64_t* t = cur;
    cur = other;
    other = t;
    len = nb;
    if (cur != sm.buf && len <= 1024) {  // the warp-level finish works in place on `cur`; keep it simple: move back
      for (int i = tid; i < len; i += kBsThreads) sm.buf[i] = cur[i];
      __syncthreads();
      cur = sm.buf;
      other = sm.bnd;
    }
  }
}

__global__ void __launch_bounds__(kBsThreads) select_block_kernel(HistSelectArgs a) {
  extern __shared__ __align__(16) unsigned char bs_raw[];
  BsSmem& sm = *reinterpret_cast<BsSmem*>(bs_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t q = blockIdx.x;
  const int k = a.k;
  const int cap = min(a.seg_cap, a.seg_stride);
  int n = 0;
  if (a.carry_in) {
    n = min(a.carry_cnt_in[q], k);
    for (int i = tid; i < n; i += kBsThreads) sm.buf[i] = a.carry_in[q * k + i];
  }
  if (const int count = front_count(a)) {
    const unsigned lt = (1u << lane) - 1u;
    __syncthreads();
    for (int base = 0; base < count; base += kBsThreads * 4) {
      if (kBsCap - n < kBsThreads * 4) {
        block_select_topk(sm, n, k, tid);
        for (int i = tid; i < k; i += kBsThreads) sm.buf[i] = sm.sel[i];
        n = k;
        __syncthreads();
      }
      if (tid == 0) sm.misc[4] = n;
      __syncthreads();
      uint64_t key[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) ok[u] = front_fetch(a, q, base + u * kBsThreads + tid, count, key[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned m = __ballot_sync(kFull, ok[u]);
        int wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(&sm.misc[4], __popc(m));
        wbase = __shfl_sync(kFull, wbase, 0);
        if (ok[u]) sm.buf[wbase + __popc(m & lt)] = key[u];
      }
      __syncthreads();
      n = sm.misc[4];
      __syncthreads();
    }
  }
  const uint64_t* qbase = a.seg_keys + q * a.nseg * static_cast<int64_t>(a.seg_stride);
  for (int g0 = 0; g0 < a.nseg; g0 += kBsPref) {
    const int ng = min(kBsPref, a.nseg - g0);
    int top = 1;
    while (top * 2 <= ng) top *= 2;
    __syncthreads();
    for (int i = tid; i < ng; i += kBsThreads) sm.pref[i] = min(__ldg(a.seg_cnt + q * a.nseg + g0 + i), cap);
    __syncthreads();
    if (warp == 0) {
      int carry = 0;
      for (int b0 = 0; b0 < ng; b0 += 32) {
        const int v = (b0 + lane < ng) ? sm.pref[b0 + lane] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += u;
        }
        if (b0 + lane < ng) sm.pref[b0 + lane] = carry + incl - v;
        carry += __shfl_sync(kFull, incl, 31);
      }
      if (lane == 0) sm.pref[ng] = carry;
    }
    __syncthreads();
    const int total = sm.pref[ng];
    int done = 0;
    while (done < total) {
      if (kBsCap - n < kBsThreads) {  // buffer full: keep the k best and go on
        block_select_topk(sm, n, k, tid);
        for (int i = tid; i < k; i += kBsThreads) sm.buf[i] = sm.sel[i];
        n = k;
        __syncthreads();
      }
      const int take = min(kBsCap - n, total - done);
      for (int i0 = 0; i0 < take; i0 += kBsThreads * 4) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kBsThreads + tid;
          v[u] = 0ull;
          if (i < take) {
            const int idx = done + i;
            int sg = 0;
            for (int step = top; step > 0; step >>= 1)
              if (sg + step < ng && sm.pref[sg + step] <= idx) sg += step;
            v[u] = __ldcs(qbase + static_cast<int64_t>(g0 + sg) * a.seg_stride + (idx - sm.pref[sg]));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kBsThreads + tid;
          if (i < take) sm.buf[n + i] = a.raw_keys ? canonical_from_raw(v[u]) : v[u];
        }
      }
      n += take;
      done += take;
      __syncthreads();
    }
  }
  __syncthreads();
  int kept = n;
  if (n > k) {
    block_select_topk(sm, n, k, tid);
    kept = k;
  } else {
    for (int i = tid; i < n; i += kBsThreads) sm.sel[i] = sm.buf[i];
    __syncthreads();
  }
  if (a.carry_out) {
    for (int i = tid; i < kept; i += kBsThreads) a.carry_out[q * k + i] = sm.sel[i];
    if (tid == 0) a.carry_cnt_out[q] = kept;
  }
  // every thread ranks one selected key (kept <= 256): the minimum is the new threshold, the ranks the output order
  const bool mine = tid < kept;
  const uint64_t key = mine ? sm.sel[tid] : 0ull;
  int rank = 0;
  if (mine)
    for (int j = 0; j < kept; ++j) rank += (sm.sel[j] > key) ? 1 : 0;
  if (a.tau_out && mine && rank == kept - 1) a.tau_out[q] = (kept >= k) ? key_score(key) : -INFINITY;
  if (a.tau_out && kept == 0 && tid == 0) a.tau_out[q] = -INFINITY;
  if (a.out_scores) {
    const float scale = a.out_scale * (a.out_qscale ? a.out_qscale[q] : 1.0f);
    if (mine) {
      a.out_scores[q * k + rank] = key_score(key) * scale;
      a.out_ids[q * k + rank] = static_cast<int64_t>(key_row(key)) + a.id_offset;
    }
    for (int i = kept + tid; i < k; i += kBsThreads) {
      a.out_scores[q * k + i] = -INFINITY;
      a.out_ids[q * k + i] = -1;
    }
  }
}

static int run_select(const HistSelectArgs& a, int64_t Q, cudaStream_t st);

// K4: merge of G shard lists of k_in (score, id) pairs per query, [G][Q][k_in] -> sorted top k_out per query
int launch_merge_lists(const float* cs, const int64_t* ci, int64_t Q, int G, int k_in, int k_out, float* os, int64_t* oi,
                       cudaStream_t st) {
  if (Q == 0) return ICR_OK;
  HistSelectArgs a{};
  a.k = k_out;
  a.Q = Q;
  a.out_scores = os;
  a.out_ids = oi;
  a.out_scale = 1.0f;
  a.list_scores = cs;
  a.list_ids = ci;
  a.list_g = G;
  a.list_k = k_in;
  return run_select(a, Q, st);
}

int launch_select_hist(const uint64_t* seg_keys, const int* seg_cnt, int64_t Q, int nseg, int seg_stride, int seg_cap,
                       const uint64_t* carry_in, const int* carry_cnt_in, uint64_t* carry_out, int* carry_cnt_out, float* tau_out,
                       float* out_scores, int64_t* out_ids, int64_t id_offset, int k, float out_scale, const float* out_qscale,
                       cudaStream_t st, const float* dense, int64_t dense_ld, int dense_rows, const uint8_t* mask) {
  if (Q == 0) return ICR_OK;
  HistSelectArgs a{};
  a.seg_keys = seg_keys;
  a.seg_cnt = seg_cnt;
  a.nseg = nseg;
  a.seg_stride = seg_stride;
  a.seg_cap = seg_cap;
  a.carry_in = carry_in;
  a.carry_cnt_in = carry_cnt_in;
  a.carry_out = carry_out;
  a.carry_cnt_out = carry_cnt_out;
  a.tau_out = tau_out;
  a.out_scores = out_scores;
  a.out_ids = out_ids;
  a.id_offset = id_offset;
  a.k = k;
  a.Q = Q;
  a.out_scale = out_scale;
  a.out_qscale = out_qscale;
  a.raw_keys = 1;  // the only producer of segments is the GEMM epilogue
  a.dense = dense;
  a.dense_ld = dense_ld;
  a.dense_rows = dense_rows;
  a.mask = mask;
  return run_select(a, Q, st);
}

static int run_select(const HistSelectArgs& a, int64_t Q, cudaStream_t st) {
  static thread_local bool attr_set = false;
  if (!attr_set) {
    ICR_CUDA_CHECK(cudaFuncSetAttribute(select_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(HsSmem))));
    attr_set = true;
  }
  static const int block_max_q = getenv("ICR_SELECT_BLOCK_MAXQ") ? atoi(getenv("ICR_SELECT_BLOCK_MAXQ")) : 512;  // tuning hook
  if (Q <= block_max_q) {
    static thread_local bool battr_set = false;
    if (!battr_set) {
      ICR_CUDA_CHECK(cudaFuncSetAttribute(select_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(BsSmem))));
      battr_set = true;
    }
    select_block_kernel<<<static_cast<unsigned>(Q), kBsThreads, sizeof(BsSmem), st>>>(a);
    ICR_LAUNCH_CHECK();
    return ICR_OK;
  }
  const unsigned grid = static_cast<unsigned>((Q + kHsWarps - 1) / kHsWarps);
  select_hist_kernel<<<grid, kHsWarps * 32, sizeof(HsSmem), st>>>(a);
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

}  // namespace icr
