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

constexpr int kHsWarps = 4;       // queries per CTA (11 KB of shared memory each: 5 CTAs = 20 warps per SM)
constexpr int kHsCap = 1024;      // keys buffered per query at once
constexpr int kHsSel = 256;       // >= ICR_MAX_K

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
};

__device__ __forceinline__ uint64_t canonical_from_raw(uint64_t raw) {
  const float f = __uint_as_float(static_cast<uint32_t>(raw >> 32));
  return (static_cast<uint64_t>(order_bits(f)) << 32) | (raw & 0xFFFFFFFFull);
}

struct HsSmem {
  uint64_t buf[kHsWarps][kHsCap];
  uint64_t sel[kHsWarps][kHsSel];
  unsigned int hist[kHsWarps][kHsBins];
  int pref[kHsWarps][33];
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
  int* pref = sm.pref[warp];
  const int k = a.k;
  const int cap = min(a.seg_cap, a.seg_stride);

  int n = 0;
  if (a.carry_in) {
    n = min(a.carry_cnt_in[q], k);
    for (int i = lane; i < n; i += 32) buf[i] = a.carry_in[q * k + i];
  }
  const uint64_t* qbase = a.seg_keys + q * a.nseg * static_cast<int64_t>(a.seg_stride);
  for (int g0 = 0; g0 < a.nseg; g0 += 32) {
    const int s_mine = g0 + lane;
    const int my_cnt = (s_mine < a.nseg) ? min(a.seg_cnt[q * a.nseg + s_mine], cap) : 0;
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(kFull, incl, 31);
    __syncwarp();
    pref[lane] = incl - my_cnt;
    if (lane == 31) pref[32] = total;
    __syncwarp();
    int done = 0;
    while (done < total) {
      if (kHsCap - n < 32) {  // buffer full: keep the k best and go on
        warp_select_topk(buf, n, k, sel, hist, lane);
        for (int i = lane; i < k; i += 32) buf[i] = sel[i];
        n = k;
        __syncwarp();
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
            // segment holding flat index idx: largest s with pref[s] <= idx (binary search over 32 entries)
            int s = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
              if (pref[s + step] <= idx) s += step;
            v[u] = __ldcs(qbase + static_cast<int64_t>(g0 + s) * a.seg_stride + (idx - pref[s]));
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
    }
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

int launch_select_hist(const uint64_t* seg_keys, const int* seg_cnt, int64_t Q, int nseg, int seg_stride, int seg_cap,
                       const uint64_t* carry_in, const int* carry_cnt_in, uint64_t* carry_out, int* carry_cnt_out, float* tau_out,
                       float* out_scores, int64_t* out_ids, int64_t id_offset, int k, float out_scale, const float* out_qscale,
                       cudaStream_t st) {
  if (Q == 0) return ICR_OK;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    ICR_CUDA_CHECK(cudaFuncSetAttribute(select_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(HsSmem))));
    attr_set = true;
  }
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
  const unsigned grid = static_cast<unsigned>((Q + kHsWarps - 1) / kHsWarps);
  select_hist_kernel<<<grid, kHsWarps * 32, sizeof(HsSmem), st>>>(a);
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

}  // namespace icr
