// IR metric arithmetic over the [Q, K] retrieved-id matrix, on the device.
//
// Replaces the per-query Python loops that follow the top-k in both consumers of the reference:
//   * sentence-transformers InformationRetrievalEvaluator.compute_metrics (built at
//     src/training/train_sbert.py:197-202): accuracy / precision / recall / MRR / NDCG / MAP @k;
//   * src/baselines/metrics.py:13-176 (compute_ir_metrics): same family, but NDCG is normalised by
//     the ideal ordering of the RETRIEVED relevances (:112-119) and AP by min(|relevant|, |ranked|)
//     (:66-72) — the *_RETRIEVED kinds.
//
// One warp per query: the lanes test 32 ranks at a time against the query's sorted relevant rows
// (binary search), ballots turn the hits into bit masks, then lane m walks the masks for metric m
// in rank order with double arithmetic (the order the reference's Python loops add in). A second
// kernel averages every metric over the queries in a fixed order (deterministic).
#include "common.cuh"

namespace icr {

struct MetricArgs {
  const int64_t* ids;
  int64_t ld;
  int64_t Q;
  int K;
  const int64_t* rel_offsets;
  const int64_t* rel_rows;
  const int32_t* n_relevant;
  int M;
  int kind[ICR_MAX_METRICS];
  int k[ICR_MAX_METRICS];
  double* per_query;
  double* means;
};

constexpr int kMaskWords = ICR_MAX_K / 32;

__device__ __forceinline__ bool contains_sorted(const int64_t* a, int64_t lo, int64_t hi, int64_t x) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = a[mid];
    if (v == x) return true;
    if (v < x) lo = mid + 1;
    else hi = mid;
  }
  return false;
}

__global__ void __launch_bounds__(256) ir_metrics_kernel(const MetricArgs g) {
  __shared__ uint32_t s_hit[8][kMaskWords], s_valid[8][kMaskWords];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 8 + wib;
  if (q >= g.Q) return;
  const int64_t* row = g.ids + q * g.ld;
  const int64_t lo = g.rel_offsets[q], hi = g.rel_offsets[q + 1];
#pragma unroll
  for (int c = 0; c < kMaskWords; ++c) {
    const int r = c * 32 + lane;
    const int64_t id = r < g.K ? row[r] : -1;
    const bool v = id >= 0;
    const bool h = v && contains_sorted(g.rel_rows, lo, hi, id);
    const uint32_t hm = __ballot_sync(kFull, h), vm = __ballot_sync(kFull, v);
    if (lane == 0) {
      s_hit[wib][c] = hm;
      s_valid[wib][c] = vm;
    }
  }
  __syncwarp();
  const double nrel = static_cast<double>(g.n_relevant[q]);
  for (int m = lane; m < g.M; m += 32) {
    const int kind = g.kind[m], k = g.k[m];
    const int kk = k < g.K ? k : g.K;
    // walk the HITS only (set bits in rank order: the same summation order as the reference's loops over ranks);
    // counts come from population counts of the masks cut at rank kk
    int cum = 0, len = 0, first = -1;
    double dcg = 0.0, sum_prec = 0.0;
#pragma unroll
    for (int c = 0; c < kMaskWords; ++c) {
      const int lo_r = c * 32;
      if (lo_r >= kk) break;
      const uint32_t cut = (kk - lo_r >= 32) ? 0xffffffffu : ((1u << (kk - lo_r)) - 1u);
      len += __popc(s_valid[wib][c] & cut);
      uint32_t h = s_hit[wib][c] & cut;
      while (h) {
        const int r = lo_r + __ffs(h) - 1;
        h &= h - 1;
        ++cum;
        if (first < 0) first = r;
        dcg += 1.0 / log2(static_cast<double>(r + 2));
        sum_prec += static_cast<double>(cum) / static_cast<double>(r + 1);
      }
    }
    double v = 0.0;
    if (nrel > 0.0) {
      switch (kind) {
        case ICR_METRIC_ACCURACY: v = cum > 0 ? 1.0 : 0.0; break;
        case ICR_METRIC_PRECISION: v = static_cast<double>(cum) / static_cast<double>(k); break;
        case ICR_METRIC_RECALL: v = static_cast<double>(cum) / nrel; break;
        case ICR_METRIC_MRR: v = first >= 0 ? 1.0 / static_cast<double>(first + 1) : 0.0; break;
        case ICR_METRIC_NDCG:
        case ICR_METRIC_NDCG_RETRIEVED: {
          const int ideal = kind == ICR_METRIC_NDCG ? static_cast<int>(nrel < k ? nrel : k) : cum;
          double idcg = 0.0;
          for (int i = 0; i < ideal; ++i) idcg += 1.0 / log2(static_cast<double>(i + 2));
          v = idcg > 0.0 ? dcg / idcg : 0.0;
          break;
        }
        case ICR_METRIC_MAP: v = sum_prec / (nrel < k ? nrel : static_cast<double>(k)); break;
        case ICR_METRIC_MAP_RETRIEVED: v = len > 0 ? sum_prec / (nrel < len ? nrel : static_cast<double>(len)) : 0.0; break;
        default: break;
      }
    }
    g.per_query[q * g.M + m] = v;
  }
}

// mean over queries of metric blockIdx.x: every thread adds its strided share in query order, then a fixed tree
__global__ void __launch_bounds__(256) ir_metrics_mean_kernel(const double* per_query, int64_t Q, int M, double* means) {
  __shared__ double s[256];
  const int m = blockIdx.x;
  double acc = 0.0;
  for (int64_t q = threadIdx.x; q < Q; q += 256) acc += per_query[q * M + m];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) means[m] = Q > 0 ? s[0] / static_cast<double>(Q) : 0.0;
}

int launch_ir_metrics(const int64_t* ids, int64_t Q, int K, int64_t ld, const int64_t* rel_offsets, const int64_t* rel_rows,
                      const int32_t* n_relevant, const int32_t* kinds, const int32_t* ks, int M, double* per_query, double* means,
                      cudaStream_t st) {
  MetricArgs g{};
  g.ids = ids;
  g.ld = ld;
  g.Q = Q;
  g.K = K;
  g.rel_offsets = rel_offsets;
  g.rel_rows = rel_rows;
  g.n_relevant = n_relevant;
  g.M = M;
  for (int m = 0; m < M; ++m) {
    g.kind[m] = kinds[m];
    g.k[m] = ks[m];
  }
  g.per_query = per_query;
  g.means = means;
  if (Q > 0) {
    ir_metrics_kernel<<<static_cast<unsigned>((Q + 7) / 8), 256, 0, st>>>(g);
    ICR_LAUNCH_CHECK();
  }
  ir_metrics_mean_kernel<<<M, 256, 0, st>>>(per_query, Q, M, means);
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

}  // namespace icr
