// Exact top-k selection over candidate segments (shared by K1, K2 and the K4 shard merge).
//
// A "segment" is a slot of 64-bit candidate keys written by one producer (a GEMV CTA, an
// epilogue thread of the GEMM kernel, a shard). One CTA selects the k best keys of a group of
// segments of one query by bitonic sort in shared memory; because keys order by (score desc,
// row asc) the result is deterministic whatever order producers appended in.
#include "common.cuh"

namespace icr {

constexpr int kSortCap = 4096;     // keys sorted at once (32 KB of shared memory)
constexpr int kSelectThreads = 512;

struct SelectArgs {
  const uint64_t* seg_keys;  // [Q][nseg][seg_stride]
  const int* seg_cnt;        // [Q][nseg] valid keys per segment (clamped to seg_cap); may be null => seg_cap
  int nseg;
  int seg_stride;
  int seg_cap;
  int group_size;            // segments per CTA; groups = ceil(nseg / group_size) = gridDim.y
  const uint64_t* carry_in;  // [Q][k] best-so-far keys from earlier phases (may be null)
  const int* carry_cnt_in;   // [Q]
  uint64_t* keys_out;        // [Q][groups][k] (may be null)
  int* cnt_out;              // [Q][groups]
  float* tau_out;            // [Q] score of the k-th best, -inf if fewer than k (groups == 1 only; may be null)
  float* out_scores;         // [Q][k] final (groups == 1 only; may be null)
  int64_t* out_ids;          // [Q][k]
  int64_t id_offset;
  int k;
};

__device__ __forceinline__ int sort_and_truncate(uint64_t* keys, int n, int k) {
  // pads [n, P) with empty keys, sorts descending, returns min(#non-empty, k)
  const int P = next_pow2(n < 2 ? 2 : n);
  for (int i = n + threadIdx.x; i < P; i += blockDim.x) keys[i] = 0ull;
  block_bitonic_sort_desc(keys, P);
  // count of non-empty keys among the first k (keys sorted: non-empty first)
  __shared__ int s_valid;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  const int lim = n < k ? n : k;
  int mine = 0;
  for (int i = threadIdx.x; i < lim; i += blockDim.x) mine += (keys[i] != 0ull);
  mine = static_cast<int>(warp_sum(static_cast<float>(mine)) + 0.5f);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_valid, mine);
  __syncthreads();
  const int v = s_valid;
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(kSelectThreads) select_topk_kernel(SelectArgs a) {
  __shared__ uint64_t keys[kSortCap];
  const int64_t q = blockIdx.x;
  const int g = blockIdx.y;
  const int groups = gridDim.y;
  const int k = a.k;
  const int s0 = g * a.group_size;
  const int s1 = min(a.nseg, s0 + a.group_size);
  const int nsg = s1 - s0;
  const int cap = min(a.seg_cap, a.seg_stride);
  const uint64_t* base = a.seg_keys + (q * a.nseg + s0) * static_cast<int64_t>(a.seg_stride);
  const int* cnts = a.seg_cnt ? a.seg_cnt + q * a.nseg + s0 : nullptr;

  int n = 0;
  if (a.carry_in && g == 0) {
    const int c = min(a.carry_cnt_in[q], k);
    for (int i = threadIdx.x; i < c; i += blockDim.x) keys[i] = a.carry_in[q * k + i];
    n = c;
  }
  if (nsg > 16 && static_cast<int64_t>(nsg) * cap + n <= kSortCap) {
    // flat gather: every slot of every segment in parallel, empty slots become key 0
    const int total = nsg * cap;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
      const int s = t / cap, i = t - s * cap;
      const int c = cnts ? min(cnts[s], cap) : cap;
      keys[n + t] = (i < c) ? base[static_cast<int64_t>(s) * a.seg_stride + i] : 0ull;
    }
    n += total;
  } else {
    // packed: append the valid keys of segment after segment (few, long segments: the GEMM epilogue's),
    // compacting to the k best whenever the buffer fills
    for (int s = 0; s < nsg; ++s) {
      const int c = cnts ? min(cnts[s], cap) : cap;
      int pos = 0;
      while (pos < c) {
        if (n == kSortCap) {
          __syncthreads();
          n = sort_and_truncate(keys, n, k);
        }
        const int take = min(kSortCap - n, c - pos);
        for (int i = threadIdx.x; i < take; i += blockDim.x) keys[n + i] = base[static_cast<int64_t>(s) * a.seg_stride + pos + i];
        n += take;
        pos += take;
      }
    }
  }
  __syncthreads();
  const int valid = sort_and_truncate(keys, n, k);

  if (a.keys_out) {
    uint64_t* o = a.keys_out + (q * groups + g) * static_cast<int64_t>(k);
    for (int i = threadIdx.x; i < valid; i += blockDim.x) o[i] = keys[i];
    if (threadIdx.x == 0) a.cnt_out[q * groups + g] = valid;
  }
  if (a.tau_out && threadIdx.x == 0) a.tau_out[q] = (valid >= k) ? key_score(keys[k - 1]) : -INFINITY;
  if (a.out_scores) {
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const bool ok = i < valid;
      a.out_scores[q * k + i] = ok ? key_score(keys[i]) : -INFINITY;
      a.out_ids[q * k + i] = ok ? static_cast<int64_t>(key_row(keys[i])) + a.id_offset : -1;
    }
  }
}

// Host driver: reduce [Q][nseg] segments (each <= seg_cap keys) to the final [Q][k] output,
// using as many levels as needed. `scratch` must hold 2 * Q * ceil(nseg/gs) * k keys + counts
// (see select_scratch_bytes).
size_t select_scratch_bytes(int64_t Q, int nseg, int seg_cap, int k) {
  int cap = seg_cap;
  int gs = kSortCap / (cap > 0 ? cap : 1);
  if (gs < 2) gs = 2;
  const int64_t groups = (nseg + gs - 1) / gs;
  if (groups <= 1) return 256;
  // level-1 output; deeper levels are strictly smaller and ping-pong between two halves
  const size_t keys = align_up(static_cast<size_t>(Q) * groups * k * sizeof(uint64_t), 256);
  const size_t cnts = align_up(static_cast<size_t>(Q) * groups * sizeof(int), 256);
  return 2 * (keys + cnts) + 256;
}

int launch_select(const uint64_t* seg_keys, const int* seg_cnt, int64_t Q, int nseg, int seg_stride, int seg_cap,
                  const uint64_t* carry_in, const int* carry_cnt_in, uint64_t* carry_out, int* carry_cnt_out,
                  float* tau_out, float* out_scores, int64_t* out_ids, int64_t id_offset, int k, void* scratch,
                  size_t scratch_bytes, cudaStream_t st) {
  if (Q == 0) return ICR_OK;
  const uint64_t* cur_keys = seg_keys;
  const int* cur_cnt = seg_cnt;
  int cur_nseg = nseg, cur_stride = seg_stride, cur_cap = seg_cap;
  const uint64_t* cin = carry_in;
  const int* ccnt = carry_cnt_in;
  int level = 0;
  for (;;) {
    int gs = kSortCap / (cur_cap > 0 ? cur_cap : 1);
    if (gs < 2) gs = 2;
    // the streaming path inside the kernel copes with any group size, but costs serial time;
    // keep groups flat-gatherable unless there is a single query block of work anyway
    int groups = (cur_nseg + gs - 1) / gs;
    if (groups < 1) groups = 1;
    SelectArgs a{};
    a.seg_keys = cur_keys;
    a.seg_cnt = cur_cnt;
    a.nseg = cur_nseg;
    a.seg_stride = cur_stride;
    a.seg_cap = cur_cap;
    a.group_size = gs;
    a.carry_in = cin;
    a.carry_cnt_in = ccnt;
    a.k = k;
    a.id_offset = id_offset;
    if (groups == 1) {
      a.keys_out = carry_out;
      a.cnt_out = carry_cnt_out;
      a.tau_out = tau_out;
      a.out_scores = out_scores;
      a.out_ids = out_ids;
      select_topk_kernel<<<dim3(static_cast<unsigned>(Q), 1), kSelectThreads, 0, st>>>(a);
      ICR_LAUNCH_CHECK();
      return ICR_OK;
    }
    const size_t keys_b = align_up(static_cast<size_t>(Q) * groups * k * sizeof(uint64_t), 256);
    const size_t cnts_b = align_up(static_cast<size_t>(Q) * groups * sizeof(int), 256);
    const size_t half = (scratch_bytes - 256) / 2;
    if (keys_b + cnts_b > half) {
      set_error("select: scratch too small (%zu needed per level, %zu available)", keys_b + cnts_b, half);
      return ICR_ERR_WORKSPACE;
    }
    char* basep = static_cast<char*>(scratch) + (level & 1) * half;
    a.keys_out = reinterpret_cast<uint64_t*>(basep);
    a.cnt_out = reinterpret_cast<int*>(basep + keys_b);
    select_topk_kernel<<<dim3(static_cast<unsigned>(Q), groups), kSelectThreads, 0, st>>>(a);
    ICR_LAUNCH_CHECK();
    cur_keys = a.keys_out;
    cur_cnt = a.cnt_out;
    cur_nseg = groups;
    cur_stride = k;
    cur_cap = k;
    cin = nullptr;  // carry was folded into group 0
    ccnt = nullptr;
    ++level;
  }
}

}  // namespace icr
