// K1: streaming cosine GEMV with fused top-k for small query batches (Q <= 7 per pass).
//
// HBM-bound by design: every catalog row is read exactly once with 128-bit streaming loads
// (ld.global.nc.L1::no_allocate), RB rows x up to 3 vectors per lane in flight; the row's squared
// norm is accumulated from the same registers, so no separate normalisation pass or inverse-norm
// array is read. Scores never reach HBM: each CTA keeps, per query, a shared-memory candidate list
// guarded by a running threshold (the CTA's k-th best so far) and emits its k best keys; the select
// kernel (select.cu) merges the per-CTA lists.
//
// Replaces, for one request: torch.tensor(catalog) + F.normalize x2 + torch.mm + argsort
// (reference src/inference/serve_recommendations.py:213-215 via sentence_transformers.util.cos_sim).
#include "common.cuh"

namespace icr {

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kGemvCand = 1024;  // candidate keys per query held in shared memory
constexpr int kGemvUnroll = 3;   // 16-byte vectors per lane per row kept in flight

struct GemvArgs {
  const void* cat;
  int64_t N;
  int64_t ldc;
  int D;
  const void* q;  // [Q][D] same dtype as the catalog
  int64_t ldq;
  int Q;
  const uint8_t* mask;  // optional exclusion mask
  int k;
  int64_t rows_per_cta;
  uint64_t* part_keys;  // [Q][gridDim.x][k]
  int* part_cnt;        // [Q][gridDim.x]
  int q0;               // first query of this pass
};

// RB rows per warp batch, QT queries per pass; V = RB * (QT + 1) partial sums per lane
// (QT dot products + the row's squared norm), V in {16, 32}.
template <typename T, int RB, int QT>
__global__ void __launch_bounds__(kGemvThreads) gemv_topk_kernel(GemvArgs a) {
  constexpr int VEC = Elem<T>::VEC;
  constexpr int V = RB * (QT + 1);
  constexpr int SH = (V == 32) ? 0 : (V == 16 ? 1 : (V == 8 ? 2 : 3));  // lane -> value index shift
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: keys [QT][kGemvCand] | qs [QT][Dv*VEC] f32 | ncand [QT] | tau [QT]
  uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw);
  const int nvec = a.D / VEC;
  const int dpad = nvec * VEC;
  float* qs = reinterpret_cast<float*>(cand + QT * kGemvCand);
  int* ncand = reinterpret_cast<int*>(qs + QT * dpad);
  float* tau = reinterpret_cast<float*>(ncand + QT);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* cat = static_cast<const T*>(a.cat);
  const T* qg = static_cast<const T*>(a.q);
  const int nq = min(QT, a.Q - a.q0);

  // ---- stage the L2-normalised queries in shared memory (fp32) ------------------------------
  for (int t = warp; t < QT; t += kGemvWarps) {
    if (t < nq) {
      const T* qrow = qg + static_cast<int64_t>(a.q0 + t) * a.ldq;
      float ss = 0.f;
      for (int e = lane; e < dpad; e += 32) {
        const float f = Elem<T>::to_f32(qrow[e]);
        qs[t * dpad + e] = f;
        ss = fmaf(f, f, ss);
      }
      ss = warp_sum(ss);
      const float inv = 1.0f / fmaxf(sqrtf(ss), kNormEps);
      __syncwarp();
      for (int e = lane; e < dpad; e += 32) qs[t * dpad + e] *= inv;
    } else {
      for (int e = lane; e < dpad; e += 32) qs[t * dpad + e] = 0.f;
    }
  }
  if (tid < QT) {
    ncand[tid] = 0;
    tau[tid] = -INFINITY;
  }
  __syncthreads();

  const int64_t row_begin = static_cast<int64_t>(blockIdx.x) * a.rows_per_cta;
  const int64_t row_end = min(a.N, row_begin + a.rows_per_cta);
  constexpr int kBlockRows = kGemvWarps * RB * 4;  // rows between two threshold refreshes

  for (int64_t blk = row_begin; blk < row_end; blk += kBlockRows) {
    const int64_t blk_end = min(row_end, blk + kBlockRows);
    for (int64_t r0 = blk + warp * RB; r0 < blk_end; r0 += kGemvWarps * RB) {
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;

      for (int vb = 0; vb < nvec; vb += 32 * kGemvUnroll) {
        uint4 c[RB][kGemvUnroll];
#pragma unroll
        for (int u = 0; u < kGemvUnroll; ++u) {
          const int v = vb + u * 32 + lane;
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            if (v < nvec && r0 + r < blk_end)
              c[r][u] = ldg_stream(cat + (r0 + r) * a.ldc + static_cast<int64_t>(v) * VEC);
            else
              c[r][u] = make_uint4(0, 0, 0, 0);
          }
        }
#pragma unroll
        for (int u = 0; u < kGemvUnroll; ++u) {
          const int v = vb + u * 32 + lane;
          if (v < nvec) {
            float qv[QT][VEC];
#pragma unroll
            for (int t = 0; t < QT; ++t) {
#pragma unroll
              for (int h = 0; h < VEC / 4; ++h) {
                const float4 x = *reinterpret_cast<const float4*>(qs + t * dpad + v * VEC + h * 4);
                qv[t][h * 4 + 0] = x.x;
                qv[t][h * 4 + 1] = x.y;
                qv[t][h * 4 + 2] = x.z;
                qv[t][h * 4 + 3] = x.w;
              }
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
              float f[VEC];
              Elem<T>::unpack(c[r][u], f);
#pragma unroll
              for (int t = 0; t < QT; ++t) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[t * RB + r] = fmaf(f[i], qv[t][i], acc[t * RB + r]);
              }
#pragma unroll
              for (int i = 0; i < VEC; ++i) acc[QT * RB + r] = fmaf(f[i], f[i], acc[QT * RB + r]);
            }
          }
        }
      }
      // scalar tail of D (D not a multiple of VEC): cooperative, rare
      if (dpad < a.D) {
        // not reachable: the host requires D % VEC == 0
      }

      warp_transpose_reduce<V>(acc, lane);
      const int idx = lane >> SH;          // value index owned by this lane
      const int t = idx / RB, r = idx % RB;
      // squared norm of row r lives in the lane group of index QT*RB + r
      const float ss = __shfl_sync(kFull, acc[0], (QT * RB + r) << SH);
      const int64_t row = r0 + r;
      if (t < nq && (lane & ((1 << SH) - 1)) == 0 && row < blk_end) {
        const bool excluded = a.mask && a.mask[row];
        const float score = acc[0] * (1.0f / fmaxf(sqrtf(ss), kNormEps));
        if (!excluded && score > tau[t]) {
          const int pos = atomicAdd(&ncand[t], 1);
          if (pos < kGemvCand) cand[t * kGemvCand + pos] = make_key(score, static_cast<uint32_t>(row));
        }
      }
    }
    __syncthreads();
    // ---- threshold refresh: if a list could overflow during the next block, keep its k best ----
    const bool last = (blk + kBlockRows >= row_end);
    for (int t = 0; t < nq; ++t) {
      const int n = min(ncand[t], kGemvCand);
      if (last || n > kGemvCand - kBlockRows) {
        uint64_t* keys = cand + t * kGemvCand;
        const int P = next_pow2(n < 2 ? 2 : n);
        for (int i = n + tid; i < P; i += kGemvThreads) keys[i] = 0ull;
        block_bitonic_sort_desc(keys, P);
        if (tid == 0) {
          const int kept = min(n, a.k);
          ncand[t] = kept;
          tau[t] = (kept >= a.k) ? key_score(keys[a.k - 1]) : -INFINITY;
        }
        __syncthreads();
      }
    }
  }

  // ---- emit this CTA's k best keys per query (sorted descending) -------------------------------
  for (int t = 0; t < nq; ++t) {
    const int n = (row_begin < row_end) ? ncand[t] : 0;
    const int64_t slot = static_cast<int64_t>(a.q0 + t) * gridDim.x + blockIdx.x;
    for (int i = tid; i < n; i += kGemvThreads) a.part_keys[slot * a.k + i] = cand[t * kGemvCand + i];
    if (tid == 0) a.part_cnt[slot] = n;
  }
}

static size_t gemv_smem_bytes(int QT, int D) {
  return static_cast<size_t>(QT) * kGemvCand * sizeof(uint64_t) + static_cast<size_t>(QT) * D * sizeof(float) + QT * 8 + 16;
}

template <typename T, int RB, int QT>
static int launch_one(const GemvArgs& a, int grid, cudaStream_t st) {
  const size_t smem = gemv_smem_bytes(QT, a.D);
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    ICR_CUDA_CHECK(cudaFuncSetAttribute(gemv_topk_kernel<T, RB, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  profile_begin(kKernelGemv, 1, st);
  gemv_topk_kernel<T, RB, QT><<<grid, kGemvThreads, smem, st>>>(a);
  profile_end(st);
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

int gemv_grid(int64_t N) {
  // 2 CTAs per SM when the catalog is large enough to give each CTA >= 64 rows
  int64_t g = (N + 63) / 64;
  if (g > 148 * 2) g = 148 * 2;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// Runs ceil(Q/7) passes (one for Q <= 7). part_keys/part_cnt sized [Q][grid][k] / [Q][grid].
int launch_gemv_topk(const void* cat, int64_t N, int64_t ldc, int D, int dtype, const void* q, int64_t ldq, int Q,
                     const uint8_t* mask, int k, uint64_t* part_keys, int* part_cnt, int grid, cudaStream_t st) {
  GemvArgs a{};
  a.cat = cat;
  a.N = N;
  a.ldc = ldc;
  a.D = D;
  a.q = q;
  a.ldq = ldq;
  a.Q = Q;
  a.mask = mask;
  a.k = k;
  a.rows_per_cta = (N + grid - 1) / grid;
  a.part_keys = part_keys;
  a.part_cnt = part_cnt;
  for (int q0 = 0; q0 < Q;) {
    a.q0 = q0;
    const int rem = Q - q0;
    int rc;
    if (dtype == ICR_F32) {
      if (rem == 1) rc = launch_one<float, 8, 1>(a, grid, st), q0 += 1;
      else if (rem <= 3) rc = launch_one<float, 8, 3>(a, grid, st), q0 += 3;
      else rc = launch_one<float, 4, 7>(a, grid, st), q0 += 7;
    } else {
      if (rem == 1) rc = launch_one<__nv_bfloat16, 8, 1>(a, grid, st), q0 += 1;
      else if (rem <= 3) rc = launch_one<__nv_bfloat16, 8, 3>(a, grid, st), q0 += 3;
      else rc = launch_one<__nv_bfloat16, 4, 7>(a, grid, st), q0 += 7;
    }
    if (rc != ICR_OK) return rc;
  }
  return ICR_OK;
}

}  // namespace icr
