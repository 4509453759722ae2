// K3t: MultipleNegativesRankingLoss on the tensor cores, for batches where the in-batch similarity is a real GEMM.
//
//   loss = mean_i [ logsumexp_j( s * <a^_i, p^_j> ) - s * <a^_i, p^_i> ],   x^ = x / max(|x|, eps)
//
// The CUDA-core kernels of mnrl.cu are latency-optimal for tiny batches but scale as B^2 * D scalar FMAs with a
// warp reduction per score (B = 4096: 7 ms). Here every B^2 * D product runs as tcgen05.mma.cta_group::2 on CTA
// pairs, with the same fp32-parity operand format as K2 (x^ * 2^8 split into fp16 hi + lo, three MMA terms,
// fp32 accumulation in TMEM) so that scale * cosine keeps the 1e-4 loss tolerance at scale 20-30:
//
//   forward   prep (normalise, planes, inverse norms)
//             -> mnrl_tc_kernel<LSE>: S tile in TMEM, epilogue thread = anchor row keeps an online log-sum-exp
//                across the column tiles of its work item and picks up the diagonal; partial (max, sum) per chunk
//             -> finish: combine the partials in a fixed order, row losses, deterministic mean
//   backward  prep (planes again + fp16 transposes A^^T, P^^T: the K-major operands of the gradient products)
//             -> mnrl_tc_kernel<G>: recompute S, epilogue writes W = softmax - I as fp16 in both orientations
//                (row i owns 32 consecutive columns -> W[i, j..j+31]; a warp's 32 rows are contiguous in W^T[j, :])
//             -> mnrl_tc_kernel<MM>: dA^ = W P^ and dP^ = W^T A^ as one-term fp16 MMAs (K = B)
//             -> jacobian: dx = coef * (dx^ - x^ <x^, dx^>) / |x|,  coef = dL/dloss * s / B
// The [B, B] score matrix never reaches HBM in fp32; W (fp16, two orientations) does, because both gradient
// products need it as a K-major operand.
//
// Replaces sentence_transformers.losses.MultipleNegativesRankingLoss.forward + autograd (constructed at reference
// src/training/train_sbert.py:182-185) for B >= kMnrlTcMinBatch; smaller batches stay on mnrl.cu.
#include "tc.cuh"

namespace icr {

namespace {

constexpr int kTcRingBytes = 192 * 1024;
constexpr int kTcEpiWarps = 8;  // warps 4-7: accumulator columns 0-127, warps 8-11: columns 128-255 (same 128 rows)
constexpr int kTcThreads = 128 + kTcEpiWarps * 32;
constexpr int kTcTmemCols = 512;
constexpr int kTcMaxStages = 6;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// MM3: the gradient products with W and x^^T split into fp16 hi + lo planes (three MMA terms, like the scores): small batches,
// whose gradient entries are large (coef = scale / B) and would carry ~2^-11 relative from single fp16 operands
enum { MODE_LSE = 0, MODE_G = 1, MODE_MM = 2, MODE_MM3 = 3, MODE_G3 = 4 };  // G3: G that also writes the lo planes

struct TcArgs {
  int B, D;     // anchors (rows of the score matrix), embedding dim
  int Bc;       // candidates (columns): B, or G*B with cross-device negatives
  int label_off;  // the positive of anchor i is candidate i + label_off (rank * B with gathered candidates)
  int kb[2];      // 64-element K blocks per MMA term (MM: of product z)
  int qblocks[2];  // blocks of 256 A-operand rows (MM: of product z)
  int rows[2];     // MM: live rows of product z's output (B, Bc)
  int plane_stride;  // element offset of the lo plane in a plane row (LSE / G)
  int mm_stride[2];  // MM3: element offset of the lo plane in the rows of product z's operands (ldw, ldwt)
  int tiles;    // 256-column tiles over the N extent (Bc for LSE / G, D for MM)
  int chunks;   // balanced tile ranges; a work item = (row block, chunk) [of product z for MM]
  float c2;     // LSE / G: scale * 2^-16 * log2(e): accumulator units -> base-2 logits
  float cnat;   // LSE: scale * 2^-16 (natural-unit logit of the diagonal)
  // LSE outputs
  float* part_m;  // [B][chunks * 2] running max (base-2 logits) of (chunk, column half)
  float* part_l;  // [B][chunks * 2] sum of exp2(logit - max)
  float* diag;    // [B] natural-unit diagonal logit
  // G
  const float* lse;  // [B] natural units
  __half* w;         // [B][ldw]   softmax - I      (ldw = Bc rounded up to 64)
  __half* wt;        // [Bc][ldwt] its transpose    (ldwt = B rounded up to 64)
  int64_t ldw, ldwt;
  // MM
  float* raw[2];  // [B][D] dA^, [Bc][D] dP^ before coef and the normalisation Jacobian
};

__device__ __forceinline__ int tc_chunk_first(const TcArgs& g, int c) { return static_cast<int>(static_cast<int64_t>(c) * g.tiles / g.chunks); }

// D[256, 256] (+)= A[256, K] * B[256, K]^T per CTA pair; warp roles as in gemm_topk.cu.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
mnrl_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_b0,
               const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1, const TcArgs g) {
  constexpr bool IS_MM = MODE == MODE_MM || MODE == MODE_MM3;
  constexpr bool IS_G = MODE == MODE_G || MODE == MODE_G3, SPLIT = MODE == MODE_G3;
  constexpr int TERMS = MODE == MODE_MM ? 1 : 3;
  constexpr int kStageTiles = TERMS == 3 ? 4 : 2;  // A_hi A_lo B_hi B_lo | A B
  constexpr int kStageBytes = kStageTiles * kTileBytes;
  constexpr int kStages = kTcRingBytes / kStageBytes;  // 3 | 6
  static_assert(kStages <= kTcMaxStages, "barrier array too small");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcRingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kTcMaxStages;
  uint64_t* tfull_bar = bars + 2 * kTcMaxStages;
  uint64_t* tempty_bar = bars + 2 * kTcMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kTcMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int items0 = g.qblocks[0] * g.chunks;  // work items of product 0 (the only one for LSE / G)
  const int items = items0 + (IS_MM ? g.qblocks[1] * g.chunks : 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 2 * kTcEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b0) : "memory");
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(tmem_ptr), kTcTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    int stage = 0;
    uint32_t phase = 0;
    for (int w = pair; w < items; w += npairs) {
      const int z = w >= items0 ? 1 : 0, rem = w - z * items0;
      const int chunk = rem / g.qblocks[z], qb = rem - chunk * g.qblocks[z];
      const int t0 = tc_chunk_first(g, chunk), t1 = tc_chunk_first(g, chunk + 1);
      const int KB = g.kb[z];
      const CUtensorMap* ma = (IS_MM && z == 1) ? &map_a1 : &map_a0;
      const CUtensorMap* mb = (IS_MM && z == 1) ? &map_b1 : &map_b0;
      const int pstride = IS_MM ? g.mm_stride[z] : g.plane_stride;
      const int arow = qb * (2 * BM) + static_cast<int>(rank) * BM;
      for (int tile = t0; tile < t1; ++tile) {
        const int brow = tile * BN + static_cast<int>(rank) * BNH;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          if (rank == 0) mbar_expect_tx(fb, 2 * kStageBytes);
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          if (TERMS == 3) {
            tma_load_2d_pair(sa, ma, kb * BK, arow, fb);
            tma_load_2d_pair(sa + kTileBytes, ma, pstride + kb * BK, arow, fb);
            tma_load_2d_pair(sa + 2 * kTileBytes, mb, kb * BK, brow, fb);
            tma_load_2d_pair(sa + 3 * kTileBytes, mb, pstride + kb * BK, brow, fb);
          } else {
            tma_load_2d_pair(sa, ma, kb * BK, arow, fb);
            tma_load_2d_pair(sa + kTileBytes, mb, kb * BK, brow, fb);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================= MMA issuer (leader CTA; whole warp walks the loop, one elected lane issues) =================
    const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>((2 * BM) >> 4) << 24);  // f16 x f16 -> f32
    const uint64_t st0 = smem_desc_sw128(smem_u32(smem));
    constexpr uint64_t kTileUnits = kTileBytes >> 4, kStageUnits = kStageBytes >> 4;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int w = pair; w < items; w += npairs) {
      const int z = w >= items0 ? 1 : 0, rem = w - z * items0;
      const int chunk = rem / g.qblocks[z];
      const int t0 = tc_chunk_first(g, chunk), t1 = tc_chunk_first(g, chunk + 1);
      const int KB = g.kb[z];
      for (int tile = t0; tile < t1; ++tile) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint64_t sd = st0 + static_cast<uint64_t>(stage) * kStageUnits;
          if (elect_one()) {
            if (TERMS == 3) {
              const uint64_t a_hi = sd, a_lo = sd + kTileUnits, b_hi = sd + 2 * kTileUnits, b_lo = sd + 3 * kTileUnits;
#pragma unroll
              for (int k4 = 0; k4 < BK / 16; ++k4) {
                const uint64_t o = static_cast<uint64_t>(k4 * 2);
                umma_f16_pair(d_tmem, a_hi + o, b_hi + o, idesc, (kb | k4) != 0 ? 1u : 0u);
                umma_f16_pair(d_tmem, a_hi + o, b_lo + o, idesc, 1u);
                umma_f16_pair(d_tmem, a_lo + o, b_hi + o, idesc, 1u);
              }
            } else {
              const uint64_t adesc = sd, bdesc = sd + kTileUnits;
#pragma unroll
              for (int k4 = 0; k4 < BK / 16; ++k4) {
                const uint64_t o = static_cast<uint64_t>(k4 * 2);
                umma_f16_pair(d_tmem, adesc + o, bdesc + o, idesc, (kb | k4) != 0 ? 1u : 0u);
              }
            }
            umma_commit_pair(smem_u32(&empty_bar[stage]));
            if (kb == KB - 1) umma_commit_pair(smem_u32(&tfull_bar[acc]));
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: thread = row of the A operand =================
    const int ew = (warp - 4) & 3;     // TMEM lane quarter (hardware rule: warp id % 4)
    const int half = (warp - 4) >> 2;  // which 128 accumulator columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = pair; w < items; w += npairs) {
      const int z = w >= items0 ? 1 : 0, rem = w - z * items0;
      const int chunk = rem / g.qblocks[z], qb = rem - chunk * g.qblocks[z];
      const int t0 = tc_chunk_first(g, chunk), t1 = tc_chunk_first(g, chunk + 1);
      const int row_w = qb * (2 * BM) + static_cast<int>(rank) * BM + ew * 32;  // first row of this warp
      const int row = row_w + lane;
      const bool live = row < (IS_MM ? g.rows[z] : g.B);
      const int pos = row + g.label_off;  // column of this row's positive
      // per-item state
      float m = -INFINITY, l = 0.f, dg = 0.f;
      bool have_diag = false;
      const float lse2 = (IS_G && live) ? g.lse[row] * kLog2e : 0.f;
      for (int tile = t0; tile < t1; ++tile) {
        const int col_base = tile * BN + half * (BN / 2);
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * BN + half * (BN / 2));
        uint32_t r[2][32];
        tmem_ld32(taddr, r[0]);
#pragma unroll
        for (int cb = 0; cb < BN / 2 / 32; ++cb) {
          uint32_t(&cur)[32] = r[cb & 1];
          tmem_ld_wait(cur);
          if (cb + 1 < BN / 2 / 32) tmem_ld32(taddr + (cb + 1) * 32, r[(cb + 1) & 1]);
          const int col0 = col_base + cb * 32;
          if (MODE == MODE_LSE) {
            if (col0 < g.Bc) {  // warp-uniform
              float v[32];
              float mx = -INFINITY;
              const bool full = col0 + 32 <= g.Bc;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[j] = __uint_as_float(cur[j]) * g.c2;
                if (!full && col0 + j >= g.Bc) v[j] = -INFINITY;
                mx = fmaxf(mx, v[j]);
              }
              const float mn = fmaxf(m, mx);
              float sum = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) sum += exp2f(v[j] - mn);
              l = l * exp2f(m - mn) + sum;
              m = mn;
              if (col0 < row_w + g.label_off + 32 && col0 + 32 > row_w + g.label_off) {  // warp-uniform: the positives of this warp's rows
                float d = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) d = (col0 + j == pos) ? __uint_as_float(cur[j]) : d;
                if (pos >= col0 && pos < col0 + 32) {
                  dg = d * g.cnat;
                  have_diag = true;
                }
              }
            }
          } else if (IS_G) {
            if (col0 < g.ldw) {  // ldw = B rounded up to 64: a 32-column group is inside or outside as a whole
              __align__(16) __half h[32], lo[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float p = exp2f(__uint_as_float(cur[j]) * g.c2 - lse2);
                if (col0 + j == pos) p -= 1.0f;
                if (col0 + j >= g.Bc) p = 0.f;
                h[j] = __float2half_rn(p);
                if (SPLIT) lo[j] = __float2half_rn(p - __half2float(h[j]));
              }
              if (live) {
                const int64_t wld = SPLIT ? 2 * g.ldw : g.ldw, wtld = SPLIT ? 2 * g.ldwt : g.ldwt;
                uint4* dst = reinterpret_cast<uint4*>(g.w + static_cast<int64_t>(row) * wld + col0);
                const uint4* srcv = reinterpret_cast<const uint4*>(h);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) dst[q4] = srcv[q4];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < g.Bc) g.wt[static_cast<int64_t>(col0 + j) * wtld + row] = h[j];
                if (SPLIT) {
                  uint4* dlo = reinterpret_cast<uint4*>(g.w + static_cast<int64_t>(row) * wld + g.ldw + col0);
                  const uint4* slo = reinterpret_cast<const uint4*>(lo);
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) dlo[q4] = slo[q4];
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (col0 + j < g.Bc) g.wt[static_cast<int64_t>(col0 + j) * wtld + g.ldwt + row] = lo[j];
                }
              }
            }
          } else {
            if (live && col0 < g.D) {
              float* dst = g.raw[z] + static_cast<int64_t>(row) * g.D + col0;
              if (col0 + 32 <= g.D) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1]),
                                                                    __uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < g.D) dst[j] = __uint_as_float(cur[j]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&tempty_bar[acc]), 0);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (MODE == MODE_LSE && live) {
        const int64_t pi = static_cast<int64_t>(row) * (g.chunks * 2) + chunk * 2 + half;
        g.part_m[pi] = m;
        g.part_l[pi] = l;
        if (have_diag) g.diag[row] = dg;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTcTmemCols);
  }
}

// ---- prep: both matrices in one launch. blockIdx.x < blocks_per -> anchors, else positives --------------------------------
// One warp per row, 32 rows per CTA, 128-bit loads: normalise in fp32, x^ * 2^8 -> fp16 hi | lo planes (prep.cu's format),
// inverse norm; TRANSPOSE also stages x^ as fp16 in shared memory and writes it transposed ([D, ldt], a warp stores the
// 32 rows of one column as 64 contiguous bytes). Block 0 also zeroes the forward's ticket counter.
constexpr int kPrepRows = 32;
constexpr int kPrepThreads = kPrepRows * 32;

template <typename T, bool TRANSPOSE>
__global__ void __launch_bounds__(kPrepThreads) mnrl_tc_prep_kernel(const T* __restrict__ a, int64_t lda, const T* __restrict__ p, int64_t ldp,
                                                                    int Ba, int Bp, int D, int dpad, __half* __restrict__ planes_a,
                                                                    __half* __restrict__ planes_p, float* __restrict__ inv_a,
                                                                    float* __restrict__ inv_p, __half* __restrict__ at, __half* __restrict__ pt,
                                                                    int64_t ldat, int64_t ldpt, unsigned int* __restrict__ ticket, int split) {
  constexpr int VEC = Elem<T>::VEC;
  extern __shared__ __align__(16) unsigned char prep_smem[];
  float* xs = reinterpret_cast<float*>(prep_smem);  // [kPrepRows][dpad + 1] normalised rows (fp32: the split transposes need hi AND lo)
  const int blocks_a = (Ba + kPrepRows - 1) / kPrepRows;
  const bool is_a = static_cast<int>(blockIdx.x) < blocks_a;
  const int row0 = (is_a ? blockIdx.x : blockIdx.x - blocks_a) * kPrepRows;
  const int B = is_a ? Ba : Bp;
  const int64_t ldt = is_a ? ldat : ldpt;
  const T* X = is_a ? a : p;
  const int64_t ldx = is_a ? lda : ldp;
  __half* planes = is_a ? planes_a : planes_p;
  float* inv_out = is_a ? inv_a : inv_p;
  __half* xt = is_a ? at : pt;
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sld = dpad + 1;
  const int row = row0 + r;
  const int nvec = D / VEC, nvec_pad = dpad / VEC;
  if (blockIdx.x == 0 && threadIdx.x == 0 && ticket) *ticket = 0u;
  if (row < B) {
    const T* src = X + static_cast<int64_t>(row) * ldx;
    float ss = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float f[VEC];
      Elem<T>::unpack(*reinterpret_cast<const uint4*>(src + static_cast<int64_t>(v) * VEC), f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) ss = fmaf(f[i], f[i], ss);
    }
    ss = warp_sum(ss);
    const float nrm = fmaxf(sqrtf(ss), kNormEps);
    const float inv = 1.0f / nrm, inv256 = 256.0f / nrm;
    if (lane == 0 && inv_out) inv_out[row] = inv;
    __half* hi = planes + static_cast<int64_t>(row) * (2 * dpad);
    __half* lo = hi + dpad;
    for (int v = lane; v < nvec_pad; v += 32) {
      float f[VEC];
      if (v < nvec) Elem<T>::unpack(*reinterpret_cast<const uint4*>(src + static_cast<int64_t>(v) * VEC), f);
      else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) f[i] = 0.f;
      }
      __align__(16) __half h[VEC], l[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float y = f[i] * inv256;
        h[i] = __float2half_rn(y);
        l[i] = __float2half_rn(y - __half2float(h[i]));
        if (TRANSPOSE && v < nvec) xs[r * sld + v * VEC + i] = f[i] * inv;
      }
      if (VEC == 8) {
        *reinterpret_cast<uint4*>(hi + v * VEC) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(lo + v * VEC) = *reinterpret_cast<const uint4*>(l);
      } else {
        *reinterpret_cast<uint2*>(hi + v * VEC) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo + v * VEC) = *reinterpret_cast<const uint2*>(l);
      }
    }
  }
  if (TRANSPOSE) {
    __syncthreads();
    if (row0 + lane < B) {
      const int64_t rowld = split ? 2 * ldt : ldt;  // split: [D][hi plane (ldt) | lo plane (ldt)]
      for (int d = r; d < D; d += kPrepRows) {
        const float x = xs[lane * sld + d];
        const __half h = __float2half_rn(x);
        xt[static_cast<int64_t>(d) * rowld + row0 + lane] = h;
        if (split) xt[static_cast<int64_t>(d) * rowld + ldt + row0 + lane] = __float2half_rn(x - __half2float(h));
      }
    }
  }
}

// ---- forward finish: combine the (chunk, half) partials per row, row losses, deterministic mean -----------------------
// thread = row; every CTA leaves the sum of its 256 row losses, the last CTA to take a ticket adds those in CTA order
__global__ void __launch_bounds__(256) mnrl_tc_finish_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                                            const float* __restrict__ diag, int B, int nparts, float* __restrict__ lse,
                                                            float* __restrict__ cta_sums, unsigned int* __restrict__ ticket,
                                                            float* __restrict__ loss) {
  __shared__ float s[256];
  __shared__ bool is_last;
  const int row = blockIdx.x * 256 + threadIdx.x;
  float v = 0.f;
  if (row < B) {
    const float* pm = part_m + static_cast<int64_t>(row) * nparts;
    const float* pl = part_l + static_cast<int64_t>(row) * nparts;
    float M = -INFINITY;
    for (int q = 0; q < nparts; ++q) M = fmaxf(M, pm[q]);
    float L = 0.f;
    for (int q = 0; q < nparts; ++q) L += pl[q] * exp2f(pm[q] - M);  // empty part: 0 * exp2(-inf) = 0
    const float x = (M + log2f(L)) * kLn2;
    lse[row] = x;
    v = x - diag[row];
  }
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    cta_sums[blockIdx.x] = s[0];
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += 32) t += __ldcg(cta_sums + i);
    t = warp_sum(t);
    if (threadIdx.x == 0) loss[0] = t / static_cast<float>(B);
  }
}

// ---- backward finish: dx = coef * (dx^ - x^ <x^, dx^>) * inv, warp per row of either matrix ----------------------------
template <typename T>
__global__ void __launch_bounds__(256) mnrl_tc_jacobian_kernel(const T* __restrict__ a, int64_t lda, const T* __restrict__ p, int64_t ldp,
                                                               int B, int Bc, int D, const float* __restrict__ inv_a, const float* __restrict__ inv_p,
                                                               const float* __restrict__ raw_a, const float* __restrict__ raw_p,
                                                               const float* __restrict__ grad_out, float scale,
                                                               T* __restrict__ grad_a, int64_t ldga, T* __restrict__ grad_p, int64_t ldgp) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= B + Bc) return;
  const int z = gw >= B ? 1 : 0, row = gw - z * B;
  const T* x = (z ? p : a) + static_cast<int64_t>(row) * (z ? ldp : lda);
  const float inv = z ? inv_p[row] : inv_a[row];
  const float* dr = (z ? raw_p : raw_a) + static_cast<int64_t>(row) * D;
  T* out = (z ? grad_p : grad_a) + static_cast<int64_t>(row) * (z ? ldgp : ldga);
  const float coef = (grad_out ? grad_out[0] : 1.0f) * scale / static_cast<float>(B);  // null: dL/dloss = 1
  float pr = 0.f;
  for (int e = lane; e < D; e += 32) pr = fmaf(Elem<T>::to_f32(x[e]) * inv, dr[e], pr);
  pr = warp_sum(pr);
  for (int e = lane; e < D; e += 32) {
    const float xh = Elem<T>::to_f32(x[e]) * inv;
    out[e] = Elem<T>::from_f32(coef * (dr[e] - xh * pr) * inv);
  }
}

constexpr size_t kTcSmemBytes = static_cast<size_t>(kTcRingBytes) + (2 * kTcMaxStages + 4) * sizeof(uint64_t) + 16 + 1024;
static_assert(kTcSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");

// below this many anchors the gradient products use hi + lo planes: gradient entries scale with scale / B, and single fp16
// operands leave ~2^-11 of the largest entry (B = 32, scale 30: 2e-4 absolute, over the 1e-4 bar; B >= 288: < 5e-5)
#ifndef ICR_MNRL_TC_SPLIT_BELOW
#define ICR_MNRL_TC_SPLIT_BELOW 288
#endif

struct TcWs {
  size_t planes_a, planes_p, at, pt, w, wt, raw_a, raw_p, part_m, part_l, diag, cta_sums, ticket, total;
  int64_t dpad, ldw, ldwt;
  int chunks;
  int split;  // gradient products with hi + lo planes (at .. wt hold two planes per row and are zero-filled first)
};

// Work items = (row block, chunk of column tiles) over 74 CTA pairs. The launch takes ceil(items / 74) rounds of the
// longest item, so the chunk count is chosen to minimise rounds x tiles-per-item (B = 4096: 4 chunks -> 64 items of 4
// tiles in one round, where ceil(148 / row blocks) = 10 chunks gave 160 items = 3 rounds of 2 tiles for 2.2 rounds of
// work); on ties fewer, longer items win (their epilogues overlap the next tile's MMAs).
int tc_chunks(int64_t B, int64_t Bc) {
  const int qblocks = static_cast<int>((B + 2 * BM - 1) / (2 * BM));
  const int tiles = static_cast<int>((Bc + BN - 1) / BN);
  const int npairs = kNumSMs / 2;
  int best = 1;
  int64_t best_cost = INT64_MAX;
  for (int c = 1; c <= tiles; ++c) {
    const int64_t items = static_cast<int64_t>(qblocks) * c;
    const int64_t rounds = (items + npairs - 1) / npairs;
    const int64_t cost = rounds * ((tiles + c - 1) / c);
    if (cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  return best;
}

TcWs tc_layout(int64_t B, int64_t Bc, int64_t D) {
  TcWs w{};
  w.dpad = (D + 63) / 64 * 64;
  w.ldw = (Bc + 63) / 64 * 64;
  w.ldwt = (B + 63) / 64 * 64;
  w.chunks = tc_chunks(B, Bc);
  w.split = B < ICR_MNRL_TC_SPLIT_BELOW ? 1 : 0;
  const size_t planes = w.split ? 2 : 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += align_up(bytes, 1024);
    return o;
  };
  w.planes_a = take(static_cast<size_t>(B) * 2 * w.dpad * 2);
  w.planes_p = take(static_cast<size_t>(Bc) * 2 * w.dpad * 2);
  w.at = take(static_cast<size_t>(D) * w.ldwt * 2 * planes);
  w.pt = take(static_cast<size_t>(D) * w.ldw * 2 * planes);
  w.w = take(static_cast<size_t>(B) * w.ldw * 2 * planes);
  w.wt = take(static_cast<size_t>(Bc) * w.ldwt * 2 * planes);
  w.raw_a = take(static_cast<size_t>(B) * D * 4);
  w.raw_p = take(static_cast<size_t>(Bc) * D * 4);
  w.part_m = take(static_cast<size_t>(B) * w.chunks * 2 * 4);
  w.part_l = take(static_cast<size_t>(B) * w.chunks * 2 * 4);
  w.diag = take(static_cast<size_t>(B) * 4);
  w.cta_sums = take(static_cast<size_t>((B + 255) / 256) * 4);
  w.ticket = take(256);
  w.total = off + 1024;
  return w;
}

template <int MODE>
int launch_tc(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1, const CUtensorMap& b1, const TcArgs& g, cudaStream_t st) {
  static thread_local SmemAttrCache attr_cache;  // per device
  {
    const int rc = ensure_dyn_smem(attr_cache, mnrl_tc_kernel<MODE>, kTcSmemBytes);
    if (rc) return rc;
  }
  const int items = (g.qblocks[0] + ((MODE == MODE_MM || MODE == MODE_MM3) ? g.qblocks[1] : 0)) * g.chunks;
  const int npairs = kNumSMs / 2;
  const int grid = 2 * (items < npairs ? items : npairs);
  mnrl_tc_kernel<MODE><<<grid, kTcThreads, kTcSmemBytes, st>>>(a0, b0, a1, b1, g);
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

template <typename T>
int launch_prep(const MnrlArgs& m, const TcWs& w, char* base, bool transpose, bool forward, cudaStream_t st) {
  const int blocks = (m.B + kPrepRows - 1) / kPrepRows + (m.Bc + kPrepRows - 1) / kPrepRows;
  __half* pa = reinterpret_cast<__half*>(base + w.planes_a);
  __half* pp = reinterpret_cast<__half*>(base + w.planes_p);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(base + w.ticket);
  const int dpad = static_cast<int>(w.dpad);
  if (transpose) {
    const size_t smem = static_cast<size_t>(kPrepRows) * (dpad + 1) * sizeof(float);
    if (smem > 48 * 1024) ICR_CUDA_CHECK(cudaFuncSetAttribute(mnrl_tc_prep_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    mnrl_tc_prep_kernel<T, true><<<blocks, kPrepThreads, smem, st>>>(static_cast<const T*>(m.a), m.lda, static_cast<const T*>(m.p), m.ldp, m.B, m.Bc,
                                                                      m.D, dpad, pa, pp, forward ? m.inv_a : nullptr, forward ? m.inv_p : nullptr,
                                                                      reinterpret_cast<__half*>(base + w.at), reinterpret_cast<__half*>(base + w.pt),
                                                                      w.ldwt, w.ldw, forward ? ticket : nullptr, w.split);
  } else {
    mnrl_tc_prep_kernel<T, false><<<blocks, kPrepThreads, 0, st>>>(static_cast<const T*>(m.a), m.lda, static_cast<const T*>(m.p), m.ldp, m.B, m.Bc,
                                                                   m.D, dpad, pa, pp, m.inv_a, m.inv_p, nullptr, nullptr, 0, 0, ticket, 0);
  }
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

}  // namespace

// batches from this size up take the tensor-core path (measured crossover, profiles/r01_notes.md)
#ifndef ICR_MNRL_TC_MIN_BATCH
#define ICR_MNRL_TC_MIN_BATCH 288
#endif

bool mnrl_tc_applies(int64_t B, int64_t Bc, int64_t D) {
  static const char* force = getenv("ICR_MNRL_PATH");  // "tc" / "simt": A/B switch for benchmarks and tests
  if (D % 8 != 0 || D > 4096 || B < 1 || Bc < 2 || B > 65536 || Bc > (1 << 20)) return false;
  if (Bc != B) return true;  // gathered candidates (cross-device negatives): only this path handles a rectangular score matrix
  if (force && force[0] == 't') return true;
  if (force && force[0] == 's') return false;
  return B >= ICR_MNRL_TC_MIN_BATCH;
}

size_t mnrl_tc_workspace_bytes(int64_t B, int64_t Bc, int64_t D) { return tc_layout(B, Bc, D).total; }

// mode 0: forward; 1: backward (lse / inverse norms are inputs); 2: both in one go (gradients for dL/dloss = grad_out or 1)
int launch_mnrl_tc(const MnrlArgs& m, int dtype, int mode, void* ws, size_t ws_bytes, cudaStream_t st) {
  const bool fwd = mode != 1, bwd = mode != 0;
  const TcWs w = tc_layout(m.B, m.Bc, m.D);
  if (ws_bytes < w.total) {
    set_error("mnrl (tensor path): workspace %zu bytes < required %zu", ws_bytes, w.total);
    return ICR_ERR_WORKSPACE;
  }
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~static_cast<uintptr_t>(1023));
  int rc;
  // split planes lie INSIDE the tensor maps' bounds (no out-of-bounds zero fill between a hi and a lo plane): padding must be zero
  if (bwd && w.split) ICR_CUDA_CHECK(cudaMemsetAsync(base + w.at, 0, w.raw_a - w.at, st));
  if (dtype == ICR_F32) rc = launch_prep<float>(m, w, base, bwd, fwd, st);
  else if (dtype == ICR_F16) rc = launch_prep<__half>(m, w, base, bwd, fwd, st);
  else rc = launch_prep<__nv_bfloat16>(m, w, base, bwd, fwd, st);
  if (rc) return rc;

  TcArgs g{};
  g.B = m.B;
  g.Bc = m.Bc;
  g.label_off = m.label_off;
  g.D = m.D;
  g.kb[0] = g.kb[1] = static_cast<int>(w.dpad / BK);
  g.plane_stride = static_cast<int>(w.dpad);
  g.qblocks[0] = g.qblocks[1] = (m.B + 2 * BM - 1) / (2 * BM);
  g.rows[0] = g.rows[1] = m.B;
  g.tiles = (m.Bc + BN - 1) / BN;
  g.chunks = w.chunks;
  g.c2 = m.scale * (1.0f / 65536.0f) * kLog2e;
  g.cnat = m.scale * (1.0f / 65536.0f);
  g.part_m = reinterpret_cast<float*>(base + w.part_m);
  g.part_l = reinterpret_cast<float*>(base + w.part_l);
  g.diag = reinterpret_cast<float*>(base + w.diag);
  g.lse = m.lse;
  g.w = reinterpret_cast<__half*>(base + w.w);
  g.wt = reinterpret_cast<__half*>(base + w.wt);
  g.ldw = w.ldw;
  g.ldwt = w.ldwt;
  g.mm_stride[0] = static_cast<int>(w.ldw);
  g.mm_stride[1] = static_cast<int>(w.ldwt);
  g.raw[0] = reinterpret_cast<float*>(base + w.raw_a);
  g.raw[1] = reinterpret_cast<float*>(base + w.raw_p);

  CUtensorMap map_a, map_p;
  if ((rc = make_map(&map_a, base + w.planes_a, m.B, 2 * w.dpad, 2 * w.dpad, false))) return rc;
  if ((rc = make_map(&map_p, base + w.planes_p, m.Bc, 2 * w.dpad, 2 * w.dpad, false))) return rc;
  if (fwd) {
    if ((rc = launch_tc<MODE_LSE>(map_a, map_p, map_a, map_p, g, st))) return rc;
    mnrl_tc_finish_kernel<<<(m.B + 255) / 256, 256, 0, st>>>(g.part_m, g.part_l, g.diag, m.B, g.chunks * 2, m.lse,
                                                             reinterpret_cast<float*>(base + w.cta_sums),
                                                             reinterpret_cast<unsigned int*>(base + w.ticket), m.loss);
    ICR_LAUNCH_CHECK();
    if (!bwd) return ICR_OK;
  }
  if (w.split) rc = launch_tc<MODE_G3>(map_a, map_p, map_a, map_p, g, st);
  else rc = launch_tc<MODE_G>(map_a, map_p, map_a, map_p, g, st);
  if (rc) return rc;
  // gradient products: dA^ [B, D] = W P^ (A operand W [B, Bc], B operand P^^T [D, Bc], K = Bc)
  //                    dP^ [Bc, D] = W^T A^ (A operand W^T [Bc, B], B operand A^^T [D, B], K = B)
  CUtensorMap map_w, map_wt, map_at, map_pt;
  if (w.split) {  // rows = [hi plane | lo plane], both inside the map
    if ((rc = make_map(&map_w, base + w.w, m.B, 2 * w.ldw, 2 * w.ldw, false))) return rc;
    if ((rc = make_map(&map_wt, base + w.wt, m.Bc, 2 * w.ldwt, 2 * w.ldwt, false))) return rc;
    if ((rc = make_map(&map_at, base + w.at, m.D, 2 * w.ldwt, 2 * w.ldwt, false))) return rc;
    if ((rc = make_map(&map_pt, base + w.pt, m.D, 2 * w.ldw, 2 * w.ldw, false))) return rc;
  } else {
    if ((rc = make_map(&map_w, base + w.w, m.B, m.Bc, w.ldw, false))) return rc;
    if ((rc = make_map(&map_wt, base + w.wt, m.Bc, m.B, w.ldwt, false))) return rc;
    if ((rc = make_map(&map_at, base + w.at, m.D, m.B, w.ldwt, false))) return rc;
    if ((rc = make_map(&map_pt, base + w.pt, m.D, m.Bc, w.ldw, false))) return rc;
  }
  TcArgs mm = g;
  mm.kb[0] = static_cast<int>(w.ldw / BK);
  mm.kb[1] = static_cast<int>(w.ldwt / BK);
  mm.qblocks[0] = (m.B + 2 * BM - 1) / (2 * BM);
  mm.qblocks[1] = (m.Bc + 2 * BM - 1) / (2 * BM);
  mm.rows[0] = m.B;
  mm.rows[1] = m.Bc;
  mm.tiles = (m.D + BN - 1) / BN;
  mm.chunks = mm.tiles;  // one N tile per work item
  if (w.split) rc = launch_tc<MODE_MM3>(map_w, map_pt, map_wt, map_at, mm, st);
  else rc = launch_tc<MODE_MM>(map_w, map_pt, map_wt, map_at, mm, st);
  if (rc) return rc;
  const int jb = (m.B + m.Bc + 7) / 8;
  if (dtype == ICR_F32)
    mnrl_tc_jacobian_kernel<float><<<jb, 256, 0, st>>>(static_cast<const float*>(m.a), m.lda, static_cast<const float*>(m.p), m.ldp, m.B, m.Bc,
                                                       m.D, m.inv_a, m.inv_p, g.raw[0], g.raw[1], m.grad_out, m.scale,
                                                       static_cast<float*>(m.grad_a), m.ldga, static_cast<float*>(m.grad_p), m.ldgp);
  else if (dtype == ICR_F16)
    mnrl_tc_jacobian_kernel<__half><<<jb, 256, 0, st>>>(static_cast<const __half*>(m.a), m.lda, static_cast<const __half*>(m.p), m.ldp, m.B, m.Bc,
                                                        m.D, m.inv_a, m.inv_p, g.raw[0], g.raw[1], m.grad_out, m.scale,
                                                        static_cast<__half*>(m.grad_a), m.ldga, static_cast<__half*>(m.grad_p), m.ldgp);
  else
    mnrl_tc_jacobian_kernel<__nv_bfloat16><<<jb, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(m.a), m.lda, static_cast<const __nv_bfloat16*>(m.p),
                                                               m.ldp, m.B, m.Bc, m.D, m.inv_a, m.inv_p, g.raw[0], g.raw[1], m.grad_out, m.scale,
                                                               static_cast<__nv_bfloat16*>(m.grad_a), m.ldga,
                                                               static_cast<__nv_bfloat16*>(m.grad_p), m.ldgp);
  ICR_LAUNCH_CHECK();
  return ICR_OK;
}

}  // namespace icr
