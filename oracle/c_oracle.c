/* CPU restatement of the retrieval arithmetic in plain C — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * An implementation of the same definitions that shares no code with torch: it cross-checks oracle/oracle.py
 * (tests/test_oracle_golden.py) so that a mistake in one restatement cannot hide in the other. Double precision
 * throughout. Only tests/ may load the shared object this builds (oracle/_build/liboracle_c.so).
 *
 *   oc_cos_topk  == torch.topk(sentence_transformers.util.cos_sim(q, c), k, dim=1)   (ST 5.2.2, restated: both
 *                   operands L2-normalised with eps 1e-12, then the dot product; reference call sites
 *                   src/inference/serve_recommendations.py:213-225, src/baselines/content_based.py:54-63),
 *                   ties broken by the lower row
 *   oc_mnrl_loss == CrossEntropyLoss()(cos_sim(A, P) * scale, arange(B))              (MultipleNegativesRankingLoss,
 *                   built at src/training/train_sbert.py:182-185)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static double inv_norm(const float* x, int64_t d) {
  double ss = 0.0;
  for (int64_t i = 0; i < d; ++i) ss += (double)x[i] * (double)x[i];
  double n = sqrt(ss);
  return 1.0 / (n > 1e-12 ? n : 1e-12);
}

static double cosine(const float* a, double ia, const float* b, double ib, int64_t d) {
  double s = 0.0;
  for (int64_t i = 0; i < d; ++i) s += ((double)a[i] * ia) * ((double)b[i] * ib);
  return s;
}

typedef struct {
  double score;
  int64_t row;
} cand_t;

static int by_score_desc_row_asc(const void* pa, const void* pb) {
  const cand_t* a = (const cand_t*)pa;
  const cand_t* b = (const cand_t*)pb;
  if (a->score != b->score) return a->score > b->score ? -1 : 1;
  return a->row < b->row ? -1 : (a->row > b->row ? 1 : 0);
}

/* queries [Q, D], catalog [N, D] row-major fp32; out_scores [Q, k] f64, out_ids [Q, k] i64; returns 0, or -1 if out of memory */
int oc_cos_topk(const float* queries, int64_t Q, const float* catalog, int64_t N, int64_t D, int64_t k, double* out_scores, int64_t* out_ids) {
  double* cinv = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(N > 0 ? N : 1));
  if (!cinv || !c) return -1;
  for (int64_t r = 0; r < N; ++r) cinv[r] = inv_norm(catalog + r * D, D);
  for (int64_t q = 0; q < Q; ++q) {
    const double qi = inv_norm(queries + q * D, D);
    for (int64_t r = 0; r < N; ++r) {
      c[r].score = cosine(queries + q * D, qi, catalog + r * D, cinv[r], D);
      c[r].row = r;
    }
    qsort(c, (size_t)N, sizeof(cand_t), by_score_desc_row_asc);
    for (int64_t j = 0; j < k; ++j) {
      out_scores[q * k + j] = j < N ? c[j].score : -INFINITY;
      out_ids[q * k + j] = j < N ? c[j].row : -1;
    }
  }
  free(cinv);
  free(c);
  return 0;
}

/* anchors, positives [B, D] fp32; returns the mean cross-entropy of scale * cos_sim against the diagonal */
double oc_mnrl_loss(const float* a, const float* p, int64_t B, int64_t D, double scale) {
  double total = 0.0;
  for (int64_t i = 0; i < B; ++i) {
    const double ia = inv_norm(a + i * D, D);
    double mx = -INFINITY, diag = 0.0;
    for (int64_t j = 0; j < B; ++j) {
      const double s = scale * cosine(a + i * D, ia, p + j * D, inv_norm(p + j * D, D), D);
      if (s > mx) mx = s;
      if (j == i) diag = s;
    }
    double sum = 0.0;
    for (int64_t j = 0; j < B; ++j) sum += exp(scale * cosine(a + i * D, ia, p + j * D, inv_norm(p + j * D, D), D) - mx);
    total += mx + log(sum) - diag;
  }
  return total / (double)B;
}
