"""Batched IR-evaluation scoring on the fused cosine top-k kernel.

Two consumers of the reference are served:

* ``InformationRetrievalEvaluator`` — same constructor / call / metric-key contract as the
  sentence-transformers evaluator the reference builds at src/training/train_sbert.py:197-202
  and reads at :220 (``eval_order-recommendation_cosine_ndcg@10``). Upstream it materialises
  cos_sim [Q, N], runs ``torch.topk(100, sorted=False)`` per 50,000-row corpus chunk and merges
  with a Python heap per query; here one fused call returns the sorted top-100 directly.
* ``rank_all`` / ``compute_ir_metrics`` — the batched ranking of
  src/baselines/content_based.py:38-64 and scripts/compare_untrained_vs_trained.py:38-85,
  whose only consumer (src/baselines/metrics.py:150-165) reads ranks <= 100, scored with the
  metric definitions of src/baselines/metrics.py:13-176.

Metric arithmetic runs on the device too (``icr_ir_metrics``: one warp per query over the [Q, k]
id matrix), so an evaluation moves M floats to the host instead of Q x k ids
plus Python loops. ``compute_metrics_from_ids`` / ``compute_ir_metrics`` keep the same arithmetic
for id matrices / string rankings that already live on the host.
"""

from __future__ import annotations

import logging
from typing import Callable

import numpy as np
import torch

from . import ops
from .similarity import cos_topk, to_device_matrix

logger = logging.getLogger(__name__)


def _encode(model, texts: list[str], batch_size: int, show_progress_bar: bool):
    try:
        return model.encode(texts, batch_size=batch_size, show_progress_bar=show_progress_bar, convert_to_tensor=True)
    except TypeError:
        return model.encode(texts, batch_size=batch_size, show_progress_bar=show_progress_bar)


def topk_ids_device(query_emb, corpus_emb, k: int, *, query_chunk: int = 16384):
    """Sorted top-k (scores f32 [Q,k], corpus rows int64 [Q,k]) as device tensors."""
    c = to_device_matrix(corpus_emb)
    q = to_device_matrix(query_emb, device=c.device)
    k = min(k, c.shape[0])
    if q.shape[0] <= query_chunk:
        return cos_topk(q, c, k)
    vals = torch.empty(q.shape[0], k, dtype=torch.float32, device=c.device)
    ids = torch.empty(q.shape[0], k, dtype=torch.int64, device=c.device)
    for s in range(0, q.shape[0], query_chunk):
        v, i = cos_topk(q[s : s + query_chunk], c, k)
        vals[s : s + query_chunk] = v
        ids[s : s + query_chunk] = i
    return vals, ids


def topk_ids(query_emb, corpus_emb, k: int, *, query_chunk: int = 16384):
    """Sorted top-k (scores f32 [Q,k], corpus rows int64 [Q,k]) as host numpy arrays."""
    v, i = topk_ids_device(query_emb, corpus_emb, k, query_chunk=query_chunk)
    return v.cpu().numpy(), i.cpu().numpy()


def _hit_matrix(ids: np.ndarray, relevant_rows: list[np.ndarray], n_corpus: int) -> np.ndarray:
    """hits[q, r] = 1 if the r-th retrieved row of query q is relevant to q."""
    Q, K = ids.shape
    if Q == 0:
        return np.zeros((0, K), dtype=bool)
    rel_keys = np.concatenate([np.asarray(r, dtype=np.int64) + qi * n_corpus for qi, r in enumerate(relevant_rows)] or [np.zeros(0, np.int64)])
    keys = ids.astype(np.int64) + (np.arange(Q, dtype=np.int64) * n_corpus)[:, None]
    return np.isin(keys, rel_keys) & (ids >= 0)


class InformationRetrievalEvaluator:
    """Cosine-similarity IR evaluator: accuracy/precision/recall@k, MRR@k, NDCG@k, MAP@k."""

    def __init__(
        self,
        queries: dict[str, str],
        corpus: dict[str, str],
        relevant_docs: dict[str, set[str]],
        corpus_chunk_size: int = 50000,
        mrr_at_k: list[int] = [10],
        ndcg_at_k: list[int] = [10],
        accuracy_at_k: list[int] = [1, 3, 5, 10],
        precision_recall_at_k: list[int] = [1, 3, 5, 10],
        map_at_k: list[int] = [100],
        show_progress_bar: bool = False,
        batch_size: int = 32,
        name: str = "",
        write_csv: bool = False,
        score_functions: dict[str, Callable] | None = None,
        main_score_function: str | None = None,
    ) -> None:
        if score_functions is not None and set(score_functions) != {"cosine"}:
            raise NotImplementedError("only the cosine score function (the reference's) is fused")
        self.queries_ids = [qid for qid in queries if qid in relevant_docs and len(relevant_docs[qid]) > 0]
        self.queries = [queries[qid] for qid in self.queries_ids]
        self.corpus_ids = list(corpus.keys())
        self.corpus = [corpus[cid] for cid in self.corpus_ids]
        self.relevant_docs = relevant_docs
        self.corpus_chunk_size = corpus_chunk_size  # kept for signature parity; the fused kernel needs no chunking
        self.mrr_at_k, self.ndcg_at_k = list(mrr_at_k), list(ndcg_at_k)
        self.accuracy_at_k, self.precision_recall_at_k, self.map_at_k = list(accuracy_at_k), list(precision_recall_at_k), list(map_at_k)
        self.show_progress_bar, self.batch_size, self.name = show_progress_bar, batch_size, name
        self.score_function_names = ["cosine"]
        self.primary_metric = f"{name + '_' if name else ''}cosine_ndcg@{max(self.ndcg_at_k)}"
        self.max_k = max(self.mrr_at_k + self.ndcg_at_k + self.accuracy_at_k + self.precision_recall_at_k + self.map_at_k)
        row_of = {cid: i for i, cid in enumerate(self.corpus_ids)}
        self._relevant_rows = [np.fromiter((row_of[d] for d in relevant_docs[qid] if d in row_of), dtype=np.int64) for qid in self.queries_ids]
        self._n_relevant = np.array([len(relevant_docs[qid]) for qid in self.queries_ids], dtype=np.float64)
        # the same metric list in the kernel's (kind, k) form, in the order compute_metrics_from_ids emits its keys
        self._metric_specs = (
            [(f"accuracy@{k}", ops.METRIC_ACCURACY, k) for k in self.accuracy_at_k]
            + [(f"precision@{k}", ops.METRIC_PRECISION, k) for k in self.precision_recall_at_k]
            + [(f"recall@{k}", ops.METRIC_RECALL, k) for k in self.precision_recall_at_k]
            + [(f"mrr@{k}", ops.METRIC_MRR, k) for k in self.mrr_at_k]
            + [(f"ndcg@{k}", ops.METRIC_NDCG, k) for k in self.ndcg_at_k]
            + [(f"map@{k}", ops.METRIC_MAP, k) for k in self.map_at_k]
        )
        self._tables: dict = {}  # device -> RelevanceTable

    def __call__(self, model, output_path: str | None = None, epoch: int = -1, steps: int = -1, *args, **kwargs) -> dict[str, float]:
        query_emb = _encode(model, self.queries, self.batch_size, self.show_progress_bar)
        corpus_emb = _encode(model, self.corpus, self.batch_size, self.show_progress_bar)
        scores = self.compute_metrics_from_embeddings(query_emb, corpus_emb)
        prefix = f"{self.name}_" if self.name else ""
        out = {f"{prefix}cosine_{k}": v for k, v in scores.items()}
        logger.info("IR evaluation%s: %s=%.4f", f" (epoch {epoch})" if epoch != -1 else "", self.primary_metric, out[self.primary_metric])
        return out

    def compute_metrics_from_embeddings(self, query_emb, corpus_emb) -> dict[str, float]:
        """Fused top-k, then the metric kernel over the device-resident id matrix; one small device->host read."""
        _, ids = topk_ids_device(query_emb, corpus_emb, self.max_k)
        return self.compute_metrics_on_device(ids)

    def relevance_table(self, device) -> ops.RelevanceTable:
        device = torch.device(device)
        if device not in self._tables:
            self._tables[device] = ops.RelevanceTable(self._relevant_rows, self._n_relevant.astype(np.int32), device=device)
        return self._tables[device]

    def compute_metrics_on_device(self, ids: torch.Tensor) -> dict[str, float]:
        """ids: CUDA int64 [Q, >=1] retrieved corpus rows, best first (-1 = none), as cos_topk returns them."""
        if ids.shape[0] == 0 or not self._metric_specs:
            return {name: 0.0 for name, _, _ in self._metric_specs}
        out: dict[str, float] = {}
        for s0 in range(0, len(self._metric_specs), ops.MAX_METRICS):
            specs = self._metric_specs[s0 : s0 + ops.MAX_METRICS]
            means, _ = ops.ir_metrics(ids, self.relevance_table(ids.device), [(kind, k) for _, kind, k in specs])
            out.update({name: float(v) for (name, _, _), v in zip(specs, means.tolist())})
        return out

    def compute_metrics_from_ids(self, ids: np.ndarray) -> dict[str, float]:
        """ids: [Q, >=max_k] retrieved corpus rows, best first (-1 = none)."""
        hits = _hit_matrix(ids, self._relevant_rows, max(len(self.corpus_ids), 1)).astype(np.float64)
        Q, K = hits.shape
        nrel = self._n_relevant
        out: dict[str, float] = {}
        cum = np.cumsum(hits, axis=1)
        ranks = np.arange(1, K + 1, dtype=np.float64)

        def at(k):
            return min(k, K)

        for k in self.accuracy_at_k:
            out[f"accuracy@{k}"] = float((cum[:, at(k) - 1] > 0).mean()) if Q and K else 0.0
        for k in self.precision_recall_at_k:
            c = cum[:, at(k) - 1] if K else np.zeros(Q)
            out[f"precision@{k}"] = float((c / k).mean()) if Q else 0.0
        for k in self.precision_recall_at_k:
            c = cum[:, at(k) - 1] if K else np.zeros(Q)
            out[f"recall@{k}"] = float((c / nrel).mean()) if Q else 0.0
        for k in self.mrr_at_k:
            h = hits[:, : at(k)]
            first = np.where(h.any(axis=1), h.argmax(axis=1) + 1, np.inf)
            out[f"mrr@{k}"] = float((1.0 / first).mean()) if Q else 0.0
        disc = 1.0 / np.log2(ranks + 1)
        for k in self.ndcg_at_k:
            kk = at(k)
            dcg = (hits[:, :kk] * disc[:kk]).sum(axis=1)
            ideal_n = np.minimum(nrel, k).astype(np.int64)
            cdisc = np.concatenate([[0.0], np.cumsum(1.0 / np.log2(np.arange(1, k + 1) + 1))])
            idcg = cdisc[ideal_n]
            out[f"ndcg@{k}"] = float((dcg / idcg).mean()) if Q else 0.0
        for k in self.map_at_k:
            kk = at(k)
            ap = (hits[:, :kk] * cum[:, :kk] / ranks[:kk]).sum(axis=1) / np.minimum(k, nrel)
            out[f"map@{k}"] = float(ap.mean()) if Q else 0.0
        return out


# --------------------------------------------------------------------------------------------------
# baselines / compare consumers
# --------------------------------------------------------------------------------------------------


def rank_all(query_embeddings, corpus_embeddings, query_ids: list[str], product_ids: list[str], limit: int = 100) -> dict[str, list[str]]:
    """query_id -> product ids ranked by cosine similarity, best first, truncated to `limit`.

    The reference builds the full N-long list per query (content_based.py:60-63); every metric it
    feeds reads at most the first 100 entries (metrics.py:150-165), so `limit` defaults to 100.
    """
    _, ids = topk_ids(query_embeddings, corpus_embeddings, limit)
    return {qid: [product_ids[j] for j in ids[i] if j >= 0] for i, qid in enumerate(query_ids)}


BASELINE_METRICS = (
    ("accuracy_at_1", ops.METRIC_ACCURACY, 1), ("accuracy_at_3", ops.METRIC_ACCURACY, 3), ("accuracy_at_5", ops.METRIC_ACCURACY, 5),
    ("accuracy_at_10", ops.METRIC_ACCURACY, 10), ("recall_at_10", ops.METRIC_RECALL, 10), ("mrr_at_10", ops.METRIC_MRR, 10),
    ("ndcg_at_10", ops.METRIC_NDCG_RETRIEVED, 10), ("map_at_100", ops.METRIC_MAP_RETRIEVED, 100),
)


def evaluate_rankings(query_embeddings, corpus_embeddings, query_ids: list[str], product_ids: list[str],
                      relevant_docs: dict[str, set[str]], limit: int = 100) -> dict[str, float]:
    """``compute_ir_metrics(rank_all(...), relevant_docs)`` without leaving the device.

    Same numbers as ranking every query (content_based.py:38-64) and scoring the rankings with
    src/baselines/metrics.py:122-176, but the [Q, 100] ids feed the metric kernel directly instead of
    becoming Q Python lists of product-id strings. Queries without relevant docs are skipped (metrics.py:137).
    """
    keep = [i for i, q in enumerate(query_ids) if q in relevant_docs and relevant_docs[q]]
    if not keep:
        return {name: 0.0 for name, _, _ in BASELINE_METRICS}
    row_of = {pid: i for i, pid in enumerate(product_ids)}
    q = to_device_matrix(query_embeddings)
    if len(keep) != len(query_ids):
        q = q[torch.as_tensor(keep, device=q.device)]
    _, ids = topk_ids_device(q, corpus_embeddings, limit)
    rel = [relevant_docs[query_ids[i]] for i in keep]
    table = ops.RelevanceTable([[row_of[d] for d in r if d in row_of] for r in rel], [len(r) for r in rel], device=ids.device)
    means, _ = ops.ir_metrics(ids, table, [(kind, k) for _, kind, k in BASELINE_METRICS])
    return {name: float(v) for (name, _, _), v in zip(BASELINE_METRICS, means.tolist())}


def compute_ir_metrics(query_rankings: dict[str, list[str]], relevant_docs: dict[str, set[str]]) -> dict[str, float]:
    """Accuracy@1/3/5/10, Recall@10, MRR@10, NDCG@10, MAP@100 as defined in src/baselines/metrics.py.

    (NDCG here normalises by the ideal ordering of the *retrieved* top-10 relevances, metrics.py:112-119.)
    """
    names = ("accuracy_at_1", "accuracy_at_3", "accuracy_at_5", "accuracy_at_10", "recall_at_10", "mrr_at_10", "ndcg_at_10", "map_at_100")
    qids = [q for q in query_rankings if q in relevant_docs and relevant_docs[q]]
    if not qids:
        return dict.fromkeys(names, 0.0)
    K = 100
    hits = np.zeros((len(qids), K))
    length = np.zeros(len(qids))
    for i, q in enumerate(qids):
        rel = relevant_docs[q]
        top = query_rankings[q][:K]
        length[i] = len(top)
        hits[i, : len(top)] = [pid in rel for pid in top]
    nrel = np.array([len(relevant_docs[q]) for q in qids], dtype=np.float64)
    h10 = hits[:, :10]
    cum10 = np.cumsum(h10, axis=1)
    disc = 1.0 / np.log2(np.arange(2, 12))
    dcg = (h10 * disc).sum(axis=1)
    idcg = (-np.sort(-h10, axis=1) * disc).sum(axis=1)
    first = np.where(h10.any(axis=1), h10.argmax(axis=1) + 1, np.inf)
    cum = np.cumsum(hits, axis=1)
    denom = np.minimum(nrel, length)
    ap = np.divide((hits * cum / np.arange(1, K + 1)).sum(axis=1), denom, out=np.zeros(len(qids)), where=denom > 0)
    vals = (
        (cum10[:, 0] > 0).mean(), (cum10[:, 2] > 0).mean(), (cum10[:, 4] > 0).mean(), (cum10[:, 9] > 0).mean(),
        (cum10[:, 9] / nrel).mean(), (1.0 / first).mean(),
        np.divide(dcg, idcg, out=np.zeros(len(qids)), where=idcg > 0).mean(), ap.mean(),
    )
    return {n: float(v) for n, v in zip(names, vals)}
