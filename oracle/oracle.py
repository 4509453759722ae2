"""CPU oracle for the retrieval hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module. The product package
(``instacart_next_order_recommendation_b200``) never does: it fails loudly when the
CUDA library is missing.

PARITY PINNING STATUS: **partially pinned**.
  * The arithmetic of ``cos_sim`` / ``MultipleNegativesRankingLoss`` /
    ``InformationRetrievalEvaluator`` lives in the un-vendored dependency
    ``sentence-transformers==5.2.2`` (reference ``uv.lock:3698-3699``), which is
    neither under /root/reference nor installed in this image. Those three functions
    are restated here from the published 5.x algorithm ("parity unpinned" for them):
    the reference's own tests hold no golden vector for this path
    (reference ``tests/conftest.py:21-50`` mocks the recommender).
  * Everything the reference implements ITSELF on the path is pinned against the
    reference's own code executed in the build container
    (``oracle/gen_golden.py`` -> ``tests/golden/*.npz|json``):
    ``Recommender.recommend`` tail (``src/inference/serve_recommendations.py:206-225``),
    ``ContentBasedBaseline.rank_all`` (``src/baselines/content_based.py:38-64``) and
    ``compute_ir_metrics`` (``src/baselines/metrics.py:122-176``).

All math is torch-CPU / numpy fp32 (fp64 where a function says so).
"""

from __future__ import annotations

import heapq
import math
from typing import Iterable, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# a1: sentence_transformers.util.cos_sim  (called at serve_recommendations.py:214,250;
#     content_based.py:54; compare_untrained_vs_trained.py:74)
# --------------------------------------------------------------------------------------


def _to_tensor(x) -> torch.Tensor:
    """ST's util converts non-tensors with torch.tensor() and 1-D inputs to [1, D]."""
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(x)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    return x


def cos_sim(a, b) -> torch.Tensor:
    """cos_sim(a, b)[i, j] = <a_i/max(|a_i|,eps), b_j/max(|b_j|,eps)>, eps=1e-12.

    Restates sentence_transformers.util.cos_sim (5.2.2): convert -> unsqueeze 1-D ->
    F.normalize(p=2, dim=1) on both -> torch.mm(a_n, b_n.T).
    """
    a = _to_tensor(a)
    b = _to_tensor(b)
    a_n = F.normalize(a, p=2, dim=1)
    b_n = F.normalize(b, p=2, dim=1)
    return torch.mm(a_n, b_n.transpose(0, 1))


def cos_sim_rounded(a, b, dtype: torch.dtype) -> torch.Tensor:
    """The low-precision oracle of SURVEY §8(c): fp32 math on inputs rounded to `dtype`."""
    a = _to_tensor(a).to(dtype).to(torch.float32)
    b = _to_tensor(b).to(dtype).to(torch.float32)
    return cos_sim(a, b)


def cos_sim_f64(a, b) -> torch.Tensor:
    """Exact-ish yardstick (fp64) used to state how far the fp32 oracle itself is off."""
    return cos_sim(_to_tensor(a).double(), _to_tensor(b).double())


# --------------------------------------------------------------------------------------
# new fused entry point == torch.topk(cos_sim(q, c), k, dim=1)  (a4's scoring step)
# --------------------------------------------------------------------------------------


def cos_topk(queries, catalog, k: int, sorted: bool = True, chunk: int = 1024):
    """(values f32 [Q,k], indices int64 [Q,k]) of torch.topk(cos_sim(q, c), k, dim=1).

    Chunked over queries only to bound memory; per-row results are those of the
    unchunked expression.
    """
    q = _to_tensor(queries)
    c = _to_tensor(catalog)
    k = min(k, c.shape[0])
    c_n = F.normalize(c, p=2, dim=1)
    vals, idxs = [], []
    for s in range(0, q.shape[0], chunk):
        q_n = F.normalize(q[s : s + chunk], p=2, dim=1)
        sc = torch.mm(q_n, c_n.transpose(0, 1))
        v, i = torch.topk(sc, k, dim=1, largest=True, sorted=sorted)
        vals.append(v)
        idxs.append(i)
    return torch.cat(vals), torch.cat(idxs)


def cos_topk_catalog_chunked(queries, catalog, k: int, chunk_rows: int = 1 << 18):
    """cos_topk for catalogs too large to score at once (BASELINE configs 4-5 at scale): the same expression,
    F.normalize -> mm -> topk, evaluated over row blocks of the catalog with the per-block top-k merged by one
    more topk. `catalog` may be bf16 (the low-precision oracle of SURVEY §8c: fp32 math on the rounded inputs);
    each block is widened to fp32 before anything is computed. Ties inside a block keep torch.topk's order; the
    callers compare ids tie-tolerantly.
    """
    q = F.normalize(_to_tensor(queries).to(torch.float32), p=2, dim=1)
    n = catalog.shape[0]
    k = min(k, n)
    best_v = torch.full((q.shape[0], 0), float("-inf"))
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64)
    for s in range(0, n, chunk_rows):
        c_n = F.normalize(catalog[s : s + chunk_rows].to(torch.float32), p=2, dim=1)
        sc = torch.mm(q, c_n.transpose(0, 1))
        v, i = torch.topk(sc, min(k, sc.shape[1]), dim=1)
        best_v = torch.cat([best_v, v], dim=1)
        best_i = torch.cat([best_i, i + s], dim=1)
        if best_v.shape[1] > k:
            best_v, pos = torch.topk(best_v, k, dim=1)
            best_i = torch.gather(best_i, 1, pos)
    order = torch.argsort(best_v, dim=1, descending=True, stable=True)
    return torch.gather(best_v, 1, order), torch.gather(best_i, 1, order)


# --------------------------------------------------------------------------------------
# a2: /recommend tail — serve_recommendations.py:213-225 (and :250-262)
# --------------------------------------------------------------------------------------


def recommend_tail(
    query_emb,
    product_embeddings,
    product_ids: Sequence[str],
    top_k: int = 10,
    exclude_product_ids: set[str] | None = None,
) -> list[tuple[str, float]]:
    """scores = cos_sim(q, E)[0]; full argsort(descending); walk, skipping excluded."""
    scores = cos_sim(query_emb, product_embeddings)[0]
    order = scores.argsort(descending=True)
    excluded = exclude_product_ids or set()
    out: list[tuple[str, float]] = []
    for idx in order:
        pid = product_ids[idx]
        if pid in excluded:
            continue
        out.append((pid, float(scores[idx])))
        if len(out) >= top_k:
            break
    return out


# --------------------------------------------------------------------------------------
# a3: batched full ranking — content_based.py:54-63, compare_untrained_vs_trained.py:74-84
# --------------------------------------------------------------------------------------


def rank_all(query_emb, corpus_emb, query_ids: Sequence[str], product_ids: Sequence[str], limit: int | None = None):
    """sim = cos_sim(Q, C).numpy(); per row np.argsort(row)[::-1] -> id lists.

    `limit` truncates each list (the only consumer, metrics.py:150-165, reads <=100).
    """
    sim = cos_sim(query_emb, corpus_emb).cpu().numpy()
    out: dict[str, list[str]] = {}
    for i, qid in enumerate(query_ids):
        order = np.argsort(np.asarray(sim[i]).flatten())[::-1]
        if limit is not None:
            order = order[:limit]
        out[qid] = [product_ids[j] for j in order]
    return out


# --------------------------------------------------------------------------------------
# a4/a6: InformationRetrievalEvaluator core (ST 5.2.2, restated): corpus chunks of 50,000,
#        topk(min(max_k, chunk), sorted=False) per chunk, per-query min-heap of size max_k.
# --------------------------------------------------------------------------------------


def ir_eval_topk(query_emb, corpus_emb, max_k: int = 100, corpus_chunk_size: int = 50_000, query_chunk: int = 1024):
    """Returns per-query list of (score, corpus_index), sorted by score descending."""
    q = _to_tensor(query_emb)
    c = _to_tensor(corpus_emb)
    heaps: list[list[tuple[float, int]]] = [[] for _ in range(q.shape[0])]
    for cs in range(0, c.shape[0], corpus_chunk_size):
        chunk = c[cs : cs + corpus_chunk_size]
        for qs in range(0, q.shape[0], query_chunk):
            ps = cos_sim(q[qs : qs + query_chunk], chunk)
            v, i = torch.topk(ps, min(max_k, ps.shape[1]), dim=1, largest=True, sorted=False)
            v = v.tolist()
            i = i.tolist()
            for r in range(len(v)):
                h = heaps[qs + r]
                for score, ci in zip(v[r], i[r]):
                    item = (score, cs + ci)
                    if len(h) < max_k:
                        heapq.heappush(h, item)
                    else:
                        heapq.heappushpop(h, item)
    return [sorted(h, key=lambda t: t[0], reverse=True) for h in heaps]


# --------------------------------------------------------------------------------------
# IR metric arithmetic — definitions of src/baselines/metrics.py:13-176 (the in-tree ones;
# note NDCG normalises by the IDCG of the *retrieved* relevances, metrics.py:112-119).
# Written independently (array form) and pinned against the reference's own module by
# tests/golden/metrics_golden.json.
# --------------------------------------------------------------------------------------


def _hits(relevant: set[str], ranked: Sequence[str], k: int) -> np.ndarray:
    return np.fromiter((1.0 if pid in relevant else 0.0 for pid in ranked[:k]), dtype=np.float64, count=min(k, len(ranked)))


def ir_metrics(query_rankings: dict[str, list[str]], relevant_docs: dict[str, set[str]]) -> dict[str, float]:
    keys = ("accuracy_at_1", "accuracy_at_3", "accuracy_at_5", "accuracy_at_10", "recall_at_10", "mrr_at_10", "ndcg_at_10", "map_at_100")
    qids = [q for q in query_rankings if q in relevant_docs and relevant_docs[q]]
    if not qids:
        return {k: 0.0 for k in keys}
    acc = {1: 0, 3: 0, 5: 0, 10: 0}
    recall = mrr = ndcg = ap = 0.0
    for q in qids:
        rel = relevant_docs[q]
        ranked = query_rankings[q]
        h100 = _hits(rel, ranked, 100)
        h10 = h100[:10]
        for k in acc:
            acc[k] += 1 if h10[:k].any() else 0
        recall += h10.sum() / len(rel)
        nz = np.flatnonzero(h10)
        mrr += 1.0 / (nz[0] + 1) if nz.size else 0.0
        disc = 1.0 / np.log2(np.arange(2, 2 + h10.size))
        dcg = float((h10 * disc).sum())
        idcg = float((np.sort(h10)[::-1] * disc).sum())
        ndcg += dcg / idcg if idcg > 0 else 0.0
        if h100.size:
            prec = np.cumsum(h100) / np.arange(1, h100.size + 1)
            ap += float((prec * h100).sum()) / min(len(rel), h100.size)
    n = len(qids)
    return {
        "accuracy_at_1": acc[1] / n,
        "accuracy_at_3": acc[3] / n,
        "accuracy_at_5": acc[5] / n,
        "accuracy_at_10": acc[10] / n,
        "recall_at_10": recall / n,
        "mrr_at_10": mrr / n,
        "ndcg_at_10": ndcg / n,
        "map_at_100": ap / n,
    }


# --------------------------------------------------------------------------------------
# a6: InformationRetrievalEvaluator.compute_metrics (ST 5.2.2, restated from the published
#     algorithm — "parity unpinned", see the header; evaluator built at train_sbert.py:197-202).
#     Loop form, per query in order: accuracy@k (any hit in the first k), precision@k = hits/k,
#     recall@k = hits/|relevant|, MRR@k = 1/rank of the first hit, NDCG@k = DCG/DCG([1]*|relevant|)
#     with discount 1/log2(rank+1), MAP@k = sum of precision at each hit / min(k, |relevant|).
#     `ranked` holds catalog rows (ints), best first; entries < 0 mean "no result".
# --------------------------------------------------------------------------------------


def st_ir_metrics(ranked: Sequence[Sequence[int]], relevant: Sequence[set], n_relevant: Sequence[int] | None = None,
                  accuracy_at_k=(1, 3, 5, 10), precision_recall_at_k=(1, 3, 5, 10), mrr_at_k=(10,), ndcg_at_k=(10,), map_at_k=(100,),
                  per_query: bool = False):
    n = len(ranked)
    nrel = [len(r) for r in relevant] if n_relevant is None else list(n_relevant)
    names = ([f"accuracy@{k}" for k in accuracy_at_k] + [f"precision@{k}" for k in precision_recall_at_k]
             + [f"recall@{k}" for k in precision_recall_at_k] + [f"mrr@{k}" for k in mrr_at_k] + [f"ndcg@{k}" for k in ndcg_at_k]
             + [f"map@{k}" for k in map_at_k])
    rows = []
    for rk, rel, nr in zip(ranked, relevant, nrel):
        rk = [r for r in rk if r >= 0]
        vals = {}
        for k in accuracy_at_k:
            vals[f"accuracy@{k}"] = 1.0 if any(r in rel for r in rk[:k]) else 0.0
        for k in precision_recall_at_k:
            c = sum(1 for r in rk[:k] if r in rel)
            vals[f"precision@{k}"] = c / k
            vals[f"recall@{k}"] = c / nr
        for k in mrr_at_k:
            vals[f"mrr@{k}"] = 0.0
            for j, r in enumerate(rk[:k]):
                if r in rel:
                    vals[f"mrr@{k}"] = 1.0 / (j + 1)
                    break
        for k in ndcg_at_k:
            dcg = 0.0
            for j, r in enumerate(rk[:k]):
                if r in rel:
                    dcg += 1.0 / math.log2(j + 2)
            idcg = 0.0
            for j in range(min(nr, k)):
                idcg += 1.0 / math.log2(j + 2)
            vals[f"ndcg@{k}"] = dcg / idcg
        for k in map_at_k:
            hits, sp = 0, 0.0
            for j, r in enumerate(rk[:k]):
                if r in rel:
                    hits += 1
                    sp += hits / (j + 1)
            vals[f"map@{k}"] = sp / min(k, nr)
        rows.append([vals[m] for m in names])
    arr = np.asarray(rows, dtype=np.float64).reshape(n, len(names))
    means = {m: (float(arr[:, i].sum() / n) if n else 0.0) for i, m in enumerate(names)}
    return (means, arr) if per_query else means


# --------------------------------------------------------------------------------------
# a5: MultipleNegativesRankingLoss.forward (ST 5.2.2, restated; built at train_sbert.py:182-185)
#     scores = cos_sim(anchors, candidates) * scale ; CrossEntropyLoss()(scores, arange(B))
# --------------------------------------------------------------------------------------


def mnrl_loss(anchors: torch.Tensor, candidates: torch.Tensor, scale: float = 20.0) -> torch.Tensor:
    scores = cos_sim(anchors, candidates) * scale
    labels = torch.arange(scores.shape[0], dtype=torch.long, device=scores.device)
    return F.cross_entropy(scores, labels)


def mnrl_loss_and_grads(anchors, candidates, scale: float = 20.0, dtype=torch.float32):
    """loss, dL/danchors, dL/dcandidates by torch.autograd, math in `dtype`."""
    a = anchors.detach().to(dtype).clone().requires_grad_(True)
    p = candidates.detach().to(dtype).clone().requires_grad_(True)
    loss = mnrl_loss(a, p, scale)
    loss.backward()
    return loss.detach(), a.grad.detach(), p.grad.detach()


def mnrl_rect_loss_and_grads(anchors, candidates, scale: float, label_offset: int):
    """B anchors vs Bc candidates, label of anchor i = i + label_offset (ST 5.x MNRL with gathered candidates, restated):
    loss = cross_entropy(cos_sim(A, C) * scale, arange(B) + label_offset); grads by autograd."""
    with torch.enable_grad():  # also callable from inside an autograd.Function (the gloo test seam)
        a = anchors.detach().float().clone().requires_grad_(True)
        c = candidates.detach().float().clone().requires_grad_(True)
        scores = cos_sim(a, c) * scale
        labels = torch.arange(a.shape[0], dtype=torch.long) + label_offset
        loss = F.cross_entropy(scores, labels)
        loss.backward()
    return loss.detach(), a.grad.detach(), c.grad.detach()


def mnrl_gathered_reference(anchors_per_rank, positives_per_rank, scale: float):
    """What G ranks running MNRL(gather_across_devices=True) compute: per rank its loss and anchor gradient, and for its
    positives the SUM over ranks of d loss_r / d P_rank (all_gather with autograd = reduce-scatter in backward)."""
    G = len(anchors_per_rank)
    B = anchors_per_rank[0].shape[0]
    cand = torch.cat([p.float() for p in positives_per_rank])
    losses, grads_a, grad_c_sum = [], [], torch.zeros_like(cand)
    for r in range(G):
        loss, ga, gc = mnrl_rect_loss_and_grads(anchors_per_rank[r], cand, scale, r * B)
        losses.append(loss)
        grads_a.append(ga)
        grad_c_sum += gc
    return losses, grads_a, [grad_c_sum[r * B : (r + 1) * B] for r in range(G)]


# --------------------------------------------------------------------------------------
# Seeded synthetic generators of SURVEY §8(d) (never Instacart data)
# --------------------------------------------------------------------------------------


def synth_isotropic(n: int, d: int, seed: int, normalize: bool = True) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g, dtype=torch.float32)
    return F.normalize(x, dim=1) if normalize else x


def synth_clustered(n: int, d: int, seed: int, n_centres: int = 134, noise: float = 1.0):
    """item = normalize(centre[c] + noise * randn): Instacart-like (134 aisles)."""
    g = torch.Generator().manual_seed(seed)
    centres = F.normalize(torch.randn(n_centres, d, generator=g), dim=1) * math.sqrt(d) * 0.25
    assign = torch.randint(0, n_centres, (n,), generator=g)
    x = centres[assign] + noise * torch.randn(n, d, generator=g)
    return F.normalize(x, dim=1), assign


def synth_queries_from_items(items: torch.Tensor, q: int, seed: int, noise: float = 0.05):
    """query = normalize(item[j] + noise*sqrt(1/d)-scaled randn); returns (queries, source rows)."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, items.shape[0], (q,), generator=g)
    x = items[src] + noise * torch.randn(q, items.shape[1], generator=g) / math.sqrt(items.shape[1]) * 4.0
    return F.normalize(x, dim=1), src


def synth_unnormalised(n: int, d: int, seed: int) -> torch.Tensor:
    """randn rows times a log-normal per-row scale: exercises the normalisation path."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    s = torch.exp(torch.randn(n, 1, generator=g))
    return x * s


# --------------------------------------------------------------------------------------
# Tie-tolerant comparison helpers shared by the parity tests
# --------------------------------------------------------------------------------------


def compare_topk(values, indices, ref_values, ref_indices, rtol: float, atol: float = 0.0, floor: float = 0.05):
    """Returns (max_rel_err_of_scores, n_id_mismatch_outside_ties).

    Scores are compared rank by rank, relative to max(|score|, floor): a cosine near 0 has unbounded
    relative error even when its absolute error is 1e-8, so the relative tolerance is
    floored at |score| = `floor`. Ids are compared only at ranks whose gap to both neighbouring
    reference scores exceeds the tolerance; elsewhere a swap is a legal tie.
    """
    v = np.asarray(values, dtype=np.float64)
    r = np.asarray(ref_values, dtype=np.float64)
    i = np.asarray(indices)
    ri = np.asarray(ref_indices)
    denom = np.maximum(np.abs(r), floor)
    rel = np.abs(v - r) / denom
    tol = rtol * denom + atol
    gap_prev = np.full_like(r, np.inf)
    gap_next = np.full_like(r, np.inf)
    gap_prev[:, 1:] = np.abs(r[:, 1:] - r[:, :-1])
    gap_next[:, :-1] = np.abs(r[:, :-1] - r[:, 1:])
    decisive = (gap_prev > 2 * tol) & (gap_next > 2 * tol)
    # the last rank competes with the (unseen) k+1-th score: never decisive
    decisive[:, -1] = False
    mism = int(((i != ri) & decisive).sum())
    return float(rel.max()) if rel.size else 0.0, mism
