"""Generate tests/golden/* by EXECUTING THE REFERENCE'S OWN CODE in the build container.

TEST INFRASTRUCTURE. Run once here (``python oracle/gen_golden.py``); the outputs are
committed because /root/reference does not exist on the GPU box.

What is executed from /root/reference, unmodified:
  * ``src.baselines.metrics.compute_ir_metrics``                (metrics.py:122-176)
  * ``src.inference.serve_recommendations.Recommender.recommend`` and
    ``MonitoredRecommender.recommend`` (:206-225, :236-279), ``EmbeddingIndex`` (:66-130)
  * ``src.baselines.content_based.ContentBasedBaseline.rank_all`` (content_based.py:38-64)

What is NOT the reference's code: the ``sentence_transformers`` package is absent from this
image (pinned 5.2.2, uv.lock:3698-3699), so a stub module is injected whose ``cos_sim`` is
the restatement in ``oracle/oracle.py`` and whose ``SentenceTransformer.encode`` returns
seeded vectors. The goldens therefore pin the reference's own tail/ranking/metric/index
logic, and record — not independently verify — the cos_sim arithmetic.
"""

from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("ICR_REFERENCE", "/root/reference"))
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))

from oracle import oracle  # noqa: E402

N, D, Q = 1500, 48, 12


class _FakeEncoder:
    """Stands in for SentenceTransformer: text 'c:<i>' / 'q:<i>' -> row i of a seeded table."""

    tables: dict[str, np.ndarray] = {}

    def __init__(self, *a, **k):
        pass

    def encode(self, texts, batch_size=64, show_progress_bar=False, normalize_embeddings=True, **kw):
        rows = []
        for t in texts:
            kind, i = t.split(":")
            rows.append(self.tables[kind][int(i)])
        return np.stack(rows).astype(np.float32)


def _install_stub():
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = _FakeEncoder
    util = types.ModuleType("sentence_transformers.util")
    util.cos_sim = oracle.cos_sim
    st.util = util
    sys.modules["sentence_transformers"] = st
    sys.modules["sentence_transformers.util"] = util


def main():
    assert REF.exists(), f"{REF} missing: goldens can only be regenerated in the build container"
    OUT.mkdir(parents=True, exist_ok=True)
    _install_stub()
    sys.path.insert(0, str(REF))
    from src.baselines import metrics as ref_metrics
    from src.baselines.content_based import ContentBasedBaseline
    from src.inference.serve_recommendations import EmbeddingIndex, MonitoredRecommender, Recommender

    items, _ = oracle.synth_clustered(N, D, seed=1234, n_centres=16)
    queries, src = oracle.synth_queries_from_items(items, Q, seed=4321)
    items = items.numpy()
    queries = queries.numpy()
    _FakeEncoder.tables = {"c": items, "q": queries}
    pids = [str(10_000 + 7 * i) for i in range(N)]

    tmp = Path(tempfile.mkdtemp(prefix="icr_golden_"))
    try:
        corpus_path = tmp / "eval_corpus.json"
        corpus_path.write_text(json.dumps({pid: f"c:{i}" for i, pid in enumerate(pids)}))
        os.utime(corpus_path, (1_700_000_000, 1_700_000_000))

        # ---- Recommender.recommend tail (reference code) ---------------------------------
        rec = Recommender(model_dir="fake-model", corpus_path=corpus_path, use_index=True)
        cases = []
        for qi in range(Q):
            for top_k, n_excl in ((10, 0), (5, 3), (100, 0), (100, 40), (1, 0)):
                base = rec.recommend(f"q:{qi}", top_k=top_k + n_excl)
                excl = sorted({pid for pid, _ in base[: 2 * n_excl : 2]})
                res = rec.recommend(f"q:{qi}", top_k=top_k, exclude_product_ids=set(excl))
                cases.append({"q": qi, "top_k": top_k, "exclude": excl, "result": [[p, s] for p, s in res]})
        mon = MonitoredRecommender(model_dir="fake-model", corpus_path=corpus_path, use_index=True)
        mres = mon.recommend("q:0", top_k=10, user_id="u1", exclude_product_ids={pids[int(src[0])]})
        mon_case = {
            "q": 0,
            "top_k": 10,
            "exclude": [pids[int(src[0])]],
            "result": [[p, s] for p, s in mres],
            "metrics_fields": sorted(vars(mon.last_metrics).keys()),
            "num_recommendations": mon.last_metrics.num_recommendations,
            "top_score": mon.last_metrics.top_score,
            "avg_score": mon.last_metrics.avg_score,
            "user_id": mon.last_metrics.user_id,
        }
        (OUT / "recommend_golden.json").write_text(json.dumps({"cases": cases, "monitored": mon_case}))

        # ---- EmbeddingIndex on-disk format written by the reference --------------------------
        idx = EmbeddingIndex(corpus_path, "fake-model")
        idx_dir = idx._dir
        gold_idx = OUT / "embedding_index"
        if gold_idx.exists():
            shutil.rmtree(gold_idx)
        shutil.copytree(idx_dir, gold_idx)
        manifest = json.loads((gold_idx / "manifest.json").read_text())
        (OUT / "embedding_index_meta.json").write_text(
            json.dumps(
                {
                    "dir_name": idx_dir.name,
                    "index_subdir": idx_dir.parent.name,
                    "canonical": f"fake-model|{corpus_path.resolve()}",
                    "manifest_keys": list(manifest.keys()),
                    "manifest": manifest,
                    "corpus_path": str(corpus_path.resolve()),
                }
            )
        )

        # ---- ContentBasedBaseline.rank_all (reference code) ----------------------------------
        eval_queries = {f"order{qi}": f"q:{qi}" for qi in range(Q)}
        eval_corpus = {pid: f"c:{i}" for i, pid in enumerate(pids)}
        cb = ContentBasedBaseline(eval_queries, eval_corpus, model_name="fake-model")
        rankings = cb.rank_all()
        rank_top = {qid: r[:100] for qid, r in rankings.items()}

        # ---- compute_ir_metrics (reference code, pure python) --------------------------------
        rng = np.random.default_rng(99)
        relevant = {}
        for qi, qid in enumerate(eval_queries):
            r = rankings[qid]
            picks = set()
            picks.update(r[j] for j in rng.choice(12, size=rng.integers(0, 4), replace=False))
            picks.update(r[j] for j in rng.choice(300, size=rng.integers(1, 6), replace=False))
            relevant[qid] = picks
        relevant["order3"] = set()  # empty relevance -> skipped (metrics.py:136)
        mvals = ref_metrics.compute_ir_metrics(rankings, relevant)
        # extra random-ranking cases to pin the metric arithmetic on its own
        extra = []
        for case in range(6):
            ids = [f"p{j}" for j in range(150)]
            qr, rel = {}, {}
            for qq in range(5):
                perm = list(rng.permutation(ids))
                qr[f"q{qq}"] = perm[: int(rng.integers(3, 150))]
                rel[f"q{qq}"] = set(rng.choice(ids, size=int(rng.integers(0, 12)), replace=False).tolist())
            extra.append(
                {
                    "rankings": qr,
                    "relevant": {k: sorted(v) for k, v in rel.items()},
                    "metrics": ref_metrics.compute_ir_metrics(qr, rel),
                }
            )
        extra.append({"rankings": {}, "relevant": {}, "metrics": ref_metrics.compute_ir_metrics({}, {})})
        (OUT / "metrics_golden.json").write_text(
            json.dumps(
                {
                    "rank_all_top100": rank_top,
                    "relevant": {k: sorted(v) for k, v in relevant.items()},
                    "metrics": mvals,
                    "extra": extra,
                }
            )
        )

        np.savez_compressed(
            OUT / "embeddings_small.npz",
            items=items,
            queries=queries,
            src=src.numpy(),
            product_ids=np.array(pids),
        )
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

    # ---- known-answer vectors (analytic) + restated-oracle records ---------------
    g = torch.Generator().manual_seed(2024)
    a = torch.randn(32, 24, generator=g)
    p = torch.randn(32, 24, generator=g)
    l20, ga20, gp20 = oracle.mnrl_loss_and_grads(a, p, 20.0, dtype=torch.float64)
    l30, ga30, gp30 = oracle.mnrl_loss_and_grads(a, p, 30.0, dtype=torch.float64)
    un_c = oracle.synth_unnormalised(300, 40, seed=7)
    un_q = oracle.synth_unnormalised(9, 40, seed=8)
    v, i = oracle.cos_topk(un_q.double(), un_c.double(), 20)
    np.savez_compressed(
        OUT / "oracle_records.npz",
        mnrl_a=a.numpy(),
        mnrl_p=p.numpy(),
        mnrl_loss20=l20.numpy(),
        mnrl_ga20=ga20.numpy(),
        mnrl_gp20=gp20.numpy(),
        mnrl_loss30=l30.numpy(),
        mnrl_ga30=ga30.numpy(),
        mnrl_gp30=gp30.numpy(),
        un_c=un_c.numpy(),
        un_q=un_q.numpy(),
        un_topk_vals_f64=v.numpy(),
        un_topk_idx=i.numpy(),
    )
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
