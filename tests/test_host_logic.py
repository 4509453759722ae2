"""CPU-side checks: the C-ABI library loads and exports the header's symbols, the host mirrors of the
reference interface behave like the reference's own code (golden fixtures), and nothing falls back to CPU."""

import json
import re
import shutil
from pathlib import Path

import numpy as np
import pytest
import torch

import instacart_next_order_recommendation_b200 as icr
from instacart_next_order_recommendation_b200 import _lib, build, evaluation, ops, sharded
from instacart_next_order_recommendation_b200.index import EmbeddingIndex

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "icr_b200.h").read_text()
    declared = set(re.findall(r"\b(icr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.icr_abi_version() == _lib.ABI_VERSION == 3
    assert lib.icr_planes_row_elems(384) == 768 and lib.icr_planes_row_elems(100) == 256
    assert lib.icr_screen_plane_row_elems(384) == 384 and lib.icr_screen_plane_row_elems(100) == 128


def test_argument_errors_are_reported_without_a_gpu(lib):
    # shape / dtype / alignment validation happens before any CUDA call
    rc = lib.icr_cos_topk(None, 1, 384, None, 10, 384, 384, 0, None, None, None, 10, 0, 0, None, None, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.icr_last_error_string()
    rc = lib.icr_cos_topk(16, 1, 384, 16, 10, 384, 384, 7, None, None, None, 10, 0, 0, None, None, None, 0, None)
    assert rc == -2
    rc = lib.icr_cos_topk(16, 1, 384, 16, 10, 384, 384, 0, None, None, None, 1000, 0, 0, None, None, None, 0, None)
    assert rc == -4
    rc = lib.icr_cos_topk(8, 1, 384, 16, 10, 384, 384, 0, None, None, None, 10, 0, 0, None, None, None, 0, None)
    assert rc == -3
    with pytest.raises(ValueError, match="ICR_ERR_K"):
        _lib.check(-4)
    assert lib.icr_cos_topk_workspace_bytes(1, 49688, 384, 0, 10, 1, 0) > 0


def test_no_cpu_fallback():
    q = torch.randn(2, 8)
    c = torch.randn(5, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.cos_topk(q, c, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.cos_sim_dense(q, c)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            icr.cos_sim(q, c)
        with pytest.raises(RuntimeError, match="CUDA"):
            icr.mnrl_loss(q, q)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.NativeLibraryMissing, match="no CPU"):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    for p in (ROOT / "instacart_next_order_recommendation_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p


def test_embedding_index_is_byte_compatible_with_the_reference(golden_dir, tmp_path):
    meta = json.loads((golden_dir / "embedding_index_meta.json").read_text())
    gold = golden_dir / "embedding_index"
    ids = json.loads((gold / "product_ids.json").read_text())
    emb = np.load(gold / "embeddings.npy")
    # same hashing of (model_dir | corpus_path) -> directory name
    idx = EmbeddingIndex(Path(meta["corpus_path"]), "fake-model")
    assert idx.directory.name == meta["dir_name"] and idx.directory.parent.name == meta["index_subdir"]
    # write with our code, compare bytes with what the reference wrote
    corpus = tmp_path / "eval_corpus.json"
    corpus.write_text("{}")
    import os

    os.utime(corpus, (1_700_000_000, 1_700_000_000))
    ours = EmbeddingIndex(corpus, "fake-model")
    ours.save(ids, emb.astype(np.float64))  # the reference casts to float32 on save
    assert (ours.directory / "embeddings.npy").read_bytes() == (gold / "embeddings.npy").read_bytes()
    assert (ours.directory / "product_ids.json").read_bytes() == (gold / "product_ids.json").read_bytes()
    m = json.loads((ours.directory / "manifest.json").read_text())
    assert list(m.keys()) == meta["manifest_keys"]
    assert m["corpus_mtime"] == meta["manifest"]["corpus_mtime"] and m["n_products"] == meta["manifest"]["n_products"]
    # an index written by the reference loads through our loader (after re-homing the manifest paths)
    shutil.rmtree(ours.directory)
    shutil.copytree(gold, ours.directory)
    man = json.loads((ours.directory / "manifest.json").read_text())
    man["corpus_path"] = str(corpus.resolve())
    (ours.directory / "manifest.json").write_text(json.dumps(man, indent=2))
    np.testing.assert_array_equal(ours.load(ids), emb)
    # invalidation rules of the reference: id list, mtime, model dir
    assert ours.load(ids[:-1]) is None
    os.utime(corpus, (1_700_000_001, 1_700_000_001))
    assert ours.load(ids) is None
    assert EmbeddingIndex(corpus, "other-model").load(ids) is None


def test_index_memmap_and_bf16_sidecar_rules(tmp_path):
    from instacart_next_order_recommendation_b200.index import chunk_spans

    corpus = tmp_path / "eval_corpus.json"
    corpus.write_text("{}")
    ids = [str(i) for i in range(300)]
    emb = np.random.default_rng(0).standard_normal((300, 24)).astype(np.float32)
    idx = EmbeddingIndex(corpus, "fake-model")
    assert idx.open(ids) is None and idx.open_bf16_sidecar(ids) is None  # nothing on disk yet
    idx.save(ids, emb)
    mm = idx.open(ids)
    assert isinstance(mm, np.memmap) and not mm.flags.writeable
    np.testing.assert_array_equal(mm, idx.load(ids))
    assert idx.open(ids[:-1]) is None  # same invalidation rules as load()
    assert idx.open_bf16_sidecar(ids) is None  # no sidecar yet
    idx.save_bf16_sidecar()
    side = idx.open_bf16_sidecar(ids)
    want = torch.from_numpy(emb).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    np.testing.assert_array_equal(side, want)
    # a sidecar of the wrong size, or older than embeddings.npy, is ignored; so is any sidecar of a stale index
    import os

    p = idx.directory / "embeddings.bf16.bin"
    p.write_bytes(p.read_bytes()[:-2])
    assert idx.open_bf16_sidecar(ids) is None
    idx.save_bf16_sidecar(normalize=True)
    side_n = idx.open_bf16_sidecar(ids)
    want_n = torch.nn.functional.normalize(torch.from_numpy(emb), dim=1).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    np.testing.assert_array_equal(side_n, want_n)
    old = (idx.directory / "embeddings.npy").stat().st_mtime - 100
    os.utime(p, (old, old))
    assert idx.open_bf16_sidecar(ids) is None
    idx.save_bf16_sidecar()
    os.utime(corpus, (1_700_000_123, 1_700_000_123))
    assert idx.open(ids) is None and idx.open_bf16_sidecar(ids) is None
    assert chunk_spans(3, 10, 4) == [(3, 7), (7, 10)] and chunk_spans(5, 5, 4) == [] and chunk_spans(0, 4, 4) == [(0, 4)]


def test_compute_ir_metrics_matches_reference_code(golden_dir):
    gold = json.loads((golden_dir / "metrics_golden.json").read_text())
    rel = {k: set(v) for k, v in gold["relevant"].items()}
    m = evaluation.compute_ir_metrics(gold["rank_all_top100"], rel)
    for key, val in gold["metrics"].items():
        assert m[key] == pytest.approx(val, abs=1e-12), key
    for case in gold["extra"]:
        m = evaluation.compute_ir_metrics(case["rankings"], {k: set(v) for k, v in case["relevant"].items()})
        for key, val in case["metrics"].items():
            assert m[key] == pytest.approx(val, abs=1e-12), key


def test_evaluator_metric_arithmetic_and_keys():
    rng = np.random.default_rng(5)
    n_corpus, n_q = 500, 40
    corpus = {f"p{i}": f"text {i}" for i in range(n_corpus)}
    queries = {f"q{i}": f"query {i}" for i in range(n_q)}
    relevant = {f"q{i}": {f"p{j}" for j in rng.choice(n_corpus, size=rng.integers(1, 15), replace=False)} for i in range(n_q)}
    relevant["q7"] = set()  # dropped, like upstream
    ev = icr.InformationRetrievalEvaluator(queries, corpus, relevant, name="order-recommendation")
    assert ev.primary_metric == "order-recommendation_cosine_ndcg@10" and ev.max_k == 100
    assert len(ev.queries_ids) == n_q - 1
    ids = np.stack([rng.permutation(n_corpus)[:100] for _ in ev.queries_ids])
    # plant some hits near the top
    for r, qid in enumerate(ev.queries_ids[:20]):
        ids[r, rng.integers(0, 10)] = int(next(iter(relevant[qid]))[1:])
    got = ev.compute_metrics_from_ids(ids)
    rel_rows = [{int(d[1:]) for d in relevant[q]} for q in ev.queries_ids]
    from oracle import oracle

    want = oracle.st_ir_metrics([list(r) for r in ids], rel_rows)
    assert set(got) == set(want)
    for k in want:
        assert got[k] == pytest.approx(want[k], abs=1e-12), k


def test_shard_bounds_and_packing():
    for n, g in ((49688, 8), (10, 4), (3, 8), (0, 2), (100, 1)):
        blocks = [sharded.shard_bounds(n, g, r) for r in range(g)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert all(lo <= hi for lo, hi in blocks)
    v = torch.tensor([[0.5, -0.25, float("-inf")]])
    i = torch.tensor([[7, 2**33, -1]])
    buf = sharded.pack_candidates(v, i)
    s2, i2 = sharded.unpack_candidates(buf.unsqueeze(0))
    assert torch.equal(s2[0], v) and torch.equal(i2[0], i)
