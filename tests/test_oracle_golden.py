"""The oracle against (a) fixtures produced by the reference's own code, (b) analytic known answers."""

import json
import math

import numpy as np
import pytest
import torch

from oracle import oracle


@pytest.fixture(scope="module")
def small(golden_dir):
    z = np.load(golden_dir / "embeddings_small.npz")
    return {k: z[k] for k in z.files}


def test_recommend_tail_matches_reference_code(golden_dir, small):
    gold = json.loads((golden_dir / "recommend_golden.json").read_text())
    pids = [str(p) for p in small["product_ids"]]
    assert len(gold["cases"]) == 60
    for case in gold["cases"]:
        got = oracle.recommend_tail(small["queries"][case["q"]], small["items"], pids, case["top_k"], set(case["exclude"]))
        assert [p for p, _ in got] == [p for p, _ in case["result"]]
        np.testing.assert_allclose([s for _, s in got], [s for _, s in case["result"]], rtol=1e-6)
        assert not set(case["exclude"]) & {p for p, _ in got}
    mon = gold["monitored"]
    got = oracle.recommend_tail(small["queries"][0], small["items"], pids, mon["top_k"], set(mon["exclude"]))
    assert [p for p, _ in got] == [p for p, _ in mon["result"]]
    assert mon["metrics_fields"] == sorted(
        ["user_id", "query_embedding_time_ms", "similarity_compute_time_ms", "total_latency_ms", "num_recommendations", "top_score", "avg_score", "timestamp"]
    )


def test_rank_all_and_metrics_match_reference_code(golden_dir, small):
    gold = json.loads((golden_dir / "metrics_golden.json").read_text())
    pids = [str(p) for p in small["product_ids"]]
    qids = list(gold["rank_all_top100"].keys())
    got = oracle.rank_all(small["queries"], small["items"], qids, pids, limit=100)
    assert got == gold["rank_all_top100"]
    rel = {k: set(v) for k, v in gold["relevant"].items()}
    m = oracle.ir_metrics(got, rel)
    for key, val in gold["metrics"].items():
        assert m[key] == pytest.approx(val, abs=1e-12), key
    for case in gold["extra"]:
        m = oracle.ir_metrics(case["rankings"], {k: set(v) for k, v in case["relevant"].items()})
        for key, val in case["metrics"].items():
            assert m[key] == pytest.approx(val, abs=1e-12), key


def test_cos_sim_known_answers():
    d = 384
    e = torch.eye(d)
    assert torch.allclose(oracle.cos_sim(e[3], e[3] * 7.5), torch.tensor([[1.0]]))
    assert torch.allclose(oracle.cos_sim(e[3], e[4]), torch.tensor([[0.0]]))
    assert torch.allclose(oracle.cos_sim(e[3], -2 * e[3]), torch.tensor([[-1.0]]))
    z = oracle.cos_sim(torch.zeros(d), e[:5])
    assert torch.isfinite(z).all() and (z == 0).all()  # eps clamp: zero vector -> 0, not NaN
    # numpy / list / 1-D inputs are accepted like ST's util
    assert oracle.cos_sim(np.ones(4, dtype=np.float32), [[1.0, 1.0, 1.0, 1.0]]).shape == (1, 1)
    # identity catalog: top-k ids are the k largest |q_i| with positive sign first
    q = torch.tensor([0.1, -0.9, 0.5, 0.3, 0.0, 0.7])
    v, i = oracle.cos_topk(q, torch.eye(6), 3)
    assert i.tolist() == [[5, 2, 3]]


def test_mnrl_known_answers_and_records(golden_dir):
    B, d, scale = 16, 32, 20.0
    q, _ = torch.linalg.qr(torch.randn(d, d, generator=torch.Generator().manual_seed(0)))
    a = q[:B]
    loss = oracle.mnrl_loss(a, a.clone(), scale)
    assert loss.item() == pytest.approx(math.log(1 + (B - 1) * math.exp(-scale)), abs=1e-6)
    same = torch.ones(B, d)
    assert oracle.mnrl_loss(same, same, scale).item() == pytest.approx(math.log(B), abs=1e-5)
    z = np.load(golden_dir / "oracle_records.npz")
    a, p = torch.from_numpy(z["mnrl_a"]), torch.from_numpy(z["mnrl_p"])
    for s in (20, 30):
        l, ga, gp = oracle.mnrl_loss_and_grads(a, p, float(s))
        assert l.item() == pytest.approx(float(z[f"mnrl_loss{s}"]), abs=1e-5)
        np.testing.assert_allclose(ga.numpy(), z[f"mnrl_ga{s}"], atol=1e-6)
        np.testing.assert_allclose(gp.numpy(), z[f"mnrl_gp{s}"], atol=1e-6)


def test_unnormalised_topk_matches_f64_record(golden_dir):
    z = np.load(golden_dir / "oracle_records.npz")
    v, i = oracle.cos_topk(torch.from_numpy(z["un_q"]), torch.from_numpy(z["un_c"]), 20)
    err, mism = oracle.compare_topk(v, i, z["un_topk_vals_f64"], z["un_topk_idx"], rtol=1e-5)
    assert err < 1e-5 and mism == 0


def test_ir_eval_core_equals_plain_topk():
    c = oracle.synth_isotropic(1200, 32, 1)
    q = oracle.synth_isotropic(7, 32, 2)
    lists = oracle.ir_eval_topk(q, c, max_k=20, corpus_chunk_size=500)
    v, i = oracle.cos_topk(q, c, 20)
    for r, lst in enumerate(lists):
        assert [ci for _, ci in lst] == i[r].tolist()
        np.testing.assert_allclose([s for s, _ in lst], v[r].numpy(), rtol=1e-6)


def test_compare_topk_is_tie_tolerant():
    rv = np.array([[0.9, 0.5, 0.5 + 1e-9, 0.1]])
    ri = np.array([[4, 7, 8, 1]])
    err, mism = oracle.compare_topk(rv, np.array([[4, 8, 7, 1]]), rv, ri, rtol=1e-5)
    assert mism == 0
    err, mism = oracle.compare_topk(rv, np.array([[5, 7, 8, 1]]), rv, ri, rtol=1e-5)
    assert mism == 1


def _c_oracle():
    """The plain-C restatement (oracle/c_oracle.c), built on demand: an implementation that shares nothing with torch."""
    import ctypes
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parents[1] / "oracle"
    subprocess.run(["make", "-s", "-C", str(root)], check=True)
    lib = ctypes.CDLL(str(root / "_build" / "liboracle_c.so"))
    lib.oc_cos_topk.restype = ctypes.c_int
    lib.oc_cos_topk.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    lib.oc_mnrl_loss.restype = ctypes.c_double
    lib.oc_mnrl_loss.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_double]
    return lib


def test_torch_oracle_agrees_with_the_plain_c_restatement(golden_dir):
    lib = _c_oracle()
    z = np.load(golden_dir / "embeddings_small.npz")
    items = np.ascontiguousarray(z["items"], dtype=np.float32)
    queries = np.ascontiguousarray(z["queries"][:40], dtype=np.float32)
    unnorm = oracle.synth_unnormalised(700, 48, seed=3).numpy()
    for q, c, k in ((queries, items, 100), (unnorm[:20].copy(), unnorm, 25), (unnorm[:3].copy(), unnorm[:10].copy(), 10)):
        Q, N, D = q.shape[0], c.shape[0], c.shape[1]
        sc = np.empty((Q, k), dtype=np.float64)
        ids = np.empty((Q, k), dtype=np.int64)
        assert lib.oc_cos_topk(q.ctypes.data, Q, c.ctypes.data, N, D, k, sc.ctypes.data, ids.ctypes.data) == 0
        v, i = oracle.cos_topk(q, c, k)
        err, mism = oracle.compare_topk(v, i, torch.from_numpy(sc).float(), torch.from_numpy(ids), rtol=1e-5)
        assert err < 1e-5 and mism == 0, (err, mism)
    g = torch.Generator().manual_seed(9)
    a = torch.randn(33, 48, generator=g).numpy()
    p = (torch.randn(33, 48, generator=g) * 3).numpy()
    for scale in (20.0, 30.0):
        want = lib.oc_mnrl_loss(a.ctypes.data, p.ctypes.data, 33, 48, scale)
        assert oracle.mnrl_loss(torch.from_numpy(a), torch.from_numpy(p), scale).item() == pytest.approx(want, abs=2e-5)
