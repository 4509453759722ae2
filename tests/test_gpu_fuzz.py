"""Seeded random shapes through every kernel family of the fused cosine top-k, against the oracle (B200 only).

Covers what the hand-picked cases of test_gpu_parity.py may miss: row lengths with every kind of vector tail, catalogs
smaller than k, ring kernels with few slabs per CTA, masks, row offsets, un-normalised inputs, resident catalogs
(precomputed norms / planes) and raw tensors, on the GEMV, GEMM and automatically chosen paths.
"""

import os

import numpy as np
import pytest
import torch

import instacart_next_order_recommendation_b200 as icr
from instacart_next_order_recommendation_b200 import ops
from oracle import oracle

pytestmark = pytest.mark.gpu

DIMS = [8, 16, 64, 72, 128, 200, 256, 384, 520, 768, 1000, 1024]
QS = [1, 1, 1, 2, 3, 5, 7, 8, 9, 33, 130, 300]
KS = [1, 5, 10, 100, 256]


def _cases(n, seed, big=False):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        D = int(rng.choice([64, 128, 200, 384] if big else DIMS))
        Q = int(rng.choice(QS))
        if big:  # many slabs per ring CTA, several phases on the tensor path
            N = int(rng.choice([150_001, 300_000]))
        else:
            N = int(rng.choice([1, 7, 100, 1000, 4097, 20000, 60001])) if i % 3 else int(rng.integers(1, 60000))
        k = int(rng.choice(KS))
        dtype = torch.float32 if rng.random() < 0.5 else torch.bfloat16
        path = int(rng.choice([ops.PATH_AUTO, ops.PATH_GEMV, ops.PATH_GEMM]))
        out.append((i, Q, N, D, k, dtype, path, bool(rng.random() < 0.3), bool(rng.random() < 0.5)))
    return out


# ICR_FUZZ_CASES / ICR_FUZZ_SEED widen the search for an occasional long run (defaults keep the suite fast)
_N_CASES = int(os.environ.get("ICR_FUZZ_CASES", "72"))
_SEED = int(os.environ.get("ICR_FUZZ_SEED", "20261018"))


@pytest.mark.parametrize("case", _cases(_N_CASES, _SEED) + [(10_000 + c[0],) + c[1:] for c in _cases(max(20, _N_CASES // 4), 7 + _SEED % 1000, big=True)], ids=lambda c: f"{c[0]}-Q{c[1]}-N{c[2]}-D{c[3]}-k{c[4]}-{str(c[5]).split('.')[-1]}-p{c[6]}-m{int(c[7])}-r{int(c[8])}")
def test_random_shape_against_oracle(case):
    i, Q, N, D, k, dtype, path, use_mask, resident = case
    items = oracle.synth_unnormalised(N, D, seed=1000 + i).to(dtype)
    queries = oracle.synth_unnormalised(Q, D, seed=2000 + i).to(dtype)
    if i % 5 == 0 and N > 3:
        items[N // 2] = 0  # a zero row scores 0, not NaN
    mask = None
    if use_mask and N > 1:
        mask = (torch.rand(N, generator=torch.Generator().manual_seed(i)) < 0.3).to(torch.uint8)
        mask[int(torch.randint(0, N, (1,)).item())] = 0  # at least one eligible row
    kk = min(k, N)
    scores = oracle.cos_sim(queries.float(), items.float())
    if mask is not None:
        scores[:, mask.bool()] = float("-inf")
    rv, ri = torch.topk(scores, kk, dim=1)
    ri = torch.where(torch.isinf(rv), torch.full_like(ri, -1), ri)
    md = mask.cuda() if mask is not None else None
    if resident:
        cat = icr.DeviceCatalog(items.cuda(), dtype=dtype, row_offset=11)
        v, ids = cat.topk(queries.cuda(), kk, exclude_mask=md, path=path)
        ids = torch.where(ids >= 0, ids - 11, ids)
    else:
        v, ids = ops.cos_topk(queries.cuda(), items.cuda(), kk, exclude_mask=md, path=path)
    rtol = 1e-5 if dtype == torch.float32 else 5e-5
    v, ids = v.cpu(), ids.cpu()
    assert v.shape == rv.shape
    live = ~torch.isinf(rv)
    assert torch.equal(torch.isinf(v), ~live), "eligible-row count differs"
    assert (ids[~live] == -1).all()
    err, mism = oracle.compare_topk(torch.where(live, v, torch.zeros_like(v)), torch.where(live, ids, torch.zeros_like(ids)),
                                    torch.where(live, rv, torch.zeros_like(rv)), torch.where(live, ri, torch.zeros_like(ri)), rtol=rtol)
    assert err <= rtol and mism == 0, (err, mism)


def _mnrl_cases(n, seed):
    rng = np.random.default_rng(seed)
    return [(i, int(rng.integers(2, 700)), int(rng.choice([8, 64, 72, 128, 200, 384, 520, 768, 1024])), float(rng.choice([10.0, 20.0, 30.0])),
             torch.float32 if rng.random() < 0.5 else torch.bfloat16) for i in range(n)]


@pytest.mark.parametrize("case", _mnrl_cases(int(os.environ.get("ICR_FUZZ_MNRL_CASES", "24")), _SEED), ids=lambda c: f"{c[0]}-B{c[1]}-D{c[2]}-s{c[3]}-{str(c[4]).split('.')[-1]}")
def test_random_mnrl_step_against_autograd(case):
    """Random batch sizes (both kernel families: the tensor path starts at B = 288), dims and scales; un-normalised inputs."""
    i, B, D, scale, dtype = case
    a = oracle.synth_unnormalised(B, D, seed=3000 + i).to(dtype)
    p = (0.7 * a.float() + 0.5 * oracle.synth_unnormalised(B, D, seed=4000 + i)).to(dtype)  # positives correlate with their anchors
    ad, pd = a.cuda().requires_grad_(True), p.cuda().requires_grad_(True)
    loss = icr.mnrl_loss(ad, pd, scale)
    loss.backward()
    rl, rga, rgp = oracle.mnrl_loss_and_grads(a.float(), p.float(), scale)
    assert abs(loss.item() - rl.item()) <= 1e-4
    for got, want in ((ad.grad, rga), (pd.grad, rgp)):
        err = (got.float().cpu() - want).abs().max().item()
        # 1e-4 absolute is the bar for unit-norm embeddings (gradient entries <= scale / B). Un-normalised rows with small
        # norms have large gradients (the normalisation Jacobian multiplies by 1 / |x|); the tensor path forms the gradient
        # products from fp16 operands (2^-11 relative each), so the error is bounded relative to the largest entry.
        bound = 1e-4 + (1e-3 + (2 ** -7 if dtype == torch.bfloat16 else 0.0)) * want.abs().max().item()
        assert err <= bound, (err, bound)


def _dense_cases(n, seed):
    rng = np.random.default_rng(seed + 1)
    return [(i, int(rng.choice([1, 3, 63, 64, 65, 200, 513])), int(rng.choice([1, 5, 1000, 1023, 1024, 1025, 3000, 5001])),
             int(rng.choice([8, 64, 72, 200, 384, 768, 1000])), torch.float32 if rng.random() < 0.5 else torch.bfloat16) for i in range(n)]


@pytest.mark.parametrize("case", _dense_cases(int(os.environ.get("ICR_FUZZ_DENSE_CASES", "16")), _SEED), ids=lambda c: f"{c[0]}-Qa{c[1]}-Nb{c[2]}-D{c[3]}-{str(c[4]).split('.')[-1]}")
def test_random_dense_cos_sim_against_oracle(case):
    """cos_sim drop-in across the CUDA-core / tensor-core crossover (64 x 1024), tails on both axes, un-normalised inputs."""
    i, Qa, Nb, D, dtype = case
    a = oracle.synth_unnormalised(Qa, D, seed=5000 + i).to(dtype)
    b = oracle.synth_unnormalised(Nb, D, seed=6000 + i).to(dtype)
    got = ops.cos_sim_dense(a.cuda(), b.cuda()).cpu()
    want = oracle.cos_sim(a.float(), b.float())
    assert got.shape == want.shape and (got - want).abs().max().item() <= (2e-6 if dtype == torch.float32 else 5e-6)


def _merge_cases(n, seed):
    rng = np.random.default_rng(seed + 2)
    return [(i, int(rng.choice([1, 2, 3, 8, 16])), int(rng.choice([1, 7, 64, 600, 5000])), int(rng.choice([1, 10, 100, 256])), int(rng.choice([1, 10, 100, 256])))
            for i in range(n)]


@pytest.mark.parametrize("case", _merge_cases(int(os.environ.get("ICR_FUZZ_MERGE_CASES", "16")), _SEED), ids=lambda c: f"{c[0]}-G{c[1]}-Q{c[2]}-kin{c[3]}-kout{c[4]}")
def test_random_shard_merge_against_sort(case):
    """icr_topk_merge on random shard lists with empty slots and duplicate scores: exact (score desc, id asc) order."""
    i, G, Q, k_in, k_out = case
    g = torch.Generator().manual_seed(7000 + i)
    scores = torch.round(torch.randn(G, Q, k_in, generator=g) * 8) / 8  # coarse grid: plenty of exact ties
    ids = torch.stack([torch.randperm(G * k_in * 4, generator=g)[: G * k_in].view(G, k_in) for _ in range(Q)], dim=1)  # distinct per query
    empty = torch.rand(G, Q, k_in, generator=g) < 0.2
    ids = torch.where(empty, torch.full_like(ids, -1), ids)
    scores = torch.where(empty, torch.full_like(scores, float("-inf")), scores)
    v, idx = ops.topk_merge(scores.cuda(), ids.cuda(), k_out)
    v, idx = v.cpu(), idx.cpu()
    s2 = scores.permute(1, 0, 2).reshape(Q, G * k_in)
    i2 = ids.permute(1, 0, 2).reshape(Q, G * k_in)
    for q in range(min(Q, 50)):
        live = [(float(s), int(d)) for s, d in zip(s2[q].tolist(), i2[q].tolist()) if d >= 0]
        live.sort(key=lambda t: (-t[0], t[1]))
        want = live[:k_out] + [(float("-inf"), -1)] * max(0, k_out - len(live))
        got = list(zip(v[q].tolist(), idx[q].tolist()))
        assert got == want, (q, got[:5], want[:5])


def _rect_cases(n, seed):
    rng = np.random.default_rng(seed + 3)
    out = []
    for i in range(n):
        B = int(rng.choice([32, 48, 64, 100, 256, 300]))
        G = int(rng.choice([1, 2, 3, 8]))
        out.append((i, B, G * B, int(rng.integers(0, G)) * B, int(rng.choice([64, 128, 384, 768])), float(rng.choice([20.0, 30.0])),
                    torch.float32 if rng.random() < 0.5 else torch.bfloat16))
    return out


@pytest.mark.parametrize("case", _rect_cases(int(os.environ.get("ICR_FUZZ_RECT_CASES", "12")), _SEED), ids=lambda c: f"{c[0]}-B{c[1]}-Bc{c[2]}-off{c[3]}-D{c[4]}-{str(c[6]).split('.')[-1]}")
def test_random_gathered_mnrl_against_autograd(case):
    """Rectangular MNRL (B anchors of rank r against the G*B gathered candidates, labels r*B + i) on unit-norm rows."""
    i, B, Bc, off, D, scale, dtype = case
    cand = oracle.synth_clustered(Bc, D, seed=8000 + i, n_centres=10)[0]
    g = torch.Generator().manual_seed(9000 + i)
    a = torch.nn.functional.normalize(cand[off : off + B] + 0.4 * torch.randn(B, D, generator=g), dim=1)
    a, cand = a.to(dtype), cand.to(dtype)
    loss, saved = ops.mnrl_forward_rect(a.cuda(), cand.cuda(), scale, off)
    ga, gc = ops.mnrl_backward_rect(a.cuda(), cand.cuda(), scale, off, saved, torch.tensor(1.0, device="cuda"))
    rl, rga, rgc = oracle.mnrl_rect_loss_and_grads(a.float(), cand.float(), scale, off)
    assert abs(loss.item() - rl.item()) <= 1e-4
    for got, want in ((ga, rga), (gc, rgc)):
        bound = 1e-4 + (2 ** -8 if dtype == torch.bfloat16 else 0.0) * want.abs().max().item()
        assert (got.float().cpu() - want).abs().max().item() <= bound


@pytest.mark.parametrize("Q", [1, 2, 255, 256, 257, 2047, 2048, 2049, 5003])
def test_host_pipeline_piece_boundaries(Q):
    """topk_host (pinned host in / out, pieces on two streams) == the device call for batch sizes around its split rules."""
    items = oracle.synth_clustered(20000, 128, seed=3)[0]
    cat = icr.DeviceCatalog(items.cuda())
    q = oracle.synth_isotropic(Q, 128, seed=Q)
    v, i = cat.topk_host(q.pin_memory(), 10)
    torch.cuda.synchronize()
    dv, di = cat.topk(q.cuda(), 10)
    assert torch.equal(v, dv.cpu()) and torch.equal(i, di.cpu())
    if Q >= 4:
        v2, i2 = cat.topk_host(q, 10, splits=[1, Q - 3, 2])  # pageable input, explicit uneven pieces
        torch.cuda.synchronize()
        # the 1- and 2-query pieces take the GEMV kernels, the device call the tensor path: equal within the fp32 tolerance
        err, mism = oracle.compare_topk(v2, i2, dv.cpu(), di.cpu(), rtol=1e-5)
        assert err <= 1e-5 and mism == 0, (err, mism)
