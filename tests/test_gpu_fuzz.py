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
