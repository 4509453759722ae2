"""Seeded random shapes through every kernel family of the fused cosine top-k, against the oracle (B200 only).

Covers what the hand-picked cases of test_gpu_parity.py may miss: row lengths with every kind of vector tail, catalogs
smaller than k, ring kernels with few slabs per CTA, masks, row offsets, un-normalised inputs, resident catalogs
(precomputed norms / planes) and raw tensors, on the GEMV, GEMM and automatically chosen paths.
"""

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


@pytest.mark.parametrize("case", _cases(72, 20261018) + [(100 + c[0],) + c[1:] for c in _cases(20, 7, big=True)], ids=lambda c: f"{c[0]}-Q{c[1]}-N{c[2]}-D{c[3]}-k{c[4]}-{str(c[5]).split('.')[-1]}-p{c[6]}-m{int(c[7])}-r{int(c[8])}")
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
