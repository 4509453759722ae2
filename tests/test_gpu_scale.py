"""Parity at the SHAPE AND SCALE of BASELINE configs 4 and 5 ("close the fixable parity gaps"):
multi-million-row bf16 catalogs, D = 768 with request-sized to 1024-query batches and D = 384 with 4096-query batches.

Two witnesses per case: the oracle (fp32 math on the bf16-rounded inputs, chunked over the catalog on the host CPU) for a
sample of 64 queries, and the same expression in plain torch fp32 on the GPU for every query. Ids are compared
tie-tolerantly, the k-th returned score is held against the witnesses' k-th (rank-k membership).
"""

import numpy as np
import pytest
import torch

import instacart_next_order_recommendation_b200 as icr
from instacart_next_order_recommendation_b200 import ops
from oracle import oracle

pytestmark = pytest.mark.gpu

K = 100
RTOL = 5e-5  # measured ~2e-6; BASELINE's bar for bf16 is 2e-3


def _make_catalog(n, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rows = torch.empty(n, d, dtype=torch.bfloat16, device="cuda")
    for s in range(0, n, 1 << 19):
        e = min(n, s + (1 << 19))
        # clustered like the aisle structure of the real catalog: near-ties at the k-th rank are common, not rare
        x = torch.randn(e - s, d, device="cuda", generator=g)
        x[:, : d // 8] += 2.0 * torch.randn(1, d // 8, device="cuda", generator=g)
        rows[s:e] = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
    return rows


def _torch_witness(q, rows, k, chunk=1 << 17):
    """torch eager fp32 on the GPU, chunked over the catalog: normalize -> mm -> topk -> merge."""
    qn = torch.nn.functional.normalize(q.float(), dim=1)
    bv = torch.full((q.shape[0], 0), float("-inf"), device="cuda")
    bi = torch.zeros((q.shape[0], 0), dtype=torch.int64, device="cuda")
    for s in range(0, rows.shape[0], chunk):
        cn = torch.nn.functional.normalize(rows[s : s + chunk].float(), dim=1)
        v, i = torch.topk(qn @ cn.T, min(k, cn.shape[0]), dim=1)
        bv, bi = torch.cat([bv, v], 1), torch.cat([bi, i + s], 1)
        if bv.shape[1] > k:
            bv, pos = torch.topk(bv, k, dim=1)
            bi = torch.gather(bi, 1, pos)
    return bv.cpu(), bi.cpu()


def _check(v, i, rv, ri, n):
    v, i = v.cpu(), i.cpu()
    assert (v[:, :-1] >= v[:, 1:]).all() and (i >= 0).all() and (i < n).all()
    err, mism = oracle.compare_topk(v, i, rv, ri, rtol=RTOL)
    assert err <= RTOL, f"max relative score error {err}"
    assert mism == 0, f"{mism} id mismatches outside ties"
    rk = np.asarray(rv, dtype=np.float64)[:, -1]
    assert (v.double().numpy()[:, -1] >= rk - RTOL * np.maximum(np.abs(rk), 0.05)).all(), "k-th returned score below the witness's k-th"
    for r in range(0, i.shape[0], max(1, i.shape[0] // 16)):
        assert len(set(i[r].tolist())) == i.shape[1]


@pytest.fixture(scope="module")
def c4_rows():
    rows = _make_catalog(2_000_000, 768, seed=41)
    yield rows, rows.cpu()
    del rows
    torch.cuda.empty_cache()


@pytest.fixture(scope="module")
def c5_rows():
    rows = _make_catalog(2_500_000, 384, seed=51)
    yield rows, rows.cpu()
    del rows
    torch.cuda.empty_cache()


@pytest.mark.parametrize("Q", [1, 16, 256, 1024])
def test_c4_shape_bf16_768_vs_oracle_and_torch(c4_rows, Q):
    rows, rows_cpu = c4_rows
    n = rows.shape[0]
    g = torch.Generator(device="cuda").manual_seed(1000 + Q)
    src = torch.randint(0, n, (Q,), device="cuda", generator=g)
    q = torch.nn.functional.normalize(rows[src].float() + 0.05 * torch.randn(Q, 768, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    cat = icr.DeviceCatalog(rows, dtype=torch.bfloat16)
    v, i = cat.topk(q, K)
    _check(v, i, *_torch_witness(q, rows, K), n)  # every query
    sample = torch.arange(0, Q, max(1, Q // 64))[:64]
    rv, ri = oracle.cos_topk_catalog_chunked(q.cpu()[sample], rows_cpu, K)
    _check(v[sample.cuda()], i[sample.cuda()], rv, ri, n)


def test_c5_shape_bf16_384_q4096_vs_oracle_and_torch(c5_rows):
    rows, rows_cpu = c5_rows
    n, Q = rows.shape[0], 4096
    g = torch.Generator(device="cuda").manual_seed(77)
    q = torch.nn.functional.normalize(torch.randn(Q, 384, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    cat = icr.DeviceCatalog(rows, dtype=torch.bfloat16)
    v, i = cat.topk(q, K)
    _check(v, i, *_torch_witness(q, rows, K), n)
    sample = torch.arange(0, Q, Q // 64)[:64]
    rv, ri = oracle.cos_topk_catalog_chunked(q.cpu()[sample], rows_cpu, K)
    _check(v[sample.cuda()], i[sample.cuda()], rv, ri, n)


def test_c2_shape_f32_every_query_vs_torch_witness():
    """The headline shape, every query: the screened fp32 path (one fp16 MMA term + exact re-scoring) against torch fp32."""
    N, D, Q = 49_688, 384, 10_000
    g = torch.Generator(device="cuda").manual_seed(1234)
    items = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1)
    queries = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1)
    v, i = icr.cos_topk(queries, items, K)
    rv, ri = _torch_witness(queries, items, K)
    v, i = v.cpu(), i.cpu()
    err, mism = oracle.compare_topk(v, i, rv, ri, rtol=1e-5)
    assert err <= 1e-5 and mism == 0, (err, mism)
    rk = rv.double().numpy()[:, -1]
    assert (v.double().numpy()[:, -1] >= rk - 1e-5 * np.maximum(np.abs(rk), 0.05)).all()
