"""icr_cos_topk_sharded / icr_peer_exchange_merge (shard search + NVLink exchange + merge behind one call) on ONE GPU.

The protocol only needs every rank's buffer to be addressable from every rank, so `world` ranks can be played by `world`
streams of one device, each with its own buffer: the kernels of the ranks then wait for each other exactly as they do across
GPUs (a real 2-GPU run is tests/test_gpu_parity.py::test_peer_memory_exchange_equals_nccl_all_gather, which compares this
path with the NCCL route). Simulated ranks must be able to run side by side: the forced GEMV path is used for the shard
search of larger batches (a tensor-core search wants every SM of the device for itself while the other rank's merge waits).
"""
import ctypes
import os

os.environ.setdefault("ICR_PEER_TIMEOUT_S", "30")  # read once by the library: a protocol bug fails the test instead of hanging it

import pytest
import torch

import instacart_next_order_recommendation_b200 as icr
from instacart_next_order_recommendation_b200 import _lib, ops
from oracle import oracle

pytestmark = pytest.mark.gpu


class FakeRanks:
    """`world` zero-filled exchange buffers on one device + one stream per rank; epochs counted like PeerExchange does."""

    def __init__(self, world: int, n_max: int):
        lib = _lib.load()
        nbytes = lib.icr_peer_buffer_bytes(n_max, world)
        self.bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
        self.ptrs = (ctypes.c_uint64 * world)(*[b.data_ptr() for b in self.bufs])
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        self.world, self.n_max, self.epoch = world, n_max, 0
        torch.cuda.synchronize()

    def call(self, fn):
        """fn(rank, peer) on every rank's stream -> list of results; checks that no wait timed out."""
        self.epoch += 1
        out = []
        for r in range(self.world):
            self.streams[r].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.streams[r]):
                out.append(fn(r, (r, self.world, ctypes.addressof(self.ptrs), self.epoch, self.n_max)))
        for s in self.streams:
            torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for r, b in enumerate(self.bufs):
            assert int(b[768:772].view(torch.int32).item()) == 0, f"rank {r} timed out waiting for a peer"
        return out


def _shards(N, D, world, dtype, seed):
    items, _ = oracle.synth_clustered(N, D, seed=seed)
    items = items.cuda().to(dtype)
    per = -(-N // world)
    cats = [icr.DeviceCatalog(items[r * per : min(N, (r + 1) * per)], dtype=dtype, row_offset=r * per) for r in range(world)]
    return items, cats


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("world,Q,k", [(1, 1, 10), (2, 1, 100), (2, 3, 100), (2, 7, 100), (4, 1, 100), (8, 1, 100), (8, 2, 100), (8, 5, 64), (3, 1, 256),
                                        (16, 1, 64), (2, 7, 256)])
def test_request_sized_sharded_call_is_one_launch_and_equals_the_unsharded_search(world, Q, k, dtype):
    """Q <= 7: the merging CTA of the GEMV kernel pushes, waits and merges - the whole sharded request is one launch per rank,
    and every rank holds the top-k of the WHOLE catalog (the oracle's, and the single-GPU call's)."""
    N, D = 30_011, 384
    items, cats = _shards(N, D, world, dtype, seed=world * 100 + Q)
    queries, _ = oracle.synth_queries_from_items(items.float().cpu(), Q, seed=5)
    q = queries.cuda().to(dtype)
    ranks = FakeRanks(world, 4096)
    launches = []

    def one(r, peer):
        res = cats[r].topk(q, k, path=ops.PATH_GEMV, peer=peer)
        launches.append(ops.last_launch_count())
        return res

    want_v, want_i = icr.DeviceCatalog(items, dtype=dtype).topk(q, k, path=ops.PATH_GEMV)
    for _ in range(3):  # epochs 1..3: both slot parities and a reused resident workspace
        launches.clear()
        res = ranks.call(one)
        assert launches == [1] * world
        for v, i in res:
            assert torch.equal(i, want_i)
            # a row's score does not depend on the shard it is scored in - up to the last bit for rows whose tail (D = 384 bf16: 16 of 48
            # vectors) is summed by the other half-warp when the row sits at an odd position of its 8-row slab
            assert torch.equal(v, want_v) if dtype == torch.float32 else torch.allclose(v, want_v, rtol=0, atol=1e-6)
    ov, oi = oracle.cos_topk(q.float().cpu(), items.float().cpu(), k)
    v, i = res[0]
    assert (i.cpu() == oi).float().mean() > 0.98  # swaps across fp32 near-ties only
    assert torch.allclose(v.cpu(), ov, rtol=0, atol=5e-6)


@pytest.mark.parametrize("world,Q,k", [(2, 8, 100), (2, 20, 10), (3, 33, 100), (4, 64, 50)])
def test_batched_sharded_call_on_the_gemv_path(world, Q, k):
    """Q > 7: the shard search, then ONE exchange + merge kernel (K4f) behind the same call."""
    N, D = 20_003, 128
    items, cats = _shards(N, D, world, torch.float32, seed=3)
    queries, _ = oracle.synth_queries_from_items(items.cpu(), Q, seed=6)
    q = queries.cuda()
    ranks = FakeRanks(world, Q * k)
    res = ranks.call(lambda r, peer: cats[r].topk(q, k, path=ops.PATH_GEMV, peer=peer))
    want_v, want_i = icr.DeviceCatalog(items).topk(q, k, path=ops.PATH_GEMV)
    for v, i in res:
        assert torch.equal(i, want_i)
        assert torch.equal(v, want_v)


@pytest.mark.parametrize("world,Q,k", [(1, 5, 10), (2, 1, 100), (2, 1000, 100), (3, 257, 256), (8, 64, 100), (16, 16, 100), (5, 100, 1)])
def test_exchange_merge_kernel_equals_gather_then_merge(world, Q, k):
    """K4f on arbitrary per-rank lists (unsorted, with (-inf, -1) padding) == the lists of all ranks concatenated and ranked by
    (score desc, id asc); identical on every rank; three calls in a row (slot reuse)."""
    g = torch.Generator().manual_seed(world * 1000 + Q)
    ranks = FakeRanks(world, Q * k)
    for rep in range(3):
        vals = torch.randn(world, Q, k, generator=g)
        vals[:, :, -1] = vals[:, :, 0]  # some exact score ties across and inside lists
        ids = torch.stack([torch.randperm(1_000_000, generator=g)[: Q * k].view(Q, k) + r * 1_000_000 for r in range(world)])
        pad = torch.rand(world, Q, k, generator=g) < 0.1
        vals[pad], ids[pad] = float("-inf"), -1
        vd, idd = vals.cuda(), ids.cuda()
        res = ranks.call(lambda r, peer: ops.peer_exchange_merge(vd[r], idd[r], peer))
        flat_v = vals.permute(1, 0, 2).reshape(Q, world * k)
        flat_i = ids.permute(1, 0, 2).reshape(Q, world * k)
        for qi in range(0, Q, max(1, Q // 50)):
            cand = sorted(((-float(s), int(i)) for s, i in zip(flat_v[qi], flat_i[qi]) if i >= 0))[:k]
            want_v = [-s for s, _ in cand] + [float("-inf")] * (k - len(cand))
            want_i = [i for _, i in cand] + [-1] * (k - len(cand))
            for v, i in res:
                assert i[qi].tolist() == want_i
                assert v[qi].tolist() == want_v
        for v, i in res[1:]:
            assert torch.equal(v, res[0][0]) and torch.equal(i, res[0][1])


def test_ranks_may_mix_the_one_launch_and_the_two_step_route():
    """A rank whose shard is shorter than k searches, pads and calls the exchange+merge kernel while its peers end the exchange
    in the tail of their GEMV kernel: same buffers, same epochs, same result."""
    D, k = 384, 100
    items, _ = oracle.synth_clustered(10_000 + 60, D, seed=8)
    items = items.cuda()
    cats = [icr.DeviceCatalog(items[:10_000], row_offset=0), icr.DeviceCatalog(items[10_000:], row_offset=10_000)]
    q = items[10_020:10_021] + 0.01
    ranks = FakeRanks(2, 4096)

    def two_step(peer):
        v, i = cats[1].topk(q, 60)
        v = torch.cat([v, torch.full((1, 40), float("-inf"), device="cuda")], dim=1)
        i = torch.cat([i, torch.full((1, 40), -1, dtype=torch.int64, device="cuda")], dim=1)
        return ops.peer_exchange_merge(v, i, peer)

    # Two ranks in ONE CUDA context: the first use of a kernel (torch's cat / fill, ours) loads its module lazily, which
    # synchronises the context - behind the other "rank's" kernel, which is waiting for this one. Real ranks have a context
    # each; here every kernel of the two-step route is used once before the ranks run side by side.
    FakeRanks(1, 4096).call(lambda r, peer: two_step(peer))
    res = ranks.call(lambda r, peer: cats[0].topk(q, k, peer=peer) if r == 0 else two_step(peer))
    want_v, want_i = ops.cos_topk(q, items, k)
    for v, i in res:
        assert torch.equal(i, want_i)
        assert torch.allclose(v, want_v, rtol=0, atol=2e-6)
