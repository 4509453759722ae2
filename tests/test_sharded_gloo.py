"""world_size-2 gloo run of the sharded-catalog host logic (partitioning, candidate packing, all-gather
layout, merge call). The CUDA kernels are replaced by the oracle through the test seams of ShardedCatalog;
the kernels themselves are covered by the -m gpu tests."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_local_topk(queries, rows, k, row_offset):
    from oracle import oracle

    v, i = oracle.cos_topk(queries, rows, k)
    return v, i + row_offset


def _oracle_merge(scores, ids, k):
    # scores/ids [G,Q,k_in] -> exact top-k by (score desc, id asc), empty slots (id < 0) last
    G, Q, kin = scores.shape
    s = scores.permute(1, 0, 2).reshape(Q, G * kin).double()
    i = ids.permute(1, 0, 2).reshape(Q, G * kin)
    s = torch.where(i < 0, torch.full_like(s, float("-inf")), s)
    order = np.lexsort((i.numpy(), -s.numpy()), axis=1)[:, :k]
    order = torch.from_numpy(order)
    return s.gather(1, order).float(), i.gather(1, order)


def _worker(rank: int, world: int, port: int, n_rows: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from instacart_next_order_recommendation_b200.sharded import ShardedCatalog
        from oracle import oracle

        full = oracle.synth_clustered(n_rows, 32, seed=11, n_centres=8)[0]
        queries = oracle.synth_isotropic(9, 32, seed=12)
        cat = ShardedCatalog.from_full(full, _local_topk=_oracle_local_topk, _merge=_oracle_merge)
        assert cat.world_size == world and cat.rank == rank
        k = 20
        v, i = cat.topk(queries, k)
        rv, ri = oracle.cos_topk(queries, full, k)
        kk = min(k, n_rows)
        assert v.shape == (9, kk)
        err, mism = oracle.compare_topk(v, i, rv, ri, rtol=1e-6)
        assert err < 1e-6 and mism == 0, (err, mism)
        torch.save({"v": v, "i": i}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _run(n_rows, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_rows, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "r0.pt")
    b = torch.load(tmp_path / "r1.pt")
    assert torch.equal(a["i"], b["i"]) and torch.equal(a["v"], b["v"])  # every rank ends with the same answer


def test_two_rank_sharded_topk_equals_unsharded(tmp_path):
    _run(501, tmp_path)  # ragged: 251 + 250 rows


def test_two_rank_short_shards(tmp_path):
    _run(15, tmp_path)  # fewer rows than k: shards pad with (-inf, -1), k clamps to the catalog size


# ---- cross-device negatives for MNRL: all-gather of positives, label offset, reduce-scatter of candidate gradients ----


def _oracle_mnrl_fwd(a, cand, scale, offset):
    from oracle import oracle

    loss, ga, gc = oracle.mnrl_rect_loss_and_grads(a, cand, scale, offset)
    return loss, torch.cat([ga.flatten(), gc.flatten()])  # "saved": the unscaled gradients


def _oracle_mnrl_bwd(a, cand, scale, offset, saved, grad_out):
    B, D = a.shape
    ga = saved[: B * D].view(B, D) * grad_out
    gc = saved[B * D :].view(cand.shape[0], D) * grad_out
    return ga, gc


def _mnrl_worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from instacart_next_order_recommendation_b200.losses import mnrl_loss_gathered
        from oracle import oracle

        B, D, scale = 6, 16, 20.0
        g = torch.Generator().manual_seed(5)
        A = [torch.randn(B, D, generator=g) for _ in range(world)]
        P = [torch.randn(B, D, generator=g) * 2 for _ in range(world)]
        a = A[rank].clone().requires_grad_(True)
        p = P[rank].clone().requires_grad_(True)
        loss = mnrl_loss_gathered(a, p, scale, _kernels=(_oracle_mnrl_fwd, _oracle_mnrl_bwd))
        (loss * 0.5).backward()
        losses, grads_a, grads_p = oracle.mnrl_gathered_reference(A, P, scale)
        assert abs(loss.item() - losses[rank].item()) < 1e-6
        assert (a.grad - 0.5 * grads_a[rank]).abs().max() < 1e-6
        assert (p.grad - 0.5 * grads_p[rank]).abs().max() < 1e-6  # sum over ranks of d loss_r / d P_rank
        torch.save({"ok": True}, os.path.join(out_dir, f"m{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_mnrl_with_gathered_negatives(tmp_path):
    port = _free_port()
    mp.spawn(_mnrl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "m0.pt").exists() and (tmp_path / "m1.pt").exists()
