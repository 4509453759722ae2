"""Row-sharded top-k on N GPUs: NCCL all-gather vs the NVLink peer-memory exchange kernel (same results, latency of each).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 benchmarks/peer_exchange_case.py
"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rows_per, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000), 768
g = torch.Generator(device=dev).manual_seed(100 + rank)
shard = torch.nn.functional.normalize(torch.randn(rows_per, D, device=dev, generator=g), dim=1).to(torch.bfloat16)
cats = {ex: icr.ShardedCatalog(shard, row_offset=rank * rows_per, total_rows=world * rows_per, dtype=torch.bfloat16, exchange=ex) for ex in ("nccl", "peer")}
ok = True
rows = []
for Q in (1, 8, 64, 1024, 4096):
    k = 100
    g2 = torch.Generator(device=dev).manual_seed(7 + Q)
    q = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=g2), dim=1).to(torch.bfloat16)
    res, ms = {}, {}
    for ex, cat in cats.items():
        for _ in range(3):
            v, i = cat.topk(q, k)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if Q <= 64 else 5
        e0.record()
        for _ in range(reps):
            v, i = cat.topk(q, k)
        e1.record(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[ex], ms[ex] = (v.clone(), i.clone()), t.item()
        # exchange + merge alone, on fixed local candidates
        lv, li = cat.local_topk(q, k)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            if ex == "peer":
                cat._peer.exchange_merge(lv, li)  # one kernel: push, flags, wait, merge
                continue
            else:
                mine = icr.sharded.pack_candidates(lv, li)
                out = torch.empty((world * 2, *mine.shape[1:]), dtype=mine.dtype, device=dev)
                dist.all_gather_into_tensor(out, mine)
                s_, g_ = icr.sharded.unpack_candidates(out.view(world, *mine.shape))
            icr.ops.topk_merge(s_, g_, k)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms[ex + "_x"] = t.item()
    same = torch.equal(res["nccl"][0], res["peer"][0]) and torch.equal(res["nccl"][1], res["peer"][1])
    ok &= same
    rows.append((Q, ms["nccl"], ms["peer"], ms["nccl_x"], ms["peer_x"], same))
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"{world} GPUs, {rows_per} x {D} bf16 rows per GPU, top-100")
    print("| Q | whole call NCCL ms | whole call peer ms | exchange+merge NCCL us | exchange+merge peer us | identical |")
    print("|---|---|---|---|---|---|")
    for Q, a, b, c, d, s in rows:
        print(f"| {Q} | {a:.3f} | {b:.3f} | {c * 1e3:.1f} | {d * 1e3:.1f} | {s} |")
    print("PEER_EXCHANGE_OK" if flag.item() == 1 else "PEER_EXCHANGE_MISMATCH")
dist.destroy_process_group()
