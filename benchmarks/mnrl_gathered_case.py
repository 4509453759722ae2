"""MNRL with cross-device in-batch negatives on N GPUs (NCCL all-gather of positives + rectangular tensor-core kernels +
reduce-scatter of candidate gradients) against the oracle's single-process restatement.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 benchmarks/mnrl_gathered_case.py
"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr
from oracle import oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for B, D, dt in ((64, 384, torch.float32), (256, 384, torch.bfloat16)):
    g = torch.Generator().manual_seed(17)
    A = [torch.randn(B, D, generator=g).to(dt) for _ in range(world)]
    P = [(torch.randn(B, D, generator=g) * 2).to(dt) for _ in range(world)]
    a = A[rank].to(dev).requires_grad_(True)
    p = P[rank].to(dev).requires_grad_(True)
    loss = icr.mnrl_loss_gathered(a, p, 20.0)
    loss.backward()
    losses, ga, gp = oracle.mnrl_gathered_reference([x.float() for x in A], [x.float() for x in P], 20.0)
    tol = 1e-4 if dt == torch.float32 else 1e-4 + 2 ** -7 * max(ga[rank].abs().max().item(), gp[rank].abs().max().item())
    e = (abs(loss.item() - losses[rank].item()), (a.grad.float().cpu() - ga[rank]).abs().max().item(), (p.grad.float().cpu() - gp[rank]).abs().max().item())
    ok &= e[0] <= 1e-4 and e[1] <= tol and e[2] <= tol
    # timing
    for _ in range(5):
        a.grad = p.grad = None; icr.mnrl_loss_gathered(a, p, 20.0).backward()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        a.grad = p.grad = None; icr.mnrl_loss_gathered(a, p, 20.0).backward()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"B={B} x {world} ranks, D={D}, {dt}: |dloss| {e[0]:.2e} |dgrad_a| {e[1]:.2e} |dgrad_p| {e[2]:.2e}; fwd+bwd {e0.elapsed_time(e1) / 20 * 1e3:.0f} us")
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MNRL_GATHERED_OK" if flag.item() == 1 else "MNRL_GATHERED_MISMATCH")
dist.destroy_process_group()
