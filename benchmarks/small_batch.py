"""Request-sized batches on the config-1 catalog (49,688 x 384): whole-call time per batch, L2 flushed before each call."""
import os
import sys

import torch

sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr  # noqa: E402
from instacart_next_order_recommendation_b200 import ops  # noqa: E402

N, D = 49_688, 384
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1234)
items = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=g), dim=1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
QS = [int(x) for x in os.environ.get("ICR_SMALL_QS", "1,2,4,7,8,16,32,64,128,256,512,1024").split(",")]


def timed(fn, n=30):
    for _ in range(5):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    return ts[len(ts) // 2]


for dt in (torch.float32, torch.bfloat16):
    cat = icr.DeviceCatalog(items, dtype=dt)
    for k in (10, 100):
        row = []
        for Q in QS:
            q = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=g), dim=1).to(dt)
            t = timed(lambda: cat.topk(q, k))
            n_launch = ops.last_launch_count()
            tg = timed(lambda: cat.topk_small(q, k, copy=False)) if Q <= 256 else float("nan")
            row.append(f"Q={Q}: {t:.0f} us, graph {tg:.0f} us ({n_launch} launches)")
        print(f"{str(dt):15s} k={k}: " + " | ".join(row))
