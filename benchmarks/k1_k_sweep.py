"""Batch-1 request on the config-1 catalog by k: device-side time (host running ahead), cold by rotation over catalog copies."""
import sys

import torch

sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
items = torch.nn.functional.normalize(torch.randn(49_688, 384, device="cuda", generator=g), dim=1)
q = torch.nn.functional.normalize(torch.randn(1, 384, device="cuda", generator=g), dim=1)
for dt in (torch.float32, torch.bfloat16):
    copies = [icr.DeviceCatalog(items.clone(), dtype=dt) for _ in range(4 if dt == torch.float32 else 8)]
    qd = q.to(dt)
    row = []
    for _ in range(2000):  # clocks up before the first measurement
        copies[0].topk_small(qd, 10, copy=False)
    torch.cuda.synchronize()
    for k in (10, 16, 17, 24, 32, 33, 48, 64, 100, 128, 10):
        for c in copies:
            c.topk_small(qd, k)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
        for i, (a, b) in enumerate(ev):
            torch.cuda._sleep(80_000)
            a.record()
            copies[i % len(copies)].topk_small(qd, k, copy=False)
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
        row.append(f"k={k}: {ts[30]:.1f}")
    print(str(dt), " | ".join(row))
    del copies
    torch.cuda.empty_cache()
