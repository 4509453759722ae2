"""Small cases of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck). Development aid.

    compute-sanitizer --tool memcheck python benchmarks/sanitize_case.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import instacart_next_order_recommendation_b200 as icr  # noqa: E402
from instacart_next_order_recommendation_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(3)
N, D = 49_688, 384
items = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1)
for dt in (torch.float32, torch.bfloat16):
    cat = icr.DeviceCatalog(items, dtype=dt)
    for Q, k in ((1, 10), (1, 100), (3, 16), (7, 10), (8, 10), (32, 100), (64, 10)):
        q = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(dt)
        v, i = cat.topk_small(q, k) if Q <= 7 else cat.topk(q, k)
        ref = torch.topk(torch.nn.functional.normalize(q.float(), dim=1) @ torch.nn.functional.normalize(cat.rows.float(), dim=1).T, k, dim=1)
        torch.cuda.synchronize()
        err = (v - ref.values).abs().max().item()
        print(dt, Q, k, "launches", ops.last_launch_count(), "max abs err", f"{err:.2e}")
        assert err < (1e-5 if dt == torch.float32 else 2e-3)
print("ok")
