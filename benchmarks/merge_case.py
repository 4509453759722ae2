"""Time the K4 shard merge (icr_topk_merge) alone. Development aid."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instacart_next_order_recommendation_b200 import ops  # noqa: E402

for G, Q, k in ((8, 4096, 100), (2, 4096, 100), (8, 1, 100), (8, 64, 10)):
    g = torch.Generator(device="cuda").manual_seed(1)
    s = torch.rand(G, Q, k, device="cuda", generator=g).sort(dim=2, descending=True).values
    i = torch.randint(0, 100_000_000, (G, Q, k), device="cuda", generator=g)
    for _ in range(3):
        v, idx = ops.topk_merge(s, i, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        v, idx = ops.topk_merge(s, i, k)
    e1.record()
    torch.cuda.synchronize()
    ref = torch.topk(s.permute(1, 0, 2).reshape(Q, G * k), k, dim=1).values
    print(f"G={G} Q={Q} k={k}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per merge; max |dv| {float((v - ref).abs().max()):.1e}")
