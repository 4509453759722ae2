"""One C4 shard as a rank of 8 sees it (1.25M x 768 bf16, 1.92 GB): whole local top-100 per batch size."""
import sys

import torch

sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr  # noqa: E402
from instacart_next_order_recommendation_b200 import ops  # noqa: E402

N, D, k = 1_250_000, 768, 100
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(5)
rows = torch.empty(N, D, dtype=torch.bfloat16, device=dev)
for s in range(0, N, 1 << 18):
    e = min(N, s + (1 << 18))
    rows[s:e] = torch.nn.functional.normalize(torch.randn(e - s, D, device=dev, generator=g), dim=1).to(torch.bfloat16)
cat = icr.DeviceCatalog(rows, dtype=torch.bfloat16)
nbytes = N * D * 2
peak = 6550.7
import os  # noqa: E402

for Q in [int(x) for x in os.environ.get("ICR_C4_QS", "1,2,4,16,64,128,256,512,1024").split(",")]:
    q = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=g), dim=1).to(torch.bfloat16)
    for _ in range(3):
        cat.topk(q, k)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        a.record()
        cat.topk(q, k)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    ms = ts[len(ts) // 2]
    n_launch = ops.last_launch_count()
    kt = ops.kernel_timing(lambda: cat.topk(q, k), 10)
    print(f"Q={Q:5d}: {ms:.3f} ms per batch ({n_launch} launches) = {nbytes / ms / 1e6 / peak:.2f} of HBM peak, "
          f"{2.0 * Q * N * D / ms / 1e9:.0f} TFLOP/s; dominant kernel {kt['ms_per_step']:.3f} ms = {nbytes / kt['ms_per_step'] / 1e6 / peak:.2f} of HBM peak")
