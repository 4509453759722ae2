"""Run one named case a few times (target for `ncu -k regex:... `). Development aid."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import instacart_next_order_recommendation_b200 as icr  # noqa: E402

CASES = {
    "c1": (49_688, 384, 1, 10, torch.float32),
    "c1bf16": (49_688, 384, 1, 10, torch.bfloat16),
    "c4q64": (1_250_000, 768, 64, 100, torch.bfloat16),
    "c4q1": (1_250_000, 768, 1, 100, torch.bfloat16),
    "c5": (4_000_000, 384, 4096, 100, torch.bfloat16),
    "c2bf16": (49_688, 384, 10_000, 100, torch.bfloat16),
    "c2": (49_688, 384, 10_000, 100, torch.float32),
    "c2half": (24_844, 384, 10_000, 100, torch.float32),  # 38 MB of fp32 rows: the re-scoring's gathers stay in one die's share of L2
    "c2quarter": (12_422, 384, 10_000, 100, torch.float32),
    "c1q4": (49_688, 384, 4, 10, torch.float32),
    "c4q16": (1_250_000, 768, 16, 100, torch.bfloat16),
    "c1q32": (49_688, 384, 32, 10, torch.float32),
    "c1q32bf16": (49_688, 384, 32, 10, torch.bfloat16),
    "c1q256": (49_688, 384, 256, 10, torch.float32),
    "c4q128": (1_250_000, 768, 128, 100, torch.bfloat16),
    "c4q256": (1_250_000, 768, 256, 100, torch.bfloat16),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="c1")
    ap.add_argument("--iters", type=int, default=4)
    a = ap.parse_args()
    N, D, Q, k, dt = CASES[a.case]
    g = torch.Generator(device="cuda").manual_seed(0)
    cat = icr.DeviceCatalog(torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1), dtype=dt)
    q = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(dt)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(a.iters):
        flush.zero_()
        v, i = cat.topk(q, k)
    torch.cuda.synchronize()
    print(a.case, "ok", v[0, :3].tolist())
