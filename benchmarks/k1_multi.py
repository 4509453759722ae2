"""Multi-query GEMV passes (K1, 3 / 7 queries per pass) against the swapped tensor-core kernel on large catalogs."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent))
import sweep  # noqa: E402
from instacart_next_order_recommendation_b200 import ops  # noqa: E402

torch.cuda.set_device(0)
for (N, D, dt) in ((1_250_000, 768, torch.bfloat16), (5_208_333, 384, torch.bfloat16), (1_302_083, 384, torch.float32)):
    for Q in (1, 2, 3, 4, 7, 8, 14):
        for path, name in ((ops.PATH_GEMV, "gemv"), (ops.PATH_GEMM, "gemm")):
            sweep.topk_case(f"N={N} D={D} {str(dt).split('.')[-1]} Q={Q} {name}", N, D, Q, 100, dt, path=path, iters=8)
for r in sweep.ROWS:
    print(f"{r['config']}: call {r['ms']:.3f} ms kernel {r['kernel_ms']:.3f} ms ({r['kernel_launches']:.0f} launches) -> call {r['roofline_frac_call']:.3f} of HBM roofline")
