"""K1 (batch-1 GEMV + fused top-k) across row shapes at ~4 GB catalogs: does streaming stay at the HBM roofline for every D / dtype?"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent))
import sweep  # noqa: E402

torch.cuda.set_device(0)
for D, dt in ((384, torch.bfloat16), (384, torch.float32), (768, torch.bfloat16), (768, torch.float32), (128, torch.bfloat16), (1024, torch.bfloat16)):
    esz = 2 if dt == torch.bfloat16 else 4
    N = int(4e9 / (D * esz))
    for Q in (1, 4):
        sweep.topk_case(f"K1 shape D={D} {str(dt).split('.')[-1]} Q={Q}", N, D, Q, 100, dt, iters=8)
for r in sweep.ROWS:
    print(f"{r['config']}: N={r['N']} call {r['ms']:.3f} ms kernel {r['kernel']} {r['kernel_ms']:.3f} ms -> {r['hbm_gbs_kernel']:.0f} GB/s ({r['roofline_frac_kernel']:.3f} of peak)")
