"""Per-section cycle breakdown of the warp-per-query select on the C2 shape (development build: ICR_NVCC_DEFS=-DICR_SELECT_TRACE).

    ICR_NVCC_DEFS=-DICR_SELECT_TRACE python -m instacart_next_order_recommendation_b200.build --force && python benchmarks/select_trace.py [f32|bf16]
"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr  # noqa: E402
from instacart_next_order_recommendation_b200 import _lib  # noqa: E402

NAMES = ["kth: min/max", "kth: hist fill", "kth: scan", "kth: bin range", "kth: refine levels", "kth: boundary list", "kth: rank list",
         "reduce: compaction", "kernel: carry + dense front", "kernel: segment gather", "kernel: outputs", "kernel: emit ranked",
         "rescore: gather + dots", "rescore: emit ranked", "kernel: final squeeze (kth + compaction)", "-"]
dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
Q, N, D, k = 10_000, 49_688, 384, 100
g = torch.Generator(device="cuda").manual_seed(1234)
items = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1)
queries = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(dtype)
cat = icr.DeviceCatalog(items, dtype=dtype)
lib = _lib.load()
fn = lib.icr_debug_select_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
for _ in range(3):
    cat.topk(queries, k)
torch.cuda.synchronize()
fn(None, 1)
steps = 5
for _ in range(steps):
    cat.topk(queries, k)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 16)()
fn(out, 0)
print(f"cycles per query-warp, summed over the selects of one {dtype} C2 step (3 selects per step):")
for name, v in zip(NAMES, out):
    if v:
        print(f"  {name:45s} {v / steps / Q:10.0f}")
