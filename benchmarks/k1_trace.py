"""Per-CTA timeline of the K1 ring kernel (needs a build with ICR_NVCC_DEFS=-DICR_TRACE). Development aid."""
import ctypes
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import instacart_next_order_recommendation_b200 as icr  # noqa: E402
from instacart_next_order_recommendation_b200 import _lib  # noqa: E402

N, D, Q, k = 49_688, 384, int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = torch.Generator(device="cuda").manual_seed(0)
cat = icr.DeviceCatalog(torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1))
q = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
# ICR_K1_ROT=1: instead of flushing L2 with a 512 MB fill (which leaves L2 full of DIRTY lines that the catalog reads then
# evict, and the kernel's own instructions in DRAM), rotate over catalog copies that together exceed L2
ROT = os.environ.get("ICR_K1_ROT") == "1"
copies = [cat] + ([icr.DeviceCatalog(cat.rows.clone()) for _ in range(3)] if ROT else [])
if ROT:
    for _ in range(3):
        for c in copies:
            c.topk(q, k)
    torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * (148 * 8))()
names = ["entry", "bar_init", "queries", "first_slab", "stream_done", "list_emitted", "ticket", "exit"]
DBG = [int(x) for x in os.environ.get("ICR_K1_DBGS", "0").split(",")]
for it in range(4 * len(DBG)):
    # trace builds read ICR_K1_DBG at every launch: 1 = no arithmetic, 2 = static row blocks, 4 = 8-slot ring
    os.environ["ICR_K1_DBG"] = str(DBG[it // 4])
    if it % 4 == 0:
        print(f"=== ICR_K1_DBG={DBG[it // 4]}")
    if not ROT:
        flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    copies[it % len(copies)].topk(q, k)
    e1.record()
    torch.cuda.synchronize()
    lib.icr_debug_read_trace(buf, 148 * 8)
    t = np.frombuffer(buf, dtype=np.uint64).reshape(148, 8).astype(np.int64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1000.0
    print(f"iter {it}: events {e0.elapsed_time(e1) * 1000:.1f} us; kernel span {rel.max():.1f} us")
    mb = (ctypes.c_ulonglong * 8)()
    lib.icr_debug_read_merge_trace(mb)
    m = (np.frombuffer(mb, dtype=np.uint64).astype(np.int64) - t0) / 1000.0
    print("   merge: start %.2f floor %.2f survivors %.2f ranked %.2f done %.2f" % tuple(m[:5]))
    for i, nm in enumerate(names):
        print(f"   {nm:13s} min {rel[:, i].min():6.2f}  median {np.median(rel[:, i]):6.2f}  max {rel[:, i].max():6.2f}")
