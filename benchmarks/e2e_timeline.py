"""Where the time of a host-to-host C2 call goes: per-piece event timeline of DeviceCatalog.topk_host's pipeline.

Re-states the loop of index.py:topk_host with an event after each stage (upload, kernels, download) and prints, for a few
splits, when each stage ended relative to the start of the call. Also prints the bare copy rates of this box.
"""
import sys, torch
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr
from instacart_next_order_recommendation_b200 import ops
Q, N, D, k = 10000, 49688, 384, 100
g = torch.Generator(device="cuda").manual_seed(0)
items = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1)
cat = icr.DeviceCatalog(items)
qh = torch.nn.functional.normalize(torch.randn(Q, D, generator=torch.Generator().manual_seed(1)), dim=1).pin_memory()
ov = torch.empty(Q, k).pin_memory(); oi = torch.empty(Q, k, dtype=torch.int64).pin_memory()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def ev():
    return torch.cuda.Event(enable_timing=True)

def copy_rates():
    qd = torch.empty(Q, D, device="cuda"); vd = torch.empty(Q, k, device="cuda"); idd = torch.empty(Q, k, dtype=torch.int64, device="cuda")
    for name, fn, nbytes in (("h2d queries", lambda: qd.copy_(qh, non_blocking=True), qh.numel() * 4),
                             ("d2h values+ids", lambda: (ov.copy_(vd, non_blocking=True), oi.copy_(idd, non_blocking=True)), Q * k * 12)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"{name}: {nbytes / 1e6:.2f} MB in {ms * 1e3:.0f} us = {nbytes / ms / 1e6:.1f} GB/s", flush=True)

def run(splits, record):
    cur = torch.cuda.current_stream()
    edges = [0]
    for n in splits: edges.append(min(Q, edges[-1] + n))
    edges[-1] = Q
    start = ev(); start.record(cur)
    marks = []
    for c, (lo, hi) in enumerate(zip(edges[:-1], edges[1:])):
        st = streams[c % 2]
        st.wait_event(start)
        with torch.cuda.stream(st):
            e0 = ev(); e1 = ev(); e2 = ev(); e3 = ev()
            if record: e0.record(st)
            qd = qh[lo:hi].to("cuda", non_blocking=True)
            if record: e1.record(st)
            v, i = ops.cos_topk(qd, cat.rows, k, cat_planes=cat.planes, cat_inv_norms=cat.inv_norms)
            if record: e2.record(st)
            ov[lo:hi].copy_(v, non_blocking=True); oi[lo:hi].copy_(i, non_blocking=True)
            if record: e3.record(st)
            marks.append((hi - lo, e0, e1, e2, e3))
    for st in streams: cur.wait_stream(st)
    end = ev(); end.record(cur)
    return start, marks, end

copy_rates()
for splits in ([10000], [5000, 5000], [7000, 3000], [3000, 7000], [2000, 6000, 2000], [1536, 6928, 1536], [2000, 5000, 3000], [4000, 4000, 2000],
               [1000, 6000, 3000], [6000, 3000, 1000], [5000, 3000, 2000], [3000, 3000, 2000, 2000]):
    for _ in range(3): run(splits, False)
    torch.cuda.synchronize()
    tot = []
    for _ in range(7):
        flush.zero_(); torch.cuda.synchronize()
        s, m, e = run(splits, False); torch.cuda.synchronize(); tot.append(s.elapsed_time(e))
    tot.sort()
    flush.zero_(); torch.cuda.synchronize()
    s, marks, e = run(splits, True); torch.cuda.synchronize()
    print(f"splits {splits}: median {tot[3]:.3f} ms min {tot[0]:.3f} (with events {s.elapsed_time(e):.3f})", flush=True)
    for n, e0, e1, e2, e3 in marks:
        print(f"   piece {n:5d}: begins {s.elapsed_time(e0) * 1e3:5.0f}  uploaded {s.elapsed_time(e1) * 1e3:5.0f}  ranked {s.elapsed_time(e2) * 1e3:5.0f}  downloaded {s.elapsed_time(e3) * 1e3:5.0f} us", flush=True)
