"""Experiment: host-to-host top-k (C2) vs number of pipeline chunks."""
import sys, torch
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr
Q, N, D, k = 10000, 49688, 384, 100
g = torch.Generator(device="cuda").manual_seed(0)
items = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1)
cat = icr.DeviceCatalog(items)
qh = torch.nn.functional.normalize(torch.randn(Q, D, generator=torch.Generator().manual_seed(1)), dim=1).pin_memory()
ov = torch.empty(Q, k).pin_memory(); oi = torch.empty(Q, k, dtype=torch.int64).pin_memory()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
cases = [dict(n_chunks=n) for n in (1, 2, 3)] + [dict(splits=s) for s in (
    [2500, 7500], [3000, 7000], [3500, 6500], [4000, 6000], [7000, 3000], [1500, 7000, 1500], [2000, 6000, 2000], [2000, 5000, 3000],
    [2500, 5000, 2500], [1500, 6000, 2500], [3000, 5000, 2000], [1000, 3000, 6000], [2000, 3000, 3000, 2000], [1500, 3500, 3500, 1500])]
for kw in cases:
    nch = kw
    for _ in range(3): cat.topk_host(qh, k, out=(ov, oi), **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cat.topk_host(qh, k, out=(ov, oi), **kw); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"{nch}: median {ts[5]:.3f} ms  min {ts[0]:.3f} ms  -> {Q / ts[5] / 1e3:.2f} M q/s", flush=True)
