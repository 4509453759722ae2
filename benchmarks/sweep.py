"""Measure every BASELINE.json config (or its per-GPU shard) on ONE B200 and write a markdown table.

    python benchmarks/sweep.py [--out gpurun_out/sweep.md] [--quick]

Not the judged bench (that is ../bench.py on config C2); this is the parameter sweep behind DESIGN.md §5 and
profiles/: C1 batch-1 HBM roofline, C2 fp32/bf16, C3 MNRL vs PyTorch eager, C4/C5 shards, plus the
PyTorch-eager-on-the-same-GPU reference point ("what you get for free").
"""

from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import instacart_next_order_recommendation_b200 as icr  # noqa: E402
from instacart_next_order_recommendation_b200 import ops  # noqa: E402

PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TC = PEAKS["hbm_gbs"] * 1e9, PEAKS["bf16_tflops"] * 1e12
DEV = torch.device("cuda", 0)
FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
    FLUSH.zero_()


def time_gpu(fn, iters=20, warmup=5, cold=True):
    for _ in range(warmup):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    torch.cuda.synchronize()
    for a, b in evs:
        if cold:
            flush_l2()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return ms[len(ms) // 2], ms[0]


def unit_rows(n, d, dtype, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    out = torch.empty(n, d, dtype=dtype, device=DEV)
    step = 1 << 20
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = torch.nn.functional.normalize(torch.randn(e - s, d, device=DEV, generator=g), dim=1).to(dtype)
    return out


ROWS = []


def record(name, **kw):
    kw = {"config": name, **kw}
    ROWS.append(kw)
    print(json.dumps(kw), flush=True)


def topk_case(name, N, D, Q, k, dtype, path=ops.PATH_AUTO, iters=20, eager=False, cold=True, graph=False):
    cat_rows = unit_rows(N, D, dtype, 1234)
    cat = icr.DeviceCatalog(cat_rows, dtype=dtype)
    q = unit_rows(Q, D, dtype, 4321)
    fn = (lambda: cat.topk_small(q, k, path=path, copy=False)) if graph else (lambda: cat.topk(q, k, path=path))  # noqa: E731
    med, best = time_gpu(fn, iters=iters, cold=cold)
    kt = ops.kernel_timing(fn, max(3, iters // 4), flush=FLUSH if cold else None)
    esz = 4 if dtype == torch.float32 else 2
    bytes_ = N * D * esz
    flops = 2.0 * Q * N * D
    floor = max(bytes_ / HBM, flops / TC)
    row = dict(N=N, D=D, Q=Q, k=k, dtype=str(dtype).split(".")[-1], kernel=kt["kernel"], ms=med, ms_min=best, qps=Q / (med * 1e-3),
               kernel_ms=kt["ms_per_step"], kernel_launches=kt["launches_per_step"], bound="hbm" if bytes_ / HBM >= flops / TC else "tensor",
               roofline_frac_call=floor / (med * 1e-3), roofline_frac_kernel=floor / (kt["ms_per_step"] * 1e-3) if kt["ms_per_step"] else None,
               hbm_gbs_kernel=bytes_ / (kt["ms_per_step"] * 1e-3) / 1e9 if kt["ms_per_step"] else None,
               tflops_kernel=flops / (kt["ms_per_step"] * 1e-3) / 1e12 if kt["ms_per_step"] else None, l2="flushed" if cold else "warm")
    if eager:
        cf = cat_rows.float()
        qf = q.float()

        def eager_fn():
            s = torch.nn.functional.normalize(qf, dim=1) @ torch.nn.functional.normalize(cf, dim=1).T
            return torch.topk(s, k, dim=1)

        try:
            emed, _ = time_gpu(eager_fn, iters=max(5, iters // 2), cold=cold)
            row["torch_eager_ms"] = emed
        except torch.OutOfMemoryError:
            row["torch_eager_ms"] = None
    record(name, **row)
    del cat, cat_rows, q
    torch.cuda.empty_cache()


def c1_rotating_case(dtype, k=10, copies=6, iters=60):
    """C1 batch-1 with the catalog HBM-cold but the L2 CLEAN: rotate through `copies` replicas of the catalog (6 x 76 MB
    fp32 >> the 126 MB L2) instead of zeroing a flush buffer, whose dirty lines have to be written back while the next
    request streams in (the flush variant charges that write traffic to the request)."""
    N, D = 49_688, 384
    rows = unit_rows(N, D, dtype, 1234)
    cats = [icr.DeviceCatalog(rows.clone(), dtype=dtype) for _ in range(copies)]
    q = unit_rows(1, D, dtype, 4321)
    state = {"i": 0}

    def fn():
        c = cats[state["i"] % copies]
        state["i"] += 1
        return c.topk(q, k)

    med, best = time_gpu(fn, iters=iters, warmup=2 * copies, cold=False)
    esz = 4 if dtype == torch.float32 else 2
    floor = N * D * esz / HBM
    record(f"C1 batch-1 top-{k} (HBM-cold, clean L2: {copies} rotating catalog copies)", N=N, D=D, Q=1, k=k, dtype=str(dtype).split(".")[-1],
           kernel="gemv_topk", ms=med, ms_min=best, qps=1 / (med * 1e-3), kernel_ms=0.0, kernel_launches=1, bound="hbm",
           roofline_frac_call=floor / (med * 1e-3), roofline_frac_kernel=None, hbm_gbs_kernel=None, tflops_kernel=None, l2="cold, clean")
    del cats
    torch.cuda.empty_cache()


def cpu_scaled_case(name, N_full, D, Q, k, N_sample=1_000_000):
    """The reference's CPU path (oracle port: cos_sim + torch.topk) on a down-scaled catalog, extrapolated linearly in N
    (C4 / C5 do not fit a CPU run at full size)."""
    import os

    from oracle import oracle

    items = oracle.synth_isotropic(N_sample, D, 1234)
    queries = oracle.synth_isotropic(Q, D, 4321)
    oracle.cos_topk(queries[: min(Q, 8)], items[:10000], k)
    t0 = time.perf_counter()
    oracle.cos_topk(queries, items, k, sorted=False)
    dt = time.perf_counter() - t0
    scale = N_full / N_sample
    record(name, cpu_ms_sample=dt * 1e3, N_sample=N_sample, N_full=N_full, D=D, Q_cpu=Q, k=k, cpu_ms_extrapolated=dt * 1e3 * scale,
           cpu_qps_extrapolated=Q / (dt * scale), threads=torch.get_num_threads(), host_cpus=os.cpu_count())


def mnrl_case(B, D, scale, dtype):
    g = torch.Generator(device=DEV).manual_seed(2024)
    a = torch.randn(B, D, device=DEV, generator=g).to(dtype).requires_grad_(True)
    p = torch.randn(B, D, device=DEV, generator=g).to(dtype).requires_grad_(True)

    def ours():
        a.grad = p.grad = None
        icr.mnrl_loss(a, p, scale).backward()

    def eager():
        a.grad = p.grad = None
        s = torch.nn.functional.normalize(a.float(), dim=1) @ torch.nn.functional.normalize(p.float(), dim=1).T * scale
        torch.nn.functional.cross_entropy(s, torch.arange(B, device=DEV)).backward()

    m1, b1 = time_gpu(ours, iters=50, warmup=10, cold=False)
    m2, b2 = time_gpu(eager, iters=50, warmup=10, cold=False)
    record(f"C3 MNRL fwd+bwd B={B}", B=B, D=D, scale=scale, dtype=str(dtype).split(".")[-1], ours_us=m1 * 1e3, ours_min_us=b1 * 1e3,
           torch_eager_us=m2 * 1e3, torch_eager_min_us=b2 * 1e3, speedup=m2 / m1, our_kernels="2 + 1 memset (+ autograd glue)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "sweep.md"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    f32, bf16 = torch.float32, torch.bfloat16
    want = set(args.only.split(",")) if args.only else None

    def on(tag):
        return want is None or tag in want

    if on("c1"):
        for dt in (f32, bf16):
            for k in (10, 100):
                topk_case(f"C1 batch-1 top-{k}", 49_688, 384, 1, k, dt, iters=50, eager=(k == 10), cold=True)
            topk_case("C1 batch-1 top-10 (L2-warm)", 49_688, 384, 1, 10, dt, iters=50, cold=False)
            c1_rotating_case(dt)
            topk_case("C1 batch-1 top-10, CUDA graph", 49_688, 384, 1, 10, dt, iters=50, cold=True, graph=True)
            topk_case("C1 batch-1 top-10, CUDA graph (L2-warm)", 49_688, 384, 1, 10, dt, iters=50, cold=False, graph=True)
        for Q in (2, 4, 7, 8, 16, 32, 64, 128, 256):
            topk_case(f"C1-size batch-{Q} top-10", 49_688, 384, Q, 10, f32, iters=30)
        for Q in (16, 64, 256):  # the same calls replayed as one CUDA graph (DeviceCatalog.topk_small): launch gaps removed
            topk_case(f"C1-size batch-{Q} top-10, CUDA graph", 49_688, 384, Q, 10, f32, iters=30, graph=True)
    if on("c2"):
        topk_case("C2 IR eval fp32", 49_688, 384, 10_000, 100, f32, eager=True)
        topk_case("C2 IR eval bf16", 49_688, 384, 10_000, 100, bf16)
        topk_case("C2 real query count", 49_688, 384, 13_120, 100, f32)
    if on("c3"):
        for dt in (bf16, f32):
            for B in (64, 256, 1024) if args.quick else (64, 256, 1024, 4096, 8192):
                mnrl_case(B, 384, 20.0, dt)
        mnrl_case(256, 384, 30.0, f32)
    if on("c4"):
        n_shard = 1_250_000  # one of 8 shards of the 10M x 768 bf16 catalog
        for Q in ((1, 8, 64, 1024) if args.quick else (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)):
            topk_case(f"C4 shard (1/8 of 10M x 768) Q={Q}", n_shard, 768, Q, 100, bf16, iters=10)
        if not args.quick:
            for Q in (1, 64, 1024):
                topk_case(f"C4 whole 10M x 768 on one GPU Q={Q}", 10_000_000, 768, Q, 100, bf16, iters=5)
    if on("c5"):
        topk_case("C5 shard (1/8 of 100M x 384) Q=4096", 12_500_000, 384, 4096, 100, bf16, iters=5)

    if on("cpu"):
        cpu_scaled_case("C4 on the CPU (1M-row sample of 10M x 768, Q=64)", 10_000_000, 768, 64, 100)
        cpu_scaled_case("C4 on the CPU (1M-row sample of 10M x 768, Q=1)", 10_000_000, 768, 1, 100)
        cpu_scaled_case("C5 on the CPU (1M-row sample of 100M x 384, Q=256 of 4096)", 100_000_000, 384, 256, 100)

    lines = ["# sweep on one B200 (`benchmarks/sweep.py`)", "", f"peaks: HBM {PEAKS['hbm_gbs']} GB/s, bf16 {PEAKS['bf16_tflops']} TFLOP/s (MEASURED_PEAKS.json)", ""]
    tk = [r for r in ROWS if "qps" in r]
    if tk:
        lines += ["| config | N | D | Q | k | dtype | L2 | kernel | call ms (median) | queries/s | kernel ms | HBM GB/s (kernel) | TFLOP/s (kernel) | bound | roofline frac (kernel / whole call) | torch eager ms |",
                  "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
        def f(x, spec):
            return "-" if x is None else format(x, spec)

        for r in tk:
            lines.append(f"| {r['config']} | {r['N']} | {r['D']} | {r['Q']} | {r['k']} | {r['dtype']} | {r['l2']} | {r['kernel']} | {r['ms']:.4f} | {r['qps']:.4g} | "
                         f"{r['kernel_ms']:.4f} | {f(r['hbm_gbs_kernel'], '.0f')} | {f(r['tflops_kernel'], '.1f')} | {r['bound']} | "
                         f"{f(r['roofline_frac_kernel'], '.3f')} / {r['roofline_frac_call']:.3f} | {f(r.get('torch_eager_ms'), '.3f')} |")
    mn = [r for r in ROWS if "ours_us" in r]
    if mn:
        lines += ["", "| config | dtype | scale | ours µs (median / min) | torch eager µs (median / min) | speed-up |", "|---|---|---|---|---|---|"]
        for r in mn:
            lines.append(f"| {r['config']} | {r['dtype']} | {r['scale']} | {r['ours_us']:.1f} / {r['ours_min_us']:.1f} | {r['torch_eager_us']:.1f} / {r['torch_eager_min_us']:.1f} | {r['speedup']:.2f}x |")
    cp = [r for r in ROWS if "cpu_ms_sample" in r]
    if cp:
        lines += ["", "| config | threads | CPU ms on the sample | extrapolated to full N (ms) | extrapolated queries/s |", "|---|---|---|---|---|"]
        for r in cp:
            lines.append(f"| {r['config']} | {r['threads']} | {r['cpu_ms_sample']:.0f} | {r['cpu_ms_extrapolated']:.0f} | {r['cpu_qps_extrapolated']:.3g} |")
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text("\n".join(lines) + "\n")
    print("wrote", args.out)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"sweep took {time.time() - t0:.1f} s")
