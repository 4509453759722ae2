"""Summarise bench.py JSON lines and an ncu launch list: python benchmarks/show_bench.py a.json b.json [--launches x.csv [first [count]]]"""
import csv
import json
import sys

args = sys.argv[1:]
launches = None
if "--launches" in args:
    i = args.index("--launches")
    launches = args[i + 1 :]
    args = args[:i]
for f in args:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{f}: {d['value']:.4g} q/s  {d['ms_per_step']:.4f} ms/step (min {d['ms_per_step_min']:.4f})  e2e {d['e2e']['value']:.4g}  launches/step {d['gpu_launches'] / d['steps']:.0f}")
    print(f"   kernel {r['kernel']} {r['kernel_ms_per_step']:.4f} ms/step  {r['achieved']:.0f} {r['unit']}  frac {r['frac']:.3f}   ir_eval {d.get('ir_eval', {}).get('ms_per_eval')}")
    for key in ("c1", "c3"):
        if key in d:
            print("   ", key, json.dumps(d[key])[:300])
if launches:
    rows = list(csv.reader(open(launches[0])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    first = int(launches[1]) if len(launches) > 1 else 0
    count = int(launches[2]) if len(launches) > 2 else 40
    tot = 0.0
    for r in rows[hdr + 1 + first : hdr + 1 + first + count]:
        us = float(r[vi].replace(",", "")) / 1e3
        tot += us
        print(f"   {us:9.1f} us  {r[ki][:90]}")
    print(f"   {tot:9.1f} us  total")
