"""Print selected metrics of an `ncu --page raw --csv` dump, one column per profiled launch.

    ncu -i x.ncu-rep --page raw --csv > raw.csv ; python benchmarks/ncu_pick.py raw.csv [substring ...]
"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak", "smsp__issue_active.avg.pct", "registers_per_thread",
           "occupancy_limit", "issue_stalled", "bank_conflicts", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "lts__t_bytes.sum", "sm__pipe_tensor", "dram__throughput", "lts__throughput", "l1tex__throughput", "sm__throughput"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pats = sys.argv[2:] or DEFAULT
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("kernels:", [r[hdr.index("Kernel Name")][:40] for r in data])
    for i, name in enumerate(hdr):
        if any(p in name for p in pats):
            vals = [r[i] for r in data]
            if all(v in ("0", "", "n/a") for v in vals):
                continue
            print(f"{name} [{units[i]}]: {vals}")


if __name__ == "__main__":
    main()
