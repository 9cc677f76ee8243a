#!/usr/bin/env python
"""Summarise ncu output into small text files for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv  > profiles/X_launch_summary.csv
    python tools/ncu_summary.py full     gpurun_out/X.ncu-rep       > profiles/X_full_summary.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
        name = r[ki].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"<.*", "", name.split("(")[0]).replace("void ", "").strip()
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v * scale
        a[1] += 1
        total += v * scale
    print(f"# gpu__time_duration.sum per kernel family from {path}; serialised cold-cache times -> compare SHARES")
    print(f"# total {total:.1f} ms over {sum(a[1] for a in agg.values())} launches")
    print("ms,share,launches,kernel")
    for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:.2f},{ms / total:.3f},{n},{name}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:120])
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m} = {r[i]} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
