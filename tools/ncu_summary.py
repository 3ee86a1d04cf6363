#!/usr/bin/env python
"""Turns an ncu report into the markdown tables kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_train_r1_v6.ncu-rep            # --set full report -> table
    python tools/ncu_summary.py --launch-list gpurun_out/launches_r1_v7.csv    # per-kernel totals of a launch list

Needs the `ncu` CLI (reads the report with `ncu -i ... --page raw --csv`).
"""
from __future__ import annotations

import collections
import csv
import re
import subprocess
import sys

METRICS = [
    ("ms", "gpu__time_duration.sum", 1e-3, "us"),
    ("SM GHz", "sm__cycles_elapsed.avg.per_second", 1.0, None),
    ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0, None),
    ("DRAM read GB", "dram__bytes_read.sum", 1.0, None),
    ("DRAM write GB", "dram__bytes_write.sum", 1.0, None),
    ("DRAM % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0, None),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0, None),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0, None),
    ("issue active %", "smsp__issue_active.avg.pct", 1.0, None),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0, None),
    ("regs/thread", "launch__registers_per_thread", 1.0, None),
    ("grid", "launch__grid_size", 1.0, None),
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("stlt::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()


def to_unit(value: str, unit: str, want: str) -> float:
    v = float(value.replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    if want == "ms":
        return v * scale.get(unit, 1.0)
    gb = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
    if unit in gb:
        return v * gb[unit]
    if unit.endswith("hz"):
        return v * {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0}.get(unit, 1.0)
    return v


def full_report(path: str) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = rows[2:]
    print("| metric | " + " | ".join(f"{i}: {short(r[idx['Kernel Name']])[:44]}" for i, r in enumerate(kernels)) + " |")
    print("|---|" + "---|" * len(kernels))
    for label, metric, _, _ in METRICS:
        if metric not in idx:
            continue
        cells = []
        for r in kernels:
            want = "ms" if label == "ms" else None
            cells.append(f"{to_unit(r[idx[metric]], units[idx[metric]], want):.4g}")
        print(f"| {label} | " + " | ".join(cells) + " |")


def launch_list(path: str) -> None:
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = None
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    count = collections.Counter()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = short(d["Kernel Name"])
        metric = d["Metric Name"]
        want = "ms" if metric.startswith("gpu__time") else None
        per[name][metric] += to_unit(d["Metric Value"], d["Metric Unit"], want)
        if metric.startswith("gpu__time"):
            count[name] += 1
    total = sum(v["gpu__time_duration.sum"] for v in per.values())
    print("| kernel | launches | ms | share | DRAM read GB | DRAM write GB |")
    print("|---|---|---|---|---|---|")
    for name, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        print(f"| {name[:70]} | {count[name]} | {v['gpu__time_duration.sum']:.3f} | "
              f"{100 * v['gpu__time_duration.sum'] / total:.1f} % | {v.get('dram__bytes_read.sum', 0):.2f} | "
              f"{v.get('dram__bytes_write.sum', 0):.2f} |")
    print(f"\ntotal {total:.3f} ms over {sum(count.values())} launches")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--launch-list":
        launch_list(sys.argv[2])
    else:
        full_report(sys.argv[1])
