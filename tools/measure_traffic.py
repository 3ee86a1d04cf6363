"""Measures `roofline.traffic` for bench.py: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of
the projection-GEMM kernels of one forward of the bench workload, from an ncu pass over this very build.

    python tools/measure_traffic.py            # on a B200 box; writes profiles/roofline_traffic.json

dram__bytes cannot be read outside a profiler, so bench.py does not measure it live: it prints the number stored
here ONLY when the stored `source_digest` equals the digest of the kernel sources it is running (build.py
_source_digest), and `traffic: null` with the reason otherwise.

Child mode (`--child PRECISION`) is what ncu profiles: two eager forwards of BASELINE configs[1] (batch 4096).
"""
from __future__ import annotations

import argparse
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

KERNELS = "regex:gemm_tcgen05_kernel|qkv_attention_kernel"


def child(precision: str, batch: int) -> None:
    import torch
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    model = stlt_b200.Stlt(cfg, precision=precision)
    model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
    model = model.to("cuda")
    model.train(False)
    data = make_batch(batch, "something", ragged=False, seed=100)
    dev = {k: data[k].cuda() for k in ("categories", "boxes", "frame_types", "lengths")}
    with torch.no_grad():
        for _ in range(2):
            model(dev)
    torch.cuda.synchronize()
    print("LAUNCHES_PER_FORWARD", model.last_launch_count())


def parse(log: Path):
    lines = [ln for ln in log.read_text().splitlines() if ln.startswith('"')]
    rows = {}
    for row in csv.DictReader(lines):
        rows.setdefault((int(row["ID"]), row["Kernel Name"]), {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    return [(k[1], v) for k, v in sorted(rows.items())]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", default=None)
    ap.add_argument("--batch", type=int, default=4096)
    args = ap.parse_args()
    if args.child:
        child(args.child, args.batch)
        return
    import __graft_entry__
    __graft_entry__.build()
    digest = __graft_entry__._load_build_module()._source_digest()
    out = {"source_digest": digest, "unit": "bytes per GEMM launch (dram__bytes_read.sum + dram__bytes_write.sum)",
           "command": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                      f"-k {KERNELS} python tools/measure_traffic.py --child <precision> (second forward, batch {args.batch})"}
    scratch = ROOT / "gpurun_out"
    scratch.mkdir(exist_ok=True)
    for precision in ("bf16", "fp32"):
        log = scratch / f"traffic_{precision}.csv"
        cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
               "--clock-control", "none", "-k", KERNELS, "--csv", "--log-file", str(log), sys.executable, __file__,
               "--child", precision, "--batch", str(args.batch)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise SystemExit(f"ncu failed: {res.stdout[-1000:]} {res.stderr[-2000:]}")
        launches = parse(log)
        per_forward = len(launches) // 2
        last = launches[per_forward:]
        total = sum(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for _, v in last)
        out[precision] = total / max(len(last), 1)
        out[f"{precision}_detail"] = {
            "gemm_launches_per_forward": len(last), "dram_bytes_per_forward_gemm_kernels": total,
            "kernel_ms_sum_under_ncu": sum(v["gpu__time_duration.sum"] for _, v in last) / 1e6,
        }
    path = ROOT / "profiles" / "roofline_traffic.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
