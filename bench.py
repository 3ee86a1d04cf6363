#!/usr/bin/env python
"""Headline benchmark: STLT inference throughput (videos/sec) on N B200s, batch-sharded.

    python bench.py --gpus N --steps K --warmup W            # this implementation (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path

One "step" is one forward of the hot path over one batch of synthetic layouts of the
Something-Else shape (BASELINE.json configs[1]: 16+1 frames x 5 slots, 174 classes, batch 4096 per
GPU, dense: every frame carries 4 boxes). Rank 0 prints ONE JSON line:
  value   : videos/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e     : videos/s through the public module call with pinned-host inputs (H2D + forward + D2H
            of the logits + a sync per step, as the reference's inference loop does)
  roofline: the tcgen05 projection GEMMs (99.8 % of the FLOPs): executed FLOPs per step / their
            summed CUDA-event time (events around every launch, in a second timed pass of the same
            K steps right after the first), vs the measured bf16 peak
  cpu_baseline: the CPU oracle (a PyTorch-CPU restatement of the reference forward) on this box's
            host cores, bounded sample.
For N > 1 launch with torchrun (one process per GPU); inference needs no collective — every rank
runs its own batch (weak scaling) and only the timing is reduced (MAX) over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

FLOPS_PER_VIDEO = {"something": 6_752_443_392, "action_genome": 12_548_941_824}
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=4096, help="videos per GPU per step")
    p.add_argument("--layout", default="something", choices=["something", "action_genome"])
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"],
                   help="headline precision; the other one is reported under 'secondary'")
    p.add_argument("--no-secondary", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--workload", default="inference", choices=["inference", "train", "cacnf"],
                   help="'train' = BASELINE configs[3]: fwd + bwd + clip + AdamW, NCCL gradient all-reduce for N > 1")
    p.add_argument("--train-batch", type=int, default=2048, help="videos per GPU per training step")
    p.add_argument("--dropout", type=float, default=None, help="training dropout (default: the reference's 0.1)")
    return p.parse_args()


def load_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            d = json.loads(path.read_text())
            return d, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(sm_max) if sm_max else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_throughput(layout: str, batch: int, seconds: float, warmup: int = 1, steps: int | None = None):
    """videos/s of the CPU oracle (PyTorch-CPU restatement of the reference forward)."""
    import torch
    import stlt_b200
    from oracle import stlt_oracle
    from stlt_b200.synthetic import make_batch, random_state_dict
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1)
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
    torch.manual_seed(0)
    sd = random_state_dict(stlt_b200.Stlt(cfg).state_dict(), seed=0)
    data = make_batch(batch, layout, ragged=False, seed=0)
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            stlt_oracle.stlt_forward(sd, data)
        t_end = time.perf_counter() + seconds
        n = 0
        while (steps is not None and n < steps) or (steps is None and (time.perf_counter() < t_end or n < 2)):
            t0 = time.perf_counter()
            stlt_oracle.stlt_forward(sd, data)
            times.append(time.perf_counter() - t0)
            n += 1
    total = sum(times)
    return {"value": batch * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "steps": len(times), "cores": torch.get_num_threads(), "host_cpus": os.cpu_count()}


def inference_config(layout: str, batch: int, world: int, L: int, S: int, num_classes: int) -> dict:
    """`config` of the inference workload; the reference arm reports the same one (it times a bounded sample of it)."""
    return {
        "workload": f"STLT inference, {layout} shape (L={L} frames x S={S} slots, "
                    f"{num_classes} classes), batch {batch} per GPU, dense layouts, random-init weights",
        "global_batch": batch * world, "parallelism": f"batch-sharded x{world}, no collective",
        "l2_policy": "activations (GBs per step) far exceed the 126 MB L2; no explicit flush",
    }


def run_reference(args, rank: int):
    """Reference arm: the reference's CPU implementation of the path, timed on the host cores.
    The reference is a Python/PyTorch repo that is not present on the GPU box, so this runs the
    oracle port (oracle/stlt_oracle.py, pinned to the reference's outputs by tests/golden)."""
    if rank != 0:
        return
    sample = 64
    r = cpu_oracle_throughput(args.layout, sample, seconds=0, warmup=max(args.warmup, 1), steps=args.steps)
    line = {
        "impl": "reference", "metric": "stlt_inference_videos_per_sec", "value": r["value"], "unit": "videos/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same config as the B200 arm; each step is a bounded sample of that workload on the host cores
        "config": inference_config(args.layout, args.batch, args.gpus, 17, 5 if args.layout == "something" else 11,
                                   174 if args.layout == "something" else 157),
        "cpu_baseline": {"value": r["value"], "unit": "videos/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{r['steps']} CPU fp32 forwards of {sample} videos of that workload (dense layouts) "
                                   f"through oracle/stlt_oracle.py on rank 0, host cpus {r['host_cpus']}"},
        "e2e": {"value": r["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_resident(model, batch_dev, steps, warmup, world, torch, dist, profile=False):
    """Device-resident timing: K forwards bracketed by barrier + synchronize, CUDA events.
    With profile=True the library also brackets every launch with CUDA events (per-category times);
    that pass is used for the roofline / breakdown only, because ~180 extra event records per step
    cost a few percent of throughput."""
    with torch.no_grad():
        for _ in range(warmup):
            model(batch_dev)
        torch.cuda.synchronize()
        if profile:
            model.set_profiling(True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            model(batch_dev)
        end.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = start.elapsed_time(end)
        prof = None
        if profile:
            prof = model.get_profile()
            model.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, prof


def time_e2e(model, batch_host, batch_dev, logits_host, steps, warmup, world, torch, dist, output_key="stlt"):
    """Public-API timing with HOST inputs. Every step uploads its inputs from pinned host memory and
    delivers its logits to host memory; stlt_b200.pipeline.HostPipeline (part of the package's public
    API) overlaps the copies of neighbouring steps with compute, the way a DataLoader with pinned
    memory feeds the reference loop. The timed region ends when the last logits are on the host."""
    from stlt_b200.pipeline import HostPipeline
    pipe = HostPipeline(model, output_key)
    checksum = 0.0

    def run(n):
        nonlocal checksum
        for host_logits in pipe.run(batch_host for _ in range(n)):
            checksum += float(host_logits[0, 0])  # touch the delivered result on the host

    with torch.no_grad():
        run(warmup)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        run(steps)
        end.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = start.elapsed_time(end)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def roofline_from_profile(prof, steps, precision, peaks, peak_src):
    gemm = prof["gemm"]
    algorithmic = gemm["flops"] / max(steps, 1)                   # 2*M*N*K of the launches actually issued
    executed = algorithmic * (3.0 if precision == "fp32" else 1.0)  # fp32 mode: 3 bf16 MMAs per algorithmic MMA
    ms = gemm["ms"] / max(steps, 1)
    achieved = algorithmic / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
    traffic = None
    tpath = ROOT / "profiles" / "roofline_traffic.json"
    if tpath.exists():
        try:
            traffic = json.loads(tpath.read_text()).get(precision)
        except Exception:
            traffic = None
    out = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic,
        "kernel": "gemm_tcgen05_kernel (all projection GEMMs of a step" +
                  ("; in bf16 mode their epilogues also carry the residual adds and LayerNorms)" if precision == "bf16" else ")"),
        "launches_per_step": gemm["launches"] / max(steps, 1), "kernel_ms_per_step": ms,
        "algorithmic_flops_per_step": algorithmic,
        "peak_source": f"bf16 dense sustained, {peak_src}",
    }
    if precision == "fp32":
        out["mma_tflops_issued"] = executed / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        out["note"] = ("fp32-parity mode issues 3 bf16 MMAs (hi*hi + lo*hi + hi*lo) per algorithmic MMA; "
                       "mma_tflops_issued / peak is the tensor-pipe fraction, frac is in algorithmic FLOPs")
    return out


def cpu_oracle_train_throughput(layout: str, batch: int, steps: int):
    """videos/s of the CPU oracle training step (autograd through the restated forward + AdamW)."""
    import torch
    import stlt_b200
    from oracle import stlt_oracle
    from stlt_b200.synthetic import make_batch, random_state_dict
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
    torch.manual_seed(0)
    sd = random_state_dict(stlt_b200.Stlt(cfg).state_dict(), seed=0)
    data = make_batch(batch, layout, ragged=False, seed=0)
    labels = torch.randint(0, spec["num_classes"], (batch,)) if layout == "something" else \
        (torch.rand(batch, spec["num_classes"]) < 0.05).float()
    loss = "cross_entropy" if layout == "something" else "bce_with_logits"
    state, times = {}, []
    for step in range(1, steps + 2):
        t0 = time.perf_counter()
        _, _, grads = stlt_oracle.loss_and_grads(sd, data, labels, loss)
        stlt_oracle.adamw_update(sd, grads, state, step, 5e-5)
        if step > 1:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * len(times) / total, "ms_per_step": 1e3 * total / len(times), "steps": len(times),
            "cores": torch.get_num_threads(), "host_cpus": os.cpu_count()}


def run_train(args, rank, local_rank, world, torch, dist):
    """BASELINE configs[3]: one optimisation step = forward (activations kept) + criterion + backward +
    gradient all-reduce (NCCL, N > 1) + global-norm clip + AdamW + bf16 re-pack, per-GPU batch fixed."""
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    from stlt_b200.training import FusedTrainStep, linear_schedule_with_warmup
    spec = stlt_b200.SOMETHING_ELSE if args.layout == "something" else stlt_b200.ACTION_GENOME
    kw = {} if args.dropout is None else {"hidden_dropout_prob": args.dropout}
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"], **kw)
    torch.manual_seed(0)
    model = stlt_b200.Stlt(cfg, precision="bf16")
    model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
    model = model.to("cuda")
    model.train(True)
    B = args.train_batch
    keys = ["categories", "boxes", "frame_types", "lengths"] + (["scores"] if spec["scores"] else [])
    full = make_batch(B, args.layout, ragged=False, seed=100 + rank)
    g = torch.Generator().manual_seed(7 + rank)
    if args.layout == "something":
        full["labels"] = torch.randint(0, spec["num_classes"], (B,), generator=g)
        loss = "cross_entropy"
    else:
        full["labels"] = (torch.rand((B, spec["num_classes"]), generator=g) < 0.05).float()
        loss = "bce_with_logits"
    keys.append("labels")
    batch_host = {k: full[k].pin_memory() for k in keys}
    batch_dev = {k: v.cuda() for k, v in batch_host.items()}
    stepper = FusedTrainStep(model, lr=5e-5, weight_decay=1e-3, clip_val=5.0, loss=loss,
                             lr_lambda=linear_schedule_with_warmup(100, 100000))
    loss_host = torch.zeros(1).pin_memory()
    peaks, peak_src = load_peaks()
    _, L, S = full["categories"].shape

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if profile:
            model.set_profiling(True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            fn()
        end.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = start.elapsed_time(end)
        prof = None
        if profile:
            prof = model.get_profile()
            model.set_profiling(False)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, prof

    def step_resident():
        stepper.step(batch_dev)

    def step_e2e():
        for k, v in batch_host.items():
            batch_dev[k].copy_(v, non_blocking=True)
        loss_host.copy_(stepper.step(batch_dev).reshape(1), non_blocking=True)
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, _ = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    prof_ms, prof = timed(step_resident, args.steps, 1, profile=True)
    e2e_ms, _ = timed(step_e2e, args.steps, max(args.warmup, 1))
    videos = B * world * args.steps
    gemm = prof["gemm"]
    gemm_ms = gemm["ms"] / args.steps
    gemm_flops = gemm["flops"] / args.steps
    peak = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_oracle_train_throughput(args.layout, 8, steps=3)
        cpu = {"value": r["value"], "unit": "videos/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['steps']} training steps (autograd + AdamW) of batch 8, oracle/stlt_oracle.py on "
                         f"{r['cores']} torch threads (host cpus {r['host_cpus']})"}
    if rank == 0:
        h2d = sum(v.numel() * v.element_size() for v in batch_host.values())
        line = {
            "metric": "stlt_train_videos_per_sec", "value": videos / (ms * 1e-3), "unit": "videos/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (fp32 master weights / grads / AdamW)",
            "data": "synthetic",
            "config": {
                "workload": f"STLT training step (fwd + bwd + clip + AdamW), {args.layout} shape (L={L} x S={S}, "
                            f"{spec['num_classes']} classes), batch {B} per GPU, dense layouts, dropout {stepper.dropout_p}",
                "global_batch": B * world,
                "parallelism": f"data-parallel x{world}" + (", NCCL gradient all-reduce in 2 buckets, first overlapped with the spatial backward" if world > 1 else ""),
                "l2_policy": "activations (GBs per step) far exceed the 126 MB L2; no explicit flush",
            },
            "e2e": {"value": videos / (e2e_ms * 1e-3), "unit": "videos/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(sum(v["launches"] for v in prof.values())),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "gemm_tcgen05_kernel (forward, data-gradient and weight-gradient GEMMs of a step)",
                         "launches_per_step": gemm["launches"] / args.steps, "kernel_ms_per_step": gemm_ms,
                         "algorithmic_flops_per_step": gemm_flops, "peak_source": f"bf16 dense sustained, {peak_src}"},
            "cpu_baseline": cpu, "clocks": clocks,
            "model_tflops": 3 * FLOPS_PER_VIDEO[args.layout] * B / (ms / args.steps * 1e-3) / 1e12,
            "breakdown_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            "profiled_pass_ms_per_step": prof_ms / args.steps,
        }
        print(json.dumps(line), flush=True)


def run_cacnf(args, rank, local_rank, world, torch, dist):
    """BASELINE configs[4]: CACNF inference on precomputed per-clip ResNet3D features, batch-sharded."""
    import stlt_b200
    from stlt_b200.synthetic import make_appearance_features, make_batch, random_state_dict
    cfg = stlt_b200.CacnfModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    model = stlt_b200.Cacnf(cfg, precision=args.dtype)
    model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
    model = model.to("cuda")
    model.train(False)
    B = args.batch
    full = make_batch(B, "something", ragged=False, seed=100 + rank)
    keys = ["categories", "boxes", "frame_types", "lengths"]
    batch_host = {k: full[k].pin_memory() for k in keys}
    batch_host["video_features"] = make_appearance_features(B, seed=200 + rank).flatten(2).contiguous().pin_memory()
    batch_dev = {k: v.cuda() for k, v in batch_host.items()}
    out_host = torch.empty((B, 174), dtype=torch.float32).pin_memory()
    peaks, peak_src = load_peaks()

    def timed(fn, steps, warmup, profile=False):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            if profile:
                model.set_profiling(True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            for _ in range(steps):
                fn()
            end.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = start.elapsed_time(end)
            prof = None
            if profile:
                prof = model.get_profile()
                model.set_profiling(False)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, prof

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, _ = timed(lambda: model(batch_dev), args.steps, args.warmup)
    clocks = sampler.stop()
    prof_ms, prof = timed(lambda: model(batch_dev), args.steps, 1, profile=True)
    e2e_ms = time_e2e(model, batch_host, batch_dev, out_host, args.steps, max(args.warmup, 1), world, torch, dist,
                      output_key="ensemble")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import stlt_oracle
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        small = make_batch(8, "something", ragged=False, seed=0)
        feats = make_appearance_features(8, seed=1)
        with torch.no_grad():
            stlt_oracle.cacnf_forward(sd, small, feats)
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < args.cpu_seconds or n < 2:
                stlt_oracle.cacnf_forward(sd, small, feats)
                n += 1
            dt = time.perf_counter() - t0
        cpu = {"value": 8 * n / dt, "unit": "videos/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n} CACNF forwards of batch 8 in {dt:.0f} s, oracle/stlt_oracle.py (ResNet trunk excluded on both sides)"}
    if rank == 0:
        videos = B * world * args.steps
        gemm = prof["gemm"]
        gemm_ms, gemm_flops = gemm["ms"] / args.steps, gemm["flops"] / args.steps
        peak = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        h2d = sum(v.numel() * v.element_size() for v in batch_host.values())
        line = {
            "metric": "cacnf_inference_videos_per_sec", "value": videos / (ms * 1e-3), "unit": "videos/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32 (3xbf16 split MMA)", "data": "synthetic",
            "config": {"workload": f"CACNF inference on precomputed ResNet3D features [B, 2048, 2x4x4] + Something-Else layouts "
                                   f"(17 x 5), batch {B} per GPU, dense layouts, random-init weights",
                       "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no collective",
                       "l2_policy": "activations (GBs per step) far exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": videos / (e2e_ms * 1e-3), "unit": "videos/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(sum(v["launches"] for v in prof.values())),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "gemm_tcgen05_kernel (all projection GEMMs of a CACNF forward)",
                         "launches_per_step": gemm["launches"] / args.steps, "kernel_ms_per_step": gemm_ms,
                         "algorithmic_flops_per_step": gemm_flops, "peak_source": f"bf16 dense sustained, {peak_src}"},
            "cpu_baseline": cpu, "clocks": clocks,
            "breakdown_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            "profiled_pass_ms_per_step": prof_ms / args.steps,
        }
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict

    if args.workload in ("train", "cacnf"):
        (run_train if args.workload == "train" else run_cacnf)(args, rank, local_rank, world, torch, dist)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    spec = stlt_b200.SOMETHING_ELSE if args.layout == "something" else stlt_b200.ACTION_GENOME
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
    torch.manual_seed(0)
    model = stlt_b200.Stlt(cfg, precision=args.dtype)
    model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
    model = model.to("cuda")
    model.train(False)

    keys = ["categories", "boxes", "frame_types", "lengths"] + (["scores"] if spec["scores"] else [])
    full = make_batch(args.batch, args.layout, ragged=False, seed=100 + rank)
    batch_host = {k: full[k].pin_memory() for k in keys}
    batch_dev = {k: v.cuda() for k, v in batch_host.items()}
    logits_host = torch.empty((args.batch, spec["num_classes"]), dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in batch_host.values())
    d2h = logits_host.numel() * logits_host.element_size()
    peaks, peak_src = load_peaks()
    B, L, S = full["categories"].shape

    def measure(precision, with_clocks):
        model.precision = precision
        with torch.no_grad():
            model(batch_dev)  # packs weights, sizes the workspace
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if with_clocks else None
        if sampler:
            sampler.start()
        ms, _ = time_resident(model, batch_dev, args.steps, args.warmup, world, torch, dist)
        clocks = sampler.stop() if sampler else None
        # end-to-end pass right after the resident one (the parts drift by a few percent over tens of seconds under
        # their power cap, so the two headline numbers are taken back to back)
        e2e_ms = time_e2e(model, batch_host, batch_dev, logits_host, args.steps, max(args.warmup, 1), world, torch, dist)
        # third timed pass of the same K steps with per-launch CUDA events -> roofline, breakdown
        prof_ms, prof = time_resident(model, batch_dev, args.steps, 1, world, torch, dist, profile=True)
        launches = model.last_launch_count() * args.steps
        videos = args.batch * world * args.steps
        res = {
            "value": videos / (ms * 1e-3), "ms_per_step": ms / args.steps,
            "e2e": {"value": videos / (e2e_ms * 1e-3), "unit": "videos/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "roofline": roofline_from_profile(prof, args.steps, precision, peaks, peak_src),
            "gpu_launches": launches,
            "breakdown_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            "profiled_pass_ms_per_step": prof_ms / args.steps,
        }
        flops = FLOPS_PER_VIDEO[args.layout] * args.batch
        res["model_tflops"] = flops / (ms / args.steps * 1e-3) / 1e12
        if precision == "bf16":
            # the same K steps with the LayerNorms in their own kernels (stlt_set_fused_ln(0)): the GEMM launches
            # then contain nothing but the projections, which is the figure comparable to a plain GEMM roofline
            model.set_fused_layer_norm(False)
            u_ms, u_prof = time_resident(model, batch_dev, args.steps, 2, world, torch, dist, profile=True)
            model.set_fused_layer_norm(True)
            u_roof = roofline_from_profile(u_prof, args.steps, precision, peaks, peak_src)
            res["separate_layernorm_kernels"] = {
                "value": args.batch * world * args.steps / (u_ms * 1e-3), "unit": "videos/s", "ms_per_step": u_ms / args.steps,
                "gemm_roofline_frac": u_roof["frac"], "gemm_ms_per_step": u_roof["kernel_ms_per_step"],
                "add_ln_ms_per_step": u_prof["add_ln"]["ms"] / args.steps,
                "note": "profiled pass (per-launch CUDA events), LayerNorm fusion off"}
        return res, clocks

    main_res, clocks = measure(args.dtype, with_clocks=True)
    secondary = None
    if not args.no_secondary:
        other = "fp32" if args.dtype == "bf16" else "bf16"
        sec, _ = measure(other, with_clocks=False)
        secondary = {"dtype": other, "value": sec["value"], "unit": "videos/s", "ms_per_step": sec["ms_per_step"],
                     "e2e": sec["e2e"], "roofline": sec["roofline"], "model_tflops": sec["model_tflops"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_oracle_throughput(args.layout, 8, seconds=args.cpu_seconds)
        cpu = {"value": r["value"], "unit": "videos/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['steps']} forwards of batch 8 (BASELINE configs[0]) in {args.cpu_seconds:.0f} s, "
                         f"oracle/stlt_oracle.py on {r['cores']} torch threads (host cpus {r['host_cpus']})"}

    if rank == 0:
        line = {
            "metric": "stlt_inference_videos_per_sec", "value": main_res["value"], "unit": "videos/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32 (3xbf16 split MMA)", "data": "synthetic",
            "config": inference_config(args.layout, args.batch, world, L, S, spec["num_classes"]),
            "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"], "roofline": main_res["roofline"],
            "cpu_baseline": cpu, "clocks": clocks, "model_tflops": main_res["model_tflops"],
            "breakdown_ms_per_step": main_res["breakdown_ms_per_step"],
            "profiled_pass_ms_per_step": main_res["profiled_pass_ms_per_step"],
            "separate_layernorm_kernels": main_res.get("separate_layernorm_kernels"), "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
