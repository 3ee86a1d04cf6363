#!/usr/bin/env python
"""Headline benchmark: STLT inference throughput (videos/sec) on N B200s, batch-sharded.

    python bench.py --gpus N --steps K --warmup W            # this implementation (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the unmodified reference on the host cores

One "step" is one forward of the hot path over one batch of synthetic layouts of the Something-Else shape
(BASELINE.json configs[1]: 16+1 frames x 5 slots, 174 classes, batch 4096 per GPU, dense: every frame carries
4 boxes). Rank 0 prints ONE JSON line:
  value    : videos/s with inputs resident in HBM (CUDA events, max over ranks); the steady-state forward is
             replayed as a CUDA graph (Stlt.enable_cuda_graphs)
  e2e      : videos/s through the public module call with pinned-host inputs (H2D + forward + D2H of the logits
             + a sync per step, as the reference's inference loop does)
  roofline : the tcgen05 projection GEMMs (99.8 % of the FLOPs; in bf16 mode their epilogues also carry the
             LayerNorms, residual adds and the attention): executed FLOPs per step / their summed CUDA-event
             time (events around every launch, in a third, eager timed pass of the same K steps), vs the
             measured bf16 peak
  parity   : logits of 256 videos of the TIMED batch vs the reference's CPU forward, both precisions (outside
             the timed region)
  cpu_baseline: the reference's CPU forward (oracle/_ref: the unmodified reference, byte-compiled; the oracle
             port only when that is absent) on this box's host cores, bounded sample, 1 thread and all threads
  other_configs: short runs of BASELINE configs[2..4] (Action-Genome shape, training step, CACNF) and of the
             raw-layouts -> metric pipeline on the same GPUs (--no-extras skips them; --workload X runs one of
             them as the headline instead)
For N > 1 launch with torchrun (one process per GPU); inference needs no collective — every rank runs its own
batch (weak scaling) and only the timing is reduced (MAX) over ranks.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

FLOPS_PER_VIDEO = {"something": 6_752_443_392, "action_genome": 12_548_941_824}
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
SHAPES = {"something": (17, 5, 174), "action_genome": (17, 11, 157)}  # frames, slots, classes
PARITY_VIDEOS = 256


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
    p.add_argument("--no-parity", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the short runs of the other BASELINE configs")
    p.add_argument("--no-graphs", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--ref-batch", type=int, default=64, help="videos per CPU step of the reference arm")
    p.add_argument("--workload", default="inference", choices=["inference", "train", "cacnf", "pipeline"],
                   help="'train' = BASELINE configs[3]: fwd + bwd + clip + AdamW, NCCL gradient all-reduce for N > 1; "
                        "'cacnf' = configs[4]; 'pipeline' = raw layouts -> batch builder -> forward -> top-k counters")
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


def host_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))


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


# ---------------------------------------------------------------------------------------------------------
# CPU side: the reference's own forward on the host cores (oracle/_ref), the oracle port as fallback
# ---------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference `Stlt` (src/modelling/models.py:166-195) on CPU with the weights of the B200 arm.
    kind = "reference": the unmodified reference, byte-compiled into oracle/_ref by oracle/build_ref.py;
    kind = "port": oracle/stlt_oracle.py (only when oracle/_ref is absent)."""

    def __init__(self, layout: str, state_dict=None):
        import torch
        import stlt_b200
        from oracle import ref_loader
        from stlt_b200.synthetic import random_state_dict
        self.torch = torch
        self.layout = layout
        spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
        self.spec = spec
        if state_dict is None:
            cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
            torch.manual_seed(0)
            state_dict = random_state_dict(stlt_b200.Stlt(cfg).state_dict(), seed=0)
        self.sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        ref = ref_loader.load()
        if ref is not None:
            models, configs = ref
            self.kind = "reference"
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.model = models.Stlt(configs.StltModelConfig(num_classes=spec["num_classes"],
                                                                 unique_categories=spec["unique_categories"]))
            self.model.load_state_dict(self.sd)  # strict: the drop-in's 174 keys are the reference's
            self.model.train(False)
            self.what = "unmodified reference Stlt (oracle/_ref, byte-compiled from /root/reference/src)"
        else:
            self.kind = "port"
            self.model = None
            self.what = "oracle/stlt_oracle.py (CPU restatement; oracle/_ref not built on this machine)"

    def forward(self, batch):
        torch = self.torch
        with torch.no_grad():
            if self.model is not None:
                return self.model(batch)["stlt"]
            from oracle import stlt_oracle
            return stlt_oracle.stlt_forward(self.sd, batch)

    def throughput(self, batch_size: int, threads: int, warmup: int = 1, steps: int | None = None,
                   seconds: float = 0.0, seed: int = 0):
        torch = self.torch
        from stlt_b200.synthetic import make_batch
        torch.set_num_threads(threads)
        data = make_batch(batch_size, self.layout, ragged=False, seed=seed)
        for _ in range(warmup):
            self.forward(data)
        times = []
        t_end = time.perf_counter() + seconds
        while (steps is not None and len(times) < steps) or \
                (steps is None and (time.perf_counter() < t_end or len(times) < 2)):
            t0 = time.perf_counter()
            self.forward(data)
            times.append(time.perf_counter() - t0)
        total = sum(times)
        return {"value": batch_size * len(times) / total, "unit": "videos/s", "ms_per_step": 1e3 * total / len(times),
                "median_ms_per_step": 1e3 * statistics.median(times), "steps": len(times), "batch": batch_size,
                "threads": torch.get_num_threads(), "host_cpus": os.cpu_count()}


def cpu_rows(cpu: CpuReference, budget_s: float):
    """the reference forward with 1 thread and with all host threads, batch 8 (configs[0]) and
    batch 64, bounded to about `budget_s` seconds in total."""
    n = host_threads()
    rows = [cpu.throughput(8, 1, warmup=1, steps=3)]
    rows.append(cpu.throughput(8, n, warmup=2, seconds=budget_s * 0.3))
    rows.append(cpu.throughput(64, n, warmup=1, seconds=budget_s * 0.4))
    return rows


def inference_config(layout: str, batch: int, world: int, L: int, S: int, num_classes: int) -> dict:
    return {
        "workload": f"STLT inference, {layout} shape (L={L} frames x S={S} slots, "
                    f"{num_classes} classes), batch {batch} per GPU, dense layouts, random-init weights",
        "global_batch": batch * world, "parallelism": f"batch-sharded x{world}, no collective",
        "l2_policy": "activations (GBs per step) far exceed the 126 MB L2; no explicit flush",
    }


def reference_pipeline_leg(cpu: CpuReference, n_videos: int, batch_size: int, threads: int):
    """The reference inference loop (src/inference.py:75-78) on CPU from RAW layouts: its own StltDataset +
    StltCollater (Python, per object) -> model -> EvaluatorSomething. Needs oracle/_ref."""
    from oracle import ref_loader
    mods = ref_loader.load_data()
    if mods is None or cpu.model is None or cpu.layout != "something":
        return None
    import importlib
    import torch
    from stlt_b200.synthetic import make_layout_dataset
    datasets, _ = mods
    configs = importlib.import_module("modelling.configs")
    evaluation = importlib.import_module("utils.evaluation")
    videos, sizes = make_layout_dataset("something", n_videos, seed=1, dense=True)
    for i, v in enumerate(videos):
        v["template"] = f"t{i % 174}"
    labels = {f"t{i}": i for i in range(174)}
    torch.set_num_threads(threads)
    with tempfile.TemporaryDirectory() as tmp:
        paths = {}
        for name, obj in (("dataset", videos), ("labels", labels), ("sizes", sizes)):
            paths[name] = os.path.join(tmp, name + ".json")
            with open(paths[name], "w") as f:
                json.dump(obj, f)
        t0 = time.perf_counter()
        cfg = configs.DataConfig(dataset_name="something", dataset_path=paths["dataset"], labels_path=paths["labels"],
                                 videoid2size_path=paths["sizes"], videos_path=None, train=False)
        ds = datasets.StltDataset(cfg)
        loader = torch.utils.data.DataLoader(ds, batch_size=batch_size, collate_fn=datasets.StltCollater(cfg))
        setup_s = time.perf_counter() - t0
        evaluator = evaluation.EvaluatorSomething(len(ds), 174, ("stlt",))
        t0 = time.perf_counter()
        t_data = 0.0
        with torch.no_grad():
            it = iter(loader)
            while True:
                td = time.perf_counter()
                try:
                    batch = next(it)
                except StopIteration:
                    break
                t_data += time.perf_counter() - td
                logits = cpu.model(batch)
                evaluator.process(logits, batch["labels"])
        total = time.perf_counter() - t0
        metrics = evaluator.evaluate()
    return {"value": n_videos / total, "unit": "videos/s", "videos": n_videos, "batch": batch_size, "threads": threads,
            "dataset_collate_fraction": t_data / total, "dataset_setup_s": setup_s,
            "top1": metrics["stlt_top1_accuracy"],
            "what": "reference StltDataset + StltCollater + Stlt + EvaluatorSomething on CPU (src/inference.py:75-78)"}


def run_reference(args, rank: int):
    """Reference arm: the reference's own CPU implementation of the path on the host cores, all threads.
    Each step is one forward of --ref-batch videos of the B200 arm's workload (a bounded sample of it)."""
    if rank != 0:
        return
    L, S, C = SHAPES[args.layout]
    cpu = CpuReference(args.layout)
    n = host_threads()
    r = cpu.throughput(args.ref_batch, n, warmup=max(args.warmup, 1), steps=args.steps, seed=100)
    rows = [cpu.throughput(8, 1, warmup=1, steps=2), cpu.throughput(8, n, warmup=1, steps=max(3, args.steps // 2))]
    config = inference_config(args.layout, args.batch, args.gpus, L, S, C)
    config["reference_sample"] = (f"each timed step is ONE CPU forward of {args.ref_batch} videos of this workload "
                                  f"(fp32, {r['threads']} threads, rank 0 only); throughput is a rate, the batch-"
                                  f"{args.batch} step itself would take {args.batch / r['value']:.0f} s per GPU-batch")
    config["executed_batch_per_step"] = args.ref_batch
    line = {
        "impl": "reference", "metric": "stlt_inference_videos_per_sec", "value": r["value"], "unit": "videos/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": r["value"], "unit": "videos/s", "cores": r["threads"], "kind": cpu.kind,
                         "sample": f"{r['steps']} fp32 forwards of {args.ref_batch} videos (dense layouts) through the "
                                   f"{cpu.what}, host cpus {r['host_cpus']}",
                         "rows": [{k: x[k] for k in ("batch", "threads", "value", "median_ms_per_step", "steps")}
                                  for x in rows + [r]]},
        "e2e": {"value": r["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_extras:
        try:
            line["other_configs"] = {"pipeline": reference_pipeline_leg(cpu, 2 * args.ref_batch, args.ref_batch, n)}
        except Exception as e:  # the headline line must survive a failure of the optional leg
            line["other_configs"] = {"pipeline": {"error": repr(e)}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------
def timed(fn, steps, warmup, world, torch, dist, before=None, after=None):
    """K calls of fn bracketed by barrier + synchronize, CUDA events on the current stream, MAX over ranks.
    Returns (max ms over ranks, this rank's ms)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if before:
        before()
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
    local = start.elapsed_time(end)
    extra = after() if after else None
    ms = local
    if world > 1:
        t = torch.tensor([local], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, local, extra


def gather_ranks(values, world, torch, dist):
    """values: list of floats of this rank -> list over ranks."""
    if world == 1:
        return [values]
    t = torch.tensor(values, dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def time_e2e(model, batch_host, steps, warmup, world, torch, dist, output_key="stlt"):
    """Public-API timing with HOST inputs. Every step uploads its inputs from pinned host memory and
    delivers its logits to host memory; stlt_b200.pipeline.HostPipeline (part of the package's public
    API) overlaps the copies of neighbouring steps with compute, the way a DataLoader with pinned
    memory feeds the reference loop. The timed region ends when the last logits are on the host."""
    from stlt_b200.pipeline import HostPipeline
    pipe = HostPipeline(model, output_key)
    checksum = [0.0]

    def run(n):
        for host_logits in pipe.run(batch_host for _ in range(n)):
            checksum[0] += float(host_logits[0, 0])  # touch the delivered result on the host

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


def measured_traffic(precision: str):
    """DRAM bytes per GEMM launch from the committed ncu capture (tools/measure_traffic.py writes
    profiles/roofline_traffic.json together with the digest of the kernel sources it was measured on).
    dram__bytes cannot be read outside a profiler, so a capture of OTHER sources is reported as null."""
    path = ROOT / "profiles" / "roofline_traffic.json"
    if not path.exists():
        return None, "no ncu capture committed (tools/measure_traffic.py)"
    try:
        d = json.loads(path.read_text())
        import __graft_entry__
        digest = __graft_entry__._load_build_module()._source_digest()
        if d.get("source_digest") != digest:
            return None, "profiles/roofline_traffic.json was captured on other kernel sources (stale digest); re-run tools/measure_traffic.py"
        v = d.get(precision)
        return v, f"ncu dram__bytes_read+write per GEMM launch, {d.get('command', 'tools/measure_traffic.py')}"
    except Exception as e:
        return None, f"unreadable traffic file: {e!r}"


KERNEL_OF_ROLE = {
    "in_proj": "gemm_tcgen05_kernel (pending LayerNorm + in-projection: NORM_A epilogue; plain for the first layer of a stack)",
    "qkv_attention": "qkv_attention_kernel (in-projection + attention; FLOPs of the in-projection)",
    "out_proj": "gemm_tcgen05_kernel (out-projection + residual add + LayerNorm statistics: RESID epilogue)",
    "linear1": "gemm_tcgen05_kernel (pending LayerNorm + linear1 + GELU: NORM_A epilogue)",
    "linear2": "gemm_tcgen05_kernel (linear2 + residual add + LayerNorm statistics: RESID epilogue)",
    "gradient": "gemm_tcgen05_kernel (data- and weight-gradient GEMMs)",
    "other_gemm": "gemm_tcgen05_kernel (other shapes)",
}


def roofline_by_kernel(roles, steps, precision, peaks):
    """Per-role split of the GEMM-class time (stlt_get_profile_by_role): ms and algorithmic TFLOP/s per step against
    the sustained bf16 peak; for the out-projection, whose fused epilogue makes it HBM-bound in bf16 mode, also the
    algorithmic bytes (context in, residual stream in / out, bf16 operand copy out) against the HBM peak."""
    peak = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
    hbm = float(peaks.get("hbm_gbs") or 0.0)
    rows = []
    for role, v in roles.items():
        ms = v["ms"] / max(steps, 1)
        flops = v["flops"] / max(steps, 1)
        tf = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        row = {"role": role, "kernel": KERNEL_OF_ROLE.get(role, role), "launches_per_step": v["launches"] / max(steps, 1),
               "ms_per_step": ms, "achieved_tflops": tf, "frac_of_tensor_peak": tf / peak}
        if precision == "fp32":  # 3 bf16 MMAs per algorithmic MMA
            row["mma_frac_of_tensor_peak"] = 3.0 * tf / peak
        if role == "out_proj" and hbm > 0 and ms > 0:
            m_rows = flops / (2.0 * 768 * 768)
            planes = 2 if precision == "fp32" else 1  # bf16 planes of the context read and of the operand copy written
            gbps = m_rows * 768 * (2 * planes + 4 + 4 + 2 * planes) / (ms * 1e-3) / 1e9
            row.update({"algorithmic_gbps": gbps, "frac_of_hbm_peak": gbps / hbm})
        rows.append(row)
    return rows


def roofline_from_profile(prof, steps, precision, peaks, peak_src, roles=None):
    gemm = prof["gemm"]
    algorithmic = gemm["flops"] / max(steps, 1)                   # 2*M*N*K of the launches actually issued
    executed = algorithmic * (3.0 if precision == "fp32" else 1.0)  # fp32 mode: 3 bf16 MMAs per algorithmic MMA
    ms = gemm["ms"] / max(steps, 1)
    achieved = algorithmic / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
    traffic, traffic_note = measured_traffic(precision)
    out = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_note,
        "kernel": "gemm_tcgen05_kernel + qkv_attention_kernel (all projection GEMMs of a step" +
                  ("; in bf16 mode their epilogues also carry the residual adds, the LayerNorms and the attention)"
                   if precision == "bf16" else ")"),
        "launches_per_step": gemm["launches"] / max(steps, 1), "kernel_ms_per_step": ms,
        "algorithmic_flops_per_step": algorithmic,
        "peak_source": f"bf16 dense sustained, {peak_src}",
    }
    if precision == "bf16":
        out["note"] = ("kernel_ms_per_step includes the attention (it runs inside the in-projection kernel and adds time but no "
                       "counted FLOPs). Round 1 quoted 0.76 for its GEMM launches alone; with its 3.3 ms of separate attention "
                       "kernels counted the same way that figure was 23.95 TFLOP / (22.58 + 3.28 ms) = 0.664 of the same peak")
    if roles:
        out["by_kernel"] = roofline_by_kernel(roles, steps, precision, peaks)
    if precision == "fp32":
        out["mma_tflops_issued"] = executed / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        out["note"] = ("fp32-parity mode issues 3 bf16 MMAs (hi*hi + lo*hi + hi*lo) per algorithmic MMA; "
                       "mma_tflops_issued / peak is the tensor-pipe fraction, frac is in algorithmic FLOPs")
    return out


def parity_stamp(model, batch_host, batch_dev, cpu: CpuReference, torch):
    """Logits of the first PARITY_VIDEOS videos of the TIMED batch, taken from a forward of the whole batch in each
    precision, against the reference's CPU forward on the same inputs and weights (outside the timed region)."""
    n = min(PARITY_VIDEOS, batch_host["categories"].shape[0])
    small = {k: v[:n].clone() for k, v in batch_host.items()}
    small["src_key_padding_mask_boxes"] = small["categories"] == 0
    small["src_key_padding_mask_frames"] = small["frame_types"] == 0
    torch.set_num_threads(host_threads())
    want = cpu.forward(small).double()
    scale = float(want.abs().max())
    out = {"videos": n, "reference": cpu.kind, "tolerance": {"fp32": 1e-4, "bf16": 2e-2}}
    keep = model.precision
    for precision in ("fp32", "bf16"):
        model.precision = precision
        with torch.no_grad():
            got = model(batch_dev)["stlt"][:n].double().cpu()
        diff = (got - want).abs()
        err = float(diff.max()) / scale
        agree = got.argmax(-1) == want.argmax(-1)
        # a flipped arg-max is a parity failure only when the reference's own top-1 / top-2 margin exceeds twice the error
        top2 = want.topk(2, dim=-1).values
        margin = (top2[:, 0] - top2[:, 1])
        unexplained = int(((~agree) & (margin > 2 * diff.max(-1).values)).sum())
        out[precision] = err
        out[f"{precision}_top1_agree"] = int(agree.sum())
        out[f"{precision}_top1_flips_beyond_error"] = unexplained
    model.precision = keep
    out["ok"] = bool(out["fp32"] < 1e-4 and out["bf16"] < 2e-2 and out["fp32_top1_flips_beyond_error"] == 0
                     and out["bf16_top1_flips_beyond_error"] == 0)
    return out


def run_inference(args, rank, local_rank, world, torch, dist, layout=None, batch=None, steps=None, warmup=None,
                  full=True):
    """BASELINE configs[1] (and [2] with layout = action_genome). full = False: the short form used for other_configs."""
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    layout = layout or args.layout
    batch = batch or args.batch
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
    torch.manual_seed(0)
    model = stlt_b200.Stlt(cfg, precision=args.dtype)
    sd = random_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    model = model.to("cuda")
    model.train(False)

    keys = ["categories", "boxes", "frame_types", "lengths"] + (["scores"] if spec["scores"] else [])
    data = make_batch(batch, layout, ragged=False, seed=100 + rank)
    batch_host = {k: data[k].pin_memory() for k in keys}
    batch_dev = {k: v.cuda() for k, v in batch_host.items()}
    h2d = sum(v.numel() * v.element_size() for v in batch_host.values())
    d2h = batch * spec["num_classes"] * 4
    peaks, peak_src = load_peaks()
    B, L, S = data["categories"].shape
    use_graphs = not args.no_graphs

    def fwd():
        with torch.no_grad():
            model(batch_dev)

    def measure(precision, with_clocks, want_unfused):
        model.precision = precision
        model.enable_cuda_graphs(use_graphs)
        fwd()  # packs weights, sizes the workspace
        fwd()  # captures the graph
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if with_clocks else None
        if sampler:
            sampler.start()
        ms, local_ms, _ = timed(fwd, steps, warmup, world, torch, dist)
        clocks = sampler.stop() if sampler else None
        # end-to-end pass right after the resident one (the parts drift by a few percent over tens of seconds under
        # their power cap, so the two headline numbers are taken back to back)
        e2e_ms = time_e2e(model, batch_host, steps, max(warmup, 1), world, torch, dist)
        # third timed pass of the same K steps, eager, with per-launch CUDA events -> roofline, breakdown
        model.enable_cuda_graphs(False)
        prof_ms, prof_local, prof = timed(fwd, steps, 1, world, torch, dist, before=lambda: model.set_profiling(True),
                                          after=lambda: (model.get_profile(), model.set_profiling(False))[0])
        roles = model.get_profile_by_role()
        launches = model.last_launch_count() * steps
        kernel_ms = sum(v["ms"] for v in prof.values()) / steps
        per_rank = gather_ranks([local_ms / steps, prof_local / steps, kernel_ms], world, torch, dist)
        videos = batch * world * steps
        res = {
            "value": videos / (ms * 1e-3), "ms_per_step": ms / steps,
            "e2e": {"value": videos / (e2e_ms * 1e-3), "unit": "videos/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / steps},
            "roofline": roofline_from_profile(prof, steps, precision, peaks, peak_src, roles),
            "gpu_launches": launches,
            "breakdown_ms_per_step": {k: v["ms"] / steps for k, v in prof.items()},
            "profiled_pass_ms_per_step": prof_ms / steps,
            # straggler visibility: per rank, the graph-replayed step, the eager profiled step and the sum of its kernels
            "per_rank": [{"rank": i, "ms_per_step": r[0], "eager_profiled_ms_per_step": r[1], "kernel_ms_per_step": r[2]}
                         for i, r in enumerate(per_rank)],
            "cuda_graph": use_graphs,
        }
        res["model_tflops"] = FLOPS_PER_VIDEO[layout] * batch / (ms / steps * 1e-3) / 1e12
        res["whole_step_frac_of_peak"] = (prof["gemm"]["flops"] / steps) / (ms / steps * 1e-3) / 1e12 / \
            float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
        if precision == "bf16" and want_unfused:
            # A/B of the epilogue fusions and of the pad-skipping layout on the same GPU: graph-replayed passes of the
            # same K steps, the variants interleaved twice (the parts drift by several percent over a run under their
            # power cap, so back-to-back blocks of one variant are not comparable); best of the two rounds
            variants = (("default", (True, True, True, False)), ("two_bf16_plane_residual_stream", (True, True, True, True)),
                        ("padded_grid", (True, True, False, False)), ("attention_unfused", (True, False, False, False)),
                        ("layernorm_and_attention_unfused", (False, False, False, False)))
            best = {}
            model.enable_cuda_graphs(use_graphs)
            for _ in range(2):
                for name, (ln, at, cp, hl) in variants:
                    model.set_fused_layer_norm(ln)
                    model.set_fused_attention(at)
                    model.set_compaction(cp)
                    model.set_hilo_residual(hl)
                    u_ms, _, _ = timed(fwd, steps, 2, world, torch, dist)
                    best[name] = min(best.get(name, u_ms), u_ms)
            model.set_fused_layer_norm(True)
            model.set_fused_attention(True)
            model.set_compaction(True)
            model.set_hilo_residual(False)
            ab = {name: {"value": videos / (u_ms * 1e-3), "unit": "videos/s", "ms_per_step": u_ms / steps}
                  for name, u_ms in best.items()}
            ab["note"] = ("graph-replayed, variants interleaved, best of 2 rounds. two_bf16_plane_residual_stream: an optional "
                          "switch that is off by default. From padded_grid on, each variant switches one more default-on thing "
                          "off: the pad-skipping row layout, the attention in the in-projection epilogue, the LayerNorms in the "
                          "GEMM epilogues (= the round-1 kernels without LayerNorm fusion)")
            res["fusion_ab"] = ab
            # the same model on a RAGGED batch of the same size (lengths ~ U{2..17}, 0..4 boxes per frame: the parity
            # batch of SURVEY.md 8(d)): what the pad-skipping layout buys on data that is not dense
            rag = make_batch(batch, layout, ragged=True, seed=300 + rank)
            rag_dev = {k: rag[k].cuda() for k in keys}

            def fwd_rag():
                with torch.no_grad():
                    model(rag_dev)

            model.enable_cuda_graphs(use_graphs)
            r_ms, _, _ = timed(fwd_rag, steps, 3, world, torch, dist)
            model.set_compaction(False)
            p_ms, _, _ = timed(fwd_rag, steps, 3, world, torch, dist)
            model.set_compaction(True)
            model.enable_cuda_graphs(False)
            res["ragged_batch"] = {"value": videos / (r_ms * 1e-3), "unit": "videos/s", "ms_per_step": r_ms / steps,
                                   "padded_grid_value": videos / (p_ms * 1e-3), "padded_grid_ms_per_step": p_ms / steps,
                                   "note": "same batch size, lengths ~ U{2..L}, 0..max objects per frame; device-resident"}
            del rag_dev
        return res, clocks

    main_res, clocks = measure(args.dtype, with_clocks=True, want_unfused=full)
    secondary = None
    if full and not args.no_secondary:
        other = "fp32" if args.dtype == "bf16" else "bf16"
        sec, _ = measure(other, with_clocks=False, want_unfused=False)
        secondary = {"dtype": other, "value": sec["value"], "unit": "videos/s", "ms_per_step": sec["ms_per_step"],
                     "e2e": sec["e2e"], "roofline": sec["roofline"], "model_tflops": sec["model_tflops"],
                     "breakdown_ms_per_step": sec["breakdown_ms_per_step"]}
        if other == "fp32" and not args.no_extras:
            # the LayerNorm-fused epilogues of the fp32-parity mode against its separate add + LayerNorm kernels:
            # graph-replayed, interleaved on this GPU, best of two rounds (as fusion_ab above)
            model.precision = "fp32"
            model.enable_cuda_graphs(use_graphs)
            best = {}
            for _ in range(2):
                for name, flag in (("default", True), ("separate_layernorm_kernels", False)):
                    model.set_fused_layer_norm(True, fp32=flag)
                    u_ms, _, _ = timed(fwd, steps, 2, world, torch, dist)
                    best[name] = min(best.get(name, u_ms), u_ms)
            model.set_fused_layer_norm(True)
            model.enable_cuda_graphs(False)
            secondary["fusion_ab"] = {name: {"value": batch * world * steps / (u_ms * 1e-3), "unit": "videos/s",
                                             "ms_per_step": u_ms / steps} for name, u_ms in best.items()}
        model.precision = args.dtype

    cpu = cpu_line = parity = None
    if rank == 0 and full and (not args.no_parity or (world == 1 and not args.no_cpu_baseline)):
        cpu = CpuReference(layout, sd)
    if cpu is not None and not args.no_parity:
        model.enable_cuda_graphs(False)
        parity = parity_stamp(model, batch_host, batch_dev, cpu, torch)
    if cpu is not None and world == 1 and not args.no_cpu_baseline:
        rows = cpu_rows(cpu, args.cpu_seconds)
        best = max(rows[1:], key=lambda r: r["value"])
        cpu_line = {"value": best["value"], "unit": "videos/s", "cores": best["threads"], "kind": cpu.kind,
                    "sample": f"{best['steps']} fp32 forwards of batch {best['batch']} of this workload through the "
                              f"{cpu.what} on {best['threads']} torch threads (host cpus {best['host_cpus']})",
                    "rows": [{k: x[k] for k in ("batch", "threads", "value", "median_ms_per_step", "steps")} for x in rows]}

    line = {
        "metric": "stlt_inference_videos_per_sec", "value": main_res["value"], "unit": "videos/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": main_res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.dtype == "bf16" else "f32 (3xbf16 split MMA)", "data": "synthetic",
        "config": inference_config(layout, batch, world, L, S, spec["num_classes"]),
        "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"], "roofline": main_res["roofline"],
        "cpu_baseline": cpu_line, "parity": parity, "clocks": clocks, "model_tflops": main_res["model_tflops"],
        "whole_step_frac_of_peak": main_res["whole_step_frac_of_peak"],
        "breakdown_ms_per_step": main_res["breakdown_ms_per_step"],
        "profiled_pass_ms_per_step": main_res["profiled_pass_ms_per_step"],
        "per_rank": main_res["per_rank"], "cuda_graph": main_res["cuda_graph"],
        "fusion_ab": main_res.get("fusion_ab"), "ragged_batch": main_res.get("ragged_batch"), "secondary": secondary,
    }
    del model, batch_dev
    gc.collect()
    torch.cuda.empty_cache()
    return line


def run_pipeline(args, rank, local_rank, world, torch, dist, steps=None):
    """Raw layouts -> metric, the whole reference inference loop (src/inference.py:75-78) on the device: per step
    LayoutStore.build_batch (StltDataset.__getitem__ + StltCollater: frame sampling, score filter, fix_box,
    normalisation, padding, masks — stage 1 of the north star) -> Stlt forward -> TopKCounter; one host
    synchronisation at the end (evaluate)."""
    import stlt_b200
    from stlt_b200 import LayoutStore, TopKCounter
    from stlt_b200.synthetic import make_layout_dataset, random_state_dict
    steps = steps or args.steps
    B = args.batch
    spec = stlt_b200.SOMETHING_ELSE
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
    torch.manual_seed(0)
    model = stlt_b200.Stlt(cfg, precision=args.dtype, cuda_graphs=not args.no_graphs)
    model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
    model = model.to("cuda")
    model.train(False)
    videos, sizes = make_layout_dataset("something", B, seed=1 + rank, dense=True)
    t0 = time.perf_counter()
    store = LayoutStore("something", videos, sizes)
    store_s = time.perf_counter() - t0
    labels = (torch.arange(B) % spec["num_classes"]).cuda()
    order = list(range(B))
    counter = TopKCounter(B * steps)

    def step():
        with torch.no_grad():
            batch = store.build_batch(order)
            counter.process(model(batch)["stlt"], labels)

    ms, _, _ = timed(step, steps, max(args.warmup, 2), world, torch, dist, before=counter.reset)
    metrics = counter.evaluate()
    objects = int(store.obj_categories.numel())
    line = {
        "metric": "stlt_pipeline_videos_per_sec", "value": B * world * steps / (ms * 1e-3), "unit": "videos/s",
        "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 2), "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"raw layouts (CSR store of {B} videos, {objects} detections per GPU) -> build_batch -> "
                               f"STLT forward -> top-1/top-5 counters, batch {B} per GPU, Something-Else shape",
                   "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no collective"},
        "top1": metrics["stlt_top1_accuracy"], "layout_store_build_s": store_s,
    }
    del model, store
    gc.collect()
    torch.cuda.empty_cache()
    return line


def cpu_oracle_train_throughput(layout: str, batch: int, steps: int):
    """videos/s of the CPU oracle training step (autograd through the restated forward + AdamW)."""
    import torch
    import stlt_b200
    from oracle import stlt_oracle
    from stlt_b200.synthetic import make_batch, random_state_dict
    torch.set_num_threads(host_threads())
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


def run_train(args, rank, local_rank, world, torch, dist, steps=None, full=True):
    """BASELINE configs[3]: one optimisation step = forward (activations kept) + criterion + backward +
    gradient all-reduce (NCCL, N > 1) + global-norm clip + AdamW + bf16 re-pack, per-GPU batch fixed."""
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    from stlt_b200.training import FusedTrainStep, linear_schedule_with_warmup
    steps = steps or args.steps
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
    data = make_batch(B, args.layout, ragged=False, seed=100 + rank)
    g = torch.Generator().manual_seed(7 + rank)
    if args.layout == "something":
        data["labels"] = torch.randint(0, spec["num_classes"], (B,), generator=g)
        loss = "cross_entropy"
    else:
        data["labels"] = (torch.rand((B, spec["num_classes"]), generator=g) < 0.05).float()
        loss = "bce_with_logits"
    keys.append("labels")
    batch_host = {k: data[k].pin_memory() for k in keys}
    batch_dev = {k: v.cuda() for k, v in batch_host.items()}
    stepper = FusedTrainStep(model, lr=5e-5, weight_decay=1e-3, clip_val=5.0, loss=loss,
                             lr_lambda=linear_schedule_with_warmup(100, 100000))
    loss_host = torch.zeros(1).pin_memory()
    peaks, peak_src = load_peaks()
    _, L, S = data["categories"].shape

    def step_resident():
        stepper.step(batch_dev)

    def step_e2e():
        for k, v in batch_host.items():
            batch_dev[k].copy_(v, non_blocking=True)
        loss_host.copy_(stepper.step(batch_dev).reshape(1), non_blocking=True)
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, local_ms, _ = timed(step_resident, steps, args.warmup, world, torch, dist)
    clocks = sampler.stop()
    prof_ms, prof_local, prof = timed(step_resident, steps, 1, world, torch, dist,
                                      before=lambda: model.set_profiling(True),
                                      after=lambda: (model.get_profile(), model.set_profiling(False))[0])
    roles = model.get_profile_by_role()
    e2e_ms, _, _ = timed(step_e2e, steps, max(args.warmup, 1), world, torch, dist)
    videos = B * world * steps
    gemm = prof["gemm"]
    gemm_ms = gemm["ms"] / steps
    gemm_flops = gemm["flops"] / steps
    peak = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    kernel_ms = sum(v["ms"] for v in prof.values()) / steps
    per_rank = gather_ranks([local_ms / steps, prof_local / steps, kernel_ms], world, torch, dist)
    cpu = None
    if rank == 0 and world == 1 and full and not args.no_cpu_baseline:
        r = cpu_oracle_train_throughput(args.layout, 8, steps=3)
        cpu = {"value": r["value"], "unit": "videos/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['steps']} training steps (autograd + AdamW) of batch 8, oracle/stlt_oracle.py on "
                         f"{r['cores']} torch threads (host cpus {r['host_cpus']})"}
    h2d = sum(v.numel() * v.element_size() for v in batch_host.values())
    line = {
        "metric": "stlt_train_videos_per_sec", "value": videos / (ms * 1e-3), "unit": "videos/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (fp32 master weights / grads / AdamW)",
        "data": "synthetic",
        "config": {
            "workload": f"STLT training step (fwd + bwd + clip + AdamW), {args.layout} shape (L={L} x S={S}, "
                        f"{spec['num_classes']} classes), batch {B} per GPU, dense layouts, dropout {stepper.dropout_p}",
            "global_batch": B * world,
            "parallelism": f"data-parallel x{world}" + (f", NCCL gradient all-reduce in {stepper.num_buckets} buckets overlapped with the backward pass" if world > 1 else ""),
            "l2_policy": "activations (GBs per step) far exceed the 126 MB L2; no explicit flush",
        },
        "e2e": {"value": videos / (e2e_ms * 1e-3), "unit": "videos/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / steps},
        "gpu_launches": int(sum(v["launches"] for v in prof.values())),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "gemm_tcgen05_kernel (forward, data-gradient and weight-gradient GEMMs of a step)",
                     "launches_per_step": gemm["launches"] / steps, "kernel_ms_per_step": gemm_ms,
                     "algorithmic_flops_per_step": gemm_flops, "peak_source": f"bf16 dense sustained, {peak_src}",
                     # forward GEMMs by role (linear1 = the dual-output epilogue: act and act' * mask), all backward GEMMs
                     "by_kernel": [{k: v for k, v in row.items() if k not in ("algorithmic_gbps", "frac_of_hbm_peak")}
                                   for row in roofline_by_kernel(roles, steps, "bf16", peaks)]},
        "cpu_baseline": cpu, "clocks": clocks,
        "model_tflops": 3 * FLOPS_PER_VIDEO[args.layout] * B / (ms / steps * 1e-3) / 1e12,
        "breakdown_ms_per_step": {k: v["ms"] / steps for k, v in prof.items()},
        "profiled_pass_ms_per_step": prof_ms / steps,
        "per_rank": [{"rank": i, "ms_per_step": r[0], "profiled_ms_per_step": r[1], "kernel_ms_per_step": r[2]}
                     for i, r in enumerate(per_rank)],
    }
    del stepper, model, batch_dev
    gc.collect()
    torch.cuda.empty_cache()
    return line


def run_cacnf(args, rank, local_rank, world, torch, dist, batch=None, steps=None, full=True):
    """BASELINE configs[4]: CACNF inference on precomputed per-clip ResNet3D features, batch-sharded."""
    import stlt_b200
    from stlt_b200.synthetic import make_appearance_features, make_batch, random_state_dict
    steps = steps or args.steps
    cfg = stlt_b200.CacnfModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    model = stlt_b200.Cacnf(cfg, precision=args.dtype)
    model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
    model = model.to("cuda")
    model.train(False)
    B = batch or args.batch
    data = make_batch(B, "something", ragged=False, seed=100 + rank)
    keys = ["categories", "boxes", "frame_types", "lengths"]
    batch_host = {k: data[k].pin_memory() for k in keys}
    batch_host["video_features"] = make_appearance_features(B, seed=200 + rank).flatten(2).contiguous().pin_memory()
    batch_dev = {k: v.cuda() for k, v in batch_host.items()}
    peaks, peak_src = load_peaks()

    def fwd():
        with torch.no_grad():
            model(batch_dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, _, _ = timed(fwd, steps, args.warmup, world, torch, dist)
    clocks = sampler.stop()
    prof_ms, _, prof = timed(fwd, steps, 1, world, torch, dist, before=lambda: model.set_profiling(True),
                             after=lambda: (model.get_profile(), model.set_profiling(False))[0])
    e2e_ms = time_e2e(model, batch_host, steps, max(args.warmup, 1), world, torch, dist, output_key="ensemble")
    cpu = None
    if rank == 0 and world == 1 and full and not args.no_cpu_baseline:
        from oracle import stlt_oracle
        torch.set_num_threads(host_threads())
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
    videos = B * world * steps
    gemm = prof["gemm"]
    gemm_ms, gemm_flops = gemm["ms"] / steps, gemm["flops"] / steps
    peak = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    h2d = sum(v.numel() * v.element_size() for v in batch_host.values())
    line = {
        "metric": "cacnf_inference_videos_per_sec", "value": videos / (ms * 1e-3), "unit": "videos/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.dtype == "bf16" else "f32 (3xbf16 split MMA)", "data": "synthetic",
        "config": {"workload": f"CACNF inference on precomputed ResNet3D features [B, 2048, 2x4x4] + Something-Else layouts "
                               f"(17 x 5), batch {B} per GPU, dense layouts, random-init weights",
                   "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2_policy": "activations (GBs per step) far exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": videos / (e2e_ms * 1e-3), "unit": "videos/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": B * 174 * 4, "ms_per_step": e2e_ms / steps},
        "gpu_launches": int(sum(v["launches"] for v in prof.values())),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "gemm_tcgen05_kernel (all projection GEMMs of a CACNF forward)",
                     "launches_per_step": gemm["launches"] / steps, "kernel_ms_per_step": gemm_ms,
                     "algorithmic_flops_per_step": gemm_flops, "peak_source": f"bf16 dense sustained, {peak_src}"},
        "cpu_baseline": cpu, "clocks": clocks,
        "breakdown_ms_per_step": {k: v["ms"] / steps for k, v in prof.items()},
        "profiled_pass_ms_per_step": prof_ms / steps,
    }
    del model, batch_dev
    gc.collect()
    torch.cuda.empty_cache()
    return line


def summary(line: dict) -> dict:
    keep = ("metric", "value", "unit", "ms_per_step", "n_gpus", "steps", "dtype")
    out = {k: line[k] for k in keep if k in line}
    out["workload"] = line["config"]["workload"]
    if "e2e" in line:
        out["e2e"] = line["e2e"]["value"]
    if "roofline" in line:
        out["gemm_roofline_frac"] = line["roofline"]["frac"]
    for k in ("top1", "per_rank"):
        if k in line:
            out[k] = line[k]
    return out


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

    ctx = (args, rank, local_rank, world, torch, dist)
    if args.workload == "train":
        line = run_train(*ctx)
    elif args.workload == "cacnf":
        line = run_cacnf(*ctx)
    elif args.workload == "pipeline":
        line = run_pipeline(*ctx)
    else:
        line = run_inference(*ctx)
        if not args.no_extras:
            # BASELINE configs[2..4] + the raw-layouts pipeline, short runs on the same GPUs so that the driver's own
            # bench / scale files carry them (full lines: --workload train | cacnf | pipeline, --layout action_genome)
            short = max(3, min(args.steps, 5))
            extras = {}
            jobs = (
                ("pipeline_raw_layouts_to_metric", lambda: run_pipeline(*ctx, steps=short)),
                ("action_genome_inference", lambda: run_inference(*ctx, layout="action_genome", steps=short, warmup=2, full=False)),
                ("train_step", lambda: run_train(*ctx, steps=short, full=False)),
                ("cacnf_inference", lambda: run_cacnf(*ctx, batch=min(args.batch, 2048), steps=short, full=False)),
            )
            for name, job in jobs:
                try:
                    extras[name] = summary(job())
                except Exception as e:  # an optional leg must not take the headline line down (all ranks fail alike)
                    extras[name] = {"error": repr(e)[:300]}
                    gc.collect()
                    torch.cuda.empty_cache()
            line["other_configs"] = extras
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
