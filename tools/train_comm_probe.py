"""Where does the data-parallel training step lose time? (torchrun, one rank per GPU)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_comm_probe.py

Prints, as the max over ranks of CUDA-event times:
  * the gradient all-reduce ALONE, per bucket of FusedTrainStep's flat layout and as one call;
  * the training step with the all-reduce switched off (N processes side by side: host / power contention only);
  * the full step for every bucket scheme the stepper offers.
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    from stlt_b200.training import FusedTrainStep

    def timed(fn, steps=10, warmup=3):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    spec = stlt_b200.SOMETHING_ELSE
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
    model = stlt_b200.Stlt(cfg, precision="bf16")
    model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
    model = model.to("cuda").train(True)
    B = 2048
    data = make_batch(B, "something", ragged=False, seed=100 + rank)
    data["labels"] = torch.randint(0, spec["num_classes"], (B,), generator=torch.Generator().manual_seed(7 + rank))
    batch = {k: data[k].cuda() for k in ("categories", "boxes", "frame_types", "lengths", "labels")}
    stepper = FusedTrainStep(model, lr=5e-5, weight_decay=1e-3, clip_val=5.0)
    out = {"world": world, "flat_gradient_mbytes": stepper.total * 4 / 1e6}

    if world > 1:
        g = stepper.flat_grads
        out["all_reduce_alone_ms"] = {"whole": timed(lambda: dist.all_reduce(g))}
        for name, (a, b) in stepper.bucket_bounds().items():
            out["all_reduce_alone_ms"][f"{name} ({(b - a) * 4 / 1e6:.1f} MB)"] = timed(lambda: dist.all_reduce(g[a:b]))
        g.zero_()

    real_world = stepper._world
    stepper._world = lambda: 1  # no all-reduce, no 1/world scaling: N independent replicas side by side
    out["step_without_all_reduce_ms"] = timed(lambda: stepper.step(batch))
    stepper._world = real_world
    if world > 1:
        # interleaved rounds, the order reversed every round: drift of the power-capped clocks shows up as a trend over the
        # rounds instead of as a difference between schemes
        out["step_ms"] = {scheme: [] for scheme in stepper.BUCKET_SCHEMES}
        out["nccl_env"] = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
        for rnd in range(4):
            order = stepper.BUCKET_SCHEMES if rnd % 2 == 0 else tuple(reversed(stepper.BUCKET_SCHEMES))
            for scheme in order:
                stepper.bucket_scheme = scheme
                out["step_ms"][scheme].append(round(timed(lambda: stepper.step(batch)), 3))
        stepper._world = lambda: 1
        out["step_without_all_reduce_ms_after"] = timed(lambda: stepper.step(batch))
        stepper._world = real_world
    else:
        out["step_ms"] = {"single": timed(lambda: stepper.step(batch))}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
