#!/usr/bin/env python
"""Launches every HBM-bound stage of the path once at BASELINE configs[1] size so that ncu can capture them:

    ncu --set full --clock-control none --import-source on \
        -k regex:"prepare_kernel|masks_kernel|embed_kernel|attention_mma|gather_rows|build_batch" \
        -o gpurun_out/prof_stages python tools/profile_stages.py

K0 (`prepare_kernel`: fix_box + normalisation + masks, via stlt_prepare), the forward with the masks requested
(`masks_kernel`, `embed_kernel`, `attention_mma_kernel`, `frame_embed_kernel`, `gather_rows_kernel`) and the
device batch builder (`build_batch_kernel`). Without ncu it prints CUDA-event times of the same calls.
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from stlt_b200 import Stlt, StltModelConfig, prepare_layout_batch
from stlt_b200.synthetic import make_batch, make_raw_boxes


def main() -> None:
    p = argparse.ArgumentParser()
    p.add_argument("--batch", type=int, default=4096)
    p.add_argument("--layout", default="something", choices=["something", "action_genome"])
    p.add_argument("--passes", type=int, default=2, help="the first pass warms up; profile from the second")
    args = p.parse_args()
    dev = torch.device("cuda:0")
    batch = make_batch(args.batch, args.layout, ragged=False, seed=0)
    raw, sizes = make_raw_boxes(batch["categories"], seed=1)
    cfg = (StltModelConfig(num_classes=174, unique_categories=4) if args.layout == "something"
           else StltModelConfig(num_classes=157, unique_categories=38))
    torch.manual_seed(0)
    model = Stlt(cfg, precision="bf16").to(dev)
    model.train(False)
    gpu = {k: v.to(dev) for k, v in batch.items() if isinstance(v, torch.Tensor)}
    raw, sizes = raw.to(dev), sizes.to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for i in range(args.passes):
        ev[0].record()
        prep = prepare_layout_batch(raw, sizes, gpu["categories"], gpu["frame_types"])
        ev[1].record()
        with torch.no_grad():
            out = model.forward_with_taps({**gpu, "boxes": prep["boxes"]})
        ev[2].record()
        torch.cuda.synchronize()
        print(f"pass {i}: stlt_prepare {ev[0].elapsed_time(ev[1]):.3f} ms, forward with taps "
              f"{ev[1].elapsed_time(ev[2]):.3f} ms, logits {tuple(out['stlt'].shape)}")


if __name__ == "__main__":
    main()
