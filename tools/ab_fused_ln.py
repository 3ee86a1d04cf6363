#!/usr/bin/env python
"""A/B timing of the bf16 forward with and without the LayerNorm-fused GEMM epilogues, same process / same GPU."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import stlt_b200  # noqa: E402
from stlt_b200.synthetic import make_batch, random_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4)
torch.manual_seed(0)
model = stlt_b200.Stlt(cfg, precision="bf16")
model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
model = model.cuda()
model.train(False)
batch = {k: v.cuda() for k, v in make_batch(B, "something", ragged=False, seed=1).items()}


def timed(steps=10):
    with torch.no_grad():
        for _ in range(3):
            model(batch)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            model(batch)
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for rnd in range(3):
    for fused in (True, False):
        model.set_fused_layer_norm(fused)
        ms = timed()
        print(f"round {rnd} fused_ln={fused}: {ms:.3f} ms/step = {B / ms * 1e3:.0f} videos/s")
