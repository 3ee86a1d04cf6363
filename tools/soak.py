#!/usr/bin/env python
"""Soak test: the deterministic inference paths must return bit-identical logits on every repetition (catches
intermittent races in the TMA / mbarrier / TMEM pipelines), the training step must stay finite."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import stlt_b200  # noqa: E402
from stlt_b200.synthetic import make_appearance_features, make_batch, random_state_dict  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4)
torch.manual_seed(0)
model = stlt_b200.Stlt(cfg)
model.load_state_dict(random_state_dict(model.state_dict(), seed=0))
model = model.cuda()
model.train(False)
for B in (4096, 37):
    batch = {k: v.cuda() for k, v in make_batch(B, "something", ragged=True, seed=B).items()}
    for precision in ("bf16", "fp32"):
        model.precision = precision
        with torch.no_grad():
            ref = model(batch)["stlt"].clone()
            bad = 0
            for _ in range(reps if B == 4096 else 4 * reps):
                out = model(batch)["stlt"]
                bad += int(not torch.equal(out, ref))
        torch.cuda.synchronize()
        print(f"B={B} {precision}: {bad} mismatching repetitions, finite={bool(torch.isfinite(ref).all())}")
        assert bad == 0

ccfg = stlt_b200.CacnfModelConfig(num_classes=174, unique_categories=4)
cm = stlt_b200.Cacnf(ccfg)
cm.load_state_dict(random_state_dict(cm.state_dict(), seed=1))
cm = cm.cuda()
cm.train(False)
batch = {k: v.cuda() for k, v in make_batch(512, "something", ragged=True, seed=5).items()}
batch["video_features"] = make_appearance_features(512, seed=6).cuda()
with torch.no_grad():
    ref = cm(batch)["ensemble"].clone()
    bad = sum(int(not torch.equal(cm(batch)["ensemble"], ref)) for _ in range(reps))
print(f"cacnf: {bad} mismatching repetitions")
assert bad == 0

tcfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4)
tm = stlt_b200.Stlt(tcfg, precision="bf16").cuda()
tm.train(True)
stepper = stlt_b200.FusedTrainStep(tm, lr=1e-4)
tb = {k: v.cuda() for k, v in make_batch(512, "something", ragged=True, seed=7).items()}
tb["labels"] = (torch.arange(512) % 174).cuda()
losses = torch.stack([stepper.step(tb) for _ in range(reps)]).cpu()
print(f"train: loss {float(losses[0]):.3f} -> {float(losses[-1]):.3f}, finite={bool(torch.isfinite(losses).all())}")
assert torch.isfinite(losses).all() and losses[-1] < losses[0]
print("SOAK_OK")
