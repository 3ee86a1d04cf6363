"""Small training step (dropout on) and CACNF forward for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tests/sanitizer_train_smoke.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import stlt_b200  # noqa: E402
from stlt_b200.synthetic import make_appearance_features, make_batch, random_state_dict  # noqa: E402
from stlt_b200.training import FusedTrainStep  # noqa: E402

cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=2, num_temporal_layers=2)
torch.manual_seed(0)
model = stlt_b200.Stlt(cfg, precision="bf16")
model.load_state_dict(random_state_dict(model.state_dict(), seed=1))
model = model.to("cuda")
model.train(True)
batch = {k: v.cuda() for k, v in make_batch(3, "something", ragged=True, seed=2).items()}
batch["labels"] = torch.tensor([3, 50, 173], device="cuda")
stepper = FusedTrainStep(model, lr=1e-4)
for _ in range(2):
    loss = stepper.step(batch)
torch.cuda.synchronize()
print("train loss", float(loss))
out = model(batch)["stlt"]  # autograd path
torch.nn.functional.cross_entropy(out, batch["labels"]).backward()
torch.cuda.synchronize()
print("autograd grad norm", float(sum(p.grad.norm() ** 2 for p in model.parameters() if p.grad is not None) ** 0.5))

ccfg = stlt_b200.CacnfModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=1, num_temporal_layers=1,
                                  num_appearance_layers=1, num_fusion_layers=1)
cm = stlt_b200.Cacnf(ccfg)
cm.load_state_dict(random_state_dict(cm.state_dict(), seed=3))
cm = cm.to("cuda")
cm.train(False)
with torch.no_grad():
    res = cm({**batch, "video_features": make_appearance_features(3, seed=4).cuda()})
torch.cuda.synchronize()
print("cacnf", float(res["ensemble"].abs().max()))
print("SANITIZER_TRAIN_SMOKE_DONE")
