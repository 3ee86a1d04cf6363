"""Small forward (both precisions) for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import stlt_b200  # noqa: E402
from stlt_b200.synthetic import make_batch, random_state_dict  # noqa: E402

cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=2, num_temporal_layers=2)
torch.manual_seed(0)
model = stlt_b200.Stlt(cfg)
model.load_state_dict(random_state_dict(model.state_dict(), seed=1))
model = model.to("cuda")
model.train(False)
batch = {k: v.cuda() for k, v in make_batch(3, "something", ragged=True, seed=2).items()}
for precision in ("fp32", "bf16"):
    model.precision = precision
    with torch.no_grad():
        out = model(batch)["stlt"]
    torch.cuda.synchronize()
    print(precision, float(out.abs().max()))
print("SANITIZER_SMOKE_DONE")
