"""Shared helpers for the parity tests."""
from __future__ import annotations

import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def nerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a - b| / max|b| (SURVEY.md §8(d) error metric)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load_golden(name: str):
    z = np.load(GOLDEN / name, allow_pickle=False)
    return {k: z[k] for k in z.files}


def weights_checksum(sd) -> float:
    return float(sum(v.double().abs().sum().item() for v in sd.values() if v.is_floating_point()))


def golden_model_case(layout: str):
    """Rebuilds inputs + weights of a tests/golden/stlt_<layout>.npz case. Returns (cfg, sd, batch, golden)."""
    import stlt_b200
    from stlt_b200.synthetic import random_state_dict
    g = load_golden(f"stlt_{layout}.npz")
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
    torch.manual_seed(0)
    shapes = stlt_b200.Stlt(cfg).state_dict()
    sd = random_state_dict(shapes, seed=int(g["weight_seed"]))
    assert abs(weights_checksum(sd) - float(g["weights_checksum"])) < 1e-6 * float(g["weights_checksum"]), \
        "seeded weights are not reproducible on this machine"
    batch = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in_")}
    return cfg, sd, batch, g


def to_cuda(batch):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
