"""CPU tests: the oracle (oracle/stlt_oracle.py) against the golden fixtures produced by the
unmodified reference (oracle/make_golden.py), and against the live reference when it is mounted."""
import json
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import stlt_oracle as O
from tests.util import golden_model_case, load_golden, nerr

REFERENCE_SRC = Path("/root/reference/src")


def test_fix_box_matches_reference_golden():
    g = load_golden("fix_box.npz")
    raw, sizes = g["raw"], g["sizes"]
    for i in range(raw.shape[0]):
        w, h = int(sizes[i, 0]), int(sizes[i, 1])
        fixed = O.fix_box([float(v) for v in raw[i]], (h, w))
        assert fixed == g["fixed"][i].tolist(), i
        norm = O.normalize_box(fixed, w, h)
        assert norm.dtype == torch.float32
        assert np.array_equal(norm.numpy().view(np.uint32), g["normalized"][i].view(np.uint32)), i
    assert O.fix_box([-3.2, 500.9, 10.5, 10.5], (240, 427)) == [0, 10, 10, 239]


def test_prepare_padded_matches_scalar_fix_box():
    g = load_golden("fix_box.npz")
    n = g["raw"].shape[0]
    raw = torch.from_numpy(g["raw"]).view(n, 1, 1, 4).expand(n, 1, 2, 4).contiguous()
    cats = torch.tensor([[[3, 2]]]).expand(n, 1, 2).contiguous()
    out = O.prepare_padded(raw, torch.from_numpy(g["sizes"]), cats, torch.full((n, 1), 2))
    got = out["boxes"][:, 0, 1, :].numpy()
    assert np.array_equal(got.view(np.uint32), g["normalized"].view(np.uint32))
    assert torch.equal(out["boxes"][:, 0, 0, :], torch.tensor([0.0, 0.0, 1.0, 1.0]).expand(n, 4))


@pytest.mark.parametrize("dataset", ["something", "action_genome"])
def test_collate_restatement_matches_reference_golden(dataset):
    g = load_golden(f"collate_{dataset}.npz")
    meta = json.loads(bytes(g["json"]).decode())
    mno = int(g["max_num_objects"])
    samples = []
    for video in meta["videos"]:
        w, h = meta["sizes"][video["id"]]
        samples.append(O.build_sample(video, w, h, dataset, mno))
    batch = O.collate(samples, dataset, mno)
    for key in ("categories", "frame_types", "lengths", "src_key_padding_mask_boxes", "src_key_padding_mask_frames"):
        assert np.array_equal(batch[key].numpy(), g[key]), key
    assert np.array_equal(batch["boxes"].numpy().view(np.uint32), g["boxes"].view(np.uint32))
    if dataset == "action_genome":
        assert np.array_equal(batch["scores"].numpy().view(np.uint32), g["scores"].view(np.uint32))
    else:
        assert "scores" not in batch and "scores" not in g


@pytest.mark.parametrize("layout", ["something", "action_genome"])
def test_forward_matches_reference_golden(layout):
    cfg, sd, batch, g = golden_model_case(layout)
    with torch.no_grad():
        taps = O.stlt_forward(sd, batch, return_taps=True)
    assert nerr(taps["stlt"], torch.from_numpy(g["logits"])) < 2e-5
    assert nerr(taps["embed"][0], torch.from_numpy(g["embed_b0"])) < 1e-5
    assert nerr(taps["frames"], torch.from_numpy(g["frames"])) < 2e-5
    # padded positions hold arbitrary finite values in both; compare valid frames only
    lengths = batch["lengths"]
    for b in range(lengths.shape[0]):
        n = int(lengths[b])
        assert nerr(taps["temporal"][b, :n], torch.from_numpy(g["temporal"][b, :n])) < 2e-5
    valid = ~batch["src_key_padding_mask_boxes"][0]
    assert nerr(taps["spatial"][0][valid], torch.from_numpy(g["spatial_b0"])[valid]) < 2e-5


@pytest.mark.parametrize("name,frames,slots", [("stlt_long_99", 100, 3), ("stlt_long_255", 256, 3), ("stlt_wide_40", 17, 41)])
def test_forward_matches_reference_golden_on_long_sequences(name, frames, slots):
    """The oracle is also pinned where the GPU path switches attention kernels: 100 and 256 frame tokens per video
    (the reference's position table ends at 256, models.py:88-96) and 41 slots per frame; logits of the unmodified
    reference module."""
    import stlt_b200
    from stlt_b200.synthetic import random_state_dict
    from tests.util import weights_checksum
    g = load_golden(f"{name}.npz")
    cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=2, num_temporal_layers=2)
    torch.manual_seed(0)
    sd = random_state_dict(stlt_b200.Stlt(cfg).state_dict(), seed=int(g["weight_seed"]))
    assert abs(weights_checksum(sd) - float(g["weights_checksum"])) < 1e-6 * float(g["weights_checksum"])
    batch = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in_")}
    assert tuple(batch["categories"].shape[1:]) == (frames, slots)
    with torch.no_grad():
        got = O.stlt_forward(sd, batch, num_spatial_layers=2, num_temporal_layers=2)
    assert nerr(got, torch.from_numpy(g["logits"])) < 2e-5


def test_forward_fp64_agrees_with_fp32():
    cfg, sd, batch, g = golden_model_case("something")
    with torch.no_grad():
        l64 = O.stlt_forward(sd, batch, dtype=torch.float64)
    assert nerr(l64, torch.from_numpy(g["logits"])) < 1e-5


def test_padding_payload_does_not_change_logits():
    """padded slots / frames never reach the logits."""
    cfg, sd, batch, g = golden_model_case("something")
    scr = {k: v.clone() for k, v in batch.items()}
    pad_slots = batch["src_key_padding_mask_boxes"]
    scr["boxes"][pad_slots] = 0.37
    with torch.no_grad():
        a = O.stlt_forward(sd, batch)
        b = O.stlt_forward(sd, scr)
    assert torch.equal(a, b)


@pytest.mark.skipif(not REFERENCE_SRC.exists(), reason="reference checkout not mounted")
def test_live_reference_default_init_and_dataset():
    sys.path.insert(0, str(REFERENCE_SRC))
    for name in ("h5py", "ffmpeg"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from modelling.configs import StltModelConfig
    from modelling.models import Stlt
    from stlt_b200.synthetic import make_batch
    torch.manual_seed(0)
    ref = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    ref.train(False)
    batch = make_batch(5, "something", ragged=True, seed=9)
    with torch.no_grad():
        want = ref({k: v.clone() for k, v in batch.items()})["stlt"]
        got = O.stlt_forward(ref.state_dict(), batch)
    assert nerr(got, want) < 2e-5


def test_oracle_charades_map_matches_reference_golden():
    """oracle.charades_map vs the reference's charades_map output (tests/golden/charades_map.npz)."""
    import numpy as np
    from oracle import stlt_oracle
    from tests.util import load_golden
    g = load_golden("charades_map.npz")
    pred = torch.sigmoid(torch.from_numpy(g["logits"])).numpy().astype(np.float64)
    m_ap, aps = stlt_oracle.charades_map(pred, g["labels"])
    assert np.isnan(m_ap) and np.isnan(g["map"])  # class 5 has no positives: np.mean propagates the nan
    assert np.allclose(aps, g["aps"], rtol=1e-12, atol=0, equal_nan=True)
    m_ap2, aps2 = stlt_oracle.charades_map(pred, g["labels2"])
    assert abs(m_ap2 - float(g["map2"])) < 1e-12 and np.allclose(aps2, g["aps2"], rtol=1e-12)
