"""CPU tests of the host-side logic: C-ABI surface, drop-in module contract, synthetic inputs,
batch sharding over a world_size-2 gloo group. No GPU compute is invoked here."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

import stlt_b200
from stlt_b200 import Stlt, StltModelConfig
from stlt_b200 import lib as L
from stlt_b200.sharding import shard_batch, shard_bounds
from stlt_b200.synthetic import make_batch, make_raw_boxes, random_state_dict

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "stlt_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stlt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = L.load_library()
    names = declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/stlt_b200.h but not exported"
    assert sorted(L.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load_library()
    dims = L.StltDims(768, 12, 4, 8, 4, 174, 256, 5, 1e-12, 1e-5)
    handle = ctypes.c_void_p()
    rc = lib.stlt_create(ctypes.byref(dims), ctypes.byref(handle))
    assert rc == L.STLT_ERR_CUDA
    assert b"CUDA" in lib.stlt_last_error(None) or b"device" in lib.stlt_last_error(None)
    model = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    model.train(False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(make_batch(2, "something"))


def test_create_rejects_unsupported_dims():
    lib = L.load_library()
    dims = L.StltDims(512, 8, 4, 8, 4, 174, 256, 5, 1e-12, 1e-5)
    handle = ctypes.c_void_p()
    assert lib.stlt_create(ctypes.byref(dims), ctypes.byref(handle)) == L.STLT_ERR_INVALID
    with pytest.raises(ValueError):
        Stlt(StltModelConfig(num_classes=10, unique_categories=4, hidden_size=512))


def test_state_dict_contract():
    """174 entries, reference names/shapes/dtypes."""
    m = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    sd = m.state_dict()
    assert len(sd) == 174
    assert sum(p.numel() for p in m.parameters()) == 93_080_238
    bfe = "backbone.frames_embeddings."
    assert sd[bfe + "position_ids"].dtype == torch.int64 and tuple(sd[bfe + "position_ids"].shape) == (1, 256)
    assert tuple(sd[bfe + "layout_embedding.category_box_embeddings.category_embeddings.weight"].shape) == (4, 768)
    assert tuple(sd[bfe + "layout_embedding.category_box_embeddings.box_embedding.weight"].shape) == (768, 4)
    assert tuple(sd[bfe + "layout_embedding.category_box_embeddings.score_embeddings.weight"].shape) == (768, 1)
    assert sum(k.startswith(bfe + "layout_embedding.encoder_layer.") for k in sd) == 12  # orphan layer
    for stack, n in ((bfe + "layout_embedding.transformer.layers.", 4), ("backbone.transformer.layers.", 8)):
        for i in range(n):
            p = f"{stack}{i}."
            assert tuple(sd[p + "self_attn.in_proj_weight"].shape) == (2304, 768)
            assert tuple(sd[p + "self_attn.out_proj.weight"].shape) == (768, 768)
            assert tuple(sd[p + "linear1.weight"].shape) == (3072, 768)
            assert tuple(sd[p + "linear2.weight"].shape) == (768, 3072)
    assert tuple(sd["prediction_head.fc2.weight"].shape) == (174, 768)
    assert tuple(sd[bfe + "frame_type_embedding.weight"].shape) == (5, 768)
    # clones at init, padding rows zero at init (nn.TransformerEncoder / nn.Embedding semantics)
    l0, l3 = (bfe + f"layout_embedding.transformer.layers.{i}.linear1.weight" for i in (0, 3))
    assert torch.equal(sd[l0], sd[l3]) and torch.equal(sd[l0], sd[bfe + "layout_embedding.encoder_layer.linear1.weight"])
    assert sd[bfe + "frame_type_embedding.weight"][0].abs().sum() == 0
    assert m.logit_names == ("stlt",)
    assert len(m.backbone.state_dict()) == 168
    m2 = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    m2.load_state_dict(random_state_dict(sd, seed=5), strict=True)
    assert m.train(False) is m and not m.training


def test_training_mode_is_not_silently_run():
    m = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    with pytest.raises(NotImplementedError):
        m(make_batch(1, "something"))


@pytest.mark.parametrize("layout,S,has_scores", [("something", 5, False), ("action_genome", 11, True)])
def test_synthetic_batch_follows_reference_layout_rules(layout, S, has_scores):
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    b = make_batch(16, layout, ragged=True, seed=1)
    B, L, S_ = b["categories"].shape
    assert (L, S_) == (17, S) and ("scores" in b) == has_scores
    assert (b["categories"][:, :, 0] == spec["cls_id"]).all()
    assert torch.equal(b["boxes"][:, :, 0], torch.tensor([0.0, 0.0, 1.0, 1.0]).expand(B, L, 4))
    for i in range(B):
        n = int(b["lengths"][i])
        assert b["frame_types"][i, n - 1] == spec["frame_types"]["extract"]
        assert (b["frame_types"][i, n:] == 0).all() and (b["frame_types"][i, :n] != 0).all()
        assert (b["categories"][i, n - 1:, 1:] == 0).all()
    assert (b["boxes"][b["categories"] == 0] == 0).all()
    assert torch.equal(b["src_key_padding_mask_boxes"], b["categories"] == 0)
    dense = make_batch(4, layout, ragged=False)
    assert (dense["lengths"] == 17).all() and (dense["categories"][:, :16] != 0).all()
    raw, sizes = make_raw_boxes(b["categories"], seed=2)
    assert raw.dtype == torch.float64 and tuple(sizes.shape) == (B, 2)


def test_shard_bounds_cover_batch():
    for B in (0, 1, 7, 8, 4096, 4099):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["STLT_ROOT"])
from stlt_b200.sharding import shard_batch, shard_bounds, gather_logits
from stlt_b200.synthetic import make_batch
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
batch = make_batch(7, "something", seed=4)
local = shard_batch(batch, world, rank)
lo, hi = shard_bounds(7, world, rank)
assert local["categories"].shape[0] == hi - lo
# stand-in for per-rank logits: a deterministic function of the local inputs
fake = local["boxes"].sum(dim=(1, 2)).float()          # [b_local, 4]
full = gather_logits(fake, 7)
want = batch["boxes"].sum(dim=(1, 2)).float()
assert torch.equal(full, want), (rank, full, want)
dist.barrier()
if rank == 0:
    print("SHARD_OK")
"""


def test_two_rank_gloo_sharding_roundtrip(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, STLT_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "SHARD_OK" in res.stdout
