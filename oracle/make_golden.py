"""Generates tests/golden/* by running the UNMODIFIED reference from /root/reference on CPU.

TEST INFRASTRUCTURE. Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

The reference ships no tests or golden vectors, so these fixtures — outputs of the
reference's own ``fix_box``, ``StltDataset`` + ``StltCollater`` and ``Stlt`` module on seeded
inputs — are what pins the oracle (oracle/stlt_oracle.py) and, through it, the CUDA path.
Inputs come from the seeded generators in the package (synthetic.py) so tests can rebuild them;
model weights are regenerated from a seed at test time (a checksum is stored here).
"""
from __future__ import annotations

import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REFERENCE_SRC = Path("/root/reference/src")
GOLDEN = ROOT / "tests" / "golden"


def import_reference():
    if not REFERENCE_SRC.exists():
        raise SystemExit("/root/reference is not available here")
    sys.path.insert(0, str(REFERENCE_SRC))
    for name in ("h5py", "ffmpeg"):  # imported by the reference data modules, unused on this path
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    from modelling import configs, datasets, models  # noqa
    from utils import data_utils  # noqa
    return configs, datasets, models, data_utils


def weights_checksum(sd) -> float:
    return float(sum(v.double().abs().sum().item() for v in sd.values() if v.is_floating_point()))


def golden_fix_box(data_utils):
    g = torch.Generator().manual_seed(11)
    n = 4096
    sizes = torch.tensor([(427, 240), (320, 240), (240, 427), (1280, 720)])[torch.randint(0, 4, (n,), generator=g)]
    u = torch.rand((n, 4), generator=g, dtype=torch.float64) * 1.4 - 0.2
    raw = u * torch.stack([sizes[:, 0], sizes[:, 1], sizes[:, 0], sizes[:, 1]], dim=1).double()
    kind = torch.randint(0, 8, (n,), generator=g)
    raw = torch.where((kind == 0).unsqueeze(1), raw.round(), raw)
    raw[:, 2] = torch.where(kind == 1, raw[:, 0], raw[:, 2])
    raw[:, 3] = torch.where(kind == 2, raw[:, 1], raw[:, 3])
    raw = torch.where((kind == 3).unsqueeze(1), torch.zeros_like(raw), raw)
    raw[0] = torch.tensor([-3.2, 500.9, 10.5, 10.5], dtype=torch.float64)  # SURVEY.md A.2 example
    sizes[0] = torch.tensor([427, 240])
    fixed = torch.zeros((n, 4), dtype=torch.int64)
    norm = torch.zeros((n, 4), dtype=torch.float32)
    for i in range(n):
        w, h = int(sizes[i, 0]), int(sizes[i, 1])
        box = data_utils.fix_box([float(v) for v in raw[i]], (h, w))       # reference call
        fixed[i] = torch.tensor(box)
        video_size = torch.tensor([w, h]).repeat(2)                       # datasets.py:54
        norm[i] = torch.tensor(box) / video_size                          # datasets.py:82
    assert fixed[0].tolist() == [0, 10, 10, 239]
    np.savez_compressed(GOLDEN / "fix_box.npz", raw=raw.numpy(), sizes=sizes.numpy(),
                        fixed=fixed.numpy(), normalized=norm.numpy())
    print("fix_box.npz", n)


def synth_dataset_json(dataset: str, n_videos: int, seed: int):
    """A small dataset in the on-disk JSON schema the reference reads (datasets.py:35-37)."""
    from oracle.stlt_oracle import DATASETS
    g = torch.Generator().manual_seed(seed)
    names = [n for n in DATASETS[dataset]["category2id"] if n not in ("pad", "cls")]
    max_obj = 4 if dataset == "something" else 6
    videos, sizes = [], {}
    for v in range(n_videos):
        vid = f"video{v}"
        w, h = [(427, 240), (320, 240), (240, 427)][int(torch.randint(0, 3, (1,), generator=g))]
        sizes[vid] = [w, h]
        n_frames = int(torch.randint(1, 40, (1,), generator=g))
        frames = []
        for _ in range(n_frames):
            objs = []
            for _ in range(int(torch.randint(0, max_obj + 1, (1,), generator=g))):
                c = torch.rand(4, generator=g, dtype=torch.float64) * 1.3 - 0.15
                objs.append({
                    "category": names[int(torch.randint(0, len(names), (1,), generator=g))],
                    "x1": float(c[0] * w), "y1": float(c[1] * h), "x2": float(c[2] * w), "y2": float(c[3] * h),
                    "score": float(torch.rand(1, generator=g, dtype=torch.float64) * 0.7 + 0.3)
                    if dataset == "action_genome" else 1.0,
                })
            frames.append({"frame_objects": objs})
        video = {"id": vid, "frames": frames}
        if dataset == "something":
            video["template"] = "Doing [something]"
        else:
            video["actions"] = ["c001", "c017"]
        videos.append(video)
    labels = {"Doing something": "7"} if dataset == "something" else {f"c{i:03d}": i for i in range(157)}
    return videos, labels, sizes


def golden_dataset(configs, datasets, dataset: str, seed: int):
    videos, labels, sizes = synth_dataset_json(dataset, n_videos=6, seed=seed)
    with tempfile.TemporaryDirectory() as td:
        paths = {}
        for name, obj in (("dataset", videos), ("labels", labels), ("sizes", sizes)):
            paths[name] = str(Path(td) / f"{name}.json")
            json.dump(obj, open(paths[name], "w"))
        cfg = configs.DataConfig(dataset_name=dataset, dataset_path=paths["dataset"], labels_path=paths["labels"],
                                 videoid2size_path=paths["sizes"], videos_path=None, train=False)
        ds = datasets.StltDataset(cfg)                 # reference, discovers max_num_objects
        batch = datasets.StltCollater(cfg)([ds[i] for i in range(len(ds))])
    out = {
        "json": np.frombuffer(json.dumps({"videos": videos, "sizes": sizes}).encode(), dtype=np.uint8),
        "max_num_objects": np.int64(cfg.max_num_objects),
        "categories": batch["categories"].numpy(), "boxes": batch["boxes"].numpy(),
        "frame_types": batch["frame_types"].numpy(), "lengths": batch["lengths"].numpy(),
        "src_key_padding_mask_boxes": batch["src_key_padding_mask_boxes"].numpy(),
        "src_key_padding_mask_frames": batch["src_key_padding_mask_frames"].numpy(),
    }
    if "scores" in batch:
        out["scores"] = batch["scores"].numpy()
    np.savez_compressed(GOLDEN / f"collate_{dataset}.npz", **out)
    print(f"collate_{dataset}.npz", tuple(batch["categories"].shape), "max_num_objects", cfg.max_num_objects)


def golden_model(configs, models, layout: str, batch_size: int, weight_seed: int, batch_seed: int):
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    cfg = configs.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"])
    torch.manual_seed(0)
    ref = models.Stlt(cfg)              # the unmodified reference module
    ref.train(False)
    sd = random_state_dict(ref.state_dict(), seed=weight_seed)
    ref.load_state_dict(sd, strict=True)
    batch = make_batch(batch_size, layout=layout, ragged=True, seed=batch_seed)
    B, L, S = batch["categories"].shape
    taps = {}
    be = ref.backbone.frames_embeddings
    hooks = [
        be.layout_embedding.category_box_embeddings.register_forward_hook(lambda m, i, o: taps.__setitem__("embed", o.detach().clone())),
        be.layout_embedding.transformer.register_forward_hook(
            lambda m, i, o: taps.__setitem__("spatial", o.detach().transpose(0, 1).reshape(B, L, S, -1).clone())),
        be.register_forward_hook(lambda m, i, o: taps.__setitem__("frames", o.detach().clone())),
        ref.backbone.register_forward_hook(lambda m, i, o: taps.__setitem__("temporal", o.detach().transpose(0, 1).clone())),
    ]
    with torch.no_grad():
        logits = ref({k: v.clone() for k, v in batch.items()})["stlt"]
    for h in hooks:
        h.remove()
    out = {
        "layout": np.array(layout), "batch_size": np.int64(batch_size), "weight_seed": np.int64(weight_seed),
        "batch_seed": np.int64(batch_seed), "weights_checksum": np.float64(weights_checksum(sd)),
        "logits": logits.numpy(),
        "embed_b0": taps["embed"][0].numpy(), "spatial_b0": taps["spatial"][0].numpy(),
        "frames": taps["frames"].numpy(), "temporal": taps["temporal"].numpy(),
    }
    for k, v in batch.items():
        out["in_" + k] = v.numpy()
    np.savez_compressed(GOLDEN / f"stlt_{layout}.npz", **out)
    print(f"stlt_{layout}.npz logits", tuple(logits.shape), "max|logit|", float(logits.abs().max()))


def golden_model_long(configs, models, frames: int, batch_size: int, weight_seed: int, batch_seed: int,
                      max_objects: int = 2, name: str = ""):
    """Long-sequence case: more sampled frames than the CLIs' default 16 (the position table allows 256,
    models.py:88-96); 2 spatial + 2 temporal layers keep the fixture and the CPU test small. Logits only."""
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    spec = stlt_b200.SOMETHING_ELSE
    cfg = configs.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"],
                                  num_spatial_layers=2, num_temporal_layers=2)
    torch.manual_seed(0)
    ref = models.Stlt(cfg)              # the unmodified reference module
    ref.train(False)
    sd = random_state_dict(ref.state_dict(), seed=weight_seed)
    ref.load_state_dict(sd, strict=True)
    batch = make_batch(batch_size, layout="something", ragged=True, seed=batch_seed, num_frames=frames,
                       max_objects=max_objects)
    with torch.no_grad():
        logits = ref({k: v.clone() for k, v in batch.items()})["stlt"]
    out = {"frames": np.int64(frames), "batch_size": np.int64(batch_size), "weight_seed": np.int64(weight_seed),
           "batch_seed": np.int64(batch_seed), "weights_checksum": np.float64(weights_checksum(sd)),
           "logits": logits.numpy()}
    for k, v in batch.items():
        out["in_" + k] = v.numpy()
    name = name or f"stlt_long_{frames}"
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    print(f"{name}.npz logits", tuple(logits.shape), "max|logit|", float(logits.abs().max()))


def golden_training(configs, models, train_utils, layout: str, batch_size: int, weight_seed: int, batch_seed: int,
                    steps: int = 3):
    """Reference training loop body (src/train.py:117-135) with dropout p = 0 on a fixed batch:
    loss, per-tensor gradient norms + a few full gradient tensors of the first step, and the loss /
    per-tensor parameter norms after each AdamW step (clip 5.0, lr 5e-5, wd 1e-3, warm-up schedule)."""
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    cfg = configs.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"],
                                  hidden_dropout_prob=0.0)
    torch.manual_seed(0)
    ref = models.Stlt(cfg)
    sd = random_state_dict(ref.state_dict(), seed=weight_seed)
    ref.load_state_dict(sd, strict=True)
    ref.train(True)
    batch = make_batch(batch_size, layout=layout, ragged=True, seed=batch_seed)
    g = torch.Generator().manual_seed(batch_seed + 100)
    if layout == "something":
        labels = torch.randint(0, spec["num_classes"], (batch_size,), generator=g)
    else:
        labels = (torch.rand((batch_size, spec["num_classes"]), generator=g) < 0.05).float()
    crit = train_utils.Criterion(layout)
    optimizer = torch.optim.AdamW(train_utils.add_weight_decay(ref, 1e-3), lr=5e-5)
    scheduler = train_utils.get_linear_schedule_with_warmup(optimizer, num_warmup_steps=2, num_training_steps=10)
    out = {"layout": np.array(layout), "batch_size": np.int64(batch_size), "weight_seed": np.int64(weight_seed),
           "batch_seed": np.int64(batch_seed), "weights_checksum": np.float64(weights_checksum(sd)),
           "labels": labels.numpy(), "steps": np.int64(steps)}
    names = [n for n, _ in ref.named_parameters()]
    out["param_names"] = np.array(names)
    losses, norms = [], []
    for step in range(steps):
        optimizer.zero_grad()
        logits = ref({k: v.clone() for k, v in batch.items()})
        loss = crit(logits, labels)
        loss.backward()
        if step == 0:
            out["logits0"] = logits["stlt"].detach().numpy()
            out["grad_norms"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0
                                          for _, p in ref.named_parameters()], dtype=np.float64)
            keep = ["prediction_head.fc2.bias", "prediction_head.fc1.bias", "prediction_head.layer_norm.weight",
                    "backbone.frames_embeddings.layout_embedding.category_box_embeddings.category_embeddings.weight",
                    "backbone.frames_embeddings.layout_embedding.category_box_embeddings.box_embedding.weight",
                    "backbone.frames_embeddings.layout_embedding.category_box_embeddings.layer_norm.bias",
                    "backbone.frames_embeddings.frame_type_embedding.weight",
                    "backbone.frames_embeddings.layer_norm.weight",
                    "backbone.frames_embeddings.layout_embedding.transformer.layers.0.self_attn.in_proj_bias",
                    "backbone.frames_embeddings.layout_embedding.transformer.layers.3.norm1.weight",
                    "backbone.transformer.layers.7.linear1.bias", "backbone.transformer.layers.0.norm2.bias"]
            if layout == "action_genome":
                keep.append("backbone.frames_embeddings.layout_embedding.category_box_embeddings.score_embeddings.weight")
            params = dict(ref.named_parameters())
            for k in keep:
                out["grad/" + k] = params[k].grad.detach().numpy().copy()
            out["grad/position_rows"] = params["backbone.frames_embeddings.position_embeddings.weight"].grad[:17].numpy().copy()
            out["grad/sp0_in_proj_rows"] = params[
                "backbone.frames_embeddings.layout_embedding.transformer.layers.0.self_attn.in_proj_weight"].grad[::96, ::8].numpy().copy()
            out["grad/tm7_linear2_rows"] = params["backbone.transformer.layers.7.linear2.weight"].grad[::32, ::64].numpy().copy()
        total = torch.nn.utils.clip_grad_norm_(ref.parameters(), 5.0)
        optimizer.step()
        scheduler.step()
        losses.append(float(loss))
        norms.append(float(total))
        out[f"param_norms_step{step + 1}"] = np.array([float(p.detach().double().norm()) for _, p in ref.named_parameters()])
        out[f"param_sums_step{step + 1}"] = np.array([float(p.detach().double().sum()) for _, p in ref.named_parameters()])
    out["losses"] = np.array(losses)
    out["total_grad_norms"] = np.array(norms)
    np.savez_compressed(GOLDEN / f"train_{layout}.npz", **out)
    print(f"train_{layout}.npz losses", losses, "grad norms", norms)


def golden_cacnf(configs, models, batch_size: int, weight_seed: int, batch_seed: int):
    """The unmodified reference CrossAttentionCentralNetFusion (models.py:504-549) in eval mode with
    Resnet3D.forward_features replaced by synthetic precomputed features (the trunk needs a Kinetics
    checkpoint and pixels, both outside this path; torch.load is stubbed for the constructor only)."""
    import stlt_b200
    from modelling import resnets3d
    from stlt_b200.synthetic import make_appearance_features, make_batch, random_state_dict
    orig_load = torch.load
    torch.load = lambda *a, **k: {"state_dict": resnets3d.generate_model(model_depth=50, n_classes=1139).state_dict()}
    try:
        cfg = configs.MultimodalModelConfig(num_classes=174, unique_categories=4, appearance_num_frames=32,
                                            resnet_model_path="synthetic")
        torch.manual_seed(0)
        ref = models.CrossAttentionCentralNetFusion(cfg)
    finally:
        torch.load = orig_load
    ref.train(False)
    full = ref.state_dict()
    own = {k: v for k, v in full.items() if ".appearance_branch.resnet." not in k}
    sd = random_state_dict(own, seed=weight_seed)
    ref.load_state_dict({**full, **sd}, strict=True)
    batch = make_batch(batch_size, layout="something", ragged=True, seed=batch_seed)
    feats = make_appearance_features(batch_size, seed=batch_seed + 50)
    ref.backbone.appearance_branch.resnet.forward_features = lambda b: feats
    with torch.no_grad():
        out = ref({**{k: v.clone() for k, v in batch.items()}, "video_frames": torch.zeros(batch_size, 1)})
    res = {"batch_size": np.int64(batch_size), "weight_seed": np.int64(weight_seed), "batch_seed": np.int64(batch_seed),
           "weights_checksum": np.float64(weights_checksum(sd)), "num_entries": np.int64(len(own))}
    for k, v in out.items():
        res["logits_" + k] = v.numpy()
    np.savez_compressed(GOLDEN / "cacnf_something.npz", **res)
    print("cacnf_something.npz", {k: float(v.abs().max()) for k, v in out.items()}, "entries", len(own))


def golden_caf_lcf(configs, models, batch_size: int, weight_seed: int, batch_seed: int):
    """The unmodified reference CrossAttentionFusion (CAF) and LateConcatenationFusion (LCF) modules, eval mode,
    ResNet features injected as in golden_cacnf."""
    import stlt_b200
    from modelling import resnets3d
    from stlt_b200.synthetic import make_appearance_features, make_batch, random_state_dict
    batch = make_batch(batch_size, layout="something", ragged=True, seed=batch_seed)
    feats = make_appearance_features(batch_size, seed=batch_seed + 50)
    res = {"batch_size": np.int64(batch_size), "weight_seed": np.int64(weight_seed), "batch_seed": np.int64(batch_seed)}
    for name, cls in (("caf", models.CrossAttentionFusion), ("lcf", models.LateConcatenationFusion)):
        orig_load = torch.load
        torch.load = lambda *a, **k: {"state_dict": resnets3d.generate_model(model_depth=50, n_classes=1139).state_dict()}
        try:
            cfg = configs.MultimodalModelConfig(num_classes=174, unique_categories=4, appearance_num_frames=32,
                                                resnet_model_path="synthetic")
            torch.manual_seed(0)
            ref = cls(cfg)
        finally:
            torch.load = orig_load
        ref.train(False)
        full = ref.state_dict()
        own = {k: v for k, v in full.items() if ".resnet." not in k}
        sd = random_state_dict(own, seed=weight_seed)
        ref.load_state_dict({**full, **sd}, strict=True)
        branch = ref.caf_backbone.appearance_branch if name == "caf" else ref.appearance_branch
        branch.resnet.forward_features = lambda b: feats
        with torch.no_grad():
            out = ref({**{k: v.clone() for k, v in batch.items()}, "video_frames": torch.zeros(batch_size, 1)})
        res[f"logits_{name}"] = out[name].numpy()
        res[f"checksum_{name}"] = np.float64(weights_checksum(sd))
        res[f"entries_{name}"] = np.int64(len(own))
        print(name, float(out[name].abs().max()), len(own))
    np.savez_compressed(GOLDEN / "caf_lcf_something.npz", **res)


def golden_charades_map():
    """The reference's own charades_map (src/utils/evaluation.py:126-132) on seeded scores / multi-hot labels,
    including videos without labels (the -inf fix) and one class without positives (nan)."""
    from utils import evaluation
    if not hasattr(np, "NINF"):  # the reference targets numpy < 2 (evaluation.py:130 uses np.NINF)
        np.NINF = -np.inf
    g = torch.Generator().manual_seed(31)
    n, c = 777, 157
    logits = torch.randn((n, c), generator=g) * 2
    labels = (torch.rand((n, c), generator=g) < 0.04).float()
    labels[::13] = 0          # videos without any action
    labels[:, 5] = 0          # a class that never occurs
    pred = torch.sigmoid(logits).numpy().astype(np.float64)
    m_ap, _, aps = evaluation.charades_map(pred, labels.numpy().astype(np.float64))
    labels2 = labels.clone()
    labels2[:, 5] = (torch.rand(n, generator=g) < 0.1).float()
    labels2[::13] = 0
    m_ap2, _, aps2 = evaluation.charades_map(pred, labels2.numpy().astype(np.float64))
    np.savez_compressed(GOLDEN / "charades_map.npz", logits=logits.numpy(), labels=labels.numpy(), labels2=labels2.numpy(),
                        map=np.float64(m_ap), aps=aps, map2=np.float64(m_ap2), aps2=aps2)
    print("charades_map.npz", m_ap, m_ap2)


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    configs, datasets, models, data_utils = import_reference()
    golden_fix_box(data_utils)
    golden_dataset(configs, datasets, "something", seed=21)
    golden_dataset(configs, datasets, "action_genome", seed=22)
    golden_model(configs, models, "something", batch_size=3, weight_seed=1, batch_seed=3)
    golden_model(configs, models, "action_genome", batch_size=2, weight_seed=2, batch_seed=4)
    golden_model_long(configs, models, frames=99, batch_size=2, weight_seed=14, batch_seed=15)
    golden_model_long(configs, models, frames=255, batch_size=2, weight_seed=16, batch_seed=17)
    # wide frames: 41 slots per frame (the 33..64-token attention tiles on the GPU side)
    golden_model_long(configs, models, frames=16, batch_size=2, weight_seed=18, batch_seed=19, max_objects=40,
                      name="stlt_wide_40")
    golden_cacnf(configs, models, batch_size=3, weight_seed=8, batch_seed=9)
    golden_charades_map()
    golden_caf_lcf(configs, models, batch_size=3, weight_seed=12, batch_seed=13)
    from utils import train_inference_utils as train_utils
    golden_training(configs, models, train_utils, "something", batch_size=4, weight_seed=5, batch_seed=6)
    golden_training(configs, models, train_utils, "action_genome", batch_size=2, weight_seed=6, batch_seed=7)


if __name__ == "__main__":
    main()
