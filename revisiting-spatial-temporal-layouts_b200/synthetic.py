"""Seeded synthetic inputs for the STLT path: layout batches of the Something-Else / Action-Genome
shapes and independently re-drawn model weights. Used by tests, smoke and bench;
there is no dataset or checkpoint access in this environment.

Layout rules follow the reference data pipeline (src/modelling/datasets.py:52-125,243-288):
slot 0 of every frame is the CLS object (box [0,0,1,1], score 1), valid frames occupy
[0, len-1), the "extract" frame sits at len-1, padded frames follow with categories [cls, 0, ...],
slot-0 box [0,0,1,1] and frame type 0.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .configs import ACTION_GENOME, SOMETHING_ELSE

LAYOUTS = {"something": SOMETHING_ELSE, "action_genome": ACTION_GENOME}
DEFAULT_MAX_OBJECTS = {"something": 4, "action_genome": 10}
VIDEO_SIZES = ((427, 240), (320, 240), (240, 427))  # (width, height)


def make_batch(batch_size: int, layout: str = "something", ragged: bool = True, seed: int = 0,
               num_frames: int = 16, max_objects: int | None = None) -> Dict[str, torch.Tensor]:
    """CPU batch dict with the keys the reference collater produces (minus video_id / labels)."""
    spec = LAYOUTS[layout]
    ft = spec["frame_types"]
    max_objects = DEFAULT_MAX_OBJECTS[layout] if max_objects is None else max_objects
    g = torch.Generator().manual_seed(seed)
    B, L, S = batch_size, num_frames + 1, max_objects + 1
    if ragged:
        lengths = torch.randint(2, L + 1, (B,), generator=g)
        n_obj = torch.randint(0, max_objects + 1, (B, L), generator=g)
    else:
        lengths = torch.full((B,), L, dtype=torch.int64)
        n_obj = torch.full((B, L), max_objects, dtype=torch.int64)
    if B > 0:
        lengths[0] = L  # the collater pads to the longest sample, so one sample has full length
    frame_idx = torch.arange(L).unsqueeze(0)
    valid = frame_idx < (lengths - 1).unsqueeze(1)
    is_extract = frame_idx == (lengths - 1).unsqueeze(1)
    is_pad = frame_idx >= lengths.unsqueeze(1)
    n_obj = n_obj * valid
    slot = torch.arange(S).view(1, 1, S)
    obj_mask = (slot >= 1) & (slot <= n_obj.unsqueeze(-1))
    ids = torch.tensor(spec["object_ids"])
    obj_ids = ids[torch.randint(0, len(ids), (B, L, S), generator=g)]
    categories = torch.where(obj_mask, obj_ids, torch.zeros_like(obj_ids))
    categories[:, :, 0] = spec["cls_id"]
    frame_types = torch.full((B, L), ft["regular"], dtype=torch.int64)
    frame_types[valid & (n_obj == 0)] = ft["empty"]
    frame_types[is_extract] = ft["extract"]
    frame_types[is_pad] = ft["pad"]
    boxes = torch.rand((B, L, S, 4), generator=g) * obj_mask.unsqueeze(-1)
    boxes[:, :, 0, :] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    batch = {
        "categories": categories,
        "boxes": boxes.to(torch.float32),
        "frame_types": frame_types,
        "lengths": lengths.to(torch.int64),
        "src_key_padding_mask_boxes": categories == 0,
        "src_key_padding_mask_frames": frame_types == ft["pad"],
    }
    if spec["scores"]:
        scores = (0.5 + 0.5 * torch.rand((B, L, S), generator=g)) * obj_mask
        scores[:, :, 0] = 1.0
        batch["scores"] = scores.to(torch.float32)
    return batch


def make_raw_boxes(categories: torch.Tensor, seed: int = 0):
    """Pixel boxes (float64, as parsed from the dataset JSON) for the object slots of a padded
    batch, including out-of-range, swapped and degenerate ones that exercise fix_box
    (src/utils/data_utils.py:205-231). Returns (raw_boxes [B,L,S,4] f64, video_sizes [B,2] i64)."""
    g = torch.Generator().manual_seed(seed)
    B, L, S = categories.shape
    sizes = torch.tensor(VIDEO_SIZES, dtype=torch.int64)[torch.randint(0, len(VIDEO_SIZES), (B,), generator=g)]
    W = sizes[:, 0].view(B, 1, 1).to(torch.float64)
    H = sizes[:, 1].view(B, 1, 1).to(torch.float64)
    u = torch.rand((B, L, S, 4), generator=g, dtype=torch.float64) * 1.3 - 0.15  # some < 0 and > size
    raw = torch.stack([u[..., 0] * W, u[..., 1] * H, u[..., 2] * W, u[..., 3] * H], dim=-1)
    kind = torch.randint(0, 10, (B, L, S), generator=g)
    raw = torch.where((kind == 0).unsqueeze(-1), raw.round(), raw)              # integer coordinates
    raw[..., 2] = torch.where(kind == 1, raw[..., 0], raw[..., 2])              # x1 == x2
    raw[..., 3] = torch.where(kind == 2, raw[..., 1], raw[..., 3])              # y1 == y2
    raw = torch.where((kind == 3).unsqueeze(-1), torch.zeros_like(raw), raw)    # all-zero box
    raw = torch.where((kind == 4).unsqueeze(-1), raw * 4.0, raw)                # far outside the frame
    return raw, sizes


def random_state_dict(model_state: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Every tensor independently re-drawn (default-init encoder layers are identical clones, so a
    default-init parity test cannot see a layer-indexing bug — SURVEY.md §7.2-6): matrices
    N(0, (0.7/sqrt(fan_in))^2), embedding tables N(0,1) incl. padding rows, LayerNorm weights
    1 + 0.1 N(0,1), all biases 0.02 N(0,1). ``model_state`` provides names / shapes / dtypes."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, ref in model_state.items():
        if not ref.is_floating_point():
            out[name] = ref.clone()
            continue
        shape = tuple(ref.shape)
        if name.endswith("embeddings.weight") and "score" not in name and "box_embedding" not in name \
                or name.endswith("frame_type_embedding.weight"):
            t = torch.randn(shape, generator=g)
        elif ("norm" in name or name.endswith(".ln.weight")) and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("cls_token") or name.endswith("pos_embed"):
            t = 0.5 * torch.randn(shape, generator=g)
        elif len(shape) > 2:  # Conv3d 1x1x1 projector of the CACNF appearance branch
            t = torch.randn(shape, generator=g) * (0.7 / math.sqrt(shape[1] * math.prod(shape[2:])))
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif len(shape) == 2:
            t = torch.randn(shape, generator=g) * (0.7 / math.sqrt(shape[1]))
        else:
            t = torch.randn(shape, generator=g)
        out[name] = t.to(ref.dtype)
    return out


def make_appearance_features(batch_size: int, seed: int = 0, channels: int = 2048, t: int = 2, hw: int = 4) -> torch.Tensor:
    """Synthetic stand-in for Resnet3D.forward_features (reference models.py:219-220): post-ReLU
    activations [B, 2048, T', H', W'] of a 32-frame 112x112 clip (2 x 4 x 4 positions)."""
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn((batch_size, channels, t, hw, hw), generator=g)).to(torch.float32)


def make_layout_dataset(dataset: str = "something", n_videos: int = 64, seed: int = 0, dense: bool = True,
                        num_frames: int = 16, max_objects: int | None = None):
    """Synthetic RAW layout dataset in the JSON schema the reference reads (written by its
    src/create_something_datasets.py:18-34 / create_action_genome_datasets.py): a list of
    ``{"id", "frames": [{"frame_objects": [{"category", "x1", "y1", "x2", "y2", "score"}]}]}`` plus
    ``videoid2size`` ``{id: [width, height]}`` — the input of ``LayoutStore`` here and of ``StltDataset`` there.
    Pixel boxes include out-of-frame, swapped and degenerate ones so that fix_box has work to do.
    ``dense``: every video has ``num_frames`` frames carrying ``max_objects`` confident detections (the throughput
    workload); otherwise frame / object counts and scores vary. Returns (videos, videoid2size)."""
    import random
    rng = random.Random(seed)
    spec = LAYOUTS[dataset]
    max_objects = DEFAULT_MAX_OBJECTS[dataset] if max_objects is None else max_objects
    names = {"something": ["hand", "object"],
             "action_genome": None}[dataset]
    if names is None:
        from .data import _AG_NAMES
        names = _AG_NAMES[2:]
    videos, sizes = [], {}
    for v in range(n_videos):
        vid = f"{dataset}_{seed}_{v}"
        w, h = VIDEO_SIZES[rng.randrange(len(VIDEO_SIZES))]
        sizes[vid] = [w, h]
        frames = []
        for _ in range(num_frames if dense else rng.randint(1, 2 * num_frames)):
            objs = []
            for _ in range(max_objects if dense else rng.randint(0, max_objects)):
                x1, x2 = rng.uniform(-0.1, 1.1) * w, rng.uniform(-0.1, 1.1) * w
                y1, y2 = rng.uniform(-0.1, 1.1) * h, rng.uniform(-0.1, 1.1) * h
                kind = rng.randrange(10)
                if kind == 0:
                    x2 = x1
                elif kind == 1:
                    y2 = y1
                elif kind == 2:
                    x1, y1, x2, y2 = round(x1), round(y1), round(x2), round(y2)
                objs.append({"category": names[rng.randrange(len(names))], "x1": x1, "y1": y1, "x2": x2, "y2": y2,
                             "score": rng.uniform(0.5, 1.0) if dense else rng.uniform(0.2, 1.0)})
            frames.append({"frame_objects": objs})
        videos.append({"id": vid, "frames": frames})
    return videos, sizes
