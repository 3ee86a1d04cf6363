"""CPU oracle for the STLT layout-encoding forward path.  TEST INFRASTRUCTURE ONLY.

This file is a plain restatement (PyTorch CPU tensor math + pure-Python integer code) of the
algorithm the reference implements for this path. It exists to CHECK the CUDA implementation:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it. The product (``stlt_b200``) never does, and has no CPU fallback.

Parity pinning: the reference ships no tests, golden vectors or checkpoints,
and its arithmetic lives in a third-party dependency (PyTorch, pinned 1.10.1 in poetry.lock:938).
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD
CONTAINER: ``oracle/make_golden.py`` imports the unmodified reference from /root/reference,
runs its ``Stlt`` module, ``StltDataset``/``StltCollater`` and ``fix_box`` on seeded inputs and
commits the results under ``tests/golden/``; ``tests/test_oracle.py`` checks this file against
those fixtures (and against the live reference when /root/reference is present).

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

HIDDEN = 768
HEADS = 12
HEAD_DIM = 64
ENCODER_LN_EPS = 1e-5  # nn.TransformerEncoderLayer default; the config eps is NOT forwarded
                       # (src/modelling/models.py:46-52,118-124)

# ---------------------------------------------------------------------------------------------------
# data side (integer / byte work): src/utils/data_utils.py, src/modelling/datasets.py
# ---------------------------------------------------------------------------------------------------
SOMETHING = {
    "category2id": {"pad": 0, "hand": 1, "object": 2, "cls": 3},  # src/modelling/configs.py:30-37
    "frame2type": {"pad": 0, "start": 1, "regular": 2, "empty": 3, "extract": 4},  # :79-87
    "scores": False,
}
_AG_NAMES = ["pad", "cls", "chair", "book", "medicine", "vacuum", "food", "groceries", "floor",
             "mirror", "closet/cabinet", "doorway", "paper/notebook", "picture", "phone/camera",
             "sofa/couch", "sandwich", "cup/glass/bottle", "towel", "box", "blanket", "television",
             "bag", "refrigerator", "table", "light", "broom", "shoe", "doorknob", "bed", "window",
             "shelf", "door", "pillow", "laptop", "dish", "clothes", "person"]
ACTION_GENOME = {
    "category2id": {n: i for i, n in enumerate(_AG_NAMES)},  # src/modelling/configs.py:38-78
    "frame2type": {"pad": 0, "regular": 1, "extract": 2, "empty": 3},  # :88
    "scores": True,
}
DATASETS = {"something": SOMETHING, "action_genome": ACTION_GENOME}


def fix_box(box: Sequence[float], video_size: Sequence[int]) -> List[int]:
    """src/utils/data_utils.py:205-231. ``video_size`` is (height, width)."""
    b = [max(0, int(v)) for v in box]
    if b[0] > b[2]:
        b[0], b[2] = b[2], b[0]
    if b[1] > b[3]:
        b[1], b[3] = b[3], b[1]
    h, w = video_size
    if b[0] >= w:
        b[0] = w - 1
    if b[1] >= h:
        b[1] = h - 1
    if b[2] >= w:
        b[2] = w - 1
    if b[3] >= h:
        b[3] = h - 1
    if b[0] == b[2] and b[0] == 0:
        b[2] = 1
    if b[1] == b[3] and b[1] == 0:
        b[3] = 1
    if b[0] == b[2]:
        b[0] -= 1
    if b[1] == b[3]:
        b[1] -= 1
    return b


def normalize_box(box: Sequence[int], width: int, height: int) -> torch.Tensor:
    """src/modelling/datasets.py:54,82: int64 tensor / int64 tensor -> fp32 true divide."""
    size = torch.tensor([width, height]).repeat(2)
    return torch.tensor(list(box)) / size


def get_test_layout_indices(coord_nr_frames: int, nr_video_frames: int) -> List[int]:
    """src/utils/data_utils.py:47-56."""
    if nr_video_frames > coord_nr_frames:
        tick = nr_video_frames * 1.0 / coord_nr_frames
        return [int(tick / 2.0 + tick * x) for x in range(coord_nr_frames)]
    return list(range(nr_video_frames))


def build_sample(video: dict, width: int, height: int, dataset: str, max_num_objects: int,
                 layout_num_frames: int = 16, score_threshold: float = 0.5,
                 indices: Optional[List[int]] = None) -> Dict[str, torch.Tensor]:
    """StltDataset.__getitem__ (src/modelling/datasets.py:52-125), test-time frame sampling."""
    cfg = DATASETS[dataset]
    c2i, f2t = cfg["category2id"], cfg["frame2type"]
    frames = video["frames"]
    if indices is None:
        indices = get_test_layout_indices(layout_num_frames, len(frames))
    S = max_num_objects + 1
    boxes, categories, scores, frame_types = [], [], [], []
    for index in indices:
        objs = frames[index]["frame_objects"]
        # frame type is decided on the UNFILTERED list (datasets.py:65-69)
        frame_types.append(f2t["empty"] if len(objs) == 0 else f2t["regular"])
        fb = [torch.tensor([0.0, 0.0, 1.0, 1.0])]
        fc = [c2i["cls"]]
        fs = [1.0]
        for e in objs:
            if e["score"] < score_threshold:
                continue
            b = fix_box([e["x1"], e["y1"], e["x2"], e["y2"]], (height, width))
            fb.append(normalize_box(b, width, height))
            fc.append(c2i[e["category"]])
            fs.append(e["score"])
        while len(fb) != S:
            fb.append(torch.full((4,), 0.0))
            fc.append(0)
            fs.append(0.0)
        categories.append(torch.tensor(fc))
        scores.append(torch.tensor(fs))
        boxes.append(torch.stack(fb, dim=0))
    # extract frame (datasets.py:97-113)
    eb = torch.full((S, 4), 0.0)
    eb[0] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    boxes.append(eb)
    ec = torch.full((S,), 0)
    ec[0] = c2i["cls"]
    categories.append(ec)
    es = torch.full((S,), 0.0)
    es[0] = 1.0
    scores.append(es)
    frame_types.append(f2t["extract"])
    return {
        "categories": torch.stack(categories, dim=0),
        "boxes": torch.stack(boxes, dim=0),
        "scores": torch.stack(scores, dim=0),
        "frame_types": torch.tensor(frame_types),
        "lengths": torch.tensor(len(categories)),
    }


def pad_sequence(sequences: List[torch.Tensor], pad_tensor: torch.Tensor) -> torch.Tensor:
    """src/utils/data_utils.py:93-102."""
    trailing = sequences[0].dim() - 1
    max_len = max(s.size(0) for s in sequences)
    out = pad_tensor.repeat((len(sequences), max_len) + (1,) * trailing)
    for i, t in enumerate(sequences):
        out[i, : t.size(0), ...] = t
    return out


def collate(samples: List[Dict[str, torch.Tensor]], dataset: str, max_num_objects: int) -> Dict[str, torch.Tensor]:
    """StltCollater.__call__ (src/modelling/datasets.py:243-288)."""
    cfg = DATASETS[dataset]
    S = max_num_objects + 1
    batch = {}
    pad_c = torch.full((S,), 0)
    pad_c[0] = cfg["category2id"]["cls"]
    batch["categories"] = pad_sequence([s["categories"] for s in samples], pad_c)
    if cfg["scores"]:
        pad_s = torch.full((S,), 0.0)
        pad_s[0] = 1.0
        batch["scores"] = pad_sequence([s["scores"] for s in samples], pad_s)
    pad_b = torch.full((S, 4), 0.0)
    pad_b[0] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    batch["boxes"] = pad_sequence([s["boxes"] for s in samples], pad_b)
    batch["frame_types"] = pad_sequence([s["frame_types"] for s in samples],
                                        torch.tensor([cfg["frame2type"]["pad"]]))
    batch["lengths"] = torch.stack([s["lengths"] for s in samples], dim=0)
    batch["src_key_padding_mask_boxes"] = batch["categories"] == 0
    batch["src_key_padding_mask_frames"] = batch["frame_types"] == cfg["frame2type"]["pad"]
    return batch


def prepare_padded(raw_boxes: torch.Tensor, video_sizes: torch.Tensor, categories: torch.Tensor,
                   frame_types: torch.Tensor) -> Dict[str, torch.Tensor]:
    """What K0 (stlt_prepare) must produce for an already padded layout: vectorised restatement
    of fix_box + normalisation (data_utils.py:205-231, datasets.py:82) applied to object slots,
    the CLS box [0,0,1,1] in slot 0 (datasets.py:70,100,263), zeros in padded slots (:91), and
    the collater masks (:274-286). raw_boxes f64 [B,L,S,4]; video_sizes i64 [B,2] = (W, H)."""
    B, L, S, _ = raw_boxes.shape
    b = raw_boxes.to(torch.float64).trunc().to(torch.int64).clamp_min(0)
    x1, y1, x2, y2 = b.unbind(-1)
    x1, x2 = torch.minimum(x1, x2), torch.maximum(x1, x2)
    y1, y2 = torch.minimum(y1, y2), torch.maximum(y1, y2)
    W = video_sizes[:, 0].view(B, 1, 1)
    H = video_sizes[:, 1].view(B, 1, 1)
    x1 = torch.where(x1 >= W, W - 1, x1)
    y1 = torch.where(y1 >= H, H - 1, y1)
    x2 = torch.where(x2 >= W, W - 1, x2)
    y2 = torch.where(y2 >= H, H - 1, y2)
    x2 = torch.where((x1 == x2) & (x1 == 0), torch.ones_like(x2), x2)
    y2 = torch.where((y1 == y2) & (y1 == 0), torch.ones_like(y2), y2)
    x1 = torch.where(x1 == x2, x1 - 1, x1)
    y1 = torch.where(y1 == y2, y1 - 1, y1)
    fixed = torch.stack([x1, y1, x2, y2], dim=-1)
    size = torch.stack([W, H, W, H], dim=-1).expand(B, L, S, 4)
    boxes = fixed / size  # int64 / int64 -> fp32 true divide, as in the reference
    is_obj = (categories != 0).unsqueeze(-1)
    boxes = torch.where(is_obj, boxes, torch.zeros_like(boxes))
    boxes[:, :, 0, :] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    return {"boxes": boxes, "src_key_padding_mask_boxes": categories == 0,
            "src_key_padding_mask_frames": frame_types == 0}


# ---------------------------------------------------------------------------------------------------
# model side (floating point): src/modelling/models.py:16-195
# ---------------------------------------------------------------------------------------------------
def _encoder_layer(x: torch.Tensor, masked: torch.Tensor, sd: Dict[str, torch.Tensor], p: str,
                   drop: Dict[str, torch.Tensor] = None) -> torch.Tensor:
    """One post-norm nn.TransformerEncoderLayer (models.py:46-52,118-124 configure it; the
    arithmetic is torch's): x [N, T, H]; masked [N, T, T] bool, True = not attended. Eval mode unless
    ``drop`` gives the dropout multipliers (0 or 1/(1-p)) of the layer's four dropout sites:
    "attn" [N, heads, T, T] on the softmax output (nn.MultiheadAttention), "branch1" [N, T, H]
    (dropout1), "ffn" [N, T, 4H] (between activation and linear2), "branch2" [N, T, H] (dropout2)."""
    drop = drop or {}
    N, T, H = x.shape
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(H, dim=-1)  # packed rows: Q; K; V
    def heads(t):
        return t.view(N, T, HEADS, HEAD_DIM).transpose(1, 2)  # [N, heads, T, d]
    q, k, v = heads(q), heads(k), heads(v)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(HEAD_DIM)
    scores = scores.masked_fill(masked.unsqueeze(1), float("-inf"))
    probs = torch.softmax(scores, dim=-1)
    if "attn" in drop:
        probs = probs * drop["attn"]
    ctx = torch.matmul(probs, v)
    ctx = ctx.transpose(1, 2).reshape(N, T, H)
    a = F.linear(ctx, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    if "branch1" in drop:
        a = a * drop["branch1"]
    x = F.layer_norm(x + a, (H,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], ENCODER_LN_EPS)
    hidden = F.gelu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
    if "ffn" in drop:
        hidden = hidden * drop["ffn"]
    f = F.linear(hidden, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    if "branch2" in drop:
        f = f * drop["branch2"]
    return F.layer_norm(x + f, (H,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], ENCODER_LN_EPS)


def stlt_forward(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], num_spatial_layers: int = 4,
                 num_temporal_layers: int = 8, layer_norm_eps: float = 1e-12,
                 dtype: torch.dtype = torch.float32, return_taps: bool = False, dropout: Dict = None):
    """Stlt.forward (models.py:185-195) and callees. Eval mode (dropout = identity) unless
    ``dropout`` gives explicit multipliers: {"embed": [B,L,S,H], "frames": [B,L,H],
    ("spatial", i): {...}, ("temporal", i): {...}} with the per-layer dicts of ``_encoder_layer``.

    ``sd`` is a reference-format state_dict. Returns logits [B, C] (or a dict of stage outputs).
    """
    dropout = dropout or {}
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    cats = batch["categories"]
    boxes = batch["boxes"].to(dtype)
    ftypes = batch["frame_types"]
    lengths = batch["lengths"]
    B, L, S = cats.shape
    H = HIDDEN
    bfe = "backbone.frames_embeddings."
    cbe = bfe + "layout_embedding.category_box_embeddings."

    # CategoryBoxEmbeddings.forward (models.py:29-39)
    e = sd[cbe + "category_embeddings.weight"][cats] + F.linear(
        boxes, sd[cbe + "box_embedding.weight"], sd[cbe + "box_embedding.bias"])
    if "scores" in batch:
        e = e + F.linear(batch["scores"].to(dtype).unsqueeze(-1), sd[cbe + "score_embeddings.weight"],
                         sd[cbe + "score_embeddings.bias"])
    e = F.layer_norm(e, (H,), sd[cbe + "layer_norm.weight"], sd[cbe + "layer_norm.bias"], layer_norm_eps)
    if "embed" in dropout:  # models.py:27,38
        e = e * dropout["embed"]
    taps = {"embed": e}

    # SpatialTransformer.forward (models.py:57-81): B*L sequences of S tokens, key padding mask
    x = e.reshape(B * L, S, H)
    key_pad = (cats == 0).reshape(B * L, 1, S).expand(B * L, S, S)  # datasets.py:277
    for i in range(num_spatial_layers):
        x = _encoder_layer(x, key_pad, sd, f"{bfe}layout_embedding.transformer.layers.{i}.",
                           dropout.get(("spatial", i)))
    taps["spatial"] = x.reshape(B, L, S, H)
    layout = x.reshape(B, L, S, H)[:, :, 0, :]  # models.py:79

    # FramesEmbeddings.forward (models.py:98-111)
    pos = sd[bfe + "position_embeddings.weight"][sd[bfe + "position_ids"][:, :L]]
    f = layout + pos + sd[bfe + "frame_type_embedding.weight"][ftypes]
    f = F.layer_norm(f, (H,), sd[bfe + "layer_norm.weight"], sd[bfe + "layer_norm.bias"], layer_norm_eps)
    if "frames" in dropout:  # models.py:93,110
        f = f * dropout["frames"]
    taps["frames"] = f

    # StltBackbone.forward (models.py:136-152): causal (model_utils.py:4-7) OR frame padding
    causal = torch.triu(torch.ones(L, L, dtype=torch.bool), diagonal=1)  # True where j > i
    masked = causal.unsqueeze(0) | (ftypes == 0).unsqueeze(1)  # [B, L, L]
    z = f
    for i in range(num_temporal_layers):
        z = _encoder_layer(z, masked, sd, f"backbone.transformer.layers.{i}.", dropout.get(("temporal", i)))
    taps["temporal"] = z

    # Stlt.forward gather (models.py:189-192) + ClassificationHead (models.py:155-163)
    h = z[torch.arange(B), lengths - 1, :]
    taps["pooled"] = h
    ph = "prediction_head."
    h = F.gelu(F.linear(h, sd[ph + "fc1.weight"], sd[ph + "fc1.bias"]))
    h = F.layer_norm(h, (H,), sd[ph + "layer_norm.weight"], sd[ph + "layer_norm.bias"], layer_norm_eps)
    logits = F.linear(h, sd[ph + "fc2.weight"], sd[ph + "fc2.bias"])
    taps["stlt"] = logits
    return taps if return_taps else logits


# ---------------------------------------------------------------------------------------------------
# training step: src/train.py:102-135, src/utils/train_inference_utils.py
# ---------------------------------------------------------------------------------------------------
def criterion(logits: torch.Tensor, labels: torch.Tensor, loss: str) -> torch.Tensor:
    """Criterion.forward for the single "stlt" logit (train_inference_utils.py:64-76): mean
    cross-entropy (Something-Else) or mean BCE-with-logits (Action Genome)."""
    if loss == "cross_entropy":
        lse = torch.logsumexp(logits, dim=-1)
        return (lse - logits.gather(1, labels.view(-1, 1)).squeeze(1)).mean()
    x, y = logits, labels.to(logits.dtype)
    return (x.clamp_min(0) - x * y + torch.log1p(torch.exp(-x.abs()))).mean()


def loss_and_grads(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], labels: torch.Tensor,
                   loss: str = "cross_entropy", dtype: torch.dtype = torch.float32, dropout: Dict = None):
    """loss.backward() of the reference loop (train.py:125-128) with dropout p = 0: returns
    (loss, logits, {name: grad}) by differentiating the restated forward. Tensors that do not reach
    the logits (orphan encoder_layer.*, score embedding without scores) have no entry, like
    ``param.grad is None`` in the reference."""
    leaves = {k: (v.detach().clone().to(dtype).requires_grad_(True) if v.is_floating_point() else v)
              for k, v in sd.items()}
    logits = stlt_forward(leaves, batch, dtype=dtype, dropout=dropout)
    value = criterion(logits, labels, loss)
    names = [k for k, v in leaves.items() if v.is_floating_point()]
    grads = torch.autograd.grad(value, [leaves[k] for k in names], allow_unused=True)
    return value.detach(), logits.detach(), {k: g for k, g in zip(names, grads) if g is not None}


def is_no_decay(name: str, tensor: torch.Tensor) -> bool:
    """add_weight_decay (train_inference_utils.py:37-54): 1-D tensors and *.bias are not decayed."""
    return tensor.dim() == 1 or name.endswith(".bias")


def linear_schedule(step: int, num_warmup_steps: int, num_training_steps: int) -> float:
    """lr_lambda of get_linear_schedule_with_warmup (train_inference_utils.py:21-34)."""
    if step < num_warmup_steps:
        return float(step) / float(max(1, num_warmup_steps))
    return max(0.0, float(num_training_steps - step) / float(max(1, num_training_steps - num_warmup_steps)))


def adamw_update(sd: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: Dict[str, dict], step: int,
                 lr: float, weight_decay: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 clip_val: float = 5.0) -> float:
    """clip_grad_norm_(parameters, clip_val) (train.py:129) followed by one torch.optim.AdamW step
    (train.py:102-104,130), in place on ``sd``; tensors without a gradient are skipped. ``step`` is
    1-based. Returns the total gradient norm before clipping."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    coef = min(1.0, clip_val / (total + 1e-6))
    b1, b2 = betas
    for name, g in grads.items():
        g = g * coef
        st = state.setdefault(name, {"m": torch.zeros_like(g), "v": torch.zeros_like(g)})
        p = sd[name]
        wd = 0.0 if is_no_decay(name, p) else weight_decay
        p.mul_(1.0 - lr * wd)
        st["m"].mul_(b1).add_(g, alpha=1.0 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = st["v"].sqrt() / math.sqrt(1.0 - b2 ** step) + eps
        p.addcdiv_(st["m"], denom, value=-(lr / (1.0 - b1 ** step)))
    return total


# ---------------------------------------------------------------------------------------------------
# CACNF on precomputed appearance features: src/modelling/models.py:232-549
# ---------------------------------------------------------------------------------------------------
def _mha(xq: torch.Tensor, xkv: torch.Tensor, sd: Dict[str, torch.Tensor], p: str,
         masked: torch.Tensor = None) -> torch.Tensor:
    """nn.MultiheadAttention (eval) with separate query / key-value streams: xq [N, Tq, H], xkv
    [N, Tk, H]; packed in-projection rows Q; K; V. masked [N, Tq, Tk] bool (True = not attended)."""
    N, Tq, H = xq.shape
    Tk = xkv.shape[1]
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(xq, w[:H], b[:H]).view(N, Tq, HEADS, HEAD_DIM).transpose(1, 2)
    k = F.linear(xkv, w[H:2 * H], b[H:2 * H]).view(N, Tk, HEADS, HEAD_DIM).transpose(1, 2)
    v = F.linear(xkv, w[2 * H:], b[2 * H:]).view(N, Tk, HEADS, HEAD_DIM).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(HEAD_DIM)
    if masked is not None:
        scores = scores.masked_fill(masked.unsqueeze(1), float("-inf"))
    ctx = torch.matmul(torch.softmax(scores, dim=-1), v).transpose(1, 2).reshape(N, Tq, H)
    return F.linear(ctx, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _attn_layer(x, context, sd, p, eps, masked=None):
    """CrossAttentionLayer / SelfAttentionLayer (models.py:328-373), eval: LN(attn(x, ctx) + x)."""
    return F.layer_norm(_mha(x, context, sd, p + "attn.", masked) + x, (x.shape[-1],), sd[p + "ln.weight"],
                        sd[p + "ln.bias"], eps)


def _head(x, sd, p, eps):
    """ClassificationHead / FusionHead (models.py:155-163, 286-294)."""
    h = F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    h = F.layer_norm(h, (h.shape[-1],), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], eps)
    return F.linear(h, sd[p + "fc2.weight"], sd[p + "fc2.bias"])


def cacnf_forward(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], features: torch.Tensor,
                  num_spatial_layers: int = 4, num_temporal_layers: int = 8, num_appearance_layers: int = 4,
                  num_fusion_layers: int = 4, layer_norm_eps: float = 1e-12) -> Dict[str, torch.Tensor]:
    """CrossAttentionCentralNetFusion.forward (models.py:526-549) in eval mode with the 3D-ResNet trunk
    factored out: ``features`` = Resnet3D.forward_features(batch) [B, 2048, T', H', W'] (models.py:219-220).
    ``sd`` is the reference CACNF state_dict (entries of the ResNet trunk are not needed)."""
    H = HIDDEN
    B, L, _ = batch["categories"].shape
    lengths, ftypes = batch["lengths"], batch["frame_types"]
    # layout branch = StltBackbone under another prefix (models.py:440, 452-453)
    lb, lc = "backbone.layout_branch.", "layout_classifier."
    stlt_sd = {"backbone." + k[len(lb):]: v for k, v in sd.items() if k.startswith(lb)}
    stlt_sd.update({"prediction_head." + k[len(lc):]: v for k, v in sd.items() if k.startswith(lc)})
    taps = stlt_forward(stlt_sd, batch, num_spatial_layers, num_temporal_layers, layer_norm_eps, return_taps=True)
    layout = taps["temporal"]                                   # [B, L, H]
    logits_stlt = taps["stlt"]                                  # layout_classifier(layout_hidden_state)

    # appearance branch: TransformerResnet.forward_features (models.py:256-276)
    ab = "backbone.appearance_branch."
    feats = features.flatten(2).transpose(1, 2)                 # [B, P, 2048]
    proj = F.linear(feats, sd[ab + "projector.weight"].flatten(1), sd[ab + "projector.bias"])  # 1x1x1 Conv3d
    app = torch.cat((sd[ab + "cls_token"].expand(B, 1, H), proj), dim=1) + sd[ab + "pos_embed"].transpose(0, 1)
    T = app.shape[1]
    for i in range(num_appearance_layers):                      # post-norm, ReLU, eps 1e-5 (nn defaults)
        p = f"{ab}transformer.layers.{i}."
        a = _mha(app, app, sd, p + "self_attn.")
        app = F.layer_norm(app + a, (H,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], ENCODER_LN_EPS)
        f = F.linear(F.relu(F.linear(app, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                     sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        app = F.layer_norm(app + f, (H,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], ENCODER_LN_EPS)
    logits_app = _head(app[:, 0], sd, "appearance_classifier.", layer_norm_eps)

    # fusion: CrossAttentionFusionBackbone.forward (models.py:464-470), CrossModalModule (:395-431)
    frame_pad = ftypes == 0                                      # src_key_padding_mask_frames
    causal = torch.triu(torch.ones(L, L, dtype=torch.bool), diagonal=1)
    self_mask = causal.unsqueeze(0) | frame_pad.unsqueeze(1)     # [B, L, L]
    cross_mask = frame_pad.unsqueeze(1).expand(B, T, L)          # appearance queries x layout keys
    for i in range(num_fusion_layers):
        p = f"backbone.mm_fusion.{i}."
        l1 = _attn_layer(layout, app, sd, p + "cross_attn.", layer_norm_eps)
        a1 = _attn_layer(app, layout, sd, p + "cross_attn.", layer_norm_eps, cross_mask)
        l2 = _attn_layer(l1, l1, sd, p + "layout_attn.", layer_norm_eps, self_mask)
        a2 = _attn_layer(a1, a1, sd, p + "appearance_attn.", layer_norm_eps)
        ff = F.linear(F.gelu(F.linear(l2, sd[p + "layout_ffn.linear1.weight"], sd[p + "layout_ffn.linear1.bias"])),
                      sd[p + "layout_ffn.linear2.weight"], sd[p + "layout_ffn.linear2.bias"])
        layout = F.layer_norm(ff + l2, (H,), sd[p + "layout_ffn.ln.weight"], sd[p + "layout_ffn.ln.bias"], layer_norm_eps)
        app = _attn_layer(a2, a2, sd, p + "appearance_ffn.", layer_norm_eps)
    fused = torch.cat((layout[torch.arange(B), lengths - 1], app[:, 0]), dim=-1)
    logits_caf = _head(fused, sd, "fusion_classifier.", layer_norm_eps)
    logits = (logits_stlt, logits_app, logits_caf)
    return {"stlt": logits_stlt, "resnet3d": logits_app, "caf": logits_caf, "ensemble": sum(logits) / 3}


# ---------------------------------------------------------------------------------------------------
# multi-label evaluation: src/utils/evaluation.py:100-132
# ---------------------------------------------------------------------------------------------------
def charades_map(predictions, ground_truths):
    """charades_map + map (evaluation.py:100-132) restated with numpy: rows without any positive label are
    ranked last (-inf), per class AP = mean over positives of (positives so far / rank), classes without
    positives give nan (and np.mean then makes the mAP nan). Returns (mAP, per-class AP)."""
    import numpy as np
    pred = np.array(predictions, dtype=np.float64, copy=True)
    gt = np.asarray(ground_truths, dtype=np.float64)
    pred[gt.sum(axis=1) == 0, :] = -np.inf
    aps = []
    for c in range(pred.shape[1]):
        order = np.argsort(-pred[:, c])
        tp = gt[order, c] == 1
        n_pos = int(tp.sum())
        if n_pos < 1:
            aps.append(float("nan"))
            continue
        ranks = np.arange(1, len(tp) + 1, dtype=np.float64)
        aps.append(float((np.cumsum(tp)[tp] / ranks[tp]).sum() / n_pos))
    aps = np.array(aps)
    return float(np.mean(aps)), aps


def caf_forward(sd, batch, features, **kw):
    """CrossAttentionFusion.forward (models.py:486-501): the CACNF backbone + the fusion classifier."""
    remap = {}
    for k, v in sd.items():
        if k.startswith("caf_backbone."):
            remap["backbone." + k[len("caf_backbone."):]] = v
        elif k.startswith("classifier."):
            remap["fusion_classifier." + k[len("classifier."):]] = v
    _add_zero_heads(remap)
    return {"caf": cacnf_forward(remap, batch, features, **kw)["caf"]}


def lcf_forward(sd, batch, features, **kw):
    """LateConcatenationFusion.forward (models.py:305-323): layout state [lengths - 1] ++ appearance CLS state ->
    FusionHead, i.e. the CACNF data flow without fusion layers."""
    remap = {}
    for k, v in sd.items():
        if k.startswith("layout_branch.") or k.startswith("appearance_branch."):
            remap["backbone." + k] = v
        elif k.startswith("classifier."):
            remap["fusion_classifier." + k[len("classifier."):]] = v
    _add_zero_heads(remap)
    return {"lcf": cacnf_forward(remap, batch, features, num_fusion_layers=0, **kw)["caf"]}


def _add_zero_heads(sd):
    c = sd["fusion_classifier.fc2.weight"].shape[0]
    for p in ("layout_classifier.", "appearance_classifier."):
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = torch.zeros(HIDDEN, HIDDEN), torch.zeros(HIDDEN)
        sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"] = torch.ones(HIDDEN), torch.zeros(HIDDEN)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = torch.zeros(c, HIDDEN), torch.zeros(c)
