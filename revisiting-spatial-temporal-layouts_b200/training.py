"""Fused STLT training step (BASELINE.json configs[3]).

The drop-in module already trains under the reference's own loop (``model(batch)`` returns logits
with a grad_fn, see module._StltTrainFunction). This module is the B200-first version of that loop
body (reference src/train.py:117-135) with no host synchronisation and no per-tensor launches:

    zero grads -> forward (activations kept) -> criterion -> backward
               -> gradient all-reduce over NCCL (data parallel; buckets leave on a communication stream as the
                  backward stages that produce them finish) -> global-norm clip + AdamW on flat fp32 buffers -> bf16 re-pack

Semantics follow the reference: ``Criterion`` (src/utils/train_inference_utils.py:64-76),
``add_weight_decay`` (:37-54; 1-D tensors and ``*.bias`` are not decayed),
``get_linear_schedule_with_warmup`` (:21-34), ``clip_grad_norm_(model.parameters(), clip_val)``
(src/train.py:129), ``optim.AdamW(lr)`` defaults (betas 0.9/0.999, eps 1e-8). Parameters that never
receive a gradient (the orphan prototype layer, models.py:46-52; the score embedding when the batch
has no ``scores``) are left untouched, exactly as AdamW skips ``grad is None``.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Callable, Dict, List, Optional

import torch

from . import lib as _lib
from .module import Stlt


def linear_schedule_with_warmup(num_warmup_steps: int, num_training_steps: int) -> Callable[[int], float]:
    """lr multiplier of get_linear_schedule_with_warmup (src/utils/train_inference_utils.py:21-34)."""

    def lr_lambda(current_step: int) -> float:
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        return max(0.0, float(num_training_steps - current_step)
                   / float(max(1, num_training_steps - num_warmup_steps)))

    return lr_lambda


def _is_no_decay(name: str, p: torch.Tensor) -> bool:  # add_weight_decay, :46
    return p.dim() == 1 or name.endswith(".bias")


_LAYER = re.compile(r"transformer\.layers\.(\d+)\.")


def backward_stage_of(name: str, num_spatial_layers: int, num_temporal_layers: int) -> int:
    """The stage of stlt_backward (include/stlt_b200.h, stlt_backward_stage_events) after which the gradient of
    parameter ``name`` is final: 0 head | 1..nt temporal layers nt-1..0 | nt+1 frame embedding |
    nt+2..nt+1+ns spatial layers ns-1..0 | nt+ns+2 category / box / score embedding."""
    ns, nt = num_spatial_layers, num_temporal_layers
    m = _LAYER.search(name)
    if name.startswith("prediction_head."):
        return 0
    if ".layout_embedding." in name:
        return nt + 2 + (ns - 1 - int(m.group(1))) if m else nt + ns + 2
    if m:
        return 1 + (nt - 1 - int(m.group(1)))
    return nt + 1


SEGMENT_ORDER = ("d", "nd", "sc_nd", "sc_d")
BUCKET_SCHEMES = ("one", "two", "per_layer")


def plan_flat_layout(named_parameters, num_spatial_layers: Optional[int] = None,
                     num_temporal_layers: Optional[int] = None):
    """Flat-buffer layout of the trainable parameters: [(name, param, offset)], {segment: (start, end)},
    total, stage_ends.

    Segments: ``d`` the weight-decayed tensors (add_weight_decay, train_inference_utils.py:37-54) ordered by the
    backward stage that finishes their gradient, so that every all-reduce bucket is one contiguous slice that can
    leave while the backward pass is still running; ``nd`` the 1-D tensors and biases (0.5 MB in all, they travel
    with the last bucket); ``sc_*`` the score embedding, which only gets a gradient when the batch carries
    ``scores``. ``stage_ends[k]`` = end offset (inside ``d``) of the tensors of stages <= k. The layer counts fix the
    stage numbering (it must be stlt_backward's); without them they are read off the parameter names, which is only
    right when every layer has a trainable tensor."""
    named = [(n, p) for n, p in named_parameters if p.requires_grad and ".encoder_layer." not in n]
    layers = {"s": 0, "t": 0}
    for n, _ in named:
        m = _LAYER.search(n)
        if m:
            key = "s" if ".layout_embedding." in n else "t"
            layers[key] = max(layers[key], int(m.group(1)) + 1)
    ns = layers["s"] if num_spatial_layers is None else int(num_spatial_layers)
    nt = layers["t"] if num_temporal_layers is None else int(num_temporal_layers)
    num_stages = ns + nt + 3
    segs: Dict[str, List] = {k: [] for k in SEGMENT_ORDER}
    for name, p in named:
        nd = _is_no_decay(name, p)
        if "score_embeddings" in name:
            segs["sc_nd" if nd else "sc_d"].append((num_stages - 1, name, p))
        else:
            segs["nd" if nd else "d"].append((backward_stage_of(name, ns, nt), name, p))
    segs["d"].sort(key=lambda t: t[0])  # stable: declaration order inside a stage
    segments, layout, off = {}, [], 0
    stage_ends = [0] * num_stages
    for key in SEGMENT_ORDER:
        start = off
        for stage, name, p in segs[key]:
            layout.append((name, p, off))
            off += (p.numel() + 3) // 4 * 4  # keep every tensor 16-byte aligned
            if key == "d":
                stage_ends[stage] = off
        segments[key] = (start, off)
    for k in range(1, num_stages):
        stage_ends[k] = max(stage_ends[k], stage_ends[k - 1])
    return layout, segments, off, stage_ends


def plan_buckets(scheme: str, stage_ends: List[int], total: int, num_temporal_stages: int,
                 min_bucket_elems: int = 2 << 20):
    """All-reduce buckets [(stage, start, end)] of the flat gradient buffer, in backward-completion order; bucket i may
    leave once backward stage ``stage`` has finished. ``one``: the whole buffer after the backward pass. ``two``: the
    temporal-phase gradients (stages < num_temporal_stages), then the rest. ``per_layer``: one bucket per stage, stages
    smaller than ``min_bucket_elems`` merged into the next one. The last bucket always runs to ``total`` (it carries
    the no-decay tensors and the score embedding)."""
    last = len(stage_ends) - 1
    if scheme == "one":
        cuts = []
    elif scheme == "two":
        cuts = [num_temporal_stages - 1]
    elif scheme == "per_layer":
        cuts, start = [], 0
        for k in range(last):
            if stage_ends[k] - start >= min_bucket_elems:
                cuts.append(k)
                start = stage_ends[k]
    else:
        raise ValueError(f"bucket scheme must be one of {BUCKET_SCHEMES}")
    buckets, start = [], 0
    for k in cuts:
        if stage_ends[k] > start:
            buckets.append((k, start, stage_ends[k]))
            start = stage_ends[k]
    buckets.append((last, start, total))
    return buckets


def all_reduce_buckets(flat_grads: torch.Tensor, buckets, group=None, before_bucket=None, works=None):
    """Issues the SUM all-reduce of every bucket asynchronously, in order; ``before_bucket(stage)`` runs first
    (FusedTrainStep makes the communication stream wait for the backward stage's event there). Returns the list of
    work handles (appended to ``works`` when given); the caller waits on them before it reads the gradients."""
    import torch.distributed as dist
    works = [] if works is None else works
    for stage, a, b in buckets:
        if before_bucket is not None:
            before_bucket(stage)
        works.append(dist.all_reduce(flat_grads[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True))
    return works


class FusedTrainStep:
    """Owns flat fp32 parameter / gradient / AdamW-state buffers of an ``Stlt`` module.

    The module's parameters are re-pointed at views of one flat buffer (state_dict keys, shapes and
    values are unchanged), laid out as [weight-decayed tensors in backward-completion order | 1-D tensors
    and biases | score-embedding bias | score-embedding weight] (plan_flat_layout), so the optimizer is two
    to four launches and each all-reduce bucket is one contiguous slice.

    Data parallel (``bucket_scheme``): stlt_backward records an event per finished stage (head, each encoder
    layer, the embeddings); a communication stream waits for the event of a bucket's last stage and the NCCL
    all-reduce of that bucket runs underneath the rest of the backward pass. ``two`` (default): the temporal-phase
    gradients (230 MB) travel under the spatial stack's backward pass, the spatial-phase gradients (114 MB, 0.3 ms on
    8 B200s over NVSwitch) at the end. ``per_layer``: 13 buckets of one encoder layer each, only the last 0.5 MB
    exposed - measured SLOWER on 8 GPUs (53.15 vs 52.56 ms per step, 51.4 without any all-reduce;
    profiles/r2_v9_train_comm_probe.md): the whole 344 MB all-reduce takes 0.9 ms alone, so there is little to hide,
    while every collective that runs beside the persistent 148-CTA GEMM grids takes SMs away from them. ``one``: a
    single all-reduce after the backward pass (52.79 ms).
    """

    BUCKET_SCHEMES = BUCKET_SCHEMES

    def __init__(self, model: Stlt, lr: float = 5e-5, weight_decay: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, clip_val: Optional[float] = 5.0, loss: str = "cross_entropy",
                 lr_lambda: Optional[Callable[[int], float]] = None, process_group=None,
                 dropout_p: Optional[float] = None, seed: int = 0):
        if loss not in ("cross_entropy", "bce_with_logits"):
            raise ValueError("loss must be 'cross_entropy' (Something-Else) or 'bce_with_logits' (Action Genome)")
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.clip_val, self.loss_kind = clip_val, loss
        self.lr_lambda = lr_lambda or (lambda step: 1.0)
        self.group = process_group
        self.dropout_p = float(model.config.hidden_dropout_prob) if dropout_p is None else float(dropout_p)
        self.seed = seed
        self.step_count = 0
        self.score_step_count = 0  # AdamW bias correction of the score embedding counts only the steps that updated it
        self._checked_batch = None
        self._ws = None
        self._ws_key = None
        self.bucket_scheme = os.environ.get("STLT_TRAIN_BUCKETS", "two")  # plan_buckets
        self._comm_stream = None

        device = next(model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("FusedTrainStep needs the module on a CUDA device (there is no CPU path)")
        self.device = device
        layout, self.segments, off, self.stage_ends = plan_flat_layout(
            model.named_parameters(), model.config.num_spatial_layers, model.config.num_temporal_layers)
        self.total = off
        self.num_temporal_stages = int(model.config.num_temporal_layers) + 2  # head, temporal layers, frame embedding
        self.flat_params = torch.zeros(off, dtype=torch.float32, device=device)
        self.flat_grads = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad_views: Dict[str, torch.Tensor] = {}
        with torch.no_grad():
            for name, p, o in layout:
                view = self.flat_params[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self.grad_views[name] = self.flat_grads[o:o + p.numel()].view(p.shape)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        self._sumsq_scratch = torch.zeros(2048, dtype=torch.float32, device=device)
        self._loss = torch.zeros(1, dtype=torch.float32, device=device)
        self._bound_scores = None

    # ------------------------------------------------------------------------------------------
    def _world(self) -> int:
        import torch.distributed as dist
        if self.group is None and not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def buckets(self, scheme: Optional[str] = None):
        """[(stage, start, end)] of the current (or given) bucket scheme."""
        return plan_buckets(scheme or self.bucket_scheme, self.stage_ends, self.total, self.num_temporal_stages)

    @property
    def num_buckets(self) -> int:
        return len(self.buckets())

    def bucket_bounds(self) -> Dict[str, tuple]:
        """{label: (start, end)} of the per-layer buckets (tools/train_comm_probe.py)."""
        return {f"stage<={k}": (a, b) for k, a, b in self.buckets("per_layer")}

    def _bind(self, has_scores: bool) -> None:
        if self._bound_scores == has_scores:
            return
        grads = {n: g for n, g in self.grad_views.items() if has_scores or "score_embeddings" not in n}
        self.model._bind_grads(grads)
        self._bound_scores = has_scores

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """One optimisation step on ``batch`` (the reference batch dict + ``labels``). Returns the
        mean loss of the local batch as a device scalar (no synchronisation)."""
        import torch.distributed as dist
        model, lib = self.model, _lib.load_library()
        cats = batch["categories"]
        B, L, S = cats.shape
        device = self.device
        inputs = (model._as_input(batch, "categories", torch.int64, (B, L, S), device),
                  model._as_input(batch, "boxes", torch.float32, (B, L, S, 4), device),
                  model._as_input(batch, "scores", torch.float32, (B, L, S), device) if "scores" in batch else None,
                  model._as_input(batch, "frame_types", torch.int64, (B, L), device),
                  model._as_input(batch, "lengths", torch.int64, (B,), device))
        labels = batch["labels"].to(device)
        C = model.config.num_classes
        if self.loss_kind == "cross_entropy":
            labels = labels.to(torch.int64).contiguous()
            kind = _lib.LOSS_CROSS_ENTROPY
        else:
            labels = labels.to(torch.float32).contiguous()
            kind = _lib.LOSS_BCE_LOGITS
        world = self._world()
        has_scores = inputs[2] is not None

        with torch.cuda.device(device):
            model._ensure_handle(device)
            stream = torch.cuda.current_stream(device).cuda_stream
            model._sync_weights(device, stream, "bf16")
            self._bind(has_scores)
            if self._ws_key != (B, L, S):
                self._ws = model._train_workspace(B, L, S, device)
                self._ws_key = (B, L, S)
            if world > 1 and self._checked_batch != B:
                # the 1/world scaling below is the global mean only when every rank holds the same number of videos
                sizes = torch.tensor([B, -B], device=device)
                dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=self.group)
                if int(sizes[0]) != B or int(-sizes[1]) != B:
                    raise RuntimeError("FusedTrainStep: every data-parallel rank must hold the same local batch size")
                self._checked_batch = B
            self.flat_grads.zero_()  # optimizer.zero_grad()
            self.step_count += 1
            if has_scores:
                self.score_step_count += 1
            # independent dropout masks per rank and step (a DDP run draws from per-rank generators)
            rank = dist.get_rank(self.group) if world > 1 else 0
            drop = (self.dropout_p if model.training else 0.0, self.seed + self.step_count * world + rank)
            logits = model._forward_train(inputs, self._ws, *drop)
            d_logits = torch.empty_like(logits)
            # mean over the GLOBAL batch: local mean gradient scaled by 1/world, summed by the all-reduce
            _lib.check(model._handle, lib.stlt_loss(model._handle, stream, kind, logits.data_ptr(),
                                                    labels.data_ptr(), B, C, 1.0 / world,
                                                    self._loss.data_ptr(), d_logits.data_ptr()))
            if world > 1:
                self._backward_overlapped(inputs, d_logits, drop, device)
            else:
                model._backward(inputs, self._ws, d_logits, _lib.BWD_ALL, *drop)
            sumsq_ptr = None
            if self.clip_val is not None:
                _lib.check(model._handle, lib.stlt_grad_sumsq(model._handle, stream, self.flat_grads.data_ptr(),
                                                              self.total, self._sumsq.data_ptr(),
                                                              self._sumsq_scratch.data_ptr(), self._sumsq_scratch.numel()))
                sumsq_ptr = self._sumsq.data_ptr()
            lr = self.lr * self.lr_lambda(self.step_count - 1)
            for key in SEGMENT_ORDER:
                a, b = self.segments[key]
                if b == a or (key.startswith("sc") and not has_scores):
                    continue
                wd = 0.0 if key.endswith("nd") else self.weight_decay
                es = 4  # bytes per element
                _lib.check(model._handle, lib.stlt_adamw_step(
                    model._handle, stream, self.flat_params.data_ptr() + a * es, self.flat_grads.data_ptr() + a * es,
                    self.exp_avg.data_ptr() + a * es, self.exp_avg_sq.data_ptr() + a * es, b - a, lr,
                    self.betas[0], self.betas[1], self.eps, wd,
                    self.score_step_count if key.startswith("sc") else self.step_count, sumsq_ptr,
                    float(self.clip_val or 0.0)))
            # the fp32 master weights changed in place behind PyTorch's back: re-pack the bf16 operands
            self._repack(stream)
        self._keepalive = (inputs, labels, d_logits, logits)
        self.last_logits = logits
        return self._loss[0].clone()  # a fresh device scalar: the accumulator is reused by the next step

    def _backward_overlapped(self, inputs, d_logits, drop, device) -> None:
        """Backward pass with the gradient all-reduce underneath it: each bucket leaves on the communication stream as soon
        as the event of its last backward stage has fired; the compute stream waits for all of them at the end. The
        temporal-phase buckets are issued before the spatial half of the backward pass is enqueued, so a host that is not
        far ahead of the GPU does not delay them."""
        model, lib = self.model, _lib.load_library()
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=device)
            _lib.check(model._handle, lib.stlt_backward_stage_events(model._handle, 1, None))
        comm = self._comm_stream

        def wait_stage(stage: int) -> None:
            _lib.check(model._handle, lib.stlt_stream_wait_backward_stage(model._handle, comm.cuda_stream, stage))

        buckets = self.buckets()
        first = [b for b in buckets if b[0] < self.num_temporal_stages]
        rest = [b for b in buckets if b[0] >= self.num_temporal_stages]
        model._backward(inputs, self._ws, d_logits, _lib.BWD_TEMPORAL, *drop)
        with torch.cuda.stream(comm):
            works = all_reduce_buckets(self.flat_grads, first, self.group, before_bucket=wait_stage)
        model._backward(inputs, self._ws, None, _lib.BWD_SPATIAL, *drop)
        with torch.cuda.stream(comm):
            all_reduce_buckets(self.flat_grads, rest, self.group, before_bucket=wait_stage, works=works)
        for w in works:  # the compute stream waits for the NCCL stream; no host synchronisation
            w.wait()

    def _repack(self, stream: int) -> None:
        model, lib = self.model, _lib.load_library()
        prec = _lib.PRECISION_BF16
        buf = model._packed[prec]
        _lib.check(model._handle, lib.stlt_pack_weights(model._handle, stream, prec, buf.data_ptr(), buf.numel()))
        # copies packed for another precision are stale now (the parameter versions did not change)
        model._packed_key = {prec: model._weights_key}
        model._last_packed = prec

    def grad_norm(self) -> float:
        """Total gradient norm of the last step (synchronises; debugging / tests)."""
        return float(self._sumsq.sqrt().item())
