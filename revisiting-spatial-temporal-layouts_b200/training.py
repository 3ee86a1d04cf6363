"""Fused STLT training step (BASELINE.json configs[3]).

The drop-in module already trains under the reference's own loop (``model(batch)`` returns logits
with a grad_fn, see module._StltTrainFunction). This module is the B200-first version of that loop
body (reference src/train.py:117-135) with no host synchronisation and no per-tensor launches:

    zero grads -> forward (activations kept) -> criterion -> backward
               -> gradient all-reduce over NCCL (data parallel; overlapped with the second half of
                  the backward pass) -> global-norm clip + AdamW on flat fp32 buffers -> bf16 re-pack

Semantics follow the reference: ``Criterion`` (src/utils/train_inference_utils.py:64-76),
``add_weight_decay`` (:37-54; 1-D tensors and ``*.bias`` are not decayed),
``get_linear_schedule_with_warmup`` (:21-34), ``clip_grad_norm_(model.parameters(), clip_val)``
(src/train.py:129), ``optim.AdamW(lr)`` defaults (betas 0.9/0.999, eps 1e-8). Parameters that never
receive a gradient (the orphan prototype layer, models.py:46-52; the score embedding when the batch
has no ``scores``) are left untouched, exactly as AdamW skips ``grad is None``.
"""
from __future__ import annotations

import ctypes
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


def _phase_of(name: str) -> int:
    """0: gradients produced by STLT_BWD_TEMPORAL (head, temporal stack, frame embedding);
    1: by STLT_BWD_SPATIAL (spatial stack, category/box embedding)."""
    return 1 if ".layout_embedding." in name else 0


SEGMENT_ORDER = ("t_nd", "t_d", "s_nd", "s_d", "sc_nd", "sc_d")


def plan_flat_layout(named_parameters):
    """Flat-buffer layout of the trainable parameters: [(name, param, offset)], {segment: (start, end)},
    total. Segments: {temporal, spatial backward phase} x {no-decay, decay} and the score embedding
    (which only gets a gradient when the batch carries ``scores``)."""
    segs: Dict[str, List] = {k: [] for k in SEGMENT_ORDER}
    for name, p in named_parameters:
        if not p.requires_grad or ".encoder_layer." in name:
            continue
        nd = _is_no_decay(name, p)
        if "score_embeddings" in name:
            segs["sc_nd" if nd else "sc_d"].append((name, p))
        else:
            segs[("t" if _phase_of(name) == 0 else "s") + ("_nd" if nd else "_d")].append((name, p))
    segments, layout, off = {}, [], 0
    for key in SEGMENT_ORDER:
        start = off
        for name, p in segs[key]:
            layout.append((name, p, off))
            off += (p.numel() + 3) // 4 * 4  # keep every tensor 16-byte aligned
        segments[key] = (start, off)
    return layout, segments, off


def all_reduce_buckets(flat_grads: torch.Tensor, segments, group=None, between=None):
    """Sums the flat gradient over the data-parallel group in two contiguous buckets. ``between`` (the
    second half of the backward pass) runs while the first bucket — the temporal-phase gradients, which
    are final by then — is in flight."""
    import torch.distributed as dist
    t0, t1 = segments["t_nd"][0], segments["t_d"][1]
    s0 = segments["s_nd"][0]
    work = dist.all_reduce(flat_grads[t0:t1], op=dist.ReduceOp.SUM, group=group, async_op=True)
    if between is not None:
        between()
    dist.all_reduce(flat_grads[s0:], op=dist.ReduceOp.SUM, group=group)
    work.wait()


class FusedTrainStep:
    """Owns flat fp32 parameter / gradient / AdamW-state buffers of an ``Stlt`` module.

    The module's parameters are re-pointed at views of one flat buffer (state_dict keys, shapes and
    values are unchanged), laid out as [temporal-phase no-decay | temporal-phase decay |
    spatial-phase no-decay | spatial-phase decay | score-embedding bias | score-embedding weight], so
    the optimizer is four launches and each all-reduce bucket is one contiguous slice.
    """

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
        self.num_buckets = 2  # all-reduce buckets per step (see all_reduce_buckets)

        device = next(model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("FusedTrainStep needs the module on a CUDA device (there is no CPU path)")
        self.device = device
        layout, self.segments, off = plan_flat_layout(model.named_parameters())
        self.total = off
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
            model._backward(inputs, self._ws, d_logits, _lib.BWD_TEMPORAL, *drop)
            spatial = lambda: model._backward(inputs, self._ws, None, _lib.BWD_SPATIAL, *drop)  # noqa: E731
            if world > 1:  # bucket 1 travels over NVLink while the spatial stack's backward runs
                all_reduce_buckets(self.flat_grads, self.segments, self.group, between=spatial)
            else:
                spatial()
            sumsq_ptr = None
            if self.clip_val is not None:
                _lib.check(model._handle, lib.stlt_grad_sumsq(model._handle, stream, self.flat_grads.data_ptr(),
                                                              self.total, self._sumsq.data_ptr(),
                                                              self._sumsq_scratch.data_ptr(), self._sumsq_scratch.numel()))
                sumsq_ptr = self._sumsq.data_ptr()
            lr = self.lr * self.lr_lambda(self.step_count - 1)
            for key in ("t_nd", "t_d", "s_nd", "s_d", "sc_nd", "sc_d"):
                a, b = self.segments[key]
                if b == a or (key.startswith("sc") and not has_scores):
                    continue
                wd = 0.0 if key.endswith("_nd") else self.weight_decay
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
