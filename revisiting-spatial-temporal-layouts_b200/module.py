"""Drop-in replacement for the reference ``Stlt`` nn.Module (reference
src/modelling/models.py:166-195): same constructor argument (an ``StltModelConfig``), same batch
dict, same ``state_dict`` keys / shapes / dtypes (174 entries for the Something-Else config,
SURVEY.md Appendix A.3), same ``{"stlt": logits}`` output — but ``forward`` is one call into
libstlt_b200.so (hand-written sm_100a CUDA behind the C ABI of include/stlt_b200.h).

The sub-modules below are *parameter holders*: they exist so that parameter names, shapes and
default initialisation match the reference; none of them has a PyTorch compute path. There is no
CPU fallback — a non-CUDA batch raises.
"""
from __future__ import annotations

import copy
import ctypes
import os
import math
from typing import Dict, Optional

import torch
from torch import nn

from . import lib as _lib


# ------------------------------------------------------------------------------------------------
# parameter holders (names follow the reference module tree)
# ------------------------------------------------------------------------------------------------
class _Affine(nn.Module):
    """weight [out, in] + bias [out] with nn.Linear's default initialisation."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_features)
        nn.init.uniform_(self.bias, -bound, bound)


class _Norm(nn.Module):
    def __init__(self, size: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(size))
        self.bias = nn.Parameter(torch.zeros(size))


class _Table(nn.Module):
    """Embedding table. ``padding_idx`` zeroes that row at initialisation and, as in nn.Embedding, the row never
    receives a gradient (the backward kernels skip row 0 of the category and frame-type tables); a loaded
    checkpoint may hold anything there and the forward does a plain gather."""

    def __init__(self, rows: int, size: int, padding_idx: Optional[int] = None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(rows, size))
        nn.init.normal_(self.weight)
        if padding_idx is not None:
            with torch.no_grad():
                self.weight[padding_idx].fill_(0)


class _SelfAttention(nn.Module):
    """Packed in-projection [3H, H] (rows Q; K; V) + out_proj, initialised like nn.MultiheadAttention."""

    def __init__(self, hidden: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * hidden, hidden))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * hidden))
        self.out_proj = _Affine(hidden, hidden)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class _EncoderLayer(nn.Module):
    """Parameters of one post-norm nn.TransformerEncoderLayer (reference models.py:46-52,118-124)."""

    def __init__(self, hidden: int):
        super().__init__()
        self.self_attn = _SelfAttention(hidden)
        self.linear1 = _Affine(hidden, 4 * hidden)
        self.linear2 = _Affine(4 * hidden, hidden)
        self.norm1 = _Norm(hidden)
        self.norm2 = _Norm(hidden)


class _Encoder(nn.Module):
    """nn.TransformerEncoder deep-copies ONE layer num_layers times (identical clones at init)."""

    def __init__(self, layer: _EncoderLayer, num_layers: int):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])


class _CategoryBoxEmbeddings(nn.Module):  # reference models.py:16-27
    def __init__(self, config):
        super().__init__()
        h = config.hidden_size
        self.category_embeddings = _Table(config.unique_categories, h, padding_idx=0)
        self.box_embedding = _Affine(4, h)
        self.score_embeddings = _Affine(1, h)
        self.layer_norm = _Norm(h)


class _SpatialTransformer(nn.Module):  # reference models.py:42-55
    def __init__(self, config):
        super().__init__()
        self.category_box_embeddings = _CategoryBoxEmbeddings(config)
        # The reference registers the prototype layer as a sub-module too; it is never used in
        # forward but its 12 tensors are part of every checkpoint.
        self.encoder_layer = _EncoderLayer(config.hidden_size)
        self.transformer = _Encoder(self.encoder_layer, config.num_spatial_layers)


class _FramesEmbeddings(nn.Module):  # reference models.py:84-96
    def __init__(self, config):
        super().__init__()
        h = config.hidden_size
        self.layout_embedding = _SpatialTransformer(config)
        self.position_embeddings = _Table(config.layout_num_frames, h)
        self.frame_type_embedding = _Table(5, h, padding_idx=0)
        self.layer_norm = _Norm(h)
        self.register_buffer("position_ids", torch.arange(config.layout_num_frames).expand((1, -1)))


class StltBackbone(nn.Module):
    """Parameter tree of the reference StltBackbone (models.py:114-152)."""

    def __init__(self, config):
        super().__init__()
        self.frames_embeddings = _FramesEmbeddings(config)
        self.transformer = _Encoder(_EncoderLayer(config.hidden_size), config.num_temporal_layers)
        object.__setattr__(self, "_config", config)   # not part of the state_dict
        object.__setattr__(self, "_runner", None)

    @classmethod
    def from_pretrained(cls, config):  # models.py:130-134
        model = cls(config)
        model.load_state_dict(torch.load(config.load_backbone_path, map_location="cpu"))
        return model

    def forward(self, batch, precision: str = "fp32"):
        """Reference StltBackbone.forward (models.py:136-152): the temporal stack's output for every frame,
        seq-first [L, B, H] as the reference returns it (its fusion models index it with lengths - 1).
        Inference only; runs the same library path as ``Stlt`` (this module's parameters + an unused head)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("StltBackbone.forward is inference-only here: use Stlt for training, or call "
                                      "it under torch.no_grad() in eval mode")
        runner = self._runner
        device = batch["categories"].device
        if runner is None or runner.precision != precision or next(runner.prediction_head.parameters()).device != device:
            runner = Stlt.__new__(Stlt)
            nn.Module.__init__(runner)
            runner.config, runner.precision = self._config, precision
            runner.backbone = self                      # shares the parameters, no copy
            runner.prediction_head = _ClassificationHead(self._config).to(device)
            runner.logit_names = ("stlt",)
            runner._handle = runner._handle_device = runner._weights_key = runner._workspace = runner._keepalive = None
            runner._packed, runner._packed_key, runner._pruning = {}, {}, True
            runner._param_cache = runner._graphs = None
            runner.train(False)
            object.__setattr__(self, "_runner", runner)
        with torch.no_grad():
            out = runner.forward_with_taps(batch)["temporal"]   # [B, L, H]
        return out.transpose(0, 1)


class _ClassificationHead(nn.Module):  # reference models.py:155-160
    def __init__(self, config):
        super().__init__()
        self.fc1 = _Affine(config.hidden_size, config.hidden_size)
        self.layer_norm = _Norm(config.hidden_size)
        self.fc2 = _Affine(config.hidden_size, config.num_classes)


# ------------------------------------------------------------------------------------------------
# the drop-in module
# ------------------------------------------------------------------------------------------------
class Stlt(nn.Module):
    """B200 implementation of the reference ``Stlt`` (models.py:166-195).

    ``precision``: "fp32" (default; 3-term bf16 split on tcgen05, logits within 1e-4 of the fp32
    reference) or "bf16" (bf16 GEMM operands, fp32 accumulate / residual / LayerNorm / softmax).

    Deviations from the reference module a caller should know (everything else is drop-in):
      * ``precision`` applies to inference. A training step (train mode with grad enabled) ALWAYS runs in bf16 mixed
        precision (bf16 GEMM operands, fp32 master weights / gradients), whatever ``precision`` says.
      * In eval mode the logits carry no ``grad_fn`` even with grad enabled (the reference is differentiable in eval
        mode); a warning is issued once. Call ``model.train(True)`` (with ``hidden_dropout_prob=0`` for deterministic
        behaviour) to differentiate through the model.
      * ``batch["src_key_padding_mask_boxes"]`` / ``["src_key_padding_mask_frames"]`` are not read: the masks are
        re-derived as ``categories == 0`` / ``frame_types == 0``, which is what the reference collater produces
        (datasets.py:274-286). ``verify_masks(batch)`` raises if a caller-supplied mask differs.
      * Parameter writes that bypass PyTorch's version counters need ``mark_weights_dirty()``.
    """

    # class-level defaults: StltBackbone builds its runner without calling __init__
    _param_cache = None     # [(name, parameter)] of the last walk over the module tree
    _graphs = None          # CUDA-graph cache of the inference forward (enable_cuda_graphs)
    _cuda_graphs = False
    _fused_ln = True
    _fused_ln_fp32 = os.environ.get("STLT_FUSED_LN_FP32", "1") != "0"
    _fused_attn = os.environ.get("STLT_FUSED_ATTENTION", "1") != "0"
    _compaction = os.environ.get("STLT_COMPACTION", "1") != "0"
    _hilo = os.environ.get("STLT_HILO_RESIDUAL", "0") != "0"

    def __init__(self, config, precision: str = "fp32", cuda_graphs: bool = False):
        super().__init__()
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        if config.hidden_size != 768 or config.num_attention_heads != 12:
            raise ValueError("the sm_100a kernels are specialised for hidden_size=768, 12 heads")
        self.config = config
        self.precision = precision
        if getattr(config, "load_backbone_path", None) is not None:
            self.backbone = StltBackbone.from_pretrained(config)
            if config.freeze_backbone:
                for param in self.backbone.parameters():
                    param.requires_grad = False
        else:
            self.backbone = StltBackbone(config)
        self.prediction_head = _ClassificationHead(config)
        self.logit_names = ("stlt",)
        # library state (not part of the state_dict)
        self._handle = None
        self._handle_device = None
        self._weights_key = None
        self._packed = {}       # precision -> packed bf16 weight buffer
        self._packed_key = {}   # precision -> weights key the buffer was packed from
        self._workspace = None
        self._keepalive = None
        self._pruning = True
        self._cuda_graphs = bool(cuda_graphs)

    # -- nn.Module protocol ------------------------------------------------------------------
    def train(self, mode: bool = True):  # reference models.py:180-183
        super().train(mode)
        if getattr(self.config, "load_backbone_path", None) and self.config.freeze_backbone:
            self.backbone.train(False)
        return self

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): parameters may be replaced
        self._param_cache = None
        self._graphs = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._param_cache = None
        return super().load_state_dict(*args, **kwargs)

    def mark_weights_dirty(self) -> None:
        """Call after changing parameters in a way PyTorch's version counters do not see — writes through
        ``param.data`` (``p.data.mul_()``, EMA swaps), raw-pointer optimizers, replacing a parameter object. The
        next forward re-binds the pointers and re-packs the bf16 copies; captured CUDA graphs are dropped."""
        self._param_cache = None
        self._weights_key = None
        self._packed_key = {}
        self._graphs = None

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load_library().stlt_destroy(self._handle)
        except Exception:
            pass

    # -- library plumbing ----------------------------------------------------------------------
    def _dims(self) -> _lib.StltDims:
        c = self.config
        return _lib.StltDims(
            hidden_size=c.hidden_size, num_heads=c.num_attention_heads,
            num_spatial_layers=c.num_spatial_layers, num_temporal_layers=c.num_temporal_layers,
            unique_categories=c.unique_categories, num_classes=c.num_classes,
            max_positions=c.layout_num_frames, num_frame_types=5,
            layer_norm_eps=c.layer_norm_eps, encoder_norm_eps=1e-5)

    def _ensure_handle(self, device: torch.device):
        if self._handle is not None and self._handle_device == device:
            return
        lib = _lib.load_library()
        if self._handle is not None:
            lib.stlt_destroy(self._handle)
            self._handle = None
        handle = ctypes.c_void_p()
        dims = self._dims()
        rc = lib.stlt_create(ctypes.byref(dims), ctypes.byref(handle))
        _lib.check(None, rc)
        self._handle = handle
        self._handle_device = device
        _lib.check(handle, lib.stlt_set_pruning(handle, int(self._pruning)))
        _lib.check(handle, lib.stlt_set_fused_ln(handle, int(self._fused_ln)))
        _lib.check(handle, lib.stlt_set_fused_ln_fp32(handle, int(self._fused_ln_fp32)))
        _lib.check(handle, lib.stlt_set_fused_attention(handle, int(self._fused_attn)))
        _lib.check(handle, lib.stlt_set_compaction(handle, int(self._compaction)))
        _lib.check(handle, lib.stlt_set_hilo_residual(handle, int(self._hilo)))
        self._weights_key = None
        self._graphs = None
        self._packed.clear()
        self._packed_key.clear()

    def _sync_weights(self, device: torch.device, stream: int, precision: Optional[str] = None):
        lib = _lib.load_library()
        named = self._param_cache
        if named is None:  # the walk over the module tree is cached (invalidated by _apply / load_state_dict)
            named = self._param_cache = [(n, p) for n, p in self.named_parameters()]
        key = tuple([(p.data_ptr(), p._version) for _, p in named])
        if key != self._weights_key:
            arr = (_lib.StltTensor * len(named))()
            keep = []
            for i, (name, p) in enumerate(named):
                if p.device != device:
                    raise RuntimeError(f"parameter {name} is on {p.device}, batch is on {device}")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError(f"parameter {name} must be contiguous float32")
                bname = name.encode()
                keep.append(bname)
                arr[i].name = bname
                arr[i].data = p.data_ptr()
                arr[i].dtype = _lib.DTYPE_F32
                arr[i].ndim = p.dim()
                for d, s in enumerate(p.shape):
                    arr[i].shape[d] = s
            _lib.check(self._handle, lib.stlt_bind_weights(self._handle, arr, len(named)))
            self._weights_key = key
            self._packed_key.clear()
        prec = _lib.PRECISIONS[precision or self.precision]
        if self._packed_key.get(prec) != key or self._last_packed != prec:
            nbytes = ctypes.c_size_t()
            _lib.check(self._handle, lib.stlt_packed_weights_bytes(self._handle, prec, ctypes.byref(nbytes)))
            buf = self._packed.get(prec)
            if buf is None or buf.numel() < nbytes.value or buf.device != device:
                buf = _lib.aligned_empty(nbytes.value, device)
                self._packed[prec] = buf
            _lib.check(self._handle, lib.stlt_pack_weights(self._handle, stream, prec, buf.data_ptr(), buf.numel()))
            self._packed_key[prec] = key
            self._last_packed = prec

    _last_packed = None

    def _get_workspace(self, B: int, L: int, S: int, device: torch.device) -> torch.Tensor:
        lib = _lib.load_library()
        nbytes = ctypes.c_size_t()
        prec = _lib.PRECISIONS[self.precision]
        _lib.check(self._handle, lib.stlt_workspace_bytes(self._handle, B, L, S, prec, ctypes.byref(nbytes)))
        ws = self._workspace
        if ws is None or ws.numel() < nbytes.value or ws.device != device:
            ws = _lib.aligned_empty(max(nbytes.value, 1024), device)
            self._workspace = ws
        return ws

    @staticmethod
    def _as_input(batch, key, dtype, shape, device):
        t = batch[key]
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"batch['{key}'] must be a tensor")
        if t.device != device:
            raise RuntimeError(f"batch['{key}'] is on {t.device}, expected {device}")
        if t.dtype != dtype:
            raise TypeError(f"batch['{key}'] must be {dtype}, got {t.dtype}")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"batch['{key}'] has shape {tuple(t.shape)}, expected {tuple(shape)}")
        return t.contiguous()

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return self._run(batch, want_taps=False)

    def forward_with_taps(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Forward that also returns the fp32 activations after every stage (parity tests)."""
        return self._run(batch, want_taps=True)

    def _run(self, batch, want_taps: bool):
        cats = batch["categories"]
        if not isinstance(cats, torch.Tensor) or cats.dim() != 3:
            raise ValueError("batch['categories'] must be an int64 tensor [B, L, S]")
        device = cats.device
        if device.type != "cuda":
            raise RuntimeError(
                "stlt_b200.Stlt has no CPU path: move the batch to a CUDA device (the CPU "
                "implementation of this path is the reference module itself)")
        B, L, S = cats.shape
        cats = self._as_input(batch, "categories", torch.int64, (B, L, S), device)
        boxes = self._as_input(batch, "boxes", torch.float32, (B, L, S, 4), device)
        ftypes = self._as_input(batch, "frame_types", torch.int64, (B, L), device)
        lengths = self._as_input(batch, "lengths", torch.int64, (B,), device)
        scores = None
        if "scores" in batch:  # presence toggles the score embedding (models.py:33-35)
            scores = self._as_input(batch, "scores", torch.float32, (B, L, S), device)

        dropout_p = float(getattr(self.config, "hidden_dropout_prob", 0.0)) if self.training else 0.0
        if getattr(self.config, "load_backbone_path", None) and self.config.freeze_backbone:
            dropout_p = 0.0  # the frozen backbone stays in eval mode (models.py:180-183); the head has no dropout
        # The autograd / mixed-precision training path is taken in train mode only. In eval mode the logits come
        # from the inference path in the configured precision and carry no grad_fn, with or without torch.no_grad().
        needs_grad = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not self.training and torch.is_grad_enabled() and not Stlt._warned_eval_grad and \
                any(p.requires_grad for p in self.parameters()):
            Stlt._warned_eval_grad = True
            import warnings
            warnings.warn("stlt_b200.Stlt in eval mode returns logits without a grad_fn even though grad is enabled "
                          "(the reference is differentiable in eval mode): wrap inference in torch.no_grad(), or call "
                          "model.train(True) to take the differentiable training path", stacklevel=3)
        if needs_grad or dropout_p > 0.0:
            # training step (src/train.py:125-127): logits carry a grad_fn whose backward is
            # stlt_backward; the reference's own criterion / clip_grad_norm_ / AdamW then work as is
            if want_taps:
                raise RuntimeError("forward_with_taps is an inference-only debugging aid")
            names, params = zip(*self.named_parameters())
            logits = _StltTrainFunction.apply(self, names, dropout_p, cats, boxes, scores, ftypes, lengths,
                                              *params)
            return {k: v for k, v in zip(self.logit_names, (logits,))}

        if self._cuda_graphs and not want_taps and B > 0 and not torch.cuda.is_current_stream_capturing():
            logits = self._run_graphed(cats, boxes, scores, ftypes, lengths)
            return {k: v for k, v in zip(self.logit_names, (logits,))}
        lib = _lib.load_library()
        with torch.cuda.device(device):
            self._ensure_handle(device)
            stream = torch.cuda.current_stream(device).cuda_stream
            self._sync_weights(device, stream)
            ws = self._get_workspace(B, L, S, device)
            logits = torch.empty((B, self.config.num_classes), dtype=torch.float32, device=device)
            out = {}
            if want_taps:
                H = self.config.hidden_size
                taps_t = {
                    "embed": torch.empty((B, L, S, H), dtype=torch.float32, device=device),
                    "spatial": torch.empty((B, L, S, H), dtype=torch.float32, device=device),
                    "frames": torch.empty((B, L, H), dtype=torch.float32, device=device),
                    "temporal": torch.empty((B, L, H), dtype=torch.float32, device=device),
                    "pooled": torch.empty((B, H), dtype=torch.float32, device=device),
                    "src_key_padding_mask_boxes": torch.empty((B, L, S), dtype=torch.bool, device=device),
                    "src_key_padding_mask_frames": torch.empty((B, L), dtype=torch.bool, device=device),
                }
                taps = _lib.StltTaps(*(taps_t[k].data_ptr() for k in ("embed", "spatial", "frames", "temporal", "pooled")))
                _lib.check(self._handle, lib.stlt_set_taps(self._handle, ctypes.byref(taps)))
                out.update(taps_t)
                mb = taps_t["src_key_padding_mask_boxes"].data_ptr()
                mf = taps_t["src_key_padding_mask_frames"].data_ptr()
            else:
                mb = mf = None
            try:
                rc = lib.stlt_forward(
                    self._handle, stream, _lib.PRECISIONS[self.precision], cats.data_ptr(),
                    boxes.data_ptr(), scores.data_ptr() if scores is not None else None,
                    ftypes.data_ptr(), lengths.data_ptr(), B, L, S, ws.data_ptr(), ws.numel(),
                    logits.data_ptr(), mb, mf)
                _lib.check(self._handle, rc)
            finally:
                if want_taps:
                    lib.stlt_set_taps(self._handle, None)
            # keep the inputs alive until the (asynchronous) kernels that read them are enqueued
            # behind the next call; contiguous() may have created temporaries
            self._keepalive = (cats, boxes, ftypes, lengths, scores)
        out["stlt"] = logits
        if want_taps:
            return out
        return {k: v for k, v in zip(self.logit_names, (logits,))}

    # -- CUDA-graph replay of the steady-state inference forward -------------------------------------
    def enable_cuda_graphs(self, enable: bool = True) -> None:
        """Inference forwards of a shape seen before replay as ONE CUDA graph launch instead of ~60 kernel launches
        through ctypes (the library enqueues on the caller's stream and never synchronises, so a forward is
        capturable as is). The first forward of a (shape, precision) runs eagerly, the second is captured; inputs
        are copied into the graph's static buffers and a fresh logits tensor is returned, so semantics are
        unchanged. Parameter updates seen by PyTorch's version counters re-pack eagerly before the replay; moved
        parameters or ``mark_weights_dirty()`` drop the graphs."""
        self._cuda_graphs = bool(enable)
        if not enable:
            self._graphs = None

    def _run_graphed(self, cats, boxes, scores, ftypes, lengths):
        device = cats.device
        B, L, S = cats.shape
        inputs = {"categories": cats, "boxes": boxes, "frame_types": ftypes, "lengths": lengths}
        if scores is not None:
            inputs["scores"] = scores
        with torch.cuda.device(device):
            self._ensure_handle(device)
            stream = torch.cuda.current_stream(device).cuda_stream
            self._sync_weights(device, stream)  # eager: re-binds / re-packs when a parameter changed
            ptrs = tuple([k[0] for k in self._weights_key])
            shape_key = (device, B, L, S, scores is not None, self.precision, self._pruning, self._fused_ln,
                         self._fused_ln_fp32, self._fused_attn, self._compaction, self._hilo)
            if self._graphs is None:
                self._graphs = {}
            entry = self._graphs.get(shape_key)
            if entry is not None and entry["ptrs"] != ptrs:  # parameters moved: the captured pointers are stale
                entry = None
            if entry is None or entry["graph"] is None:
                if entry is None:  # first sight of this shape: plain eager forward (sizes the workspace, sets attributes)
                    self._graphs[shape_key] = {"graph": None, "ptrs": ptrs}
                    return self._eager(inputs)
                static = {k: v.clone() for k, v in inputs.items()}
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self._eager(static)
                entry = self._graphs[shape_key] = {"graph": graph, "ptrs": ptrs, "static": static, "out": out}
            for k, v in inputs.items():
                entry["static"][k].copy_(v, non_blocking=True)
            entry["graph"].replay()
            return entry["out"].clone()

    def _eager(self, inputs):
        graphs, self._cuda_graphs = self._cuda_graphs, False
        try:
            with torch.no_grad():
                return self._run(inputs, want_taps=False)["stlt"]
        finally:
            self._cuda_graphs = graphs

    # -- training plumbing (used by _StltTrainFunction and training.FusedTrainStep) ----------------
    def _bind_grads(self, grads: Dict[str, torch.Tensor]) -> None:
        lib = _lib.load_library()
        arr = (_lib.StltTensor * max(len(grads), 1))()
        keep = []
        for i, (name, g) in enumerate(grads.items()):
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise RuntimeError(f"gradient buffer of {name} must be contiguous float32")
            bname = name.encode()
            keep.append(bname)
            arr[i].name = bname
            arr[i].data = g.data_ptr()
            arr[i].dtype = _lib.DTYPE_F32
            arr[i].ndim = g.dim()
            for d, sz in enumerate(g.shape):
                arr[i].shape[d] = sz
        _lib.check(self._handle, lib.stlt_bind_grads(self._handle, arr, len(grads)))

    def _train_workspace(self, B: int, L: int, S: int, device: torch.device) -> torch.Tensor:
        lib = _lib.load_library()
        nbytes = ctypes.c_size_t()
        _lib.check(self._handle, lib.stlt_train_workspace_bytes(self._handle, B, L, S, ctypes.byref(nbytes)))
        return _lib.aligned_empty(nbytes.value, device)

    def _forward_train(self, inputs, ws: torch.Tensor, dropout_p: float, seed: int) -> torch.Tensor:
        cats, boxes, scores, ftypes, lengths = inputs
        B, L, S = cats.shape
        lib = _lib.load_library()
        logits = torch.empty((B, self.config.num_classes), dtype=torch.float32, device=cats.device)
        stream = torch.cuda.current_stream(cats.device).cuda_stream
        _lib.check(self._handle, lib.stlt_forward_train(
            self._handle, stream, cats.data_ptr(), boxes.data_ptr(),
            scores.data_ptr() if scores is not None else None, ftypes.data_ptr(), lengths.data_ptr(),
            B, L, S, ws.data_ptr(), ws.numel(), dropout_p, seed, logits.data_ptr()))
        return logits

    def _backward(self, inputs, ws: torch.Tensor, d_logits: Optional[torch.Tensor], phases: int,
                  dropout_p: float = 0.0, seed: int = 0) -> None:
        cats, boxes, scores, ftypes, lengths = inputs
        B, L, S = cats.shape
        lib = _lib.load_library()
        stream = torch.cuda.current_stream(cats.device).cuda_stream
        _lib.check(self._handle, lib.stlt_backward(
            self._handle, stream, cats.data_ptr(), boxes.data_ptr(),
            scores.data_ptr() if scores is not None else None, ftypes.data_ptr(), lengths.data_ptr(),
            B, L, S, ws.data_ptr(), ws.numel(), dropout_p, seed,
            d_logits.data_ptr() if d_logits is not None else None, phases))

    def make_graphed(self, example_batch: Dict[str, torch.Tensor], warmup: int = 2):
        """Captures the inference forward for ``example_batch``'s shapes in a CUDA graph (the library
        enqueues everything on the caller's stream and never synchronises, so the ~130 launches of a
        forward replay as one graph launch — the latency-bound small-batch regime of the reference's
        batch-8 config). Returns ``run(batch) -> {"stlt": logits}``; inputs are copied into the graph's
        static buffers, the returned logits tensor is reused by the next call."""
        if self.training:
            raise RuntimeError("make_graphed is for inference: call model.train(False) first")
        keys = [k for k in ("categories", "boxes", "frame_types", "lengths", "scores") if k in example_batch]
        static = {k: example_batch[k].clone() for k in keys}
        side = torch.cuda.Stream(static["categories"].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # packs weights, sizes the workspace, sets kernel attributes
                self(static)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            out = self(static)["stlt"]
        weights_key = self._weights_key

        def run(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
            named = self._param_cache or [(n, p) for n, p in self.named_parameters()]
            if self._weights_key != weights_key or \
                    tuple([(p.data_ptr(), p._version) for _, p in named]) != weights_key:
                raise RuntimeError("parameters changed since capture: call make_graphed again")
            for k in keys:
                static[k].copy_(batch[k], non_blocking=True)
            graph.replay()
            return {"stlt": out}

        run.graph, run.static_inputs = graph, static
        return run

    _warned_eval_grad = False

    @staticmethod
    def verify_masks(batch: Dict[str, torch.Tensor]) -> None:
        """Raises if the batch carries padding masks that differ from the ones this module derives
        (``categories == 0``, ``frame_types == 0`` — the reference collater's, datasets.py:274-286). Synchronises."""
        for key, src in (("src_key_padding_mask_boxes", "categories"), ("src_key_padding_mask_frames", "frame_types")):
            if key in batch and not torch.equal(batch[key].to(torch.bool), batch[src] == 0):
                raise ValueError(f"batch['{key}'] differs from batch['{src}'] == 0: custom padding masks are not "
                                 f"supported by stlt_b200.Stlt (it derives the masks itself)")

    def check_inputs(self) -> None:
        """Synchronises and raises if the last forward saw an out-of-range index (debug aid)."""
        if self._handle is None or self._workspace is None:
            return
        lib = _lib.load_library()
        stream = torch.cuda.current_stream(self._workspace.device).cuda_stream
        _lib.check(self._handle, lib.stlt_check_errors(self._handle, stream, self._workspace.data_ptr()))

    def set_pruning(self, enable: bool) -> None:
        """Last-layer row pruning (default on; logits are bit-identical either way)."""
        self._pruning = bool(enable)
        if self._handle is not None:
            _lib.check(self._handle, _lib.load_library().stlt_set_pruning(self._handle, int(self._pruning)))

    def set_fused_layer_norm(self, enable: bool, fp32: Optional[bool] = None) -> None:
        """LayerNorm folded into the GEMM epilogues (default on; bf16 mode, and the fp32-parity mode on split operands;
        ``fp32`` sets the switch of the fp32-parity mode separately). Off = separate add+LN kernels."""
        self._fused_ln = bool(enable)
        self._fused_ln_fp32 = bool(enable if fp32 is None else fp32)
        if self._handle is not None:
            lib = _lib.load_library()
            _lib.check(self._handle, lib.stlt_set_fused_ln(self._handle, int(self._fused_ln)))
            _lib.check(self._handle, lib.stlt_set_fused_ln_fp32(self._handle, int(self._fused_ln_fp32)))

    def set_fused_attention(self, enable: bool) -> None:
        """bf16 mode: attention folded into the epilogue of the in-projection GEMM (default on; needs the fused
        LayerNorm path and sequences of at most 32 tokens). Off = separate in-projection GEMM + attention kernel."""
        self._fused_attn = bool(enable)
        if self._handle is not None:
            _lib.check(self._handle, _lib.load_library().stlt_set_fused_attention(self._handle, int(self._fused_attn)))

    def set_compaction(self, enable: bool) -> None:
        """Pad-skipping row layout of the spatial stack (default on, both precision modes): padding frames and the padded
        slots of one-token frames are not computed. Off = the whole padded [B, L, S] grid, as the reference computes it."""
        self._compaction = bool(enable)
        if self._handle is not None:
            _lib.check(self._handle, _lib.load_library().stlt_set_compaction(self._handle, int(self._compaction)))

    def set_hilo_residual(self, enable: bool) -> None:
        """bf16 mode: residual stream of the encoder stacks stored as two bf16 planes instead of fp32 + a bf16 copy.
        Off by default (fewer bytes, but measured slower: see DESIGN.md "Measured and rejected")."""
        self._hilo = bool(enable)
        if self._handle is not None:
            _lib.check(self._handle, _lib.load_library().stlt_set_hilo_residual(self._handle, int(self._hilo)))

    def set_profiling(self, enable: bool) -> None:
        """Per-category CUDA-event timing of the kernels launched by forward (bench / profiles)."""
        if self._handle is None:
            raise RuntimeError("run one forward before enabling profiling")
        _lib.check(self._handle, _lib.load_library().stlt_set_profiling(self._handle, int(enable)))

    def get_profile(self) -> Dict[str, Dict[str, float]]:
        prof = _lib.StltProfile()
        _lib.check(self._handle, _lib.load_library().stlt_get_profile(self._handle, ctypes.byref(prof)))
        return {name: {"ms": prof.ms[i], "flops": prof.flops[i], "launches": int(prof.launches[i])}
                for i, name in enumerate(_lib.PROF_CATEGORIES)}

    def get_profile_by_role(self) -> Dict[str, Dict[str, float]]:
        """The "gemm" entry of the most recent get_profile(), split by the role of the launch in the encoder layer."""
        prof = _lib.StltRoleProfile()
        _lib.check(self._handle, _lib.load_library().stlt_get_profile_by_role(self._handle, ctypes.byref(prof)))
        return {name: {"ms": prof.ms[i], "flops": prof.flops[i], "launches": int(prof.launches[i])}
                for i, name in enumerate(_lib.PROF_ROLES) if prof.launches[i]}

    def last_launch_count(self) -> int:
        if self._handle is None:
            return 0
        return int(_lib.load_library().stlt_last_launch_count(self._handle))


class _StltTrainFunction(torch.autograd.Function):
    """Stlt.forward with a grad_fn: forward = stlt_forward_train (activations kept in a workspace
    owned by the autograd context), backward = stlt_backward. Parameters are passed as inputs so
    autograd accumulates into ``param.grad`` and DDP-style hooks fire as usual."""

    @staticmethod
    def forward(ctx, module, names, dropout_p, cats, boxes, scores, ftypes, lengths, *params):
        device = cats.device
        with torch.cuda.device(device):
            module._ensure_handle(device)
            stream = torch.cuda.current_stream(device).cuda_stream
            module._sync_weights(device, stream, "bf16")  # mixed precision: bf16 operands, fp32 master
            B, L, S = cats.shape
            ws = module._train_workspace(B, L, S, device)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if dropout_p > 0 else 0
            inputs = (cats, boxes, scores, ftypes, lengths)
            logits = module._forward_train(inputs, ws, dropout_p, seed)
        ctx.module, ctx.names, ctx.ws, ctx.inputs = module, names, ws, inputs
        ctx.dropout = (dropout_p, seed)
        ctx.param_shapes = [p.shape for p in params]
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        module, names = ctx.module, ctx.names
        device = d_logits.device
        # ctx.needs_input_grad[8:] lines up with *params
        want = [ctx.needs_input_grad[8 + i] for i in range(len(names))]
        has_scores = ctx.inputs[2] is not None
        sizes, total = [], 0
        for name, shape, w in zip(names, ctx.param_shapes, want):
            # the orphan prototype layer (models.py:46-52) and, without scores, the score embedding
            # never reach the logits: their gradient is None in the reference too
            used = w and ".encoder_layer." not in name and (has_scores or "score_embeddings" not in name)
            n = int(torch.Size(shape).numel()) if used else 0
            sizes.append(n)
            total += (n + 3) // 4 * 4
        flat = torch.zeros(max(total, 4), dtype=torch.float32, device=device)
        grads, out, off = {}, [], 0
        for name, shape, n in zip(names, ctx.param_shapes, sizes):
            if n == 0:
                out.append(None)
                continue
            g = flat[off:off + n].view(shape)
            off += (n + 3) // 4 * 4
            grads[name] = g
            out.append(g)
        with torch.cuda.device(device):
            module._bind_grads(grads)
            module._backward(ctx.inputs, ctx.ws, d_logits.contiguous().float(), _lib.BWD_ALL, *ctx.dropout)
        ctx.ws = None
        return (None,) * 8 + tuple(out)


models_factory = {"stlt": Stlt}
