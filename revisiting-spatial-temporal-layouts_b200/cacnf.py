"""Drop-in for the reference ``CrossAttentionCentralNetFusion`` (CACNF, reference
src/modelling/models.py:504-549) on PRECOMPUTED appearance features — SURVEY.md §8(f) rank 2,
BASELINE.json configs[4]. The 3D-ResNet trunk (``Resnet3D.forward_features``, models.py:219-220) is
outside this library: the batch carries its output as ``batch["video_features"]``
([B, 2048, T', H', W'] or [B, 2048, P], P = T'·H'·W' <= 32) instead of ``video_frames``.

Parameter names, shapes and initialisation follow the reference module tree, so a reference checkpoint
loads with ``load_reference_state_dict`` (the ``backbone.appearance_branch.resnet.resnet.*`` trunk
entries are skipped). ``forward`` is one call into libstlt_b200.so (``stlt_cacnf_forward``); there is no
PyTorch compute path and no CPU fallback. Inference only.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch
from torch import nn

from . import lib as _lib
from .module import (StltBackbone, _Affine, _ClassificationHead, _Encoder, _EncoderLayer, _Norm,
                     _SelfAttention)


class _AttnLayer(nn.Module):  # CrossAttentionLayer / SelfAttentionLayer (models.py:328-373)
    def __init__(self, hidden: int):
        super().__init__()
        self.attn = _SelfAttention(hidden)
        self.ln = _Norm(hidden)


class _FeedforwardModule(nn.Module):  # models.py:313-325
    def __init__(self, hidden: int):
        super().__init__()
        self.linear1 = _Affine(hidden, 4 * hidden)
        self.linear2 = _Affine(4 * hidden, hidden)
        self.ln = _Norm(hidden)


class _CrossModalModule(nn.Module):  # models.py:376-393
    def __init__(self, hidden: int):
        super().__init__()
        self.cross_attn = _AttnLayer(hidden)
        self.layout_attn = _AttnLayer(hidden)
        self.layout_ffn = _FeedforwardModule(hidden)
        self.appearance_attn = _AttnLayer(hidden)
        self.appearance_ffn = _AttnLayer(hidden)  # a SelfAttentionLayer in the reference (models.py:386)


class _Projector(nn.Module):  # nn.Conv3d(2048, hidden, kernel_size=(1, 1, 1)) (models.py:236-238)
    def __init__(self, channels: int, hidden: int):
        super().__init__()
        lin = _Affine(channels, hidden)
        self.weight = nn.Parameter(lin.weight.detach().view(hidden, channels, 1, 1, 1).clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())


class _AppearanceBranch(nn.Module):  # TransformerResnet minus the ResNet trunk (models.py:232-254)
    def __init__(self, config):
        super().__init__()
        h = config.hidden_size
        self.projector = _Projector(config.feature_channels, h)
        self.transformer = _Encoder(_EncoderLayer(h), config.num_appearance_layers)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, h))
        self.pos_embed = nn.Parameter(torch.zeros(config.appearance_num_frames + 1, 1, h))
        self.classifier = _Affine(h, config.num_classes)  # unused by CACNF, part of the checkpoint


class _FusionBackbone(nn.Module):  # CrossAttentionFusionBackbone (models.py:434-446)
    def __init__(self, config):
        super().__init__()
        self.layout_branch = StltBackbone(config)
        self.appearance_branch = _AppearanceBranch(config)
        self.mm_fusion = nn.ModuleList([_CrossModalModule(config.hidden_size) for _ in range(config.num_fusion_layers)])


class _FusionHead(nn.Module):  # models.py:286-291
    def __init__(self, config):
        super().__init__()
        self.fc1 = _Affine(config.hidden_size * 2, config.hidden_size)
        self.layer_norm = _Norm(config.hidden_size)
        self.fc2 = _Affine(config.hidden_size, config.num_classes)


class Cacnf(nn.Module):
    """``precision``: "bf16" (default) or "fp32" (3-term bf16 split on tcgen05, logits within 1e-4 of the fp32
    reference), as for ``Stlt``."""

    # reference state_dict prefix -> name the library binds (identity for CACNF; see Caf / Lcf below)
    _NAME_MAP = ()
    _FUSION_LAYERS = None  # None: config.num_fusion_layers

    def __init__(self, config, precision: str = "bf16"):
        super().__init__()
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.precision = precision
        self._build(config)

    def _build(self, config):
        if config.hidden_size != 768 or config.num_attention_heads != 12:
            raise ValueError("the sm_100a kernels are specialised for hidden_size=768, 12 heads")
        if config.appearance_num_frames > 32:
            raise ValueError("at most 32 appearance positions (+1 CLS token) are supported")
        self.config = config
        self._make_parameters(config)
        self._handle = None
        self._handle_device = None
        self._weights_key = None
        self._packed = None
        self._workspace = None
        self._keepalive = None

    def _make_parameters(self, config):
        self.backbone = _FusionBackbone(config)
        self.layout_classifier = _ClassificationHead(config)
        self.appearance_classifier = _ClassificationHead(config)
        self.fusion_classifier = _FusionHead(config)
        self.logit_names = ("stlt", "resnet3d", "caf", "ensemble")

    def _bound_tensors(self):
        """(library name, tensor) for every tensor the library binds."""
        out = []
        for name, p in self.named_parameters():
            for src, dst in self._NAME_MAP:
                if name.startswith(src):
                    name = dst + name[len(src):]
                    break
            out.append((name, p))
        return out

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load_library().stlt_destroy(self._handle)
        except Exception:
            pass

    def load_reference_state_dict(self, state_dict):
        """Loads a reference CACNF checkpoint; the 3D-ResNet trunk's entries are not part of this module."""
        own = {k: v for k, v in state_dict.items() if ".appearance_branch.resnet." not in k}
        return self.load_state_dict(own, strict=True)

    # -- library plumbing ----------------------------------------------------------------------
    def _ensure_handle(self, device):
        if self._handle is not None and self._handle_device == device:
            return
        lib = _lib.load_library()
        if self._handle is not None:
            lib.stlt_destroy(self._handle)
        c = self.config
        dims = _lib.StltDims(hidden_size=c.hidden_size, num_heads=c.num_attention_heads,
                             num_spatial_layers=c.num_spatial_layers, num_temporal_layers=c.num_temporal_layers,
                             unique_categories=c.unique_categories, num_classes=c.num_classes,
                             max_positions=c.layout_num_frames, num_frame_types=5,
                             layer_norm_eps=c.layer_norm_eps, encoder_norm_eps=1e-5)
        handle = ctypes.c_void_p()
        _lib.check(None, lib.stlt_create(ctypes.byref(dims), ctypes.byref(handle)))
        self._handle, self._handle_device, self._weights_key = handle, device, None

    def _sync_weights(self, device, stream):
        lib = _lib.load_library()
        named = self._bound_tensors()
        prec = _lib.PRECISIONS[self.precision]
        key = (prec,) + tuple((p.data_ptr(), p._version) for _, p in named)
        if key == self._weights_key:
            return
        arr = (_lib.StltTensor * len(named))()
        keep = []
        for i, (name, p) in enumerate(named):
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"parameter {name} must be contiguous float32 on {device}")
            shape = tuple(p.shape)
            if len(shape) > 4:  # Conv3d projector [768, C, 1, 1, 1]
                shape = shape[:2] + (1, 1)
            bname = name.encode()
            keep.append(bname)
            arr[i].name, arr[i].data, arr[i].dtype, arr[i].ndim = bname, p.data_ptr(), _lib.DTYPE_F32, len(shape)
            for d, s in enumerate(shape):
                arr[i].shape[d] = s
        c = self.config
        fusion_layers = c.num_fusion_layers if self._FUSION_LAYERS is None else self._FUSION_LAYERS
        _lib.check(self._handle, lib.stlt_cacnf_bind_weights(self._handle, arr, len(named), c.num_appearance_layers,
                                                             fusion_layers, c.appearance_num_frames,
                                                             c.feature_channels))
        n1, n2 = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(self._handle, lib.stlt_packed_weights_bytes(self._handle, prec, ctypes.byref(n1)))
        _lib.check(self._handle, lib.stlt_cacnf_packed_weights_bytes(self._handle, prec, ctypes.byref(n2)))
        a = (n1.value + 1023) // 1024 * 1024
        if self._packed is None or self._packed.numel() < a + n2.value or self._packed.device != device:
            self._packed = _lib.aligned_empty(a + n2.value, device)
        _lib.check(self._handle, lib.stlt_pack_weights(self._handle, stream, prec, self._packed.data_ptr(), n1.value))
        _lib.check(self._handle, lib.stlt_cacnf_pack_weights(self._handle, stream, prec, self._packed.data_ptr() + a,
                                                             n2.value))
        self._weights_key = key

    def forward(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if torch.is_grad_enabled() and self.training:
            raise NotImplementedError("the CACNF path of this library is inference-only; call model.train(False) "
                                      "under torch.no_grad()")
        cats = batch["categories"]
        device = cats.device
        if device.type != "cuda":
            raise RuntimeError("stlt_b200.Cacnf has no CPU path: move the batch to a CUDA device")
        B, L, S = cats.shape
        feats = batch["video_features"]
        if feats.dim() > 3:
            feats = feats.flatten(2)
        c = self.config
        if tuple(feats.shape) != (B, c.feature_channels, c.appearance_num_frames) or feats.dtype != torch.float32:
            raise ValueError(f"batch['video_features'] must be float32 [B, {c.feature_channels}, "
                             f"{c.appearance_num_frames}] (got {tuple(feats.shape)}, {feats.dtype})")
        feats = feats.contiguous()
        cats = cats.contiguous()
        boxes = batch["boxes"].contiguous()
        ftypes = batch["frame_types"].contiguous()
        lengths = batch["lengths"].contiguous()
        scores = batch["scores"].contiguous() if "scores" in batch else None
        lib = _lib.load_library()
        with torch.cuda.device(device):
            self._ensure_handle(device)
            stream = torch.cuda.current_stream(device).cuda_stream
            self._sync_weights(device, stream)
            nbytes = ctypes.c_size_t()
            prec = _lib.PRECISIONS[self.precision]
            _lib.check(self._handle, lib.stlt_cacnf_workspace_bytes(self._handle, B, L, S, prec, ctypes.byref(nbytes)))
            if self._workspace is None or self._workspace.numel() < nbytes.value or self._workspace.device != device:
                self._workspace = _lib.aligned_empty(max(nbytes.value, 1024), device)
            out = [torch.empty((B, c.num_classes), dtype=torch.float32, device=device) for _ in range(4)]
            _lib.check(self._handle, lib.stlt_cacnf_forward(
                self._handle, stream, prec, cats.data_ptr(), boxes.data_ptr(),
                scores.data_ptr() if scores is not None else None, ftypes.data_ptr(), lengths.data_ptr(),
                feats.data_ptr(), B, L, S, self._workspace.data_ptr(), self._workspace.numel(),
                *(t.data_ptr() for t in out)))
            self._keepalive = (cats, boxes, ftypes, lengths, scores, feats)
        return self._select_logits(dict(zip(("stlt", "resnet3d", "caf", "ensemble"), out)))

    def _select_logits(self, logits):
        return logits

    def set_profiling(self, enable: bool) -> None:
        _lib.check(self._handle, _lib.load_library().stlt_set_profiling(self._handle, int(enable)))

    def get_profile(self):
        prof = _lib.StltProfile()
        _lib.check(self._handle, _lib.load_library().stlt_get_profile(self._handle, ctypes.byref(prof)))
        return {name: {"ms": prof.ms[i], "flops": prof.flops[i], "launches": int(prof.launches[i])}
                for i, name in enumerate(_lib.PROF_CATEGORIES)}


class _UnusedHeads(nn.Module):
    """The library's CACNF entry point always evaluates the layout / appearance classifiers; the two-logit-less
    variants below feed it zero heads that are NOT part of their state_dict."""

    def _add_unused_heads(self, config):
        for prefix in ("layout_classifier", "appearance_classifier"):
            head = _ClassificationHead(config)
            for name, p in head.named_parameters():
                self.register_buffer(f"_unused_{prefix}_{name.replace('.', '_')}", torch.zeros_like(p), persistent=False)

    def _unused_head_tensors(self):
        out = []
        for prefix in ("layout_classifier", "appearance_classifier"):
            for name in ("fc1.weight", "fc1.bias", "layer_norm.weight", "layer_norm.bias", "fc2.weight", "fc2.bias"):
                out.append((f"{prefix}.{name}", getattr(self, f"_unused_{prefix}_{name.replace('.', '_')}")))
        return out


class Caf(Cacnf, _UnusedHeads):
    """Drop-in for the reference ``CrossAttentionFusion`` (CAF, models.py:486-501) on precomputed features:
    the CACNF backbone with the fusion classifier only; state_dict prefixes ``caf_backbone.`` / ``classifier.``."""

    _NAME_MAP = (("caf_backbone.", "backbone."), ("classifier.", "fusion_classifier."))

    def _make_parameters(self, config):
        self.caf_backbone = _FusionBackbone(config)
        self.classifier = _FusionHead(config)
        self._add_unused_heads(config)
        self.logit_names = ("caf",)

    def _bound_tensors(self):
        return super()._bound_tensors() + self._unused_head_tensors()

    def _select_logits(self, logits):
        return {"caf": logits["caf"]}


class Lcf(Cacnf, _UnusedHeads):
    """Drop-in for the reference ``LateConcatenationFusion`` (LCF, models.py:297-323) on precomputed features:
    layout state [lengths - 1] and appearance CLS state concatenated into the fusion head — the CACNF pipeline with
    zero fusion layers; state_dict prefixes ``layout_branch.`` / ``appearance_branch.`` / ``classifier.``."""

    _NAME_MAP = (("layout_branch.", "backbone.layout_branch."), ("appearance_branch.", "backbone.appearance_branch."),
                 ("classifier.", "fusion_classifier."))
    _FUSION_LAYERS = 0

    def _make_parameters(self, config):
        self.layout_branch = StltBackbone(config)
        self.appearance_branch = _AppearanceBranch(config)
        self.classifier = _FusionHead(config)
        self._add_unused_heads(config)
        self.logit_names = ("lcf",)

    def _bound_tensors(self):
        return super()._bound_tensors() + self._unused_head_tensors()

    def _select_logits(self, logits):
        return {"lcf": logits["caf"]}
