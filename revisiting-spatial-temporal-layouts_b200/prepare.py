"""Host-side entry for K0: box repair, normalisation by the video size and the padding masks.

Replaces, for an already padded layout batch, the per-object Python in the reference data path:
``fix_box`` (reference src/utils/data_utils.py:205-231), ``torch.tensor(box) / video_size``
(src/modelling/datasets.py:54,82) and the two masks of ``StltCollater``
(src/modelling/datasets.py:274-286). All arithmetic runs in one CUDA kernel via the C ABI
(``stlt_prepare`` in include/stlt_b200.h); the results are bit-identical to the reference.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch

from . import lib as _lib

_handles = {}


def _prep_handle(device: torch.device):
    """stlt_prepare does not depend on the model; a minimal handle per device is enough."""
    key = (device.type, device.index)
    if key not in _handles:
        lib = _lib.load_library()
        dims = _lib.StltDims(768, 12, 0, 0, 1, 1, 1, 1, 1e-12, 1e-5)
        handle = ctypes.c_void_p()
        _lib.check(None, lib.stlt_create(ctypes.byref(dims), ctypes.byref(handle)))
        _handles[key] = handle
    return _handles[key]


def prepare_layout_batch(raw_boxes: torch.Tensor, video_sizes: torch.Tensor, categories: torch.Tensor,
                         frame_types: torch.Tensor) -> Dict[str, torch.Tensor]:
    """raw_boxes f64 [B,L,S,4] pixel (x1,y1,x2,y2); video_sizes i64 [B,2] (width, height);
    categories i64 [B,L,S]; frame_types i64 [B,L] — all on one CUDA device.

    Returns ``boxes`` f32 [B,L,S,4], ``src_key_padding_mask_boxes`` bool [B,L,S] and
    ``src_key_padding_mask_frames`` bool [B,L] exactly as the reference dataset + collater would.
    """
    device = categories.device
    if device.type != "cuda":
        raise RuntimeError("prepare_layout_batch has no CPU path (the CPU path is the reference)")
    B, L, S = categories.shape
    for name, t, dt, shape in (("raw_boxes", raw_boxes, torch.float64, (B, L, S, 4)),
                               ("video_sizes", video_sizes, torch.int64, (B, 2)),
                               ("categories", categories, torch.int64, (B, L, S)),
                               ("frame_types", frame_types, torch.int64, (B, L))):
        if t.device != device or t.dtype != dt or tuple(t.shape) != shape:
            raise ValueError(f"{name}: expected {dt} {shape} on {device}, got {t.dtype} {tuple(t.shape)} on {t.device}")
    raw_boxes, video_sizes = raw_boxes.contiguous(), video_sizes.contiguous()
    categories, frame_types = categories.contiguous(), frame_types.contiguous()
    boxes = torch.empty((B, L, S, 4), dtype=torch.float32, device=device)
    mask_boxes = torch.empty((B, L, S), dtype=torch.bool, device=device)
    mask_frames = torch.empty((B, L), dtype=torch.bool, device=device)
    lib = _lib.load_library()
    with torch.cuda.device(device):
        handle = _prep_handle(device)
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = lib.stlt_prepare(handle, stream, raw_boxes.data_ptr(), video_sizes.data_ptr(),
                              categories.data_ptr(), frame_types.data_ptr(), B, L, S,
                              boxes.data_ptr(), mask_boxes.data_ptr(), mask_frames.data_ptr())
        _lib.check(handle, rc)
    return {"boxes": boxes, "src_key_padding_mask_boxes": mask_boxes,
            "src_key_padding_mask_frames": mask_frames}
