"""ctypes binding of libstlt_b200.so (C ABI declared in include/stlt_b200.h).

There is deliberately no fallback: if the shared library is missing or cannot be loaded the
import of the product path fails loudly (the CPU path is the reference itself, not this package).
"""
from __future__ import annotations

import ctypes
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t,
                    c_uint8, c_uint64, c_void_p)
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libstlt_b200.so"

STLT_OK = 0
STLT_ERR_INVALID = -1
STLT_ERR_CUDA = -2
STLT_ERR_STATE = -3
STLT_ERR_INPUT = -4

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISIONS = {"fp32": PRECISION_FP32, "bf16": PRECISION_BF16}

DTYPE_F32 = 0
DTYPE_I64 = 1

GEMM_OUT_F32 = 0
GEMM_OUT_BF16 = 1
GEMM_OUT_BF16_SPLIT = 2
GEMM_OUT_BF16_DUAL = 3
GEMM_OUT_HILO = 6
GEMM_NT, GEMM_NN, GEMM_TN_RED = 0, 1, 2
GEMM_EPI_PLAIN, GEMM_EPI_NORM_A, GEMM_EPI_RESID = 0, 1, 2

BWD_TEMPORAL, BWD_SPATIAL, BWD_ALL = 1, 2, 3
LOSS_CROSS_ENTROPY, LOSS_BCE_LOGITS = 0, 1


class StltDims(Structure):
    _fields_ = [
        ("hidden_size", c_int32),
        ("num_heads", c_int32),
        ("num_spatial_layers", c_int32),
        ("num_temporal_layers", c_int32),
        ("unique_categories", c_int32),
        ("num_classes", c_int32),
        ("max_positions", c_int32),
        ("num_frame_types", c_int32),
        ("layer_norm_eps", c_float),
        ("encoder_norm_eps", c_float),
    ]


class StltTensor(Structure):
    _fields_ = [
        ("name", c_char_p),
        ("data", c_void_p),
        ("dtype", c_int32),
        ("ndim", c_int32),
        ("shape", c_int64 * 4),
    ]


class StltTaps(Structure):
    _fields_ = [
        ("embed", c_void_p),
        ("spatial", c_void_p),
        ("frames", c_void_p),
        ("temporal", c_void_p),
        ("pooled", c_void_p),
    ]


class StltLayoutStore(Structure):
    _fields_ = [
        ("video_frame_offsets", c_void_p),
        ("frame_object_offsets", c_void_p),
        ("obj_boxes", c_void_p),
        ("obj_categories", c_void_p),
        ("obj_scores", c_void_p),
        ("video_sizes", c_void_p),
    ]


class StltLayoutIds(Structure):
    _fields_ = [("cls", c_int64), ("ft_pad", c_int64), ("ft_regular", c_int64), ("ft_empty", c_int64),
                ("ft_extract", c_int64)]


PROF_CATEGORIES = ("gemm", "attention", "add_ln", "other")
PROF_ROLES = ("in_proj", "qkv_attention", "out_proj", "linear1", "linear2", "gradient", "other_gemm")


class StltProfile(Structure):
    _fields_ = [
        ("ms", c_double * 4),
        ("flops", c_double * 4),
        ("launches", c_int64 * 4),
    ]


class StltRoleProfile(Structure):
    _fields_ = [
        ("ms", c_double * 7),
        ("flops", c_double * 7),
        ("launches", c_int64 * 7),
    ]


# name -> (restype, argtypes); mirrors include/stlt_b200.h one to one.
SIGNATURES = {
    "stlt_create": (c_int32, [POINTER(StltDims), POINTER(c_void_p)]),
    "stlt_destroy": (c_int32, [c_void_p]),
    "stlt_last_error": (c_char_p, [c_void_p]),
    "stlt_bind_weights": (c_int32, [c_void_p, POINTER(StltTensor), c_int32]),
    "stlt_packed_weights_bytes": (c_int32, [c_void_p, c_int32, POINTER(c_size_t)]),
    "stlt_pack_weights": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_size_t]),
    "stlt_workspace_bytes": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_size_t)]),
    "stlt_prepare": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                               c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "stlt_build_batch": (c_int32, [c_void_p, c_void_p, POINTER(StltLayoutStore), POINTER(StltLayoutIds),
                                   c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_double,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "stlt_topk_count": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "stlt_map_accumulate": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "stlt_charades_map": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "stlt_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p,
                               c_void_p, c_void_p]),
    "stlt_check_errors": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "stlt_last_launch_count": (c_int32, [c_void_p]),
    "stlt_set_taps": (c_int32, [c_void_p, POINTER(StltTaps)]),
    "stlt_set_pruning": (c_int32, [c_void_p, c_int32]),
    "stlt_set_fused_ln": (c_int32, [c_void_p, c_int32]),
    "stlt_set_fused_ln_fp32": (c_int32, [c_void_p, c_int32]),
    "stlt_set_fused_attention": (c_int32, [c_void_p, c_int32]),
    "stlt_set_compaction": (c_int32, [c_void_p, c_int32]),
    "stlt_set_hilo_residual": (c_int32, [c_void_p, c_int32]),
    "stlt_set_profiling": (c_int32, [c_void_p, c_int32]),
    "stlt_get_profile": (c_int32, [c_void_p, POINTER(StltProfile)]),
    "stlt_get_profile_by_role": (c_int32, [c_void_p, POINTER(StltRoleProfile)]),
    "stlt_bind_grads": (c_int32, [c_void_p, POINTER(StltTensor), c_int32]),
    "stlt_train_workspace_bytes": (c_int32, [c_void_p, c_int32, c_int32, c_int32, POINTER(c_size_t)]),
    "stlt_forward_train": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int32, c_int32, c_int32, c_void_p, c_size_t, c_float, c_uint64,
                                     c_void_p]),
    "stlt_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int32, c_int32, c_int32, c_void_p, c_size_t, c_float, c_uint64, c_void_p,
                                c_int32]),
    "stlt_op_dropout_mask": (c_int32, [c_void_p, c_float, c_uint64, c_int32, c_int64, c_int64, c_void_p]),
    "stlt_loss": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_float,
                            c_void_p, c_void_p]),
    "stlt_backward_stage_events": (c_int32, [c_void_p, c_int32, c_void_p]),
    "stlt_stream_wait_backward_stage": (c_int32, [c_void_p, c_void_p, c_int32]),
    "stlt_grad_sumsq": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32]),
    "stlt_adamw_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_float, c_float, c_float, c_float, c_float, c_int32, c_void_p, c_float]),
    "stlt_cacnf_bind_weights": (c_int32, [c_void_p, POINTER(StltTensor), c_int32, c_int32, c_int32, c_int32, c_int32]),
    "stlt_cacnf_packed_weights_bytes": (c_int32, [c_void_p, c_int32, POINTER(c_size_t)]),
    "stlt_cacnf_pack_weights": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_size_t]),
    "stlt_cacnf_workspace_bytes": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_size_t)]),
    "stlt_cacnf_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "stlt_op_gemm": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                               c_int32, c_int32, c_int32, c_int32, c_int32]),
    "stlt_op_gemm_grad": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                    c_int32, c_int64, c_int32]),
    "stlt_op_gemm_fused": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_float, c_int32]),
    "stlt_op_gemm_fused_split": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_float, c_int32]),
    "stlt_op_gemm_resid_hilo": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int32]),
    "stlt_op_pack_folded": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                      c_void_p, c_void_p, c_void_p, c_int32]),
    "stlt_op_qkv_attention": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_float, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    "stlt_op_attention_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                        c_int32, c_void_p, c_int32, c_void_p]),
    "stlt_op_attention_cross": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                          c_int32, c_void_p]),
    "stlt_op_gemm_simt": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int32, c_int32, c_int32, c_int32]),
    "stlt_op_attention": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_int32,
                                    c_int32, c_void_p, c_int32, c_int64]),
    "stlt_op_add_ln": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                 c_int64, c_void_p, c_void_p, c_int32, c_int64]),
    "stlt_op_pack_bf16": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32]),
}

_lib = None


def load_library() -> ctypes.CDLL:
    """Loads libstlt_b200.so and declares every entry point. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python {PKG_DIR / 'build.py'}` "
            "(or __graft_entry__.build()). There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def aligned_empty(nbytes: int, device, alignment: int = 1024):
    """uint8 device buffer of ``nbytes`` whose data pointer is ``alignment``-byte aligned. The library requires 1 KiB
    aligned workspaces (128-byte-swizzled TMA tiles); PyTorch's caching allocator only guarantees 512 B, so the
    buffer is over-allocated and a view starting at the aligned offset is returned (the view keeps it alive)."""
    import torch
    raw = torch.empty(max(int(nbytes), 1) + alignment, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % alignment
    return raw[off:off + max(int(nbytes), 1)]


class StltError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libstlt_b200 error {code}: {message}")
        self.code = code


def check(handle, rc: int) -> None:
    if rc != STLT_OK:
        msg = load_library().stlt_last_error(handle)
        raise StltError(rc, msg.decode() if msg else "unknown error")
