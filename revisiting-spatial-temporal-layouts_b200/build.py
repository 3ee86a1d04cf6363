"""Builds libstlt_b200.so (hand-written sm_100a CUDA + the C ABI of include/stlt_b200.h) in-tree.

Usage: ``python revisiting-spatial-temporal-layouts_b200/build.py [--force]``.
nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libstlt_b200.so"
BUILD_DIR = PKG_DIR / "build"
SOURCES = ["gemm_tcgen05.cu", "gemm_qkv_attn.cu", "compact.cu", "gemm_simt.cu", "elementwise.cu", "attention.cu", "attention_mma.cu",
           "attention_bwd.cu", "attention_bwd_mma.cu", "train_kernels.cu", "stlt_api.cu", "stlt_train.cu",
           "attention_cross.cu", "attention_long.cu", "cacnf_kernels.cu", "stlt_cacnf.cu", "eval_kernels.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found; libstlt_b200.so cannot be built")
    return nvcc


def _source_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG_DIR.parent / "include" / "stlt_b200.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _source_digest()
    stamp = BUILD_DIR / "digest.txt"
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    BUILD_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> Path:
        obj = BUILD_DIR / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr, file=sys.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
            *map(str, objs), "-cudart", "static"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
