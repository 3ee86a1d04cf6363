"""Test infrastructure, not product code: builds oracle/_ref/ — the UNMODIFIED reference, byte-compiled.

The reference (gorjanradevski/revisiting-spatial-temporal-layouts) is pure Python, so "compiling it from the
sources where they lie" means CPython byte-compilation: every module of /root/reference/src that the STLT path
touches is compiled with py_compile straight from /root/reference into a sourceless ``.pyc`` under
``oracle/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like the built ``.so``; the GPU
box runs the same image, hence the same bytecode magic). No reference source file is copied into the repo.

``oracle/_ref`` is what ``bench.py --impl reference`` times (``cpu_baseline.kind = "reference"``) and what
``tests/test_oracle.py`` pins the oracle restatement against on machines without /root/reference.

    python oracle/build_ref.py            # (re)build if /root/reference is present; no-op otherwise
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import py_compile
import sys
from pathlib import Path

REFERENCE_SRC = Path("/root/reference/src")
REF_DIR = Path(__file__).resolve().parent / "_ref"
# modules on (or feeding) the STLT path: the model, its configs, the mask helper, the dataset / collater that build
# the batch dict, the criterion / schedule of the training loop and the evaluators
MODULES = [
    "modelling/__init__.py", "modelling/configs.py", "modelling/models.py", "modelling/resnets3d.py",
    "modelling/datasets.py", "utils/__init__.py", "utils/model_utils.py", "utils/data_utils.py",
    "utils/train_inference_utils.py", "utils/evaluation.py",
]


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is usable afterwards."""
    stamp = REF_DIR / "MANIFEST.json"
    if not REFERENCE_SRC.exists():
        return stamp.exists()
    digest = hashlib.sha256()
    for rel in MODULES:
        digest.update(rel.encode())
        digest.update((REFERENCE_SRC / rel).read_bytes())
    tag = f"{sys.implementation.cache_tag}:{importlib.util.MAGIC_NUMBER.hex()}:{digest.hexdigest()}"
    if not force and stamp.exists() and json.loads(stamp.read_text()).get("tag") == tag:
        return True
    for rel in MODULES:
        dst = (REF_DIR / rel).with_suffix(".pyc")
        dst.parent.mkdir(parents=True, exist_ok=True)
        py_compile.compile(str(REFERENCE_SRC / rel), cfile=str(dst), dfile=f"<reference>/src/{rel}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    stamp.write_text(json.dumps({"tag": tag, "modules": MODULES, "source": str(REFERENCE_SRC),
                                 "note": "byte-compiled, unmodified; see oracle/build_ref.py"}, indent=1))
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print(f"oracle/_ref {'ready' if ok else 'unavailable (no /root/reference and no earlier build)'}: {REF_DIR}")
