"""Test infrastructure, not product code: builds oracle/_ref/ — the UNMODIFIED reference, byte-compiled.

The reference (gorjanradevski/revisiting-spatial-temporal-layouts) is pure Python, so "compiling it from the
sources where they lie" means CPython byte-compilation: every module of /root/reference/src that the STLT path
touches is compiled with the builtin compile() straight from /root/reference and the code objects are marshalled
into ONE binary, ``oracle/_ref/stlt_reference.bin`` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like
the built ``.so``; the GPU box runs the same image, hence the same bytecode magic — which the loader verifies). No
reference source file is copied into the repo. (Loose ``.pyc`` files do not survive the snapshot to the GPU box.)

``oracle/_ref`` is what ``bench.py --impl reference`` times (``cpu_baseline.kind = "reference"``) and what
``tests/test_host.py`` pins the oracle restatement against on machines without /root/reference.

    python oracle/build_ref.py            # (re)build if /root/reference is present; no-op otherwise
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import marshal
import pickle
import sys
import warnings
from pathlib import Path

REFERENCE_SRC = Path("/root/reference/src")
REF_DIR = Path(__file__).resolve().parent / "_ref"
BLOB = REF_DIR / "stlt_reference.bin"
# modules on (or feeding) the STLT path: the model, its configs, the mask helper, the dataset / collater that build
# the batch dict, the criterion / schedule of the training loop and the evaluators
MODULES = [
    "modelling/__init__.py", "modelling/configs.py", "modelling/models.py", "modelling/resnets3d.py",
    "modelling/datasets.py", "utils/__init__.py", "utils/model_utils.py", "utils/data_utils.py",
    "utils/train_inference_utils.py", "utils/evaluation.py",
]


def _tag(digest: str) -> str:
    return f"{sys.implementation.cache_tag}:{importlib.util.MAGIC_NUMBER.hex()}:{digest}"


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is usable afterwards."""
    stamp = REF_DIR / "MANIFEST.json"
    if not REFERENCE_SRC.exists():
        return stamp.exists() and BLOB.exists()
    digest = hashlib.sha256()
    for rel in MODULES:
        digest.update(rel.encode())
        digest.update((REFERENCE_SRC / rel).read_bytes())
    tag = _tag(digest.hexdigest())
    if not force and stamp.exists() and BLOB.exists() and json.loads(stamp.read_text()).get("tag") == tag:
        return True
    REF_DIR.mkdir(parents=True, exist_ok=True)
    code = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)  # the reference has two invalid escape sequences in regexes
        for rel in MODULES:
            name = rel[:-3].replace("/", ".")
            is_pkg = name.endswith(".__init__")
            if is_pkg:
                name = name[: -len(".__init__")]
            src = (REFERENCE_SRC / rel).read_text()
            code[name] = (is_pkg, marshal.dumps(compile(src, f"<reference>/src/{rel}", "exec", dont_inherit=True)))
    BLOB.write_bytes(pickle.dumps({"magic": importlib.util.MAGIC_NUMBER, "modules": code}))
    for old in REF_DIR.rglob("*.pyc"):  # layout of an earlier recipe
        old.unlink()
    stamp.write_text(json.dumps({"tag": tag, "modules": MODULES, "source": str(REFERENCE_SRC), "blob": BLOB.name,
                                 "note": "marshalled code objects of the unmodified modules; see oracle/build_ref.py"},
                                indent=1))
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print(f"oracle/_ref {'ready' if ok else 'unavailable (no /root/reference and no earlier build)'}: {REF_DIR}")
