"""Test infrastructure, not product code: imports the unmodified reference from oracle/_ref (see build_ref.py).

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may use this module."""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF_DIR / "MANIFEST.json").exists() and (REF_DIR / "modelling" / "models.pyc").exists()


def load():
    """Returns (models, configs) = the reference's modelling.models / modelling.configs modules, or None when
    oracle/_ref has not been built (no /root/reference at build() time)."""
    if not available():
        return None
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    # modules the reference imports at module scope but never uses on the layout path
    for name in ("h5py", "ffmpeg"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    models = importlib.import_module("modelling.models")
    configs = importlib.import_module("modelling.configs")
    if not str(getattr(models, "__file__", "")).startswith(str(REF_DIR)):
        raise RuntimeError(f"modelling.models resolved to {models.__file__}, not to oracle/_ref")
    return models, configs


def load_data():
    """(datasets, data_utils) of the reference, for the dataset + collater leg."""
    if load() is None:
        return None
    return importlib.import_module("modelling.datasets"), importlib.import_module("utils.data_utils")
