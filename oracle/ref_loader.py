"""Test infrastructure, not product code: imports the unmodified reference from oracle/_ref (see build_ref.py).

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may use this module."""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import marshal
import pickle
import sys
import types
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"
BLOB = REF_DIR / "stlt_reference.bin"
_TOP = ("modelling", "utils")


class _BlobFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serves the reference's `modelling.*` / `utils.*` modules from the marshalled code objects."""

    def __init__(self, modules):
        self.modules = modules

    def find_spec(self, fullname, path=None, target=None):
        if fullname not in self.modules:
            return None
        is_pkg, _ = self.modules[fullname]
        return importlib.util.spec_from_loader(fullname, self, origin=f"{BLOB}::{fullname}", is_package=is_pkg)

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        _, blob = self.modules[module.__name__]
        module.__file__ = f"{BLOB}::{module.__name__}"
        exec(marshal.loads(blob), module.__dict__)


_finder = None


def available() -> bool:
    return (REF_DIR / "MANIFEST.json").exists() and BLOB.exists()


def load():
    """Returns (models, configs) = the reference's modelling.models / modelling.configs modules, or None when
    oracle/_ref has not been built (no /root/reference at build() time)."""
    global _finder
    if not available():
        return None
    if _finder is None:
        data = pickle.loads(BLOB.read_bytes())
        if data["magic"] != importlib.util.MAGIC_NUMBER:
            raise RuntimeError("oracle/_ref was byte-compiled by another Python version: rebuild it (oracle/build_ref.py)")
        for name in _TOP:
            if name in sys.modules and not str(getattr(sys.modules[name], "__file__", "")).startswith(str(BLOB)):
                raise RuntimeError(f"a foreign top-level module '{name}' is already imported")
        _finder = _BlobFinder(data["modules"])
        sys.meta_path.insert(0, _finder)
    # modules the reference imports at module scope but never uses on the layout path
    for name in ("h5py", "ffmpeg"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    models = importlib.import_module("modelling.models")
    configs = importlib.import_module("modelling.configs")
    if not str(getattr(models, "__file__", "")).startswith(str(BLOB)):
        raise RuntimeError(f"modelling.models resolved to {models.__file__}, not to oracle/_ref")
    return models, configs


def load_data():
    """(datasets, data_utils) of the reference, for the dataset + collater leg."""
    if load() is None:
        return None
    return importlib.import_module("modelling.datasets"), importlib.import_module("utils.data_utils")
