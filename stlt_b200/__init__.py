"""Import alias for the package directory ``revisiting-spatial-temporal-layouts_b200/`` (whose
name is not a valid Python identifier). ``import stlt_b200`` resolves sub-modules from there."""
from pathlib import Path as _Path

_PKG_DIR = _Path(__file__).resolve().parent.parent / "revisiting-spatial-temporal-layouts_b200"
__path__.append(str(_PKG_DIR))

from .configs import ACTION_GENOME, SOMETHING_ELSE, CacnfModelConfig, StltModelConfig  # noqa: E402
from .module import Stlt, StltBackbone, models_factory  # noqa: E402
from .cacnf import Cacnf, Caf, Lcf  # noqa: E402

models_factory.update({"cacnf": Cacnf, "caf": Caf, "lcf": Lcf})
from .prepare import prepare_layout_batch  # noqa: E402
from .data import CharadesMapEvaluator, LayoutStore, TopKCounter  # noqa: E402
from .pipeline import HostPipeline  # noqa: E402
from .training import FusedTrainStep, linear_schedule_with_warmup  # noqa: E402

__all__ = ["Stlt", "StltBackbone", "StltModelConfig", "Cacnf", "Caf", "Lcf", "CacnfModelConfig", "models_factory", "prepare_layout_batch", "LayoutStore", "TopKCounter", "CharadesMapEvaluator", "HostPipeline", "FusedTrainStep",
           "linear_schedule_with_warmup",
           "SOMETHING_ELSE", "ACTION_GENOME"]
