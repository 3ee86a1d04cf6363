"""Model configuration of the STLT path.

Schema = the fields of the reference's GeneralModelConfig + StltModelConfig
(reference src/modelling/configs.py:92-111) that the forward path reads, with the same names and defaults, so
callers can pass either this class or the reference's own instance (attributes are read duck-typed). The
classes are table-driven: keyword arguments not in the table are ignored, as the reference's kwargs.pop
objects do; the two required fields raise when missing."""
from __future__ import annotations

_REQUIRED = object()


class _TableConfig:
    """Keyword-constructed config whose fields and defaults come from the class attribute ``FIELDS``."""
    FIELDS: dict = {}

    def __init__(self, **kwargs):
        for name, default in self.FIELDS.items():
            value = kwargs.get(name, None if default is _REQUIRED else default)
            if default is _REQUIRED and not value:
                raise AssertionError(f"{name} must not be None!")
            setattr(self, name, value)

    def to_dict(self) -> dict:
        return {name: getattr(self, name) for name in self.FIELDS}

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in self.to_dict().items())})"

    def __eq__(self, other) -> bool:
        return isinstance(other, _TableConfig) and self.to_dict() == other.to_dict()


class StltModelConfig(_TableConfig):
    FIELDS = {
        "num_classes": _REQUIRED, "unique_categories": _REQUIRED,
        "hidden_size": 768, "num_attention_heads": 12, "hidden_dropout_prob": 0.1, "layer_norm_eps": 1e-12,
        "num_spatial_layers": 4, "num_temporal_layers": 8,
        "layout_num_frames": 256,  # rows of the position table; the data side samples 16
        "load_backbone_path": None, "freeze_backbone": False,
    }


class CacnfModelConfig(StltModelConfig):
    """Fields of the reference MultimodalModelConfig (configs.py:150-175) + AppearanceModelConfig (:128-147) that
    the CACNF path on precomputed features reads; ``resnet_model_path`` is accepted and ignored because the
    3D-ResNet trunk is outside this library."""
    FIELDS = {**StltModelConfig.FIELDS, "appearance_num_frames": 32, "num_appearance_layers": 4,
              "num_fusion_layers": 4, "feature_channels": 2048}

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.stlt_config = self
        self.appearance_config = self


# Dataset-level constants of the two supported layouts (reference src/modelling/configs.py:30-88).
SOMETHING_ELSE = {
    "unique_categories": 4, "num_classes": 174, "cls_id": 3, "object_ids": (1, 2),
    "frame_types": {"pad": 0, "start": 1, "regular": 2, "empty": 3, "extract": 4}, "scores": False,
}
ACTION_GENOME = {
    "unique_categories": 38, "num_classes": 157, "cls_id": 1, "object_ids": tuple(range(2, 38)),
    "frame_types": {"pad": 0, "regular": 1, "extract": 2, "empty": 3}, "scores": True,
}
