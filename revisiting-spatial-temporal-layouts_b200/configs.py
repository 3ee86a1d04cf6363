"""Model configuration of the STLT path; mirrors the reference's kwargs-popping config objects
(reference src/modelling/configs.py:92-126: GeneralModelConfig + StltModelConfig) so callers can
pass either this class or the reference's own instance (attributes are read duck-typed)."""
from __future__ import annotations


class StltModelConfig:
    def __init__(self, **kwargs):
        # GeneralModelConfig (configs.py:92-99)
        self.num_classes = kwargs.pop("num_classes", None)
        assert self.num_classes, "num_classes must not be None!"
        self.hidden_size = kwargs.pop("hidden_size", 768)
        self.hidden_dropout_prob = kwargs.pop("hidden_dropout_prob", 0.1)
        self.layer_norm_eps = kwargs.pop("layer_norm_eps", 1e-12)
        self.num_attention_heads = kwargs.pop("num_attention_heads", 12)
        # StltModelConfig (configs.py:102-111)
        self.unique_categories = kwargs.pop("unique_categories", None)
        assert self.unique_categories, "unique_categories must not be None!"
        self.num_spatial_layers = kwargs.pop("num_spatial_layers", 4)
        self.num_temporal_layers = kwargs.pop("num_temporal_layers", 8)
        self.layout_num_frames = kwargs.pop("layout_num_frames", 256)
        self.load_backbone_path = kwargs.pop("load_backbone_path", None)
        self.freeze_backbone = kwargs.pop("freeze_backbone", False)

    def __repr__(self):
        return (
            f"- Unique categories: {self.unique_categories}\n"
            f"- Number of classes: {self.num_classes}\n"
            f"- Hidden size: {self.hidden_size}\n"
            f"- Hidden dropout probability: {self.hidden_dropout_prob}\n"
            f"- Layer normalization epsilon: {self.layer_norm_eps}\n"
            f"- Number of attention heads: {self.num_attention_heads}\n"
            f"- Number of spatial layers: {self.num_spatial_layers}\n"
            f"- Number of temporal layers: {self.num_temporal_layers}\n"
            f"- Max number of layout frames: {self.layout_num_frames}\n"
            f"- The backbone path is: {self.load_backbone_path}\n"
            f"- Freezing the backbone: {self.freeze_backbone}"
        )


class CacnfModelConfig(StltModelConfig):
    """Mirrors the reference MultimodalModelConfig (configs.py:150-175) + AppearanceModelConfig
    (:128-147) for the CACNF path on precomputed features; no ``resnet_model_path`` is needed because
    the 3D-ResNet trunk is outside this library."""

    def __init__(self, **kwargs):
        self.appearance_num_frames = kwargs.pop("appearance_num_frames", 32)
        self.num_appearance_layers = kwargs.pop("num_appearance_layers", 4)
        self.num_fusion_layers = kwargs.pop("num_fusion_layers", 4)
        self.feature_channels = kwargs.pop("feature_channels", 2048)
        kwargs.pop("resnet_model_path", None)
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
