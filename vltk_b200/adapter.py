"""Drop-in for the reference's extraction plugin `vltk.adapters.frcnn.FRCNN`
(vltk/adapters/frcnn.py:10-64): same hooks — `setup()`, `schema()`, `forward(model, entry)`
— and the same row dict {object_ids, attr_ids, box, features}, so `VisnExtraction.extract`
(vltk/abc/extraction.py:94-248) can call it unchanged.

Differences, all documented in SURVEY.md Appendix B.15: the model contract is followed
((h,w) image_shapes, legacy `Preprocess` arithmetic) instead of the live adapter's
self-declared-incorrect preset (adapters/frcnn.py:12); `setup()` needs a state_dict because
there is no hub access.
"""
from __future__ import annotations

import torch

from .config import FRCNNConfig
from .frcnn import FRCNN as FasterRCNN

# column-name constants of vltk/vars.py:46-57
FEATURES = "features"
BOX = "box"
IMG = "image"
SIZE = "size"
SCALE = "wh_scale"
RAWSIZE = "rawsize"


def rescale_box(boxes: torch.Tensor, wh_scale) -> torch.Tensor:
    """vltk/utils/adapters.py:205-216 (in place, x by wh_scale[0], y by wh_scale[1])."""
    boxes[:, 0] *= wh_scale[0]
    boxes[:, 1] *= wh_scale[1]
    boxes[:, 2] *= wh_scale[0]
    boxes[:, 3] *= wh_scale[1]
    return boxes


class FRCNN:
    """Extraction adapter.  Registered by lowercase class name, as the reference does
    (vltk/adapters/__init__.py:13-17)."""

    name = "frcnn"
    # what the reference's default_processor encodes (adapters/frcnn.py:13-23), expressed as the
    # Preprocess arguments actually honoured here
    default_processor = {"size": 800, "max_size": 1333, "mode": "bilinear", "pad_value": 0.0,
                         "mean": [102.9801, 115.9465, 122.7717], "std": [1.0, 1.0, 1.0]}
    _state_dict = None
    _config = None
    _mode = "exact_tc"

    @classmethod
    def configure(cls, state_dict, config: FRCNNConfig = None, mode: str = "exact_tc"):
        cls._state_dict, cls._config, cls._mode = state_dict, config or FRCNNConfig(), mode

    @staticmethod
    def setup():
        """-> (model, model_config) (adapters/frcnn.py:25-32)."""
        cls = FRCNN
        if cls._state_dict is None:
            raise RuntimeError("FRCNN.configure(state_dict, config) first: no hub access here")
        model = FasterRCNN.from_pretrained(state_dict=cls._state_dict, config=cls._config, mode=cls._mode)
        return model, cls._config

    @staticmethod
    def schema(max_detections=36, visual_dim=2048):
        """Arrow column types (adapters/frcnn.py:34-41; vltk/features.py:13-16,81-95)."""
        import pyarrow as pa
        return {
            "attr_ids": pa.list_(pa.float32()),
            "object_ids": pa.list_(pa.float32()),
            FEATURES: pa.list_(pa.list_(pa.float32(), visual_dim), max_detections),
            BOX: pa.list_(pa.list_(pa.float32())),
        }

    @staticmethod
    def forward(model, entry):
        """One image -> one row (adapters/frcnn.py:43-64)."""
        size = torch.as_tensor(entry[SIZE])
        scale_wh = torch.as_tensor(entry[SCALE], dtype=torch.float32)
        image = entry[IMG]
        model_out = model(
            images=image.unsqueeze(0),
            image_shapes=size.unsqueeze(0),
            padding="max_detections",
            pad_value=0.0,
            location="cpu",
        )
        boxes = torch.round(rescale_box(model_out["boxes"][0].clone(), 1 / scale_wh))
        return {
            "object_ids": [model_out["obj_ids"][0].tolist()],
            "attr_ids": [model_out["attr_ids"][0].tolist()],
            BOX: [boxes.tolist()],
            FEATURES: [model_out["roi_features"][0]],
        }
