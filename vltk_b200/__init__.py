"""B200-native drop-in for vltk's Faster R-CNN region-feature extraction path."""
from .config import FRCNNConfig  # noqa: F401

__all__ = ["FRCNNConfig", "FRCNN", "Preprocess"]


def __getattr__(name):  # lazy: importing the package must not need torch/CUDA
    if name == "FRCNN":
        from .frcnn import FRCNN
        return FRCNN
    if name == "Preprocess":
        from .preprocess import Preprocess
        return Preprocess
    raise AttributeError(name)
