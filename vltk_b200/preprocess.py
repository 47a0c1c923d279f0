"""Drop-in for the reference's legacy `Preprocess` (vltk/legacy/processing.py:76-150):
BGR u8 image -> shortest-edge bilinear resize (800/1333 rule, `int(x+0.5)` rounding) ->
(x-mean)/std -> zero-pad to the batch max, returning (ids, images, sizes, scales_yx).
Resize + normalise + pad run as ONE CUDA kernel per image (csrc/elementwise.cu); only file
decoding stays on the host, as in the reference (compat.py:573-579)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib
from .config import FRCNNConfig
from .synthetic import resized_hw


def _decode_bgr(path: str) -> np.ndarray:
    try:
        import cv2
        img = cv2.imread(path)
        if img is None:
            raise IOError(path)
        return img
    except ImportError:
        from PIL import Image
        return np.asarray(Image.open(path).convert("RGB"))[:, :, ::-1].copy()


class Preprocess:
    def __init__(self, cfg: FRCNNConfig = None, device: int = 0):
        self.cfg = cfg or FRCNNConfig()
        self.device = torch.device("cuda", int(device))
        self._lib = _lib.lib()
        self._mean = (C.c_float * 3)(*self.cfg.pixel_mean)
        self._std = (C.c_float * 3)(*self.cfg.pixel_std)

    def __call__(self, images, img_ids=None):
        if not isinstance(images, (list, tuple)):
            images = [images]
        if img_ids is None:
            img_ids = list(range(len(images)))
        raws: List[torch.Tensor] = []
        good_ids = []
        for img_id, img in zip(img_ids, images):
            if isinstance(img, str):
                img = torch.from_numpy(_decode_bgr(img))
            img = torch.as_tensor(img)
            if img.dtype != torch.uint8:
                raise TypeError("Preprocess expects decoded uint8 BGR images [h,w,3]")
            assert img.dim() == 3 and img.shape[2] == 3, img.shape
            raws.append(img.contiguous())
            good_ids.append(img_id)
        if not raws:
            return [], [], [], []
        raw_sizes = [(int(r.shape[0]), int(r.shape[1])) for r in raws]
        new_sizes = [resized_hw(h, w, self.cfg) for h, w in raw_sizes]
        hm = max(s[0] for s in new_sizes)
        wm = max(s[1] for s in new_sizes)
        n = len(raws)
        with torch.cuda.device(self.device):
            out = torch.empty((n, 3, hm, wm), dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            keep = []
            for i, (r, (rh, rw), (nh, nw)) in enumerate(zip(raws, raw_sizes, new_sizes)):
                d = r if r.is_cuda else r.to(self.device, non_blocking=True)
                keep.append(d)
                _lib.check(self._lib.vltk_frcnn_preprocess(
                    d.data_ptr(), rh, rw, nh, nw, self._mean, self._std, float(self.cfg.pad_value),
                    out.data_ptr(), i, hm, wm, stream), "vltk_frcnn_preprocess")
            torch.cuda.current_stream(self.device).synchronize()
        sizes = torch.tensor(new_sizes)
        scales_yx = torch.true_divide(torch.tensor(raw_sizes), sizes)
        return good_ids, out, sizes, scales_yx
