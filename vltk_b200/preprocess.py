"""Drop-in for the reference's legacy `Preprocess` (vltk/legacy/processing.py:76-150):
BGR u8 image -> shortest-edge bilinear resize (800/1333 rule, `int(x+0.5)` rounding) ->
(x-mean)/std -> zero-pad to the batch max, returning (ids, images, sizes, scales_yx).
Resize + normalise + pad run as ONE CUDA kernel per image (csrc/elementwise.cu).  JPEG files / byte strings are
decoded by the GPU front end (vltk_b200/jpeg.py: the decode inside the reference's `cv2.imread`, compat.py:573-579,
bit-exact); other formats are read on the host with cv2 exactly as the reference does."""
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


def _is_jpeg(head: bytes) -> bool:
    return len(head) >= 3 and head[0] == 0xFF and head[1] == 0xD8 and head[2] == 0xFF


class Preprocess:
    """images: list of decoded BGR u8 tensors/arrays [h,w,3] (host or device), file paths, or encoded JPEG bytes.

    host_decode_unsupported: JPEG files outside the GPU front end's coverage (CMYK, 12-bit, ...) raise
    `jpeg.UnsupportedJpeg` by default; set True to read those files with cv2 on the host like the reference."""

    def __init__(self, cfg: FRCNNConfig = None, device: int = 0, host_decode_unsupported: bool = False):
        self.cfg = cfg or FRCNNConfig()
        self.device = torch.device("cuda", int(device))
        self._lib = _lib.lib()
        self._mean = (C.c_float * 3)(*self.cfg.pixel_mean)
        self._std = (C.c_float * 3)(*self.cfg.pixel_std)
        self.host_decode_unsupported = host_decode_unsupported
        self._jpeg = None

    def _decode_jpegs(self, datas):
        from . import jpeg
        if self._jpeg is None:
            self._jpeg = jpeg.JpegDecoder(self.device)
        try:
            return self._jpeg.decode(datas)
        except jpeg.UnsupportedJpeg:
            if len(datas) == 1:
                raise
            out = []                                     # isolate the offending file(s)
            for d in datas:
                out.extend(self._decode_jpegs([d]))
            return out

    def __call__(self, images, img_ids=None, sync: bool = True):
        if not isinstance(images, (list, tuple)):
            images = [images]
        if img_ids is None:
            img_ids = list(range(len(images)))
        raws: List[torch.Tensor] = []
        good_ids = []
        images = list(images)
        # encoded JPEGs (bytes, or files that start with an SOI marker) -> ONE batched GPU decode
        jp_idx, jp_data = [], []
        for i, img in enumerate(images):
            data = None
            if isinstance(img, (bytes, bytearray, memoryview)):
                data = bytes(img)
            elif isinstance(img, str):
                with open(img, "rb") as f:
                    head = f.read(3)
                    if _is_jpeg(head):
                        data = head + f.read()
            if data is not None:
                jp_idx.append(i)
                jp_data.append(data)
        if jp_data:
            from . import jpeg
            try:
                decoded = self._decode_jpegs(jp_data)
            except jpeg.UnsupportedJpeg:
                if not self.host_decode_unsupported:
                    raise
                decoded = []
                for i, d in zip(jp_idx, jp_data):
                    try:
                        decoded.extend(self._decode_jpegs([d]))
                    except jpeg.UnsupportedJpeg:
                        import cv2
                        arr = cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_COLOR)
                        if arr is None:
                            raise
                        decoded.append(torch.from_numpy(arr))
            for i, t in zip(jp_idx, decoded):
                images[i] = t
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
            if sync:      # sync=False: stream-ordered, the caller keeps `images` alive until the stream has run
                torch.cuda.current_stream(self.device).synchronize()
            else:
                self._keep = keep
        sizes = torch.tensor(new_sizes)
        scales_yx = torch.true_divide(torch.tensor(raw_sizes), sizes)
        return good_ids, out, sizes, scales_yx
