"""JPEG front end: `cv2.imread` of the reference's image-loading call site
(vltk/compat.py:573-579 `img_tensorize`, used by legacy/processing.py:119-129) split in two — the serial
Huffman entropy decoder on host threads (C++, vltk_b200/csrc/jpeg_host.cpp, GIL released), and dequantisation +
inverse DCT + chroma upsampling + colour conversion on the GPU (vltk_b200/csrc/jpeg.cu), bit-identical to
libjpeg-turbo's default pipeline.  The decoded BGR u8 image never exists in host memory: coefficients go up
(int16, about the size of the u8 image), the existing fused resize/normalise/pad kernel consumes the device image.

Files the front end does not cover (progressive, 12-bit, CMYK, exotic sampling) raise `UnsupportedJpeg`: the
caller decides (there is no silent CPU decode behind this API).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib


JpegInfo = _lib.JpegInfo


class UnsupportedJpeg(ValueError):
    """A valid JPEG outside the front end's coverage (progressive, 12-bit, CMYK, unusual sampling)."""


def _check(rc: int, what: str):
    if rc == 0:
        return
    msg = _lib.lib().vltk_frcnn_last_error().decode()
    if rc == -3:
        raise UnsupportedJpeg(msg)
    raise _lib.LibraryError(f"{what} failed ({rc}): {msg}")


def parse(data: bytes) -> JpegInfo:
    L = _lib.lib()
    info = JpegInfo()
    _check(L.vltk_jpeg_parse(data, len(data), C.byref(info)), "vltk_jpeg_parse")
    return info


def coefficients(data: bytes):
    """Host-only: (info, int16 numpy array of info.coef_count quantised coefficients)."""
    info = parse(data)
    out = np.empty(int(info.coef_count), np.int16)
    _check(_lib.lib().vltk_jpeg_decode_coefficients(data, len(data), out.ctypes.data, out.size), "vltk_jpeg_decode_coefficients")
    return info, out


def apply_orientation(img: torch.Tensor, orientation: int) -> torch.Tensor:
    """EXIF orientation as cv2.imread applies it (IMREAD_COLOR without IMREAD_IGNORE_ORIENTATION)."""
    if orientation in (0, 1) or orientation > 8:
        return img
    if orientation == 2:
        return img.flip(1)
    if orientation == 3:
        return img.flip(0).flip(1)
    if orientation == 4:
        return img.flip(0)
    if orientation == 5:
        return img.transpose(0, 1)
    if orientation == 6:
        return img.transpose(0, 1).flip(1)
    if orientation == 7:
        return img.transpose(0, 1).flip(0).flip(1)
    return img.transpose(0, 1).flip(0)     # 8


class JpegDecoder:
    """Decodes batches of JPEG byte strings to BGR u8 [h, w, 3] tensors on `device`.

    decode(list_of_bytes) -> list of device tensors (one per image), stream-ordered on the current stream.
    Host work (marker parsing + Huffman decoding of every image of the batch) runs on `threads` C++ threads."""

    def __init__(self, device=None, threads: Optional[int] = None, honor_orientation: bool = True):
        self._L = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.LibraryError("JpegDecoder needs a CUDA device: the IDCT/colour stage has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.threads = int(threads or min(16, os.cpu_count() or 1))
        self.honor_orientation = honor_orientation
        self._pinned = None        # staging for the coefficients of one batch
        self._ev = None            # the H2D copy that last read the staging buffer

    def decode(self, datas: Sequence[bytes]) -> List[torch.Tensor]:
        n = len(datas)
        if n == 0:
            return []
        infos = [parse(d) for d in datas]
        # one pinned staging buffer + one device buffer for the whole batch; every image 16-byte aligned
        offs, tot = [], 0
        for inf in infos:
            offs.append(tot)
            tot += (int(inf.coef_count) + 7) // 8 * 8
        if self._ev is not None:
            self._ev.synchronize()                      # the previous batch's H2D has drained the staging buffer
        if self._pinned is None or self._pinned.numel() < tot:
            self._pinned = torch.empty(max(tot, 1), dtype=torch.int16).pin_memory()
        base = self._pinned.data_ptr()
        arr_d = (C.c_char_p * n)(*datas)
        arr_l = (C.c_size_t * n)(*[len(d) for d in datas])
        arr_c = (C.c_void_p * n)(*[base + 2 * o for o in offs])
        arr_cap = (C.c_int64 * n)(*[int(inf.coef_count) for inf in infos])
        status = (C.c_int * n)()
        _check(self._L.vltk_jpeg_decode_coefficients_batch(n, arr_d, arr_l, arr_c, arr_cap, self.threads, status),
               "vltk_jpeg_decode_coefficients_batch")
        out = []
        with torch.cuda.device(self.device):
            dev = self._pinned[:tot].to(self.device, non_blocking=True)
            self._ev = torch.cuda.Event()
            self._ev.record()
            st = torch.cuda.current_stream(self.device).cuda_stream
            planes = torch.empty(max(int(inf.plane_bytes) for inf in infos), dtype=torch.uint8, device=self.device)
            for inf, o in zip(infos, offs):
                img = torch.empty((inf.height, inf.width, 3), dtype=torch.uint8, device=self.device)
                _check(self._L.vltk_jpeg_reconstruct(dev.data_ptr() + 2 * o, C.byref(inf), planes.data_ptr(), img.data_ptr(), st),
                       "vltk_jpeg_reconstruct")
                if self.honor_orientation:
                    img = apply_orientation(img, inf.orientation)
                out.append(img)
        return out

    def decode_files(self, paths: Sequence[Union[str, os.PathLike]]) -> List[torch.Tensor]:
        datas = []
        for p in paths:
            with open(p, "rb") as f:
                datas.append(f.read())
        return self.decode(datas)
