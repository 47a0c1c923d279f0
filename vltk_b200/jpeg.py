"""JPEG front end: `cv2.imread` of the reference's image-loading call site
(vltk/compat.py:573-579 `img_tensorize`, used by legacy/processing.py:119-129) split in two — the serial
Huffman entropy decoder on host threads (C++, vltk_b200/csrc/jpeg_host.cpp, GIL released), and dequantisation +
inverse DCT + chroma upsampling + colour conversion on the GPU (vltk_b200/csrc/jpeg.cu), bit-identical to
libjpeg-turbo's default pipeline.  The decoded BGR u8 image never exists in host memory: coefficients go up
(int16, about the size of the u8 image), the existing fused resize/normalise/pad kernel consumes the device image.

Progressive files take the host entropy decoder (their scans are serial) and the same GPU stages.  Files the front end
does not cover (12-bit, CMYK, arithmetic coding, exotic sampling) raise `UnsupportedJpeg`: the
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
    """A valid JPEG outside the front end's coverage (12-bit, CMYK, arithmetic coding, unusual sampling)."""


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
    """Decodes batches of JPEG byte strings to BGR u8 [h, w, 3] tensors on `device`, stream-ordered on the current
    stream; `decode(list_of_bytes)` returns one device tensor per image.

    entropy="gpu" (default): the host only parses markers and strips the byte stuffing (~0.1 ms per image); the
    Huffman decoding runs on the device, one CTA per image (self-synchronising subsequences; streams with restart
    intervals need no synchronisation: one thread per interval).
    entropy="host": every image is Huffman-decoded on `threads` C++ host threads (GIL released)."""

    def __init__(self, device=None, threads: Optional[int] = None, honor_orientation: bool = True, entropy: str = "gpu"):
        self._L = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.LibraryError("JpegDecoder needs a CUDA device: the decode stages have no CPU fallback")
        assert entropy in ("gpu", "host")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.threads = int(threads or min(16, os.cpu_count() or 1))
        self.honor_orientation = honor_orientation
        self.entropy = entropy
        self._pinned = None        # staging for one batch (coefficients or blob)
        self._ev = None            # the H2D copy that last read the staging buffer
        self.last_iterations = None

    def _staging(self, nbytes: int) -> torch.Tensor:
        if self._ev is not None:
            self._ev.synchronize()                      # the previous batch's H2D has drained the staging buffer
        if self._pinned is None or self._pinned.numel() < nbytes:
            self._pinned = torch.empty(max(nbytes, 16), dtype=torch.uint8).pin_memory()
        return self._pinned

    def _coefficients_host(self, datas, infos, offs, tot):
        """Huffman-decodes on host threads into pinned memory, one async H2D -> device int16 buffer."""
        n = len(datas)
        pin = self._staging(tot * 2)
        base = pin.data_ptr()
        arr_d = (C.c_char_p * n)(*datas)
        arr_l = (C.c_size_t * n)(*[len(d) for d in datas])
        arr_c = (C.c_void_p * n)(*[base + 2 * o for o in offs])
        arr_cap = (C.c_int64 * n)(*[int(inf.coef_count) for inf in infos])
        status = (C.c_int * n)()
        _check(self._L.vltk_jpeg_decode_coefficients_batch(n, arr_d, arr_l, arr_c, arr_cap, self.threads, status),
               "vltk_jpeg_decode_coefficients_batch")
        dev = pin[: tot * 2].to(self.device, non_blocking=True)
        self._ev = torch.cuda.Event()
        self._ev.record()
        return dev.view(torch.int16)

    def _coefficients_gpu(self, datas):
        n = len(datas)
        arr_d = (C.c_char_p * n)(*datas)
        arr_l = (C.c_size_t * n)(*[len(d) for d in datas])
        cap = int(self._L.vltk_jpeg_gpu_blob_bound(n, arr_l))
        pin = self._staging(cap)
        infos = (JpegInfo * n)()
        used, coef_total = C.c_size_t(0), C.c_int64(0)
        offs = (C.c_int64 * n)()
        on_gpu = (C.c_int * n)()
        _check(self._L.vltk_jpeg_gpu_prepare_batch(n, arr_d, arr_l, infos, pin.data_ptr(), cap, C.byref(used), offs,
                                                   C.byref(coef_total), on_gpu), "vltk_jpeg_gpu_prepare_batch")
        blob = pin[: used.value].to(self.device, non_blocking=True)
        self._ev = torch.cuda.Event()
        self._ev.record()
        tot = int(coef_total.value)
        coef = torch.empty(max(tot, 8), dtype=torch.int16, device=self.device)
        iters = torch.zeros(n, dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _check(self._L.vltk_jpeg_gpu_entropy_decode(n, blob.data_ptr(), coef.data_ptr(), tot, iters.data_ptr(), st),
               "vltk_jpeg_gpu_entropy_decode")
        self.last_iterations = iters
        for i in range(n):                               # restart-interval streams: host decoder, copied in place
            if not on_gpu[i]:
                _, co = coefficients(datas[i])
                coef[int(offs[i]): int(offs[i]) + co.size].copy_(torch.from_numpy(co), non_blocking=False)
        return [infos[i] for i in range(n)], [int(o) for o in offs], coef, blob

    def coefficients(self, datas: Sequence[bytes]):
        """(infos, offsets, device int16 coefficient buffer) — the stage between entropy decoding and the IDCT."""
        with torch.cuda.device(self.device):
            if self.entropy == "gpu":
                infos, offs, coef, _ = self._coefficients_gpu(datas)
                return infos, offs, coef
            infos = [parse(d) for d in datas]
            offs, tot = [], 0
            for inf in infos:
                offs.append(tot)
                tot += (int(inf.coef_count) + 7) // 8 * 8
            return infos, offs, self._coefficients_host(datas, infos, offs, tot)

    def decode(self, datas: Sequence[bytes]) -> List[torch.Tensor]:
        if len(datas) == 0:
            return []
        out = []
        with torch.cuda.device(self.device):
            infos, offs, dev = self.coefficients(datas)
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
