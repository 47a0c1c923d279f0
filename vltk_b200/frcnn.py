"""Drop-in for `vltk.modeling.frcnn.FRCNN` (reference: vltk/modeling/frcnn.py:1743-2004).

Same constructor / `from_pretrained` / `forward(images, image_shapes, scales_yx=..., **kw)`
contract and the same mutable `roi_outputs.{nms_thresh,min_detections,max_detections}`
knobs (tests/frcnn_test.py:16-31), but everything below the call is the C-ABI library
(include/vltk_frcnn.h): hand-written sm_100a kernels, no torch ops, no CPU fallback.
torch is used only for device memory, streams and host<->device copies.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .config import FRCNNConfig


class ROIOutputs:
    """The knobs callers poke on `model.roi_outputs` (frcnn.py:1229-1240)."""

    def __init__(self, cfg: FRCNNConfig):
        nms = cfg.nms_thresh_test
        self.nms_thresh = list(nms) if isinstance(nms, (list, tuple)) else [nms]
        self.score_thresh = cfg.score_thresh_test  # accepted, never used — as in the reference
        self.min_detections = cfg.min_detections
        self.max_detections = cfg.max_detections


class FRCNN:
    def __init__(self, cfg: Optional[FRCNNConfig] = None, mode: str = "exact_tc", device: int = 0):
        self.config = cfg or FRCNNConfig()
        self.mode = mode
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self.roi_outputs = ROIOutputs(self.config)
        self.min_detections = self.config.min_detections
        self.max_detections = self.config.max_detections
        self.training = False
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        ccfg = _lib.make_config(self.config, mode)
        _lib.check(self._lib.vltk_frcnn_create(C.byref(ccfg), self.device_index, C.byref(self._h)),
                   "vltk_frcnn_create")
        self._finalized = False
        self._workspace = None
        self._pinned: Dict[tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------ lifecycle
    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.vltk_frcnn_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, *model_args, **kwargs):
        """Reference: frcnn.py:1757-1922.  There is no hub access here: a `state_dict` (the
        reference's 640-key layout) or a local checkpoint path must be
        given; `config` may be an FRCNNConfig."""
        config = kwargs.pop("config", None)
        if config is None and model_args:
            config = model_args[0]
        state_dict = kwargs.pop("state_dict", None)
        mode = kwargs.pop("mode", "exact_tc")
        device = kwargs.pop("device", 0)
        if state_dict is None:
            if pretrained_model_name_or_path is None:
                raise ValueError("from_pretrained needs state_dict= or a local checkpoint path")
            state_dict = torch.load(pretrained_model_name_or_path, map_location="cpu")
        model = cls(config if isinstance(config, FRCNNConfig) else None, mode=mode, device=device)
        model.load_state_dict(state_dict)
        return model

    def load_state_dict(self, state_dict, strict: bool = True):
        if self._finalized:
            raise RuntimeError("weights already loaded into this engine")
        for key, t in state_dict.items():
            # old-format BN names, as the reference renames them (frcnn.py:1862-1872)
            key = key.replace("gamma", "weight") if "gamma" in key else key
            key = key.replace("beta", "bias") if "beta" in key else key
            if key.endswith("num_batches_tracked"):
                continue
            a = np.ascontiguousarray(t.detach().cpu().float().numpy())
            _lib.check(self._lib.vltk_frcnn_load_tensor(self._h, key.encode(), a.ctypes.data, a.size),
                       f"load_tensor({key})")
        _lib.check(self._lib.vltk_frcnn_finalize(self._h), "vltk_frcnn_finalize")
        self._finalized = True
        return self

    # -------------------------------------------------------------------- forward
    def _ws(self, n, h, w, slot=0):
        """One scratch buffer per in-flight slot (forwards on different streams must not share one)."""
        need = int(self._lib.vltk_frcnn_workspace_bytes(self._h, n, h, w))
        if self._workspace is None:
            self._workspace = {}
        cur = self._workspace.get(slot)
        if cur is None or cur.numel() < need:
            self._workspace[slot] = None
            self._workspace[slot] = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace[slot]

    def _alloc_out(self, n, md, d):
        dev = self.device
        t = OrderedDict(
            boxes=torch.empty((n, md, 4), dtype=torch.float32, device=dev),
            normalized_boxes=torch.empty((n, md, 4), dtype=torch.float32, device=dev),
            obj_ids=torch.empty((n, md), dtype=torch.int64, device=dev),
            obj_probs=torch.empty((n, md), dtype=torch.float32, device=dev),
            attr_ids=torch.empty((n, md), dtype=torch.int64, device=dev),
            attr_probs=torch.empty((n, md), dtype=torch.float32, device=dev),
            roi_features=torch.empty((n, md, d), dtype=torch.float32, device=dev),
            preds_per_image=torch.empty((n,), dtype=torch.int32, device=dev),
            keep_idx=torch.empty((n, md), dtype=torch.int32, device=dev),
        )
        o = _lib.Out()
        for k, v in t.items():
            setattr(o, k, v.data_ptr())
        return t, o

    def run(self, images: torch.Tensor, sizes_hw: np.ndarray, scales_yx: Optional[np.ndarray],
            max_detections: int, min_detections: int, nms_thresh, pad_value: float = 0.0, slot: int = 0,
            ignorey: Optional[np.ndarray] = None):
        """Enqueues one forward on the current stream; returns the dense device tensors.  The call
        never synchronises, so forwards on different streams overlap if they use different `slot`s
        (each slot owns a workspace)."""
        if not self._finalized:
            raise RuntimeError("load_state_dict() has not been called")
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
        n, c, h, w = images.shape
        assert c == 3
        sizes_hw = np.ascontiguousarray(sizes_hw, dtype=np.int32).reshape(n, 2)
        sc_ptr = None
        if scales_yx is not None:
            scales_yx = np.ascontiguousarray(scales_yx, dtype=np.float32).reshape(n, 2)
            sc_ptr = scales_yx.ctypes.data
        ws = self._ws(n, h, w, slot)
        d = self.config.res2_out_channels * 8
        tensors, out = self._alloc_out(n, max_detections, d)
        if ignorey is not None:      # [N, J, 2] y-ranges (frcnn.py:328-366); only honoured together with scales_yx
            ignorey = np.ascontiguousarray(ignorey, dtype=np.float32)
            if ignorey.ndim != 3 or ignorey.shape[0] != n or ignorey.shape[2] != 2:
                raise ValueError(f"ignorey must be [N={n}, J, 2], got {ignorey.shape}")   # reference: assert ndim == 3
        knobs = _lib.make_knobs(nms_thresh, min_detections, max_detections, pad_value, ignorey)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.vltk_frcnn_forward(
            self._h, images.data_ptr(), sizes_hw.ctypes.data, sc_ptr, n, h, w, C.byref(knobs),
            C.byref(out), ws.data_ptr(), ws.numel(), stream), "vltk_frcnn_forward")
        tensors["_keepalive"] = (images, sizes_hw, scales_yx, ignorey)
        return tensors

    def forward(self, images, image_shapes, gt_boxes=None, proposals=None, scales_yx=None,
                ignorey=None, **kwargs):
        """kwargs (frcnn.py:1924-1929): max_detections, return_tensors in {"np","pt",None},
        padding in {None,"max_detections"}, pad_value, location in {"cuda","cpu"}.

        Differences from the live reference, on purpose: (i) the `max_detections` kwarg overrides
        `roi_outputs.max_detections` for this call, as in v1.0.0 and as the docstring of the reference promises — the
        live reference ignores it (its handling is commented out, frcnn.py:1975-1994); (ii) padding="max_detections"
        returns v1.0.0's dense layout, which the live reference also dropped.  Index-exact agreement with the
        reference needs mode="exact_tc" (or "fp32"); mode="bf16" is the fast mode (stage-level bf16 bounds only).

        padding=None returns the live reference's ragged lists of per-image tensors;
        padding="max_detections" returns the v1.0.0 dense layout [N,max_det,...] plus `sizes`
        and `normalized_boxes`.

        ignorey [N,J,2] (frcnn.py:328-366, only honoured together with scales_yx, like the reference): per image J
        y-ranges in raw-image coordinates; RPN proposals spanning a range are dropped and the others clipped to its
        nearer end before clipping/NMS.  The reference's branch only survives a batch of one image (it overwrites
        the shared level_ids); here every image of the batch gets the same per-image rule."""
        if self.training:
            raise NotImplementedError()
        if gt_boxes is not None or proposals is not None:
            raise NotImplementedError("gt_boxes / proposals are not part of the extraction path")
        padding = kwargs.get("padding", None)
        return_tensors = kwargs.get("return_tensors", None)
        pad_value = kwargs.get("pad_value", 0)
        location = kwargs.get("location", None) or "cpu"
        assert padding in (None, "max_detections"), padding
        assert return_tensors in (None, "np", "pt"), return_tensors
        ro = self.roi_outputs
        max_det = int(kwargs.get("max_detections", None) or ro.max_detections)
        min_det = int(ro.min_detections)   # not clamped: with min > max no threshold satisfies the window and the LAST one is kept, like do_nms (frcnn.py:1273-1278)

        with torch.cuda.device(self.device):
            x = torch.as_tensor(images)
            if not x.is_cuda:
                x = x.float().contiguous()
                x = (x if x.is_pinned() else x.pin_memory()).to(self.device, non_blocking=True)
            else:
                x = x.to(self.device).float().contiguous()
            sizes = np.asarray(torch.as_tensor(image_shapes).cpu().numpy(), dtype=np.int32)
            scales = None if scales_yx is None else \
                np.asarray(torch.as_tensor(scales_yx).cpu().numpy(), dtype=np.float32)
            ign = None if ignorey is None else np.asarray(torch.as_tensor(ignorey).cpu().numpy(), dtype=np.float32)
            t = self.run(x, sizes, scales, max_det, min_det, ro.nms_thresh, float(pad_value), ignorey=ign)
            t.pop("_keepalive")
            keep = t.pop("keep_idx")
            counts = t["preds_per_image"].cpu().to(torch.int64)  # the one sync; int64 like frcnn.py:1985

        def place(v):
            return v.cpu() if (location == "cpu" or return_tensors == "np") else v

        keys = ("obj_ids", "obj_probs", "attr_ids", "attr_probs", "boxes", "roi_features")
        if padding is None:
            out = OrderedDict()
            for k in keys[:5]:
                out[k] = [place(t[k][i, : int(c)]) for i, c in enumerate(counts)]
            out["preds_per_image"] = counts
            out["roi_features"] = [place(t["roi_features"][i, : int(c)]) for i, c in enumerate(counts)]
            out["keep_idx"] = [place(keep[i, : int(c)]).to(torch.int64) for i, c in enumerate(counts)]
            return out
        out = OrderedDict()
        for k in keys:
            out[k] = place(t[k])
        out["preds_per_image"] = counts
        out["sizes"] = torch.as_tensor(sizes.astype(np.int64))
        out["normalized_boxes"] = place(t["normalized_boxes"])
        out["keep_idx"] = place(keep)
        if return_tensors == "np":
            out = OrderedDict((k, v.numpy()) for k, v in out.items())
        return out

    __call__ = forward
    inference = forward

    def forward_stream(self, batches, max_detections=None, pad_value=0.0, depth=3, compute_streams=2):
        """Pipelined `forward(..., padding="max_detections", return_tensors="np")` over an iterable of host
        batches `(images, image_shapes, scales_yx)`: the host->device copy of batch i+1 and the device->host
        copy of batch i-1 run on their own streams while batch i computes (the forward itself never
        synchronises), and consecutive batches alternate between `compute_streams` CUDA streams, each with its
        own workspace, so one batch's few-CTA selection kernels (RPN top-k, NMS, detection tail: ~1 ms on 8 SMs)
        overlap the other batch's convolutions (+4.5 % measured).  Yields one dict of numpy arrays per batch,
        in order, bit-identical to `forward` on the same batch.  Every batch's images are copied from (pinned)
        host memory and every result is read back to the host, exactly like `forward`."""
        if not self._finalized:
            raise RuntimeError("load_state_dict() has not been called")
        ro = self.roi_outputs
        md = int(max_detections or ro.max_detections)
        mind = int(ro.min_detections)
        dev = self.device
        depth = max(int(depth), 1)
        keys = ("obj_ids", "obj_probs", "attr_ids", "attr_probs", "boxes", "roi_features", "preds_per_image",
                "normalized_boxes", "keep_idx")
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            cs = [main] + [torch.cuda.Stream(device=dev) for _ in range(max(int(compute_streams), 1) - 1)]
            for s_ in cs[1:]:                        # ordered after whatever the caller enqueued before this call
                s_.wait_stream(main)
            s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            slots = [dict(x=None, ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(), ev_out=torch.cuda.Event(),
                          host=None, dev_out=None, sizes=None, busy=False) for _ in range(depth)]

            def collect(sl):
                sl["ev_out"].synchronize()
                out = OrderedDict((k, sl["host"][k].numpy().copy()) for k in keys)   # staging buffers are reused
                out["preds_per_image"] = out["preds_per_image"].astype(np.int64)
                out["sizes"] = sl["sizes"].astype(np.int64)
                sl["busy"] = False
                return out

            pending = []
            try:
                for i, (images, image_shapes, scales_yx) in enumerate(batches):
                    sl = slots[i % depth]
                    if sl["busy"]:                       # its previous result has not been handed out yet
                        yield collect(pending.pop(0))
                    k = i % len(cs)                      # compute stream and workspace slot of this batch
                    compute = cs[k]
                    x = torch.as_tensor(images)
                    if x.is_cuda:
                        raise ValueError("forward_stream takes host batches; use forward() for device tensors")
                    x = x.float().contiguous()
                    if not x.is_pinned():
                        x = x.pin_memory()
                    if sl["x"] is None or sl["x"].shape != x.shape:
                        sl["x"] = torch.empty(x.shape, dtype=torch.float32, device=dev)
                    sizes = np.asarray(torch.as_tensor(image_shapes).cpu().numpy(), dtype=np.int32)
                    scales = None if scales_yx is None else np.asarray(torch.as_tensor(scales_yx).cpu().numpy(), dtype=np.float32)
                    with torch.cuda.stream(s_in):
                        s_in.wait_event(sl["ev_done"])   # the forward that last read this input buffer is finished
                        sl["x"].copy_(x, non_blocking=True)
                        sl["ev_in"].record(s_in)
                    compute.wait_event(sl["ev_in"])
                    compute.wait_event(sl["ev_out"])     # the D2H that last read this slot's outputs is finished
                    with torch.cuda.stream(compute):
                        t = self.run(sl["x"], sizes, scales, md, mind, ro.nms_thresh, float(pad_value), slot=k)
                    keep_alive = t.pop("_keepalive")
                    sl["ev_done"].record(compute)
                    if sl["host"] is None or sl["host"]["roi_features"].shape != t["roi_features"].shape:
                        sl["host"] = {k_: torch.empty(t[k_].shape, dtype=t[k_].dtype).pin_memory() for k_ in keys}
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(sl["ev_done"])
                        for k_ in keys:
                            sl["host"][k_].copy_(t[k_], non_blocking=True)
                        sl["ev_out"].record(s_out)
                    sl["dev_out"], sl["sizes"], sl["busy"], sl["_x_host"], sl["_ka"] = t, sizes, True, x, keep_alive
                    pending.append(sl)
                    if len(pending) >= depth:            # hand out the oldest result while newer batches run
                        yield collect(pending.pop(0))
                while pending:
                    yield collect(pending.pop(0))
            finally:
                for s_ in cs[1:]:                        # later work on the caller's stream stays ordered after ours
                    main.wait_stream(s_)

    def forward_jpeg_stream(self, batches, preprocess, group: int = 8, max_detections=None, pad_value=0.0,
                            depth: int = 3, compute_streams: int = 2, on_device: bool = False):
        """Raw-image front door of the extraction path: `batches` yields lists whose entries are JPEG byte strings
        and/or decoded BGR u8 [h,w,3] arrays/tensors (one list = one model batch).  The encoded entries of `group`
        batches at a time are decoded by ONE call of the GPU JPEG front end (one CTA per image: the more images
        per call, the better its few-SM kernels amortise), then each batch is resized/normalised/padded by the
        fused preprocess kernel, run, and read back asynchronously; consecutive batches alternate between
        `compute_streams` streams (own workspace each) like `forward_stream`; the group decode itself runs with both
        streams drained.  Yields one dict of numpy arrays per batch, in order (plus `scales_yx`).
        on_device=True skips the device->host copy and yields the dense DEVICE tensors instead (the single-file
        extraction packs them on the device and hands them to NCCL without a host round trip)."""
        if not self._finalized:
            raise RuntimeError("load_state_dict() has not been called")
        ro = self.roi_outputs
        md = int(max_detections or ro.max_detections)
        mind = int(ro.min_detections)
        dev = self.device
        depth = max(int(depth), 1)
        keys = ("obj_ids", "obj_probs", "attr_ids", "attr_probs", "boxes", "roi_features", "preds_per_image",
                "normalized_boxes", "keep_idx")
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            cs = [main] + [torch.cuda.Stream(device=dev) for _ in range(max(int(compute_streams), 1) - 1)]
            for s_ in cs[1:]:
                s_.wait_stream(main)
            s_out = torch.cuda.Stream(device=dev)
            slots = [dict(ev_done=torch.cuda.Event(), ev_out=torch.cuda.Event(), host=None, busy=False) for _ in range(depth)]
            ev_dec = torch.cuda.Event()
            pending = []

            def collect(sl):
                if on_device:                            # fresh tensors per forward (run() allocates its outputs): no copy needed
                    sl["ev_done"].synchronize()
                    out = OrderedDict((k, sl["dev_out"][k]) for k in keys)
                else:
                    sl["ev_out"].synchronize()
                    out = OrderedDict((k, sl["host"][k].numpy().copy()) for k in keys)
                    out["preds_per_image"] = out["preds_per_image"].astype(np.int64)
                out["sizes"] = sl["sizes"].astype(np.int64)
                out["scales_yx"] = sl["scales"]
                sl["busy"] = False
                return out

            it = iter(batches)
            i = 0
            try:
                while True:
                    grp = []
                    for _ in range(group):
                        b = next(it, None)
                        if b is None:
                            break
                        grp.append(list(b))
                    if not grp:
                        break
                    flat = [d for b in grp for d in b]
                    enc = [j for j, d in enumerate(flat) if isinstance(d, (bytes, bytearray, memoryview))]
                    imgs = list(flat)
                    if enc:                                          # one front-end call for the whole group (caller's stream)
                        # The Huffman kernel is 1-CTA-per-image and runs ~2 ms; the convolutions are persistent kernels
                        # with a static tile schedule and a full register file per SM, so a conv launch that finds 64
                        # SMs held by the decoder waits for them (measured: 485 -> 272 images/s when the two overlap).
                        # The decode therefore runs between groups, after both compute streams have drained.
                        for s_ in cs[1:]:
                            main.wait_stream(s_)
                        for j, t in zip(enc, preprocess._decode_jpegs([bytes(flat[j]) for j in enc])):
                            imgs[j] = t
                    ev_dec.record(main)
                    o = 0
                    for b in grp:
                        sl = slots[i % depth]
                        k = i % len(cs)
                        compute = cs[k]
                        i += 1
                        if sl["busy"]:
                            yield collect(pending.pop(0))
                        part = imgs[o:o + len(b)]
                        o += len(b)
                        compute.wait_event(ev_dec)               # this group's decoded images
                        compute.wait_event(sl["ev_out"])
                        with torch.cuda.stream(compute):
                            _, x, sizes_t, scales_t = preprocess(part, sync=False)
                            sizes = np.asarray(sizes_t.numpy(), dtype=np.int32)
                            scales = np.asarray(scales_t.numpy(), dtype=np.float32)
                            t = self.run(x, sizes, scales, md, mind, ro.nms_thresh, float(pad_value), slot=k)
                        sl["_ka"] = (t.pop("_keepalive"), part)  # decoded images stay alive until this slot is reused
                        sl["ev_done"].record(compute)
                        if not on_device:
                            if sl["host"] is None or sl["host"]["roi_features"].shape != t["roi_features"].shape:
                                sl["host"] = {k_: torch.empty(t[k_].shape, dtype=t[k_].dtype).pin_memory() for k_ in keys}
                            with torch.cuda.stream(s_out):
                                s_out.wait_event(sl["ev_done"])
                                for k_ in keys:
                                    sl["host"][k_].copy_(t[k_], non_blocking=True)
                                sl["ev_out"].record(s_out)
                        sl["dev_out"], sl["sizes"], sl["scales"], sl["busy"] = t, sizes, scales, True
                        pending.append(sl)
                        if len(pending) >= depth:
                            yield collect(pending.pop(0))
                while pending:
                    yield collect(pending.pop(0))
            finally:
                for s_ in cs[1:]:
                    main.wait_stream(s_)

    forward_raw_stream = forward_jpeg_stream

    # ------------------------------------------------------------------ test taps
    def debug_read(self, name: str, dtype=np.float32) -> np.ndarray:
        """Copies an intermediate of the last forward to the host (tests only)."""
        cap = 1 << 20
        while True:
            buf = np.empty(cap, dtype=np.float32)
            n = int(self._lib.vltk_frcnn_debug_read(self._h, name.encode(), buf.ctypes.data, cap))
            if n >= 0:
                return buf[:n].view(dtype).copy()
            msg = self._lib.vltk_frcnn_last_error().decode()
            if "capacity" in msg and cap < (1 << 33):
                cap *= 8
                continue
            raise _lib.LibraryError(msg)

    def run_part(self, part: int, x: torch.Tensor, blocks=(0, -1)) -> torch.Tensor:
        """One part of the model on a DEVICE fp32 tensor through this engine's layers and arithmetic mode (tests only;
        include/vltk_frcnn.h vltk_frcnn_run_part): part 0 stem (x = images NCHW) -> pooled NHWC; 2-4 res2-res4
        (NHWC -> NHWC; only blocks [blocks[0], blocks[1]) of the stage); 5 RPN head (res4 NHWC -> fp32 rows [n, h, w, ld], columns [0,60) deltas, [60,75) logits)."""
        x = x.to(self.device).float().contiguous()
        n = x.shape[0]
        hh, ww = (x.shape[2], x.shape[3]) if part == 0 else (x.shape[1], x.shape[2])
        cap = int(n) * hh * ww * 2048
        y = torch.empty(cap, dtype=torch.float32, device=self.device)
        dims = (C.c_int32 * 3)()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vltk_frcnn_run_part(self._h, int(part), int(blocks[0]), int(blocks[1]), x.data_ptr(), n, hh, ww, y.data_ptr(), cap,
                                                     C.cast(dims, C.c_void_p), torch.cuda.current_stream(self.device).cuda_stream),
                       "vltk_frcnn_run_part")
        oh, ow, ch = int(dims[0]), int(dims[1]), int(dims[2])
        return y[: n * oh * ow * ch].view(n, oh, ow, ch).clone()

    def launch_count(self) -> int:
        return int(self._lib.vltk_frcnn_launch_count(self._h))

    def profile(self, enable: bool):
        _lib.check(self._lib.vltk_frcnn_profile_enable(self._h, int(enable)), "profile_enable")

    def profile_read(self, want_csv: bool = False):
        """-> ({'tcgen05': (ms, flops, launches), 'simt': (...)}, csv or None); clears the log."""
        agg = (C.c_double * 6)()
        buf = C.create_string_buffer(1 << 22) if want_csv else None
        _lib.check(self._lib.vltk_frcnn_profile_read(self._h, agg, buf, (1 << 22) if want_csv else 0),
                   "profile_read")
        d = {"tcgen05": tuple(agg[0:3]), "simt": tuple(agg[3:6])}
        return d, (buf.value.decode() if want_csv else None)
