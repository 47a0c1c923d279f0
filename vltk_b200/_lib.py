"""ctypes binding of libvltk_frcnn.so (include/vltk_frcnn.h).

The library is built in-tree by `vltk_b200/csrc/build.sh` (see __graft_entry__.build).
There is no CPU or PyTorch fallback: if the library is missing, or no sm_100 device is
present when an engine is created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VLTK_LIB selects another in-tree build of the same sources (the -DVLTK_TC_TRACE diagnosis build); never a fallback
LIB_PATH = os.path.join(_HERE, os.environ.get("VLTK_LIB", "libvltk_frcnn.so"))

MODE_FP32 = 0
MODE_BF16 = 1
MODE_EXACT_TC = 2
MODES = {"fp32": MODE_FP32, "bf16": MODE_BF16, "exact_tc": MODE_EXACT_TC}


class Config(C.Structure):
    _fields_ = [
        ("stem_out_channels", C.c_int), ("res2_out_channels", C.c_int), ("blocks", C.c_int * 3),
        ("res5_blocks", C.c_int), ("num_anchors", C.c_int), ("anchor_stride", C.c_int),
        ("rpn_hidden", C.c_int), ("rpn_nms_thresh", C.c_float), ("rpn_pre_nms_topk", C.c_int),
        ("rpn_post_nms_topk", C.c_int), ("rpn_min_size", C.c_float),
        ("rpn_bbox_weights", C.c_float * 4), ("pooler_resolution", C.c_int),
        ("num_classes", C.c_int), ("num_attrs", C.c_int), ("roi_bbox_weights", C.c_float * 4),
        ("mode", C.c_int),
    ]


class Knobs(C.Structure):
    _fields_ = [("nms_thresh", C.c_float * 4), ("n_nms_thresh", C.c_int),
                ("min_detections", C.c_int), ("max_detections", C.c_int), ("pad_value", C.c_float),
                ("ignorey", C.c_void_p), ("n_ignorey", C.c_int)]


class Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "boxes", "normalized_boxes", "obj_ids", "obj_probs", "attr_ids", "attr_probs",
        "roi_features", "preds_per_image", "keep_idx")]


class JpegInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("ncomp", C.c_int),
                ("comp_id", C.c_int * 3), ("hs", C.c_int * 3), ("vs", C.c_int * 3),
                ("hmax", C.c_int), ("vmax", C.c_int), ("mcus_x", C.c_int), ("mcus_y", C.c_int),
                ("blocks_w", C.c_int * 3), ("blocks_h", C.c_int * 3), ("comp_w", C.c_int * 3), ("comp_h", C.c_int * 3),
                ("coef_offset", C.c_int64 * 3), ("coef_count", C.c_int64),
                ("plane_offset", C.c_int64 * 3), ("plane_bytes", C.c_int64),
                ("qt", (C.c_uint16 * 64) * 3),
                ("restart_interval", C.c_int), ("orientation", C.c_int), ("progressive", C.c_int),
                ("color_transform", C.c_int)]


class LibraryError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/vltk_frcnn.h declares
SYMBOLS = {
    "vltk_frcnn_last_error": (C.c_char_p, []),
    "vltk_frcnn_version": (C.c_char_p, []),
    "vltk_frcnn_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]),
    "vltk_frcnn_destroy": (None, [C.c_void_p]),
    "vltk_frcnn_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "vltk_frcnn_finalize": (C.c_int, [C.c_void_p]),
    "vltk_frcnn_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vltk_frcnn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.POINTER(Knobs), C.POINTER(Out),
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "vltk_frcnn_preprocess": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vltk_conv2d_nhwc": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 13 + [C.c_void_p]),
    "vltk_conv2d_dual_nhwc": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 10 + [C.c_void_p]),
    "vltk_conv_tc_set_cta_pairs": (C.c_int, [C.c_int, C.c_int]),
    "vltk_conv_tcx_set_cta_pairs": (C.c_int, [C.c_int]),
    "vltk_conv_tc_set_trace": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vltk_conv2d_meanpool_nhwc": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 11 + [C.c_void_p]),
    "vltk_conv2d_meanpool_exact_nhwc": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 7 + [C.c_void_p]),
    "vltk_linear_tc3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p]),
    "vltk_rpn_proposals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 7
                           + [C.c_float, C.c_float, C.POINTER(C.c_float), C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "vltk_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p,
                           C.c_void_p, C.c_void_p]),
    "vltk_roi_pool_nchw": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "vltk_roi_outputs": (C.c_int, [C.c_void_p] * 8 + [C.c_int] * 5
                         + [C.POINTER(C.c_float), C.POINTER(Knobs), C.POINTER(Out), C.c_void_p]),
    "vltk_frcnn_run_part": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p]),
    "vltk_frcnn_debug_read": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "vltk_frcnn_launch_count": (C.c_int64, [C.c_void_p]),
    "vltk_frcnn_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "vltk_gather_rows_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p]),
    "vltk_jpeg_parse": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(JpegInfo)]),
    "vltk_jpeg_decode_coefficients": (C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int64]),
    "vltk_jpeg_decode_coefficients_batch": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                                      C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
                                                      C.POINTER(C.c_int)]),
    "vltk_jpeg_gpu_blob_bound": (C.c_size_t, [C.c_int, C.POINTER(C.c_size_t)]),
    "vltk_jpeg_gpu_prepare_batch": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.POINTER(JpegInfo),
                                              C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int64),
                                              C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "vltk_jpeg_gpu_entropy_decode": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vltk_jpeg_reconstruct": (C.c_int, [C.c_void_p, C.POINTER(JpegInfo), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vltk_frcnn_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_char_p, C.c_size_t]),
}


def lib():
    """Loads the shared library (once) and declares every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryError(
            f"{LIB_PATH} not found: build it with vltk_b200/csrc/build.sh "
            "(or __graft_entry__.build()). There is no CPU fallback for this path.")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(handle, name)  # AttributeError if the header and the .so drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().vltk_frcnn_last_error().decode("utf-8", "replace")
        raise LibraryError(f"{what} failed ({rc}): {msg}")


def make_config(cfg, mode: str) -> Config:
    c = Config()
    c.stem_out_channels = cfg.stem_out_channels
    c.res2_out_channels = cfg.res2_out_channels
    c.blocks = (C.c_int * 3)(*cfg.blocks_per_stage)
    c.res5_blocks = cfg.res5_blocks
    c.num_anchors = cfg.num_anchors
    c.anchor_stride = cfg.anchor_stride
    c.rpn_hidden = cfg.rpn_hidden
    c.rpn_nms_thresh = cfg.rpn_nms_thresh
    c.rpn_pre_nms_topk = cfg.rpn_pre_nms_topk
    c.rpn_post_nms_topk = cfg.rpn_post_nms_topk
    c.rpn_min_size = cfg.rpn_min_size
    c.rpn_bbox_weights = (C.c_float * 4)(*cfg.rpn_bbox_weights)
    c.pooler_resolution = cfg.pooler_resolution
    c.num_classes = cfg.num_classes
    c.num_attrs = cfg.num_attrs
    c.roi_bbox_weights = (C.c_float * 4)(*cfg.roi_bbox_weights)
    c.mode = MODES[mode]
    return c


def make_knobs(nms_thresh, min_det: int, max_det: int, pad_value: float = 0.0, ignorey=None) -> Knobs:
    """ignorey: C-contiguous float32 numpy [N, J, 2] (kept alive by the caller for the duration of the call)."""
    nms_thresh = list(nms_thresh) if isinstance(nms_thresh, (list, tuple)) else [nms_thresh]
    if not 1 <= len(nms_thresh) <= 4:
        raise ValueError("between 1 and 4 NMS thresholds are supported")
    k = Knobs()
    k.nms_thresh = (C.c_float * 4)(*(nms_thresh + [0.0] * (4 - len(nms_thresh))))
    k.n_nms_thresh = len(nms_thresh)
    k.min_detections = int(min_det)
    k.max_detections = int(max_det)
    k.pad_value = float(pad_value)
    if ignorey is not None and ignorey.shape[1] > 0:
        k.ignorey = ignorey.ctypes.data
        k.n_ignorey = int(ignorey.shape[1])
    return k
