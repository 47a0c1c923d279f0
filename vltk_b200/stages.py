"""Torch-tensor wrappers over the stage entry points of include/vltk_frcnn.h — used by the
teacher-forced parity tests and the config-4 microbenchmarks.  Each mirrors the reference
call it replaces; all tensors must live on the CUDA device (no CPU fallback)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def conv2d_nhwc(x, weight, scale=None, shift=None, residual=None, stride=1, pad=0, dil=1, relu=False,
                mode="fp32", tensor_cores=False):
    """x [N,H,W,Cin] (f32 or bf16 per `mode`), weight [Cout,Cin,k,k] f32 (reference layout);
    returns y [N,OH,OW,Cout] = act(conv(x,w)*scale + shift + residual)
    (reference: Conv2d+BN+ReLU, frcnn.py:794-822; bottleneck add, :963-979).
    mode="exact_tc": the fp32-faithful tensor-core kernel (csrc/conv_tcx.cu); tensor_cores=2 takes its fp32-output
    epilogue (cout % 128 == 0, no residual) instead of the split-fp16 one."""
    L = _lib.lib()
    dt = torch.bfloat16 if mode == "bf16" else torch.float32     # exact_tc: fp32 tensors, split / widened inside
    assert x.is_cuda and x.dtype == dt and x.is_contiguous()
    n, h, w, cin = x.shape
    cout, cin2, kh, kw = weight.shape
    assert cin2 == cin
    oh = (h + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    ow = (w + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    y = torch.empty((n, oh, ow, cout), dtype=dt, device=x.device)
    wt = weight.to(x.device, torch.float32).contiguous()
    sc = None if scale is None else scale.to(x.device, torch.float32).contiguous()
    sh = None if shift is None else shift.to(x.device, torch.float32).contiguous()
    if residual is not None:
        assert residual.dtype == dt and residual.shape == y.shape and residual.is_contiguous()
    with torch.cuda.device(x.device):
        _lib.check(L.vltk_conv2d_nhwc(x.data_ptr(), wt.data_ptr(), _ptr(sc), _ptr(sh), _ptr(residual),
                                      y.data_ptr(), n, h, w, cin, cout, kh, kw, stride, pad, dil,
                                      int(relu), _lib.MODES[mode], int(tensor_cores), _stream(x)),
                   "vltk_conv2d_nhwc")
    return y


def conv2d_meanpool_nhwc(x, weight, scale, shift, residual, pool_rows, stride=1, pad=0, dil=1, relu=True):
    """The res5 tail on the tensor pipe (frcnn.py:1389, 1401): act(conv(x,w)*scale + shift + residual) whose
    output is reduced to the mean of every `pool_rows` consecutive pixels instead of being stored.
    x, residual: bf16 NHWC; returns f32 [N*OH*OW/pool_rows, Cout]."""
    L = _lib.lib()
    assert x.dtype == torch.bfloat16 and residual.dtype == torch.bfloat16 and x.is_contiguous() and residual.is_contiguous()
    n, h, w, cin = x.shape
    cout, _, k, _ = weight.shape
    oh = (h + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    ow = (w + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    m = n * oh * ow
    assert m % pool_rows == 0
    out = torch.empty((m // pool_rows, cout), dtype=torch.float32, device=x.device)
    wt = weight.to(x.device, torch.float32).contiguous()
    sc, sh = scale.to(x.device, torch.float32).contiguous(), shift.to(x.device, torch.float32).contiguous()
    with torch.cuda.device(x.device):
        _lib.check(L.vltk_conv2d_meanpool_nhwc(x.data_ptr(), wt.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                                               residual.data_ptr(), out.data_ptr(), n, h, w, cin, cout, k,
                                               stride, pad, dil, int(relu), pool_rows, _stream(x)),
                   "vltk_conv2d_meanpool_nhwc")
    return out


def conv2d_meanpool_exact_nhwc(x, weight, scale, shift, residual, pool_rows, relu=True):
    """The same fused tail in exact_tc mode (vltk_conv2d_meanpool_exact_nhwc): fp32 NHWC x / residual, 1x1 weight
    [cout, cin, 1, 1]; returns f32 [N*h*w/pool_rows, cout]."""
    L = _lib.lib()
    assert x.dtype == torch.float32 and residual.dtype == torch.float32 and x.is_contiguous() and residual.is_contiguous()
    n, h, w, cin = x.shape
    cout = weight.shape[0]
    m = n * h * w
    assert m % pool_rows == 0
    out = torch.empty((m // pool_rows, cout), dtype=torch.float32, device=x.device)
    wt = weight.to(x.device, torch.float32).reshape(cout, cin).contiguous()
    sc, sh = scale.to(x.device, torch.float32).contiguous(), shift.to(x.device, torch.float32).contiguous()
    with torch.cuda.device(x.device):
        _lib.check(L.vltk_conv2d_meanpool_exact_nhwc(x.data_ptr(), wt.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                                                     residual.data_ptr(), out.data_ptr(), n, h, w, cin, cout, int(relu),
                                                     pool_rows, _stream(x)), "vltk_conv2d_meanpool_exact_nhwc")
    return out


def conv2d_dual_nhwc(x, weight, x2, weight2, shift=None, stride2=1, relu=True):
    """A projection bottleneck's tail on the tensor pipe (frcnn.py:918-925, 971-979): conv3(x) + shortcut(x2)
    as ONE K-concatenated GEMM, y = act(x.w^T + x2[:, ::stride2, ::stride2].w2^T + shift); BN scales are
    expected to be folded into the weights already.  x [N,h,w,cin], x2 [N,h2,w2,cin2] bf16 NHWC;
    weight [cout,cin], weight2 [cout,cin2] f32."""
    L = _lib.lib()
    assert x.dtype == torch.bfloat16 and x2.dtype == torch.bfloat16 and x.is_contiguous() and x2.is_contiguous()
    n, h, w, cin = x.shape
    _, h2, w2, cin2 = x2.shape
    cout = weight.shape[0]
    assert (h2 - 1) // stride2 + 1 == h and (w2 - 1) // stride2 + 1 == w
    y = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=x.device)
    wa = weight.to(x.device, torch.float32).reshape(cout, cin).contiguous()
    wb = weight2.to(x.device, torch.float32).reshape(cout, cin2).contiguous()
    sh = None if shift is None else shift.to(x.device, torch.float32).contiguous()
    with torch.cuda.device(x.device):
        _lib.check(L.vltk_conv2d_dual_nhwc(x.data_ptr(), wa.data_ptr(), x2.data_ptr(), wb.data_ptr(), _ptr(sh),
                                           y.data_ptr(), n, h, w, cin, h2, w2, cin2, stride2, cout, int(relu),
                                           _stream(x)), "vltk_conv2d_dual_nhwc")
    return y


def set_cta_pairs(min_pixels=-1, residual_layers=-1):
    """Process-wide kernel selection for the tcgen05 convolutions (include/vltk_frcnn.h,
    vltk_conv_tc_set_cta_pairs): layers with >= `min_pixels` output pixels run on CTA pairs (0 = never);
    `residual_layers`=0 keeps shortcut / fused-mean layers on the single-CTA kernel.  -1 = unchanged."""
    _lib.lib().vltk_conv_tc_set_cta_pairs(int(min_pixels), int(residual_layers))


def set_cta_pairs_exact(min_pixels=-1):
    """The same switch for the exact_tc kernels (vltk_conv_tcx_set_cta_pairs); 0 = never, -1 = unchanged."""
    _lib.lib().vltk_conv_tcx_set_cta_pairs(int(min_pixels))


def linear_tc3(x, weight, bias=None, relu=False):
    """F.linear(x, weight, bias) on the tensor pipe with split-bf16 (hi*hi + lo*hi + hi*lo) operands
    and fp32 accumulate/output — how the predictor linears (frcnn.py:1729-1737) run in bf16 mode."""
    L = _lib.lib()
    m, k = x.shape
    n = weight.shape[0]
    y = torch.empty((m, n), dtype=torch.float32, device=x.device)
    xx, ww = x.float().contiguous(), weight.float().contiguous()
    bb = None if bias is None else bias.float().contiguous()
    with torch.cuda.device(x.device):
        _lib.check(L.vltk_linear_tc3(xx.data_ptr(), ww.data_ptr(), _ptr(bb), y.data_ptr(), m, k, n, int(relu),
                                     _stream(x)), "vltk_linear_tc3")
    return y


def rpn_proposals(logits, deltas, cell_anchors, image_shapes, cfg):
    """find_top_rpn_proposals + RPN.inference (frcnn.py:264-390, 1615-1638) on NCHW head
    outputs.  Returns (proposals [N,post,4], logits [N,post], counts [N])."""
    L = _lib.lib()
    assert logits.is_cuda and deltas.is_cuda
    n, a, h4, w4 = logits.shape
    post = cfg.rpn_post_nms_topk
    props = torch.empty((n, post, 4), dtype=torch.float32, device=logits.device)
    plog = torch.empty((n, post), dtype=torch.float32, device=logits.device)
    counts = torch.empty((n,), dtype=torch.int32, device=logits.device)
    cell = np.ascontiguousarray(cell_anchors.detach().cpu().numpy(), dtype=np.float32)
    sizes = np.ascontiguousarray(np.asarray(image_shapes), dtype=np.int32).reshape(n, 2)
    wts = (C.c_float * 4)(*cfg.rpn_bbox_weights)
    lg = logits.float().contiguous()
    dl = deltas.float().contiguous()
    with torch.cuda.device(logits.device):
        _lib.check(L.vltk_rpn_proposals(lg.data_ptr(), dl.data_ptr(), cell.ctypes.data, sizes.ctypes.data,
                                        n, a, h4, w4, cfg.anchor_stride, cfg.rpn_pre_nms_topk, post,
                                        cfg.rpn_nms_thresh, cfg.rpn_min_size, wts, props.data_ptr(),
                                        plog.data_ptr(), counts.data_ptr(), _stream(logits)),
                   "vltk_rpn_proposals")
    return props, plog, counts


def nms(boxes, scores, thresh, max_keep=None):
    """torchvision.ops.nms(boxes, scores, thresh)[:max_keep] (frcnn.py:132-133, 383-384)."""
    L = _lib.lib()
    k = boxes.shape[0]
    max_keep = max(1, min(max_keep or k, 8192))
    keep = torch.empty((max_keep,), dtype=torch.int32, device=boxes.device)
    count = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
    b = boxes.float().contiguous()
    s = scores.float().contiguous()
    with torch.cuda.device(boxes.device):
        _lib.check(L.vltk_nms(b.data_ptr(), s.data_ptr(), k, float(thresh), max_keep, keep.data_ptr(),
                              count.data_ptr(), _stream(boxes)), "vltk_nms")
    return keep[: int(count.item())].to(torch.int64)


def roi_pool(feat, rois, output_size, spatial_scale):
    """torchvision.ops.RoIPool(output_size, spatial_scale)(feat, rois) (frcnn.py:1179, 1198)."""
    L = _lib.lib()
    n, c, h, w = feat.shape
    r = rois.shape[0]
    out = torch.empty((r, c, output_size, output_size), dtype=torch.float32, device=feat.device)
    f = feat.float().contiguous()
    rr = rois.float().contiguous()
    with torch.cuda.device(feat.device):
        _lib.check(L.vltk_roi_pool_nchw(f.data_ptr(), n, c, h, w, rr.data_ptr(), r, output_size,
                                        float(spatial_scale), out.data_ptr(), _stream(feat)),
                   "vltk_roi_pool_nchw")
    return out


def roi_outputs(obj_logits, attr_logits, box_deltas, feats, proposals, counts, image_shapes, scales_yx,
                cfg, nms_thresh=None, min_det=None, max_det=None, pad_value=0.0):
    """ROIOutputs.inference (frcnn.py:1262-1294) in the padded layout.  proposals [N,R,4]."""
    L = _lib.lib()
    dev = obj_logits.device
    n, r = proposals.shape[0], proposals.shape[1]
    d = feats.shape[1]
    md = max_det or cfg.max_detections
    t = dict(
        boxes=torch.empty((n, md, 4), dtype=torch.float32, device=dev),
        normalized_boxes=torch.empty((n, md, 4), dtype=torch.float32, device=dev),
        obj_ids=torch.empty((n, md), dtype=torch.int64, device=dev),
        obj_probs=torch.empty((n, md), dtype=torch.float32, device=dev),
        attr_ids=torch.empty((n, md), dtype=torch.int64, device=dev),
        attr_probs=torch.empty((n, md), dtype=torch.float32, device=dev),
        roi_features=torch.empty((n, md, d), dtype=torch.float32, device=dev),
        preds_per_image=torch.empty((n,), dtype=torch.int32, device=dev),
        keep_idx=torch.empty((n, md), dtype=torch.int32, device=dev))
    o = _lib.Out()
    for k, v in t.items():
        setattr(o, k, v.data_ptr())
    knobs = _lib.make_knobs(nms_thresh or cfg.nms_thresh_test, min_det if min_det is not None else cfg.min_detections,
                            md, pad_value)
    sizes = np.ascontiguousarray(np.asarray(image_shapes), dtype=np.int32).reshape(n, 2)
    sc = None if scales_yx is None else np.ascontiguousarray(np.asarray(scales_yx), dtype=np.float32).reshape(n, 2)
    wts = (C.c_float * 4)(*cfg.roi_bbox_weights)
    args = [x.float().contiguous() for x in (obj_logits, attr_logits, box_deltas, feats, proposals)]
    cn = counts.to(dev, torch.int32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(L.vltk_roi_outputs(*[a.data_ptr() for a in args], cn.data_ptr(), sizes.ctypes.data,
                                      None if sc is None else sc.ctypes.data, n, r, cfg.num_classes,
                                      cfg.num_attrs, d, wts, C.byref(knobs), C.byref(o), _stream(obj_logits)),
                   "vltk_roi_outputs")
    return t
