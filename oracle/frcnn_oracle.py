"""TEST INFRASTRUCTURE — CPU restatement (torch fp32 + numpy) of the reference's
Faster R-CNN R101-C4 VG extraction path.  Never imported by the product package:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may use it, and only as the checker / CPU baseline.

Parity status: PINNED — tests/test_oracle_vs_golden.py checks every stage of this
file against tests/golden/*.npz, which oracle/make_goldens.py produced by running
the UNMODIFIED reference (/root/reference/vltk/modeling/frcnn.py) in the authoring
container on the same seeded weights/images.  (The reference's own tests pin no
values for this path — SURVEY.md §4/§8c.)

Third-party arithmetic the reference reaches through compiled wheels and that is
therefore restated from its published algorithm here: torchvision.ops.RoIPool and
torchvision.ops.nms (torchvision, unpinned in the reference's requirements.txt:53;
0.26.0 in this image).  Dense layers use torch's CPU conv2d/linear directly (the
same third-party kernels the reference calls).

All `frcnn.py:a-b` citations are /root/reference/vltk/modeling/frcnn.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default (frcnn.py:163-173)
SCALE_CLAMP = math.log(1000.0 / 16)  # frcnn.py:510


# --------------------------------------------------------------------------- dense
def conv_bn(sd, prefix, x, stride=1, pad=0, dil=1, relu=False):
    """Conv2d(bias=False) -> frozen BatchNorm2d [-> ReLU] (frcnn.py:794-822)."""
    y = F.conv2d(x, sd[prefix + ".weight"], None, stride, pad, dil)
    n = prefix + ".norm"
    y = F.batch_norm(y, sd[n + ".running_mean"], sd[n + ".running_var"], sd[n + ".weight"],
                     sd[n + ".bias"], False, 0.0, BN_EPS)
    return F.relu_(y) if relu else y


def bottleneck(sd, prefix, x, stride, dil):
    """BottleneckBlock.forward with stride in the 1x1 (frcnn.py:932, 963-979)."""
    out = conv_bn(sd, prefix + ".conv1", x, stride=stride, relu=True)
    out = conv_bn(sd, prefix + ".conv2", out, pad=dil, dil=dil, relu=True)
    out = conv_bn(sd, prefix + ".conv3", out)
    if (prefix + ".shortcut.weight") in sd:
        sc = conv_bn(sd, prefix + ".shortcut", x, stride=stride)
    else:
        sc = x
    out += sc
    return F.relu_(out)


def stem(sd, images):
    """BasicStem.forward, caffe max-pool variant (frcnn.py:872-879)."""
    x = conv_bn(sd, "backbone.stem.conv1", images, stride=2, pad=3, relu=True)
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=0, ceil_mode=True)


def _count_blocks(sd, stage_prefix):
    n = 0
    while f"{stage_prefix}.{n}.conv1.weight" in sd:
        n += 1
    return n


def backbone(sd, images, return_all=False):
    """ResNet.forward up to res4 (frcnn.py:1076-1090, make_stage :1101-1143)."""
    feats = {}
    x = stem(sd, images)
    feats["stem"] = x
    for name, first_stride in (("res2", 1), ("res3", 2), ("res4", 2)):
        p = "backbone." + name
        for b in range(_count_blocks(sd, p)):
            x = bottleneck(sd, f"{p}.{b}", x, first_stride if b == 0 else 1, 1)
        feats[name] = x
    return feats if return_all else x


def rpn_head(sd, res4):
    """RPNHead.forward (frcnn.py:1561-1572) -> NCHW logits [N,A,H,W], deltas [N,4A,H,W]."""
    p = "proposal_generator.rpn_head."
    t = F.relu(F.conv2d(res4, sd[p + "conv.weight"], sd[p + "conv.bias"], 1, 1))
    logits = F.conv2d(t, sd[p + "objectness_logits.weight"], sd[p + "objectness_logits.bias"])
    deltas = F.conv2d(t, sd[p + "anchor_deltas.weight"], sd[p + "anchor_deltas.bias"])
    return logits, deltas


def res5_head(sd, pooled, chunk=64):
    """Res5ROIHeads._shared_roi_transform + mean (frcnn.py:1345-1355, 1387-1401):
    three bottlenecks, stride 1, conv2 dilation 2 / pad 2; then mean over 14x14."""
    outs = []
    nb = _count_blocks(sd, "roi_heads.res5")
    for s in range(0, pooled.shape[0], chunk):
        x = pooled[s:s + chunk]
        for b in range(nb):
            x = bottleneck(sd, f"roi_heads.res5.{b}", x, 1, 2)
        outs.append(x.mean(dim=[2, 3]))
    if not outs:
        return pooled.new_zeros((0, sd["roi_heads.res5.0.conv3.weight"].shape[0]))
    return torch.cat(outs, 0)


def box_predictor(sd, feats):
    """FastRCNNOutputLayers.forward with the VG attribute head (frcnn.py:1726-1740)."""
    p = "roi_heads.box_predictor."
    scores = F.linear(feats, sd[p + "cls_score.weight"], sd[p + "cls_score.bias"])
    deltas = F.linear(feats, sd[p + "bbox_pred.weight"], sd[p + "bbox_pred.bias"])
    max_class = scores.argmax(-1)  # over all classes incl. background (frcnn.py:1732)
    emb = sd[p + "cls_embedding.weight"][max_class]
    h = F.relu(F.linear(torch.cat([feats, emb], -1), sd[p + "fc_attr.weight"], sd[p + "fc_attr.bias"]))
    attr = F.linear(h, sd[p + "attr_score.weight"], sd[p + "attr_score.bias"])
    return scores, attr, deltas


# ------------------------------------------------- bf16-operand emulation ("bf16op")
# What the bf16 tensor-core mode computes, restated on the CPU: operands rounded to bf16, fp32 accumulation, the
# frozen BN as y = acc * scale + shift with the engine's folded constants, the sum rounded to bf16 on store.  The only
# differences left between this and the GPU are the fp32 summation order and the occasional bf16 rounding flip it
# causes, so teacher-forced stage tests can hold the bf16 kernels to a few bf16 ulps instead of a statistical bound.
def rb(t: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even to bf16, kept in fp32"""
    return t.to(torch.bfloat16).to(torch.float32)


def bn_fold(sd, n):
    """frozen BN (eps 1e-5, frcnn.py:163-173) as y = x * scale + shift, fp32 like the engine's pack_layer"""
    scale = sd[n + ".weight"] * (1.0 / torch.sqrt(sd[n + ".running_var"] + BN_EPS))
    return scale, sd[n + ".bias"] - sd[n + ".running_mean"] * scale


def conv_bn_bf16(sd, prefix, x, stride=1, pad=0, dil=1, relu=False, residual=None):
    """Conv2d + frozen BN (+ residual) (+ ReLU) the way conv_tc computes it; x (and residual) hold bf16 values."""
    acc = F.conv2d(x, rb(sd[prefix + ".weight"]), None, stride, pad, dil)
    scale, shift = bn_fold(sd, prefix + ".norm")
    y = torch.addcmul(shift.view(1, -1, 1, 1), acc, scale.view(1, -1, 1, 1))
    if residual is not None:
        y = y + residual
    return rb(F.relu_(y) if relu else y)


def bottleneck_bf16(sd, prefix, x, stride, dil):
    """BottleneckBlock.forward (frcnn.py:963-979) in the bf16 mode's arithmetic.  Projection blocks run conv3 and the
    shortcut as ONE K-concatenated GEMM with both BN scales folded into the bf16 weights (csrc/conv_tc.cuh TcConcat)."""
    t1 = conv_bn_bf16(sd, prefix + ".conv1", x, stride=stride, relu=True)
    t2 = conv_bn_bf16(sd, prefix + ".conv2", t1, pad=dil, dil=dil, relu=True)
    if (prefix + ".shortcut.weight") in sd:
        s3, b3 = bn_fold(sd, prefix + ".conv3.norm")
        ss, bs = bn_fold(sd, prefix + ".shortcut.norm")
        w3 = rb(s3.view(-1, 1, 1, 1) * sd[prefix + ".conv3.weight"])
        ws = rb(ss.view(-1, 1, 1, 1) * sd[prefix + ".shortcut.weight"])
        y = F.conv2d(t2, w3) + F.conv2d(x, ws, None, stride) + (b3 + bs).view(1, -1, 1, 1)
        return rb(F.relu_(y))
    return conv_bn_bf16(sd, prefix + ".conv3", t2, relu=True, residual=x)


def stem_bf16(sd, images):
    """BasicStem (frcnn.py:872-879): the bf16 mode rounds the fp32 pixels to bf16 in its im2col rows."""
    x = conv_bn_bf16(sd, "backbone.stem.conv1", rb(images), stride=2, pad=3, relu=True)
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=0, ceil_mode=True)


def stage_bf16(sd, name, x, blocks=None):
    """one residual stage (`res2` | `res3` | `res4`), or its blocks [blocks[0], blocks[1]), on a bf16-valued NCHW tensor"""
    p = "backbone." + name
    first_stride = 1 if name == "res2" else 2
    b0, b1 = blocks or (0, _count_blocks(sd, p))
    for b in range(b0, b1):
        x = bottleneck_bf16(sd, f"{p}.{b}", x, first_stride if b == 0 else 1, 1)
    return x


def rpn_head_bf16(sd, res4):
    """RPNHead (frcnn.py:1561-1572) in the bf16 mode: 3x3 conv on bf16 operands, hidden map stored in bf16; the 1x1
    head keeps fp32 weights (bf16 hi + lo planes on the tensor pipe) and fp32 outputs."""
    p = "proposal_generator.rpn_head."
    t = rb(F.relu(F.conv2d(res4, rb(sd[p + "conv.weight"]), sd[p + "conv.bias"], 1, 1)))
    logits = F.conv2d(t, sd[p + "objectness_logits.weight"], sd[p + "objectness_logits.bias"])
    deltas = F.conv2d(t, sd[p + "anchor_deltas.weight"], sd[p + "anchor_deltas.bias"])
    return logits, deltas


# --------------------------------------------------------------------- box algebra
def grid_anchors(cell: torch.Tensor, h: int, w: int, stride: int) -> torch.Tensor:
    """[h*w*A, 4]; index (y*w + x)*A + a, offset 0 (frcnn.py:176-197, 1463-1477)."""
    sx = torch.arange(0, w * stride, stride, dtype=torch.float32)
    sy = torch.arange(0, h * stride, stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), 1)
    return (shifts[:, None, :] + cell[None, :, :]).reshape(-1, 4)


def apply_deltas(deltas: torch.Tensor, boxes: torch.Tensor, weights) -> torch.Tensor:
    """Box2BoxTransform.apply_deltas for k=1 (frcnn.py:548-584): no +1 widths, deltas
    divided by weights, dw/dh clamped from above at log(1000/16)."""
    wx, wy, ww, wh = weights
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * widths
    cy = boxes[:, 1] + 0.5 * heights
    dx = deltas[:, 0] / wx
    dy = deltas[:, 1] / wy
    dw = torch.clamp(deltas[:, 2] / ww, max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3] / wh, max=SCALE_CLAMP)
    pcx = dx * widths + cx
    pcy = dy * heights + cy
    pw = torch.exp(dw) * widths
    ph = torch.exp(dh) * heights
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), 1)


def clip_boxes_(boxes: torch.Tensor, hw) -> torch.Tensor:
    """_clip_box (frcnn.py:147-153): x in [0,w], y in [0,h], in place."""
    h, w = float(hw[0]), float(hw[1])
    boxes[:, 0].clamp_(0, w)
    boxes[:, 1].clamp_(0, h)
    boxes[:, 2].clamp_(0, w)
    boxes[:, 3].clamp_(0, h)
    return boxes


# The restatements below (nms_np, roi_pool_np) are the CHECKER: slow Python/numpy loops, bit-exact with the compiled
# torchvision ops (tests/test_host_logic.py).  When this file is TIMED as the CPU baseline (bench.py --impl reference,
# cpu_baseline) the two ops are taken from torchvision itself, exactly as the reference calls them
# (frcnn.py:30-31, 132, 383, 1179), so the baseline is not slowed down by the restatement.
_THIRD_PARTY_OPS = False


def use_torchvision_ops(flag: bool):
    """True: nms / RoIPool run through torchvision.ops like the reference does (timing); False: the restatements."""
    global _THIRD_PARTY_OPS
    _THIRD_PARTY_OPS = bool(flag)


def nms_keep(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, max_keep: Optional[int] = None) -> torch.Tensor:
    """Kept indices (int64 tensor, score order) of greedy NMS, cut to max_keep."""
    if _THIRD_PARTY_OPS:
        import torchvision
        keep = torchvision.ops.nms(boxes.float(), scores.float(), float(thresh))
        return keep if max_keep is None else keep[:max_keep]
    return torch.from_numpy(nms_np(boxes.numpy(), scores.numpy(), thresh, max_keep))


def roi_pool(feat: torch.Tensor, rois: torch.Tensor, out: int, scale: float) -> torch.Tensor:
    if _THIRD_PARTY_OPS:
        import torchvision
        return torchvision.ops.RoIPool((out, out), scale)(feat, rois)
    return roi_pool_np(feat, rois, out, scale)


def nms_np(boxes: np.ndarray, scores: np.ndarray, thresh: float, max_keep: Optional[int] = None):
    """Greedy NMS as published for torchvision.ops.nms (CPU kernel): stable
    score-descending order; j suppressed iff IoU(i,j) > thresh; area=(x2-x1)*(y2-y1);
    NaN IoU (0/0) never suppresses.  All arithmetic float32.  Returns kept indices into
    `boxes`, in score order.  Stopping after max_keep survivors equals slicing the
    full result (frcnn.py:383-384, 132-133)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = (boxes[order, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, bool)
    keep = []
    thr = np.float32(thresh)
    zero = np.float32(0)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(order[i])
        if max_keep is not None and len(keep) >= max_keep:
            break
        j = slice(i + 1, n)
        w = np.maximum(zero, np.minimum(x2[i], x2[j]) - np.maximum(x1[i], x1[j]))
        h = np.maximum(zero, np.minimum(y2[i], y2[j]) - np.maximum(y1[i], y1[j]))
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):
            ovr = inter / (areas[i] + areas[j] - inter)
        dead[j] |= ovr > thr
    return np.asarray(keep, np.int64)


def roi_pool_np(feat: torch.Tensor, rois: torch.Tensor, out: int, scale: float) -> torch.Tensor:
    """torchvision.ops.RoIPool restated: feat [N,C,H,W]; rois [R,5]=(batch,x1,y1,x2,y2).
    start/end = round-half-away(coord*scale); size=max(end-start+1,1); bin p covers
    [floor(p*size/out), ceil((p+1)*size/out)) + start, clipped to the map; empty -> 0
    (frcnn.py:1179, 1195-1198)."""
    n, c, hh, ww = feat.shape
    r = rois.shape[0]
    res = feat.new_zeros((r, c, out, out))
    rr = rois.detach().cpu().numpy().astype(np.float32)
    fscale = np.float32(scale)

    def rnd(v):  # C roundf: half away from zero, on the float32 product
        v = np.float32(v) * fscale
        return int(np.sign(v) * np.floor(np.abs(v) + np.float32(0.5)))

    for i in range(r):
        b = int(rr[i, 0])
        sw, sh, ew, eh = rnd(rr[i, 1]), rnd(rr[i, 2]), rnd(rr[i, 3]), rnd(rr[i, 4])
        rw = max(ew - sw + 1, 1)
        rh = max(eh - sh + 1, 1)
        bh = np.float32(rh) / np.float32(out)
        bw = np.float32(rw) / np.float32(out)
        fm = feat[b]
        for ph in range(out):
            hs = min(max(int(np.floor(np.float32(ph) * bh)) + sh, 0), hh)
            he = min(max(int(np.ceil(np.float32(ph + 1) * bh)) + sh, 0), hh)
            if he <= hs:
                continue
            strip = fm[:, hs:he, :].amax(1)  # [C, W]
            for pw in range(out):
                ws = min(max(int(np.floor(np.float32(pw) * bw)) + sw, 0), ww)
                we = min(max(int(np.ceil(np.float32(pw + 1) * bw)) + sw, 0), ww)
                if we <= ws:
                    continue
                res[i, :, ph, pw] = strip[:, ws:we].amax(1)
    return res


# ----------------------------------------------------------------------- RPN select
def apply_ignorey(boxes: torch.Tensor, ranges: torch.Tensor, scale_x: torch.Tensor):
    """The `ignorey` branch of find_top_rpn_proposals (frcnn.py:328-366) for ONE image, on the decoded,
    not yet clipped top-k boxes.  ranges [J,2] are caller-given (y0, y1) pairs; the reference divides them
    by scales_yx[n, 1] (the X scale — reproduced as written).  Per range, in order:
      * drop every box that spans the whole range (y1 <= r0 and y2 >= r1)          (:333-336)
      * of the rest, boxes NOT lying entirely past it (not (y1 > r1 and y2 > r0)) are clipped: the box edge
        nearer to its range end moves there — y2 = int(r0) when |r1-y2| < |r0-y1|, y1 = int(r1) when
        |r0-y1| < |r1-y2| (`box_ignore_below` is a contradiction and never fires)  (:342-366)
    Returns (boxes after clipping [K,4], alive mask [K]); dropped rows keep their last coordinates.
    The reference compacts the arrays instead of masking and, because it also overwrites the shared
    `level_ids` (:340), only survives batches of one image; semantics per image are identical."""
    boxes = boxes.clone()
    alive = torch.ones(len(boxes), dtype=torch.bool)
    for r in ranges:
        r = (r * 1) / scale_x                       # `ignoreyij * 1/(scales_yx[n, 1])` (:331)
        r0, r1 = r[0], r[1]
        y1, y2 = boxes[:, 1], boxes[:, 3]
        alive &= ~((r1 <= y2) & (r0 >= y1))
        past = (y1 > r1) & (y2 > r0)
        to_clip = alive & ~past
        clip_top = to_clip & ((r1 - y2).abs() < (r0 - y1).abs())
        clip_bottom = to_clip & ((r0 - y1).abs() < (r1 - y2).abs())
        boxes[clip_bottom, 1] = float(int(r1))
        boxes[clip_top, 3] = float(int(r0))
    return boxes, alive


def rpn_select(cfg, logits_nchw, deltas_nchw, cell, image_shapes, return_debug=False, ignorey=None,
               scales_yx=None):
    """predict_proposals + predict_objectness_logits + find_top_rpn_proposals +
    RPN.inference re-sort (frcnn.py:748-781, 264-390, 1615-1638), single level.

    Returns per image (boxes [n<=post,4], logits [n]); with return_debug also the
    top-k anchor indices, their decoded+clipped boxes and the kept positions.
    ignorey [N,J,2] is honoured only together with scales_yx (frcnn.py:328)."""
    n, a, h, w = logits_nchw.shape
    logits = logits_nchw.permute(0, 2, 3, 1).reshape(n, -1)
    deltas = deltas_nchw.view(n, a, 4, h, w).permute(0, 3, 4, 1, 2).reshape(n, -1, 4)
    anchors = grid_anchors(cell, h, w, cfg.anchor_stride)
    k = min(cfg.rpn_pre_nms_topk, logits.shape[1])
    out, dbg = [], []
    for i in range(n):
        # sort desc; ties resolve lower-index-first (what torch's CPU sort does)
        srt, idx = torch.sort(logits[i], descending=True, stable=True)
        idx = idx[:k]
        sc = srt[:k]
        boxes = apply_deltas(deltas[i][idx], anchors[idx], cfg.rpn_bbox_weights)
        alive = torch.ones(len(boxes), dtype=torch.bool)
        if ignorey is not None and scales_yx is not None:
            boxes, alive = apply_ignorey(boxes, torch.as_tensor(ignorey, dtype=torch.float32)[i],
                                         torch.as_tensor(scales_yx, dtype=torch.float32)[i, 1])
        clip_boxes_(boxes, image_shapes[i])
        ok = alive & ((boxes[:, 2] - boxes[:, 0]) > cfg.rpn_min_size) & \
             ((boxes[:, 3] - boxes[:, 1]) > cfg.rpn_min_size)
        pos = torch.nonzero(ok).squeeze(1)
        keep = nms_keep(boxes[pos], sc[pos], cfg.rpn_nms_thresh, cfg.rpn_post_nms_topk)
        kept_pos = pos[keep]  # positions inside the sorted top-k list
        out.append((boxes[kept_pos], sc[kept_pos]))
        dbg.append({"topk_idx": idx, "topk_boxes": boxes, "topk_scores": sc, "kept_pos": kept_pos})
    return (out, dbg) if return_debug else out


# --------------------------------------------------------------------- ROI outputs
def roi_outputs(cfg, obj_logits, attr_logits, box_deltas, proposals: Sequence[torch.Tensor],
                feats, image_shapes, scales_yx=None, return_keep=False):
    """ROIOutputs.inference + do_nms (frcnn.py:1262-1294, 116-143).

    softmax over all classes then drop the (last) background column; attributes drop
    the last column then softmax; per-ROI best foreground class picks the box; clip;
    class-agnostic NMS; first max_detections; thresholds tried in order until the
    count lies in [min,max] (the last attempt is returned regardless)."""
    counts = [int(p.shape[0]) for p in proposals]
    probs_all = F.softmax(obj_logits, dim=-1)
    attr_p_all, attr_i_all = attr_logits[..., :-1].softmax(-1).max(-1)
    res = {k: [] for k in ("boxes", "obj_ids", "obj_probs", "attr_ids", "attr_probs",
                           "roi_features", "keep", "all_boxes", "all_scores")}
    s = 0
    for i, cnt in enumerate(counts):
        sl = slice(s, s + cnt)
        s += cnt
        probs = probs_all[sl][:, :-1]
        max_scores, max_classes = probs.max(1)
        d = box_deltas[sl].view(cnt, -1, 4)
        sel = d[torch.arange(cnt), max_classes]  # only the winning class's deltas matter
        boxes = apply_deltas(sel, proposals[i], cfg.roi_bbox_weights)
        clip_boxes_(boxes, image_shapes[i])
        keep = None
        for thr in cfg.nms_thresh_test:
            keep = nms_keep(boxes, max_scores, thr)[: cfg.max_detections]
            if cfg.min_detections <= len(keep) <= cfg.max_detections:
                break
        kb = boxes[keep].clone()
        if scales_yx is not None:
            kb[:, 0::2] *= scales_yx[i][1]
            kb[:, 1::2] *= scales_yx[i][0]
        res["boxes"].append(kb)
        res["obj_ids"].append(max_classes[keep])
        res["obj_probs"].append(max_scores[keep])
        res["attr_ids"].append(attr_i_all[sl][keep])
        res["attr_probs"].append(attr_p_all[sl][keep])
        res["roi_features"].append(feats[sl][keep])
        res["keep"].append(keep)
        res["all_boxes"].append(boxes)      # every ROI's winning-class box (clipped) and score:
        res["all_scores"].append(max_scores)  # what the final NMS ranked (margin analysis)
    if not return_keep:
        for k in ("keep", "all_boxes", "all_scores"):
            res.pop(k)
    return res


# ------------------------------------------------------------------------- forward
@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], cfg, images: torch.Tensor, image_shapes,
            scales_yx=None, stages: Optional[dict] = None, res5_chunk: int = 64, ignorey=None):
    """FRCNN.inference (frcnn.py:1942-2004).  images [N,3,H,W] f32 normalised+padded;
    image_shapes [N,2] resized (h,w); returns the reference's ragged dict plus `keep`
    (indices into each image's proposal list).  `stages`, if given, is filled with the
    intermediate tensors used for teacher-forced per-stage parity tests."""
    image_shapes = [(int(s[0]), int(s[1])) for s in image_shapes]
    res4 = backbone(sd, images)
    logits, deltas = rpn_head(sd, res4)
    cell = sd["proposal_generator.anchor_generator.cell_anchors.0"]
    props, dbg = rpn_select(cfg, logits, deltas, cell, image_shapes, return_debug=True, ignorey=ignorey,
                            scales_yx=scales_yx)
    boxes = [p[0] for p in props]
    rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1)
                      for i, b in enumerate(boxes)], 0)
    pooled = roi_pool(res4, rois, cfg.pooler_resolution, 1.0 / cfg.anchor_stride)
    feats = res5_head(sd, pooled, chunk=res5_chunk)
    obj_logits, attr_logits, box_deltas = box_predictor(sd, feats)
    out = roi_outputs(cfg, obj_logits, attr_logits, box_deltas, boxes, feats, image_shapes,
                      scales_yx, return_keep=True)
    out["preds_per_image"] = torch.tensor([len(b) for b in out["boxes"]])
    if stages is not None:
        stages.update(res4=res4, rpn_logits=logits, rpn_deltas=deltas, rpn_debug=dbg,
                      proposals=boxes, proposal_logits=[p[1] for p in props], pooled=pooled,
                      feats=feats, obj_logits=obj_logits, attr_logits=attr_logits,
                      box_deltas=box_deltas)
    return out


def pad_outputs(out: dict, image_shapes, scales_yx, max_det: int, pad_value: float = 0.0):
    """v1.0.0 `padding="max_detections"` contract (the code survives
    commented-out at frcnn.py:40-113, 1979-1995): dense [N,max_det,...] tensors,
    plus sizes and normalized_boxes = boxes / (image_shapes*scales_yx), x by width."""
    n = len(out["boxes"])

    def pad(lst, tail, dtype):
        t = torch.full((n, max_det) + tail, pad_value, dtype=dtype)
        for i, v in enumerate(lst):
            t[i, : v.shape[0]] = v
        return t

    d = out["roi_features"][0].shape[-1] if n else 0
    res = {
        "obj_ids": pad(out["obj_ids"], (), torch.int64),
        "obj_probs": pad(out["obj_probs"], (), torch.float32),
        "attr_ids": pad(out["attr_ids"], (), torch.int64),
        "attr_probs": pad(out["attr_probs"], (), torch.float32),
        "boxes": pad(out["boxes"], (4,), torch.float32),
        "roi_features": pad(out["roi_features"], (d,), torch.float32),
        "preds_per_image": out["preds_per_image"].clone(),
        "sizes": torch.as_tensor(image_shapes).clone(),
    }
    raw = torch.as_tensor(image_shapes, dtype=torch.float32)
    if scales_yx is not None:
        raw = raw * torch.as_tensor(scales_yx, dtype=torch.float32)
    nb = res["boxes"].clone()
    nb[:, :, 0::2] /= raw[:, 1].view(-1, 1, 1)
    nb[:, :, 1::2] /= raw[:, 0].view(-1, 1, 1)
    res["normalized_boxes"] = nb
    return res


# ---------------------------------------------------------------------- preprocess
def preprocess(cfg, raw_images: List[torch.Tensor]):
    """Preprocess.__call__ for tensor inputs (legacy/processing.py:112-150): each raw
    image is [h,w,3] (BGR, any dtype) -> float -> bilinear shortest-edge resize
    (align_corners=False, no antialias; :40-73) -> (x-mean)/std -> zero-pad bottom/right
    to the batch max (:98-110).  Returns images [N,3,H,W], sizes [N,2], scales_yx."""
    from vltk_b200.synthetic import resized_hw
    mean = torch.tensor(cfg.pixel_mean).view(3, 1, 1)
    std = torch.tensor(cfg.pixel_std).view(3, 1, 1)
    outs, raw_sizes = [], []
    for im in raw_images:
        im = im.float()
        h, w = im.shape[:2]
        raw_sizes.append((h, w))
        nh, nw = resized_hw(h, w, cfg)
        x = F.interpolate(im.permute(2, 0, 1).unsqueeze(0), (nh, nw), mode="bilinear",
                          align_corners=False).squeeze(0)
        outs.append((x - mean) / std)
    hm = max(o.shape[1] for o in outs)
    wm = max(o.shape[2] for o in outs)
    sizes = torch.tensor([o.shape[-2:] for o in outs])
    batch = torch.stack([F.pad(o, [0, wm - o.shape[2], 0, hm - o.shape[1]], value=cfg.pad_value)
                         for o in outs])
    scales = torch.true_divide(torch.tensor(raw_sizes), sizes)
    return batch, sizes, scales
