"""Seeded synthetic weights and images (there is no network for the hub checkpoint).

`make_state_dict` emits the reference's 640-key state_dict layout
(vltk/modeling/frcnn.py:1743-1755 builds the modules whose
parameters these are) with the *engineered* initialisation of SURVEY.md Appendix E:
plain He / default inits explode or give near-uniform posteriors through 101 layers,
which makes index parity vacuous.  `make_image` is the SURVEY §8(d) recipe.

Everything is drawn from a CPU `torch.Generator`, so the same seed gives the same
tensors here and on the GPU box.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch

from . import arch
from .config import FRCNNConfig

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def cell_anchors(cfg: FRCNNConfig) -> torch.Tensor:
    """15 cell anchors, size-major then ratio; float32 of python-double math
    (reference: frcnn.py:1479-1497)."""
    rows = []
    for size in cfg.anchor_sizes:
        area = size ** 2.0
        for r in cfg.anchor_ratios:
            w = math.sqrt(area / r)
            h = r * w
            rows.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(rows, dtype=torch.float64).float()


def _normal(g, shape, std):
    return torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=g)


def _uniform(g, shape, lo, hi):
    return torch.empty(shape, dtype=torch.float32).uniform_(lo, hi, generator=g)


def make_state_dict(cfg: FRCNNConfig, seed: int = 0, cls_bias: Optional[torch.Tensor] = "auto"):
    """Engineered random weights in the reference state_dict layout.

    cls_bias: "auto" loads the committed calibration vector
    (`vltk_b200/data/cls_bias_seed{seed}.npy`, produced by oracle/make_goldens.py with
    one calibration forward: bias = -W . mean(pooled features)); None leaves zeros.
    """
    g = torch.Generator().manual_seed(1000 + seed)
    sd = OrderedDict()

    def bn(prefix, c, glo, ghi):
        sd[prefix + ".weight"] = _uniform(g, (c,), glo, ghi)
        sd[prefix + ".bias"] = _normal(g, (c,), 0.1)
        sd[prefix + ".running_mean"] = _normal(g, (c,), 0.1)
        sd[prefix + ".running_var"] = _uniform(g, (c,), 0.5, 1.5)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)

    for spec in arch.all_bn_convs(cfg):
        fan_in = spec.cin * spec.k * spec.k
        sd[spec.name + ".weight"] = _normal(g, (spec.cout, spec.cin, spec.k, spec.k),
                                            math.sqrt(2.0 / fan_in))
        if spec.role == "stem":
            bn(spec.name + ".norm", spec.cout, 0.01, 0.03)   # inputs are +-128
        elif spec.role == "conv3":
            bn(spec.name + ".norm", spec.cout, 0.1, 0.3)     # damp the residual branch
        elif spec.role == "shortcut":
            bn(spec.name + ".norm", spec.cout, 0.6, 1.0)
        else:
            bn(spec.name + ".norm", spec.cout, 0.5, 1.5)

    a = cfg.num_anchors
    c4 = cfg.res2_out_channels * 4
    hid = cfg.rpn_hidden
    sd["proposal_generator.anchor_generator.cell_anchors.0"] = cell_anchors(cfg)
    p = "proposal_generator.rpn_head."
    sd[p + "conv.weight"] = _normal(g, (hid, c4, 3, 3), math.sqrt(2.0 / (c4 * 9)))
    sd[p + "conv.bias"] = _normal(g, (hid,), 0.05)
    sd[p + "objectness_logits.weight"] = _normal(g, (a, hid, 1, 1), 0.05)
    sd[p + "objectness_logits.bias"] = _normal(g, (a,), 0.05)
    sd[p + "anchor_deltas.weight"] = _normal(g, (a * 4, hid, 1, 1), 0.01)
    sd[p + "anchor_deltas.bias"] = _normal(g, (a * 4,), 0.01)

    d = arch.feature_dim(cfg)
    nc, na = cfg.num_classes, cfg.num_attrs
    p = "roi_heads.box_predictor."
    sd[p + "cls_score.weight"] = _normal(g, (nc + 1, d), 0.04)
    sd[p + "cls_score.bias"] = torch.zeros(nc + 1)
    sd[p + "bbox_pred.weight"] = _normal(g, (nc * 4, d), 0.01)
    sd[p + "bbox_pred.bias"] = _normal(g, (nc * 4,), 0.01)
    sd[p + "cls_embedding.weight"] = _normal(g, (nc + 1, d // 8), 1.0)
    sd[p + "fc_attr.weight"] = _normal(g, (d // 4, d + d // 8), math.sqrt(2.0 / (d + d // 8)))
    sd[p + "fc_attr.bias"] = _normal(g, (d // 4,), 0.05)
    sd[p + "attr_score.weight"] = _normal(g, (na + 1, d // 4), 0.05)
    sd[p + "attr_score.bias"] = _normal(g, (na + 1,), 0.05)

    if isinstance(cls_bias, str) and cls_bias == "auto":
        path = os.path.join(_DATA, f"cls_bias_seed{seed}.npy")
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `python oracle/make_goldens.py --calibrate {seed}` "
                "(needs the oracle; product code never calibrates)")
        cls_bias = torch.from_numpy(np.load(path))
    if cls_bias is not None:
        assert cls_bias.shape == (nc + 1,)
        sd[p + "cls_score.bias"] = cls_bias.float().clone()
    return sd


def make_raw_image(h: int, w: int, seed: int) -> torch.Tensor:
    """Raw BGR u8 [h, w, 3]: clamp(round(128 + 48*lowpass(N) + 16*N))."""
    g = torch.Generator().manual_seed(2000 + seed)
    ch, cw = max(h // 32, 2), max(w // 32, 2)
    coarse = torch.empty((1, 3, ch, cw), dtype=torch.float32).normal_(generator=g)
    # ATen's CPU bilinear kernel takes a different code path (results 1 ulp apart on half of the elements, a few dozen
    # flipped u8 pixels per image) when the process has ONE intra-op thread — which is what torchrun sets up
    # (OMP_NUM_THREADS=1).  The goldens were made with the multi-thread path (identical for 2, 3, 8, 16 threads), so pin it:
    # the same seed must give the same image in every launch mode.
    nthr = torch.get_num_threads()
    if nthr < 2:
        torch.set_num_threads(2)
    try:
        low = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear",
                                              align_corners=False)[0]
    finally:
        if nthr < 2:
            torch.set_num_threads(nthr)
    fine = torch.empty((3, h, w), dtype=torch.float32).normal_(generator=g)
    img = torch.round(128.0 + 48.0 * low + 16.0 * fine).clamp_(0, 255)
    return img.permute(1, 2, 0).contiguous().to(torch.uint8)


def resized_hw(h: int, w: int, cfg: FRCNNConfig):
    """ResizeShortestEdge size rule (reference: legacy/processing.py:48-60)."""
    size = cfg.min_size_test
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > cfg.max_size_test:
        scale = cfg.max_size_test * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)
