"""Static layer table of the R101-C4 detector.

Everything that needs to walk the network (synthetic weights, the weight packer,
the FLOP model used by bench.py, the oracle) iterates this table instead of
re-deriving shapes.  Names are the reference's state_dict prefixes
(reference: vltk/modeling/frcnn.py:857-979, 1101-1143,
1345-1385, 1545-1555, 1705-1719).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

from .config import FRCNNConfig


@dataclass(frozen=True)
class ConvSpec:
    name: str          # state_dict prefix, e.g. "backbone.res2.0.conv1"
    cin: int
    cout: int
    k: int
    stride: int = 1
    pad: int = 0
    dil: int = 1
    bn: bool = True    # frozen BN follows (no conv bias) — else plain bias
    relu: bool = True  # ReLU directly after BN/bias
    role: str = ""     # conv1|conv2|conv3|shortcut|stem|rpn


@dataclass(frozen=True)
class BlockSpec:
    name: str
    conv1: ConvSpec
    conv2: ConvSpec
    conv3: ConvSpec
    shortcut: Optional[ConvSpec]


def _stage(prefix: str, nblocks: int, cin: int, mid: int, cout: int, first_stride: int,
           dil: int) -> List[BlockSpec]:
    blocks = []
    for b in range(nblocks):
        s = first_stride if b == 0 else 1
        bin_ = cin if b == 0 else cout
        p = f"{prefix}.{b}"
        sc = None
        if bin_ != cout:
            sc = ConvSpec(f"{p}.shortcut", bin_, cout, 1, stride=s, relu=False, role="shortcut")
        blocks.append(BlockSpec(
            p,
            ConvSpec(f"{p}.conv1", bin_, mid, 1, stride=s, role="conv1"),
            ConvSpec(f"{p}.conv2", mid, mid, 3, pad=dil, dil=dil, role="conv2"),
            # conv3's ReLU comes after the residual add (frcnn.py:977-978)
            ConvSpec(f"{p}.conv3", mid, cout, 1, relu=False, role="conv3"),
            sc,
        ))
    return blocks


def stem_spec(cfg: FRCNNConfig) -> ConvSpec:
    return ConvSpec("backbone.stem.conv1", 3, cfg.stem_out_channels, 7, stride=2, pad=3, role="stem")


def backbone_stages(cfg: FRCNNConfig) -> List[List[BlockSpec]]:
    c2 = cfg.res2_out_channels
    n2, n3, n4 = cfg.blocks_per_stage
    return [
        _stage("backbone.res2", n2, cfg.stem_out_channels, c2 // 4, c2, 1, 1),
        _stage("backbone.res3", n3, c2, c2 // 2, c2 * 2, 2, 1),
        _stage("backbone.res4", n4, c2 * 2, c2, c2 * 4, 2, 1),
    ]


def res5_stage(cfg: FRCNNConfig) -> List[BlockSpec]:
    """VG head: first-block stride forced to 1, every conv2 dilation 2 / pad 2
    (frcnn.py:1345-1355)."""
    c2 = cfg.res2_out_channels
    return _stage("roi_heads.res5", cfg.res5_blocks, c2 * 4, c2 * 2, c2 * 8, 1, 2)


def rpn_conv_spec(cfg: FRCNNConfig) -> ConvSpec:
    return ConvSpec("proposal_generator.rpn_head.conv", cfg.res2_out_channels * 4, cfg.rpn_hidden,
                    3, pad=1, bn=False, relu=True, role="rpn")


def all_bn_convs(cfg: FRCNNConfig) -> List[ConvSpec]:
    out = [stem_spec(cfg)]
    for stage in backbone_stages(cfg) + [res5_stage(cfg)]:
        for blk in stage:
            if blk.shortcut is not None:
                out.append(blk.shortcut)
            out += [blk.conv1, blk.conv2, blk.conv3]
    return out


def feature_dim(cfg: FRCNNConfig) -> int:
    return cfg.res2_out_channels * 8


def flops_per_image(cfg: FRCNNConfig, h: int, w: int, rois: int) -> dict:
    """Algorithmic FLOPs (2*MACs) per image, per stage."""
    f = {}
    hs, ws = cfg.stem_conv_out(h), cfg.stem_conv_out(w)
    f["stem"] = 2 * hs * ws * 3 * cfg.stem_out_channels * 49
    ch, cw = cfg.stem_pool_out(hs), cfg.stem_pool_out(ws)
    for name, stage in zip(("res2", "res3", "res4"), backbone_stages(cfg)):
        tot = 0
        for blk in stage:
            oh = (ch - 1) // blk.conv1.stride + 1
            ow = (cw - 1) // blk.conv1.stride + 1
            for c in (blk.conv1, blk.conv2, blk.conv3, blk.shortcut):
                if c is not None:
                    tot += 2 * oh * ow * c.cin * c.cout * c.k * c.k
            ch, cw = oh, ow
        f[name] = tot
    a = cfg.num_anchors
    c4 = cfg.res2_out_channels * 4
    f["rpn"] = 2 * ch * cw * (c4 * cfg.rpn_hidden * 9 + cfg.rpn_hidden * a * 5)
    p = cfg.pooler_resolution
    tot = 0
    for blk in res5_stage(cfg):
        for c in (blk.conv1, blk.conv2, blk.conv3, blk.shortcut):
            if c is not None:
                tot += 2 * p * p * c.cin * c.cout * c.k * c.k
    f["res5"] = tot * rois
    d = feature_dim(cfg)
    f["predictor"] = 2 * rois * (d * (cfg.num_classes + 1) + d * cfg.num_classes * 4
                                 + (d + d // 8) * (d // 4) + (d // 4) * (cfg.num_attrs + 1))
    f["total"] = sum(f.values())
    return f
