"""Architecture + runtime knobs of the R101-C4 Visual-Genome bottom-up extractor.

One frozen description replaces the ~60 ``cfg.*`` keys the reference reads
(reference: vltk/modeling/frcnn.py:201-223, 1230-1240, 1314-1336, 1367-1372,
1414-1417, 1537-1540, 1583-1607, 1747-1755; vltk/legacy/processing.py:78-94).
The values are the reconstructed config of SURVEY.md Appendix A.
"""
from __future__ import annotations

import copy
import dataclasses
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass
class FRCNNConfig:
    # model / preprocessing (legacy/processing.py:78-94)
    pixel_mean: Tuple[float, float, float] = (102.9801, 115.9465, 122.7717)
    pixel_std: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    min_size_test: int = 800
    max_size_test: int = 1333
    pad_value: float = 0.0
    # backbone (frcnn.py:200-261)
    stem_out_channels: int = 64
    res2_out_channels: int = 256
    blocks_per_stage: Tuple[int, int, int] = (3, 4, 23)  # res2, res3, res4 (depth 101)
    res5_blocks: int = 3
    # RPN (frcnn.py:1406-1673)
    anchor_sizes: Tuple[float, ...] = (32.0, 64.0, 128.0, 256.0, 512.0)
    anchor_ratios: Tuple[float, ...] = (0.5, 1.0, 2.0)
    anchor_stride: int = 16
    rpn_hidden: int = 512
    rpn_nms_thresh: float = 0.7
    rpn_pre_nms_topk: int = 6000
    rpn_post_nms_topk: int = 300
    rpn_min_size: float = 0.0
    rpn_bbox_weights: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)
    # ROI head (frcnn.py:1305-1403, 1676-1740)
    pooler_resolution: int = 14
    num_classes: int = 1600
    num_attrs: int = 400
    roi_bbox_weights: Tuple[float, float, float, float] = (10.0, 10.0, 5.0, 5.0)
    # ROIOutputs knobs (frcnn.py:1229-1240); callers mutate these on the model
    nms_thresh_test: List[float] = field(default_factory=lambda: [0.3])
    score_thresh_test: float = 0.2  # accepted, never used (frcnn.py:116)
    min_detections: int = 36
    max_detections: int = 36

    @property
    def num_anchors(self) -> int:
        return len(self.anchor_sizes) * len(self.anchor_ratios)

    def replace(self, **kw) -> "FRCNNConfig":
        return dataclasses.replace(copy.deepcopy(self), **kw)

    # ---- shape arithmetic shared by host code, oracle and tests -------------
    @staticmethod
    def stem_conv_out(n: int) -> int:  # 7x7 s2 p3
        return (n + 6 - 7) // 2 + 1

    @staticmethod
    def stem_pool_out(n: int) -> int:
        """max_pool2d(k=3, s=2, p=0, ceil_mode=True) (frcnn.py:875-876)."""
        o = -(-(n - 3) // 2) + 1
        if (o - 1) * 2 >= n:  # last window must start inside the input
            o -= 1
        return o

    @staticmethod
    def stride2_out(n: int) -> int:  # 1x1 s2 p0 (stride_in_1x1, frcnn.py:932)
        return (n - 1) // 2 + 1

    def res4_hw(self, h: int, w: int) -> Tuple[int, int]:
        f = lambda n: self.stride2_out(self.stride2_out(self.stem_pool_out(self.stem_conv_out(n))))
        return f(h), f(w)

    def to_reference_dict(self) -> dict:
        """Nested dict accepted by the reference's compat.Config."""
        return {
            "model": {"device": "cpu", "pixel_mean": list(self.pixel_mean),
                      "pixel_std": list(self.pixel_std), "max_pool": True},
            "backbone": {"freeze_at": 2},
            "resnets": {"depth": 101, "norm": "BN", "num_groups": 1, "width_per_group": 64,
                        "out_features": ["res4"], "res2_out_channels": self.res2_out_channels,
                        "res5_dilation": 1, "stem_out_channels": self.stem_out_channels,
                        "stride_in_1x1": True},
            "anchor_generator": {"sizes": [list(self.anchor_sizes)],
                                 "aspect_ratios": [list(self.anchor_ratios)], "offset": 0.0},
            "proposal_generator": {"hidden_channels": self.rpn_hidden, "min_size": self.rpn_min_size},
            "rpn": {"in_features": ["res4"], "nms_thresh": self.rpn_nms_thresh,
                    "pre_nms_topk_test": self.rpn_pre_nms_topk,
                    "post_nms_topk_test": self.rpn_post_nms_topk,
                    "pre_nms_topk_train": 12000, "post_nms_topk_train": 2000,
                    "bbox_reg_weights": list(self.rpn_bbox_weights),
                    "batch_size_per_image": 256, "positive_fraction": 0.5,
                    "smooth_l1_beta": 0.0, "loss_weight": 1.0, "boundary_thresh": -1,
                    "iou_thresholds": [0.3, 0.7], "iou_labels": [0, -1, 1]},
            "roi_heads": {"in_features": ["res4"], "num_classes": self.num_classes,
                          "score_thresh_test": self.score_thresh_test,
                          "nms_thresh_test": list(self.nms_thresh_test),
                          "proposal_append_gt": True, "positive_fraction": 0.25,
                          "iou_thresholds": [0.5], "iou_labels": [0, 1]},
            "roi_box_head": {"pooler_resolution": self.pooler_resolution,
                             "pooler_sampling_ratio": 2,
                             "bbox_reg_weights": list(self.roi_bbox_weights),
                             "cls_agnostic_bbox_reg": False, "smooth_l1_beta": 0.0,
                             "res5halve": False, "attr": True, "num_attrs": self.num_attrs},
            "min_detections": self.min_detections,
            "max_detections": self.max_detections,
            "input": {"min_size_test": self.min_size_test, "max_size_test": self.max_size_test,
                      "format": "BGR"},
            "size_divisibility": 0,
            "pad_value": self.pad_value,
        }
