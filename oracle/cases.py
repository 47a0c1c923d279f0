"""TEST INFRASTRUCTURE — the seeded synthetic parity cases.

A case = config overrides + weight seed + raw-image (h, w, seed) list.  Inputs are
regenerated from the seeds everywhere (here, in CPU tests, on the GPU box); only the
reference's OUTPUTS are committed under tests/golden/.

The image seeds are MARGIN-CERTIFIED (oracle/certify.py, oracle/margins.py): every selection
decision that affects which boxes/ids come out sits >= ~10x fp32 round-off away from its
threshold, so exact-index parity between two fp32 implementations is a meaningful claim.  The
margins of each case are stored in its golden's metadata.
"""
from __future__ import annotations

from vltk_b200.config import FRCNNConfig

# name -> (cfg overrides, weight seed, [(raw_h, raw_w, image seed), ...])
CASES = {
    # CPU-quick: identity resize, reduced proposal counts
    "tiny": (dict(min_size_test=192, max_size_test=256, rpn_pre_nms_topk=600,
                  rpn_post_nms_topk=40, min_detections=12, max_detections=12),
             0, [(192, 256, 1)]),
    # real bilinear resize, two aspect ratios -> bottom/right padding, non-unit scales_yx,
    # nms_thresh list retry, variable preds_per_image
    "mixed": (dict(min_size_test=192, max_size_test=288, rpn_pre_nms_topk=600,
                   rpn_post_nms_topk=48, min_detections=8, max_detections=20,
                   nms_thresh_test=[0.05, 0.3]),
              0, [(150, 200, 2), (240, 160, 3)]),
    # degenerate inputs: a constant image yields FEWER detections than min_detections (the nms_thresh
    # list is exhausted and the last attempt is kept, frcnn.py:1274-1278) -> zero-padded tail with
    # preds_per_image < max_detections; hard-edged stripes exercise saturated activations
    "constant": (dict(min_size_test=192, max_size_test=256, rpn_pre_nms_topk=600,
                      rpn_post_nms_topk=40, min_detections=12, max_detections=12),
                 0, [(192, 256, "const117")]),
    "stripes": (dict(min_size_test=192, max_size_test=256, rpn_pre_nms_topk=600,
                     rpn_post_nms_topk=40, min_detections=12, max_detections=12),
                0, [(192, 256, "stripes16")]),
    # FEWER survivors than the proposal budget: 6x8 res4 cells = 720 anchors -> top 600 -> NMS leaves
    # < 300, so ROI slots past the count are zero-filled and masked all the way to the tail
    "few": (dict(min_size_test=96, max_size_test=128, rpn_pre_nms_topk=600, rpn_post_nms_topk=300,
                 min_detections=10, max_detections=36), 0, [(96, 128, 3005)]),
    # full proposal/detection counts on a mid-size image (6000 -> 300 -> 36)
    "full36": (dict(min_size_test=384, max_size_test=576), 0, [(384, 576, 4)]),
    # BASELINE.json configs[0]: 1 image 800x1333, 36 boxes
    "cfg1": (dict(), 0, [(800, 1333, 6000)]),
    # BASELINE.json configs[1] (2 of its 8 images — the bench runs all 8): 600x1000
    "cfg2x2": (dict(min_size_test=600, max_size_test=1000), 0, [(600, 1000, 4010), (600, 1000, 4011)]),
    # BASELINE.json configs[2] flavour: mixed aspect ratios with padding, max_detections=100
    "cfg3x2": (dict(min_detections=10, max_detections=100), 0, [(600, 800, 20), (1000, 750, 21)]),
    # the `ignorey` branch of find_top_rpn_proposals (frcnn.py:328-366): two caller-given y-ranges (raw-image
    # coordinates) on one resized image (scale 0.78); proposals spanning a range are dropped, others clipped.
    # One image only: the reference's branch overwrites the shared level_ids and cannot run a larger batch.
    "ignorey": (dict(min_size_test=192, max_size_test=288, rpn_pre_nms_topk=600, rpn_post_nms_topk=48,
                     min_detections=8, max_detections=20), 0, [(150, 200, 2)]),
}

# case -> ignorey [N, J, 2] passed to forward (None for every other case)
IGNOREY = {"ignorey": [[[40.0, 70.0], [100.0, 118.0]]]}

CPU_CASES = ("tiny", "mixed", "constant", "stripes", "few", "ignorey")   # cheap enough for the no-GPU suite
GPU_CASES = ("tiny", "mixed", "constant", "stripes", "few", "ignorey", "full36", "cfg1", "cfg2x2", "cfg3x2")


def case_config(name: str) -> FRCNNConfig:
    return FRCNNConfig().replace(**CASES[name][0])


def case_ignorey(name: str):
    """-> ignorey tensor [N,J,2] f32 of the case, or None."""
    import torch
    v = IGNOREY.get(name)
    return None if v is None else torch.tensor(v, dtype=torch.float32)


def case_inputs(name: str):
    """-> (cfg, weight_seed, [raw BGR u8 images])."""
    over, wseed, imgs = CASES[name]
    cfg = FRCNNConfig().replace(**over)
    raws = [raw_image(h, w, s) for (h, w, s) in imgs]
    return cfg, wseed, raws


def raw_image(h: int, w: int, spec):
    """Raw BGR u8 [h,w,3]: an int spec is a seed of the SURVEY §8d noise recipe; a string names a
    deterministic pattern."""
    import torch
    from vltk_b200 import synthetic
    if isinstance(spec, int):
        return synthetic.make_raw_image(h, w, spec)
    if spec.startswith("const"):
        return torch.full((h, w, 3), int(spec[5:]), dtype=torch.uint8)
    if spec.startswith("stripes"):
        p = int(spec[7:])
        return ((torch.arange(h).view(-1, 1, 1) // p % 2) * 200).expand(h, w, 3).to(torch.uint8).contiguous()
    raise ValueError(spec)
