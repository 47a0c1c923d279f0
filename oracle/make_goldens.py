"""TEST INFRASTRUCTURE — produces tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference) on the seeded cases of oracle/cases.py, and the class-bias
calibration vector vltk_b200/data/cls_bias_seed*.npy.  Run in the authoring container:

    python oracle/make_goldens.py --calibrate 0        # once per weight seed
    python oracle/make_goldens.py [case ...]           # default: all cases

The reference is driven through its own modules (FRCNN.backbone, .proposal_generator,
.roi_heads, .roi_outputs and legacy Preprocess), stage by stage, so that stage tensors
can be committed for teacher-forced parity tests.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, frcnn_oracle as O, margins, ref_loader  # noqa: E402
from vltk_b200 import synthetic  # noqa: E402
from vltk_b200.config import FRCNNConfig  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "vltk_b200", "data")


def calibrate(seed: int):
    """cls_score.bias = -W . mean(pooled features) on one calibration image, so the class
    posteriors are not dominated by one class."""
    cfg = FRCNNConfig().replace(min_size_test=256, max_size_test=384, rpn_post_nms_topk=100)
    sd = synthetic.make_state_dict(cfg, seed, cls_bias=None)
    raw = synthetic.make_raw_image(256, 384, 99)
    imgs, sizes, scales = O.preprocess(cfg, [raw])
    st = {}
    O.forward(sd, cfg, imgs, sizes, scales, stages=st)
    w = sd["roi_heads.box_predictor.cls_score.weight"]
    bias = -(w.double() @ st["feats"].double().mean(0)).float()
    os.makedirs(DATA, exist_ok=True)
    np.save(os.path.join(DATA, f"cls_bias_seed{seed}.npy"), bias.numpy())
    print(f"calibrated seed {seed}: |bias| mean {bias.abs().mean():.3f}")


def checksum(t: torch.Tensor):
    t = t.double()
    return np.array([t.sum().item(), t.abs().sum().item(), float(t.numel())])


@torch.no_grad()
def run_reference(name: str):
    cfg, wseed, raws = cases.case_inputs(name)
    sd = synthetic.make_state_dict(cfg, wseed)
    model = ref_loader.build_reference_model(cfg, sd)
    frcnn, compat = ref_loader.load_reference()
    Preprocess = ref_loader.load_preprocess()
    pre = Preprocess(compat.Config(cfg.to_reference_dict()))
    t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(0)  # ResizeShortestEdge draws randint(min,min+1) (processing.py:44)
        ids, images, sizes, scales = pre([r.clone() for r in raws], list(range(len(raws))))
        feats = model.backbone(images)
        res4 = feats["res4"]
        logits, deltas = model.proposal_generator.rpn_head([res4])
        ign = cases.case_ignorey(name)
        boxes, plogits = model.proposal_generator(images, sizes, feats, None, ign, scales)
        obj_logits, attr_logits, box_deltas, pooled = model.roi_heads(feats, boxes, None)
        # end-to-end through the public call as well (must equal the staged run)
        out = model(images, sizes, scales_yx=scales, ignorey=ign)
    dt = time.time() - t0
    g = {
        "images_ck": checksum(images), "sizes": sizes.numpy(), "scales_yx": scales.numpy(),
        "res4_ck": checksum(res4), "res4_sub": res4[:, ::16].numpy(),
        "rpn_logits": logits[0].numpy(), "rpn_deltas_ck": checksum(deltas[0]),
        "rpn_deltas_sub": deltas[0][:, :, ::2, ::2].numpy(),
        "n_props": np.array([len(b) for b in boxes]),
        "proposals": torch.cat(boxes).numpy(), "proposal_logits": torch.cat(plogits).numpy(),
        "feats_ck": checksum(pooled), "feats_sub": pooled[:, ::8].numpy(),
        "obj_argmax_all": obj_logits.argmax(-1).numpy(),
        "obj_fg_argmax_all": obj_logits[:, :-1].argmax(-1).numpy(),
        "obj_lse": torch.logsumexp(obj_logits, -1).numpy(),
        "obj_logits_sub": obj_logits[:, ::16].numpy(),
        "attr_logits_sub": attr_logits[:, ::8].numpy(),
        "box_deltas_sub": box_deltas[:, ::64].numpy(),
        "preds_per_image": out["preds_per_image"].numpy(),
        "boxes": torch.cat(out["boxes"]).numpy(),
        "obj_ids": torch.cat(out["obj_ids"]).numpy(),
        "obj_probs": torch.cat(out["obj_probs"]).numpy(),
        "attr_ids": torch.cat(out["attr_ids"]).numpy(),
        "attr_probs": torch.cat(out["attr_probs"]).numpy(),
    }
    rf = torch.cat(out["roi_features"])
    g["roi_features_ck"] = checksum(rf)
    g["roi_features"] = rf.numpy() if rf.numel() <= 40 * 2048 else rf[:, ::8].numpy()
    g["roi_features_stride"] = np.array(1 if rf.numel() <= 40 * 2048 else 8)
    if name == "tiny":  # full stage tensors only where they are small
        g["res4"] = res4.numpy()
        g["rpn_deltas"] = deltas[0].numpy()
        g["feats"] = pooled.numpy()
        g["obj_logits"] = obj_logits.numpy()
        g["attr_logits"] = attr_logits.numpy()
        g["box_deltas_f16"] = box_deltas.numpy().astype(np.float16)
    # decision margins of this case (computed with the oracle port on the same inputs)
    ost = {}
    oout = O.forward(sd, cfg, images, sizes, scales, stages=ost, ignorey=ign)
    mg = margins.margins(cfg, ost, oout)
    assert margins.certified(mg), f"{name}: seeds are not margin-certified: {mg}"
    g["rpn_topk_anchor_idx"] = torch.stack([d["topk_idx"] for d in ost["rpn_debug"]]).numpy().astype(np.int32)
    meta = {"case": name, "overrides": cases.CASES[name][0], "weight_seed": wseed, "margins": mg,
            "margin_thresholds": margins.THRESHOLDS,
            "images": cases.CASES[name][2], "torch": torch.__version__,
            "reference_seconds": round(dt, 2), "threads": torch.get_num_threads()}
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), meta=json.dumps(meta), **g)
    sz = os.path.getsize(os.path.join(GOLD, f"{name}.npz")) / 1e6
    print(f"{name}: reference {dt:.1f}s, preds {g['preds_per_image'].tolist()}, "
          f"props {g['n_props'].tolist()}, distinct obj_ids {len(set(g['obj_ids'].tolist()))}, "
          f"{sz:.2f} MB")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--calibrate", type=int, default=None)
    ap.add_argument("cases", nargs="*")
    a = ap.parse_args()
    if a.calibrate is not None:
        calibrate(a.calibrate)
    else:
        for c in (a.cases or list(cases.CASES)):
            run_reference(c)
