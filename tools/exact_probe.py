"""Bring-up and evidence for the exact_tc mode (csrc/conv_tcx.cu): run on the GPU box.

1. stage check: the fp32-faithful tensor-core conv against an fp64 convolution, next to the CUDA-core fp32
   kernel on the same inputs (error in units of 2^-24 of sum |x||w|, i.e. "fp32 ulps of the reduction");
2. every golden case end to end in exact_tc: are preds_per_image / obj_ids / attr_ids / proposal set exact,
   and how far are res4 / feats from the CUDA-core fp32 mode;
3. img/s of a batch of 8 600x1000 images.
Writes gpurun_out/exact_probe.json.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import cases  # noqa: E402
from tests.util import load_golden, weights  # noqa: E402
from vltk_b200 import stages  # noqa: E402

CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad, dil, residual, relu
    (1, 12, 16, 64, 64, 1, 1, 0, 1, False, True),
    (2, 14, 14, 64, 128, 3, 1, 2, 2, False, True),
    (1, 25, 33, 128, 256, 1, 2, 0, 1, False, False),
    (3, 14, 14, 128, 256, 1, 1, 0, 1, True, True),
    (1, 19, 23, 64, 64, 3, 1, 1, 1, False, True),
    (2, 14, 14, 256, 512, 3, 1, 2, 2, True, True),
    (1, 13, 17, 1024, 512, 3, 1, 1, 1, False, True),
    (5, 14, 14, 512, 2048, 1, 1, 0, 1, True, True),
    (40, 14, 14, 512, 512, 3, 1, 2, 2, False, True),     # res5 conv2, K = 4608, many tiles per CTA
    (60, 14, 14, 2048, 512, 1, 1, 0, 1, False, True),
]


def conv_check(dev):
    out = []
    for case in CONV_CASES:
        n, h, w, cin, cout, k, s, p, d, use_res, relu = case
        g = torch.Generator().manual_seed(abs(hash(case)) % (2 ** 31))
        x = torch.randn(n, h, w, cin, generator=g).clamp_min(-0.5).to(dev)
        wt = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(dev)
        sc = (torch.rand(cout, generator=g) + 0.5).to(dev)
        sh = (torch.randn(cout, generator=g) * 0.1).to(dev)
        oh = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
        ow = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
        res = torch.randn(n, oh, ow, cout, generator=g).to(dev) if use_res else None
        xd, wd = x.permute(0, 3, 1, 2).double(), wt.double()
        y64 = F.conv2d(xd, wd, None, s, p, d)
        mag = F.conv2d(xd.abs(), wd.abs(), None, s, p, d) * sc.double().view(1, -1, 1, 1)     # sum |x||w| * scale
        y64 = y64 * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
        y64, mag = y64.permute(0, 2, 3, 1), mag.permute(0, 2, 3, 1)
        if res is not None:
            y64 = y64 + res.double()
        if relu:
            y64 = F.relu(y64)
        rec = {"case": list(case)}
        for name, kw in (("exact_tc", dict(mode="exact_tc")), ("simt_fp32", dict(mode="fp32"))):
            y = stages.conv2d_nhwc(x, wt, sc, sh, res, s, p, d, relu, **kw).double()
            e = (y - y64).abs()
            rec[name] = {"max_abs": float(e.max()), "max_ulp24_of_sumabs": float((e / (mag + 1e-3)).max() * 2 ** 24),
                         "rms_ulp24_of_sumabs": float(((e / (mag + 1e-3)) ** 2).mean().sqrt() * 2 ** 24),
                         "rms_rel_of_y": float((e.pow(2).mean() / y64.pow(2).mean()).sqrt())}
        if cout % 128 == 0 and not use_res:
            y = stages.conv2d_nhwc(x, wt, sc, sh, None, s, p, d, relu, mode="exact_tc", tensor_cores=2).double()
            e = (y - y64).abs()
            rec["exact_tc_f32out"] = {"max_abs": float(e.max()), "max_ulp24_of_sumabs": float((e / (mag + 1e-3)).max() * 2 ** 24)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    return out


def golden_check(which):
    from vltk_b200.frcnn import FRCNN
    from vltk_b200.preprocess import Preprocess
    res = {}
    models = {}
    for case in which:
        cfg = cases.case_config(case)
        g = load_golden(case)
        per = {}
        outs = {}
        for mode in ("exact_tc", "fp32"):
            key = (cfg.rpn_pre_nms_topk, cfg.rpn_post_nms_topk, mode)
            if key not in models:
                models[key] = FRCNN.from_pretrained(state_dict=weights(cases.CASES[case][1]), config=cfg, mode=mode)
            m = models[key]
            m.roi_outputs.nms_thresh = list(cfg.nms_thresh_test)
            m.roi_outputs.min_detections = cfg.min_detections
            m.roi_outputs.max_detections = cfg.max_detections
            _, _, raws = cases.case_inputs(case)
            ids, images, sizes, scales = Preprocess(cfg)(raws)
            torch.cuda.synchronize()
            t0 = time.time()
            out = m(images, sizes, scales_yx=scales, ignorey=cases.case_ignorey(case))
            torch.cuda.synchronize()
            dt = time.time() - t0
            n = images.shape[0]
            cat = lambda x: torch.cat(list(x)).cpu().numpy()
            tk = m.debug_read("topk_anchor_idx", np.int32).reshape(n, -1)
            outs[mode] = {"res4": m.debug_read("res4"), "feats": m.debug_read("feats"), "head": m.debug_read("rpn_head"),
                          "props": m.debug_read("proposals")}
            per[mode] = {
                "preds_per_image": out["preds_per_image"].tolist() == g["preds_per_image"].tolist(),
                "obj_ids": bool(np.array_equal(cat(out["obj_ids"]), g["obj_ids"])) if out["preds_per_image"].tolist() == g["preds_per_image"].tolist() else False,
                "attr_ids": bool(np.array_equal(cat(out["attr_ids"]), g["attr_ids"])) if out["preds_per_image"].tolist() == g["preds_per_image"].tolist() else False,
                "n_props": m.debug_read("proposal_count", np.int32).tolist() == g["n_props"].tolist(),
                "topk_set": all(set(tk[i].tolist()) == set(g["rpn_topk_anchor_idx"][i].tolist()) for i in range(n)),
                "first_call_s": dt,
            }
            if per[mode]["preds_per_image"]:
                per[mode]["boxes_max_abs"] = float(np.abs(cat(out["boxes"]) - g["boxes"]).max())
                s = int(g["roi_features_stride"])
                a, b = cat(out["roi_features"])[:, ::s], g["roi_features"]
                per[mode]["roi_features_max_rel"] = float((np.abs(a - b) / (np.abs(b) + 1e-3)).max())
                per[mode]["obj_probs_max_abs"] = float(np.abs(cat(out["obj_probs"]) - g["obj_probs"]).max())
        a, b = outs["exact_tc"], outs["fp32"]
        per["res4_rel_rms_vs_fp32"] = float(np.sqrt(((a["res4"] - b["res4"]) ** 2).mean() / (b["res4"] ** 2).mean()))
        per["res4_max_abs_vs_fp32"] = float(np.abs(a["res4"] - b["res4"]).max())
        n75 = 75
        ha = a["head"].reshape(-1, a["head"].size // (b["head"].size // 76))[:, :n75]
        hb = b["head"].reshape(-1, 76)[:, :n75]
        per["rpn_head_max_abs_vs_fp32"] = float(np.abs(ha - hb).max())
        per["rpn_logit_rel_rms_vs_fp32"] = float(np.sqrt(((ha[:, 60:] - hb[:, 60:]) ** 2).mean() / (hb[:, 60:] ** 2).mean()))
        if a["feats"].shape == b["feats"].shape and np.abs(a["props"] - b["props"]).max() < 1e-2:
            per["feats_rel_rms_vs_fp32"] = float(np.sqrt(((a["feats"] - b["feats"]) ** 2).mean() / (b["feats"] ** 2).mean()))
        print(case, json.dumps(per), flush=True)
        res[case] = per
    return res


def speed(n_img=8, iters=5):
    from vltk_b200.frcnn import FRCNN
    from vltk_b200.config import FRCNNConfig
    cfg = FRCNNConfig().replace(min_size_test=600, max_size_test=1000)
    out = {}
    g = torch.Generator().manual_seed(5)
    images = (torch.randn(n_img, 3, 600, 1000, generator=g) * 40).cuda()
    sizes = torch.tensor([[600, 1000]] * n_img)
    for mode in ("exact_tc", "bf16"):
        m = FRCNN.from_pretrained(state_dict=weights(0), config=cfg, mode=mode)
        for _ in range(2):
            m(images, sizes)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(iters):
            m(images, sizes)
        torch.cuda.synchronize()
        out[mode] = {"img_per_s": n_img * iters / (time.time() - t0)}
        print(mode, out[mode], flush=True)
        del m
    return out


def main():
    dev = torch.device("cuda:0")
    what = sys.argv[1:] or ["conv", "golden", "speed"]
    res = {}
    if "conv" in what:
        res["conv"] = conv_check(dev)
    if "golden_small" in what:
        res["golden"] = golden_check(["tiny", "mixed", "few"])
    if "golden" in what:
        res["golden"] = golden_check(list(cases.GPU_CASES))
    if "speed" in what:
        res["speed"] = speed()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/exact_probe.json", "w"), indent=1)


if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    main()
