"""BASELINE.json configs[3]: the HBM/ALU-bound pieces at exactly config 4's shape — one 800x1333 image:
63 000 anchors -> top 6 000 -> NMS 0.7 -> first 300 -> ROIPool 300 ROIs on the [1,1024,50,84] map —
timed kernel-only with the engine's own per-launch CUDA events inside real forwards (the stage entry
points allocate scratch and synchronise per call, so they cannot give kernel times).

Algorithmic work: decode/top-k 63 000*(4+16) B in + 6 000*20 B out = 1.38 MB;
NMS 6 000*20 B in + 300*20 B out = 126 KB but 18.0 M IoU pairs (latency/ALU-bound: pairs/s is the
meaningful rate); ROIPool = map read + 300*196*1024*esz B write.  Prints one JSON object.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vltk_b200 import synthetic  # noqa: E402
from vltk_b200.config import FRCNNConfig  # noqa: E402
from vltk_b200.frcnn import FRCNN  # noqa: E402

H, W, STEPS = 800, 1333, 20


def gather_bench(hbm):
    """Device-resident feature column (20 000 images x 36 x 2048 f32 = 5.9 GB), shuffled batches of 256 rows
    gathered by vltk_gather_rows_f32: algorithmic bytes = read + write of the gathered rows."""
    from vltk_b200 import _lib
    L = _lib.lib()
    n, width, b = 20000, 36 * 2048, 256
    table = torch.empty((n, width), dtype=torch.float32, device="cuda").normal_()
    g = torch.Generator().manual_seed(0)
    idxs = [torch.randperm(n, generator=g)[:b].int().cuda() for _ in range(8)]
    outb = torch.empty((b, width), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for i in range(3):
        _lib.check(L.vltk_gather_rows_f32(table.data_ptr(), n, width, idxs[i].data_ptr(), b, width, outb.data_ptr(), st), "gather")
    assert torch.equal(outb, table[idxs[2].long()])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        _lib.check(L.vltk_gather_rows_f32(table.data_ptr(), n, width, idxs[i % 8].data_ptr(), b, width, outb.data_ptr(), st), "gather")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    by = 2.0 * b * width * 4
    return {"rows": b, "row_bytes": width * 4, "ms": ms, "algorithmic_bytes": by, "gbs": by / ms / 1e6,
            "frac_of_hbm_peak": by / ms / 1e6 / hbm, "images_per_s": b / ms * 1e3}


def main():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    cfg = FRCNNConfig()
    out = {"shape": "1 x 800x1333 (res4 50x84, 63000 anchors, 6000 pre-NMS, 300 post-NMS)", "hbm_peak_gbs": hbm}
    for mode in ("bf16", "fp32"):
        model = FRCNN.from_pretrained(state_dict=synthetic.make_state_dict(cfg, 0), config=cfg, mode=mode)
        mean = torch.tensor(cfg.pixel_mean).view(1, 3, 1, 1)
        xs = [(synthetic.make_raw_image(H, W, 7000 + i).permute(2, 0, 1).float().unsqueeze(0) - mean).contiguous().cuda() for i in range(4)]
        sizes, scales = np.array([[H, W]], np.int32), np.ones((1, 2), np.float32)
        ro = model.roi_outputs
        for i in range(3):
            t = model.run(xs[i % 4], sizes, scales, ro.max_detections, ro.min_detections, ro.nms_thresh)
        torch.cuda.synchronize()
        model.profile(True)
        for i in range(STEPS):
            t = model.run(xs[i % 4], sizes, scales, ro.max_detections, ro.min_detections, ro.nms_thresh)
        _, csv = model.profile_read(want_csv=True)
        model.profile(False)
        ms, by = {}, {}
        for ln in csv.splitlines():
            k, m_, k_, c_, t_ = ln.split(",")
            ms[k] = ms.get(k, 0.0) + float(t_) / STEPS
            if int(k_) == 0:
                by[k] = float(m_)
        pairs = 6000 * 5999 / 2
        out[mode] = {
            "rpn_select (anchors+top-k sort+decode+clip)": {"ms": ms["rpn_select"], "algorithmic_bytes": by["rpn_select"], "gbs": by["rpn_select"] / ms["rpn_select"] / 1e6},
            "rpn_nms (IoU bitmask + on-device scan, stop at 300)": {"ms": ms["rpn_nms"], "iou_pairs": pairs, "gpairs_per_s": pairs / ms["rpn_nms"] / 1e6, "algorithmic_bytes": by["rpn_nms"]},
            "roi_pool (300 ROIs)": {"ms": ms["roi_pool"], "algorithmic_bytes": by["roi_pool"], "gbs": by["roi_pool"] / ms["roi_pool"] / 1e6, "frac_of_hbm_peak": by["roi_pool"] / ms["roi_pool"] / 1e6 / hbm},
            "mean_rows (14x14 mean)": {"ms": ms["mean_rows"], "algorithmic_bytes": by["mean_rows"], "gbs": by["mean_rows"] / ms["mean_rows"] / 1e6, "frac_of_hbm_peak": by["mean_rows"] / ms["mean_rows"] / 1e6 / hbm},
            "roi_tail": {"ms": ms["roi_tail"]}, "preds": t["preds_per_image"].cpu().tolist(),
        }
        del model
    out["feature_gather (reader, SURVEY 8 f3)"] = gather_bench(hbm)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
