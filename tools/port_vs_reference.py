"""Authoring-container check (needs /root/reference): how far is the TIMED oracle port (torchvision nms / RoIPool,
res5 in one pass) from the real reference's FRCNN.forward on the same image and threads?  bench.py --impl reference
runs the port on the GPU box (the Python reference cannot travel); this states the gap.  Writes profiles/r02_port_vs_reference.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import frcnn_oracle as O, ref_loader  # noqa: E402
from vltk_b200 import synthetic  # noqa: E402
from vltk_b200.config import FRCNNConfig  # noqa: E402


def main():
    H, W = 600, 1000
    threads = int(os.environ.get("THREADS", os.cpu_count()))
    torch.set_num_threads(threads)
    cfg = FRCNNConfig().replace(min_size_test=H, max_size_test=W)
    sd = synthetic.make_state_dict(cfg, 0)
    ref = ref_loader.build_reference_model(cfg, sd)
    ref.roi_outputs.nms_thresh = list(cfg.nms_thresh_test)
    ref.roi_outputs.min_detections = cfg.min_detections
    ref.roi_outputs.max_detections = cfg.max_detections
    O.use_torchvision_ops(True)
    t_ref, t_port = [], []
    for i in range(4):
        raw = synthetic.make_raw_image(H, W, 900 + i)
        imgs, sizes, scales = O.preprocess(cfg, [raw])
        t0 = time.time()
        with torch.no_grad():
            r = ref(imgs, torch.tensor(sizes), scales_yx=torch.as_tensor(scales))
        t1 = time.time()
        o = O.forward(sd, cfg, imgs, sizes, scales, res5_chunk=1 << 30)
        t2 = time.time()
        same = torch.equal(r["obj_ids"][0], o["obj_ids"][0])
        if i:
            t_ref.append(t1 - t0)
            t_port.append(t2 - t1)
        print(i, f"reference {t1 - t0:.2f}s port {t2 - t1:.2f}s ids equal {same}", flush=True)
    res = {"threads": threads, "image": [H, W], "reference_s_per_img": sum(t_ref) / len(t_ref), "port_s_per_img": sum(t_port) / len(t_port),
           "port_over_reference": sum(t_port) / sum(t_ref)}
    print(json.dumps(res))
    json.dump(res, open("profiles/r02_port_vs_reference.json", "w"), indent=1)


if __name__ == "__main__":
    main()
