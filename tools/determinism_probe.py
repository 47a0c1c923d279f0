"""Runs the headline batch through FRCNN.forward repeatedly and compares every output bit for bit with the first run
(and the parity numbers against the cfg2x2 golden): a timing-dependent bug in the kernels shows up as a drifting result.
   python tools/determinism_probe.py [mode] [reps] [--stress]   (--stress: host threads hog the CPU so launch timing varies)"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vltk_b200 import synthetic
from vltk_b200.config import FRCNNConfig
from vltk_b200.frcnn import FRCNN

mode = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "exact_tc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 20
stress = "--stress" in sys.argv
cfg = FRCNNConfig().replace(min_size_test=bench.H, max_size_test=bench.W)
sd = synthetic.make_state_dict(cfg, 0)
model = FRCNN.from_pretrained(state_dict=sd, config=cfg, mode=mode, device=0)
host = [b.pin_memory() for b in bench.synthetic_batches(cfg, 2, 0)]
sizes = np.tile(np.array([[bench.H, bench.W]], np.int32), (bench.BATCH, 1))
sizes_t, scales_t = torch.from_numpy(sizes), torch.ones((bench.BATCH, 2))
stop = False
if stress:
    def hog():
        x = 0
        while not stop:
            x += 1
    for _ in range(os.cpu_count() or 8):
        threading.Thread(target=hog, daemon=True).start()
first = None
bad = 0
for r in range(reps):
    if r % 3 == 2:       # another batch in between, so buffers hold other data
        model(host[1], sizes_t, scales_yx=scales_t, padding="max_detections", return_tensors="np")
    d = model(host[0], sizes_t, scales_yx=scales_t, padding="max_detections", return_tensors="np")
    feats = model.debug_read("feats").copy()
    res4 = model.debug_read("res4").copy()
    cur = {k: np.asarray(v).copy() for k, v in d.items()}
    cur["feats_all"], cur["res4"] = feats, res4
    if first is None:
        first = cur
        p = bench.parity_check(model, mode, host[0], sizes_t, scales_t)
        print("parity", {k: p[k] for k in p if k not in ("golden", "rule")}, flush=True)
        continue
    diff = {k: float(np.abs(cur[k].astype(np.float64) - first[k].astype(np.float64)).max()) for k in first if not np.array_equal(cur[k], first[k])}
    if diff:
        bad += 1
        print(f"rep {r}: DIFFERS {diff}", flush=True)
stop = True
print(f"{mode}: {reps} runs, {bad} differ from the first", flush=True)
