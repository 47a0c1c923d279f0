"""BASELINE.json configs[4] (scaled by --images): a COCO-style extraction sweep through the public
driver `vltk_b200.extract.extract` — mixed aspect-ratio synthetic images (SURVEY.md §8d config-3 size
set), batch 8, images sharded by index over the ranks, Arrow feature write.

    python tools/sweep.py --images 512                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sweep.py --images 1024 [--single-file]

Times the WHOLE job on the wall clock after a barrier (host u8 images or JPEG bytes -> H2D -> [GPU JPEG decode]
-> fused preprocess -> forward -> D2H -> Arrow IPC write on a background thread), max over ranks; rank 0 prints
one JSON line.  Raw images come from
a small pre-generated pool (generating 5000 distinct noise images on the host would dominate)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZES = [(800, 1067), (800, 1333), (1067, 800), (1333, 800), (600, 1000), (800, 800), (704, 1333), (800, 1200)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--single-file", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--source", default="raw", choices=["raw", "jpeg"], help="raw: decoded u8 BGR images in pinned host memory; jpeg: encoded JPEG bytes (GPU decode)")
    ap.add_argument("--no-bucket", action="store_true", help="batch in index order (pads every batch to the mixed-aspect maximum)")
    a = ap.parse_args()
    import torch.distributed as dist
    from vltk_b200 import synthetic
    from vltk_b200.config import FRCNNConfig
    from vltk_b200.extract import extract, read_arrow
    from vltk_b200.frcnn import FRCNN
    from vltk_b200.preprocess import Preprocess
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = FRCNNConfig()
    model = FRCNN.from_pretrained(state_dict=synthetic.make_state_dict(cfg, 0), config=cfg, mode=a.mode, device=local)
    pre = Preprocess(cfg, device=local)
    pool = [synthetic.make_raw_image(h, w, 3000 + i).pin_memory() for i, (h, w) in enumerate(SIZES)]
    if a.source == "jpeg":
        import cv2
        jp = [cv2.imencode(".jpg", p.numpy(), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for p in pool]
        source = lambda i: jp[i % len(jp)]  # noqa: E731
    else:
        source = lambda i: pool[i % len(pool)]  # noqa: E731
    ids = [f"img{i:06d}" for i in range(a.images)]
    out_dir = a.out or tempfile.mkdtemp(prefix="vltk_sweep_")
    extract(source, ids[: 8 * a.batch * world], model, pre, os.path.join(out_dir, "warm"), batch_size=a.batch, rank=rank, world=world,
            bucket=not a.no_bucket)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    path = extract(source, ids, model, pre, out_dir, split="train", batch_size=a.batch, rank=rank, world=world,
                   single_file=a.single_file, meta={"dataset": "synthetic-sweep", "model_config": {"mode": a.mode}},
                   bucket=not a.no_bucket)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        rows = None
        if path and os.path.exists(path):
            table, meta = read_arrow(path)
            rows = table.num_rows
        print(json.dumps({"workload": "configs[4] sweep: mixed-aspect synthetic images, batch 8, Arrow write",
                          "images": a.images, "n_gpus": world, "mode": a.mode, "source": a.source, "bucketed": not a.no_bucket, "seconds": dt, "images_per_sec": a.images / dt,
                          "single_file": a.single_file, "rank0_file": path, "rank0_rows": rows,
                          "file_mb": round(os.path.getsize(path) / 1e6, 1) if path and os.path.exists(path) else None}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
