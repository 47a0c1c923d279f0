#!/usr/bin/env python
"""Headline benchmark: images/sec of FRCNN R101-C4 VG region-feature extraction, 36 boxes/image.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode bf16|fp32] [--impl reference]

A step = one pass of the hot path over one batch: BASELINE.json configs[1] — 8 synthetic
600x1000 images per GPU, random-init (engineered, seeded) R101-C4 VG weights, 300 proposals ->
36 detections per image.  One process per GPU; images shard by rank with no collective in the
data path (weak scaling: every rank runs its own batches).  Rank 0 prints ONE JSON line.

  value        images/s with the normalised batch already resident in HBM, CUDA-event timed; --streams batches are kept in
               flight (default 2: batch i on stream i % 2 with its own workspace, outputs bit-identical to one at a time)
  e2e          same metric through the public streaming call FRCNN.forward_stream() with pinned HOST tensors in and HOST
               arrays out (H2D / D2H inside the timed region); the synchronous FRCNN.forward() is timed beside it
  roofline     tcgen05 implicit-GEMM kernels: algorithmic conv FLOPs / busy time (union of their launches' CUDA-event
               intervals on their own streams)
  cpu_baseline the oracle port (CPU restatement of the reference) on a 1-image sample
  --impl reference : times that CPU implementation on all host threads instead (1 image/step)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec FRCNN region-feature extraction (36 boxes) at 1/2/4/8 B200 vs host-CPU ref"
BATCH, H, W = 8, 600, 1000
WORKLOAD = "configs[1]: batch 8 synthetic 600x1000 images/GPU, 36 boxes/image, 300 proposals, R101-C4 VG random-init"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        try:  # one long-lived nvidia-smi sampling every 50 ms (spawning one per sample takes ~150 ms each)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for ln in self.proc.stdout:
                if self.stop_flag:
                    break
                parts = [x.strip() for x in ln.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        p = getattr(self, "proc", None)
        if p is not None:  # the exact child this object started
            p.terminate()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.samples)}


def synthetic_batches(cfg, n_batches, seed0):
    """n_batches distinct [8,3,600,1000] normalised batches via the oracle-free host recipe
    (raw == target size, so Preprocess reduces to mean subtraction)."""
    import torch
    from vltk_b200 import synthetic
    mean = torch.tensor(cfg.pixel_mean).view(1, 3, 1, 1)
    out = []
    for b in range(n_batches):
        raws = [synthetic.make_raw_image(H, W, seed0 + b * BATCH + i) for i in range(BATCH)]
        x = torch.stack([r.permute(2, 0, 1).float() for r in raws]) - mean
        out.append(x.contiguous())
    return out


def cpu_oracle_images_per_sec(cfg, sd, seconds_budget=30.0, timed=3, threads=None):
    """Times the oracle port (CPU restatement of the reference) on single 600x1000 images: one warm-up,
    then up to `timed` runs within the budget; reports the mean."""
    import torch
    from oracle import frcnn_oracle as O
    from vltk_b200 import synthetic
    torch.set_num_threads(threads or os.cpu_count() or 1)
    times = []
    t_start = time.time()
    for i in range(timed + 1):
        raw = synthetic.make_raw_image(H, W, 900 + i)
        imgs, sizes, scales = O.preprocess(cfg, [raw])
        t0 = time.time()
        O.forward(sd, cfg, imgs, sizes, scales)
        if i > 0:
            times.append(time.time() - t0)
        if times and time.time() - t_start > seconds_budget:
            break
    return len(times) / sum(times), len(times), torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port — the
    Python reference cannot travel to the GPU box) on all host threads, 1 image per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import frcnn_oracle as O
    from vltk_b200 import synthetic
    from vltk_b200.config import FRCNNConfig
    cfg = FRCNNConfig().replace(min_size_test=H, max_size_test=W)
    sd = synthetic.make_state_dict(cfg, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    budget = 240.0
    done, t_total = 0, 0.0
    t_begin = time.time()
    for s in range(args.warmup + args.steps):
        raw = synthetic.make_raw_image(H, W, 900 + s)
        imgs, sizes, scales = O.preprocess(cfg, [raw])
        t0 = time.time()
        O.forward(sd, cfg, imgs, sizes, scales)
        dt = time.time() - t0
        if s >= min(args.warmup, 1):  # CPU: one warm-up pass is enough to page everything in
            done += 1
            t_total += dt
        if time.time() - t_begin > budget and done >= 1:
            break
    v = done / t_total
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/sec", "n_gpus": args.gpus,
            "steps": args.steps, "steps_executed": done, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / done,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "1 image per step"},
            "cpu_baseline": {"value": v, "unit": "images/sec", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{done} single 600x1000 images through oracle/frcnn_oracle.py (torch fp32 CPU)"},
            "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("VLTK_BENCH_MODE", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=2, help="batches kept in flight on separate CUDA streams (each with its own workspace): one batch's few-CTA selection kernels overlap the other's convolutions")
    ap.add_argument("--profile-csv", default=None, help="write per-launch conv timings here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from vltk_b200 import arch, synthetic
    from vltk_b200.config import FRCNNConfig
    from vltk_b200.frcnn import FRCNN

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` relaunches itself one-process-per-GPU (the driver
        # launches it under torch.distributed.run directly, which skips this branch)
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(args.warmup, 3)

    cfg = FRCNNConfig().replace(min_size_test=H, max_size_test=W)
    sd = synthetic.make_state_dict(cfg, 0)
    model = FRCNN.from_pretrained(state_dict=sd, config=cfg, mode=args.mode, device=local)
    n_rot = 4  # rotate distinct input batches: 4 x 57.6 MB > 126 MB L2 (activations are GBs anyway)
    host = [b.pin_memory() for b in synthetic_batches(cfg, n_rot, seed0=10000 * rank)]
    devb = [b.to(dev) for b in host]
    sizes = np.tile(np.array([[H, W]], np.int32), (BATCH, 1))
    scales = np.ones((BATCH, 2), np.float32)
    ro = model.roi_outputs

    side = [torch.cuda.Stream(device=dev) for _ in range(max(args.streams - 1, 0))]

    def step_resident(i):
        if args.streams <= 1:
            return model.run(devb[i % n_rot], sizes, scales, ro.max_detections, ro.min_detections, ro.nms_thresh)
        k = i % args.streams          # batch i runs on stream k with its own workspace slot
        if k == 0:
            return model.run(devb[i % n_rot], sizes, scales, ro.max_detections, ro.min_detections, ro.nms_thresh, slot=0)
        with torch.cuda.stream(side[k - 1]):
            return model.run(devb[i % n_rot], sizes, scales, ro.max_detections, ro.min_detections, ro.nms_thresh, slot=k)

    def join_streams():
        cur = torch.cuda.current_stream(dev)
        for s_ in side:
            cur.wait_stream(s_)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    for i in range(W_ * max(args.streams, 1)):
        t = step_resident(i)
    join_streams()
    barrier()
    preds = t["preds_per_image"].cpu().tolist()
    l0 = model.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)  # let the sampler attach before the timed region
    model.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    for s_ in side:
        s_.wait_stream(torch.cuda.current_stream(dev))
    e0.record()
    for i in range(args.steps):
        t = step_resident(W_ + i)
    join_streams()                    # the timed region ends when EVERY stream has drained
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = model.launch_count() - l0
    prof, csv = model.profile_read(want_csv=True)
    stage_ms, stage_bytes = {}, {}
    for ln in (csv or "").splitlines():   # live per-kernel-class breakdown of the timed region
        kind, m_, k_, c_, ms_ = ln.split(",")
        if kind == "tcgen05":             # overlapping launches: reported from the interval union below
            continue
        stage_ms[kind] = stage_ms.get(kind, 0.0) + float(ms_) / args.steps
        if int(k_) == 0:                  # non-GEMM kernels log their algorithmic bytes in column 2
            stage_bytes[kind] = stage_bytes.get(kind, 0.0) + float(m_) / args.steps
    model.profile(False)
    if sampler:
        sampler.stop()
        sampler.join(timeout=3)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = world * BATCH * args.steps / (ms / 1e3)

    # ---------------- end-to-end through the public API, host buffers (e2e) ----------------
    # Every step: that step's batch is copied from pinned host memory, run, and its result dict read back to
    # host numpy arrays.  Headline = the streaming call FRCNN.forward_stream (copies of neighbouring batches
    # overlap the compute); the plain synchronous FRCNN.forward is timed beside it.
    sizes_t, scales_t = torch.from_numpy(sizes.astype(np.int64)), torch.from_numpy(scales)

    def e2e_batches(k):
        for i in range(k):
            yield host[i % n_rot], sizes_t, scales_t

    for o in model.forward_stream(e2e_batches(3)):
        pass
    barrier()
    t0 = time.perf_counter()
    n_out = 0
    for o in model.forward_stream(e2e_batches(args.steps)):
        n_out += int(o["roi_features"].shape[0])
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3     # wall clock: the last result is on the host
    assert n_out == BATCH * args.steps

    def step_sync(i):
        return model(host[i % n_rot], sizes_t, scales_yx=scales_t, padding="max_detections", return_tensors="np")
    for i in range(2):
        step_sync(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_sync(i)
    torch.cuda.synchronize()
    ms_sync = (time.perf_counter() - t0) * 1e3
    if world > 1:
        tt = torch.tensor([ms_e2e, ms_sync], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e, ms_sync = float(tt[0].item()), float(tt[1].item())
    # ---------------- encoded-bytes front door: JPEG bytes in host memory -> features on the host ----------------
    jpeg_info = None
    try:
        import cv2
        from vltk_b200.preprocess import Preprocess
        pre = Preprocess(cfg, device=local)
        jb = []
        for b in range(n_rot):    # same synthetic images, cv2-encoded at quality 90 (4:2:0), kept as bytes
            jb.append([cv2.imencode(".jpg", synthetic.make_raw_image(H, W, 10000 * rank + 100 * b + j).numpy(),
                                    [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for j in range(BATCH)])

        def jpeg_batches(k):
            for i in range(k):
                yield jb[i % n_rot]
        for o in model.forward_jpeg_stream(jpeg_batches(8), pre, group=8):
            pass
        barrier()
        t0 = time.perf_counter()
        n_out = 0
        for o in model.forward_jpeg_stream(jpeg_batches(args.steps), pre, group=8):
            n_out += int(o["roi_features"].shape[0])
        torch.cuda.synchronize()
        ms_jpeg = (time.perf_counter() - t0) * 1e3
        assert n_out == BATCH * args.steps
        if world > 1:
            tt = torch.tensor([ms_jpeg], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_jpeg = float(tt.item())
        jpeg_info = {"value": world * BATCH * args.steps / (ms_jpeg / 1e3), "unit": "images/sec", "ms_per_step": ms_jpeg / args.steps,
                     "h2d_bytes_per_step": int(sum(len(d) for d in jb[0])),
                     "api": "FRCNN.forward_jpeg_stream(lists of 8 JPEG byte strings, 600x1000 q90 4:2:0) -> numpy dicts: host parses markers and strips byte stuffing, GPU does Huffman + IDCT + colour + resize/normalise/pad + the model; 8 batches decoded per front-end call",
                     "reference_equivalent": "cv2.imread on the host (vltk/compat.py:573-579) + Preprocess + forward"}
    except ImportError:
        pass
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)
    sync_value = world * BATCH * args.steps / (ms_sync / 1e3)
    h2d = host[0].numel() * 4
    d2h = int(sum(v.nbytes for k, v in o.items() if k != "sizes"))

    if rank == 0:
        pk, pk_src = peaks()
        fl = arch.flops_per_image(cfg, H, W, cfg.rpn_post_nms_topk)
        tc_ms, tc_fl, tc_n = prof["tcgen05"]
        if tc_n:
            stage_ms["tcgen05"] = tc_ms / args.steps
        si_ms, si_fl, si_n = prof["simt"]
        dom = "tcgen05" if tc_n else "simt"
        d_ms, d_fl, d_n = prof[dom]
        achieved = d_fl / (d_ms / 1e3) / 1e12 if d_ms else 0.0
        peak = pk["bf16_tflops_sustained"] if "bf16_tflops_sustained" in pk else pk["bf16_tflops"]
        traffic, traffic_note = None, None
        ncu_json = os.path.join(ROOT, "profiles", "r01_conv_tc_traffic_v19.json")
        if dom == "tcgen05" and os.path.exists(ncu_json):  # dram bytes/launch from the committed ncu --set full capture
            nj = json.load(open(ncu_json))
            traffic = nj["traffic_bytes_per_launch"]
            traffic_note = (f"dram__bytes_read+write per launch, avg of {len(nj['launches'])} res5 launches in {os.path.basename(ncu_json)} "
                            f"(algorithmic {nj['algorithmic_bytes_per_launch']:.4g} B)")
        line = {
            "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": W_, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * BATCH, "mode": args.mode, "streams_in_flight": args.streams,
                       "value_timed_with_profiling_events": "the K timed steps carry ~250 cudaEventRecords/step (per-launch roofline timing); e2e loops do not",
                       "arithmetic": ("bf16 operands / fp32 accumulate on tcgen05 (stem, res2-res5, RPN 3x3); predictor linears fp32-faithful on tcgen05 (3-pass split-bf16, fp32 logits); RPN 1x1 head and the whole selection tail in fp32"
                                      if args.mode == "bf16" else "fp32 FMA (CUDA cores), index-exact parity mode"),
                       "parallelism": f"images sharded by rank, dp{world}, no data-path collective",
                       "l2": "4 rotating input batches (230 MB) and multi-GB activations exceed the 126 MB L2",
                       "preds_per_image": preds},
            "e2e": {"value": e2e_value, "unit": "images/sec", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "FRCNN.forward_stream(batches of host pinned f32 [8,3,600,1000]) -> numpy dicts; H2D/D2H of neighbouring batches overlap compute",
                    "timing": "wall clock over the K steps, last result on the host, max over ranks",
                    "sync_forward_value": sync_value, "sync_forward_ms_per_step": ms_sync / args.steps,
                    "sync_api": "FRCNN.forward(host pinned f32, padding='max_detections', return_tensors='np'), one call per step"},
            "e2e_jpeg": jpeg_info,
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "conv_tc3_kernel (CTA pairs, tcgen05 cta_group::2) + conv_tc2_kernel: tcgen05/TMEM implicit GEMM, TMA im2col" if dom == "tcgen05" else "conv_simt_kernel",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": pk_src + (", sustained bf16" if "bf16_tflops_sustained" in pk else ""),
                         "timing": "CUDA events around every launch on its own stream; busy time = union of the launch intervals (res2-res4 run as two image halves on two streams, overlapped time counted once)",
                         "launches_per_step": d_n / args.steps, "ms_per_step": d_ms / args.steps,
                         "flops_per_step": d_fl / args.steps,
                         "share_of_step": (d_ms / args.steps) / (ms / args.steps),
                         "other_dense_ms_per_step": (si_ms if dom == "tcgen05" else tc_ms) / args.steps},
            "stage_ms_per_step": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])},
            "hbm_kernels": {k: {"ms_per_step": round(stage_ms[k], 4), "algorithmic_bytes_per_step": b,
                                "gbs": b / (stage_ms[k] / 1e3) / 1e9, "frac_of_hbm_peak": b / (stage_ms[k] / 1e3) / 1e9 / pk["hbm_gbs"]}
                            for k, b in stage_bytes.items() if stage_ms.get(k)},
            "flops_per_image": fl["total"], "step_tflops": BATCH * fl["total"] / (ms / args.steps / 1e3) / 1e12,
            "clocks": sampler.summary() if sampler else None,
        }
        if not args.no_cpu_baseline and world == 1:    # reported baseline: rank 0 at N=1 only (the other N repeat the same CPU number)
            v, nimg, cores = cpu_oracle_images_per_sec(cfg, sd)
            line["cpu_baseline"] = {"value": v, "unit": "images/sec", "cores": cores, "kind": "port",
                                    "sample": f"mean of {nimg} single 600x1000 images after 1 warm-up, oracle/frcnn_oracle.py (torch fp32 CPU, all host threads)"}
        if args.profile_csv and csv:
            with open(args.profile_csv, "w") as f:
                f.write("kind,M,K,Cout,ms\n" + csv)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
