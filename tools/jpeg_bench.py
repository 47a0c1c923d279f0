"""JPEG front end throughput on BASELINE.json configs[1]'s image size: batches of 8 synthetic
600x1000 JPEGs (cv2-encoded, 4:2:0) -> BGR u8 device images, bit-checked against cv2.imdecode.  Reports, per
quality: host time per image of (a) cv2.imdecode (what the reference pays, one core), (b) the C++ host entropy
decoder, (c) the marker-parse + destuff preparation of the GPU entropy path; and device time per batch (CUDA
events) of the GPU entropy decoder and of IDCT + colour.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys
import time

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vltk_b200 import _lib, jpeg, synthetic  # noqa: E402

N, H, W = 8, 600, 1000


def main():
    out = {"batch": N, "size": [H, W], "host_cores": os.cpu_count()}
    cv2.setNumThreads(1)
    raws = [synthetic.make_raw_image(H, W, 900 + i).numpy() for i in range(N)]
    # a smoother, photo-like variant as well: the noise recipe compresses ~2x worse than natural images
    smooth = [cv2.GaussianBlur(r, (0, 0), 2.0) for r in raws]
    for label, imgs, q in (("noise_q90", raws, 90), ("noise_q75", raws, 75), ("smooth_q90", smooth, 90)):
        datas = [cv2.imencode(".jpg", r, [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes() for r in imgs]
        arrs = [np.frombuffer(d, np.uint8) for d in datas]
        t0 = time.perf_counter()
        for _ in range(3):
            refs = [cv2.imdecode(a, cv2.IMREAD_COLOR) for a in arrs]
        t_cv2 = (time.perf_counter() - t0) / (3 * N) * 1e3
        L = _lib.lib()
        buf = np.empty(int(jpeg.parse(datas[0]).coef_count), np.int16)
        t0 = time.perf_counter()
        for _ in range(3):
            for d in datas:
                L.vltk_jpeg_decode_coefficients(d, len(d), buf.ctypes.data, buf.size)
        t_host = (time.perf_counter() - t0) / (3 * N) * 1e3
        res = {"bytes_per_image": int(np.mean([len(d) for d in datas])), "cv2_imdecode_ms_per_image_1core": t_cv2,
               "host_entropy_ms_per_image_1core": t_host}
        for mode in ("gpu", "host"):
            dec = jpeg.JpegDecoder(entropy=mode, threads=min(8, os.cpu_count() or 1))
            for _ in range(3):
                outs = dec.decode(datas)
            torch.cuda.synchronize()
            for o, r in zip(outs, refs):
                assert np.array_equal(o.cpu().numpy(), r)
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 10
            for _ in range(reps):
                outs = dec.decode(datas)
            e1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / reps * 1e3
            res[f"{mode}_entropy"] = {"wall_ms_per_batch": wall, "images_per_s": N / wall * 1e3,
                                      "device_span_ms_per_batch": e0.elapsed_time(e1) / reps}
            if mode == "gpu":
                res[f"{mode}_entropy"]["sync_iterations"] = dec.last_iterations.cpu().tolist()
                # stage split: preparation on the host, entropy kernels, reconstruction
                n = len(datas)
                arr_d = (C.c_char_p * n)(*datas)
                arr_l = (C.c_size_t * n)(*[len(d) for d in datas])
                cap = int(L.vltk_jpeg_gpu_blob_bound(n, arr_l))
                pin = torch.empty(cap, dtype=torch.uint8).pin_memory()
                infos = (jpeg.JpegInfo * n)()
                used, tot = C.c_size_t(0), C.c_int64(0)
                offs, on_gpu = (C.c_int64 * n)(), (C.c_int * n)()
                t0 = time.perf_counter()
                for _ in range(10):
                    L.vltk_jpeg_gpu_prepare_batch(n, arr_d, arr_l, infos, pin.data_ptr(), cap, C.byref(used), offs, C.byref(tot), on_gpu)
                res["gpu_entropy"]["host_prepare_ms_per_image_1core"] = (time.perf_counter() - t0) / (10 * n) * 1e3
                blob = pin[: used.value].cuda()
                coef = torch.empty(int(tot.value), dtype=torch.int16, device="cuda")
                st = torch.cuda.current_stream().cuda_stream
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                planes = torch.empty(max(int(i.plane_bytes) for i in infos), dtype=torch.uint8, device="cuda")
                img = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
                for rep in range(3):
                    ev[0].record()
                    L.vltk_jpeg_gpu_entropy_decode(n, blob.data_ptr(), coef.data_ptr(), int(tot.value), None, st)
                    ev[1].record()
                    for i in range(n):
                        L.vltk_jpeg_reconstruct(coef.data_ptr() + 2 * int(offs[i]), C.byref(infos[i]), planes.data_ptr(), img.data_ptr(), st)
                    ev[2].record()
                torch.cuda.synchronize()
                res["gpu_entropy"]["entropy_kernels_ms_per_batch"] = ev[0].elapsed_time(ev[1])
                res["gpu_entropy"]["idct_color_ms_per_batch"] = ev[1].elapsed_time(ev[2])
                res["gpu_entropy"]["h2d_bytes_per_batch"] = int(used.value)
        out[label] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
