"""Bring-up diagnostics run on the GPU box (prints, never asserts): python tools/gpu_first.py <what>"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vltk_b200 import stages  # noqa: E402


def ref_conv(x_nhwc, w, scale, shift, res, stride, pad, dil, relu):
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2).float(), w.float(), None, stride, pad, dil)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + res.float()
    return F.relu(y) if relu else y


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad, dil, res, relu
    (1, 12, 16, 64, 64, 1, 1, 0, 1, False, True),
    (2, 14, 14, 64, 128, 3, 1, 2, 2, False, True),
    (1, 25, 33, 128, 256, 1, 2, 0, 1, False, False),
    (3, 14, 14, 128, 256, 1, 1, 0, 1, True, True),
    (1, 19, 23, 64, 64, 3, 1, 1, 1, False, True),
    (2, 14, 14, 256, 512, 3, 1, 2, 2, True, True),
    (1, 50, 84, 1024, 512, 3, 1, 1, 1, False, True),
    (5, 14, 14, 1024, 512, 1, 1, 0, 1, False, True),
    (5, 14, 14, 512, 2048, 1, 1, 0, 1, True, True),
]


def run_conv_cases(mode, tc):
    torch.manual_seed(0)
    dev = torch.device("cuda")
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    for (n, h, w, cin, cout, k, s, p, d, use_res, relu) in CONV_CASES:
        x = torch.randn(n, h, w, cin, device=dev).to(dt)
        wt = torch.randn(cout, cin, k, k, device=dev) * (2.0 / (cin * k * k)) ** 0.5
        if mode == "bf16":
            wt = wt.bfloat16().float()
        sc = torch.rand(cout, device=dev) + 0.5
        sh = torch.randn(cout, device=dev) * 0.1
        oh = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
        ow = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
        res = torch.randn(n, oh, ow, cout, device=dev).to(dt) if use_res else None
        t0 = time.time()
        try:
            y = stages.conv2d_nhwc(x, wt, sc, sh, res, s, p, d, relu, mode=mode, tensor_cores=tc)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"conv {mode} tc={tc} case {(n, h, w, cin, cout, k, s, p, d)} FAILED: {e}", flush=True)
            return
        ref = ref_conv(x, wt, sc, sh, res, s, p, d, relu)
        err = (y.float() - ref).abs()
        denom = ref.abs().max().item() + 1e-6
        bad = (err > (1e-3 if mode == "fp32" else 3e-2) * (ref.abs() + 1.0)).sum().item()
        print(f"conv {mode} tc={tc} n{n} {h}x{w} cin{cin} cout{cout} k{k} s{s} p{p} d{d} res{int(use_res)}: "
              f"max_abs_err {err.max().item():.3e} (ref max {denom:.2f}) bad {bad}/{err.numel()} "
              f"{(time.time() - t0) * 1e3:.1f} ms", flush=True)
        if bad and tc:
            # locate the damage: which rows / channels are wrong
            e2 = err.reshape(-1, cout)
            rows = (e2.max(1).values > 3e-2 * (ref.abs().max() + 1)).nonzero().flatten()
            cols = (e2.max(0).values > 3e-2 * (ref.abs().max() + 1)).nonzero().flatten()
            print(f"   bad rows {rows.numel()} first {rows[:12].tolist()} | bad cols {cols.numel()} first {cols[:12].tolist()}")
            print("   y[0,:8]  ", y.reshape(-1, cout)[0, :8].float().tolist())
            print("   ref[0,:8]", ref.reshape(-1, cout)[0, :8].tolist())


def run_e2e(case, mode):
    from oracle import cases
    from tests.util import load_golden
    from vltk_b200 import synthetic
    from vltk_b200.frcnn import FRCNN
    from vltk_b200.preprocess import Preprocess
    cfg, wseed, raws = cases.case_inputs(case)
    t0 = time.time()
    sd = synthetic.make_state_dict(cfg, wseed)
    print(f"[{case}/{mode}] weights {time.time() - t0:.1f}s", flush=True)
    t0 = time.time()
    model = FRCNN.from_pretrained(state_dict=sd, config=cfg, mode=mode)
    print(f"[{case}/{mode}] engine {time.time() - t0:.1f}s", flush=True)
    ids, images, sizes, scales = Preprocess(cfg)(raws)
    g = load_golden(case)
    ck = np.array([images.double().sum().item(), images.double().abs().sum().item(), images.numel()])
    print(f"[{case}] images ck {ck} golden {g['images_ck']}", flush=True)
    t0 = time.time()
    out = model(images, sizes, scales_yx=scales)
    torch.cuda.synchronize()
    print(f"[{case}/{mode}] forward {time.time() - t0:.3f}s preds {out['preds_per_image'].tolist()} golden {g['preds_per_image'].tolist()}")
    res4 = model.debug_read("res4")
    n, h4, w4 = len(raws), *cfg.res4_hw(images.shape[2], images.shape[3])
    res4 = torch.from_numpy(res4).view(n, h4, w4, -1).permute(0, 3, 1, 2)
    gs = torch.from_numpy(g["res4_sub"])
    e = (res4[:, ::16] - gs).abs()
    print(f"[{case}/{mode}] res4 max err {e.max():.3e} (ref max {gs.abs().max():.2f}, mean {gs.abs().mean():.3f})")
    cnt = model.debug_read("proposal_count", np.int32)
    print(f"[{case}/{mode}] proposal counts {cnt.tolist()} golden {g['n_props'].tolist()}")
    props = torch.from_numpy(model.debug_read("proposals")).view(n, -1, 4)
    gp = torch.from_numpy(g["proposals"])
    mine = torch.cat([props[i, : int(cnt[i])] for i in range(n)])
    if mine.shape == gp.shape:
        pe = (mine - gp).abs().max(1).values
        print(f"[{case}/{mode}] proposals: max err {pe.max():.3e}, rows >1e-2: {(pe > 1e-2).sum().item()}/{len(pe)}")
    oid = torch.cat(out["obj_ids"]).numpy()
    aid = torch.cat(out["attr_ids"]).numpy()
    bx = torch.cat(out["boxes"]).numpy()
    if oid.shape == g["obj_ids"].shape:
        print(f"[{case}/{mode}] obj_ids equal {(oid == g['obj_ids']).sum()}/{len(oid)} attr_ids equal {(aid == g['attr_ids']).sum()}/{len(aid)} "
              f"boxes max err {np.abs(bx - g['boxes']).max():.3e} "
              f"probs max err {np.abs(torch.cat(out['obj_probs']).numpy() - g['obj_probs']).max():.3e}")
        s = int(g["roi_features_stride"])
        rf = torch.cat(out["roi_features"]).numpy()[:, ::s]
        print(f"[{case}/{mode}] roi_features max err {np.abs(rf - g['roi_features']).max():.3e} (ref max {np.abs(g['roi_features']).max():.2f})")
    print(f"[{case}/{mode}] launches {model.launch_count()}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "simt":
        run_conv_cases("fp32", False)
        run_conv_cases("bf16", False)
    elif what == "tc":
        run_conv_cases("bf16", True)
    elif what == "e2e":
        run_e2e(sys.argv[2], sys.argv[3])
