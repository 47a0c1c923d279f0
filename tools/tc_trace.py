"""Pipeline trace of one CTA of the single-CTA tcgen05 conv kernel (diagnosis build only):

    VLTK_TRACE=1 bash vltk_b200/csrc/build.sh
    VLTK_LIB=libvltk_frcnn_trace.so python tools/tc_trace.py [--res 1] [--cin 512] [--cout 2048] [--out trace.json]

Each role of CTA `--cta` records clock64 at its synchronisation points (include/vltk_frcnn.h, vltk_conv_tc_set_trace).
Prints, per role, the mean cycles between consecutive events in steady state — i.e. where each warp role waits.
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vltk_b200 import stages, _lib

ROLES = ["tma_producer", "mma_issuer", "res_producer", "epilogue_g0", "epilogue_g1"]
EV = {
    0: {1: "empty slot acquired"},
    1: {0: "tile start", 1: "accumulator acquired (tempty)", 2: "operands landed (full)", 3: "tile committed",
        4: "kernel entry", 5: "set-up done + predecessor complete"},
    2: {1: "ring slot acquired (rempty)"},
    3: {0: "tile start", 1: "accumulator ready (tfull)", 2: "tcgen05.ld done", 3: "residual landed", 4: "staging drained (bulk wait)",
        5: "barrier 1", 6: "slab computed + written", 7: "barrier 2 + store issued", 8: "all stores complete"},
}
EV[4] = EV[3]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rois", type=int, default=2400, help="images / ROIs (N of the NHWC input)")
    ap.add_argument("--h", type=int, default=14)
    ap.add_argument("--w", type=int, default=14)
    ap.add_argument("--pad", type=int, default=-1)
    ap.add_argument("--dil", type=int, default=-1)
    ap.add_argument("--timeline", type=int, default=0, help="print the first N raw records of every role")
    ap.add_argument("--cin", type=int, default=512)
    ap.add_argument("--cout", type=int, default=2048)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--res", type=int, default=1)
    ap.add_argument("--cta", type=int, default=5)
    ap.add_argument("--pairs", type=int, default=0, help="1: trace the CTA-pair kernel (conv_tc3_kernel; --cta even = the leader CTA)")
    ap.add_argument("--cap", type=int, default=8192)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    x = torch.randn(a.rois, a.h, a.w, a.cin, device=dev).bfloat16()
    wt = torch.randn(a.cout, a.cin, a.k, a.k, device=dev) * (2.0 / (a.cin * a.k * a.k)) ** 0.5
    sc = torch.ones(a.cout, device=dev); sh = torch.zeros(a.cout, device=dev)
    res = torch.randn(a.rois, a.h, a.w, a.cout, device=dev).bfloat16() if a.res else None
    pad = a.pad if a.pad >= 0 else (0 if a.k == 1 else 2)
    dil = a.dil if a.dil >= 0 else (1 if a.k == 1 else 2)
    stages.set_cta_pairs(1 if a.pairs else 0, 1 if a.pairs else 0)
    buf = torch.zeros(5 * a.cap * 2, dtype=torch.int64, device=dev)
    L = _lib.lib()
    for _ in range(2):   # warm-up launch, then the traced one (the buffer is overwritten)
        buf.zero_()
        _lib.check(L.vltk_conv_tc_set_trace(buf.data_ptr(), a.cap, a.cta), "set_trace")
        stages.conv2d_nhwc(x, wt, sc, sh, res, 1, pad, dil, True, mode="bf16", tensor_cores=True)
        L.vltk_conv_tc_set_trace(None, 0, 0)
    t = buf.cpu().reshape(5, a.cap, 2).numpy()
    summary = {}
    allclk = [int(c) for r in range(5) for _, c in t[r] if c]
    t0 = min(allclk)
    print(f"CTA {a.cta}: first record .. last record = {max(allclk) - t0} cycles")
    if a.timeline:
        for r, name in enumerate(ROLES):
            recs = [(int(tag) >> 40, (int(tag) >> 16) & 0xFFFFFF, int(tag) & 0xFFFF, int(clk)) for tag, clk in t[r] if clk]
            print(f"-- {name}")
            for ev, it, idx, clk in recs[: a.timeline]:
                print(f"   +{clk - t0:8d}  tile {it:3d} idx {idx:3d}  {EV[r][ev]}")
    for r, name in enumerate(ROLES):
        rec = [(int(tag) >> 40, (int(tag) >> 16) & 0xFFFFFF, int(tag) & 0xFFFF, int(clk)) for tag, clk in t[r] if clk]
        if not rec:
            continue
        n = len(rec)
        lo, hi = n // 4, n - n // 8           # steady state: skip the ramp and the tail
        gaps = {}
        for i in range(max(lo, 1), hi):
            key = f"{EV[r][rec[i - 1][0]]} -> {EV[r][rec[i][0]]}"
            gaps.setdefault(key, []).append(rec[i][3] - rec[i - 1][3])
        span = rec[hi - 1][3] - rec[lo][3]
        tiles = len({x[1] for x in rec[lo:hi]}) if r in (1, 3, 4) else None
        print(f"== {name}: {n} records, steady window {span} cycles" + (f", {tiles} tiles -> {span / max(tiles, 1):.0f} cycles/tile" if tiles else ""))
        summary[name] = {"records": n, "window_cycles": span, "tiles": tiles, "gaps": {}}
        for k, v in sorted(gaps.items(), key=lambda kv: -sum(kv[1])):
            v.sort()
            print(f"   {k:70s} n={len(v):5d} mean {sum(v) / len(v):8.0f}  median {v[len(v) // 2]:6d}  p90 {v[int(len(v) * 0.9)]:6d}  share {100.0 * sum(v) / span:5.1f}%")
            summary[name]["gaps"][k] = {"n": len(v), "mean": sum(v) / len(v), "median": v[len(v) // 2], "p90": v[int(len(v) * 0.9)], "share": sum(v) / span}
    if a.out:
        json.dump({"args": vars(a), "summary": summary}, open(a.out, "w"), indent=1)
    stages.set_cta_pairs(32768, 0)


if __name__ == "__main__":
    main()
