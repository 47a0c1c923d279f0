"""How does tcgen05 accumulate in fp32?  (diagnosis for the exact tensor-core mode)

With bf16-EXACT operands the lo planes of vltk_linear_tc3 are zero, so its fp32 output is the raw
tensor-core accumulation of exact products: any difference from the fp64 dot product is accumulation
rounding.  Reports the signed bias (towards zero = truncation) and the spread, next to a plain fp32
torch.matmul of the same operands.
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from vltk_b200 import stages  # noqa: E402


def stats(y, ref):
    err = (y.double() - ref)
    rel = err / ref.abs().clamp_min(1e-30)
    toward_zero = -(err * torch.sign(ref)) / ref.abs().clamp_min(1e-30)   # > 0: magnitude shrank
    big = ref.abs() > ref.abs().median()
    return {"rel_rms": float(rel[big].pow(2).mean().sqrt()), "rel_max": float(rel[big].abs().max()),
            "shrink_mean_ulp24": float(toward_zero[big].mean() * 2 ** 24),
            "rel_rms_ulp24": float(rel[big].pow(2).mean().sqrt() * 2 ** 24)}


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    out = {}
    for K in (256, 1152, 2304, 4608):
        for kind in ("relu_x_signed_w", "all_positive", "signed"):
            M, N = 512, 256
            x = torch.randn(M, K, generator=g)
            w = torch.randn(N, K, generator=g) * (2.0 / K) ** 0.5
            if kind == "relu_x_signed_w":
                x = x.clamp_min(0)
            elif kind == "all_positive":
                x, w = x.abs(), w.abs()
            xb, wb = x.bfloat16().float(), w.bfloat16().float()     # bf16-exact operands
            ref = xb.double() @ wb.double().t()
            y_tc = stages.linear_tc3(xb.to(dev), wb.to(dev)).cpu()
            y_f32 = (xb.to(dev) @ wb.to(dev).t()).cpu()
            # fp32 operands: the 3-pass split's representation error on top
            ref_full = x.double() @ w.double().t()
            y_tc_full = stages.linear_tc3(x.to(dev), w.to(dev)).cpu()
            out[f"K{K}_{kind}"] = {"tc_bf16exact": stats(y_tc, ref), "torch_fp32": stats(y_f32, ref),
                                   "tc3_fp32_operands": stats(y_tc_full, ref_full)}
            print(K, kind, json.dumps(out[f"K{K}_{kind}"]), flush=True)
    json.dump(out, open("gpurun_out/acc_probe.json", "w"), indent=1)


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = False
    main()
