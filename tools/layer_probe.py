"""Runs single res5-shaped layers through the stage entry points, for `ncu` (per-launch time / DRAM bytes):
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_tc[23]_kernel \\
       python tools/layer_probe.py [layer ...]      # layers: concat conv1 conv2 conv3res conv3 c512
The stage entry points pack weights and allocate scratch per call, so wall-clock timing here is meaningless: read the
kernel durations from ncu."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vltk_b200 import stages

dev = torch.device("cuda", 0)
torch.manual_seed(0)
R = int(os.environ.get("PROBE_ROIS", "2400"))
REPS = int(os.environ.get("PROBE_REPS", "3"))
MODE = os.environ.get("PROBE_MODE", "bf16")        # exact_tc: fp32 tensors through conv_tcx (concat has no stage entry there)


def act(c):
    x = torch.randn(R, 14, 14, c, device=dev)
    return x if MODE == "exact_tc" else x.bfloat16()


def w(cout, cin, k=1):
    return torch.randn(cout, cin, k, k, device=dev) * (2.0 / (cin * k * k)) ** 0.5


def run(name):
    one = lambda c: (torch.ones(c, device=dev), torch.zeros(c, device=dev))
    if name == "concat":
        x, x2, w1, w2 = act(512), act(1024), w(2048, 512).reshape(2048, 512), w(2048, 1024).reshape(2048, 1024)
        f = lambda: stages.conv2d_dual_nhwc(x, w1, x2, w2, None, stride2=1, relu=True)
    elif name == "conv1":
        x, wt = act(2048), w(512, 2048); sc, sh = one(512)
        f = lambda: stages.conv2d_nhwc(x, wt, sc, sh, None, 1, 0, 1, True, mode=MODE, tensor_cores=True)
    elif name == "conv2":
        x, wt = act(512), w(512, 512, 3); sc, sh = one(512)
        f = lambda: stages.conv2d_nhwc(x, wt, sc, sh, None, 1, 2, 2, True, mode=MODE, tensor_cores=True)
    elif name in ("conv3res", "conv3"):
        x, wt = act(512), w(2048, 512); sc, sh = one(2048)
        res = act(2048) if name == "conv3res" else None
        f = lambda: stages.conv2d_nhwc(x, wt, sc, sh, res, 1, 0, 1, True, mode=MODE, tensor_cores=True)
    elif name == "c512":
        x, wt = act(512), w(512, 512); sc, sh = one(512)
        f = lambda: stages.conv2d_nhwc(x, wt, sc, sh, None, 1, 0, 1, True, mode=MODE, tensor_cores=True)
    else:
        raise SystemExit(f"unknown layer {name}")
    for _ in range(REPS):
        f()
    torch.cuda.synchronize()
    print("ran", name, flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["concat", "conv1", "conv2", "conv3res"]):
        run(n)
