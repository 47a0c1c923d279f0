"""TEST INFRASTRUCTURE — searches image seeds for which every decision margin of a case
clears oracle/margins.THRESHOLDS (the "margin-certified goldens" of SURVEY.md §7):

    python oracle/certify.py cfg1 [max_tries]

prints the margins per try and the first passing seed list to paste into oracle/cases.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import cases, frcnn_oracle as O, margins  # noqa: E402
from vltk_b200 import synthetic  # noqa: E402
from vltk_b200.config import FRCNNConfig  # noqa: E402


def main():
    name = sys.argv[1]
    tries = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    over, wseed, imgs = cases.CASES[name]
    cfg = FRCNNConfig().replace(**over)
    sd = synthetic.make_state_dict(cfg, wseed)
    for t in range(tries):
        seeds = [(h, w, s + 1000 * t if isinstance(s, int) else s) for (h, w, s) in imgs]
        raws = [cases.raw_image(h, w, s) for (h, w, s) in seeds]
        images, sizes, scales = O.preprocess(cfg, raws)
        st = {}
        out = O.forward(sd, cfg, images, sizes, scales, stages=st)
        m = margins.margins(cfg, st, out)
        ok = margins.certified(m)
        bad = {k: f"{m[k]:.1e}" for k, v in margins.THRESHOLDS.items() if m[k] < v}
        print(f"{name} try {t} seeds {seeds} -> {'CERTIFIED' if ok else 'fail ' + str(bad)}", flush=True)
        if ok:
            print({k: f"{v:.2e}" for k, v in m.items()})
            return


if __name__ == "__main__":
    main()
