"""Shared helpers for the parity tests (golden loading, cached oracle runs)."""
import functools
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLD, f"{name}.npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["meta"] = json.loads(str(g["meta"]))
    return g


def checksum(t):
    t = torch.as_tensor(t).double()
    return np.array([t.sum().item(), t.abs().sum().item(), float(t.numel())])


@functools.lru_cache(maxsize=None)
def weights(seed=0):
    from vltk_b200 import synthetic
    from vltk_b200.config import FRCNNConfig
    return synthetic.make_state_dict(FRCNNConfig(), seed)


@functools.lru_cache(maxsize=None)
def oracle_run(name):
    """(cfg, images, sizes, scales, out, stages) of the oracle port on a named case."""
    from oracle import cases, frcnn_oracle as O
    cfg, wseed, raws = cases.case_inputs(name)
    sd = weights(wseed)
    images, sizes, scales = O.preprocess(cfg, raws)
    st = {}
    out = O.forward(sd, cfg, images, sizes, scales, stages=st, ignorey=cases.case_ignorey(name))
    return cfg, images, sizes, scales, out, st


def match_proposals(mine_boxes, mine_logits, gold_boxes, gold_logits, box_tol=1e-2, tie_rel=1e-5, window=8):
    """Maps each golden RPN proposal i to the position perm[i] of the same box in `mine`.

    The 300 survivors carry near-uniform random logits whose smallest consecutive gap is ~1e-6
    relative on every seed (oracle/margins.py): two fp32 implementations with different summation
    order legitimately disagree on the ORDER of such near-ties, never on the membership.  So the
    lists must be equal up to a permutation that only ever exchanges entries whose golden logits
    are tied within `tie_rel`; anything else fails."""
    n = len(gold_boxes)
    assert len(mine_boxes) == n, (len(mine_boxes), n)
    perm = np.full(n, -1, np.int64)
    used = np.zeros(n, bool)
    for i in range(n):
        lo, hi = max(0, i - window), min(n, i + window + 1)
        d = np.abs(mine_boxes[lo:hi] - gold_boxes[i]).max(1)
        d[used[lo:hi]] = np.inf
        j = int(np.argmin(d))
        assert d[j] <= box_tol, f"golden proposal {i} has no counterpart within {box_tol}px (best {d[j]:.3g})"
        perm[i] = lo + j
        used[lo + j] = True
    assert used.all()
    moved = np.nonzero(perm != np.arange(n))[0]
    for i in moved:
        a, b = float(gold_logits[i]), float(gold_logits[perm[i]])
        assert abs(a - b) <= tie_rel * max(abs(a), 1e-6), \
            f"proposal {i} moved to rank {perm[i]} but its logit is not tied ({a} vs {b})"
    np.testing.assert_allclose(mine_logits[perm], gold_logits, rtol=1e-4, atol=1e-4)
    return perm
