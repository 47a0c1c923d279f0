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
    out = O.forward(sd, cfg, images, sizes, scales, stages=st)
    return cfg, images, sizes, scales, out, st
