"""Pins oracle/frcnn_oracle.py (the CPU restatement) against tests/golden/*.npz, which
oracle/make_goldens.py produced by running the UNMODIFIED reference.  Same CPU kernels
underneath (torch conv/linear), so dense stages agree to fp32 round-off and every
index is exact."""
import numpy as np
import pytest
import torch

from oracle import cases
from tests.util import checksum, load_golden, oracle_run


def close(a, b, rtol=1e-5, atol=1e-5):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", cases.CPU_CASES)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    cfg, images, sizes, scales, out, st = oracle_run(name)
    # inputs: the oracle's Preprocess restatement vs the reference's Preprocess
    close(checksum(images), g["images_ck"], rtol=1e-6)
    assert np.array_equal(sizes.numpy(), g["sizes"])
    close(scales.numpy(), g["scales_yx"], rtol=0, atol=0)
    # dense stages
    close(st["res4"][:, ::16], g["res4_sub"], rtol=1e-4, atol=1e-4)
    close(st["rpn_logits"], g["rpn_logits"], rtol=1e-4, atol=1e-4)
    close(st["rpn_deltas"][:, :, ::2, ::2], g["rpn_deltas_sub"], rtol=1e-4, atol=1e-4)
    # RPN selection: exact count, boxes to sub-pixel round-off
    assert [len(p) for p in st["proposals"]] == g["n_props"].tolist()
    close(torch.cat(st["proposals"]), g["proposals"], rtol=0, atol=1e-3)
    close(torch.cat(st["proposal_logits"]), g["proposal_logits"], rtol=1e-4, atol=1e-4)
    # ROI head
    close(st["feats"][:, ::8], g["feats_sub"], rtol=1e-4, atol=1e-4)
    assert np.array_equal(st["obj_logits"].argmax(-1).numpy(), g["obj_argmax_all"])
    assert np.array_equal(st["obj_logits"][:, :-1].argmax(-1).numpy(), g["obj_fg_argmax_all"])
    close(st["obj_logits"][:, ::16], g["obj_logits_sub"], rtol=1e-4, atol=1e-4)
    close(st["attr_logits"][:, ::8], g["attr_logits_sub"], rtol=1e-4, atol=1e-4)
    close(st["box_deltas"][:, ::64], g["box_deltas_sub"], rtol=1e-4, atol=1e-4)
    # final outputs: ids/counts exact, floats tight
    assert out["preds_per_image"].tolist() == g["preds_per_image"].tolist()
    assert np.array_equal(torch.cat(out["obj_ids"]).numpy(), g["obj_ids"])
    assert np.array_equal(torch.cat(out["attr_ids"]).numpy(), g["attr_ids"])
    close(torch.cat(out["boxes"]), g["boxes"], rtol=0, atol=1e-3)
    close(torch.cat(out["obj_probs"]), g["obj_probs"], rtol=0, atol=1e-5)
    close(torch.cat(out["attr_probs"]), g["attr_probs"], rtol=0, atol=1e-5)
    s = int(g["roi_features_stride"])
    close(torch.cat(out["roi_features"])[:, ::s], g["roi_features"], rtol=1e-4, atol=1e-4)


def test_oracle_tiny_full_stage_tensors():
    g = load_golden("tiny")
    _, _, _, _, _, st = oracle_run("tiny")
    close(st["res4"], g["res4"], rtol=1e-4, atol=1e-4)
    close(st["feats"], g["feats"], rtol=1e-4, atol=1e-4)
    close(st["obj_logits"], g["obj_logits"], rtol=1e-4, atol=1e-4)
    close(st["attr_logits"], g["attr_logits"], rtol=1e-4, atol=1e-4)
