"""End-to-end parity of the CUDA path (through the C ABI / FRCNN.forward) against
  (a) tests/golden/*.npz — outputs of the UNMODIFIED reference, generated in the authoring
      container by oracle/make_goldens.py, and
  (b) the oracle port run on this box for the small cases (full tensors).

fp32 mode: kept-proposal indices, obj_ids, attr_ids, preds_per_image EXACT; boxes <= 1e-2 px;
probs <= 1e-5; roi_features rel 1e-4 (SURVEY.md §8c parity rules).
bf16 mode (single-pass tensor-core operands): dense stages within bf16 tolerance; selection
parity is asserted per stage with teacher forcing (test_gpu_stages.py) and end-to-end agreement
is checked statistically, as SURVEY.md Appendix E explains random-init scores make rank flips
unavoidable at bf16 operand precision."""
import numpy as np
import pytest
import torch

from oracle import cases
from tests.util import load_golden, match_proposals, oracle_run, weights

pytestmark = pytest.mark.gpu

_models = {}


def get_model(case, mode):
    """One engine per (selection config, mode): weights are shared by every case."""
    from vltk_b200.frcnn import FRCNN
    cfg = cases.case_config(case)
    key = (cfg.rpn_pre_nms_topk, cfg.rpn_post_nms_topk, mode)
    if key not in _models:
        _models[key] = FRCNN.from_pretrained(state_dict=weights(cases.CASES[case][1]), config=cfg, mode=mode)
    m = _models[key]
    m.roi_outputs.nms_thresh = list(cfg.nms_thresh_test)
    m.roi_outputs.min_detections = cfg.min_detections
    m.roi_outputs.max_detections = cfg.max_detections
    return m, cfg


def run_case(case, mode, **kw):
    from vltk_b200.preprocess import Preprocess
    model, cfg = get_model(case, mode)
    _, _, raws = cases.case_inputs(case)
    ids, images, sizes, scales = Preprocess(cfg)(raws)
    out = model(images, sizes, scales_yx=scales, ignorey=cases.case_ignorey(case), **kw)
    return model, cfg, images, sizes, scales, out


def cat(x):
    return torch.cat(list(x)).cpu().numpy()


# the two index-exact modes: fp32 FMA chains on the CUDA cores, and the fp32-faithful tensor-core mode
# (csrc/conv_tcx.cu: split-fp16 operands, three tcgen05 passes, chunk sums promoted to fp32 registers)
EXACT_MODES = ("fp32", "exact_tc")


@pytest.mark.parametrize("mode", EXACT_MODES)
@pytest.mark.parametrize("case", cases.GPU_CASES)
def test_fp32_matches_reference_golden(case, mode):
    g = load_golden(case)
    model, cfg, images, sizes, scales, out = run_case(case, mode)
    n = images.shape[0]
    ck = np.array([images.double().sum().item(), images.double().abs().sum().item(), images.numel()])
    np.testing.assert_allclose(ck, g["images_ck"], rtol=1e-6)
    # backbone
    h4, w4 = cfg.res4_hw(images.shape[2], images.shape[3])
    res4 = torch.from_numpy(model.debug_read("res4")).view(n, h4, w4, -1).permute(0, 3, 1, 2)
    np.testing.assert_allclose(res4[:, ::16].numpy(), g["res4_sub"], rtol=1e-3, atol=1e-3)
    # proposals: same count, same boxes; order equal up to near-tied logits (tests/util.py)
    cnt = model.debug_read("proposal_count", np.int32)
    assert cnt.tolist() == g["n_props"].tolist()
    props = torch.from_numpy(model.debug_read("proposals")).view(n, -1, 4)
    plog = torch.from_numpy(model.debug_read("proposal_logits")).view(n, -1)
    feats = torch.from_numpy(model.debug_read("feats")).view(n, -1, 2048)
    # logits rows are padded (to a multiple of 4 on the CUDA cores, of 128 on the tensor pipe)
    cl = torch.from_numpy(model.debug_read("cls_logits")).view(n, props.shape[1], -1)[:, :, : cfg.num_classes + 1]
    s0 = 0
    for i in range(n):
        c = int(cnt[i])
        gs = slice(s0, s0 + c)
        s0 += c
        perm = match_proposals(props[i, :c].numpy(), plog[i, :c].numpy(), g["proposals"][gs], g["proposal_logits"][gs])
        # pooled res5 features of every proposal + per-ROI class decisions
        np.testing.assert_allclose(feats[i, :c].numpy()[perm][:, ::8], g["feats_sub"][gs], rtol=1e-3, atol=1e-3)
        assert np.array_equal(cl[i, :c].argmax(-1).numpy()[perm], g["obj_argmax_all"][gs])
        assert np.array_equal(cl[i, :c, :-1].argmax(-1).numpy()[perm], g["obj_fg_argmax_all"][gs])
    # the pre-NMS top-k anchor SET (order inside it is the same near-tie story)
    tk = model.debug_read("topk_anchor_idx", np.int32).reshape(n, -1)
    for i in range(n):
        assert set(tk[i].tolist()) == set(g["rpn_topk_anchor_idx"][i].tolist())
    # final detections
    assert out["preds_per_image"].tolist() == g["preds_per_image"].tolist()
    assert np.array_equal(cat(out["obj_ids"]), g["obj_ids"])
    assert np.array_equal(cat(out["attr_ids"]), g["attr_ids"])
    np.testing.assert_allclose(cat(out["boxes"]), g["boxes"], rtol=0, atol=1e-2)
    np.testing.assert_allclose(cat(out["obj_probs"]), g["obj_probs"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(cat(out["attr_probs"]), g["attr_probs"], rtol=0, atol=1e-5)
    s = int(g["roi_features_stride"])
    np.testing.assert_allclose(cat(out["roi_features"])[:, ::s], g["roi_features"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("mode", EXACT_MODES)
@pytest.mark.parametrize("case", ["tiny", "mixed", "ignorey"])
def test_fp32_matches_oracle_full_tensors(case, mode):
    cfg, oimg, osz, osc, oout, st = oracle_run(case)
    model, cfg, images, sizes, scales, out = run_case(case, mode)
    for i, k in enumerate(oout["keep"]):
        assert torch.equal(out["keep_idx"][i].cpu(), k)   # kept-box indices into the proposal list
        assert torch.equal(out["obj_ids"][i].cpu(), oout["obj_ids"][i])
        assert torch.equal(out["attr_ids"][i].cpu(), oout["attr_ids"][i])
        np.testing.assert_allclose(out["roi_features"][i].cpu().numpy(), oout["roi_features"][i].numpy(),
                                   rtol=1e-4, atol=1e-4)
    # padded contract (v1.0.0): dense tensors + sizes + normalized_boxes
    from oracle import frcnn_oracle as O
    dense = model(images, sizes, scales_yx=scales, padding="max_detections", return_tensors="np",
                  ignorey=cases.case_ignorey(case))
    ref = O.pad_outputs(oout, osz, osc, cfg.max_detections)
    assert dense["roi_features"].shape == (images.shape[0], cfg.max_detections, 2048)
    assert np.array_equal(dense["obj_ids"], ref["obj_ids"].numpy())
    assert np.array_equal(dense["sizes"], ref["sizes"].numpy())
    np.testing.assert_allclose(dense["normalized_boxes"], ref["normalized_boxes"].numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(dense["boxes"], ref["boxes"].numpy(), rtol=0, atol=1e-2)


def test_exact_tc_matches_the_oracle_on_fresh_uncommitted_seeds():
    """Index parity that cannot be an artefact of the committed cases: ten image seeds no golden uses, the oracle run here
    on the box's CPU, its decision margins measured (oracle/margins.py), and — on every seed whose margins clear the
    certification thresholds, i.e. where two fp32 implementations are REQUIRED to agree — kept indices, ids and counts
    must equal the oracle's in exact_tc mode.  Uncertified seeds (a near-tie somewhere) are reported, not asserted."""
    from oracle import frcnn_oracle as O, margins
    from vltk_b200.preprocess import Preprocess
    model, cfg = get_model("tiny", "exact_tc")
    certified = exact = 0
    for seed in range(7100, 7110):
        raws = [cases.raw_image(192, 256, seed)]
        oimg, osz, osc = O.preprocess(cfg, raws)
        st = {}
        ref = O.forward(weights(0), cfg, oimg, osz, osc, stages=st)
        ok = margins.certified(margins.margins(cfg, st, ref))
        ids, images, sizes, scales = Preprocess(cfg)(raws)
        out = model(images, sizes, scales_yx=scales)
        same = (out["preds_per_image"].tolist() == ref["preds_per_image"].tolist()
                and torch.equal(out["keep_idx"][0].cpu(), ref["keep"][0])
                and torch.equal(out["obj_ids"][0].cpu(), ref["obj_ids"][0])
                and torch.equal(out["attr_ids"][0].cpu(), ref["attr_ids"][0]))
        certified += ok
        exact += same
        print(f"seed {seed}: margins {'certified' if ok else 'near-tie'}, exact_tc {'== oracle' if same else 'differs'}")
        if ok:
            assert same, f"seed {seed}: margins are certified but exact_tc differs from the oracle"
            np.testing.assert_allclose(out["boxes"][0].cpu().numpy(), ref["boxes"][0].numpy(), rtol=0, atol=1e-2)
            np.testing.assert_allclose(out["roi_features"][0].cpu().numpy(), ref["roi_features"][0].numpy(), rtol=1e-4, atol=1e-4)
    print(f"{certified} of 10 fresh seeds certified; exact_tc equals the oracle on {exact} of 10")
    assert certified >= 2, "too few certified seeds for the test to mean anything"


@pytest.mark.parametrize("mode", ["exact_tc", "bf16", "fp32"])
def test_engine_roi_pool_equals_torchvision_semantics_on_its_own_inputs(mode):
    """The RoIPool kernels the engine actually runs (roi_pool_h2_kernel on split-fp16 planes in exact_tc — the maximum is
    selected with packed fp16 compares of the (hi, lo') pair —, roi_pool_kernel<bf16 / f32> otherwise) against the oracle's
    restatement of torchvision.ops.RoIPool applied to the engine's OWN res4 map and proposals: a max copies values, so the
    pooled tensor must be bit-identical, including the zero rows past the proposal count."""
    from oracle import frcnn_oracle as O
    model, cfg, images, sizes, scales, out = run_case("few", mode)      # fewer proposals than the 300-slot budget
    n = images.shape[0]
    h4, w4 = cfg.res4_hw(images.shape[2], images.shape[3])
    res4 = torch.from_numpy(model.debug_read("res4")).view(n, h4, w4, -1).permute(0, 3, 1, 2).contiguous()
    cnt = model.debug_read("proposal_count", np.int32)
    props = torch.from_numpy(model.debug_read("proposals")).view(n, -1, 4)
    R = props.shape[1]
    pooled = torch.from_numpy(model.debug_read("pooled")).view(n, R, 14, 14, -1)
    assert 0 < int(cnt[0]) < R
    for i in range(n):
        c = int(cnt[i])
        rois = torch.cat([torch.full((c, 1), float(i)), props[i, :c]], 1)
        ref = O.roi_pool_np(res4, rois, 14, 1.0 / cfg.anchor_stride).permute(0, 2, 3, 1)
        assert torch.equal(pooled[i, :c], ref), mode
        assert not pooled[i, c:].any()


def test_ignorey_on_a_batch_applies_the_per_image_rule():
    """forward(..., ignorey=[N,J,2]) on a 2-image batch with different ranges per image.  The reference's branch
    cannot run this (it overwrites the shared level_ids after the first image, frcnn.py:340); the engine applies
    the same per-image rule to every image, checked against the oracle restatement (itself pinned to the real
    reference on the one-image `ignorey` golden)."""
    from oracle import frcnn_oracle as O
    from vltk_b200.preprocess import Preprocess
    model, cfg = get_model("mixed", "fp32")
    _, _, raws = cases.case_inputs("mixed")
    ids, images, sizes, scales = Preprocess(cfg)(raws)
    ign = torch.tensor([[[30.0, 55.0], [90.0, 120.0]], [[100.0, 140.0], [10.0, 20.0]]])
    oi, osz, osc = O.preprocess(cfg, raws)
    ref = O.forward(weights(0), cfg, oi, osz, osc, ignorey=ign)
    base = O.forward(weights(0), cfg, oi, osz, osc)
    assert any(not torch.equal(a, b) for a, b in zip(ref["boxes"], base["boxes"]))   # the ranges do change the result
    out = model(images, sizes, scales_yx=scales, ignorey=ign)
    assert out["preds_per_image"].tolist() == ref["preds_per_image"].tolist()
    for i in range(2):
        assert torch.equal(out["keep_idx"][i].cpu(), ref["keep"][i])
        assert torch.equal(out["obj_ids"][i].cpu(), ref["obj_ids"][i])
        np.testing.assert_allclose(out["boxes"][i].cpu().numpy(), ref["boxes"][i].numpy(), rtol=0, atol=1e-2)
    # without scales_yx the reference skips the branch entirely (frcnn.py:328)
    plain = model(images, sizes, ignorey=ign)
    noign = model(images, sizes)
    for i in range(2):
        assert torch.equal(plain["boxes"][i], noign["boxes"][i])


def _iou(a, b):
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / np.maximum(aa[:, None] + ab[None, :] - inter, 1e-9)


def _proposal_view(model, n):
    cnt = model.debug_read("proposal_count", np.int32)
    topk = model.debug_read("topk_anchor_idx", np.int32).reshape(n, -1)
    pos = model.debug_read("proposal_pos", np.int32).reshape(n, -1)
    props = model.debug_read("proposals").reshape(n, -1, 4)
    feats = model.debug_read("feats").reshape(n, -1, 2048)
    out = []
    for i in range(n):
        c = int(cnt[i])
        out.append(dict(anchor=topk[i][pos[i, :c]], box=props[i, :c], feat=feats[i, :c]))
    return out


@pytest.mark.parametrize("case", ["tiny", "full36", "cfg1", "cfg2x2"])
def test_bf16_tensor_core_mode_agrees_statistically(case):
    """bf16 mode vs fp32 mode (which the golden tests pin to the reference, ids exact), compared BY ANCHOR
    IDENTITY so that 'the same proposal' is a fact, not an IoU guess.  Stated bf16 tolerances:
      * res4 mean relative error < 2e-2 after ~100 bf16 layers
      * >= 85 % of the reference's proposals (same anchors) are selected, their boxes at IoU >= 0.8 on average
        (RPN deltas have std ~0.7 with these weights, so a bf16 error of ~1e-2 moves a box by ~1 %)
      * pooled 2048-d features of same-anchor proposals: median cosine >= 0.995 and >= 90 % above 0.98.  The
        tail is RoIPool's quantisation, not arithmetic: a 0.1 px shift across a half-cell rounding boundary
        moves the pooled window by a whole 16 px cell (torchvision semantics, reproduced bit-exactly).
    Final detections are reported, not asserted beyond their count: SURVEY.md Appendix E shows bf16 operand
    rounding re-ranks the near-uniform class scores of random-init weights."""
    g = load_golden(case)
    mf, cfg, images, sizes, scales, out_f = run_case(case, "fp32")
    n = images.shape[0]
    ref = _proposal_view(mf, n)
    h4, w4 = cfg.res4_hw(images.shape[2], images.shape[3])
    res4_f = mf.debug_read("res4")
    mb, _, _, _, _, out = run_case(case, "bf16")
    got = _proposal_view(mb, n)
    rel = np.abs(mb.debug_read("res4") - res4_f).mean() / np.abs(res4_f).mean()
    assert rel < 2e-2, rel
    assert out["preds_per_image"].tolist() == g["preds_per_image"].tolist()
    shared, ious, coss = [], [], []
    for r, b in zip(ref, got):
        idx_b = {a: k for k, a in enumerate(b["anchor"].tolist())}
        pairs = [(k, idx_b[a]) for k, a in enumerate(r["anchor"].tolist()) if a in idx_b]
        shared.append(len(pairs) / max(len(r["anchor"]), 1))
        ka = np.array([p[0] for p in pairs]); kb = np.array([p[1] for p in pairs])
        ious.append(np.diag(_iou(r["box"][ka], b["box"][kb])))
        fa, fb = r["feat"][ka], b["feat"][kb]
        coss.append((fa * fb).sum(1) / (np.linalg.norm(fa, axis=1) * np.linalg.norm(fb, axis=1) + 1e-12))
    ious, coss = np.concatenate(ious), np.concatenate(coss)
    iou_final = _iou(cat(out["boxes"]), cat(out_f["boxes"]))
    print(f"[{case}] bf16 vs reference-exact fp32: res4 rel err {rel:.2e}; same-anchor proposals {np.mean(shared):.2f}, their box IoU "
          f"mean {ious.mean():.3f} min {ious.min():.3f}, pooled-feature cosine median {np.median(coss):.4f} mean {coss.mean():.4f} "
          f">=0.98: {(coss >= 0.98).mean():.2f} min {coss.min():.3f}; final boxes re-found (IoU>=0.7) {(iou_final.max(0) >= 0.7).mean():.2f}, "
          f"obj_ids equal at rank {(cat(out['obj_ids']) == cat(out_f['obj_ids'])).mean():.2f}")
    assert np.mean(shared) >= 0.85, shared
    assert ious.mean() >= 0.8, ious.mean()
    assert np.median(coss) >= 0.995 and (coss >= 0.98).mean() >= 0.9, (np.median(coss), (coss >= 0.98).mean())
    # final detections against the REFERENCE golden: most of its boxes are re-found, and a re-found box carries the
    # reference's class id (a box only changes class when its top-2 logit gap is inside bf16 noise)
    gb, mb = g["boxes"], cat(out["boxes"])
    gi, mi = g["obj_ids"], cat(out["obj_ids"])
    found, same = [], []
    o0 = m0 = 0
    for i in range(n):
        ng, nm = int(g["preds_per_image"][i]), int(out["preds_per_image"][i])
        iou = _iou(gb[o0:o0 + ng], mb[m0:m0 + nm])
        j, hit = iou.argmax(1), iou.max(1) >= 0.5
        found.append(hit)
        same.append(mi[m0:m0 + nm][j][hit] == gi[o0:o0 + ng][hit])
        o0, m0 = o0 + ng, m0 + nm
    found, same = np.concatenate(found), np.concatenate(same)
    print(f"[{case}] bf16 vs reference golden: {found.mean():.2f} of the final boxes re-found at IoU >= 0.5, obj_ids equal on {same.mean():.2f} of them")
    assert found.mean() >= 0.6 and same.mean() >= 0.75, (found.mean(), same.mean())


@pytest.mark.parametrize("mode", EXACT_MODES)
def test_determinism_and_batch_invariance(mode):
    """Idempotence / order properties at full size (no oracle needed): the same input gives
    bit-identical outputs run to run, and an image's detections do not depend on its batch
    neighbours or its position in the batch."""
    from vltk_b200.preprocess import Preprocess
    model, cfg = get_model("cfg3x2", mode)
    _, _, raws = cases.case_inputs("cfg3x2")
    ids, images, sizes, scales = Preprocess(cfg)(raws)
    a = model(images, sizes, scales_yx=scales, padding="max_detections")
    b = model(images, sizes, scales_yx=scales, padding="max_detections")
    for k in ("boxes", "obj_ids", "attr_ids", "roi_features", "obj_probs"):
        assert torch.equal(a[k], b[k]), k
    # swapped batch order
    idx = torch.tensor([1, 0])
    sw = model(images[idx].contiguous(), sizes[idx], scales_yx=scales[idx], padding="max_detections")
    assert torch.equal(sw["obj_ids"][1], a["obj_ids"][0]) and torch.equal(sw["obj_ids"][0], a["obj_ids"][1])
    assert torch.equal(sw["roi_features"][0], a["roi_features"][1])


@pytest.mark.parametrize("mode", ["fp32", "bf16", "exact_tc"])
def test_full_batch8_equals_smaller_batches(mode):
    """BASELINE.json configs[1] at FULL size (batch 8 x 600x1000, the benchmarked workload), checked through a
    size-independent property: with identical canvases a batch is bit-identical to the same images run in
    smaller batches (every kernel treats images / ROI rows independently and deterministically).  Images 0-1
    are the `cfg2x2` golden pair, so the full batch is chained to the reference's outputs: golden == batch of 2
    (test_fp32_matches_reference_golden) == rows 0-1 of the batch of 8."""
    from vltk_b200 import synthetic
    model, cfg = get_model("cfg2x2", mode)
    mean = torch.tensor(cfg.pixel_mean).view(1, 3, 1, 1)
    raws = [synthetic.make_raw_image(600, 1000, 4010 + i) for i in range(8)]
    x = (torch.stack([r.permute(2, 0, 1).float() for r in raws]) - mean).contiguous()
    sizes = torch.tensor([[600, 1000]] * 8)
    scales = torch.ones(8, 2)
    keys = ("obj_ids", "attr_ids", "boxes", "obj_probs", "attr_probs", "roi_features", "preds_per_image", "keep_idx")
    full = model(x, sizes, scales_yx=scales, padding="max_detections", return_tensors="np")
    assert full["roi_features"].shape == (8, 36, 2048)
    for lo, hi in ((0, 2), (2, 3), (3, 8)):
        part = model(x[lo:hi].contiguous(), sizes[lo:hi], scales_yx=scales[lo:hi], padding="max_detections", return_tensors="np")
        for k in keys:
            assert np.array_equal(part[k], full[k][lo:hi]), (mode, lo, hi, k)
    if mode in EXACT_MODES:
        g = load_golden("cfg2x2")
        assert np.array_equal(full["obj_ids"][:2].reshape(-1), g["obj_ids"])
        assert np.array_equal(full["attr_ids"][:2].reshape(-1), g["attr_ids"])
        np.testing.assert_allclose(full["boxes"][:2].reshape(-1, 4), g["boxes"], rtol=0, atol=1e-2)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_maximum_canvas_1333x1333(mode):
    """Maximum sizes: a landscape 800x1333 and a portrait 1333x800 image in one batch pad to the largest canvas
    the 800/1333 resize rule can produce (1333x1333, res4 84x84 = 105 840 anchors per image).  Checked through
    properties: every image equals itself run alone on the same canvas (bit-exact), clipping respects each
    image's own (h, w), detections are full, and the result is reproducible after interleaved smaller calls."""
    from vltk_b200 import synthetic
    model, cfg = get_model("cfg1", mode)
    mean = torch.tensor(cfg.pixel_mean).view(3, 1, 1)
    x = torch.zeros(2, 3, 1333, 1333)
    shapes = [(800, 1333), (1333, 800)]
    for i, (h, w) in enumerate(shapes):
        x[i, :, :h, :w] = synthetic.make_raw_image(h, w, 60 + i).permute(2, 0, 1).float() - mean
    sizes = torch.tensor(shapes)
    scales = torch.tensor([[0.5, 0.5], [1.25, 1.25]])
    both = model(x, sizes, scales_yx=scales, padding="max_detections", return_tensors="np")
    assert both["preds_per_image"].tolist() == [36, 36]
    for i, (h, w) in enumerate(shapes):
        one = model(x[i:i + 1].contiguous(), sizes[i:i + 1], scales_yx=scales[i:i + 1], padding="max_detections", return_tensors="np")
        for k in ("obj_ids", "attr_ids", "boxes", "roi_features", "obj_probs", "keep_idx"):
            assert np.array_equal(one[k][0], both[k][i]), (mode, i, k)
        b = both["boxes"][i] / np.array([scales[i, 1], scales[i, 0], scales[i, 1], scales[i, 0]], np.float32)
        assert (b[:, 0::2] >= 0).all() and (b[:, 0::2] <= w + 1e-3).all()     # clipped to THIS image's width ...
        assert (b[:, 1::2] >= 0).all() and (b[:, 1::2] <= h + 1e-3).all()     # ... and height, not the canvas
        assert (both["normalized_boxes"][i] <= 1.0 + 1e-5).all()
    again = model(x, sizes, scales_yx=scales, padding="max_detections", return_tensors="np")
    assert np.array_equal(again["roi_features"], both["roi_features"])      # reproducible after the smaller calls above


def test_forward_stream_equals_forward():
    """The pipelined public API (H2D / compute / D2H of neighbouring batches overlapped on three streams)
    must return, in order, exactly what one synchronous forward() per batch returns — including across
    batches of different canvas shapes (slot buffers are reallocated) and more batches than pipeline depth."""
    from vltk_b200.preprocess import Preprocess
    model, cfg = get_model("mixed", "fp32")
    pre = Preprocess(cfg)
    host = []
    for name in ("mixed", "tiny", "mixed", "stripes", "tiny", "mixed", "constant"):
        _, _, raws = cases.case_inputs(name)
        _, images, sizes, scales = pre(raws)
        host.append((images.cpu(), sizes, scales if name != "stripes" else None))   # one batch without scales_yx
    ref = [model(x, sz, scales_yx=sc, padding="max_detections", return_tensors="np") for x, sz, sc in host]
    for depth, cstreams in ((1, 1), (2, 1), (3, 1), (1, 2), (2, 2), (3, 2), (4, 3)):
        got = list(model.forward_stream(iter(host), depth=depth, compute_streams=cstreams))
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            for k in ("obj_ids", "attr_ids", "boxes", "normalized_boxes", "obj_probs", "attr_probs", "roi_features",
                      "preds_per_image", "sizes", "keep_idx"):
                assert np.array_equal(a[k], b[k]), (depth, cstreams, k)


def test_batches_in_flight_on_two_streams_are_bit_identical_at_full_size():
    """The benchmarked configuration keeps TWO batches in flight (bench.py --streams 2; FRCNN.forward_stream's
    compute_streams): batch i runs on stream i % 2 with its own workspace and its own backbone side stream, so one
    batch's few-CTA selection kernels overlap the other's convolutions.  At BASELINE size (8 x 600x1000, bf16 mode:
    CTA pairs, two backbone halves) every batch must come out bit-identical to the same batch run alone."""
    from vltk_b200 import synthetic
    model, cfg = get_model("cfg2x2", "bf16")
    mean = torch.tensor(cfg.pixel_mean).view(1, 3, 1, 1)
    sizes = np.tile(np.array([[600, 1000]], np.int32), (8, 1))
    scales = np.ones((8, 2), np.float32)
    ro = model.roi_outputs
    keys = ("obj_ids", "attr_ids", "boxes", "obj_probs", "attr_probs", "roi_features", "preds_per_image", "keep_idx")
    dev = model.device
    xs = []
    for b in range(3):
        raws = [synthetic.make_raw_image(600, 1000, 5000 + 8 * b + i) for i in range(8)]
        xs.append((torch.stack([r.permute(2, 0, 1).float() for r in raws]) - mean).contiguous().to(dev))
    alone = []
    for x in xs:
        t = model.run(x, sizes, scales, ro.max_detections, ro.min_detections, ro.nms_thresh)
        torch.cuda.synchronize(dev)
        alone.append({k: t[k].cpu().numpy() for k in keys})
    side = torch.cuda.Stream(device=dev)
    for rep in range(2):
        side.wait_stream(torch.cuda.current_stream(dev))
        outs = []
        for i in range(6):                                  # batches 0,1,2,0,1,2 alternating between the two streams
            if i % 2 == 0:
                outs.append(model.run(xs[i % 3], sizes, scales, ro.max_detections, ro.min_detections, ro.nms_thresh, slot=0))
            else:
                with torch.cuda.stream(side):
                    outs.append(model.run(xs[i % 3], sizes, scales, ro.max_detections, ro.min_detections, ro.nms_thresh, slot=1))
        torch.cuda.synchronize(dev)
        for i, t in enumerate(outs):
            for k in keys:
                assert np.array_equal(t[k].cpu().numpy(), alone[i % 3][k]), (rep, i, k)


def test_two_devices_in_one_process():
    """The C ABI allows one process to own engines on several GPUs: the >48 KB shared-memory opt-in of the
    big kernels is per device and must be repeated on each (a per-process flag would make the second device's
    first launch fail)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from vltk_b200.frcnn import FRCNN
    from vltk_b200.preprocess import Preprocess
    cfg = cases.case_config("tiny")
    _, _, raws = cases.case_inputs("tiny")
    outs = []
    for dev in (0, 1):
        for mode in ("bf16", "fp32"):
            m = FRCNN.from_pretrained(state_dict=weights(0), config=cfg, mode=mode, device=dev)
            _, images, sizes, scales = Preprocess(cfg, device=dev)(raws)
            o = m(images, sizes, scales_yx=scales, padding="max_detections", return_tensors="np")
            outs.append((dev, mode, o))
    for mode in ("bf16", "fp32"):
        a = next(o for d, m_, o in outs if d == 0 and m_ == mode)
        b = next(o for d, m_, o in outs if d == 1 and m_ == mode)
        for k in ("obj_ids", "attr_ids", "boxes", "roi_features", "preds_per_image"):
            assert np.array_equal(a[k], b[k]), (mode, k)    # same kernels, same inputs: bit-identical across GPUs


def test_error_behaviour():
    from vltk_b200 import _lib
    from vltk_b200.frcnn import FRCNN
    model, cfg = get_model("tiny", "fp32")
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(_lib.LibraryError):  # max_detections above the proposal budget
        model(x, torch.tensor([[64, 64]]), max_detections=10 ** 6)
    with pytest.raises(_lib.LibraryError):  # image_shapes outside the padded canvas
        model(x, torch.tensor([[65, 64]]))
    with pytest.raises(NotImplementedError):
        model(x, torch.tensor([[64, 64]]), proposals=[torch.zeros(1, 4)])
    m2 = FRCNN(cfg, mode="fp32")
    with pytest.raises(_lib.LibraryError):  # strict state_dict, like load_state_dict(strict=True)
        m2.load_state_dict({"backbone.stem.conv1.weight": torch.zeros(64, 3, 7, 7)})
