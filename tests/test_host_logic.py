"""CPU tests of the host-side logic: shape arithmetic, synthetic recipe, FLOP model, the
oracle's restated third-party ops vs the real torchvision, sharding and the Arrow contract."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import frcnn_oracle as O
from vltk_b200 import arch, synthetic
from vltk_b200.config import FRCNNConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [7, 8, 150, 151, 300, 333, 400, 667, 800, 1333])
def test_shape_arithmetic_matches_torch_ops(n):
    cfg = FRCNNConfig()
    x = torch.zeros(1, 1, n, 9)
    s = F.conv2d(x, torch.zeros(1, 1, 7, 7), stride=2, padding=3)
    assert s.shape[2] == cfg.stem_conv_out(n)
    p = F.max_pool2d(s, 3, 2, 0, ceil_mode=True)
    assert p.shape[2] == cfg.stem_pool_out(s.shape[2])
    assert F.conv2d(p, torch.zeros(1, 1, 1, 1), stride=2).shape[2] == cfg.stride2_out(p.shape[2])


def test_res4_sizes_of_the_baseline_configs():
    cfg = FRCNNConfig()
    assert cfg.res4_hw(800, 1333) == (50, 84)
    assert cfg.res4_hw(600, 1000) == (38, 63)


def test_flop_model_matches_baseline_md():
    cfg = FRCNNConfig()
    f1 = arch.flops_per_image(cfg, 800, 1333, 300)
    f2 = arch.flops_per_image(cfg, 600, 1000, 300)
    assert abs(f1["total"] / 1e9 - 2100.26) < 0.5
    assert abs(f2["total"] / 1e9 - 1956.75) < 0.5
    assert abs(f1["res5"] / 1e9 - 1757.20) < 0.1


def test_state_dict_layout_and_determinism():
    cfg = FRCNNConfig()
    a = synthetic.make_state_dict(cfg, 0)
    b = synthetic.make_state_dict(cfg, 0)
    assert len(a) == 640
    assert sum(v.numel() for k, v in a.items() if "running" not in k and "tracked" not in k) == 65447577
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["roi_heads.box_predictor.cls_score.bias"].abs().sum() > 0   # calibrated bias shipped


def test_resize_rule():
    cfg = FRCNNConfig()
    assert synthetic.resized_hw(800, 1333, cfg) == (800, 1333)
    assert synthetic.resized_hw(480, 640, cfg) == (800, 1067)
    assert synthetic.resized_hw(375, 1242, cfg) == (402, 1333)   # long side capped
    assert synthetic.resized_hw(1000, 750, cfg) == (1067, 800)


def test_oracle_nms_and_roipool_match_torchvision():
    tv = pytest.importorskip("torchvision")
    g = torch.Generator().manual_seed(0)
    xy = torch.rand(400, 2, generator=g) * 300
    boxes = torch.cat([xy, xy + torch.rand(400, 2, generator=g) * 90 + 1], 1)
    boxes[5] = boxes[4]                      # duplicate
    boxes[9, 2:] = boxes[9, :2]              # zero area
    scores = torch.rand(400, generator=g)
    scores[5] = scores[4]
    for thr in (0.3, 0.7):
        ref = tv.ops.nms(boxes, scores, thr).numpy()
        assert np.array_equal(O.nms_np(boxes.numpy(), scores.numpy(), thr), ref)
        assert np.array_equal(O.nms_np(boxes.numpy(), scores.numpy(), thr, 25), ref[:25])
    feat = torch.randn(2, 8, 19, 27, generator=g)
    rois = torch.cat([torch.randint(0, 2, (40, 1), generator=g).float(), boxes[:40] * 1.4 - 20], 1)
    assert torch.equal(O.roi_pool_np(feat, rois, 14, 1 / 16), tv.ops.RoIPool((14, 14), 1 / 16)(feat, rois))


def test_oracle_preprocess_matches_reference_when_present():
    from oracle import cases, ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    import warnings
    cfg, _, raws = cases.case_inputs("mixed")
    _, compat = ref_loader.load_reference()
    pre = ref_loader.load_preprocess()(compat.Config(cfg.to_reference_dict()))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(0)
        _, images, sizes, scales = pre([r.clone() for r in raws], [0, 1])
    oi, osz, osc = O.preprocess(cfg, raws)
    assert torch.equal(images, oi) and torch.equal(sizes, osz) and torch.equal(scales, osc)


def test_shard_indices_partition():
    from vltk_b200.extract import shard_indices
    for n, w in [(0, 1), (5, 2), (17, 4), (5000, 8)]:
        parts = [shard_indices(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_arrow_contract_roundtrip_and_reference_fixture(tmp_path):
    """Our writer produces the reference's columns (adapters/frcnn.py:35-41) in an Arrow IPC
    stream with its metadata keys; the reference's own fixture (tests/visualgenome/frcnn/
    train.arrow: 10 rows, 36 boxes) pins those names/shapes when the tree is present."""
    pa = pytest.importorskip("pyarrow")
    from vltk_b200.extract import _rows, read_arrow, write_arrow
    n, md, d = 3, 36, 2048
    rng = np.random.default_rng(0)
    dense = dict(boxes=rng.random((n, md, 4), np.float32) * 500, normalized_boxes=rng.random((n, md, 4), np.float32),
                 obj_ids=rng.integers(0, 1600, (n, md)), obj_probs=rng.random((n, md), np.float32),
                 attr_ids=rng.integers(0, 400, (n, md)), attr_probs=rng.random((n, md), np.float32),
                 roi_features=rng.random((n, md, d), np.float32), preds_per_image=np.full(n, md))
    cols = _rows(["a", "b", "c"], dense, np.full((n, 2), 600), np.ones((n, 2), np.float32))
    path = str(tmp_path / "train.arrow")
    write_arrow(path, cols, {"dataset": "t", "model_config": {}, "processor_args": {}})
    table, meta = read_arrow(path)
    assert table.num_rows == n
    for c in ("imgid", "attr_ids", "object_ids", "features", "box"):
        assert c in table.column_names
    assert {"img_to_row_map", "model_config", "dataset", "processor_args"} <= set(meta)
    assert json.loads(meta["img_to_row_map"]) == {"a": 0, "b": 1, "c": 2}
    f = np.asarray(table.column("features").to_pylist(), np.float32)
    assert f.shape == (n, md, d) and np.array_equal(f, dense["roi_features"])
    fixture = "/root/reference/tests/visualgenome/frcnn/train.arrow"
    if os.path.exists(fixture):
        ref, rmeta = read_arrow(fixture)
        assert {"attr_ids", "box", "features", "imgid", "object_ids"} <= set(ref.column_names)
        assert {"img_to_row_map", "model_config", "dataset", "processor_args"} <= set(rmeta)
        row = ref.slice(0, 1).to_pylist()[0]
        assert np.asarray(row["features"]).shape == (36, 2048) and np.asarray(row["box"]).shape == (36, 4)
        assert len(row["attr_ids"]) == 36 and len(row["object_ids"]) == 36


def test_oracle_ignorey_rules():
    """apply_ignorey (frcnn.py:328-366): span -> dropped; nearer end clipped with int() truncation; boxes lying
    entirely past the range untouched; ranges are divided by the X scale as the reference writes it."""
    from oracle import frcnn_oracle as O
    boxes = torch.tensor([[0., 10., 5., 90.],     # spans [40.5, 60.5] -> dropped
                          [0., 45., 5., 58.],     # inside: |60.5-58| < |40.5-45| -> y2 = int(40.5) = 40
                          [0., 41., 5., 50.],     # |40.5-41| < |60.5-50| -> y1 = int(60.5) = 60
                          [0., 70., 5., 95.],     # y1 > r1 and y2 > r0 -> untouched
                          [0., 5., 5., 30.]])     # before the range, still 'to clip': |60.5-30| < |40.5-5| -> y2 = 40 (sic)
    out, alive = O.apply_ignorey(boxes, torch.tensor([[81.0, 121.0]]), torch.tensor(2.0))
    assert alive.tolist() == [False, True, True, True, True]
    assert out[1].tolist() == [0., 45., 5., 40.]
    assert out[2].tolist() == [0., 60., 5., 50.]
    assert out[3].tolist() == [0., 70., 5., 95.]
    assert out[4].tolist() == [0., 5., 5., 40.]


def test_oracle_timing_path_with_torchvision_ops_equals_the_restatement():
    """bench.py times the oracle with nms / RoIPool taken from torchvision itself (what the reference calls,
    frcnn.py:132, 383, 1179) instead of the slow restatements: both paths must give identical outputs."""
    from oracle import cases
    cfg, wseed, raws = cases.case_inputs("tiny")
    sd = synthetic.make_state_dict(FRCNNConfig(), wseed)
    images, sizes, scales = O.preprocess(cfg, raws)
    a = O.forward(sd, cfg, images, sizes, scales)
    O.use_torchvision_ops(True)
    try:
        b = O.forward(sd, cfg, images, sizes, scales, res5_chunk=1 << 30)
    finally:
        O.use_torchvision_ops(False)
    for k in ("keep", "obj_ids", "attr_ids"):
        assert all(torch.equal(x, y) for x, y in zip(a[k], b[k])), k
    for k in ("boxes", "roi_features", "obj_probs"):
        for x, y in zip(a[k], b[k]):
            np.testing.assert_allclose(x.numpy(), y.numpy(), rtol=1e-5, atol=1e-5)


def test_our_arrow_files_load_through_the_reference_adapter(tmp_path):
    """The consumer side of the drop-in: a `{split}.arrow` written by vltk_b200.extract must load through the
    reference's OWN loader — Adapter.load -> _load_one_arrow -> datasets.Dataset(arrow_table) (abc/adapter.py:381-462)
    — and give back features / box / object_ids / attr_ids per image id via Adapter.get (adapter.py:186-196).
    Authoring container only (needs /root/reference); the reference's own fixture goes through the same code for
    comparison of the column types."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    pytest.importorskip("datasets")
    from vltk_b200.extract import _rows, write_arrow
    Adapter = ref_loader.load_adapter()

    class FRCNNFeatures(Adapter):
        _is_feature = True

        def forward(*a, **k):
            raise NotImplementedError

        def schema(*a, **k):
            return {}

        def _meta_names():
            return []

    rng = np.random.default_rng(0)
    n, md, d = 5, 36, 2048
    dense = dict(boxes=rng.random((n, md, 4), np.float32) * 300, normalized_boxes=rng.random((n, md, 4), np.float32),
                 obj_ids=rng.integers(0, 1600, (n, md)), obj_probs=rng.random((n, md), np.float32),
                 attr_ids=rng.integers(0, 400, (n, md)), attr_probs=rng.random((n, md), np.float32),
                 roi_features=rng.random((n, md, d), np.float32), preds_per_image=np.full(n, md))
    ids = [f"{1000 + i}" for i in range(n)]
    cols = _rows(ids, dense, np.tile([[600, 800]], (n, 1)))
    path = str(tmp_path / "train.arrow")
    write_arrow(path, cols, {"dataset": "synthetic", "model_config": {}, "processor_args": {"size": [800, 1333]}})
    ours = FRCNNFeatures.load(path)
    assert ours.img_to_row_map == {k: i for i, k in enumerate(ids)}
    # a plain-string value comes back as raw bytes, exactly like `dataset` in the reference's own file (adapter.py:400-404)
    assert ours.meta_dataset in ("synthetic", b"synthetic") and ours.meta_processor_args == {"size": [800, 1333]}
    for i, k in enumerate(ids):
        row = ours.get(k)
        assert row["imgid"] == k
        np.testing.assert_array_equal(np.asarray(row["features"], np.float32), dense["roi_features"][i])
        np.testing.assert_array_equal(np.asarray(row["box"], np.float32), np.round(dense["boxes"][i]))
        assert row["object_ids"] == dense["obj_ids"][i].astype(np.float32).tolist()
        assert row["attr_ids"] == dense["attr_ids"][i].astype(np.float32).tolist()
    # the reference's own extracted-features file through the same loader: same feature types for its columns
    theirs = FRCNNFeatures.load(os.path.join(ref_loader.REF_ROOT, "tests", "visualgenome", "frcnn", "train.arrow"))
    for k in ("imgid", "attr_ids", "object_ids", "features", "box"):
        assert ours.features[k] == theirs.features[k], (k, ours.features[k], theirs.features[k])
        assert ours.data.schema.field(k).type == theirs.data.schema.field(k).type      # same Arrow storage type
    r0 = theirs.get(theirs.imgids[0])
    assert np.asarray(r0["features"], np.float32).shape == (36, 2048) and np.asarray(r0["box"], np.float32).shape == (36, 4)


def test_synthetic_images_do_not_depend_on_the_thread_count():
    """torchrun starts every rank with OMP_NUM_THREADS=1, and ATen's single-thread bilinear path rounds differently from
    the multi-thread one: the same seed must still give the golden's image (vltk_b200/synthetic.py::make_raw_image) —
    otherwise rank 0's in-run parity check of `bench.py --gpus N` compares against the wrong pixels."""
    import hashlib
    import os
    import subprocess
    import sys
    from vltk_b200 import synthetic
    code = ("import hashlib, sys; sys.path.insert(0, %r); from vltk_b200 import synthetic; "
            "print(hashlib.md5(synthetic.make_raw_image(150, 200, 4010).numpy().tobytes()).hexdigest())") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    here = hashlib.md5(synthetic.make_raw_image(150, 200, 4010).numpy().tobytes()).hexdigest()
    for n in ("1", "2"):
        env = dict(os.environ, OMP_NUM_THREADS=n)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env).stdout.strip().splitlines()[-1]
        assert out == here, (n, out, here)
