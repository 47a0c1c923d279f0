"""Rows a14 and f1 of SURVEY.md §8 on the real engine: the adapter hook produces the reference's
row dict, and the sharded/batched extraction driver writes Arrow files whose rows equal direct
model calls on the same batches (config-5 flavour, scaled down)."""
import json

import numpy as np
import pytest
import torch

from oracle import cases, frcnn_oracle as O
from tests.util import weights

pytestmark = pytest.mark.gpu

SIZES = [(150, 200), (240, 160), (192, 256), (170, 230), (200, 150)]


@pytest.fixture(scope="module")
def engine():
    from vltk_b200.frcnn import FRCNN
    cfg = cases.case_config("mixed")
    model = FRCNN.from_pretrained(state_dict=weights(0), config=cfg, mode="exact_tc")
    return model, cfg


def _source(i):
    from vltk_b200 import synthetic
    h, w = SIZES[i % len(SIZES)]
    return synthetic.make_raw_image(h, w, 500 + i).numpy()


def test_adapter_forward_row_matches_oracle(engine):
    """adapters/frcnn.py:43-64: one preprocessed image in -> {object_ids, attr_ids, box, features}."""
    from vltk_b200 import adapter
    from vltk_b200.preprocess import Preprocess
    model, cfg = engine
    raw = torch.from_numpy(_source(0))
    ids, images, sizes, scales = Preprocess(cfg)([raw])
    entry = {adapter.IMG: images[0], adapter.SIZE: sizes[0], adapter.SCALE: torch.tensor([1.0, 1.0])}
    row = adapter.FRCNN.forward(model, entry)
    oi, osz, osc = O.preprocess(cfg, [raw])
    ref = O.forward(weights(0), cfg, oi, osz, None)
    n = int(ref["preds_per_image"][0])
    md = cfg.max_detections
    assert len(row["object_ids"][0]) == md and len(row["attr_ids"][0]) == md
    assert row["object_ids"][0][:n] == ref["obj_ids"][0].tolist()
    assert row["attr_ids"][0][:n] == ref["attr_ids"][0].tolist()
    assert tuple(row[adapter.FEATURES][0].shape) == (md, 2048)
    np.testing.assert_allclose(row[adapter.FEATURES][0][:n].numpy(), ref["roi_features"][0].numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(np.asarray(row[adapter.BOX][0])[:n], torch.round(ref["boxes"][0]).numpy(), atol=1.0)
    assert set(adapter.FRCNN.schema()) == {"attr_ids", "object_ids", "features", "box"}


def test_extract_writes_arrow_rows_equal_to_direct_calls(engine, tmp_path):
    from vltk_b200.extract import extract, read_arrow
    from vltk_b200.preprocess import Preprocess
    model, cfg = engine
    pre = Preprocess(cfg)
    ids = [f"img{i:03d}" for i in range(7)]
    path = extract(_source, ids, model, pre, str(tmp_path), split="train", batch_size=3, bucket=False,
                   meta={"dataset": "synthetic", "model_config": {"max_detections": cfg.max_detections}})
    table, meta = read_arrow(path)
    assert table.num_rows == 7
    assert json.loads(meta["img_to_row_map"]) == {k: i for i, k in enumerate(ids)}
    rows = table.to_pylist()

    def check_against_direct_batches(rows, order, batch, batches=None):
        """Every Arrow row must equal a direct model call on THE SAME BATCH.  (Not on the image
        alone: like the reference's Preprocess.pad, a batch is zero-padded to its largest member,
        and the padded border legitimately changes features near the image edge.)"""
        for idx in (batches or [order[s0:s0 + batch] for s0 in range(0, len(order), batch)]):
            _, images, sizes, scales = pre([torch.from_numpy(_source(i)) for i in idx])
            d = model(images, sizes, scales_yx=scales, padding="max_detections", return_tensors="np")
            for j, i in enumerate(idx):
                r = rows[ids[i]]
                assert r["preds_per_image"] == int(d["preds_per_image"][j])
                assert r["object_ids"] == d["obj_ids"][j].astype(np.float32).tolist()
                assert r["attr_ids"] == d["attr_ids"][j].astype(np.float32).tolist()
                assert np.array_equal(np.asarray(r["features"], np.float32), d["roi_features"][j])
                assert np.array_equal(np.asarray(r["boxes"], np.float32), d["boxes"][j])
                assert np.array_equal(np.asarray(r["normalized_boxes"], np.float32), d["normalized_boxes"][j])
            # the reference column, computed the reference's way (adapters/frcnn.py:50-57): model WITHOUT scales_yx ->
            # boxes in resized pixels -> round(boxes * 1/wh_scale), wh_scale = resized (w,h) / raw (w,h) -> raw pixels
            d0 = model(images, sizes, padding="max_detections", return_tensors="np")
            for j, i in enumerate(idx):
                rh, rw = _source(i).shape[:2]
                sh, sw = (int(v) for v in sizes[j])
                inv = 1.0 / torch.tensor([sw / rw, sh / rh], dtype=torch.float32)
                ref_box = torch.round(torch.from_numpy(d0["boxes"][j].copy()) * torch.stack([inv[0], inv[1], inv[0], inv[1]])).numpy()
                got = np.asarray(rows[ids[i]]["box"], np.float32)
                np.testing.assert_allclose(got, ref_box, atol=1.0)     # the two float paths may round a x.5 differently
                assert (got == ref_box).mean() > 0.9

    assert [r["imgid"] for r in rows] == ids
    check_against_direct_batches({r["imgid"]: r for r in rows}, list(range(7)), 3)
    # sharding: rank r of 2 owns images i with i mod 2 == r and writes exactly those rows
    for rank in range(2):
        pth = extract(_source, ids, model, pre, str(tmp_path / "s"), batch_size=2, rank=rank, world=2, bucket=False)
        t, _ = read_arrow(pth)
        part = t.to_pylist()
        assert [r["imgid"] for r in part] == ids[rank::2]
        check_against_direct_batches({r["imgid"]: r for r in part}, list(range(rank, 7, 2)), 2)
    # default bucketing: batches are formed by plan_batches (equal resized sizes together); rows stay in id order
    from vltk_b200.extract import plan_batches
    pth = extract(_source, ids, model, pre, str(tmp_path / "b"), batch_size=3)
    t, _ = read_arrow(pth)
    rows_b = t.to_pylist()
    assert [r["imgid"] for r in rows_b] == ids
    plan = plan_batches([_source(i).shape[:2] for i in range(7)], cfg, 3, True)
    assert sorted(j for b in plan for j in b) == list(range(7))
    check_against_direct_batches({r["imgid"]: r for r in rows_b}, None, 3, batches=plan)


def test_single_file_gather_device_path_writes_the_same_file_as_the_direct_writer(engine, tmp_path):
    """single_file=True on GPUs: windows are packed on the device, gathered with NCCL into rank 0's device buffers, split
    into column-contiguous tensors on a copy stream and handed to the writer thread through cycling pinned buffers
    (vltk_b200/extract.py::_WindowGather).  A 1-rank NCCL group runs exactly that machinery on one GPU; the file must
    equal the direct writer's, column for column and bit for bit, with several windows (buffer reuse) and a partial last
    window."""
    import socket
    import torch.distributed as dist
    from vltk_b200.extract import extract, read_arrow
    from vltk_b200.preprocess import Preprocess
    model, cfg = engine
    pre = Preprocess(cfg)
    ids = [f"img{i:03d}" for i in range(21)]
    direct = extract(_source, ids, model, pre, str(tmp_path / "direct"), batch_size=4, window=8, bucket=False)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        gathered = extract(_source, ids, model, pre, str(tmp_path / "gather"), batch_size=4, window=8, bucket=False,
                           single_file=True, _force_gather=True)
        again = extract(_source, ids, model, pre, str(tmp_path / "gather2"), batch_size=4, window=8, bucket=False,
                        single_file=True, _force_gather=True)      # second job of the process: pinned sets are reused
    finally:
        dist.destroy_process_group()
    ta, ma = read_arrow(direct)
    for pth in (gathered, again):
        tb, mb = read_arrow(pth)
        assert tb.num_rows == ta.num_rows == 21 and tb.schema.names == ta.schema.names
        assert json.loads(mb["img_to_row_map"]) == json.loads(ma["img_to_row_map"])
        for name in ta.schema.names:
            assert tb.column(name).combine_chunks().equals(ta.column(name).combine_chunks()), name
