"""Reader fast path (vltk_b200/reader.py) against
  (a) the reference's own extracted-features fixture (first 3 rows, tests/golden/ref_frcnn_train_head3.arrow,
      derived by oracle/make_reader_fixture.py) with per-row checksums taken through pyarrow's per-row path, and
  (b) files written by this package's extractor layout (fixed-size lists, per-rank shards)."""
import json
import os

import numpy as np
import pytest

from tests.util import GOLD

FIX = os.path.join(GOLD, "ref_frcnn_train_head3.arrow")


def test_reads_the_reference_fixture_layout():
    from vltk_b200.reader import FeatureTable
    chk = json.load(open(FIX.replace(".arrow", ".json")))["checks"]
    t = FeatureTable.load(FIX)
    assert len(t) == 3 and t.n_imgs == 3
    assert t.row_shape("features") == [36, 2048] and t.row_shape("box") == [36, 4]
    # metadata keys of extraction.py:230-233 decoded like adapter.py:395-407
    assert t.meta_processor_args["size"] == [800, 1333] and t.meta_model_config in (None, "None")
    assert not hasattr(t, "meta_huggingface")
    for r, c in chk.items():
        row = t.get(c["imgid"])                          # Adapter.get: img_to_row_map lookup
        assert t.get_idx(c["imgid"]) == int(r) and t.has_id(c["imgid"])
        for name in ("features", "box", "attr_ids", "object_ids"):
            a = np.asarray(row[name], np.float64)
            assert list(a.shape) == c[name]["shape"]
            assert a.sum() == pytest.approx(c[name]["sum"], rel=1e-12)
            assert np.abs(a).sum() == pytest.approx(c[name]["abs_sum"], rel=1e-12)
            assert a.reshape(-1)[0] == c[name]["first"] and a.reshape(-1)[-1] == c[name]["last"]
    # batch gather across record batches (the fixture has one row per batch), any order, repeats
    f = t.features([2, 0, 2, 1])
    assert f.shape == (4, 36, 2048) and f.dtype == np.float32
    assert np.array_equal(f[0], f[2]) and np.array_equal(f[1], t.get(chk["0"]["imgid"])["features"])
    # one contiguous slice of one record batch is a zero-copy view of the mapped file
    v = t.features([1])
    assert not v.flags.owndata and not v.flags.writeable
    with pytest.raises(IndexError):
        t.features([3])
    assert not t.has_id("nope")
    with pytest.raises(KeyError):
        t.get("nope")


def test_reads_this_packages_writer_layout_and_rank_shards(tmp_path):
    from vltk_b200 import extract
    from vltk_b200.reader import FeatureTable
    rng = np.random.default_rng(0)

    def shard(ids):
        n = len(ids)
        dense = dict(boxes=rng.random((n, 36, 4), np.float32) * 100, normalized_boxes=rng.random((n, 36, 4), np.float32),
                     obj_ids=rng.integers(0, 1600, (n, 36)), obj_probs=rng.random((n, 36), np.float32),
                     attr_ids=rng.integers(0, 400, (n, 36)), attr_probs=rng.random((n, 36), np.float32),
                     roi_features=rng.random((n, 36, 2048), np.float32), preds_per_image=np.full(n, 36))
        return extract._rows(ids, dense, np.tile([[600, 1000]], (n, 1)), np.ones((n, 2), np.float32))

    ids = [f"img{i}" for i in range(300)]                 # > 128 rows: several record batches per file
    a, b = shard(ids[0::2]), shard(ids[1::2])
    pa_, pb_ = str(tmp_path / "train.rank0.arrow"), str(tmp_path / "train.rank1.arrow")
    extract.write_arrow(pa_, a, {"dataset": "synthetic", "model_config": {}, "processor_args": {}})
    extract.write_arrow(pb_, b, {"dataset": "synthetic", "model_config": {}, "processor_args": {}})
    t = FeatureTable.load_many([pa_, pb_])
    assert len(t) == 300 and t.n_imgs == 300 and t.meta_dataset == "synthetic"
    want = {**{k: a["features"][i] for i, k in enumerate(ids[0::2])}, **{k: b["features"][i] for i, k in enumerate(ids[1::2])}}
    pick = [ids[i] for i in rng.permutation(300)[:64]]
    rows = t.rows_of(pick)
    got = t.features(rows)
    for k, g in zip(pick, got):
        assert np.array_equal(g, want[k])
    assert np.array_equal(t.get("img7")["features"], want["img7"])
    assert t.get("img7")["imgid"] == "img7" and t.get("img8")["preds_per_image"] == 36
    # a slice inside one 128-row record batch: zero copy
    assert not t.features(np.arange(130, 140)).flags.owndata
    # pinned staging for the H2D copy
    import torch
    p = t.pinned("features", rows[:5])
    assert p.is_pinned() == torch.cuda.is_available() and p.shape == (5, 36, 2048)
    assert np.array_equal(p.numpy(), got[:5])


def test_device_column_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vltk_b200 import _lib
    from vltk_b200.reader import FeatureTable
    with pytest.raises(_lib.LibraryError):
        FeatureTable.load(FIX).to_device("features")


@pytest.mark.gpu
def test_device_resident_column_gather_is_bit_exact():
    import torch
    from vltk_b200.reader import FeatureTable
    t = FeatureTable.load(FIX)
    col = t.to_device("features")
    rows = [2, 0, 1, 1, 2]
    got = col.gather(rows)
    assert got.is_cuda and got.shape == (5, 36, 2048)
    assert np.array_equal(got.cpu().numpy(), t.features(rows))
    box = t.to_device("box")
    assert np.array_equal(box.gather([1, 2]).cpu().numpy(), t.column("box", [1, 2]))
    with pytest.raises(IndexError):
        col.gather([3])
    # a larger synthetic table: 4096 rows x 8192 floats, shuffled gather of 1024 rows
    g = torch.Generator().manual_seed(0)
    col.data = torch.randn(4096, 8192, generator=g).to(col.device)
    col.n_rows, col.width, col.shape = 4096, 8192, [8192]
    idx = torch.randperm(4096, generator=g)[:1024]
    out = col.gather(idx.numpy())
    assert torch.equal(out, col.data[idx.to(col.device)])
