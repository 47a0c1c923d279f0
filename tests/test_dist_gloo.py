"""The N>1 path on CPU: two gloo ranks shard images by index with no data-path collective,
and the optional single-file mode does exactly one gather to the writer rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakePreprocess:
    def __call__(self, raws, ids):
        n = len(raws)
        return list(ids), torch.stack([r.float().mean().expand(3, 4, 4) for r in raws]), \
            torch.full((n, 2), 4), torch.ones(n, 2)


class FakeModel:
    """Deterministic stand-in with the FRCNN.forward contract (the CUDA engine needs a GPU)."""
    md, d = 4, 8

    def __call__(self, images, sizes, scales_yx=None, padding=None, return_tensors=None, **kw):
        n = images.shape[0]
        v = images.reshape(n, -1)[:, 0].numpy()
        base = v[:, None, None] + np.arange(self.md)[None, :, None] + np.zeros((1, 1, self.d))
        return dict(boxes=base[:, :, :4].astype(np.float32), normalized_boxes=base[:, :, :4].astype(np.float32) / 4,
                    obj_ids=(v[:, None] + np.arange(self.md)).astype(np.int64), obj_probs=np.ones((n, self.md), np.float32),
                    attr_ids=np.zeros((n, self.md), np.int64), attr_probs=np.ones((n, self.md), np.float32),
                    roi_features=base.astype(np.float32), preds_per_image=np.full(n, self.md))


def _source(i):
    return np.full((5, 6, 3), i, np.uint8)


def _worker(rank, world, port, out_dir, single):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from vltk_b200.extract import extract
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = [f"img{i}" for i in range(11)]
    extract(_source, ids, FakeModel(), FakePreprocess(), out_dir, batch_size=3, rank=rank, world=world,
            single_file=single)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("single", [False, True])
def test_two_rank_sharded_extraction(tmp_path, single):
    pytest.importorskip("pyarrow")
    from vltk_b200.extract import extract, read_arrow
    port = 29500 + (os.getpid() % 2000) + (1 if single else 0)
    mp.spawn(_worker, args=(2, port, str(tmp_path), single), nprocs=2, join=True)
    ref_dir = tmp_path / "ref"
    ids = [f"img{i}" for i in range(11)]
    extract(_source, ids, FakeModel(), FakePreprocess(), str(ref_dir), batch_size=4)
    ref, _ = read_arrow(str(ref_dir / "train.arrow"))
    ref_rows = {r["imgid"]: r for r in ref.to_pylist()}
    if single:
        got, meta = read_arrow(str(tmp_path / "train.arrow"))
        rows = got.to_pylist()
        assert [r["imgid"] for r in rows] == ids          # global index order restored
    else:
        rows = []
        for r in range(2):
            t, _ = read_arrow(str(tmp_path / f"train.rank{r}.arrow"))
            part = t.to_pylist()
            assert [x["imgid"] for x in part] == ids[r::2]   # i mod world == rank
            rows += part
    assert len(rows) == 11
    for r in rows:
        assert r["features"] == ref_rows[r["imgid"]]["features"]
        assert r["object_ids"] == ref_rows[r["imgid"]]["object_ids"]
