"""Sharded, batched replacement of the reference's extraction driver loop
(vltk/abc/extraction.py:94-248) for the FRCNN path: images are sharded by index across
ranks (one process per GPU, no collective in the hot path), run through the model in
batches, and written as Arrow IPC *stream* files with the reference's columns
(vltk/adapters/frcnn.py:35-41; `imgid` from vltk/abc/adapter.py:42) plus the wider model-dict
columns, and the reference's schema metadata keys (extraction.py:230-233).

Per-rank shard files `{split}.rank{r}.arrow` are the default; `single_file=True` gathers the
fixed-size tensors to rank 0 with ONE torch.distributed gather (NCCL over NVLink on GPUs,
gloo in the CPU tests) and writes `{split}.arrow`.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

FEATURE_KEYS = ("roi_features", "boxes", "normalized_boxes", "obj_ids", "obj_probs", "attr_ids",
                "attr_probs", "preds_per_image")


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Rank r owns items i with i mod world == r."""
    return list(range(rank, n_items, world))


def _rows(ids: Sequence[str], dense: Dict[str, np.ndarray], sizes, scales_yx):
    """Dense model outputs -> column arrays, reference columns first."""
    n = len(ids)
    boxes = np.asarray(dense["boxes"], np.float32)
    # adapter epilogue: round(boxes / wh_scale) (adapters/frcnn.py:57, utils/adapters.py:205-216);
    # boxes are already multiplied by scales_yx, dividing returns resized-image pixels
    sc = np.asarray(scales_yx, np.float32).reshape(n, 2)
    box = boxes.copy()
    box[:, :, 0::2] /= sc[:, None, 1:2]
    box[:, :, 1::2] /= sc[:, None, 0:1]
    box = np.round(box)
    cols = {
        "imgid": np.asarray([str(i) for i in ids], dtype=object),
        "attr_ids": np.asarray(dense["attr_ids"]).astype(np.float32),
        "object_ids": np.asarray(dense["obj_ids"]).astype(np.float32),
        "features": np.asarray(dense["roi_features"], np.float32),
        "box": box.astype(np.float32),
        "boxes": boxes,
        "normalized_boxes": np.asarray(dense["normalized_boxes"], np.float32),
        "obj_probs": np.asarray(dense["obj_probs"], np.float32),
        "attr_probs": np.asarray(dense["attr_probs"], np.float32),
        "preds_per_image": np.asarray(dense["preds_per_image"]).astype(np.int32),
        "sizes": np.asarray(sizes).astype(np.int32).reshape(n, 2),
    }
    return cols


def _fixed(arr: np.ndarray):
    """[n, a, b] / [n, a] float/int arrays -> nested fixed-size-list Arrow arrays."""
    import pyarrow as pa
    flat = pa.array(np.ascontiguousarray(arr).reshape(-1))
    out = flat
    for dim in reversed(arr.shape[1:]):
        out = pa.FixedSizeListArray.from_arrays(out, int(dim))
    return out


def write_arrow(path: str, cols: Dict[str, np.ndarray], meta: Dict[str, object]):
    """Arrow IPC stream, one record batch per 128 rows like the reference's flush cadence
    (extraction.py:26, 206-219); metadata values are json strings (utils/base.py:71-88)."""
    import pyarrow as pa
    n = len(cols["imgid"])
    arrays, names = [], []
    for k, v in cols.items():
        names.append(k)
        if k == "imgid":
            arrays.append(pa.array([str(x) for x in v], type=pa.string()))
        elif v.ndim == 1:
            arrays.append(pa.array(v))
        else:
            arrays.append(_fixed(v))
    table = pa.Table.from_arrays(arrays, names=names)
    md = {k: (v if isinstance(v, str) else json.dumps(v)) for k, v in meta.items()}
    md["img_to_row_map"] = json.dumps({str(i): r for r, i in enumerate(cols["imgid"])})
    table = table.replace_schema_metadata(md)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with pa.OSFile(path, "wb") as sink:
        with pa.ipc.new_stream(sink, table.schema) as w:
            for b in table.to_batches(max_chunksize=128):
                w.write_batch(b)
    return n


def read_arrow(path: str):
    """What Adapter._load_one_arrow does (vltk/abc/adapter.py:381-409): open the IPC stream,
    read everything, decode the schema metadata."""
    import pyarrow as pa
    with pa.memory_map(path, "r") as src:
        table = pa.ipc.open_stream(src).read_all()
    meta = {k.decode(): v.decode() for k, v in (table.schema.metadata or {}).items()}
    return table, meta


def extract(image_source: Callable[[int], np.ndarray], image_ids: Sequence[str], model, preprocess,
            out_dir: str, split: str = "train", batch_size: int = 8, rank: int = 0, world: int = 1,
            single_file: bool = False, max_detections: Optional[int] = None,
            meta: Optional[dict] = None, progress: Optional[Callable[[int], None]] = None) -> Optional[str]:
    """Runs this rank's shard.  image_source(i) -> raw BGR u8 [h,w,3] for global index i.
    Returns the path written by this rank (None on non-writer ranks with single_file)."""
    mine = shard_indices(len(image_ids), rank, world)
    chunks: List[Dict[str, np.ndarray]] = []
    for s in range(0, len(mine), batch_size):
        idx = mine[s:s + batch_size]
        raws = [torch.as_tensor(image_source(i)) for i in idx]
        ids, images, sizes, scales = preprocess(raws, [image_ids[i] for i in idx])
        kw = {} if max_detections is None else {"max_detections": max_detections}
        dense = model(images, sizes, scales_yx=scales, padding="max_detections", return_tensors="np", **kw)
        chunks.append(_rows(ids, dense, np.asarray(sizes), np.asarray(scales)))
        if progress:
            progress(len(idx))
    keys = list(chunks[0].keys()) if chunks else []
    cols = {k: np.concatenate([c[k] for c in chunks], 0) for k in keys}
    meta = dict(meta or {})
    meta.setdefault("dataset", "synthetic")
    meta.setdefault("model_config", {})
    meta.setdefault("processor_args", {})
    if not single_file or world == 1:
        name = f"{split}.arrow" if world == 1 else f"{split}.rank{rank}.arrow"
        path = os.path.join(out_dir, name)
        if keys:
            write_arrow(path, cols, meta)
        return path
    return _gather_and_write(cols, keys, mine, len(image_ids), out_dir, split, rank, world, meta)


def _gather_and_write(cols, keys, mine, n_total, out_dir, split, rank, world, meta):
    """The one collective of the path: fixed-size per-image tensors -> writer rank 0."""
    import torch.distributed as dist
    assert dist.is_initialized(), "single_file=True needs torch.distributed"
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    per_rank = -(-n_total // world)  # pad every rank to the same row count
    gathered = {}
    for k in keys:
        if k == "imgid":
            continue
        a = torch.from_numpy(np.ascontiguousarray(cols[k])) if len(mine) else None
        shape = (per_rank,) + tuple(a.shape[1:]) if a is not None else None
        # shapes are identical on every rank except for the row count; broadcast them from rank 0
        meta_t = [shape, str(a.dtype) if a is not None else None]
        lst = [None] * world
        dist.all_gather_object(lst, meta_t)
        shape, dt = next((s, d) for s, d in lst if s is not None)
        buf = torch.zeros(shape, dtype=getattr(torch, dt.split(".")[-1]), device=dev)
        if a is not None:
            buf[: a.shape[0]] = a.to(dev)
        out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, out, dst=0)
        if rank == 0:
            gathered[k] = [o.cpu().numpy() for o in out]
    ids_all = [None] * world
    dist.all_gather_object(ids_all, [str(x) for x in cols.get("imgid", [])])
    if rank != 0:
        return None
    # interleave back to global index order: global i lives at rank i % world, row i // world
    order = [(i % world, i // world) for i in range(n_total)]
    final = {"imgid": np.asarray([ids_all[r][j] for r, j in order], dtype=object)}
    for k, parts in gathered.items():
        final[k] = np.stack([parts[r][j] for r, j in order], 0)
    path = os.path.join(out_dir, f"{split}.arrow")
    write_arrow(path, final, meta)
    return path
