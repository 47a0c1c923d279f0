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
    # adapter epilogue (adapters/frcnn.py:50-57, utils/adapters.py:205-216): the reference calls the model WITHOUT
    # scales_yx (boxes in resized-image pixels) and writes round(boxes * 1/wh_scale) with wh_scale = resized/raw
    # (processing/image.py:128-135), i.e. RAW-image pixels.  `boxes` here was already multiplied by
    # scales_yx = raw/resized by the model (frcnn.py:1280-1283), so it is in that frame: only the rounding is left.
    box = np.round(boxes)
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


class _AsyncArrowWriter:
    """Arrow IPC stream written by a background thread while the GPU keeps working: `put(cols)` appends rows (one
    record batch per 128 rows, the reference's flush cadence, extraction.py:26); the schema metadata —
    including img_to_row_map, which Arrow stores in the schema message at the START of the stream — is fixed up
    front, so rows must arrive in the announced id order."""

    def __init__(self, path: str, ids_in_order: Sequence[str], meta: Dict[str, object]):
        import queue
        import threading
        self.path, self.meta = path, dict(meta)
        self.meta["img_to_row_map"] = {str(i): r for r, i in enumerate(ids_in_order)}
        self.q = queue.Queue(maxsize=4)
        self.err = None
        self.rows = 0
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        import pyarrow as pa
        sink = writer = None
        try:
            while True:
                cols = self.q.get()
                if cols is None:
                    break
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
                if writer is None:
                    md = {k: (v if isinstance(v, str) else json.dumps(v)) for k, v in self.meta.items()}
                    schema = table.schema.with_metadata(md)
                    os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
                    sink = pa.OSFile(self.path, "wb")
                    writer = pa.ipc.new_stream(sink, schema)
                for b in table.to_batches(max_chunksize=128):
                    writer.write_batch(b)
                self.rows += table.num_rows
        except Exception as e:  # surfaced by close()
            self.err = e
            while self.q.get() is not None:
                pass
        finally:
            if writer is not None:
                writer.close()
            if sink is not None:
                sink.close()

    def put(self, cols):
        if self.err:
            raise self.err
        self.q.put(cols)

    def close(self):
        self.q.put(None)
        self.t.join()
        if self.err:
            raise self.err
        return self.rows


def _stream(model, preprocess, batches, group, **kw):
    """The model's pipelined raw-image stream API when it has one (vltk_b200.frcnn.FRCNN); otherwise — duck-typed
    models with just the reference's forward contract — one synchronous preprocess + forward per batch."""
    if hasattr(model, "forward_raw_stream"):
        yield from model.forward_raw_stream(batches, preprocess, group=group, **kw)
        return
    for raws in batches:
        ids, images, sizes, scales = preprocess([torch.as_tensor(r) for r in raws], list(range(len(raws))))
        dense = dict(model(images, sizes, scales_yx=scales, padding="max_detections", return_tensors="np", **kw))
        dense["sizes"], dense["scales_yx"] = np.asarray(sizes), np.asarray(scales)
        yield dense


def plan_batches(raw_hw: Sequence[tuple], cfg, batch_size: int, bucket: bool = True) -> List[List[int]]:
    """Batches (lists of positions) for one window of images with raw sizes `raw_hw`.  bucket=False: index order.
    bucket=True: positions sorted (stably) by the resized (h/w, h, w), so images of equal size share a batch and
    mixed batches pad as little as possible."""
    from .synthetic import resized_hw
    order = list(range(len(raw_hw)))
    if bucket and cfg is not None:
        key = []
        for h, w in raw_hw:
            nh, nw = resized_hw(int(h), int(w), cfg)
            key.append((nh / nw, nh, nw))
        order.sort(key=lambda j: key[j])
    return [order[k:k + batch_size] for k in range(0, len(order), batch_size)]


def _entry_hw(x):
    """(h, w) of a raw entry: decoded [h,w,3] array/tensor, or an encoded JPEG (header parse only)."""
    if isinstance(x, (bytes, bytearray, memoryview)):
        from . import jpeg
        inf = jpeg.parse(bytes(x))
        return (inf.width, inf.height) if inf.orientation in (5, 6, 7, 8) else (inf.height, inf.width)
    return int(x.shape[0]), int(x.shape[1])


def extract(image_source: Callable[[int], np.ndarray], image_ids: Sequence[str], model, preprocess,
            out_dir: str, split: str = "train", batch_size: int = 8, rank: int = 0, world: int = 1,
            single_file: bool = False, max_detections: Optional[int] = None,
            meta: Optional[dict] = None, progress: Optional[Callable[[int], None]] = None,
            window: int = 64, bucket: bool = True) -> Optional[str]:
    """Runs this rank's shard.  image_source(i) -> raw BGR u8 [h,w,3] (array / tensor) or the encoded JPEG bytes
    of global index i.  Returns the path written by this rank (None on non-writer ranks with single_file).

    The shard is processed in windows of `window` images.  Inside a window the images are batched by
    `plan_batches`: with `bucket` (default) images of equal resized size share a batch and mixed batches pad as
    little as possible.  Like the reference's Preprocess.pad (legacy/processing.py:98-110) a batch is zero-padded
    to its largest member and that border does influence features near the image edge, so batch composition is
    part of the result; the reference's own driver is batch-1 and pads nothing (abc/extraction.py:142-199), which
    bucketing approaches (equal-size batches reproduce it exactly).  Encoded entries are decoded by one GPU
    front-end call per window, the batches run through the model's pipelined stream API, and a background thread
    appends finished windows — restored to shard order — to the Arrow file."""
    mine = shard_indices(len(image_ids), rank, world)
    meta = dict(meta or {})
    meta.setdefault("dataset", "synthetic")
    meta.setdefault("model_config", {})
    meta.setdefault("processor_args", {})
    direct = (not single_file) or world == 1
    path = os.path.join(out_dir, f"{split}.arrow" if world == 1 else f"{split}.rank{rank}.arrow")
    writer = _AsyncArrowWriter(path, [image_ids[i] for i in mine], meta) if (direct and mine) else None
    plan: List[tuple] = []          # (window number, positions inside the window) per batch, in feed order
    chunks: List[Dict[str, np.ndarray]] = []
    cfg = getattr(preprocess, "cfg", None)
    bucket = bucket and cfg is not None

    def feed():
        for wn, w0 in enumerate(range(0, len(mine), window)):
            widx = mine[w0:w0 + window]
            raws = [image_source(i) for i in widx]
            raws = [r if isinstance(r, (bytes, bytearray, memoryview)) else torch.as_tensor(r) for r in raws]
            for pos in plan_batches([_entry_hw(r) for r in raws] if bucket else [(1, 1)] * len(raws), cfg, batch_size, bucket):
                plan.append((wn, w0, len(widx), pos))
                yield [raws[j] for j in pos]

    kw = {} if max_detections is None else {"max_detections": max_detections}
    pending: Dict[int, list] = {}
    try:
        for dense in _stream(model, preprocess, feed(), max(1, window // batch_size), **kw):
            wn, w0, wlen, pos = plan.pop(0)
            ids = [image_ids[mine[w0 + j]] for j in pos]
            rows = _rows(ids, dense, dense["sizes"], dense["scales_yx"])
            slot = pending.setdefault(wn, [None, 0, wlen, []])
            slot[3].append((pos, rows))
            slot[1] += len(pos)
            if progress:
                progress(len(pos))
            if slot[1] == wlen:                      # window complete: back to shard order
                parts = pending.pop(wn)[3]
                keys = list(parts[0][1].keys())
                where = np.concatenate([np.asarray(p, np.int64) for p, _ in parts])
                inv = np.argsort(where, kind="stable")
                cols = {k: np.concatenate([r[k] for _, r in parts], 0)[inv] for k in keys}
                if writer is not None:
                    writer.put(cols)
                else:
                    chunks.append(cols)
    finally:
        if writer is not None:
            writer.close()
    if direct:
        return path
    keys = list(chunks[0].keys()) if chunks else []
    cols = {k: np.concatenate([c[k] for c in chunks], 0) for k in keys}
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
