"""Reader fast path for the extracted feature files.

Mirrors what the reference does on the consuming side of the Arrow files this package (and the
reference's own extractor) writes:

  * `Adapter._load_one_arrow` (vltk/abc/adapter.py:381-409): memory-map the IPC *stream*, read every
    record batch, json-decode the schema metadata (the `huggingface` key is skipped);
  * `Adapter.__init__` (adapter.py:70-80): `img_to_row_map` becomes the id -> row index, the other keys
    become `meta_*`;
  * `Adapter.get(img_id)` / `get_idx` / `imgids` / `n_imgs` / `has_id` (adapter.py:183-260);
  * the loaders then tensorise the `Array2D((36, 2048))` features (dataset/visnlangdataset.py:400-405).

The reference materialises every row through `datasets` as nested Python lists.  Here a column is a
zero-copy numpy view of the Arrow value buffer ([rows, 36, 2048] f32), batches are gathered with one
vectorised copy into pinned memory, and `to_device()` keeps the whole column resident in HBM (COCO-scale:
123 k images x 295 KB = 36 GB of a B200's 180 GB) so that shuffled batches are gathered by row index on the
device at HBM bandwidth (`vltk_gather_rows_f32`).  Both on-disk layouts are read: the reference's
`list<list<float>>` (datasets Array2D, one row per record batch) and this package's fixed-size lists.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np


def _flat_values(arr):
    """(values ndarray, row_width) of one nested-list Arrow chunk whose rows all have the same number of
    leaf values; raises if the rows are ragged."""
    import pyarrow as pa
    n = len(arr)
    a = arr
    width = 1
    while pa.types.is_list(a.type) or pa.types.is_large_list(a.type) or pa.types.is_fixed_size_list(a.type):
        if pa.types.is_fixed_size_list(a.type):
            width *= a.type.list_size
            a = a.flatten()           # honours the chunk's offset
        else:
            off = a.offsets.to_numpy()
            step = np.diff(off)
            if len(step) and not (step == step[0]).all():
                raise ValueError("ragged rows: this column cannot be viewed as a dense tensor")
            width *= int(step[0]) if len(step) else 0
            a = a.flatten()
    if a.null_count:
        raise ValueError("null values in a dense feature column")
    vals = a.to_numpy(zero_copy_only=True) if len(a) else np.zeros((0,), dtype=a.type.to_pandas_dtype())
    if n and len(vals) != n * width:
        raise ValueError("ragged rows: this column cannot be viewed as a dense tensor")
    return vals, width


def _shape_of(field_type, arr) -> List[int]:
    """Per-row shape of a nested list column (e.g. [36, 2048]) taken from its first row."""
    import pyarrow as pa
    shape = []
    a = arr
    while len(a) and (pa.types.is_list(a.type) or pa.types.is_large_list(a.type) or pa.types.is_fixed_size_list(a.type)):
        if pa.types.is_fixed_size_list(a.type):
            shape.append(a.type.list_size)
            a = a.flatten()
        else:
            off = a.offsets.to_numpy()
            shape.append(int(off[1] - off[0]))
            a = a.flatten()
    return shape


class FeatureTable:
    """An extracted-features Arrow file (or several per-rank shards) opened for reading."""

    def __init__(self, table, meta: Dict[str, object], paths: Sequence[str]):
        self.table = table
        self.paths = list(paths)
        self._meta_dict = meta
        m = meta.get("img_to_row_map")
        if not isinstance(m, dict):
            # files without the metadata key: fall back to the imgid column order (adapter.py:42 base schema)
            m = {str(i): r for r, i in enumerate(table.column("imgid").to_pylist())}
        self._img_to_row_map = m
        for k, v in meta.items():                     # adapter.py:75-79
            if k not in ("img_to_row_map", "vocab"):
                setattr(self, "meta_" + k, v)
        self._dense: Dict[str, List[np.ndarray]] = {}
        self._row_shape: Dict[str, List[int]] = {}
        self._chunk_start: Dict[str, np.ndarray] = {}

    # ------------------------------------------------------------------ loading
    @staticmethod
    def _read(path: str):
        import pyarrow as pa
        src = pa.memory_map(path, "r")                # stays mapped: numpy views point into it
        table = pa.ipc.open_stream(src).read_all()
        meta = {}
        for k, v in (table.schema.metadata or {}).items():
            k = k.decode()
            if k == "huggingface":
                continue
            try:
                meta[k] = json.loads(v)
            except Exception:
                meta[k] = v.decode() if isinstance(v, bytes) else v
        return table, meta

    @classmethod
    def load(cls, path: str) -> "FeatureTable":
        table, meta = cls._read(path)
        return cls(table, meta, [path])

    @classmethod
    def load_many(cls, paths: Iterable[str]) -> "FeatureTable":
        """Per-rank shards `{split}.rank{r}.arrow` (vltk_b200.extract) as one table: rows are concatenated in
        the given order and every shard's img_to_row_map is shifted by the rows before it."""
        import pyarrow as pa
        tables, merged, meta0, base = [], {}, None, 0
        paths = list(paths)
        for p in paths:
            t, m = cls._read(p)
            mp = m.get("img_to_row_map")
            if not isinstance(mp, dict):
                mp = {str(i): r for r, i in enumerate(t.column("imgid").to_pylist())}
            for k, r in mp.items():
                merged[k] = ([x + base for x in r] if isinstance(r, list) else r + base)
            base += t.num_rows
            tables.append(t.replace_schema_metadata(None))
            meta0 = meta0 or m
        meta = dict(meta0 or {})
        meta["img_to_row_map"] = merged
        return cls(pa.concat_tables(tables), meta, paths)

    # --------------------------------------------------- Adapter-compatible surface
    @property
    def img_to_row_map(self):
        return self._img_to_row_map

    @property
    def imgids(self):
        return tuple(self._img_to_row_map.keys())

    @property
    def n_imgs(self):
        return len(self._img_to_row_map)

    def __len__(self):
        return self.table.num_rows

    def has_id(self, img_id) -> bool:
        return str(img_id) in self._img_to_row_map

    def get_idx(self, img_id):
        return self._img_to_row_map[str(img_id)]

    def rows_of(self, img_ids: Sequence) -> np.ndarray:
        out = np.empty(len(img_ids), np.int64)
        for i, k in enumerate(img_ids):
            r = self._img_to_row_map[str(k)]
            out[i] = r[0] if isinstance(r, list) else r
        return out

    def get(self, img_id) -> Dict[str, object]:
        """One row as a dict, like `Adapter.get` (adapter.py:186-192) — dense columns come back as numpy
        arrays of their per-row shape instead of nested lists."""
        r = self.get_idx(img_id)
        if isinstance(r, list):
            r = r[0]
        out = {}
        for name in self.table.column_names:
            typ = self.table.schema.field(name).type
            import pyarrow as pa
            if pa.types.is_list(typ) or pa.types.is_fixed_size_list(typ) or pa.types.is_large_list(typ):
                out[name] = self.column(name, [r])[0]
            else:
                out[name] = self.table.column(name)[int(r)].as_py()
        return out

    # ------------------------------------------------------------- dense columns
    def _prepare(self, name: str):
        if name in self._dense:
            return
        col = self.table.column(name)
        chunks, starts, pos, shape = [], [], 0, None
        for ch in col.chunks:
            if len(ch) == 0:
                continue
            if shape is None:
                shape = _shape_of(ch.type, ch)
            vals, width = _flat_values(ch)
            if width != int(np.prod(shape)):
                raise ValueError(f"column {name!r}: rows of different sizes")
            chunks.append(vals.reshape([len(ch)] + shape))       # zero-copy view of the mapped file
            starts.append(pos)
            pos += len(ch)
        self._dense[name] = chunks
        self._row_shape[name] = shape or []
        self._chunk_start[name] = np.asarray(starts + [pos], np.int64)

    def row_shape(self, name: str) -> List[int]:
        self._prepare(name)
        return list(self._row_shape[name])

    def column(self, name: str, rows: Optional[Sequence[int]] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Dense [len(rows)] + row_shape array of a nested-list column (all rows when `rows` is None).  A request
        that is one contiguous slice of one record batch is returned as a zero-copy view; anything else is
        gathered with one vectorised copy per record batch touched (into `out` when given — e.g. pinned
        memory)."""
        self._prepare(name)
        chunks, starts, shape = self._dense[name], self._chunk_start[name], self._row_shape[name]
        n_all = int(starts[-1])
        rows = np.arange(n_all, dtype=np.int64) if rows is None else np.asarray(rows, np.int64).reshape(-1)
        if len(rows) and (rows.min() < 0 or rows.max() >= n_all):
            raise IndexError(f"row index out of range for a table of {n_all} rows")
        ci = np.searchsorted(starts, rows, side="right") - 1
        if out is None and len(rows) and (ci == ci[0]).all() and (np.diff(rows) == 1).all():
            lo = int(rows[0] - starts[ci[0]])
            return chunks[int(ci[0])][lo:lo + len(rows)]
        dtype = chunks[0].dtype if chunks else np.float32
        if out is None:
            out = np.empty([len(rows)] + shape, dtype)
        assert list(out.shape) == [len(rows)] + shape and out.dtype == dtype, (out.shape, out.dtype)
        for c in np.unique(ci):
            sel = np.nonzero(ci == c)[0]
            out[sel] = chunks[int(c)][rows[sel] - starts[c]]
        return out

    def features(self, rows=None, out=None) -> np.ndarray:
        """[B, 36, 2048] f32 of the reference's `features` column (adapters/frcnn.py:39)."""
        return self.column("features", rows, out)

    def pinned(self, name: str, rows: Sequence[int]):
        """torch tensor in pinned host memory holding `column(name, rows)` (ready for an async H2D copy)."""
        rows = np.asarray(rows, np.int64).reshape(-1)
        import torch
        self._prepare(name)
        dt = self._dense[name][0].dtype if self._dense[name] else np.float32
        t = torch.empty([len(rows)] + self._row_shape[name], dtype=torch.from_numpy(np.zeros(0, dt)).dtype)
        if torch.cuda.is_available():                 # page-locking needs a CUDA context; plain memory otherwise
            t = t.pin_memory()
        self.column(name, rows, out=t.numpy())
        return t

    def to_device(self, name: str = "features", device=None) -> "DeviceColumn":
        return DeviceColumn(self, name, device)


class DeviceColumn:
    """One dense f32 column resident in HBM; `gather(rows)` returns a [B] + row_shape device tensor.
    No CPU fallback: needs the CUDA library and a device."""

    def __init__(self, table: FeatureTable, name: str, device=None, chunk_rows: int = 512):
        import torch
        from . import _lib
        self._lib = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.LibraryError("DeviceColumn needs a CUDA device: there is no CPU fallback for the device gather")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = table.row_shape(name)
        self.n_rows = len(table)
        self.width = int(np.prod(self.shape))
        if self.width % 4:
            raise ValueError("row width must be a multiple of 4 floats")
        self.data = torch.empty((self.n_rows, self.width), dtype=torch.float32, device=self.device)
        stage = [torch.empty((chunk_rows, self.width), dtype=torch.float32).pin_memory() for _ in range(2)]
        evs = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.device(self.device):
            for i, lo in enumerate(range(0, self.n_rows, chunk_rows)):   # double-buffered pinned staging
                hi = min(lo + chunk_rows, self.n_rows)
                b = i & 1
                evs[b].synchronize()
                table.column(name, np.arange(lo, hi), out=stage[b].numpy()[: hi - lo].reshape([hi - lo] + self.shape))
                self.data[lo:hi].copy_(stage[b][: hi - lo], non_blocking=True)
                evs[b].record()
            torch.cuda.synchronize(self.device)

    def gather(self, rows, out=None):
        import torch
        from . import _lib
        rows_np = np.asarray(rows, np.int64).reshape(-1)
        if len(rows_np) and (rows_np.min() < 0 or rows_np.max() >= self.n_rows):
            raise IndexError(f"row index out of range for a table of {self.n_rows} rows")
        idx = torch.as_tensor(rows_np.astype(np.int32)).to(self.device, non_blocking=True)
        if out is None:
            out = torch.empty((len(rows_np), self.width), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            for lo in range(0, len(rows_np), 65535):
                hi = min(lo + 65535, len(rows_np))
                _lib.check(self._lib.vltk_gather_rows_f32(self.data.data_ptr(), self.n_rows, self.width,
                                                          idx[lo:hi].data_ptr(), hi - lo, self.width,
                                                          out[lo:hi].data_ptr(), st), "vltk_gather_rows_f32")
        return out.view([len(rows_np)] + self.shape)
