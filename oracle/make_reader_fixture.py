"""TEST INFRASTRUCTURE — derives tests/golden/ref_frcnn_train_head3.arrow from the reference's own extracted-
features fixture (/root/reference/tests/visualgenome/frcnn/train.arrow, written by the reference's extractor):
the first 3 rows, re-written with the SAME schema (incl. the datasets Array2D field metadata), the same schema
metadata keys (img_to_row_map cut to the kept rows) and the same one-row-per-record-batch cadence, plus per-row
checksums computed through pyarrow's slow per-row path (`as_py()`, what `datasets` formatting does).  The
reference tree does not exist on the GPU box, so the reader tests use this 0.9 MB copy.

    python oracle/make_reader_fixture.py
"""
import json
import os

import numpy as np
import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/tests/visualgenome/frcnn/train.arrow"
DST = os.path.join(ROOT, "tests", "golden", "ref_frcnn_train_head3.arrow")
KEEP = 3

if __name__ == "__main__":
    table = pa.ipc.open_stream(pa.memory_map(SRC)).read_all()
    head = table.slice(0, KEEP)
    md = dict(table.schema.metadata)
    full_map = json.loads(md[b"img_to_row_map"])
    md[b"img_to_row_map"] = json.dumps({k: v for k, v in full_map.items() if v < KEEP}).encode()
    head = head.replace_schema_metadata(md)
    with pa.OSFile(DST, "wb") as sink:
        with pa.ipc.new_stream(sink, head.schema) as w:
            for b in head.to_batches(max_chunksize=1):
                w.write_batch(b)
    sums = {}
    for r in range(KEEP):
        row = {}
        for name in ("features", "box", "attr_ids", "object_ids"):
            a = np.asarray(table.column(name)[r].as_py(), np.float64)
            row[name] = {"shape": list(a.shape), "sum": float(a.sum()), "abs_sum": float(np.abs(a).sum()),
                         "first": float(a.reshape(-1)[0]), "last": float(a.reshape(-1)[-1])}
        row["imgid"] = table.column("imgid")[r].as_py()
        sums[str(r)] = row
    with open(DST.replace(".arrow", ".json"), "w") as f:
        json.dump({"source": SRC, "rows": KEEP, "schema": str(table.schema), "checks": sums}, f, indent=1)
    print(f"wrote {DST} ({os.path.getsize(DST) / 1e6:.2f} MB)")
