"""Host-side probe: how fast can ONE process append Arrow IPC record batches of extracted features to a fresh file?
   python tools/writer_probe.py [dir ...]
Compares the sequential stream writer with pre-serialised batches written by a pool of pwrite threads (same bytes)."""
import os, sys, time, tempfile, shutil, threading, queue
import numpy as np, pyarrow as pa
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vltk_b200.extract import _to_table

n = 512
cols = {"imgid": np.asarray([f"img{i:06d}" for i in range(n)], dtype=object),
        "features": np.random.rand(n, 36, 2048).astype(np.float32), "boxes": np.random.rand(n, 36, 4).astype(np.float32)}
WINDOWS = 10


def sequential(path):
    t = _to_table(cols)
    with pa.OSFile(path, "wb") as sink:
        with pa.ipc.new_stream(sink, t.schema) as w:
            for _ in range(WINDOWS):
                for b in _to_table(cols).to_batches(max_chunksize=128):
                    w.write_batch(b)


def parallel(path, threads):
    t = _to_table(cols)
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    q = queue.Queue(maxsize=2 * threads)

    def work():
        while True:
            it = q.get()
            if it is None:
                return
            off, batch = it
            buf = batch.serialize()
            os.pwrite(fd, buf, off)
    th = [threading.Thread(target=work) for _ in range(threads)]
    [x.start() for x in th]
    head = t.schema.serialize()
    os.pwrite(fd, head, 0)
    off = head.size
    for _ in range(WINDOWS):
        for b in _to_table(cols).to_batches(max_chunksize=128):
            sz = pa.ipc.get_record_batch_size(b)
            q.put((off, b))
            off += sz
    for _ in th:
        q.put(None)
    [x.join() for x in th]
    os.pwrite(fd, b"\xff\xff\xff\xff\x00\x00\x00\x00", off)
    os.close(fd)


for base in (sys.argv[1:] or ["/tmp", "/dev/shm"]):
    d = tempfile.mkdtemp(dir=base)
    try:
        for name, fn in [("sequential", sequential), ("pwrite x1", lambda p: parallel(p, 1)), ("pwrite x2", lambda p: parallel(p, 2)),
                         ("pwrite x4", lambda p: parallel(p, 4)), ("pwrite x8", lambda p: parallel(p, 8)), ("sequential again", sequential)]:
            path = os.path.join(d, name.replace(" ", "_") + ".arrow")
            t0 = time.perf_counter()
            fn(path)
            dt = time.perf_counter() - t0
            sz = os.path.getsize(path)
            with pa.memory_map(path, "r") as src:
                rows = pa.ipc.open_stream(src).read_all().num_rows
            print(f"{base:9s} {name:18s} {sz / 1e9:.2f} GB in {dt:.2f} s = {sz / 1e9 / dt:.2f} GB/s, {rows} rows ({rows / dt:.0f} img/s)", flush=True)
            os.remove(path)
    finally:
        shutil.rmtree(d, ignore_errors=True)
print("cores", os.cpu_count())
