"""CPU-side checks of the drop-in boundary: the C-ABI library loads here (no GPU), exports
every symbol include/vltk_frcnn.h declares, and refuses to run without a device — there is
no CPU fallback to fall into."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vltk_frcnn.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vltk_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def library():
    from vltk_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_and_binding_agree(library):
    syms = header_symbols()
    assert len(syms) >= 15
    assert sorted(library.SYMBOLS) == syms


def test_library_exports_every_declared_symbol(library):
    lib = library.lib()
    for s in header_symbols():
        assert hasattr(lib, s), s
    nm = subprocess.run(["nm", "-D", "--defined-only", library.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vltk_\w+)", nm))
    assert set(header_symbols()) <= exported
    assert lib.vltk_frcnn_version().startswith(b"vltk_b200")


def test_library_is_torch_free_and_sm100a(library):
    ldd = subprocess.run(["ldd", library.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "cudnn" not in ldd and "cublas" not in ldd
    sass = subprocess.run(["cuobjdump", "-lelf", library.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_create_fails_loudly_without_gpu(library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vltk_b200.config import FRCNNConfig
    h = C.c_void_p()
    rc = library.lib().vltk_frcnn_create(C.byref(library.make_config(FRCNNConfig(), "bf16")), 0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in library.lib().vltk_frcnn_last_error()
    from vltk_b200.frcnn import FRCNN
    with pytest.raises(library.LibraryError):
        FRCNN(FRCNNConfig())


def test_struct_layouts_match_header(library):
    # field order/width of the ctypes mirrors vs the C structs (compiled probe)
    probe = r'''
    #include "%s"
    #include <stdio.h>
    int main(){ printf("%%zu %%zu %%zu\n", sizeof(vltk_frcnn_config), sizeof(vltk_frcnn_knobs), sizeof(vltk_frcnn_out)); return 0; }
    ''' % HEADER
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "p.c")
        open(src, "w").write(probe)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", src, "-o", exe])
        sizes = list(map(int, subprocess.check_output([exe]).split()))
    assert sizes == [C.sizeof(library.Config), C.sizeof(library.Knobs), C.sizeof(library.Out)]
