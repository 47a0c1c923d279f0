"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference from /root/reference.

Only usable in the authoring container (the GPU box has no /root/reference); used
by oracle/make_goldens.py to produce tests/golden/* and by the CPU tests that pin
oracle/frcnn_oracle.py against the real reference when it is present.

Recipe: `import vltk` itself fails on the installed
`datasets`, so a namespace stub package is registered, `wget` is stubbed, and
vltk/future/decorators.py is loaded as `vltk.decorators` (imported, never used, by
vltk/modeling/frcnn.py:32).  Nothing is copied out of the reference.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("VLTK_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "vltk", "modeling", "frcnn.py"))


def load_reference():
    """Returns (frcnn_module, compat_module) of the real reference."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    if "vltk.modeling.frcnn" in sys.modules and getattr(sys.modules["vltk"], "_b200_stub", False):
        return sys.modules["vltk.modeling.frcnn"], sys.modules["vltk.compat"]
    pkg = types.ModuleType("vltk")
    pkg.__path__ = [os.path.join(REF_ROOT, "vltk")]
    pkg._b200_stub = True
    sys.modules["vltk"] = pkg
    sys.modules.setdefault("wget", types.ModuleType("wget"))
    spec = importlib.util.spec_from_file_location(
        "vltk.decorators", os.path.join(REF_ROOT, "vltk", "future", "decorators.py"))
    dec = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dec)
    sys.modules["vltk.decorators"] = dec
    pkg.decorators = dec
    mod = types.ModuleType("vltk.modeling")
    mod.__path__ = [os.path.join(REF_ROOT, "vltk", "modeling")]
    sys.modules["vltk.modeling"] = mod
    compat = importlib.import_module("vltk.compat")
    frcnn = importlib.import_module("vltk.modeling.frcnn")
    return frcnn, compat


def build_reference_model(cfg, state_dict):
    """Reference FRCNN(cfg).eval() with `state_dict` loaded strictly."""
    import warnings
    frcnn, compat = load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = frcnn.FRCNN(compat.Config(cfg.to_reference_dict())).eval()
    model.load_state_dict(state_dict, strict=True)
    return model


def load_preprocess():
    """The reference's legacy `Preprocess` class (vltk/legacy/processing.py:76-150).
    Its module imports a `.transformers_compat` that does not exist in the tree
    (SURVEY.md §0); a stub exposing compat.img_tensorize stands in (tensor inputs
    never reach it, legacy/processing.py:120-122)."""
    _, compat = load_reference()
    if "vltk.legacy.processing" not in sys.modules:
        leg = types.ModuleType("vltk.legacy")
        leg.__path__ = [os.path.join(REF_ROOT, "vltk", "legacy")]
        sys.modules["vltk.legacy"] = leg
        tc = types.ModuleType("vltk.legacy.transformers_compat")
        tc.img_tensorize = getattr(compat, "img_tensorize", None)
        sys.modules["vltk.legacy.transformers_compat"] = tc
    return importlib.import_module("vltk.legacy.processing").Preprocess


def load_adapter():
    """The reference's `Adapter` base class (vltk/abc/adapter.py) — the loader of extracted-feature Arrow files
    (`Adapter.load` -> `_load_one_arrow` -> `datasets.Dataset(arrow_table)`, adapter.py:381-462).  Shims, all for
    imports the loading path never executes: `jsonlines` (vltk/utils/base.py:17, absent here) is stubbed and
    `datasets.ArrowWriter` (top-level name in the pinned datasets==1.9.0) is aliased to datasets.arrow_writer's."""
    load_reference()
    import datasets
    if not hasattr(datasets, "ArrowWriter"):
        from datasets.arrow_writer import ArrowWriter
        datasets.ArrowWriter = ArrowWriter
    sys.modules.setdefault("jsonlines", types.ModuleType("jsonlines"))
    for name in ("vltk.abc", "vltk.utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
            sys.modules[name] = m
    return importlib.import_module("vltk.abc.adapter").Adapter
