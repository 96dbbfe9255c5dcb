"""The C-ABI library loads without a GPU and exports every symbol include/y3b200.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from yolo_v3_tf2_b200 import _lib
    header = open(os.path.join(ROOT, "include", "y3b200.h")).read()
    declared = set(re.findall(r"\b(y3_[a-z0-9_]+)\s*\(", header))
    declared -= {"y3_net_plan_"}
    assert len(declared) >= 20
    handle = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in y3b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().y3_version() == 100


def test_struct_layouts_match_header(built):
    from yolo_v3_tf2_b200 import _lib
    assert C.sizeof(_lib.LayerDesc) == 9 * 4
    assert C.sizeof(_lib.LayerPlan) == 14 * 4 + 8


def test_no_cpu_fallback(built):
    """Compute entry points refuse to run on a planning-only context; on a machine without CUDA the device context
    itself cannot be created."""
    import torch
    from yolo_v3_tf2_b200 import _lib
    ctx = _lib.Context(-1)
    rc = _lib.lib().y3_class_reduce(ctx.handle, 1, 1, 1, 1, 1, 1, 1, None)
    assert rc == _lib.Y3_ERR_STATE
    assert b"no CPU fallback" in _lib.lib().y3_last_error()
    if not torch.cuda.is_available():
        with pytest.raises(_lib.Y3Error):
            _lib.Context(0)
        with pytest.raises(_lib.Y3Error):
            _lib.context()


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "yolo_v3_tf2_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
