"""CPU checks of the drop-in boundary: libot_b200.so builds/loads without a GPU and exports every symbol that
include/ot_b200.h declares; compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from onnx_transformer_b200 import _lib
    return _lib


def test_header_symbols_are_exported_and_bound(lib):
    header = open(os.path.join(ROOT, "include", "ot_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(ot_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 20
    handle = lib.load()
    for name in declared:
        assert hasattr(handle, name), "missing export " + name
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert handle.ot_version() >= 100
    assert ctypes.sizeof(lib.OtFault) == 32


def test_compute_entry_points_fail_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = lib.load()
    assert handle.ot_device_ok() == 0
    rc = handle.ot_residual_add(None, None, None, 0, None)
    assert rc == lib.OT_ENODEV and b"no CPU fallback" in handle.ot_last_error()
    from onnx_transformer_b200 import weights as W
    from onnx_transformer_b200.engine import QuantizedTransformer
    with pytest.raises(lib.OtError):
        QuantizedTransformer(W.init_float_weights(0, 11, 13, 1), n_layers=1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "onnx-transformer_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
