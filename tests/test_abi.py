"""The C-ABI library loads on a CPU-only box and exports every symbol include/elvis_b200.h
declares (no compute calls); argument validation that needs no device memory is exercised."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from elvis_b200 import build
    build.build()
    from elvis_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "elvis_b200.h")).read()
    declared = set(re.findall(r"ELVIS_API\s+(?:const\s+char\*|int64_t|int)\s+(elvis_\w+)\s*\(", header))
    assert len(declared) >= 19
    assert declared == set(lib.EXPORTS)
    for name in declared:
        assert hasattr(lib.lib, name), name


def test_version_and_error_strings(lib):
    assert lib.lib.elvis_abi_version() == lib.ABI_VERSION
    assert lib.lib.elvis_error_string(0) == b"ok"
    assert b"divisible" in lib.lib.elvis_error_string(lib.ERR_SHAPE)


def test_argument_validation_without_a_device(lib):
    null = ctypes.c_void_p(0)
    assert lib.lib.elvis_minmax(null, 0, 10, null, null) == lib.ERR_INVALID_ARG
    assert lib.lib.elvis_select_rows(null, 1, 1, 1, null, 0, 0, null, null) == lib.ERR_INVALID_ARG
    pl = lib.Plane(1, 64, 8, 8, 8, 1, 0)     # a fake non-null pointer is never dereferenced on these paths
    one = ctypes.c_void_p(1)
    assert lib.lib.elvis_score_sc_tc(ctypes.byref(pl), 1, null, 12, 8, one, one, null, 0, 1, null) == lib.ERR_UNSUPPORTED
    assert lib.lib.elvis_score_sc_tc(ctypes.byref(pl), 1, null, 16, 8, one, one, null, 0, 1, null) == lib.ERR_SHAPE
    with pytest.raises(ValueError):
        lib.call("elvis_score_sc_tc", ctypes.byref(pl), 1, null, 16, 8, one, one, null, 0, 1, null)
    with pytest.raises(lib.ElvisError):
        lib.call("elvis_minmax", null, 0, 10, null, null)


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from elvis_b200 import ops
    with pytest.raises(TypeError):
        ops.score_sc_tc(torch.zeros((1, 16, 16), dtype=torch.uint8), 16)
    with pytest.raises(TypeError):
        ops.select_rows(torch.zeros((1, 2, 3), dtype=torch.float64), 1)
