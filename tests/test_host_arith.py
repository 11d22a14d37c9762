"""CPU check of the device arithmetic headers (AAN DCT butterflies, scale tables, folded score
weights): the same header code the kernels compile is built for the host with g++ and compared
with the NumPy specs."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import spec_dct_dampen, spec_scoring

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostlib") / "libhost_dct.so"
    cmd = ["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", os.path.join(HERE, "host_dct_check.cpp"), "-o", str(out)]
    subprocess.run(cmd, check=True)
    return ctypes.CDLL(str(out))


def test_score_weights_inc_is_current():
    """score_weights.inc must be what tools/gen_score_weights.py generates."""
    root = os.path.dirname(HERE)
    path = os.path.join(root, "elvis_b200", "csrc", "score_weights.inc")
    before = open(path).read()
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_score_weights.py")], check=True, cwd=root)
    assert open(path).read() == before


def test_host_score_tile_matches_spec(hostlib):
    rng = np.random.default_rng(0)
    T = 12
    for trial in range(20):
        y = rng.integers(0, 256, (T, 8, 8), dtype=np.uint8)
        if trial % 3 == 0:
            y[5] = y[4]
            y[7] = y[6]
            y[7, 3, 3] ^= 1
        sc = np.zeros(T, np.float32)
        tc = np.zeros(T, np.float32)
        hostlib.host_score_tile(y.ctypes.data_as(ctypes.c_void_p), T, sc.ctypes.data_as(ctypes.c_void_p),
                                tc.ctypes.data_as(ctypes.c_void_p))
        rsc, rtc = spec_scoring.sc_tc(y, 8)          # one 8x8 block per frame; spec divides by 64
        np.testing.assert_allclose(sc / 64, rsc[:, 0, 0], rtol=1e-5)
        np.testing.assert_allclose(tc / 64, rtc[:, 0, 0], rtol=1e-5)


def test_host_dampen_tile_matches_spec(hostlib):
    rng = np.random.default_rng(1)
    for s in (0.0, 0.3, 1.0):
        blk = rng.integers(0, 256, (8, 8), dtype=np.uint8)
        g = (np.exp2(-4.0 * s * np.arange(15) / 14.0) / 64.0).astype(np.float32)
        out = np.zeros((8, 8), np.float32)
        hostlib.host_dampen_tile(blk.ctypes.data_as(ctypes.c_void_p), g.ctypes.data_as(ctypes.c_void_p),
                                 out.ctypes.data_as(ctypes.c_void_p))
        ref = spec_dct_dampen.dampen_plane(blk, np.array([[s]]), 8, return_float=True)
        np.testing.assert_allclose(out, ref, atol=2e-3)
