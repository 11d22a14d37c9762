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


def test_rowcol_dims_match_the_oracle_plan():
    """Host-side pass structure of the row+column shrink (elvis_b200.ops.rowcol_dims) against the
    oracle's simulation (sizes and counts depend on the grid and the target only)."""
    import numpy as np
    from elvis_b200 import ops
    from oracle import ref_port as P
    rng = np.random.default_rng(0)
    for by, bx, amount in [(4, 6, 0.3), (5, 8, 0.55), (9, 7, 0.85), (1, 8, 0.5), (8, 1, 0.5), (6, 6, 1.0), (6, 6, 0.0), (17, 30, 0.4)]:
        mask, pmap, passes = P.rowcol_plan(rng.random((by, bx)), amount)
        fby, fbx, counts = ops.rowcol_dims(by, bx, int(by * bx * amount))
        assert (fby, fbx) == pmap.shape[:2] and counts == [len(a) for a in passes] and sum(counts) == mask.sum()


def test_roi_host_logic_matches_the_oracle():
    """x265 CTU choice and the cv2 float-area tables built on the host."""
    import numpy as np
    from elvis_b200 import _tables as T
    from elvis_b200 import elvis as E
    from oracle import ref_port as P
    from oracle import spec_cv
    for bs in (4, 8, 16, 32, 64, 128):
        for (w, h) in ((320, 240), (1920, 1080), (3840, 2160), (2160, 3840), (7680, 4320)):
            assert E.x265_ctu_size(bs, w, h) == P.x265_ctu_size(bs, w, h)
    for ssize, dsize in ((240, 60), (135, 34), (67, 17), (13, 5), (9, 9)):
        start, src, alpha = T.area_f32_tables(ssize, dsize)
        tab = spec_cv.area_table(ssize, dsize)
        assert [s for s, _, _ in tab] == src.tolist() and np.array_equal(np.array([a for _, _, a in tab], np.float32), alpha)
        assert np.bincount([d for _, d, _ in tab], minlength=dsize).cumsum().tolist() == start[1:].tolist()
    assert T.area_f32_plan(136, 240, 34, 60) == (4, 4, 0) and T.area_f32_plan(16, 24, 8, 12) == (2, 2, 12)
    assert T.area_f32_plan(135, 240, 34, 60) == (0, 0, 0)


def test_nearest_index_matches_cv2():
    import numpy as np
    cv2 = __import__("pytest").importorskip("cv2")
    from elvis_b200 import _tables as T
    rng = np.random.default_rng(3)
    for _ in range(200):
        ssize, dsize = (int(v) for v in rng.integers(1, 300, 2))
        row = np.arange(ssize, dtype=np.float32)[None]
        ref = cv2.resize(row, (dsize, 1), interpolation=cv2.INTER_NEAREST)[0].astype(np.int64)
        assert np.array_equal(T.nearest_index(ssize, dsize), ref), (ssize, dsize)


def test_fast_levels_flag():
    """The `fast_tables_ok` argument of elvis_degrade_downsample: 1 = every level is a power-of-two reduction, 0 = none (or
    not describable), otherwise an even value whose bit l + 1 marks level l as one for the table-driven kernel."""
    from elvis_b200 import _tables as T
    assert T.fast_levels_flag(16, (16, 8, 4, 2, 1)) == 1
    assert T.fast_levels_flag(8, (8, 4, 2, 2)) == 1
    assert T.fast_levels_flag(16, (16, 8, 5, 4)) == 1 << 3            # utils.py:1142-1148: only 16 -> 5 is fractional
    assert T.fast_levels_flag(16, (16, 3, 8, 7, 2, 1)) == (1 << 2) | (1 << 4)
    assert T.fast_levels_flag(16, (5, 3)) == 0                        # nothing for the closed form
    assert T.fast_levels_flag(32, (32, 16)) == 0                      # only 8- and 16-pixel blocks have the fast kernels
    assert T.fast_levels_flag(16, (16, 8, 5, 4)) % 2 == 0 and T.all_fast(16, (16, 8, 4)) and not T.all_fast(16, (16, 5))
