"""-m gpu: randomized differential tests -- many small random geometries per kernel family against the
oracle (seeded, so a failure is reproducible from the printed case)."""
import os

import numpy as np
import pytest

from _util import synth_luma
from oracle import ref_port as P
from oracle import spec_scoring

pytestmark = pytest.mark.gpu
RTOL = 1e-4
SCALE = int(os.environ.get("ELVIS_FUZZ_SCALE", "1"))     # multiply the number of random cases (soak runs)


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available()
    return torch.device("cuda")


def _to(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("impl", ["umma", "simt"])
def test_fuzz_scoring(dev, monkeypatch, impl):
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_SCORE_IMPL", impl)
    rng = np.random.default_rng(2024)
    for case in range(24 * SCALE):
        bs = int(rng.choice([8, 16, 32]))
        by, bx = int(rng.integers(1, 6)), int(rng.integers(1, 40))
        # widths that are / are not multiples of 16 bytes, with spare columns and rows beyond the block grid
        H, W = by * bs + int(rng.integers(0, bs)), bx * bs + int(rng.choice([0, 0, 3, 8, 13]))
        T = int(rng.integers(1, 15))
        chunk = str(int(rng.choice([3, 5, 64])))
        monkeypatch.setenv("ELVIS_SCORE_CHUNK", chunk)
        y = synth_luma(T + 1, H, W, seed=case)
        if case % 3 == 0:
            y[-1] = rng.integers(0, 256, (H, W), dtype=np.uint8)
        halo = case % 2 == 1
        yd = _to(y, dev)
        sc, tc, mm = ops.score_sc_tc(yd[1:], bs, prev_halo=yd[0] if halo else None)
        rsc, rtc = spec_scoring.sc_tc(y[1:], bs, prev=y[0] if halo else None)
        info = (impl, case, bs, H, W, T, chunk, halo)
        np.testing.assert_allclose(sc.cpu().numpy(), rsc, rtol=RTOL, atol=1e-6 * max(1.0, rsc.max()), err_msg=str(info))
        np.testing.assert_allclose(tc.cpu().numpy(), rtc, rtol=RTOL, atol=1e-6 * max(1.0, rtc.max()), err_msg=str(info))
        s, t = sc.cpu().numpy(), tc.cpu().numpy()
        assert mm.cpu().numpy().tolist() == [s.min(), s.max(), t.min(), t.max()], info


def test_fuzz_shrink_stretch_and_rowcol(dev):
    from elvis_b200 import elvis as E, utils as U
    rng = np.random.default_rng(2025)
    for case in range(24 * SCALE):
        bs = int(rng.choice([4, 8, 16]))
        by, bx = int(rng.integers(1, 7)), int(rng.integers(2, 12))
        img = rng.integers(0, 256, (by * bs, bx * bs, 3), dtype=np.uint8)
        scores = np.round(rng.random((by, bx)) * 8) / 8
        amount = float(rng.choice([0.0, 0.2, 0.5, 0.8, 2, 0.999]))
        g, r = E.apply_selective_removal(img, scores, bs, amount), P.apply_selective_removal(img, scores, bs, amount)
        assert np.array_equal(g[0], r[0]) and np.array_equal(g[1], r[1]) and g[2] == r[2], (case, bs, by, bx, amount)
        assert np.array_equal(E.stretch_frame(g[0], g[1], bs), P.stretch_frame(r[0], r[1], bs))
        frac = float(rng.choice([0.0, 0.15, 0.4, 0.7, 0.95]))
        crop = img if case % 2 else np.pad(img, ((0, 3), (0, bs - 1), (0, 0)))      # spare pixels, less than a block
        g, r = U.shrink_frame_row_only(crop, scores, bs, frac), P.shrink_frame_row_only(crop, scores, bs, frac)
        assert np.array_equal(g[0], r[0]) and np.array_equal(g[1], r[1]), (case, "row_only", frac)
        g, r = U.shrink_frame_position_map(crop, scores, bs, frac), P.shrink_frame_position_map(crop, scores, bs, frac)
        assert all(np.array_equal(a, b) for a, b in zip(g, r)), (case, "position_map", frac)
        assert np.array_equal(U.stretch_frame_position_map(*g, bs), P.stretch_frame_position_map(*r, bs))
        g, r = U.shrink_frame_removal_indices(crop, scores, bs, frac), P.shrink_frame_removal_indices(crop, scores, bs, frac)
        assert np.array_equal(g[0], r[0]) and all(np.array_equal(a, b) for a, b in zip(g[2], r[2]))
        assert np.array_equal(U.stretch_frame_removal_indices(g[0], g[2], by, bx, bs),
                              P.stretch_frame_removal_indices(r[0], r[2], by, bx, bs)), (case, "removal_indices", frac)


def test_fuzz_degradations(dev):
    from elvis_b200 import elvis as E, utils as U
    rng = np.random.default_rng(2026)
    for case in range(16 * SCALE):
        bs = int(rng.choice([8, 16, 32]))
        by, bx = int(rng.integers(1, 5)), int(rng.integers(1, 7))
        img = rng.integers(0, 256, (by * bs, bx * bs, 3), dtype=np.uint8)
        scores = rng.random((by, bx))
        for fn in ("filter_frame_downsample", "filter_frame_gaussian"):
            g, r = getattr(E, fn)(img, scores, bs), getattr(P, fn)(img, scores, bs)
            assert np.array_equal(g[1], r[1]) and np.array_equal(g[0], r[0]), (case, fn, bs)
        crop = np.pad(img, ((0, int(rng.integers(0, bs))), (0, int(rng.integers(0, bs))), (0, 0)))
        for fn in ("degrade_adaptive_downsample", "degrade_adaptive_blur"):
            g, r = getattr(U, fn)(crop, scores, bs), getattr(P, fn)(crop, scores, bs)
            assert np.array_equal(g[1], r[1]) and np.array_equal(g[0], r[0]), (case, fn, bs)
        lv = rng.integers(0, 7, (by, bx))
        assert np.array_equal(E.restore_blur_opencv_unsharp_mask(img, lv, bs), P.restore_blur_opencv_unsharp_mask(img, lv, bs)), (case, "unsharp")
        dl = rng.integers(0, int(np.log2(bs)) + 1, (by, bx))
        assert np.array_equal(E.restore_downsample_opencv_lanczos(img, dl, bs), P.restore_downsample_opencv_lanczos(img, dl, bs)), (case, "lanczos")


def test_fuzz_planar_pipeline_and_side_channels(dev):
    """Planar YUV 4:2:0 clips of random geometry through ElvisV1 (fused Y+U+V move kernel for
    16-multiples, per-plane kernels otherwise), the pipelined variants, and the bit packers."""
    import torch
    from _util import synth_yuv420
    from elvis_b200 import ops
    from elvis_b200.pipeline import ElvisV1, ElvisV1Pipelined, Yuv420
    rng = np.random.default_rng(2027)
    for case in range(10 * SCALE):
        bs = int(rng.choice([8, 16, 32]))
        by, bx = int(rng.integers(1, 6)), int(rng.integers(2, 14))
        T, H, W = int(rng.integers(1, 7)), by * bs, bx * bs
        amount = float(rng.choice([0.0, 0.25, 0.5, 0.9]))
        y, u, v = synth_yuv420(T, H, W, seed=100 + case)
        clip = Yuv420(_to(y, dev), _to(u, dev), _to(v, dev))
        scores, mask, shrunk, full = ElvisV1(bs, amount, 0.4, 0.6).run(clip)
        s_np, m_np = scores.cpu().numpy(), mask.cpu().numpy()
        k = P.blocks_to_remove_elvis(amount, bx)
        assert np.array_equal(m_np, P.select_rows(s_np, k, P.REMOVE_HIGH)), (case, bs, by, bx, T, amount)
        for name, plane, pb in (("y", y, bs), ("u", u, bs // 2), ("v", v, bs // 2)):
            for t in range(T):
                rs = P.shrink_plane(plane[t], m_np[t], pb)
                assert np.array_equal(getattr(shrunk, name)[t].cpu().numpy(), rs), (case, name, t)
                assert np.array_equal(getattr(full, name)[t].cpu().numpy(), P.stretch_plane(rs, m_np[t], pb)), (case, name, t)
        for depth, split in ((2, False), (3, True)):
            slot = ElvisV1Pipelined(T, H, W, bs, amount, 0.4, 0.6, dev, depth=depth, split_stretch=split).submit(clip)
            slot["done"].synchronize()
            assert torch.equal(slot["mask"], mask) and torch.equal(slot["scores"], scores)
            assert all(torch.equal(a, b) for a, b in zip(slot["full"].planes, full.planes))
        packed, shape = P.pack_masks(m_np)
        assert np.array_equal(ops.pack_mask_bits(mask).cpu().numpy(), packed)
        assert np.array_equal(ops.unpack_mask_bits(_to(packed, dev), shape).cpu().numpy(), m_np)
        lv = rng.integers(0, 4, (T, by, bx)).astype(np.int32)
        p2 = ops.pack_levels_2bit(_to(lv, dev))
        assert np.array_equal(p2.cpu().numpy(), P.pack_levels_2bit(lv))
        assert np.array_equal(ops.unpack_levels_2bit(p2, bx).cpu().numpy(), lv)


def test_fuzz_roi_and_i420(dev, tmp_path):
    from elvis_b200 import utils as U
    rng = np.random.default_rng(2028)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    for case in range(8 * SCALE):
        h, w = 16 * int(rng.integers(3, 40)), 16 * int(rng.integers(5, 70))
        maps = [rng.random((h // 16, w // 16)) for _ in range(2)]
        crf, rq = int(rng.integers(0, 64)), int(rng.integers(0, 20))
        U.create_svtav1_roi_file(maps, a, crf, rq, w, h)
        P.create_svtav1_roi_file(maps, b, crf, rq, w, h)
        assert open(a).read() == open(b).read(), (case, h, w, crf, rq)
        qp, rr = int(rng.integers(0, 52)), int(rng.integers(0, 26))
        U.create_kvazaar_roi_file(maps, a, qp, rr)
        P.create_kvazaar_roi_file(maps, b, qp, rr)
        assert open(a, "rb").read() == open(b, "rb").read(), (case, qp, rr)
        fh, fw = 2 * int(rng.integers(1, 40)), 2 * int(rng.integers(1, 60))
        frames = [rng.integers(0, 256, (fh, fw, 3), dtype=np.uint8) for _ in range(2)]
        U.write_y4m(frames, a, 25.0)
        P.write_y4m(frames, b, 25.0)
        assert open(a, "rb").read() == open(b, "rb").read(), (case, fh, fw)
