"""-m gpu: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.
Bit-exact for masks, frames, level maps, blurred/downsampled pixels and -- given identical
SC/TC -- float64 scores; <= 1e-4 relative for SC/TC and dampened pixels."""
import numpy as np
import pytest

from _util import random_scores, synth_luma, synth_yuv420
from oracle import ref_port as P
from oracle import spec_dct_dampen, spec_scoring

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # north_star tolerance for float scores / dampened pixels


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import elvis_b200.ops  # noqa: F401  (raises if libelvis_b200.so is missing)
    return torch.device("cuda")


def to_dev(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ---------------------------------------------------------------------------------- a1
IMPLS = ["umma", "simt"]      # tcgen05 kernel (default for large aligned clips) / CUDA-core kernel


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("bs,H,W,T", [(16, 64, 96, 7), (8, 40, 72, 5), (32, 64, 128, 4), (16, 48, 272, 30), (16, 1080, 1920, 3),
                                      (8, 72, 1000, 14), (32, 96, 528, 13)])
def test_sc_tc_vs_spec(dev, monkeypatch, impl, bs, H, W, T):
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_SCORE_IMPL", impl)
    y = synth_luma(T, H, W, seed=bs + T)
    sc, tc, mm = ops.score_sc_tc(to_dev(y, dev), bs)
    rsc, rtc = spec_scoring.sc_tc(y, bs)
    sc, tc, mm = sc.cpu().numpy(), tc.cpu().numpy(), mm.cpu().numpy()
    assert sc.shape == rsc.shape
    np.testing.assert_allclose(sc, rsc, rtol=RTOL, atol=0)
    np.testing.assert_allclose(tc, rtc, rtol=RTOL, atol=0)
    assert np.all(tc[0] == 0)
    assert mm.tolist() == [sc.min(), sc.max(), tc.min(), tc.max()]


@pytest.mark.parametrize("T", [1, 2, 6, 7, 12, 13, 19])
def test_sc_tc_tcgen05_frame_loop_boundaries(dev, monkeypatch, T):
    """The tcgen05 kernel unrolls its frame loop by the ring depth (6) and buffers A x2 / D x3: run
    lengths around the multiples, with and without a halo (which adds a priming frame), one chunk
    and several."""
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_SCORE_IMPL", "umma")
    y = synth_luma(T + 1, 48, 144, seed=T)
    yd = to_dev(y, dev)
    for chunk in ("64", "5"):
        monkeypatch.setenv("ELVIS_SCORE_CHUNK", chunk)
        sc, tc, _ = ops.score_sc_tc(yd[1:], 16)
        rsc, rtc = spec_scoring.sc_tc(y[1:], 16)
        np.testing.assert_allclose(sc.cpu().numpy(), rsc, rtol=RTOL, atol=0)
        np.testing.assert_allclose(tc.cpu().numpy(), rtc, rtol=RTOL, atol=0)
        sch, tch, _ = ops.score_sc_tc(yd[1:], 16, prev_halo=yd[0])
        rsc_h, rtc_h = spec_scoring.sc_tc(y[1:], 16, prev=y[0])
        np.testing.assert_allclose(sch.cpu().numpy(), rsc_h, rtol=RTOL, atol=0)
        np.testing.assert_allclose(tch.cpu().numpy(), rtc_h, rtol=RTOL, atol=0)
        assert np.array_equal(sch.cpu().numpy(), sc.cpu().numpy())      # SC does not depend on the halo


@pytest.mark.parametrize("impl", IMPLS)
def test_sc_tc_random_noise_and_static(dev, monkeypatch, impl):
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_SCORE_IMPL", impl)
    rng = np.random.default_rng(5)
    y = rng.integers(0, 256, (6, 64, 64), dtype=np.uint8)
    y[3] = y[2]                      # a repeated frame: TC must be exactly 0
    y[5] = y[4]
    y[5, 10, 10] ^= 1                # a single-LSB change: the smallest non-zero TC
    sc, tc, _ = ops.score_sc_tc(to_dev(y, dev), 16)
    rsc, rtc = spec_scoring.sc_tc(y, 16)
    np.testing.assert_allclose(sc.cpu().numpy(), rsc, rtol=RTOL, atol=0)
    np.testing.assert_allclose(tc.cpu().numpy(), rtc, rtol=RTOL, atol=0)
    assert np.all(tc.cpu().numpy()[3] == 0)


@pytest.mark.parametrize("impl", IMPLS)
def test_sc_tc_chunking_and_halo(dev, monkeypatch, impl):
    """Chunk boundaries and the sharding halo must not change any value."""
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_SCORE_IMPL", impl)
    y = synth_luma(21, 32, 64, seed=9)
    yd = to_dev(y, dev)
    monkeypatch.setenv("ELVIS_SCORE_CHUNK", "64")
    sc1, tc1, _ = ops.score_sc_tc(yd, 16)
    monkeypatch.setenv("ELVIS_SCORE_CHUNK", "4")
    sc4, tc4, _ = ops.score_sc_tc(yd, 16)
    rsc, rtc = spec_scoring.sc_tc(y, 16)
    for a in (sc1, sc4):
        np.testing.assert_allclose(a.cpu().numpy(), rsc, rtol=RTOL, atol=0)
    for a in (tc1, tc4):
        np.testing.assert_allclose(a.cpu().numpy(), rtc, rtol=RTOL, atol=0)
    if impl == "umma":      # coefficients are computed afresh per frame: chunking cannot change a bit
        assert np.array_equal(sc1.cpu().numpy(), sc4.cpu().numpy()) and np.array_equal(tc1.cpu().numpy(), tc4.cpu().numpy())
    # halo: frames 8.. scored alone with frame 7 as prev_halo == the tail of the full run
    sch, tch, mm = ops.score_sc_tc(yd[8:], 16, prev_halo=yd[7], minmax_range=(2, 5))
    rsc_h, rtc_h = spec_scoring.sc_tc(y[8:], 16, prev=y[7])
    np.testing.assert_allclose(sch.cpu().numpy(), rsc_h, rtol=RTOL, atol=0)
    np.testing.assert_allclose(tch.cpu().numpy(), rtc_h, rtol=RTOL, atol=0)
    s, t = sch.cpu().numpy()[2:5], tch.cpu().numpy()[2:5]
    assert mm.cpu().numpy().tolist() == [s.min(), s.max(), t.min(), t.max()]


def test_sc_tc_strided_rows(dev):
    """A cropped view (row stride > width, width not a multiple of the warp tile)."""
    from elvis_b200 import ops
    y = synth_luma(4, 48, 112, seed=3)
    yd = to_dev(y, dev)[:, :, :88]          # 88 = 5.5 blocks of 16 -> Bx = 5
    sc, tc, _ = ops.score_sc_tc(yd, 16)
    rsc, rtc = spec_scoring.sc_tc(y[:, :, :88], 16)
    np.testing.assert_allclose(sc.cpu().numpy(), rsc, rtol=RTOL, atol=0)
    np.testing.assert_allclose(tc.cpu().numpy(), rtc, rtol=RTOL, atol=0)
    yo = to_dev(y, dev)[:, :, 3:3 + 80]     # misaligned base pointer -> byte-load path
    sc, tc, _ = ops.score_sc_tc(yo, 16)
    rsc, rtc = spec_scoring.sc_tc(y[:, :, 3:83], 16)
    np.testing.assert_allclose(sc.cpu().numpy(), rsc, rtol=RTOL, atol=0)
    np.testing.assert_allclose(tc.cpu().numpy(), rtc, rtol=RTOL, atol=0)


# ---------------------------------------------------------------------------------- a2 / a3
@pytest.mark.parametrize("beta", [1.0, 0.5, 0.25])
@pytest.mark.parametrize("with_bg", [False, True])
def test_removability_bit_exact(dev, beta, with_bg):
    from elvis_b200 import elvis as E
    rng = np.random.default_rng(11)
    T, By, Bx = 9, 7, 13
    sc, tc = rng.random((T, By, Bx)) * 60, rng.random((T, By, Bx)) * 30
    tc[0] = 0
    bg = rng.random((T, By, Bx)) > 0.6 if with_bg else None
    got = E.removability_from_features(sc, tc, 0.3, beta, bg)
    ref = P.combine_removability(sc, tc, 0.3, beta, bg)
    assert got.dtype == np.float64 and np.array_equal(got, ref)


def test_removability_flat_inputs(dev):
    from elvis_b200 import elvis as E
    sc = np.full((3, 4, 5), 2.5)
    tc = np.zeros((3, 4, 5))
    assert np.array_equal(E.removability_from_features(sc, tc, 0.5, 0.5), P.combine_removability(sc, tc, 0.5, 0.5))
    one = np.random.default_rng(0).random((1, 4, 5))
    assert np.array_equal(E.removability_from_features(one, one * 0, 0.5, 0.5), P.combine_removability(one, one * 0, 0.5, 0.5))


def test_importance_bit_exact(dev):
    from elvis_b200 import utils as U
    rng = np.random.default_rng(12)

    class Cx:
        pass
    T, By, Bx = 6, 9, 11
    cx = Cx()
    cx.SC, cx.TC = rng.random((T, By, Bx)) * 50, rng.random((T, By, Bx)) * 20
    fg = rng.random((T, By, Bx))
    got = np.stack(U.calculate_importance_scores(None, 16, 0.3, 0.6, cx, fg))
    assert np.array_equal(got, P.importance_scores(cx.SC, cx.TC, 0.3, 0.6, fg))
    cx.SC, cx.TC = cx.SC[:1], cx.TC[:1]
    got = np.stack(U.calculate_importance_scores(None, 16, 0.3, 0.6, cx, fg[:1]))
    assert np.array_equal(got, P.importance_scores(cx.SC, cx.TC, 0.3, 0.6, fg[:1]))


# ---------------------------------------------------------------------------------- a4-a7
@pytest.mark.parametrize("ties", ["none", "quantised", "all"])
@pytest.mark.parametrize("bx", [5, 64, 120, 240, 300, 700])
def test_select_rows_bit_exact(dev, ties, bx):
    from elvis_b200 import ops
    rng = np.random.default_rng(bx)
    s = random_scores(rng, (3, 6, bx), ties)
    s[0, 0, : min(3, bx)] = [0.0, -0.0, 0.0][: min(3, bx)]
    for pol in (P.REMOVE_HIGH, P.REMOVE_LOW):
        for k in (0, 1, bx // 2, bx - 1, bx):
            got = ops.select_rows(to_dev(s, dev), k, pol).cpu().numpy()
            assert np.array_equal(got, P.select_rows(s, k, pol)), (pol, k)
    kk = rng.integers(0, bx + 1, 6).astype(np.int32)
    got = ops.select_rows(to_dev(s, dev), to_dev(kk, dev), P.REMOVE_LOW).cpu().numpy()
    assert np.array_equal(got, P.select_rows(s, np.broadcast_to(kk, (3, 6)), P.REMOVE_LOW))


@pytest.mark.parametrize("H,W,bs,shrink", [(64, 96, 16, 0.5), (48, 80, 8, 0.25), (32, 64, 16, 3), (32, 64, 16, 0.0),
                                          (32, 64, 16, 0.999), (32, 64, 16, 1.0), (64, 64, 32, 0.5), (24, 36, 12, 0.34),
                                          (272, 480, 4, 0.4)])
def test_elvis_shrink_stretch_packed(dev, H, W, bs, shrink):
    from elvis_b200 import elvis as E
    rng = np.random.default_rng(H + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    sc = rng.random((H // bs, W // bs))
    g_img, g_mask, g_coords = E.apply_selective_removal(img, sc, bs, shrink)
    r_img, r_mask, r_coords = P.apply_selective_removal(img, sc, bs, shrink)
    assert g_img.shape == r_img.shape and np.array_equal(g_img, r_img)
    assert g_mask.dtype == np.int8 and np.array_equal(g_mask, r_mask) and g_coords == r_coords
    g_full = E.stretch_frame(g_img, g_mask, bs)
    assert np.array_equal(g_full, P.stretch_frame(r_img, r_mask, bs))
    keep = np.kron(1 - r_mask, np.ones((bs, bs), np.uint8)).astype(bool)
    assert np.array_equal(g_full[keep], img[keep]) and not g_full[~keep].any()   # round trip


def test_elvis_errors(dev):
    from elvis_b200 import elvis as E
    img = np.zeros((50, 64, 3), np.uint8)
    with pytest.raises(ValueError):
        E.apply_selective_removal(img, np.zeros((3, 4)), 16, 0.5)
    with pytest.raises(ValueError):
        E.split_image_into_blocks(img, 16)
    with pytest.raises(ValueError):
        E.filter_frame_gaussian(img, np.zeros((3, 4)), 16)


@pytest.mark.parametrize("H,W,bs,shrink", [(64, 96, 16, 0.5), (50, 85, 8, 0.25), (80, 128, 16, 0.3), (80, 128, 16, 0.0),
                                          (80, 128, 16, 0.99), (40, 64, 8, 0.3)])
def test_row_only_shrink_stretch(dev, H, W, bs, shrink):
    from elvis_b200 import utils as U
    rng = np.random.default_rng(H * W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    imp = np.round(rng.random((H // bs, W // bs)) * 8) / 8
    g_img, g_mask = U.shrink_frame_row_only(img, imp, bs, shrink)
    r_img, r_mask = P.shrink_frame_row_only(img, imp, bs, shrink)
    assert g_mask.dtype == bool and np.array_equal(g_mask, r_mask)
    assert g_img.shape == r_img.shape and np.array_equal(g_img, r_img)
    assert np.array_equal(U.stretch_frame_row_only(g_img, g_mask, bs), P.stretch_frame_row_only(r_img, r_mask, bs))


@pytest.mark.parametrize("H,W,bs,shrink", [(64, 96, 16, 0.5), (51, 85, 8, 0.3), (80, 128, 16, 0.9), (80, 128, 16, 0.0),
                                          (40, 64, 8, 0.62), (16, 128, 16, 0.5), (128, 16, 16, 0.5), (48, 48, 8, 1.0),
                                          (272, 480, 4, 0.45), (360, 640, 8, 0.5)])
def test_rowcol_shrink_stretch(dev, H, W, bs, shrink):
    """8f rank 2 (utils.py:763-1018): both bookkeeping variants, through the mirrors."""
    from elvis_b200 import utils as U
    rng = np.random.default_rng(H * W + bs)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    by, bx = H // bs, W // bs
    imp = np.round(rng.random((by, bx)) * 16) / 16 if bs > 4 else rng.random((by, bx))
    g = U.shrink_frame_position_map(img, imp, bs, shrink)
    r = P.shrink_frame_position_map(img, imp, bs, shrink)
    assert g[1].dtype == bool and np.array_equal(g[1], r[1])
    assert g[2].shape == r[2].shape and np.array_equal(g[2], r[2])
    assert g[0].shape == r[0].shape and np.array_equal(g[0], r[0])
    assert np.array_equal(U.stretch_frame_position_map(*g, bs), P.stretch_frame_position_map(*r, bs))
    g = U.shrink_frame_removal_indices(img, imp, bs, shrink)
    r = P.shrink_frame_removal_indices(img, imp, bs, shrink)
    assert np.array_equal(g[0], r[0]) and np.array_equal(g[1], r[1]) and len(g[2]) == len(r[2])
    assert all(a.dtype == np.int32 and np.array_equal(a, b) for a, b in zip(g[2], r[2]))
    assert np.array_equal(U.stretch_frame_removal_indices(g[0], g[2], by, bx, bs),
                          P.stretch_frame_removal_indices(r[0], r[2], by, bx, bs))
    if g[2]:
        bad = [a[:max(1, len(a) - 2)] + 3 for a in g[2]]
        assert np.array_equal(U.stretch_frame_removal_indices(g[0], bad, by, bx, bs),
                              P.stretch_frame_removal_indices(r[0], bad, by, bx, bs))


def test_rowcol_position_map_duplicates_last_wins(dev):
    """stretch_frame_position_map scatters in row-major order (utils.py:851-855): a later shrunk
    block overwrites an earlier one aimed at the same original position."""
    from elvis_b200 import utils as U
    rng = np.random.default_rng(5)
    small = rng.integers(0, 256, (24, 32, 3), dtype=np.uint8)
    pm = rng.integers(0, 4, (3, 4, 2))
    mask = np.zeros((4, 4), bool)
    assert np.array_equal(U.stretch_frame_position_map(small, mask, pm, 8), P.stretch_frame_position_map(small, mask, pm, 8))


def test_rowcol_plan_batched_matches_per_frame(dev):
    """The plan kernel runs one CTA per frame: a clip gives the per-frame results."""
    from elvis_b200 import ops
    rng = np.random.default_rng(9)
    imp = rng.random((5, 17, 30))
    target = int(17 * 30 * 0.4)
    import torch
    mask, pos, pidx, pcnt, meta = ops.rowcol_plan(torch.from_numpy(imp).to(dev), target)
    fby, fbx, counts = ops.rowcol_dims(17, 30, target)
    for t in range(5):
        m, pm, passes = P.rowcol_plan(imp[t], 0.4)
        assert meta[t].tolist() == [len(passes), fby, fbx, target]
        assert np.array_equal(mask[t].cpu().numpy().astype(bool), m)
        lin = pos[t, :fby, :fbx].cpu().numpy()
        assert np.array_equal(np.stack([lin // 30, lin % 30], -1), pm)
        assert pcnt[t, :len(passes)].tolist() == counts == [len(a) for a in passes]
        for i, a in enumerate(passes):
            assert np.array_equal(pidx[t, i, :len(a)].cpu().numpy(), a)


def test_roi_side_files_and_y4m(dev, tmp_path):
    """8f rank 3 (utils.py:453-462, 1026-1092; elvis.py:2027-2090): files byte for byte vs the oracle port."""
    from elvis_b200 import elvis as E, utils as U
    rng = np.random.default_rng(21)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    imps = [rng.random((135, 240)) for _ in range(2)] + [rng.random((17, 30))]
    imps[0][0, :5] = [0.0, 1.0, 0.5, 0.125, 0.9999]
    for base_qp, rq in ((48, 15), (3, 15), (30, 6)):
        U.create_kvazaar_roi_file(imps, a, base_qp, rq)
        P.create_kvazaar_roi_file(imps, b, base_qp, rq)
        assert open(a, "rb").read() == open(b, "rb").read()
    for (w, h, crf, rq) in ((3840, 2160, 35, 10), (480, 272, 60, 15), (1920, 1088, 2, 7), (2176, 1152, 30, 12)):
        maps = [rng.random((h // 16, w // 16)) for _ in range(2)]
        U.create_svtav1_roi_file(maps, a, crf, rq, w, h)
        P.create_svtav1_roi_file(maps, b, crf, rq, w, h)
        assert open(a).read() == open(b).read(), (w, h)
    for (w, h, bs) in ((480, 272, 16), (3840, 2160, 16), (256, 128, 32), (4352, 2304, 32), (1024, 576, 8)):
        scores = rng.random((2, h // bs, w // bs))
        scores[0, 0, :2] = [0.0, 1.0]
        E.write_per_block_qpfile(scores, bs, w, h, a)
        P.write_per_block_qpfile(scores, bs, w, h, b)
        assert open(a).read() == open(b).read(), (w, h, bs)
    for (h, w) in ((34, 50), (64, 96), (2, 2), (270, 482)):
        frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(3)]
        frames[0][0, :2] = [[255, 255, 255], [0, 0, 0]]
        U.write_y4m(frames, a, 29.97)
        P.write_y4m(frames, b, 29.97)
        assert open(a, "rb").read() == open(b, "rb").read(), (h, w)
    with pytest.raises(ValueError):
        U.write_y4m([np.zeros((5, 8, 3), np.uint8)], a, 30.0)


def test_rgb_to_i420_strided_and_resize_area_shapes(dev):
    import torch
    from elvis_b200 import ops
    from oracle import spec_cv
    rng = np.random.default_rng(22)
    big = rng.integers(0, 256, (2, 40, 70, 3), dtype=np.uint8)
    view = to_dev(big, dev)[:, 3:35, 5:59]                     # cropped, misaligned view: byte path
    got = ops.rgb_to_i420(view).cpu().numpy()
    for t in range(2):
        assert np.array_equal(got[t], spec_cv.rgb_to_i420(big[t, 3:35, 5:59]).reshape(-1))
    for (sh, sw, dh, dw) in [(135, 240, 34, 60), (136, 240, 34, 60), (16, 16, 8, 8), (34, 60, 17, 30), (10, 34, 5, 17), (9, 13, 4, 5), (7, 9, 7, 9)]:
        a = (rng.random((3, sh, sw)) * 2 - 1).astype(np.float32)
        got = ops.resize_area_f32(to_dev(a, dev), dh, dw).cpu().numpy()
        for t in range(3):
            assert np.array_equal(got[t], spec_cv.resize_area_f32(a[t], dh, dw)), (sh, sw, dh, dw)
    with pytest.raises(NotImplementedError):
        ops.resize_area_f32(torch.zeros((1, 4, 4), dtype=torch.float32, device=dev), 8, 4)


def test_pyramid_reconstruction_and_map_video(dev):
    """8f rank 4 (elvis.py:2522-2600, 2198-2245) through the mirrors, against the oracle port."""
    import cv2
    from elvis_b200 import elvis as E, ops
    rng = np.random.default_rng(23)

    def up_cubic(im):
        return cv2.resize(im, None, fx=2, fy=2, interpolation=cv2.INTER_CUBIC)

    def up_repeat(im):
        return np.ascontiguousarray(im.repeat(2, 0).repeat(2, 1))

    for (H, W, bs, top) in [(64, 96, 16, 4), (64, 96, 16, 2), (48, 80, 8, 3), (64, 64, 32, 5), (32, 48, 16, 0), (272, 480, 16, 3)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        maps = rng.integers(0, top + 1, (H // bs, W // bs))
        maps.flat[0] = top
        for up in (up_cubic, up_repeat):
            got = E.upscale_realesrgan_adaptive(img, maps.copy(), bs, upsample_fn=up)
            assert np.array_equal(got, P.upscale_realesrgan_adaptive(img, maps.copy(), bs, up)), (H, W, bs, top)
    with pytest.raises(NotImplementedError):
        E.upscale_realesrgan_adaptive(img, maps, 16)
    for f in (1, 2, 4, 8, 16):
        clip = rng.integers(0, 256, (2, 64, 96, 3), dtype=np.uint8)
        got = ops.area_downscale(to_dev(clip, dev), f).cpu().numpy()
        for t in range(2):
            assert np.array_equal(got[t], P.area_downscale(clip[t], f)), f
    for hi, rng_max in ((10, 10.0), (4, 4), (3, 3)):
        m = rng.integers(0, hi + 1, (3, 17, 30)).astype(np.int32)
        m.flat[:2] = [0, hi]
        gray = E.strength_maps_to_gray(m)
        assert gray.dtype == np.uint8 and np.array_equal(gray, P.strength_maps_to_gray(m))
        assert np.array_equal(E.gray_to_strength_maps(gray, 0.0, rng_max), m)
        noisy = rng.integers(0, 256, gray.shape, dtype=np.uint8)          # what a lossy codec hands back
        assert np.array_equal(E.gray_to_strength_maps(noisy, 0.0, rng_max), P.gray_to_strength_maps(noisy, 0.0, rng_max))


def test_resize_nearest_matches_cv2(dev):
    """Map resize of the mirrors (elvis.py:1189-1193, utils.py:1343-1345): cv2's INTER_NEAREST index rule."""
    import cv2
    from elvis_b200 import ops
    rng = np.random.default_rng(31)
    for (sh, sw, dh, dw) in [(1080, 1920, 67, 120), (17, 30, 34, 60), (33, 58, 34, 60), (5, 7, 11, 13), (100, 37, 9, 80), (64, 64, 64, 64)]:
        a = rng.integers(0, 256, (2, sh, sw), dtype=np.uint8)
        f = rng.random((2, sh, sw)).astype(np.float32)
        d = rng.random((2, sh, sw))
        for arr in (a, f, d):
            got = ops.resize_nearest(to_dev(arr, dev), dh, dw).cpu().numpy()
            for t in range(2):
                assert np.array_equal(got[t], cv2.resize(arr[t], (dw, dh), interpolation=cv2.INTER_NEAREST)), (sh, sw, dh, dw, arr.dtype)


def test_planar_pipeline_matches_per_plane_oracle(dev):
    """Planar YUV 4:2:0: mask from luma scores, applied to chroma at half block size."""
    from elvis_b200.pipeline import ElvisV1, Yuv420
    T, H, W, bs = 6, 96, 160, 16
    y, u, v = synth_yuv420(T, H, W, seed=21)
    clip = Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev))
    pipe = ElvisV1(bs, 0.5, 0.5, 0.5)
    scores, mask, shrunk, stretched = pipe.run(clip)
    scores, mask = scores.cpu().numpy(), mask.cpu().numpy()
    rsc, rtc = spec_scoring.sc_tc(y, bs)
    ref_scores = P.combine_removability(rsc, rtc, 0.5, 0.5)
    np.testing.assert_allclose(scores, ref_scores, rtol=RTOL, atol=1e-9)
    k = P.blocks_to_remove_elvis(0.5, W // bs)
    # the mask must be the exact top-k of the scores the GPU itself produced ...
    assert np.array_equal(mask, P.select_rows(scores, k, P.REMOVE_HIGH))
    # ... and equal to the oracle's wherever the oracle's decision margin exceeds the tolerance
    ref_mask = P.select_rows(ref_scores, k, P.REMOVE_HIGH)
    srt = np.sort(ref_scores, axis=-1)[..., ::-1]
    margin = srt[..., k - 1] - srt[..., k]
    safe = margin > 2 * RTOL
    assert safe.mean() > 0.5 and np.array_equal(mask[safe], ref_mask[safe])
    for name, plane, pb in (("y", y, bs), ("u", u, bs // 2), ("v", v, bs // 2)):
        g_s = getattr(shrunk, name).cpu().numpy()
        g_f = getattr(stretched, name).cpu().numpy()
        for t in range(T):
            r_s = P.shrink_plane(plane[t], mask[t], pb)
            assert np.array_equal(g_s[t], r_s), (name, t)
            assert np.array_equal(g_f[t], P.stretch_plane(r_s, mask[t], pb)), (name, t)


# ---------------------------------------------------------------------------------- a8-a12
@pytest.mark.parametrize("bs", [8, 16, 32])
def test_elvis_filters_packed(dev, bs):
    from elvis_b200 import elvis as E
    rng = np.random.default_rng(bs)
    H, W = bs * 5, bs * 7
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    sc = rng.random((5, 7))
    sc[0, :5] = [0.5, 0.125, 1.0, 0.0, 0.25]     # exact .5 products -> round-half-even cases
    for name in ("filter_frame_downsample", "filter_frame_gaussian"):
        g_img, g_map = getattr(E, name)(img, sc, bs)
        r_img, r_map = getattr(P, name)(img, sc, bs)
        assert g_map.dtype == np.int32 and np.array_equal(g_map, r_map), name
        assert np.array_equal(g_img, r_img), (name, np.argwhere(g_img != r_img)[:5])


@pytest.mark.parametrize("bs", [8, 16])
def test_utils_degrade_packed_with_crop(dev, bs):
    from elvis_b200 import utils as U
    rng = np.random.default_rng(bs + 1)
    img = rng.integers(0, 256, (bs * 5 + 3, bs * 7 + 5, 3), dtype=np.uint8)
    imp = rng.random((5, 7))
    for name in ("degrade_adaptive_downsample", "degrade_adaptive_blur"):
        g_img, g_map = getattr(U, name)(img, imp, bs)
        r_img, r_map = getattr(P, name)(img, imp, bs)
        assert np.array_equal(g_map, r_map), name
        assert np.array_equal(g_img, r_img), name


def test_presley_degrade_video(dev):
    from elvis_b200 import presley as Pr
    rng = np.random.default_rng(77)
    frames = [rng.integers(0, 256, (64, 96, 3), dtype=np.uint8) for _ in range(3)]
    imps = [rng.random((4, 6)) for _ in range(3)]
    for method, tag in ((Pr.downscale_block, "downscale"), (Pr.blur_block, "blur")):
        out, maps = Pr.degrade_video_adaptive(frames, imps, 16, 4, method)
        for f, imp, o, m in zip(frames, imps, out, maps):
            rm = P.levels_inverted_round(imp, 4)
            assert np.array_equal(m, rm)
            assert np.array_equal(o, P.presley_degrade_frame(f, rm, 16, tag)), tag
    blk = frames[0][:16, :16]
    assert np.array_equal(Pr.downscale_block(blk, 3), P.presley_degrade_frame(blk, np.array([[3]]), 16, "downscale"))
    assert np.array_equal(Pr.blur_block(blk, 2), P.presley_degrade_frame(blk, np.array([[2]]), 16, "blur"))


@pytest.mark.parametrize("movers", ["tma", "cp.async", "tma-per-warp"])
@pytest.mark.parametrize("W", [48, 96, 128, 272, 400])       # 3 / 6 / 8 / 17 / 25 blocks per row: partial tiles, several tiles, odd chroma pitch
def test_fused_downsample_movers(dev, monkeypatch, movers, W):
    """The fused Y+U+V power-of-two downsample through both tile movers (TMA boxes with swizzle; cp.async pieces),
    levels 0..4, on contiguous planes and on planes that are windows of wider buffers."""
    import torch
    from elvis_b200.pipeline import PresleyV2, Yuv420
    monkeypatch.setenv("ELVIS_DOWNSAMPLE_TMA", {"tma": "1", "cp.async": "0", "tma-per-warp": "2"}[movers])
    T, H, bs = 2, 48, 16
    y, u, v = synth_yuv420(T, H, W, seed=W)
    rng = np.random.default_rng(W)
    y[1] = rng.integers(0, 256, y[1].shape, dtype=np.uint8)
    levels = rng.integers(0, 5, (T, H // bs, W // bs)).astype(np.int32)
    v2 = PresleyV2(bs)

    def windowed(a, pad):                    # the same pixels as a window of a wider, taller buffer (row pitch > width)
        big = torch.full((a.shape[0], a.shape[1] + 2, a.shape[2] + pad), 7, dtype=torch.uint8, device=dev)
        big[:, 1:-1, 16:16 + a.shape[2]] = to_dev(a, dev)
        return big[:, 1:-1, 16:16 + a.shape[2]]

    def check(d):
        for name, plane, pb, cap in (("y", y, bs, 4), ("u", u, bs // 2, 3), ("v", v, bs // 2, 3)):
            for t in range(T):
                small = np.maximum(1, pb >> np.minimum(levels[t], cap))
                assert np.array_equal(getattr(d, name)[t].cpu().numpy(), P.downsample_plane(plane[t], small, pb)), (name, t)

    check(v2.downsample_pow2(Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev)), to_dev(levels, dev), 4))
    # source AND destination as windows of wider buffers: the result is right and nothing outside the windows is written
    src = Yuv420(windowed(y, 48), windowed(u, 32), windowed(v, 32))
    dst = Yuv420(windowed(np.zeros_like(y), 48), windowed(np.zeros_like(u), 32), windowed(np.zeros_like(v), 32))
    check(v2.downsample_pow2(src, to_dev(levels, dev), 4, dst))
    for plane in dst.planes:
        big = plane._base if plane._base is not None else plane
        plane.fill_(7)
        assert bool((big == 7).all())


@pytest.mark.parametrize("movers", ["tma", "direct"])
@pytest.mark.parametrize("pb,by,bx", [(16, 3, 6), (16, 1, 17), (8, 5, 7), (8, 2, 16), (8, 1, 1)])
def test_blur_tensor_core_movers(dev, monkeypatch, movers, pb, by, bx):
    """The tensor-core blur with TMA boxes (per-warp pipeline) and with direct loads / stores: rounds 0..10, odd block
    counts (half-empty chroma tiles), contiguous planes and windows of wider buffers."""
    import torch
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_BLUR_TMA", "1" if movers == "tma" else "0")
    T, H, W = 3, by * pb, bx * pb
    rng = np.random.default_rng(pb * 1000 + by * 31 + bx)
    plane = rng.integers(0, 256, (T, H, W), dtype=np.uint8)
    plane[0] = synth_luma(1, H, W, seed=bx)[0]
    rounds = rng.integers(0, 11, (T, by, bx)).astype(np.int32)
    rounds[0, 0, 0] = 0
    want = np.stack([P.blur_plane(plane[t], rounds[t], pb) for t in range(T)])
    got = ops.degrade_blur(to_dev(plane, dev), to_dev(rounds, dev), pb)
    assert np.array_equal(got.cpu().numpy(), want)
    big = torch.full((T, H + 3, (W + 48 + 15) // 16 * 16), 9, dtype=torch.uint8, device=dev)   # row pitch a multiple of 16: TMA-capable even for odd W
    big[:, 2:2 + H, 32:32 + W] = to_dev(plane, dev)
    out = torch.zeros_like(big)
    ops.degrade_blur(big[:, 2:2 + H, 32:32 + W], to_dev(rounds, dev), pb, out=out[:, 2:2 + H, 32:32 + W])
    assert np.array_equal(out[:, 2:2 + H, 32:32 + W].cpu().numpy(), want)
    out[:, 2:2 + H, 32:32 + W] = 0
    assert int(out.count_nonzero()) == 0                                          # nothing written outside the window


@pytest.mark.parametrize("pb,smalls", [(16, (16, 8, 5, 4)), (16, (16, 3, 8, 7, 2, 1)), (8, (8, 5, 3, 2)), (8, (8, 4, 2, 2))])
def test_downsample_mixed_levels(dev, pb, smalls):
    """utils' level set (utils.py:1142-1148: block -> 8, 5, 4 pixels) mixes power-of-two reductions with fractional ones:
    the former run through the closed-form kernel, the latter through the table-driven one, in the same call."""
    from elvis_b200 import ops
    T, by, bx = 2, 5, 9
    rng = np.random.default_rng(pb + len(smalls))
    plane = rng.integers(0, 256, (T, by * pb, bx * pb), dtype=np.uint8)
    plane[0] = synth_luma(1, by * pb, bx * pb, seed=3)[0]
    levels = rng.integers(0, len(smalls), (T, by, bx)).astype(np.int32)
    got = ops.degrade_downsample(to_dev(plane, dev), to_dev(levels, dev), pb, smalls).cpu().numpy()
    small_map = np.asarray(smalls)[levels]
    for t in range(T):
        assert np.array_equal(got[t], P.downsample_plane(plane[t], small_map[t], pb)), t


@pytest.mark.parametrize("H,W", [(5, 7), (16, 64), (33, 130), (8, 4)])
def test_split_and_merge_channels(dev, H, W):
    """Packed (T, H, W, 3) <-> three planes: the word path (aligned, width % 4 == 0), the byte path, and windows."""
    import torch
    from elvis_b200 import ops
    rng = np.random.default_rng(H * W)
    clip = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    planes = ops.split_channels3(to_dev(clip, dev))
    assert np.array_equal(planes.cpu().numpy(), np.moveaxis(clip, 3, 0))
    assert np.array_equal(ops.merge_channels3(planes).cpu().numpy(), clip)
    big = torch.zeros((2, H + 2, W + 5, 3), dtype=torch.uint8, device=dev)
    big[:, 1:-1, 2:2 + W] = to_dev(clip, dev)
    assert np.array_equal(ops.split_channels3(big[:, 1:-1, 2:2 + W]).cpu().numpy(), np.moveaxis(clip, 3, 0))
    out = torch.full_like(big, 9)
    ops.merge_channels3(planes, out=out[:, 1:-1, 2:2 + W])
    assert np.array_equal(out[:, 1:-1, 2:2 + W].cpu().numpy(), clip)
    out[:, 1:-1, 2:2 + W] = 9
    assert bool((out == 9).all())


def test_planar_degrade(dev):
    from elvis_b200 import ops
    from elvis_b200.pipeline import PresleyV2, Yuv420
    T, H, W, bs = 3, 64, 96, 16
    y, u, v = synth_yuv420(T, H, W, seed=4)
    clip = Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev))
    rng = np.random.default_rng(8)
    rounds = rng.integers(0, 11, (T, H // bs, W // bs)).astype(np.int32)
    levels = rng.integers(0, 4, (T, H // bs, W // bs)).astype(np.int32)
    v2 = PresleyV2(bs)
    b = v2.blur(clip, to_dev(rounds, dev))
    d = v2.downsample_pow2(clip, to_dev(levels, dev), 3)
    for name, plane, pb in (("y", y, bs), ("u", u, bs // 2), ("v", v, bs // 2)):
        for t in range(T):
            assert np.array_equal(getattr(b, name)[t].cpu().numpy(), P.blur_plane(plane[t], rounds[t], pb)), name
            small = np.maximum(1, pb >> levels[t])
            assert np.array_equal(getattr(d, name)[t].cpu().numpy(), P.downsample_plane(plane[t], small, pb)), name
    packed = ops.pack_levels_2bit(to_dev(levels, dev))
    assert np.array_equal(packed.cpu().numpy(), P.pack_levels_2bit(levels))
    assert np.array_equal(ops.unpack_levels_2bit(packed, W // bs).cpu().numpy(), levels)


# ---------------------------------------------------------------------------------- a13 / a14
def test_mask_bit_packing(dev):
    from elvis_b200 import elvis as E
    rng = np.random.default_rng(2)
    m = (rng.random((5, 7, 13)) > 0.5).astype(np.int8)
    packed, shape = E.pack_removal_masks(m)
    ref, _ = P.pack_masks(m)
    assert np.array_equal(packed, ref) and shape == m.shape
    assert np.array_equal(E.unpack_removal_masks(packed, shape), m)


@pytest.mark.parametrize("impl", ["packed", "pair", "scalar", "hmma"])      # packed fp32 / two tiles per thread / round-1 scalar kernel / tensor cores
@pytest.mark.parametrize("pb", [8, 16])
def test_dct_dampen(dev, monkeypatch, pb, impl):
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_DAMPEN_IMPL", impl)
    T, H, W = 2, pb * 4, pb * (7 if impl == "hmma" else 6)        # an odd block count leaves the last tile of the hmma path half empty
    y = synth_luma(T, H, W, seed=pb)
    rng = np.random.default_rng(pb)
    s = rng.random((T, H // pb, W // pb)).astype(np.float32)
    s[0, 0, 0], s[0, 0, 1] = 0.0, 1.0
    out = ops.dct_dampen(to_dev(y, dev), to_dev(s, dev), pb).cpu().numpy()
    for t in range(T):
        ref_f = spec_dct_dampen.dampen_plane(y[t], s[t], pb, return_float=True)
        # the u8 result is the rounding of a value within RTOL (relative to full scale) of the spec
        assert np.all(np.abs(out[t].astype(np.float64) - np.clip(ref_f, 0, 255)) <= 0.5 + RTOL * 255)
        ref = spec_dct_dampen.dampen_plane(y[t], s[t], pb)
        assert (out[t] != ref).mean() < 1e-3
    assert np.array_equal(out[0, :pb, :pb], y[0, :pb, :pb])     # strength 0 is the identity
    packed = np.repeat(y[..., None], 3, axis=-1)                  # packed 3-channel path
    outp = ops.dct_dampen(to_dev(packed, dev), to_dev(s, dev), pb).cpu().numpy()
    assert np.array_equal(outp[..., 1], out)


def test_invariant_division_is_the_correctly_rounded_quotient(dev):
    """normalize_ divides by a per-launch constant with a reciprocal + two FMA corrections (csrc/common.cuh
    InvariantDivisor): it must equal IEEE division bit for bit, including divisors whose significand is all ones
    or a power of two, numerators at the extremes, zeros, NaN and infinity, and spans that take the general path."""
    import torch
    from elvis_b200 import ops
    rng = np.random.default_rng(2024)
    n = 1 << 20
    ones = np.frombuffer(np.uint64(0x3FFFFFFFFFFFFFFF).tobytes(), dtype=np.float64)[0]      # 1.999..., significand all ones
    spans = [1.0, ones, ones * 2.0 ** -7, 3.0, 1.0 / 3.0, 0.1, float(np.nextafter(1.0, 2.0)), 7.3e12, 2.0 ** -400, 2.0 ** 499,
             2.0 ** -600, 2.0 ** 600, 5e-324 * 2 ** 20] + list(np.exp(rng.uniform(-40, 40, 12)))
    for i, span in enumerate(spans):
        lo = [0.0, -1.5, 0.37][i % 3] if 2.0 ** -60 < span < 2.0 ** 60 else 0.0
        hi = lo + span
        x = lo + rng.random(n) * (hi - lo)
        bits = rng.integers(0, 1 << 52, n // 4, dtype=np.uint64) | (np.uint64(1023) << np.uint64(52))
        x[: n // 4] = lo + (bits.view(np.float64) - 1.0) * (hi - lo)            # full-width significands
        x[-8:] = [lo, hi, 0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324]
        x[-16:-8] = lo + np.array([2.0 ** -700, 2.0 ** -510, 2.0 ** -490, 2.0 ** 490, 2.0 ** 510, 2.0 ** 700, 1e-310, 1.0])
        with np.errstate(all="ignore"):
            want = (x - lo) / (hi - lo)
        got = ops.normalize_(to_dev(x, dev), torch.tensor([lo, hi], dtype=torch.float64, device=dev)).cpu().numpy()
        same = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (span, x[~same][:4], got[~same][:4], want[~same][:4])


@pytest.mark.parametrize("sort_path", [False, True])
def test_normalize_select_rows_equals_the_two_calls(dev, monkeypatch, sort_path):
    """elvis_normalize_select_rows == elvis_normalize then elvis_select_rows: same normalised scores (bits), same mask."""
    import torch
    from elvis_b200 import ops
    if sort_path:
        monkeypatch.setenv("ELVIS_SELECT_SORT", "1")
    rng = np.random.default_rng(5)
    for bx, ties in ((240, "none"), (120, "quantised"), (37, "all"), (300, "none")):
        raw = random_scores(rng, (3, 7, bx), ties) * 3.7 - 1.2
        mm = torch.tensor([raw.min(), raw.max()], dtype=torch.float64, device=dev)
        for pol in (P.REMOVE_HIGH, P.REMOVE_LOW):
            for k in (0, 1, bx // 2, bx):
                two = ops.normalize_(to_dev(raw, dev), mm)
                mask2 = ops.select_rows(two, k, pol)
                one = to_dev(raw, dev)
                mask1 = ops.select_rows(one, k, pol, normalize_with=mm)
                assert torch.equal(one, two) and torch.equal(mask1, mask2)
                ref = (raw - raw.min()) / (raw.max() - raw.min()) if raw.max() > raw.min() else raw
                assert np.array_equal(one.cpu().numpy(), ref)
                assert np.array_equal(mask1.cpu().numpy(), P.select_rows(ref, k, pol))
        kr = rng.integers(0, bx + 1, 7).astype(np.int32)                      # per-row k
        one = to_dev(raw, dev)
        mask1 = ops.select_rows(one, to_dev(kr, dev), P.REMOVE_LOW, normalize_with=mm)
        assert np.array_equal(mask1.cpu().numpy(), P.select_rows((raw - raw.min()) / (raw.max() - raw.min()) if raw.max() > raw.min() else raw, kr, P.REMOVE_LOW))


def test_select_rows_cta_sort_path(dev, monkeypatch):
    """Short rows normally take the warp quickselect; force the CTA bitonic path as well."""
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_SELECT_SORT", "1")
    rng = np.random.default_rng(99)
    for bx, ties in ((120, "none"), (240, "quantised"), (37, "all")):
        s = random_scores(rng, (2, 5, bx), ties)
        for pol in (P.REMOVE_HIGH, P.REMOVE_LOW):
            for k in (1, bx // 2, bx - 1):
                assert np.array_equal(ops.select_rows(to_dev(s, dev), k, pol).cpu().numpy(), P.select_rows(s, k, pol))


def test_select_rows_sorted_and_adversarial_rows(dev):
    """Quickselect must not depend on input order: sorted, reversed, constant-with-one-outlier."""
    from elvis_b200 import ops
    bx = 240
    base = np.linspace(0, 1, bx)
    rows = np.stack([base, base[::-1], np.full(bx, 0.25), np.r_[np.full(bx - 1, 0.25), 0.9], np.r_[0.1, np.full(bx - 1, 0.25)],
                     np.where(np.arange(bx) % 2 == 0, 0.0, 1.0)])[None]
    for pol in (P.REMOVE_HIGH, P.REMOVE_LOW):
        for k in (1, 60, 120, 239):
            assert np.array_equal(ops.select_rows(to_dev(rows, dev), k, pol).cpu().numpy(), P.select_rows(rows, k, pol))


@pytest.mark.parametrize("impl", ["simt", "mma", "tma", "umma"])
@pytest.mark.parametrize("H,W,T", [(48, 128, 5), (64, 96, 7), (112, 400, 9), (1080, 1920, 3), (16, 16, 2)])
def test_score_impls_agree_with_spec(dev, monkeypatch, impl, H, W, T):
    """The CUDA-core kernel, the tensor-core kernel with direct loads and the tensor-core
    kernel fed by TMA must all match the spec (16x16 blocks; edge tiles, partial block rows)."""
    from elvis_b200 import ops
    monkeypatch.setenv("ELVIS_SCORE_IMPL", impl)
    monkeypatch.setenv("ELVIS_SCORE_CHUNK", "4")
    y = synth_luma(T, H, W, seed=H + W + T)
    rng = np.random.default_rng(T)
    y[-1] = rng.integers(0, 256, (H, W), dtype=np.uint8)      # full-range noise frame
    yd = to_dev(y, dev)
    sc, tc, mm = ops.score_sc_tc(yd, 16)
    rsc, rtc = spec_scoring.sc_tc(y, 16)
    np.testing.assert_allclose(sc.cpu().numpy(), rsc, rtol=RTOL, atol=0)
    np.testing.assert_allclose(tc.cpu().numpy(), rtc, rtol=RTOL, atol=0)
    s, t = sc.cpu().numpy(), tc.cpu().numpy()
    assert mm.cpu().numpy().tolist() == [s.min(), s.max(), t.min(), t.max()]
    if T > 2:
        sch, tch, _ = ops.score_sc_tc(yd[2:], 16, prev_halo=yd[1])
        rsc_h, rtc_h = spec_scoring.sc_tc(y[2:], 16, prev=y[1])
        np.testing.assert_allclose(sch.cpu().numpy(), rsc_h, rtol=RTOL, atol=0)
        np.testing.assert_allclose(tch.cpu().numpy(), rtc_h, rtol=RTOL, atol=0)


def test_score_tightness(dev, monkeypatch):
    """Report (and bound) the actual relative error of each implementation: far below RTOL."""
    from elvis_b200 import ops
    y = synth_luma(12, 96, 256, seed=77)
    rsc, rtc = spec_scoring.sc_tc(y, 16)
    for impl in ("simt", "mma", "tma", "umma"):
        monkeypatch.setenv("ELVIS_SCORE_IMPL", impl)
        sc, tc, _ = ops.score_sc_tc(to_dev(y, dev), 16)
        e_sc = np.abs(sc.cpu().numpy() - rsc).max() / np.abs(rsc).max()
        nz = rtc > 0
        e_tc = (np.abs(tc.cpu().numpy() - rtc)[nz] / rtc[nz]).max()
        print(f"{impl}: max rel err SC {e_sc:.2e} TC {e_tc:.2e}")
        assert e_sc < 2e-5 and e_tc < 2e-5


@pytest.mark.parametrize("prepared", [True, False])
@pytest.mark.parametrize("depth,split", [(2, False), (3, True)])
def test_pipelined_equals_serial(dev, depth, split, prepared):
    """Clips in flight on two (score | move) or three (score | shrink | stretch) streams must give
    exactly the serial results."""
    import torch
    from elvis_b200.pipeline import ElvisV1, ElvisV1Pipelined, Yuv420
    T, H, W, bs = 5, 96, 160, 16
    clips = []
    for seed in (1, 2, 3):
        y, u, v = synth_yuv420(T, H, W, seed=seed)
        clips.append(Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev)))
    serial = ElvisV1(bs, 0.5, 0.5, 0.5)
    ref = [serial.run(c) for c in clips]
    pp = ElvisV1Pipelined(T, H, W, bs, 0.5, 0.5, 0.5, dev, depth=depth, split_stretch=split, prepared=prepared)
    for rep in range(2):            # the second round replays the prepared calls of (clip, slot) pairs seen before
        for i, c in enumerate(clips * 2):
            slot = pp.submit(c)
            slot["done"].synchronize()
            assert torch.equal(slot["mask"], ref[i % 3][1]) and torch.equal(slot["full"].y, ref[i % 3][3].y)
    assert (len(pp._programs) > 0) == prepared
    for i, c in enumerate(clips):
        slot = pp.submit(c)
        slot["done"].synchronize()      # read the slot before it is reused
        scores, mask, shrunk, full = ref[i]
        assert torch.equal(slot["scores"], scores) and torch.equal(slot["mask"], mask)
        for a, b in zip(slot["shrunk"].planes, shrunk.planes):
            assert torch.equal(a, b)
        for a, b in zip(slot["full"].planes, full.planes):
            assert torch.equal(a, b)
    # back-to-back submissions without waiting: the last two clips must still be intact
    s1 = pp.submit(clips[0])
    s2 = pp.submit(clips[1])
    pp.join()
    torch.cuda.synchronize()
    assert torch.equal(s1["mask"], ref[0][1]) and torch.equal(s2["mask"], ref[1][1])
    assert torch.equal(s1["full"].y, ref[0][3].y) and torch.equal(s2["full"].y, ref[1][3].y)


@pytest.mark.parametrize("bs,T,amount", [(8, 3, 0.25), (32, 2, 0.5), (16, 1, 0.5), (16, 3, 0.0), (16, 2, 1.0), (16, 2, 7)])
def test_planar_pipeline_edge_configs(dev, bs, T, amount):
    """Other block sizes (4x4 / 16x16 chroma blocks -> per-plane kernels), single-frame clips (no
    smoothing, elvis.py:1202), nothing / one block / an absolute count removed (elvis.py:1392-1396)."""
    from elvis_b200.pipeline import ElvisV1, Yuv420
    H, W = bs * 4, bs * 10
    y, u, v = synth_yuv420(T, H, W, seed=bs + T)
    clip = Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev))
    scores, mask, shrunk, full = ElvisV1(bs, amount, 0.4, 0.5).run(clip)
    scores, mask = scores.cpu().numpy(), mask.cpu().numpy()
    rsc, rtc = spec_scoring.sc_tc(y, bs)
    np.testing.assert_allclose(scores, P.combine_removability(rsc, rtc, 0.4, 0.5), rtol=RTOL, atol=1e-9)
    k = P.blocks_to_remove_elvis(amount, W // bs)
    assert np.array_equal(mask, P.select_rows(scores, k, P.REMOVE_HIGH)) and (mask.sum(-1) == k).all()
    for name, plane, pb in (("y", y, bs), ("u", u, bs // 2), ("v", v, bs // 2)):
        for t in range(T):
            r_s = P.shrink_plane(plane[t], mask[t], pb)
            assert np.array_equal(getattr(shrunk, name)[t].cpu().numpy(), r_s), (name, t)
            assert np.array_equal(getattr(full, name)[t].cpu().numpy(), P.stretch_plane(r_s, mask[t], pb)), (name, t)


def test_host_pipeline_matches_device_pipeline(dev):
    """HostElvisV1 (pinned host buffers in/out, two streams) returns what ElvisV1 computes."""
    import torch
    from elvis_b200.pipeline import ElvisV1, HostElvisV1, Yuv420
    T, H, W, bs = 4, 64, 160, 16
    host = HostElvisV1(T, H, W, bs, 0.5, 0.5, 0.5, dev, depth=2)
    outs = [host.host_buffers(pinned=True) for _ in range(2)]
    clips, events = [], []
    for i in range(3):
        y, u, v = synth_yuv420(T, H, W, seed=40 + i)
        i420 = torch.from_numpy(np.concatenate([y.reshape(T, -1), u.reshape(T, -1), v.reshape(T, -1)], axis=1)).pin_memory()
        clips.append((y, u, v, i420))
    for i, (y, u, v, i420) in enumerate(clips):
        ev = host.process(i420, *outs[i % 2])
        ev.synchronize()
        shrunk_h, full_h, mask_h = outs[i % 2]
        ref = ElvisV1(bs, 0.5, 0.5, 0.5).run(Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev)))
        assert torch.equal(mask_h, ref[1].cpu())
        sw = ref[2].y.shape[2]
        s = Yuv420.from_i420(shrunk_h, H, sw)
        f = Yuv420.from_i420(full_h, H, W)
        for a, b in zip(s.planes, ref[2].planes):
            assert torch.equal(a, b.cpu())
        for a, b in zip(f.planes, ref[3].planes):
            assert torch.equal(a, b.cpu())
    assert host.h2d_bytes == T * H * W * 3 // 2


def test_sharding_world_one_equals_local(dev):
    """The sharded scorers with a single rank (no communication) equal the local pipeline."""
    import torch
    from elvis_b200 import sharding
    from elvis_b200.pipeline import ElvisV1, Yuv420
    T, H, W, bs = 5, 48, 96, 16
    y, u, v = synth_yuv420(T, H, W, seed=77)
    halo = sharding.HaloClip(T, H, W, dev)
    halo.owned.copy_(to_dev(y, dev))
    got = sharding.sharded_removability(halo, T, bs, 0.5, 0.5, 0, 1)
    ref = ElvisV1(bs, 0.5, 0.5, 0.5).score(Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev)))
    assert torch.equal(got, ref)
    imp = sharding.sharded_importance(halo, bs, 0.5, 0.5, 0, 1).cpu().numpy()
    rsc, rtc = spec_scoring.sc_tc(y, bs)
    np.testing.assert_allclose(imp, P.importance_scores(rsc, rtc, 0.5, 0.5, np.ones_like(rsc)), rtol=RTOL, atol=1e-7)
    # the block-sized transform (the reference's EVCA call) through the sharded scorers
    got16 = sharding.sharded_removability(halo, T, bs, 0.5, 0.5, 0, 1, dct_size=bs)
    assert torch.equal(got16, ElvisV1(bs, 0.5, 0.5, 0.5, dct_size=bs).score(Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev))))
    assert not torch.equal(got16, got)
    imp16 = sharding.sharded_importance(halo, bs, 0.5, 0.5, 0, 1, dct_size=bs).cpu().numpy()
    rsc, rtc = spec_scoring.sc_tc(y, bs, bs)
    np.testing.assert_allclose(imp16, P.importance_scores(rsc, rtc, 0.5, 0.5, np.ones_like(rsc)), rtol=RTOL, atol=1e-7)


def test_api_errors_and_strided_inputs(dev):
    import torch
    from elvis_b200 import ops
    y = to_dev(synth_luma(2, 32, 64, seed=1), dev)
    with pytest.raises(Exception):
        ops.score_sc_tc(y, 12)                        # unsupported block size
    with pytest.raises(ValueError):
        ops.shrink(y, torch.zeros((2, 2, 4), dtype=torch.uint8, device=dev), 12, 2)   # 32 % 12 != 0
    with pytest.raises(TypeError):
        ops.select_rows(torch.zeros((1, 2, 3), dtype=torch.float32, device=dev), 1)
    # a non-contiguous (column-cropped) clip goes through the strided path of shrink/stretch
    wide = to_dev(np.random.default_rng(0).integers(0, 256, (2, 32, 96), dtype=np.uint8), dev)
    crop = wide[:, :, 16:80]
    mask = (torch.rand((2, 2, 4), device=dev) > 0.5).to(torch.uint8)
    mask[:, :, 0] = 0
    mask[:, :, 1] = 1
    mask[:, :, 2] = 0
    mask[:, :, 3] = 1
    s = ops.shrink(crop, mask, 16, 2)
    for t in range(2):
        assert np.array_equal(s[t].cpu().numpy(), P.shrink_plane(crop[t].cpu().numpy(), mask[t].cpu().numpy(), 16))


# ---------------------------------------------------------------------------------- 8f rank 1
@pytest.mark.parametrize("bs", [8, 16])
def test_unsharp_restorer_elvis(dev, bs):
    from elvis_b200 import elvis as E
    rng = np.random.default_rng(bs + 100)
    img = rng.integers(0, 256, (bs * 3, bs * 5, 3), dtype=np.uint8)
    maps = rng.integers(0, 11, (3, 5))
    maps[0, 0] = 0
    assert np.array_equal(E.restore_blur_opencv_unsharp_mask(img, maps, bs), P.restore_blur_opencv_unsharp_mask(img, maps, bs))


@pytest.mark.parametrize("halo,tb", [(0, 0.0), (4, 0.0), (8, 0.1), (0, 0.3), (20, 0.25)])
def test_unsharp_restorer_utils(dev, halo, tb):
    from elvis_b200 import utils as U
    rng = np.random.default_rng(halo + 7)
    frames = [rng.integers(0, 256, (53, 85, 3), dtype=np.uint8) for _ in range(4)]
    maps = rng.integers(0, 5, (4, 3, 5))
    got = U.restore_with_opencv_unsharp(frames, maps, 16, halo=halo, temporal_blend=tb)
    ref = P.restore_with_opencv_unsharp(frames, maps, 16, halo, tb)
    assert len(got) == 4 and all(np.array_equal(a, b) for a, b in zip(got, ref))
    assert all(np.array_equal(a, b) for a, b in zip(U.restore_with_opencv_lanczos(frames, maps, 16, halo=halo, temporal_blend=tb), ref))
    # fewer maps than frames: the rest is returned untouched (utils.py:1341)
    got = U.restore_with_opencv_unsharp(frames, maps[:2], 16)
    assert np.array_equal(got[3], frames[3]) and np.array_equal(got[2], frames[2])


def test_mirrors_resize_mismatched_maps_like_the_reference(dev, tmp_path):
    """Maps that arrive at another resolution than the frame's block grid are brought to it with
    INTER_NEAREST on the GPU (utils.py:1343-1345; UFO masks elvis.py:1186-1193)."""
    import cv2
    from elvis_b200 import elvis as E, utils as U
    rng = np.random.default_rng(41)
    frames = [rng.integers(0, 256, (48, 80, 3), dtype=np.uint8) for _ in range(2)]
    odd = rng.integers(0, 5, (2, 7, 9))                     # the block grid is 3 x 5
    fixed = np.stack([cv2.resize(m.astype(np.float32), (5, 3), interpolation=cv2.INTER_NEAREST).astype(np.int32) for m in odd])
    got = U.restore_with_opencv_unsharp(frames, odd, 16)
    ref = P.restore_with_opencv_unsharp(frames, fixed, 16, 0, 0.0)
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    # calculate_removability_scores with UFO masks on disk
    T, H, W, bs = 4, 64, 96, 16
    y, u, v = synth_yuv420(T, H, W, seed=5)
    raw = tmp_path / "reference_raw.yuv"
    with open(raw, "wb") as f:
        for t in range(T):
            f.write(y[t].tobytes() + u[t].tobytes() + v[t].tobytes())
    mdir = tmp_path / "maps" / "ufo_masks"
    mdir.mkdir(parents=True)
    bgs = []
    for t in range(T):
        m = (rng.random((H, W)) > 0.5).astype(np.uint8) * 255
        cv2.imwrite(str(mdir / f"{t + 1:05d}.png"), m)
        bgs.append(cv2.resize(m, (W // bs, H // bs), interpolation=cv2.INTER_NEAREST) == 0)
    for dct, n in ((None, bs), (8, 8), (16, 16)):       # default: the reference's `evca.main -b block_size` call
        got = E.calculate_removability_scores(str(raw), "", W, H, bs, alpha=0.4, working_dir=str(tmp_path), smoothing_beta=0.5, dct_size=dct)
        rsc, rtc = spec_scoring.sc_tc(y, bs, n)
        np.testing.assert_allclose(got, P.combine_removability(rsc, rtc, 0.4, 0.5, np.stack(bgs)), rtol=RTOL, atol=1e-9)


def test_unsharp_restorer_planar(dev):
    from elvis_b200 import ops
    y, u, _ = synth_yuv420(3, 64, 96, seed=12)
    rng = np.random.default_rng(3)
    lv = rng.integers(0, 5, (3, 4, 6)).astype(np.int32)
    out_y = ops.restore_unsharp(to_dev(y, dev), to_dev(lv, dev), 16).cpu().numpy()
    out_u = ops.restore_unsharp(to_dev(u, dev), to_dev(lv, dev), 8, halo=2).cpu().numpy()
    for t in range(3):
        assert np.array_equal(out_y[t], P.unsharp_plane(y[t], lv[t], 16))
        assert np.array_equal(out_u[t], P.unsharp_plane(u[t], lv[t], 8, halo=2))


@pytest.mark.parametrize("bs", [8, 16, 32])
def test_lanczos_restorer_elvis(dev, bs):
    from elvis_b200 import elvis as E
    rng = np.random.default_rng(bs + 200)
    img = rng.integers(0, 256, (bs * 3, bs * 5, 3), dtype=np.uint8)
    maps = rng.integers(0, 7, (3, 5))
    assert np.array_equal(E.restore_downsample_opencv_lanczos(img, maps, bs), P.restore_downsample_opencv_lanczos(img, maps, bs))
    zero = np.zeros((3, 5), int)
    assert E.restore_downsample_opencv_lanczos(img, zero, bs) is img


# ---------------------------------------------------------------------------------- round 2
def test_analyze_frames_rgb_and_luma(dev):
    """presley.analyze_frames (the stand-in for evca.analyze_frames, presley.py:202): RGB input goes
    through the cv2-exact luma kernel, luma input is used as is; SC/TC against the spec."""
    from elvis_b200 import presley
    from oracle import spec_cv
    rng = np.random.default_rng(21)
    rgb = np.clip(rng.normal(128, 50, (5, 64, 96, 3)), 0, 255).astype(np.uint8)
    rgb[1:] = np.roll(rgb[:-1], 3, axis=2) // 2 + rgb[1:] // 2
    luma = spec_cv.rgb_to_gray(rgb)
    got_luma = presley.rgb_to_luma(to_dev(rgb, dev)).cpu().numpy()
    assert np.array_equal(got_luma, luma)
    for bs, dct in ((8, None), (16, None), (16, 8), (32, None)):
        cx = presley.analyze_frames(list(rgb), presley.EVCAConfig(block_size=bs, dct_size=dct))
        rsc, rtc = spec_scoring.sc_tc(luma, bs, bs if dct is None else dct)
        assert cx.SC.dtype == np.float64 and cx.SC.shape == rsc.shape
        np.testing.assert_allclose(cx.SC, rsc, rtol=RTOL, atol=0)
        np.testing.assert_allclose(cx.TC, rtc, rtol=RTOL, atol=0)
        cy = presley.analyze_frames(luma, presley.EVCAConfig(block_size=bs, dct_size=dct))
        assert np.array_equal(cx.SC, cy.SC) and np.array_equal(cx.TC, cy.TC)
    # odd widths take the byte path of the luma kernel; strided frames
    odd = rng.integers(0, 256, (2, 9, 13, 3), dtype=np.uint8)
    assert np.array_equal(presley.rgb_to_luma(to_dev(odd, dev)).cpu().numpy(), spec_cv.rgb_to_gray(odd))
    wide = to_dev(rng.integers(0, 256, (2, 10, 40, 3), dtype=np.uint8), dev)
    view = wide[:, 1:9, 4:36]
    from elvis_b200 import ops
    assert np.array_equal(ops.rgb_to_gray(view).cpu().numpy(), spec_cv.rgb_to_gray(view.cpu().numpy()))


@pytest.mark.parametrize("by,bx,bs,amount", [(5, 7, 8, 0.25), (5, 7, 8, 0.33), (5, 7, 8, 0.5), (4, 6, 4, 0.5), (3, 9, 8, 0.9),
                                             (6, 5, 8, 0.0), (2, 2, 8, 0.99), (9, 20, 16, 0.37)])
def test_presley_batch_shrink_stretch(dev, by, bx, bs, amount):
    """presley.shrink_video_frames / stretch_video_frames (presley.py:761-827): the stretch refills the
    kept positions in row-major order over the frame, which differs from the per-row refill after a
    partial last pass."""
    from elvis_b200 import presley
    rng = np.random.default_rng(by * 100 + bx)
    frames = [rng.integers(0, 256, (by * bs + 3, bx * bs + 2, 3), dtype=np.uint8) for _ in range(3)]
    imps = [np.round(rng.random((by, bx)) * 8) / 8 for _ in range(3)]
    small, masks = presley.shrink_video_frames(frames, imps, bs, amount, presley.shrink_frame_row_only)
    for f, i, s, m in zip(frames, imps, small, masks):
        rs, rm = P.shrink_frame_row_only(f, i, bs, amount)
        assert np.array_equal(s, rs) and np.array_equal(m, rm)
    full = presley.stretch_video_frames(small, masks, bs)
    ref = P.stretch_video_frames(small, masks, bs)
    assert all(np.array_equal(a, b) for a, b in zip(full, ref))
    # a shrunk frame smaller than the mask implies (the reference's bounds guard), and masks that keep more
    # blocks than the shrunk frame holds
    cut = [s[:, :max(bs, s.shape[1] - bs)] for s in small]
    assert all(np.array_equal(a, b) for a, b in zip(presley.stretch_video_frames(cut, masks, bs), P.stretch_video_frames(cut, masks, bs)))


def test_resize_linear_float_vs_spec(dev):
    from elvis_b200 import ops
    from oracle import spec_cv
    rng = np.random.default_rng(5)
    for dt in (np.float32, np.float64):
        for sh, sw, dh, dw in ((9, 16, 27, 33), (5, 7, 8, 9), (34, 60, 40, 70), (135, 240, 68, 120), (12, 16, 3, 4), (2, 2, 5, 1)):
            a = (rng.random((3, sh, sw)) * 2 - 0.5).astype(dt)
            got = ops.resize_linear_float(to_dev(a, dev), dh, dw).cpu().numpy()
            assert got.dtype == dt
            for t in range(3):
                assert np.array_equal(got[t], spec_cv.resize_linear_float(a[t], dh, dw)), (dt.__name__, sh, sw, dh, dw)
    with pytest.raises(NotImplementedError):
        ops.resize_linear_float(to_dev(np.zeros((1, 1, 5)), dev), 3, 3)


def test_two_devices_in_one_process():
    """Opt-in kernel attributes (the tcgen05 kernel's dynamic shared memory) are per device, and every
    operator must launch on the device and stream of its tensors, not of the current device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import os
    from elvis_b200 import ops
    from elvis_b200.pipeline import ElvisV1, Yuv420
    y, u, v = synth_yuv420(6, 128, 512, seed=9)
    rsc, rtc = spec_scoring.sc_tc(y, 16)
    os.environ["ELVIS_SCORE_IMPL"] = "umma"
    try:
        for d in (0, 1, 0):
            dev = torch.device("cuda", d)
            assert torch.cuda.current_device() == 0
            sc, tc, _ = ops.score_sc_tc(torch.from_numpy(y).to(dev), 16)
            assert sc.device == dev
            np.testing.assert_allclose(sc.cpu().numpy(), rsc, rtol=RTOL, atol=0)
            np.testing.assert_allclose(tc.cpu().numpy(), rtc, rtol=RTOL, atol=0)
    finally:
        del os.environ["ELVIS_SCORE_IMPL"]
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        clip = Yuv420(*(torch.from_numpy(p).to(dev) for p in (y, u, v)))
        scores, mask, shrunk, full = ElvisV1(16, 0.5, 0.5, 0.5).run(clip)
        torch.cuda.synchronize(dev)
        outs.append([t.cpu().numpy() for t in (scores, mask, shrunk.y, shrunk.u, full.y, full.v)])
    assert all(np.array_equal(a, b) for a, b in zip(*outs))


def test_pack_levels_2bit_saturates(dev):
    from elvis_b200 import ops
    lv = np.array([[[4, -1, 3, 7, 2, 0, 1]]], np.int32)
    packed = ops.pack_levels_2bit(to_dev(lv, dev))
    assert np.array_equal(packed.cpu().numpy(), P.pack_levels_2bit(lv))
    assert ops.unpack_levels_2bit(packed, 7).cpu().numpy().tolist() == [[[3, 0, 3, 3, 2, 0, 1]]]


@pytest.mark.parametrize("bs,H,W,T", [(16, 64, 96, 7), (32, 64, 128, 4), (16, 48, 272, 30), (16, 1080, 1920, 3), (32, 96, 528, 13),
                                      (16, 50, 75, 5)])
@pytest.mark.parametrize("impl", ["tensor", "simt"])
def test_sc_tc_dct_size_equals_block_size(dev, monkeypatch, bs, H, W, T, impl):
    """dct_size = block_size (one 16 x 16 / 32 x 32 transform per block: the size the reference passes to
    EVCA, elvis.py:1022-1023) against the spec with n = block_size; chunked runs and halos give the
    same bits; unaligned planes take the byte-load path.  16 x 16 runs on the tensor cores (score_dct16.cu:
    partial groups of blocks, several groups per row) and, forced, on the CUDA cores (score_dctn.cu)."""
    from elvis_b200 import ops
    if impl == "simt":
        if bs != 16:
            pytest.skip("32 x 32 has one implementation")
        monkeypatch.setenv("ELVIS_SCORE_DCT16", "simt")
    y = synth_luma(T + 1, H, W, seed=bs + T)
    yd = to_dev(y, dev)
    rsc, rtc = spec_scoring.sc_tc(y[1:], bs, bs)
    sc, tc, mm = ops.score_sc_tc(yd[1:], bs, dct_size=bs)
    sc_n, tc_n, mm = sc.cpu().numpy(), tc.cpu().numpy(), mm.cpu().numpy()
    np.testing.assert_allclose(sc_n, rsc, rtol=RTOL, atol=1e-6 * rsc.max())
    np.testing.assert_allclose(tc_n, rtc, rtol=RTOL, atol=1e-6 * rtc.max())
    assert np.all(tc_n[0] == 0) and mm.tolist() == [sc_n.min(), sc_n.max(), tc_n.min(), tc_n.max()]
    import torch
    monkeypatch.setenv("ELVIS_SCORE_CHUNK", "3")
    sc2, tc2, _ = ops.score_sc_tc(yd[1:], bs, dct_size=bs)
    assert torch.equal(sc2, sc) and torch.equal(tc2, tc)
    hsc, htc = spec_scoring.sc_tc(y[1:], bs, bs, prev=y[0])
    sc3, tc3, _ = ops.score_sc_tc(yd[1:], bs, prev_halo=yd[0], dct_size=bs)
    assert torch.equal(sc3, sc)
    np.testing.assert_allclose(tc3.cpu().numpy(), htc, rtol=RTOL, atol=1e-6 * htc.max())
    with pytest.raises(Exception):
        ops.score_sc_tc(yd, 32 if bs == 16 else 16, dct_size=bs)          # dct_size must be 8 or the block size


def test_pipeline_with_block_sized_transform(dev):
    from elvis_b200.pipeline import ElvisV1, Yuv420
    T, H, W, bs = 4, 64, 160, 16
    y, u, v = synth_yuv420(T, H, W, seed=31)
    clip = Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev))
    scores, mask, shrunk, full = ElvisV1(bs, 0.5, 0.4, 0.5, dct_size=bs).run(clip)
    rsc, rtc = spec_scoring.sc_tc(y, bs, bs)
    scores, mask = scores.cpu().numpy(), mask.cpu().numpy()
    np.testing.assert_allclose(scores, P.combine_removability(rsc, rtc, 0.4, 0.5), rtol=RTOL, atol=1e-9)
    assert np.array_equal(mask, P.select_rows(scores, W // bs // 2, P.REMOVE_HIGH))
    for t in range(T):
        assert np.array_equal(full.y[t].cpu().numpy(), P.stretch_plane(P.shrink_plane(y[t], mask[t], bs), mask[t], bs))


def test_host_server_and_client_legs(dev):
    """HostElvisV1(outputs=("mask", "shrunk"), pack_masks=True) is the reference's server (elvis.py:4389-4418: shrunk
    frames + np.packbits mask side channel) and HostElvisClient its client (elvis.py:4537-4557): together they must
    reproduce the composite path, and the packed masks must be np.packbits' bytes."""
    import torch
    from elvis_b200.pipeline import ElvisV1, HostElvisClient, HostElvisV1, Yuv420
    T, H, W, bs = 4, 64, 160, 16
    y, u, v = synth_yuv420(T, H, W, seed=52)
    i420 = torch.from_numpy(np.concatenate([y.reshape(T, -1), u.reshape(T, -1), v.reshape(T, -1)], axis=1)).pin_memory()
    ref = ElvisV1(bs, 0.5, 0.5, 0.5).run(Yuv420(to_dev(y, dev), to_dev(u, dev), to_dev(v, dev)))
    ref_mask = ref[1].cpu().numpy()
    server = HostElvisV1(T, H, W, bs, 0.5, 0.5, 0.5, dev, depth=2, outputs=("mask", "shrunk"), pack_masks=True)
    shrunk_h, full_h, bits_h = server.host_buffers()
    assert full_h is None and bits_h.numel() == (T * (H // bs) * (W // bs) + 7) // 8
    assert server.d2h_bytes == shrunk_h.numel() + bits_h.numel() and "full" not in server.slots[0]
    for _ in range(3):                     # the slots are reused
        server.process(i420, shrunk_h, None, bits_h).synchronize()
    assert np.array_equal(bits_h.numpy(), np.packbits(ref_mask))
    sw = ref[2].y.shape[2]
    for a, b in zip(Yuv420.from_i420(shrunk_h, H, sw).planes, ref[2].planes):
        assert torch.equal(a, b.cpu())
    client = HostElvisClient(T, H, W, bs, 0.5, dev, depth=2)
    out = torch.empty((T, H * W * 3 // 2), dtype=torch.uint8).pin_memory()
    for _ in range(3):
        client.process(shrunk_h, bits_h, out).synchronize()
    for a, b in zip(Yuv420.from_i420(out, H, W).planes, ref[3].planes):
        assert torch.equal(a, b.cpu())
    assert client.h2d_bytes == shrunk_h.numel() + bits_h.numel() and client.d2h_bytes == out.numel()
    # unpacked masks, stretched only
    only = HostElvisV1(T, H, W, bs, 0.5, 0.5, 0.5, dev, depth=1, outputs=("stretched",))
    _, full2, _ = only.host_buffers()
    only.process(i420, None, full2, None).synchronize()
    assert torch.equal(full2, out)
    with pytest.raises(ValueError):
        HostElvisV1(T, H, W, bs, outputs=("scores",))
