"""Pins the oracle (CPU only): spec_cv against the real cv2 of the image, ref_port against
the unmodified reference modules when the reference tree is present (build container), the
coefficient tables against spec_cv, and domain properties with hypothesis."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import ref_import
from oracle import ref_port as P
from oracle import spec_cv, spec_dct_dampen, spec_scoring

cv2 = pytest.importorskip("cv2")
needs_reference = pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")


def _blocks(rng, bs, n):
    out = []
    for i in range(n):
        k = i % 3
        if k == 0:
            out.append(rng.integers(0, 256, (bs, bs), dtype=np.uint8))
        elif k == 1:
            out.append((rng.integers(0, 2, (bs, bs)) * 255).astype(np.uint8))
        else:
            out.append(np.clip(rng.normal(128, 20, (bs, bs)), 0, 255).astype(np.uint8))
    return out


@pytest.mark.parametrize("bs", [4, 8, 16, 32, 12, 20])
def test_spec_cv_bit_exact_vs_cv2(bs):
    rng = np.random.default_rng(bs)
    for b in _blocks(rng, bs, 24):
        ref = mine = b
        for _ in range(3):
            ref = cv2.GaussianBlur(ref, (5, 5), sigmaX=1.0)
            mine = spec_cv.gaussian_blur5(mine)
            assert np.array_equal(ref, mine)
        for small in sorted({max(1, bs // k) for k in (1, 2, 3, 4, 5, 8, 16)}):
            ra = cv2.resize(b, (small, small), interpolation=cv2.INTER_AREA)
            assert np.array_equal(ra, spec_cv.resize_area(b, small)), (bs, small)
            ru = cv2.resize(ra, (bs, bs), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(ru, spec_cv.resize_linear(ra, bs)), (bs, small)


def test_spec_cv_channels_are_independent():
    rng = np.random.default_rng(0)
    blk = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    r3 = cv2.GaussianBlur(blk, (5, 5), sigmaX=1.0)
    assert np.array_equal(r3, np.stack([spec_cv.gaussian_blur5(blk[..., c]) for c in range(3)], -1))
    for small in (8, 5, 4, 1):
        r3 = cv2.resize(cv2.resize(blk, (small, small), interpolation=cv2.INTER_AREA), (16, 16), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(r3, np.stack([spec_cv.down_up(blk[..., c], small) for c in range(3)], -1))


def test_tables_match_spec_cv():
    from elvis_b200 import _tables as T
    for pb in (4, 8, 16, 32, 12, 20):
        for small in range(1, pb):
            for h in (True, False):
                assert all(np.array_equal(a, b) for a, b in zip(T._linear_taps(small, pb, h), spec_cv.linear_coeffs(small, pb, h)))
            if pb % small:
                start, src, alpha = T._area_entries(pb, small)
                tab = spec_cv.area_table(pb, small)
                assert [t[0] for t in tab] == src.tolist()
                assert np.array_equal(np.array([t[2] for t in tab], np.float32), alpha)
                assert [t[1] for t in tab] == [d for d in range(small) for _ in range(start[d + 1] - start[d])]


@needs_reference
def test_port_matches_reference_v1():
    E, U = ref_import.load("elvis"), ref_import.load("utils")
    rng = np.random.default_rng(1)
    for (H, W, bs, sh) in [(64, 96, 16, 0.5), (48, 80, 8, 0.25), (32, 64, 16, 3), (32, 64, 16, 0.0), (32, 64, 16, 0.999)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        sc = rng.random((H // bs, W // bs))
        r, m = E.apply_selective_removal(img, sc, bs, sh), P.apply_selective_removal(img, sc, bs, sh)
        assert np.array_equal(r[0], m[0]) and np.array_equal(r[1], m[1]) and r[1].dtype == m[1].dtype and r[2] == m[2]
        assert np.array_equal(E.stretch_frame(r[0], r[1], bs), P.stretch_frame(m[0], m[1], bs))
    for (H, W, bs, sh) in [(64, 96, 16, 0.5), (50, 85, 8, 0.25), (80, 128, 16, 0.3), (80, 128, 16, 0.0), (80, 128, 16, 0.99), (40, 64, 8, 0.3)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        imp = np.round(rng.random((H // bs, W // bs)) * 8) / 8
        r, m = U.shrink_frame_row_only(img, imp, bs, sh), P.shrink_frame_row_only(img, imp, bs, sh)
        assert np.array_equal(r[1], m[1]) and np.array_equal(r[0], m[0])
        assert np.array_equal(U.stretch_frame_row_only(r[0], r[1], bs), P.stretch_frame_row_only(m[0], m[1], bs))


@needs_reference
def test_port_matches_reference_rowcol():
    """8f rank 2: utils.py:763-1018, incl. ties, partial passes, crop and 1-wide grids."""
    U = ref_import.load("utils")
    rng = np.random.default_rng(7)
    for (H, W, bs, sh) in [(64, 96, 16, 0.5), (51, 85, 8, 0.3), (80, 128, 16, 0.9), (80, 128, 16, 0.0), (40, 64, 8, 0.62),
                           (16, 128, 16, 0.5), (128, 16, 16, 0.5), (48, 48, 8, 1.0)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        imp = np.round(rng.random((H // bs, W // bs)) * 8) / 8
        r, m = U.shrink_frame_position_map(img, imp, bs, sh), P.shrink_frame_position_map(img, imp, bs, sh)
        assert all(np.array_equal(a, b) for a, b in zip(r, m)), (H, W, bs, sh)
        assert np.array_equal(U.stretch_frame_position_map(*r, bs), P.stretch_frame_position_map(*m, bs))
        r, m = U.shrink_frame_removal_indices(img, imp, bs, sh), P.shrink_frame_removal_indices(img, imp, bs, sh)
        assert np.array_equal(r[0], m[0]) and np.array_equal(r[1], m[1]) and len(r[2]) == len(m[2])
        assert all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(r[2], m[2]))
        by, bx = H // bs, W // bs
        assert np.array_equal(U.stretch_frame_removal_indices(r[0], r[2], by, bx, bs),
                              P.stretch_frame_removal_indices(m[0], m[2], by, bx, bs))
        # malformed side info: short / out-of-range records are clamped the same way
        bad = [a[:max(1, len(a) - 2)] + 3 for a in r[2]]
        assert np.array_equal(U.stretch_frame_removal_indices(r[0], bad, by, bx, bs),
                              P.stretch_frame_removal_indices(m[0], bad, by, bx, bs))


def test_spec_cv_float_area_and_i420_vs_cv2():
    """8f rank 3 arithmetic: float32 INTER_AREA (general, integer-ratio and 2x2 vector paths) and
    RGB -> I420, bit-exact against the cv2 of this image."""
    rng = np.random.default_rng(11)
    for (sh, sw, dh, dw) in [(135, 240, 34, 60), (136, 240, 34, 60), (67, 120, 17, 30), (8, 12, 4, 3), (9, 13, 4, 5),
                             (135, 240, 68, 120), (16, 16, 8, 8), (34, 60, 17, 30), (10, 34, 5, 17), (24, 36, 8, 12), (7, 9, 7, 9)]:
        a = (rng.random((sh, sw)) * 2 - 1).astype(np.float32)
        assert np.array_equal(cv2.resize(a, (dw, dh), interpolation=cv2.INTER_AREA), spec_cv.resize_area_f32(a, dh, dw)), (sh, sw, dh, dw)
    for (h, w) in [(6, 8), (64, 96), (2, 2), (34, 50)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        img[0, 0], img[0, 1] = (255, 255, 255), (0, 0, 0)
        assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_RGB2YUV_I420), spec_cv.rgb_to_i420(img))


def test_spec_cv_float_linear_and_gray_vs_cv2():
    """cv2.resize(float map, INTER_LINEAR) -- fused lerp, fma-contracted coordinate on the float64 path
    only -- and COLOR_RGB2GRAY's 15-bit fixed point, restated bit for bit."""
    from elvis_b200 import _tables as T
    rng = np.random.default_rng(13)
    for dt in (np.float32, np.float64):
        shapes = [tuple(int(v) for v in rng.integers(2, 40, 4)) for _ in range(40)]
        shapes += [(9, 16, 27, 33), (10, 16, 5, 8), (12, 16, 3, 4), (135, 240, 68, 120), (17, 30, 34, 60), (3, 5, 6, 10)]
        for sh, sw, dh, dw in shapes:
            a = (rng.random((sh, sw)) * 2 - 0.5).astype(dt)
            ref = cv2.resize(a, (dw, dh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(ref, spec_cv.resize_linear_float(a, dh, dw)), (dt.__name__, sh, sw, dh, dw)
            fused = dt == np.float64
            idx, frac = T.linear_float_index(sw, dw, fused)
            cs = spec_cv.linear_float_coords(sw, dw, fused)
            assert idx.tolist() == [c[0] for c in cs] and frac.tolist() == [c[1] for c in cs]
    for w in (1, 3, 7, 16, 33, 257):
        a = rng.integers(0, 256, (5, w, 3), dtype=np.uint8)
        assert np.array_equal(cv2.cvtColor(a, cv2.COLOR_RGB2GRAY), spec_cv.rgb_to_gray(a))


@needs_reference
def test_port_matches_reference_presley_stretch_and_mismatched_maps(tmp_path):
    """presley.py's batch wrappers (compiled from the unmodified file) incl. partial last passes, where
    its row-major refill differs from utils.stretch_frame_row_only; utils degradations with an
    importance grid that differs from the frame's; the qpfile's INTER_LINEAR branch."""
    from _ref_drive import reference_qpfile
    E, U = ref_import.load("elvis"), ref_import.load("utils")
    PF = ref_import.load_presley_functions("shrink_frame_row_only", "shrink_video_frames", "stretch_video_frames")
    rng = np.random.default_rng(3)
    differs = 0
    for (by, bx, bs, sh) in ((5, 7, 8, 0.25), (5, 7, 8, 0.33), (5, 7, 8, 0.5), (4, 6, 4, 0.5), (3, 9, 8, 0.9), (6, 5, 8, 0.0), (2, 2, 8, 0.99)):
        frames = [rng.integers(0, 256, (by * bs + 3, bx * bs + 2, 3), dtype=np.uint8) for _ in range(2)]
        imps = [np.round(rng.random((by, bx)) * 8) / 8 for _ in range(2)]
        small, masks = PF["shrink_video_frames"](frames, imps, bs, sh, PF["shrink_frame_row_only"])
        ref = PF["stretch_video_frames"](small, masks, bs)
        mine = P.stretch_video_frames(small, masks, bs)
        assert all(np.array_equal(a, b) for a, b in zip(ref, mine)), (by, bx, bs, sh)
        differs += any(not np.array_equal(a, P.stretch_frame_row_only(s, m, bs)) for a, s, m in zip(ref, small, masks))
    assert differs >= 3       # the partial-pass cases are really exercised
    frame = rng.integers(0, 256, (16 * 3 + 3, 16 * 4 + 5, 3), dtype=np.uint8)
    for shape in ((2, 3), (7, 9), (3, 9), (5, 4)):
        imp = rng.random(shape)
        for name in ("degrade_adaptive_downsample", "degrade_adaptive_blur"):
            r, m = getattr(U, name)(frame, imp, 16), getattr(P, name)(frame, imp, 16)
            assert np.array_equal(r[0], m[0]) and np.array_equal(r[1], m[1]), (name, shape)
    qs = rng.random((2, 3, 5))
    P.write_per_block_qpfile(qs, 128, 640, 384, str(tmp_path / "q.txt"))
    assert open(tmp_path / "q.txt").read() == reference_qpfile(E, qs, 128, 640, 384, str(tmp_path / "ref"))


@needs_reference
def test_port_matches_reference_roi_files(tmp_path):
    """8f rank 3: the side files byte for byte (utils.py:453-462, 1026-1092; elvis.py:2027-2090)."""
    E, U = ref_import.load("elvis"), ref_import.load("utils")
    rng = np.random.default_rng(12)
    imps = [rng.random((17, 30)) for _ in range(3)]
    imps[1][0, :4] = [0.0, 1.0, 0.5, 0.125]
    for base_qp, rng_qp in ((48, 15), (3, 15), (30, 6)):
        U.create_kvazaar_roi_file(imps, str(tmp_path / "a.bin"), base_qp, rng_qp)
        P.create_kvazaar_roi_file(imps, str(tmp_path / "b.bin"), base_qp, rng_qp)
        assert (tmp_path / "a.bin").read_bytes() == (tmp_path / "b.bin").read_bytes()
    for (w, h, crf, rq) in ((480, 272, 35, 10), (480, 272, 60, 15), (1920, 1088, 2, 7)):
        maps = [rng.random((h // 16, w // 16)) for _ in range(2)]
        U.create_svtav1_roi_file(maps, str(tmp_path / "a.txt"), crf, rq, w, h)
        P.create_svtav1_roi_file(maps, str(tmp_path / "b.txt"), crf, rq, w, h)
        assert (tmp_path / "a.txt").read_text() == (tmp_path / "b.txt").read_text()
    frames = [rng.integers(0, 256, (32, 48, 3), dtype=np.uint8) for _ in range(2)]
    U.write_y4m(frames, str(tmp_path / "a.y4m"), 29.97)
    P.write_y4m(frames, str(tmp_path / "b.y4m"), 29.97)
    assert (tmp_path / "a.y4m").read_bytes() == (tmp_path / "b.y4m").read_bytes()
    # the qpfile of encode_with_roi: capture it where the reference hands it to the encoder
    from _ref_drive import reference_qpfile
    for (w, h, bs) in ((480, 272, 16), (3840, 2160, 16), (256, 128, 32), (4352, 2304, 32)):   # same grid, general, same, 2x2
        scores = rng.random((2, h // bs, w // bs))
        scores[0, 0, :3] = [0.0, 1.0, 0.5]
        P.write_per_block_qpfile(scores, bs, w, h, str(tmp_path / "q.txt"))
        assert reference_qpfile(E, scores, bs, w, h, str(tmp_path / "ref")) == (tmp_path / "q.txt").read_text(), (w, h, bs)


def _up2_cubic(im):
    return cv2.resize(im, None, fx=2, fy=2, interpolation=cv2.INTER_CUBIC)


def _up2_repeat(im):
    return np.ascontiguousarray(im.repeat(2, 0).repeat(2, 1))


@needs_reference
def test_port_matches_reference_pyramid_and_map_video(tmp_path):
    """8f rank 4: upscale_realesrgan_adaptive with a deterministic 2x upsampler standing in for the
    external one (elvis.py:2522-2600), and the arithmetic of the map video (elvis.py:2198-2245)."""
    from _ref_drive import reference_map_video_arithmetic
    E = ref_import.load("elvis")
    rng = np.random.default_rng(13)
    for (H, W, bs, top) in [(64, 96, 16, 4), (64, 96, 16, 2), (48, 80, 8, 3), (64, 64, 32, 5), (32, 48, 16, 0)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        maps = rng.integers(0, top + 1, (H // bs, W // bs))
        maps.flat[0] = top
        for up in (_up2_cubic, _up2_repeat):
            ref = E.upscale_realesrgan_adaptive(img, maps.copy(), bs, upsample_fn=up)
            assert np.array_equal(ref, P.upscale_realesrgan_adaptive(img, maps.copy(), bs, up)), (H, W, bs, top)
    for kind, bs, hi in (("gaussian", 16, 10), ("downsample", 16, 4), ("downsample", 8, 3)):
        m = rng.integers(0, hi + 1, (3, 17, 30)).astype(np.int32)
        m.flat[:2] = [0, hi]
        gray, decoded = reference_map_video_arithmetic(E, m, kind, bs, str(tmp_path / f"{kind}{bs}"))
        assert np.array_equal(gray, P.strength_maps_to_gray(m))
        assert np.array_equal(decoded, P.gray_to_strength_maps(gray, 0.0, 10.0 if kind == "gaussian" else int(np.log2(bs))))
        assert np.array_equal(decoded, m)          # lossless when the video codec is


@needs_reference
def test_port_matches_reference_v2_and_scores():
    from _ref_drive import run_reference_removability
    E, U = ref_import.load("elvis"), ref_import.load("utils")
    rng = np.random.default_rng(2)
    for bs in (8, 16):
        img = rng.integers(0, 256, (bs * 4, bs * 5, 3), dtype=np.uint8)
        sc = rng.random((4, 5))
        sc[0, :4] = [0.5, 0.125, 1.0, 0.0]
        for fn in ("filter_frame_downsample", "filter_frame_gaussian"):
            r, m = getattr(E, fn)(img, sc, bs), getattr(P, fn)(img, sc, bs)
            assert np.array_equal(r[0], m[0]) and np.array_equal(r[1], m[1]) and r[1].dtype == m[1].dtype
        img2 = rng.integers(0, 256, (bs * 4 + 3, bs * 5 + 5, 3), dtype=np.uint8)
        for fn in ("degrade_adaptive_downsample", "degrade_adaptive_blur"):
            r, m = getattr(U, fn)(img2, sc, bs), getattr(P, fn)(img2, sc, bs)
            assert np.array_equal(r[0], m[0]) and np.array_equal(r[1], m[1])
    T, By, Bx = 6, 5, 9
    s, t = rng.random((T, By, Bx)) * 60, rng.random((T, By, Bx)) * 30
    t[0] = 0
    fg = (rng.random((T, By, Bx)) > 0.4).astype(np.uint8) * 255
    for beta in (1, 0.5):
        for m in (None, fg):
            ref = run_reference_removability(s, t, m, 0.3, beta, 16)
            assert np.array_equal(ref, P.combine_removability(s, t, 0.3, beta, None if m is None else m == 0))

    class Cx:
        pass
    cx = Cx()
    cx.SC, cx.TC = s, t
    f = rng.random((T, By, Bx))
    assert np.array_equal(np.stack(U.calculate_importance_scores(None, 16, 0.3, 0.6, cx, f)), P.importance_scores(s, t, 0.3, 0.6, f))


def test_sigma_gaussian_and_unsharp_bit_exact_vs_cv2():
    rng = np.random.default_rng(5)
    for level in range(1, 11):
        for shape in ((16, 16), (8, 8), (24, 20), (16, 27)):
            t = rng.integers(0, 256, shape, dtype=np.uint8)
            bl = cv2.GaussianBlur(t, (0, 0), max(1, level))
            assert np.array_equal(bl, spec_cv.gaussian_blur_sigma(t, max(1, level))), (level, shape)
            ref = np.clip(cv2.addWeighted(t, 1.0 + level * 0.5, bl, -level * 0.5, 0), 0, 255).astype(np.uint8)
            assert np.array_equal(ref, spec_cv.unsharp(t, level)), (level, shape)
    from elvis_b200 import _tables as T
    tab = T.gaussian_kernels(10)
    for level in range(1, 11):
        q = spec_cv.gaussian_kernel_q8(level)
        assert q.sum() == 256 and tab[level, 0] == len(q) and np.array_equal(tab[level, 1:1 + len(q)], q)


def test_lanczos4_bit_exact_vs_cv2():
    rng = np.random.default_rng(8)
    from elvis_b200 import _tables as T
    for bs, small in ((16, 8), (16, 4), (16, 2), (16, 1), (8, 4), (8, 2), (8, 1), (32, 8), (12, 6), (24, 3)):
        s = rng.integers(0, 256, (4, small, small), dtype=np.uint8)
        mine = spec_cv.resize_lanczos4(s, bs)
        for i in range(4):
            assert np.array_equal(cv2.resize(s[i], (bs, bs), interpolation=cv2.INTER_LANCZOS4), mine[i]), (bs, small)
        a, b = T._lanczos4_taps(small, bs), spec_cv.lanczos4_taps(small, bs)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@needs_reference
def test_port_matches_reference_restorers():
    E, U = ref_import.load("elvis"), ref_import.load("utils")
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (48, 80, 3), dtype=np.uint8)
    maps = rng.integers(0, 5, (3, 5))
    assert np.array_equal(E.restore_blur_opencv_unsharp_mask(img, maps, 16), P.restore_blur_opencv_unsharp_mask(img, maps, 16))
    for bs in (8, 16):
        im = rng.integers(0, 256, (bs * 3, bs * 5, 3), dtype=np.uint8)
        lm = rng.integers(0, 6, (3, 5))
        assert np.array_equal(E.restore_downsample_opencv_lanczos(im, lm, bs), P.restore_downsample_opencv_lanczos(im, lm, bs))
    frames = [rng.integers(0, 256, (53, 85, 3), dtype=np.uint8) for _ in range(3)]
    dm = rng.integers(0, 5, (3, 3, 5))
    for halo, tb in ((0, 0.0), (8, 0.1), (20, 0.25)):
        ref = U.restore_with_opencv_unsharp(frames, dm, 16, halo=halo, temporal_blend=tb)
        ref2 = U.restore_with_opencv_lanczos(frames, dm, 16, halo=halo, temporal_blend=tb)
        mine = P.restore_with_opencv_unsharp(frames, dm, 16, halo, tb)
        assert all(np.array_equal(a, b) for a, b in zip(ref, mine)) and all(np.array_equal(a, b) for a, b in zip(ref2, mine))


# ---------------------------------------------------------------------------- properties
@settings(max_examples=40, deadline=None)
@given(st.integers(1, 6), st.integers(2, 12), st.sampled_from([4, 8, 16]), st.floats(0, 0.99), st.integers(0, 2 ** 31))
def test_shrink_stretch_round_trip(by, bx, bs, amount, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (by * bs, bx * bs, 3), dtype=np.uint8)
    scores = np.round(rng.random((by, bx)) * 5) / 5
    small, mask, coords = P.apply_selective_removal(img, scores, bs, amount)
    k = P.blocks_to_remove_elvis(amount, bx)
    assert (mask.sum(axis=1) == k).all() and small.shape == (by * bs, (bx - k) * bs, 3)
    full = P.stretch_frame(small, mask, bs)
    keep = np.kron(1 - mask, np.ones((bs, bs), np.int8)).astype(bool)
    assert np.array_equal(full[keep], img[keep]) and not full[~keep].any()
    # the removed set is the stable top-k
    for j in range(by):
        assert coords[j] == sorted(np.argsort(-scores[j], kind="stable")[:k].tolist())


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 7), st.integers(2, 12), st.floats(0, 0.99))
def test_row_only_plan_counts(by, bx, amount):
    k, out_bx = P.row_only_plan(by, bx, amount)
    target = int(by * bx * amount)
    assert k.sum() == min(target, by * (bx - 1)) and out_bx == bx - k.max() and k.max() - k.min() <= 1


def test_level_rules_and_2bit_packing():
    s = np.array([[0.0, 0.125, 0.5, 0.375, 0.625, 1.0]])
    assert P.levels_elvis_downsample(s, 16).tolist() == [[0, 0, 2, 2, 2, 4]]       # round half to even
    assert P.levels_elvis_blur(np.array([0.05, 0.15, 0.25])).tolist() == [0, 2, 2]
    assert P.levels_utils_downsample(np.array([1.0, 0.74, 0.5, 0.26, 0.0])).tolist() == [0, 2, 3, 3, 4]
    rng = np.random.default_rng(0)
    lv = rng.integers(0, 4, (2, 3, 13))
    packed = P.pack_levels_2bit(lv)
    assert packed.shape == (2, 3, 4) and np.array_equal(P.unpack_levels_2bit(packed, 13), lv)
    # saturating: the reference's level 4 (16x at bs 16) is kept as 8x in the 2-bit map; negatives as 0
    assert P.unpack_levels_2bit(P.pack_levels_2bit(np.array([[4, -1, 3, 7, 2]])), 5).tolist() == [[3, 0, 3, 3, 2]]


def test_scoring_spec_properties():
    rng = np.random.default_rng(3)
    y = rng.integers(0, 256, (4, 32, 48), dtype=np.uint8)
    y[2] = y[1]
    sc, tc = spec_scoring.sc_tc(y, 16)
    assert np.all(tc[0] == 0) and np.all(tc[2] == 0) and np.all(sc >= 0)
    flat = np.full((2, 16, 16), 77, np.uint8)
    s2, t2 = spec_scoring.sc_tc(flat, 16)
    assert np.allclose(s2, 0, atol=1e-9) and np.all(t2 == 0)      # DC carries no energy
    s3, t3 = spec_scoring.sc_tc(y[1:], 16, prev=y[0])
    assert np.allclose(t3, tc[1:]) and np.allclose(s3, sc[1:])
    p = y[0]
    assert np.array_equal(spec_dct_dampen.dampen_plane(p, np.zeros((2, 3)), 16), p)


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 6), st.integers(2, 7), st.integers(0, 2), st.integers(0, 2 ** 31 - 1))
def test_rowcol_properties(by, bx, whole_passes, seed):
    """Row+column shrink (utils.py:763-1018) when the target is met by whole passes (a partial last
    pass leaves stale blocks in the grid -- a quirk of the reference that the port reproduces and
    test_port_matches_reference_rowcol pins): the removed blocks are exactly the ones missing from
    the position map, and stretching by position map puts every survivor back and zeros the rest."""
    rng = np.random.default_rng(seed)
    bs = 4
    img = rng.integers(1, 256, (by * bs, bx * bs, 3), dtype=np.uint8)         # no zeros: zeros mark removed blocks
    imp = np.round(rng.random((by, bx)) * 4) / 4
    target = [0, by, by + (bx - 1)][whole_passes]
    amount = (target + 1e-6) / (by * bx)
    assert int(by * bx * amount) == target
    small, mask, pmap = P.shrink_frame_position_map(img, imp, bs, amount)
    assert mask.sum() == target
    assert pmap.shape[:2] == (by - (whole_passes == 2), bx - (whole_passes >= 1))
    kept = {(int(y), int(x)) for y, x in pmap.reshape(-1, 2)}
    assert len(kept) == pmap.shape[0] * pmap.shape[1] == by * bx - target
    assert all(not mask[y, x] for y, x in kept)
    full = P.stretch_frame_position_map(small, mask, pmap, bs)
    for y in range(by):
        for x in range(bx):
            blk = full[y * bs:(y + 1) * bs, x * bs:(x + 1) * bs]
            if mask[y, x]:
                assert not blk.any()
            else:
                assert np.array_equal(blk, img[y * bs:(y + 1) * bs, x * bs:(x + 1) * bs])
    small2, mask2, passes = P.shrink_frame_removal_indices(img, imp, bs, amount)
    assert np.array_equal(small2, small) and np.array_equal(mask2, mask) and sum(len(p) for p in passes) == target
    assert P.stretch_frame_removal_indices(small2, passes, by, bx, bs).shape == img.shape
