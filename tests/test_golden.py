"""Golden vectors frozen from the unmodified reference (tests/golden/gen_golden.py):
  * CPU: the oracle port reproduces them bit for bit (pins the oracle);
  * GPU (-m gpu): the CUDA path, called through the reference-signature mirrors, reproduces
    them bit for bit."""
import os

import numpy as np
import pytest

from oracle import ref_port as P

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


class _Impl:
    """Uniform view over the oracle port and the CUDA mirrors."""

    def __init__(self, kind):
        self.kind = kind
        if kind == "gpu":
            import torch
            assert torch.cuda.is_available()
            from elvis_b200 import elvis, presley, utils
            self.E, self.U, self.Pr = elvis, utils, presley

    def removability(self, sc, tc, alpha, beta, bg):
        if self.kind == "gpu":
            return self.E.removability_from_features(sc, tc, alpha, beta, bg)
        return P.combine_removability(sc, tc, alpha, beta, bg)

    def importance(self, sc, tc, alpha, beta, fg):
        if self.kind == "gpu":
            class Cx:
                pass
            cx = Cx()
            cx.SC, cx.TC = sc, tc
            return np.stack(self.U.calculate_importance_scores(None, 16, alpha, beta, cx, fg))
        return P.importance_scores(sc, tc, alpha, beta, fg)

    def __getattr__(self, name):
        if self.kind == "gpu":
            for mod in (self.E, self.U, self.Pr):
                if hasattr(mod, name):
                    return getattr(mod, name)
            raise AttributeError(name)
        return getattr(P, name)


def _check_all(impl):
    sc, tc, fg = G["a2_sc"], G["a2_tc"], G["a2_fg"]
    for tag, beta, m in (("b1", 1, None), ("b05", 0.5, None), ("b05m", 0.5, fg), ("b025m", 0.25, fg)):
        got = impl.removability(sc, tc, 0.3, beta, None if m is None else m == 0)
        assert np.array_equal(got, G[f"a2_out_{tag}"]), tag
    assert np.array_equal(impl.importance(sc, tc, 0.3, 0.6, G["a3_fg"]), G["a3_out"])

    img = G["a4_img"]
    for tag, bs, amount in (("s25", 8, 0.25), ("s50", 8, 0.5), ("n2", 8, 2)):
        small, mask, coords = impl.apply_selective_removal(img, G[f"a4_scores_{tag}"], bs, amount)
        assert small.dtype == np.uint8 and np.array_equal(small, G[f"a4_small_{tag}"]), tag
        assert mask.dtype == np.int8 and np.array_equal(mask, G[f"a4_mask_{tag}"]), tag
        assert coords == [np.flatnonzero(r).tolist() for r in G[f"a4_mask_{tag}"]]
        assert np.array_equal(impl.stretch_frame(small, mask, bs), G[f"a5_full_{tag}"]), tag

    for tag, amount in (("q30", 0.3), ("q50", 0.5)):
        small, mask = impl.shrink_frame_row_only(G["a6_img"], G[f"a6_imp_{tag}"], 8, amount)
        assert mask.dtype == bool and np.array_equal(mask, G[f"a6_mask_{tag}"]), tag
        assert np.array_equal(small, G[f"a6_small_{tag}"]), tag
        assert np.array_equal(impl.stretch_frame_row_only(small, mask, 8), G[f"a7_full_{tag}"]), tag

    for tag, amount in (("q30", 0.3), ("q55", 0.55), ("q85", 0.85)):
        small, mask, pmap = impl.shrink_frame_position_map(G["f2_img"], G[f"f2_imp_{tag}"], 8, amount)
        assert mask.dtype == bool and np.array_equal(mask, G[f"f2_mask_{tag}"]), tag
        assert np.array_equal(small, G[f"f2_small_{tag}"]) and np.array_equal(pmap, G[f"f2_pmap_{tag}"]), tag
        assert np.array_equal(impl.stretch_frame_position_map(small, mask, pmap, 8), G[f"f2_full_pm_{tag}"]), tag
        small2, mask2, passes = impl.shrink_frame_removal_indices(G["f2_img"], G[f"f2_imp_{tag}"], 8, amount)
        assert np.array_equal(small2, small) and np.array_equal(mask2, mask), tag
        assert [len(a) for a in passes] == G[f"f2_npass_{tag}"].tolist(), tag
        assert all(a.dtype == np.int32 for a in passes) and np.array_equal(np.concatenate(passes), G[f"f2_passes_{tag}"]), tag
        assert np.array_equal(impl.stretch_frame_removal_indices(small2, passes, 6, 8, 8), G[f"f2_full_ri_{tag}"]), tag

    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "f")
        imps = list(G["f3_imps"])
        impl.create_kvazaar_roi_file(imps, path, 40, 15)
        assert open(path, "rb").read() == G["f3_kvazaar"].tobytes()
        impl.create_svtav1_roi_file(imps, path, 35, 10, 480, 272)
        assert open(path, "rb").read() == G["f3_svtav1"].tobytes()
        impl.write_y4m(list(G["f3_rgb"]), path, 30.0)
        assert open(path, "rb").read() == G["f3_y4m"].tobytes()
        impl.write_per_block_qpfile(G["f3_scores"], 16, 960, 544, path)
        assert open(path, "rb").read() == G["f3_qpfile"].tobytes()
        s4k = G["f3_scores"][:1, :, :60].repeat(4, 1).repeat(4, 2)[:, :135, :240].copy()
        impl.write_per_block_qpfile(s4k, 16, 3840, 2160, path)
        assert open(path, "rb").read() == G["f3_qpfile_4k"].tobytes()

    up2 = lambda im: np.ascontiguousarray(im.repeat(2, 0).repeat(2, 1))      # noqa: E731
    if impl.kind == "gpu":
        got = impl.upscale_realesrgan_adaptive(G["f4_img"], G["f4_maps"].copy(), 16, upsample_fn=up2)
    else:
        got = impl.upscale_realesrgan_adaptive(G["f4_img"], G["f4_maps"].copy(), 16, up2)
    assert np.array_equal(got, G["f4_out"])
    gray = impl.strength_maps_to_gray(G["f4_levels"])
    assert gray.dtype == np.uint8 and np.array_equal(gray, G["f4_gray"])
    assert np.array_equal(impl.gray_to_strength_maps(G["f4_gray"], 0.0, 10.0), G["f4_decoded"])

    assert np.array_equal(impl.restore_blur_opencv_unsharp_mask(G["f1_img"], G["f1_map"], 16), G["f1_out"])
    for tag, halo, tb in (("h0", 0, 0.0), ("h6b", 6, 0.2)):
        got = impl.restore_with_opencv_unsharp(list(G["f1_frames"]), G["f1_maps"], 16, halo=halo, temporal_blend=tb)
        assert np.array_equal(np.stack(got), G[f"f1_out_{tag}"]), tag

    for bs in (8, 16):
        s = G[f"a8_scores_{bs}"]
        for fn, key in (("filter_frame_downsample", "a8"), ("filter_frame_gaussian", "a9")):
            out, lv = getattr(impl, fn)(G[f"a8_img_{bs}"], s, bs)
            assert lv.dtype == np.int32 and np.array_equal(lv, G[f"{key}_map_{bs}"]), (fn, bs)
            assert np.array_equal(out, G[f"{key}_out_{bs}"]), (fn, bs)
        for fn, key in (("degrade_adaptive_downsample", "a10"), ("degrade_adaptive_blur", "a11")):
            out, lv = getattr(impl, fn)(G[f"a10_img_{bs}"], s, bs)
            assert np.array_equal(lv, G[f"{key}_map_{bs}"]), (fn, bs)
            assert np.array_equal(out, G[f"{key}_out_{bs}"]), (fn, bs)


def _check_round2(impl):
    """Vectors added in round 2: presley batch stretch (row-major refill, partial passes), utils
    degradations with a mismatched importance grid, the qpfile's INTER_LINEAR branch."""
    frames = list(G["p7_frames"])
    for tag, amount in (("q25", 0.25), ("q33", 0.33), ("q50", 0.5), ("q60", 0.6)):
        small, masks = list(G[f"p7_small_{tag}"]), list(G[f"p7_mask_{tag}"])
        if impl.kind == "gpu":
            got_small, got_masks = impl.shrink_video_frames(frames, list(G[f"p7_imp_{tag}"]), 8, amount, impl.Pr.shrink_frame_row_only)
            assert np.array_equal(np.stack(got_small), G[f"p7_small_{tag}"]), tag
            assert got_masks[0].dtype == bool and np.array_equal(np.stack(got_masks), G[f"p7_mask_{tag}"]), tag
        assert np.array_equal(np.stack(impl.stretch_video_frames(small, masks, 8)), G[f"p7_full_{tag}"]), tag
    for tag in ("up", "down"):
        for fn, key in (("degrade_adaptive_downsample", "a10m"), ("degrade_adaptive_blur", "a11m")):
            out, lv = getattr(impl, fn)(G["a10m_img"], G[f"a10m_imp_{tag}"], 16)
            assert np.array_equal(lv, G[f"{key}_map_{tag}"]), (fn, tag)
            assert np.array_equal(out, G[f"{key}_out_{tag}"]), (fn, tag)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "f")
        impl.write_per_block_qpfile(G["f3_scores_lin"], 128, 640, 384, path)
        assert open(path, "rb").read() == G["f3_qpfile_lin"].tobytes()


def test_oracle_reproduces_golden_vectors():
    _check_all(_Impl("oracle"))
    _check_round2(_Impl("oracle"))
    from oracle import spec_cv
    for tag in ("f32", "f64"):
        assert np.array_equal(spec_cv.resize_linear_float(G[f"lin_src_{tag}"], 27, 33), G[f"lin_up_{tag}"])
        assert np.array_equal(spec_cv.resize_linear_float(G[f"lin_src_{tag}"], 4, 7), G[f"lin_down_{tag}"])
    assert np.array_equal(spec_cv.rgb_to_gray(G["gray_rgb"]), G["gray_out"])
    packed, shape = P.pack_masks(G["a13_masks"])
    assert np.array_equal(packed, G["a13_packed"])
    assert np.array_equal(P.unpack_masks(packed, shape), G["a13_masks"])


@pytest.mark.gpu
def test_cuda_reproduces_golden_vectors():
    impl = _Impl("gpu")
    _check_all(impl)
    _check_round2(impl)
    import torch
    from elvis_b200 import ops
    for tag in ("f32", "f64"):
        src = torch.from_numpy(G[f"lin_src_{tag}"]).cuda()[None]
        assert np.array_equal(ops.resize_linear_float(src, 27, 33)[0].cpu().numpy(), G[f"lin_up_{tag}"])
        assert np.array_equal(ops.resize_linear_float(src, 4, 7)[0].cpu().numpy(), G[f"lin_down_{tag}"])
    assert np.array_equal(ops.rgb_to_gray(torch.from_numpy(G["gray_rgb"]).cuda()).cpu().numpy(), G["gray_out"])
    packed, shape = impl.E.pack_removal_masks(G["a13_masks"])
    assert np.array_equal(packed, G["a13_packed"])
    assert np.array_equal(impl.E.unpack_removal_masks(packed, shape), G["a13_masks"])
