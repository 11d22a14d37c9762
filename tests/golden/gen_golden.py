"""Freeze golden input/output vectors from the UNMODIFIED reference (imported from
/root/reference through oracle/ref_import.py; calculate_removability_scores is driven with
EVCA/UFO faked, see tests/_ref_drive.py).  Run in the build container:

    python tests/golden/gen_golden.py

Library versions of the run are recorded in the file (the reference pins numpy<2 and
opencv-python 4.8.0.76; this image has numpy 2.3 / cv2 4.13)."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from _ref_drive import run_reference_removability  # noqa: E402
from oracle import ref_import  # noqa: E402


def main():
    import cv2
    E, U = ref_import.load("elvis"), ref_import.load("utils")
    rng = np.random.default_rng(20260101)
    g = {"versions": np.array([np.__version__, cv2.__version__])}

    # a2 -- in-tree tail of calculate_removability_scores
    T, By, Bx = 5, 4, 6
    sc, tc = rng.random((T, By, Bx)) * 60, rng.random((T, By, Bx)) * 30
    tc[0] = 0
    fg = (rng.random((T, By, Bx)) > 0.4).astype(np.uint8) * 255
    g.update(a2_sc=sc, a2_tc=tc, a2_fg=fg)
    for tag, beta, masks in (("b1", 1, None), ("b05", 0.5, None), ("b05m", 0.5, fg), ("b025m", 0.25, fg)):
        g[f"a2_out_{tag}"] = run_reference_removability(sc, tc, masks, 0.3, beta, 16)

    # a3 -- calculate_importance_scores
    class Cx:
        pass
    cx = Cx()
    cx.SC, cx.TC = sc, tc
    fgf = rng.random((T, By, Bx))
    g["a3_fg"] = fgf
    g["a3_out"] = np.stack(U.calculate_importance_scores(None, 16, 0.3, 0.6, cx, fgf))

    # a4 / a5 -- elvis shrink + stretch
    img = rng.integers(0, 256, (32, 48, 3), dtype=np.uint8)
    scores = rng.random((4, 6))
    g.update(a4_img=img, a4_scores=scores)
    for tag, bs, amount in (("s25", 8, 0.25), ("s50", 8, 0.5), ("n2", 8, 2)):
        s8 = rng.random((32 // bs, 48 // bs))
        g[f"a4_scores_{tag}"] = s8
        small, mask, _ = E.apply_selective_removal(img, s8, bs, amount)
        g[f"a4_small_{tag}"], g[f"a4_mask_{tag}"] = small, mask
        g[f"a5_full_{tag}"] = E.stretch_frame(small, mask, bs)

    # a6 / a7 -- utils row-only shrink (incl. the partial-last-pass quirk) + stretch
    img2 = rng.integers(0, 256, (43, 67, 3), dtype=np.uint8)
    g["a6_img"] = img2
    for tag, amount in (("q30", 0.3), ("q50", 0.5)):
        imp = np.round(rng.random((5, 8)) * 8) / 8      # ties on purpose
        g[f"a6_imp_{tag}"] = imp
        small, mask = U.shrink_frame_row_only(img2, imp, 8, amount)
        g[f"a6_small_{tag}"], g[f"a6_mask_{tag}"] = small, mask
        g[f"a7_full_{tag}"] = U.stretch_frame_row_only(small, mask, 8)

    # a8 / a9 -- elvis filters; a10 / a11 -- utils degradations (with crop)
    for bs in (8, 16):
        im = rng.integers(0, 256, (bs * 3, bs * 4, 3), dtype=np.uint8)
        s = rng.random((3, 4))
        s[0, :4] = [0.5, 0.125, 1.0, 0.0]
        g[f"a8_img_{bs}"], g[f"a8_scores_{bs}"] = im, s
        g[f"a8_out_{bs}"], g[f"a8_map_{bs}"] = E.filter_frame_downsample(im, s, bs)
        g[f"a9_out_{bs}"], g[f"a9_map_{bs}"] = E.filter_frame_gaussian(im, s, bs)
        im2 = rng.integers(0, 256, (bs * 3 + 3, bs * 4 + 5, 3), dtype=np.uint8)
        g[f"a10_img_{bs}"] = im2
        g[f"a10_out_{bs}"], g[f"a10_map_{bs}"] = U.degrade_adaptive_downsample(im2, s, bs)
        g[f"a11_out_{bs}"], g[f"a11_map_{bs}"] = U.degrade_adaptive_blur(im2, s, bs)

    # 8f rank 1 -- OpenCV client restorers (unsharp mask; halo + temporal blend variants)
    rimg = rng.integers(0, 256, (48, 80, 3), dtype=np.uint8)
    rmap = rng.integers(0, 11, (3, 5))
    g.update(f1_img=rimg, f1_map=rmap, f1_out=E.restore_blur_opencv_unsharp_mask(rimg, rmap, 16))
    rframes = np.stack([rng.integers(0, 256, (37, 53, 3), dtype=np.uint8) for _ in range(3)])
    rmaps = rng.integers(0, 5, (3, 2, 3))
    g.update(f1_frames=rframes, f1_maps=rmaps)
    for tag, halo, tb in (("h0", 0, 0.0), ("h6b", 6, 0.2)):
        g[f"f1_out_{tag}"] = np.stack(U.restore_with_opencv_unsharp(list(rframes), rmaps, 16, halo=halo, temporal_blend=tb))

    # a13 -- mask side channel (elvis.py:4412-4418)
    masks = (rng.random((3, 5, 7)) > 0.5).astype(np.int8)
    g["a13_masks"] = masks
    g["a13_packed"] = np.packbits(masks.astype(np.uint8))

    # 8f rank 2 -- row+column shrink variants (utils.py:763-1018); ties and a partial last pass
    img3 = rng.integers(0, 256, (51, 70, 3), dtype=np.uint8)
    g["f2_img"] = img3
    for tag, amount in (("q30", 0.3), ("q55", 0.55), ("q85", 0.85)):
        imp = np.round(rng.random((6, 8)) * 8) / 8
        g[f"f2_imp_{tag}"] = imp
        small, mask, pmap = U.shrink_frame_position_map(img3, imp, 8, amount)
        g[f"f2_small_{tag}"], g[f"f2_mask_{tag}"], g[f"f2_pmap_{tag}"] = small, mask, pmap
        g[f"f2_full_pm_{tag}"] = U.stretch_frame_position_map(small, mask, pmap, 8)
        small2, mask2, passes = U.shrink_frame_removal_indices(img3, imp, 8, amount)
        assert np.array_equal(small, small2) and np.array_equal(mask, mask2)
        g[f"f2_npass_{tag}"] = np.array([len(a) for a in passes], np.int32)
        g[f"f2_passes_{tag}"] = np.concatenate(passes).astype(np.int32)
        g[f"f2_full_ri_{tag}"] = U.stretch_frame_removal_indices(small2, passes, 6, 8, 8)

    # 8f rank 3 -- ROI side files and the Y4M frame conversion, as file bytes
    import tempfile
    from _ref_drive import reference_qpfile
    with tempfile.TemporaryDirectory() as td:
        imps = [rng.random((17, 30)) for _ in range(2)]
        imps[0][0, :4] = [0.0, 1.0, 0.5, 0.125]
        g["f3_imps"] = np.stack(imps)
        U.create_kvazaar_roi_file(imps, os.path.join(td, "k.bin"), 40, 15)
        g["f3_kvazaar"] = np.frombuffer(open(os.path.join(td, "k.bin"), "rb").read(), np.uint8)
        U.create_svtav1_roi_file(imps, os.path.join(td, "s.txt"), 35, 10, 480, 272)
        g["f3_svtav1"] = np.frombuffer(open(os.path.join(td, "s.txt"), "rb").read(), np.uint8)
        rgb = [rng.integers(0, 256, (34, 50, 3), dtype=np.uint8) for _ in range(2)]
        g["f3_rgb"] = np.stack(rgb)
        U.write_y4m(rgb, os.path.join(td, "v.y4m"), 30.0)
        g["f3_y4m"] = np.frombuffer(open(os.path.join(td, "v.y4m"), "rb").read(), np.uint8)
        qs = rng.random((2, 34, 60))
        g["f3_scores"] = qs
        g["f3_qpfile"] = np.frombuffer(reference_qpfile(E, qs, 16, 960, 544, os.path.join(td, "q")).encode(), np.uint8)
        g["f3_qpfile_4k"] = np.frombuffer(reference_qpfile(E, qs[:1, :, :60].repeat(4, 1).repeat(4, 2)[:, :135, :240].copy(), 16, 3840, 2160,
                                                           os.path.join(td, "q4")).encode(), np.uint8)

    # 8f rank 4 -- pyramid reconstruction with a deterministic 2x upsampler (pixel repetition) standing in for
    # the external one, and the arithmetic of the level-map video
    from _ref_drive import reference_map_video_arithmetic
    pimg = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    pmaps = rng.integers(0, 5, (4, 6))
    pmaps.flat[0] = 4
    g.update(f4_img=pimg, f4_maps=pmaps)
    g["f4_out"] = E.upscale_realesrgan_adaptive(pimg, pmaps.copy(), 16, upsample_fn=lambda im: np.ascontiguousarray(im.repeat(2, 0).repeat(2, 1)))
    with tempfile.TemporaryDirectory() as td:
        lv = rng.integers(0, 11, (2, 17, 30)).astype(np.int32)
        lv.flat[:2] = [0, 10]
        g["f4_levels"] = lv
        g["f4_gray"], g["f4_decoded"] = reference_map_video_arithmetic(E, lv, "gaussian", 16, td)

    # ---- round 2 additions (appended so that the random stream of everything above is unchanged)
    # a7 -- presley batch shrink / stretch incl. partial last passes (row-major refill, presley.py:787-827)
    PF = ref_import.load_presley_functions("shrink_frame_row_only", "shrink_video_frames", "stretch_video_frames")
    pframes = [rng.integers(0, 256, (43, 59, 3), dtype=np.uint8) for _ in range(2)]
    g["p7_frames"] = np.stack(pframes)
    for tag, amount in (("q25", 0.25), ("q33", 0.33), ("q50", 0.5), ("q60", 0.6)):
        imps = [np.round(rng.random((5, 7)) * 16) / 16 for _ in range(2)]
        g[f"p7_imp_{tag}"] = np.stack(imps)
        small, masks = PF["shrink_video_frames"](pframes, imps, 8, amount, PF["shrink_frame_row_only"])
        g[f"p7_small_{tag}"], g[f"p7_mask_{tag}"] = np.stack(small), np.stack(masks)
        g[f"p7_full_{tag}"] = np.stack(PF["stretch_video_frames"](small, masks, 8))

    # utils degradations with an importance map whose grid differs from the frame's (cv2 float64 INTER_LINEAR)
    im3 = rng.integers(0, 256, (16 * 3 + 3, 16 * 4 + 5, 3), dtype=np.uint8)
    g["a10m_img"] = im3
    for tag, shape in (("up", (2, 3)), ("down", (7, 9))):
        imp = rng.random(shape)
        g[f"a10m_imp_{tag}"] = imp
        g[f"a10m_out_{tag}"], g[f"a10m_map_{tag}"] = U.degrade_adaptive_downsample(im3, imp, 16)
        g[f"a11m_out_{tag}"], g[f"a11m_map_{tag}"] = U.degrade_adaptive_blur(im3, imp, 16)

    # x265 qpfile with a CTU grid finer than the block grid (block 128 > CTU 64: cv2 float32 INTER_LINEAR)
    with tempfile.TemporaryDirectory() as td:
        ql = rng.random((2, 3, 5))
        g["f3_scores_lin"] = ql
        g["f3_qpfile_lin"] = np.frombuffer(reference_qpfile(E, ql, 128, 640, 384, os.path.join(td, "ql")).encode(), np.uint8)

    # cv2 primitives the round-2 kernels restate, frozen from the cv2 of this image
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        m = (rng.random((9, 16)) * 2 - 0.5).astype(dt)
        g[f"lin_src_{tag}"] = m
        g[f"lin_up_{tag}"] = cv2.resize(m, (33, 27), interpolation=cv2.INTER_LINEAR)
        g[f"lin_down_{tag}"] = cv2.resize(m, (7, 4), interpolation=cv2.INTER_LINEAR)
    rgbg = rng.integers(0, 256, (2, 21, 37, 3), dtype=np.uint8)
    g["gray_rgb"] = rgbg
    g["gray_out"] = np.stack([cv2.cvtColor(f, cv2.COLOR_RGB2GRAY) for f in rgbg])

    out = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(out, **g)
    print(f"wrote {out}: {len(g)} arrays, {os.path.getsize(out) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
