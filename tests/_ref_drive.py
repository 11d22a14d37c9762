"""Drive the unmodified reference `calculate_removability_scores` (elvis.py:968-1224) with
its two external tools faked, so that its in-tree tail (elvis.py:1160-1220) can be used as
a checker.  EVCA (`python -m evca.main`, elvis.py:1014-1031) is replaced by a function
that writes the caller's SC/TC as the CSVs the tail reads; UFO (elvis.py:1109-1113) by a
function that writes the caller's foreground masks as PNGs.  Build-container only."""
from __future__ import annotations

import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

from oracle import ref_import


def run_reference_removability(sc, tc, fg_masks, alpha, beta, block_size):
    """sc, tc: (T, By, Bx) float64.  fg_masks: None or (T, By, Bx) uint8 (0 = background),
    written at block resolution so the reference's NEAREST resize is the identity."""
    import cv2
    E = ref_import.load("elvis")
    T, By, Bx = sc.shape
    width, height = Bx * block_size, By * block_size
    with tempfile.TemporaryDirectory() as tmp:
        pkg_dir = os.path.join(tmp, "evca")
        os.makedirs(pkg_dir)
        frames_dir = os.path.join(tmp, "frames")
        os.makedirs(frames_dir)
        for i in range(T):
            open(os.path.join(frames_dir, f"{i + 1:05d}.png"), "wb").close()
        raw = os.path.join(tmp, "raw.yuv")
        open(raw, "wb").close()
        fake_evca = types.ModuleType("evca")
        fake_evca.__file__ = os.path.join(pkg_dir, "__init__.py")

        def fake_run(cmd, *a, **k):
            for name, arr in (("evca_SC_blocks.csv", sc), ("evca_TC_blocks.csv", tc)):
                table = arr.reshape(T, By * Bx).T        # rows = blocks, cols = frames
                with open(os.path.join(pkg_dir, name), "w") as f:
                    f.write(",".join(f"f{i}" for i in range(T)) + "\n")
                    for row in table:
                        f.write(",".join(repr(float(v)) for v in row) + "\n")
            return types.SimpleNamespace(returncode=0, stdout="", stderr="")

        def fake_ufo(device, model, datapath, save_root, *a):
            if fg_masks is None:
                return
            for i in range(T):
                cv2.imwrite(os.path.join(save_root[0], f"{i + 1:05d}.png"), fg_masks[i])

        fake_ufo_pkg = types.ModuleType("ufo")
        fake_ufo_pkg.__file__ = os.path.join(tmp, "ufo", "__init__.py")
        fake_ufo_test = types.ModuleType("ufo.test")
        fake_ufo_test.debug_test = fake_ufo
        with mock.patch.dict(sys.modules, {"evca": fake_evca, "ufo": fake_ufo_pkg, "ufo.test": fake_ufo_test}), \
                mock.patch.object(E.subprocess, "run", fake_run), \
                mock.patch("builtins.print", lambda *a, **k: None):
            out = E.calculate_removability_scores(raw, frames_dir, width, height, block_size,
                                                  alpha=alpha, working_dir=tmp, smoothing_beta=beta)
    return np.asarray(out)


def reference_qpfile(E, scores, block_size, width, height, workdir):
    """Text of the per-block qpfile the reference's encode_with_roi (elvis.py:2013-2139) writes,
    read at the point where it is handed to encode_video (which is replaced: no ffmpeg here)."""
    import os
    os.makedirs(workdir, exist_ok=True)
    captured = {}

    def fake_encode_video(**kwargs):
        with open(kwargs["qpfile"]) as f:
            captured["text"] = f.read()

    real = E.encode_video
    E.encode_video = fake_encode_video
    try:
        E.encode_with_roi(workdir, os.path.join(workdir, "out.mp4"), scores, block_size, 30.0, width, height)
    finally:
        E.encode_video = real
    return captured["text"]


def reference_map_video_arithmetic(E, strength_maps, kind, block_size, workdir):
    """Runs the reference's encode_strength_maps / decode_strength_maps (elvis.py:2198-2245) with the
    external video encode / decode replaced by nothing: returns (the uint8 frames it would encode,
    the maps it reconstructs from exactly those frames)."""
    import os
    import cv2
    video = os.path.join(workdir, f"maps_{kind}.mp4")      # decode keys its range off the file name
    real_enc, real_dec = E.encode_video, E.decode_video
    E.encode_video = lambda **kw: None
    E.decode_video = lambda *a, **kw: None
    try:
        E.encode_strength_maps(strength_maps, video, 30.0)
        frames_dir = os.path.splitext(video)[0]
        files = sorted(f for f in os.listdir(frames_dir) if f.endswith(".png"))
        gray = np.stack([cv2.imread(os.path.join(frames_dir, f), cv2.IMREAD_GRAYSCALE) for f in files])
        decoded = E.decode_strength_maps(video, block_size, frames_dir)
    finally:
        E.encode_video, E.decode_video = real_enc, real_dec
    return gray, decoded
