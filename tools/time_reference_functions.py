"""Times the UNMODIFIED reference functions of the v1 path (single process, packed BGR frames, the way run_elvis
calls them: elvis.py:4389-4394, 4550-4557) in the build container, where /root/reference is mounted -- the GPU box
does not have the reference tree, so bench.py's CPU arm times the oracle port there.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import, spec_scoring  # noqa: E402


def main():
    import cv2
    E = ref_import.load("elvis")
    H, W, bs, n = 2160, 3840, 16, 3
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(n)]
    luma = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in frames])
    t0 = time.perf_counter()
    sc, tc = spec_scoring.sc_tc(luma, bs)
    t_score = (time.perf_counter() - t0) / n
    scores = rng.random((n, H // bs, W // bs))
    t0 = time.perf_counter()
    shrunk = [E.apply_selective_removal(f, s, bs, 0.5) for f, s in zip(frames, scores)]
    t_shrink = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    _ = [E.stretch_frame(sf, m, bs) for sf, m, _ in shrunk]
    t_stretch = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    _ = E.filter_frame_downsample(frames[0], scores[0], bs)
    t_down = time.perf_counter() - t0
    t0 = time.perf_counter()
    _ = E.filter_frame_gaussian(frames[0], scores[0], bs)
    t_blur = time.perf_counter() - t0
    per_frame = t_score + t_shrink + t_stretch
    print(json.dumps({"what": "unmodified reference functions, single process, packed BGR 4K frames, 16x16 blocks, 50 % removal",
                      "cpu": "build container, %d cores" % (os.cpu_count() or 1), "numpy": np.__version__, "cv2": cv2.__version__,
                      "ms_per_frame": {"spec_scoring.sc_tc (stands in for EVCA)": t_score * 1e3,
                                       "elvis.apply_selective_removal": t_shrink * 1e3, "elvis.stretch_frame": t_stretch * 1e3,
                                       "elvis.filter_frame_downsample": t_down * 1e3, "elvis.filter_frame_gaussian": t_blur * 1e3},
                      "v1_frames_per_s_single_process": 1.0 / per_frame}))


if __name__ == "__main__":
    main()
