"""Times this package's MIRRORS of the reference functions the way a drop-in user calls them -- NumPy frames in, NumPy
frames out, one packed BGR 4K frame per call (host <-> device copies included) -- to be read beside
tools/time_reference_functions.py (the unmodified reference functions on the CPU).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elvis_b200 import elvis as E  # noqa: E402
from elvis_b200 import presley as Pr  # noqa: E402
from elvis_b200 import utils as U  # noqa: E402


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    H, W, bs = 2160, 3840, 16
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    scores = rng.random((H // bs, W // bs))
    shrunk, mask, _ = E.apply_selective_removal(frame, scores, bs, 0.5)
    frames30 = [frame] * 8
    imps = [scores] * 8
    out = {
        "elvis.apply_selective_removal": timed(lambda: E.apply_selective_removal(frame, scores, bs, 0.5)),
        "elvis.stretch_frame": timed(lambda: E.stretch_frame(shrunk, mask, bs)),
        "elvis.filter_frame_downsample": timed(lambda: E.filter_frame_downsample(frame, scores, bs)),
        "elvis.filter_frame_gaussian": timed(lambda: E.filter_frame_gaussian(frame, scores, bs)),
        "utils.degrade_adaptive_downsample": timed(lambda: U.degrade_adaptive_downsample(frame, scores, bs)),
        "utils.degrade_adaptive_blur": timed(lambda: U.degrade_adaptive_blur(frame, scores, bs)),
        "utils.shrink_frame_row_only": timed(lambda: U.shrink_frame_row_only(frame, scores, bs, 0.5)),
        "presley.degrade_video_adaptive(blur, 8 frames) per frame": timed(lambda: Pr.degrade_video_adaptive(frames30, imps, bs, 4, Pr.blur_block), 2) / 8,
        "presley.degrade_video_adaptive(downscale, 8 frames) per frame": timed(lambda: Pr.degrade_video_adaptive(frames30, imps, bs, 4, Pr.downscale_block), 2) / 8,
    }
    print(json.dumps({"what": "elvis_b200 mirrors, NumPy in / NumPy out, packed BGR 4K frames, 16x16 blocks (host<->device copies included)",
                      "gpu": torch.cuda.get_device_name(0), "ms_per_frame": {k: round(v, 3) for k, v in out.items()}}))


if __name__ == "__main__":
    main()
