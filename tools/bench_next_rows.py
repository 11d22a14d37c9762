"""Times the kernels of the SURVEY 8f rows 2-3 at 4K: row+column plan / gather / expand (bs16, 240 x 135 blocks),
RGB -> I420 and the ROI maps."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from elvis_b200 import ops  # noqa: E402

dev = torch.device("cuda")
T, by, bx, bs = 16, 135, 240, 16
rng = np.random.default_rng(0)
imp = torch.from_numpy(rng.random((T, by, bx))).to(dev)
clip = torch.randint(0, 256, (T, by * bs, bx * bs, 3), dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for amount in (0.1, 0.5):
    target = int(by * bx * amount)
    fby, fbx, counts = ops.rowcol_dims(by, bx, target)
    t_plan = timed(lambda: ops.rowcol_plan(imp, target))
    mask, pos, pidx, pcnt, meta = ops.rowcol_plan(imp, target)
    out = torch.empty((T, fby * bs, fbx * bs, 3), dtype=torch.uint8, device=dev)
    t_g = timed(lambda: ops.gather_blocks(clip, pos, bs, fby, fbx, out=out))
    P = len(counts)
    t_e = timed(lambda: ops.rowcol_expand(pidx[:, :P].contiguous(), pcnt[:, :P].contiguous(), fby, fbx))
    print(json.dumps({"shrink": amount, "frames": T, "passes": P, "final": [fby, fbx], "plan_ms": round(t_plan, 3),
                      "gather_ms": round(t_g, 3), "gather_GBps": round(2 * out.numel() / t_g / 1e6, 1), "expand_ms": round(t_e, 3)}))

# 8f rank 3: RGB -> I420 of 4K frames (37.3 MB of traffic per frame)
rgb = torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device=dev)
i420 = torch.empty((8, 2160 * 3840 * 3 // 2), dtype=torch.uint8, device=dev)
t_c = timed(lambda: ops.rgb_to_i420(rgb, out=i420), n=10)
print(json.dumps({"rgb_to_i420_ms_per_8_4k_frames": round(t_c, 4), "GBps": round((rgb.numel() + i420.numel()) / t_c / 1e6, 1)}))
imp = torch.rand((120, 135, 240), dtype=torch.float64, device=dev)
t_k = timed(lambda: ops.roi_kvazaar(imp, 40, 15), n=10)
t_s = timed(lambda: ops.roi_svtav1_offsets(ops.resize_area_f32(ops.roi_prepare_f32(imp, 0), 34, 60), 35, 10), n=10)
print(json.dumps({"kvazaar_dqp_ms_120_frames": round(t_k, 4), "svtav1_offsets_ms_120_frames": round(t_s, 4)}))
