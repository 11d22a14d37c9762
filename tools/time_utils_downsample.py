"""Time the utils-mode adaptive downsample (utils.py:1101-1168: scales {0, 2, 3, 4} -> block sizes 16, 8, 5, 4; the
16 -> 5 level has a fractional INTER_AREA factor) next to the power-of-two one, per plane, 30 4K frames."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elvis_b200 import ops
from elvis_b200.synth import synth_yuv420
dev = torch.device("cuda")
T, H, W = 30, 2160, 3840
clip = synth_yuv420(T, H, W, device=dev)
g = torch.Generator(device=dev).manual_seed(3)
lv = torch.randint(0, 4, (T, H // 16, W // 16), generator=g, device=dev, dtype=torch.int32)
out = [torch.empty_like(p) for p in clip.planes]
def run(smalls_of):
    for p, o, pb in zip(clip.planes, out, (16, 8, 8)):
        ops.degrade_downsample(p, lv, pb, smalls_of(pb), out=o)
for name, f in (("utils {16, 8, 5, 4}", lambda pb: [pb, max(1, pb // 2), max(1, pb // 3), max(1, pb // 4)]),
                ("pow2 {16, 8, 4, 2}", lambda pb: [pb, pb // 2, pb // 4, pb // 8])):
    run(f); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): run(f)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(name, round(ms, 3), "ms per 30 4K frames Y+U+V,", round(3 * W * H * T / ms / 1e6), "GB/s")
