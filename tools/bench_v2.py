"""Throughput of the v2 per-block degradations on a planar 4K clip:
ms per clip, frames/s and GB/s against the algorithmic bytes (read + write the frame = 3*W*H
for YUV 4:2:0).  Prints one JSON line per kernel.  python tools/bench_v2.py [frames]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elvis_b200 import ops
from elvis_b200.pipeline import PresleyV2, Yuv420
from elvis_b200.synth import synth_yuv420

T = int(sys.argv[1]) if len(sys.argv) > 1 else 30
H, W, BS = 2160, 3840, 16
dev = torch.device("cuda")
clip = synth_yuv420(T, H, W, device=dev)
out = Yuv420.empty(T, H, W, dev)
g = torch.Generator(device=dev).manual_seed(7)
scores = torch.rand((T, H // BS, W // BS), generator=g, device=dev, dtype=torch.float64)
v2 = PresleyV2(BS)
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6650.0) if os.path.exists("MEASURED_PEAKS.json") else 6650.0
bytes_per_clip = 3 * W * H * T

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

cases = {}
rounds10 = ops.levels_from_scores(scores, ops.LEVELS_ROUND, 10)           # elvis.py:2176  (0..10 rounds)
rounds4 = ops.levels_from_scores(scores, ops.LEVELS_INVERTED_ROUND, 4)    # presley.py:1483 (0..4 rounds)
lv5 = ops.levels_from_scores(scores, ops.LEVELS_ROUND, 4)                 # elvis.py:2146 at bs=16 (levels 0..4)
lv2bit = lv5.clamp(max=3)                                                 # the 2-bit 1x/2x/4x/8x map
strength = scores.float()
cases["blur rounds 0..10 (elvis filter_frame_gaussian)"] = lambda: v2.blur(clip, rounds10, out)
cases["blur rounds 0..4 (presley blur_block)"] = lambda: v2.blur(clip, rounds4, out)
cases["downsample pow2 levels 0..4 (elvis filter_frame_downsample)"] = lambda: v2.downsample_pow2(clip, lv5, 4, out)
cases["downsample pow2 levels 0..3 (2-bit map)"] = lambda: v2.downsample_pow2(clip, lv2bit, 3, out)
def dampen_with(impl):
    def run():
        os.environ["ELVIS_DAMPEN_IMPL"] = impl
        v2.dampen(clip, strength, out)
        os.environ.pop("ELVIS_DAMPEN_IMPL")
    return run
cases["dct dampen"] = lambda: v2.dampen(clip, strength, out)
cases["dct dampen (ELVIS_DAMPEN_IMPL=pair)"] = dampen_with("pair")
cases["dct dampen (ELVIS_DAMPEN_IMPL=packed)"] = dampen_with("packed")
cases["levels_from_scores + pack 2-bit"] = lambda: ops.pack_levels_2bit(ops.levels_from_scores(scores, ops.LEVELS_ROUND, 3))
for name, fn in cases.items():
    ms = timed(fn)
    print(json.dumps({"kernel": name, "frames": T, "ms_per_clip": round(ms, 3), "frames_per_s": round(T / ms * 1e3),
                      "algorithmic_gbs": round(bytes_per_clip / ms / 1e6, 1), "frac_of_measured_hbm_peak": round(bytes_per_clip / ms / 1e6 / peak, 4)}))
