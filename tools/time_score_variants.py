"""Time the tcgen05 scoring kernel alone (120 4K frames) under the environment's ELVIS_UMMA_* knobs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elvis_b200 import ops
from elvis_b200.synth import synth_yuv420
dev = torch.device("cuda")
clip = synth_yuv420(120, 2160, 3840, device=dev)
ops.score_sc_tc(clip.y, 16); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.score_sc_tc(clip.y, 16)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 10)
print({k: v for k, v in os.environ.items() if k.startswith("ELVIS_")}, round(best, 4), "ms per 120 4K frames")
