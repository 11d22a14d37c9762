"""Debug helper: time elvis_score_sc_tc alone for several ELVIS_SCORE_CHUNK settings."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from elvis_b200 import ops
from elvis_b200.synth import synth_yuv420
dev = torch.device("cuda"); T, H, W = 120, 2160, 3840
clip = synth_yuv420(T, H, W, device=dev)
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
for chunk in sys.argv[1:]:
    os.environ["ELVIS_SCORE_CHUNK"] = chunk
    print("chunk", chunk, round(timed(lambda: ops.score_sc_tc(clip.y, 16)), 4), "ms")
