"""Time elvis_score_sc_tc on 120 4K frames with the 8 x 8 transform (tcgen05 kernel) and the block-sized 16 x 16 one."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elvis_b200 import ops
from elvis_b200.synth import synth_yuv420
dev = torch.device("cuda")
clip = synth_yuv420(120, 2160, 3840, device=dev)
for ds in (8, 16):
    ops.score_sc_tc(clip.y, 16, dct_size=ds); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): ops.score_sc_tc(clip.y, 16, dct_size=ds)
    b.record(); torch.cuda.synchronize()
    print("dct_size", ds, a.elapsed_time(b) / 5, "ms per 120 4K frames")
