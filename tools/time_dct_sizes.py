import torch, time, sys
sys.path.insert(0, "/root/repo")
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
