"""Time the degradations on PACKED 3-channel frames (the reference's own frame format, H x W x 3), 8 4K frames."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elvis_b200 import ops
dev = torch.device("cuda")
T, H, W, bs = 8, 2160, 3840, 16
g = torch.Generator(device=dev).manual_seed(1)
clip = torch.randint(0, 256, (T, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
out = torch.empty_like(clip)
rounds = torch.randint(0, 11, (T, H // bs, W // bs), generator=g, device=dev, dtype=torch.int32)
levels = torch.randint(0, 5, (T, H // bs, W // bs), generator=g, device=dev, dtype=torch.int32)
strength = torch.rand((T, H // bs, W // bs), generator=g, device=dev)
def timed(name, fn):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(name, round(ms, 3), "ms per", T, "packed 4K frames =", round(ms / T * 1e3, 1), "us per frame,", round(2 * clip.numel() / ms / 1e6), "GB/s")
timed("blur rounds 0..10", lambda: ops.degrade_blur(clip, rounds, bs, out=out))
timed("downsample pow2 0..4", lambda: ops.degrade_downsample(clip, levels, bs, [16, 8, 4, 2, 1], out=out))
timed("dct dampen", lambda: ops.dct_dampen(clip, strength, bs, out=out))
