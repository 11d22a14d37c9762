"""Debug helper: do the scoring kernel and the block-move kernels run concurrently on two streams?"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from elvis_b200 import ops
from elvis_b200.pipeline import ElvisV1, Yuv420
from elvis_b200.synth import synth_yuv420
dev = torch.device("cuda"); T, H, W = 120, 2160, 3840
clip = synth_yuv420(T, H, W, device=dev)
pipe = ElvisV1(16, 0.5, 0.5, 0.5)
scores = pipe.score(clip); shrunk, mask = pipe.shrink(clip, scores); full = pipe.stretch(shrunk, mask)
s1, s2 = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
def timed(fn, reps=5):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
def score(): ops.score_sc_tc(clip.y, 16)
def move():
    ops.shrink(clip.y, mask, 16, 120, out=shrunk.y); ops.shrink(clip.u, mask, 8, 120, out=shrunk.u); ops.shrink(clip.v, mask, 8, 120, out=shrunk.v)
    ops.stretch(shrunk.y, mask, 16, out=full.y); ops.stretch(shrunk.u, mask, 8, out=full.u); ops.stretch(shrunk.v, mask, 8, out=full.v)
def both(order):
    cur = torch.cuda.current_stream(); e = torch.cuda.Event(); e.record()
    s1.wait_event(e); s2.wait_event(e)
    def a():
        with torch.cuda.stream(s1): score()
    def b():
        with torch.cuda.stream(s2): move()
    (a(), b()) if order == 0 else (b(), a())
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    with torch.cuda.stream(s1): e1.record()
    with torch.cuda.stream(s2): e2.record()
    cur.wait_event(e1); cur.wait_event(e2)
print("score alone", round(timed(score), 3), "move alone", round(timed(move), 3))
print("both, score launched first", round(timed(lambda: both(0)), 3), "move first", round(timed(lambda: both(1)), 3))
for cap in (8, 4, 3, 2, 1):
    os.environ["ELVIS_MOVE_CTAS_PER_SM"] = str(cap)
    print("move CTAs/SM", cap, "move alone", round(timed(move), 3), "both", round(timed(lambda: both(0)), 3), round(timed(lambda: both(1)), 3))
