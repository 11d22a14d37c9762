"""Debug helper (not part of the product): CPU-side enqueue time vs GPU time of the sharded
step at N ranks, to see whether the step is launch bound."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from elvis_b200 import sharding, ops
from elvis_b200.pipeline import ElvisV1, Yuv420
from elvis_b200.synth import synth_yuv420
world=int(os.environ["WORLD_SIZE"]); rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
T,H,W=120,2160,3840
halo=sharding.HaloClip(T,H,W,dev); chroma=torch.empty((2,T,H//2,W//2),dtype=torch.uint8,device=dev)
clip=Yuv420(halo.owned,chroma[0],chroma[1]); synth_yuv420(T,H,W,device=dev,out=clip,frame_offset=rank*T)
pipe=ElvisV1(16,0.5,0.5,0.5)
shrunk=Yuv420.empty(T,H,W//2,dev); full=Yuv420.empty(T,H,W,dev)
def step(mode):
    c=[time.perf_counter()]
    if mode=="sharded": scores=sharding.sharded_removability(halo,T*world,16,0.5,0.5,rank,world)
    else: scores=pipe.score(clip)
    c.append(time.perf_counter())
    _,mask=pipe.shrink(clip,scores,shrunk); c.append(time.perf_counter())
    pipe.stretch(shrunk,mask,full); c.append(time.perf_counter())
    return [c[i+1]-c[i] for i in range(3)]
for mode in ("local","sharded"):
    for _ in range(3): step(mode)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); a.record(); cpu=[]
    for _ in range(10): cpu.append(step(mode))
    t1=time.perf_counter(); b.record(); torch.cuda.synchronize(); t2=time.perf_counter()
    if rank==0:
        print(mode, "gpu ms/step", round(a.elapsed_time(b)/10,3), "cpu enqueue ms/step", round((t1-t0)*100,3), "wall", round((t2-t0)*100,3),
              "cpu split ms", [round(sum(x[i] for x in cpu)*100,3) for i in range(3)])
dist.destroy_process_group()
