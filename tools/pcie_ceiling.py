"""Plain pinned-copy ceiling of the box for the e2e leg of bench.py: every rank moves the byte volumes of one
v1 step (1.49 GB host->device, 2.24 GB device->host per 120-frame 4K clip) between pinned host memory and its
GPU on two streams, nothing else.  Launch like bench.py (python tools/pcie_ceiling.py, or torchrun with N ranks);
rank 0 prints one JSON line: aggregate GB/s in each direction and the frames/s those copies alone would allow."""
import json
import os
import time

import torch
import torch.distributed as dist

H2D, D2H, FRAMES = 1_492_992_000, 2_243_376_000, 120


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h_in = torch.empty(H2D, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(D2H, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(H2D, dtype=torch.uint8, device=dev)
    d_out = torch.empty(D2H, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    results = {}
    for mode in ("h2d", "d2h", "both"):
        def step():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        n = 6
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        results[mode] = sec / n
    if rank == 0:
        line = {"n_gpus": world, "h2d_alone_gbs": H2D * world / results["h2d"] / 1e9, "d2h_alone_gbs": D2H * world / results["d2h"] / 1e9,
                "both_h2d_gbs": H2D * world / results["both"] / 1e9, "both_d2h_gbs": D2H * world / results["both"] / 1e9,
                "ms_per_step_both": results["both"] * 1e3, "copy_only_frames_per_s": FRAMES * world / results["both"],
                "note": "pinned host memory, one H2D and one D2H stream per rank, byte volumes of one v1 e2e step"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
