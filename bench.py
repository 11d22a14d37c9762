#!/usr/bin/env python
"""Headline benchmark: 4K frames/s for score + shrink + stretch (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): ELVIS v1 on a synthetic planar YUV 4:2:0 clip, 3840x2160,
120 frames per GPU, 16x16 blocks, 50 % removal, alpha = beta = 0.5 (BASELINE.json
configs[1]).  One step = one pass of the whole hot path over the clip: SC/TC scoring ->
elvis combine/normalise -> per-row top-k mask -> shrink -> stretch.  With N > 1 (torchrun,
one rank per GPU) the ranks own consecutive 120-frame ranges of ONE 120*N-frame clip: a
one-frame luma halo is exchanged with the neighbours and the two global min/max
normalisations are all-reduced (weak scaling).

value  = frames/s with the clip resident in HBM (CUDA events, max over ranks)
e2e    = frames/s through elvis_b200.pipeline.HostElvisV1 with host buffers (pinned),
         H2D of the clip and D2H of masks + shrunk + stretched clips inside the timed region
--impl reference = the CPU arm: oracle port of the same path on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, FRAMES, BLOCK, SHRINK, ALPHA, BETA = 3840, 2160, 120, 16, 0.5, 0.5, 0.5
METRIC = "4K frames/s, score+shrink+stretch"
UNIT = "frames/s"


def algorithmic_bytes_per_frame(w=WIDTH, h=HEIGHT, s=SHRINK):
    """score reads Y once; shrink reads+writes the kept blocks; stretch reads
    the kept blocks and writes the full frame (YUV 4:2:0 = 1.5 bytes/pixel)."""
    score = w * h
    shrink = 2 * (1 - s) * 1.5 * w * h
    stretch = (1 - s) * 1.5 * w * h + 1.5 * w * h
    return {"score": score, "shrink": shrink, "stretch": stretch, "total": score + shrink + stretch}


def config_dict(n_gpus, frames=FRAMES):
    return {"workload": f"ELVIS v1 score+shrink+stretch, synthetic planar YUV420 {WIDTH}x{HEIGHT}, "
                        f"{frames} frames per GPU, {BLOCK}x{BLOCK} blocks, {int(SHRINK * 100)}% removal, "
                        f"alpha={ALPHA} beta={BETA} (BASELINE.json configs[1])",
            "frames_per_gpu": frames, "width": WIDTH, "height": HEIGHT, "block_size": BLOCK,
            "shrink_amount": SHRINK, "sharding": f"contiguous frame ranges x{n_gpus}, 1-frame luma halo",
            "l2": "inputs larger than L2 (1.49 GB clip per GPU vs 126 MB L2); no explicit flush",
            "pipelining": ("3 clips in flight on 3 CUDA streams: scoring of clip i+2, shrink of clip i+1 and stretch of clip i overlap"
                           if n_gpus == 1 else
                           "2 clips in flight on 2 CUDA streams (+1 for NCCL): scoring of clip i+1 overlaps shrink+stretch of clip i")
                          + " (elvis_b200.pipeline.ElvisV1Pipelined); every clip takes the full serial path; "
                            "roofline.serial_step is the un-overlapped figure"}


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~2 ms from a thread while the
    timed region runs (nvidia-smi's own loop is too coarse for a millisecond-scale step)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.sm, self.bits, self.max_mhz, self._stop, self.thread = [], 0, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = None
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def reset(self):
        """Drop the samples taken so far (warm-up); called right before the timed region."""
        self.sm, self.bits = [], 0

    def stop(self):
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._stop = True
        self.thread.join()
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for n, b in self.REASONS.items() if self.bits & b), "samples": len(self.sm)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_arm(sample_frames: int, steps: int, warmup: int):
    """Oracle port on the host cores, on the first `sample_frames` frames of the workload."""
    import torch
    from elvis_b200.synth import synth_yuv420
    from oracle.cpu_baseline import CpuElvisV1
    cores = os.cpu_count() or 1
    clip = synth_yuv420(sample_frames, HEIGHT, WIDTH, seed=1234, device="cpu")
    cpu = CpuElvisV1(clip.y.numpy(), clip.u.numpy(), clip.v.numpy(), BLOCK, SHRINK, ALPHA, BETA, workers=cores)
    try:
        for _ in range(warmup):
            cpu.step()
        times = [cpu.step() for _ in range(steps)]
    finally:
        cpu.close()
    sec = sum(times) / len(times)
    return {"value": sample_frames / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample_frames} frames of the 4K workload per step, {steps} steps after {warmup} warm-up; "
                      f"oracle/cpu_baseline.py (NumPy/SciPy port, fork pool over {cores} cores)",
            "ms_per_step": sec * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 16
    cb = cpu_arm(sample, max(1, args.steps), max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from elvis_b200 import ops, sharding
    from elvis_b200.pipeline import ElvisV1, ElvisV1Pipelined, HostElvisV1, Yuv420
    from elvis_b200.synth import synth_yuv420

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's internal streams at high priority: its small kernels (halo send/recv, 16-byte all-reduces) then
        # take the first SM slot that frees up instead of queueing behind the move kernels (N=2: 1.088 -> 1.080 ms)
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)
    T = args.frames
    pipe = ElvisV1(BLOCK, SHRINK, ALPHA, BETA)
    by, bx = HEIGHT // BLOCK, WIDTH // BLOCK
    k = int(SHRINK * bx)

    # rank r owns frames [r*T, (r+1)*T) of one global clip; luma lives in a halo buffer
    halo = sharding.HaloClip(T, HEIGHT, WIDTH, dev)
    chroma = torch.empty((2, T, HEIGHT // 2, WIDTH // 2), dtype=torch.uint8, device=dev)
    clip = Yuv420(halo.owned, chroma[0], chroma[1])
    synth_yuv420(T, HEIGHT, WIDTH, seed=1234, device=dev, out=clip, frame_offset=rank * T)
    shrunk = Yuv420.empty(T, HEIGHT, (bx - k) * BLOCK, dev)
    full = Yuv420.empty(T, HEIGHT, WIDTH, dev)

    def step(ev=None):
        if ev:
            ev[0].record()
        if world > 1:
            scores = sharding.sharded_removability(halo, T * world, BLOCK, ALPHA, BETA, rank, world)
        else:
            scores = pipe.score(clip)
        if ev:
            ev[1].record()
        _, mask = pipe.shrink(clip, scores, shrunk)
        if ev:
            ev[2].record()
        pipe.stretch(shrunk, mask, full)
        if ev:
            ev[3].record()
        return mask

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # (1) serial steps: one clip at a time on one stream -- per-stage breakdown and clip latency
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for i in range(args.steps):
        step(evs[i])
    barrier()
    stage_ms = [sum(e[j].elapsed_time(e[j + 1]) for e in evs) / args.steps for j in range(3)]
    serial_ms = sum(e[0].elapsed_time(e[3]) for e in evs) / args.steps

    # (2) the timed region: the same steps through the two-stream pipeline (scoring of clip i+1
    # overlaps shrink + stretch of clip i; every clip goes through the full serial path)
    score_fn = comm_fn = None
    if world > 1:   # halo exchange ahead of time on the communication stream, all-reduces inside the score stage
        comm_fn = lambda c: sharding.exchange_halo(halo, rank, world)  # noqa: E731
        score_fn = lambda c, slot: sharding.sharded_removability(halo, T * world, BLOCK, ALPHA, BETA, rank, world,  # noqa: E731
                                                                 exchange=False)
    # one GPU: three stages, three clips in flight (1.01 ms per clip; two stages: 1.07).  Sharded: two stages,
    # two clips (measured at N=2: 1.09 ms; three stages 1.11-1.35 ms -- the NCCL kernels wait for SM slots)
    split = (world == 1) if args.split_stretch is None else args.split_stretch
    depth = (3 if split else 2) if args.depth is None else args.depth
    pp = ElvisV1Pipelined(T, HEIGHT, WIDTH, BLOCK, SHRINK, ALPHA, BETA, dev, depth=depth, score_fn=score_fn,
                          move_ctas_per_sm=args.move_ctas, comm_fn=comm_fn, split_stretch=split,
                          stretch_ctas_per_sm=args.stretch_ctas)
    if args.serial:
        def run(n):
            for _ in range(n):
                step()
    else:
        def run(n):
            for _ in range(n):
                pp.submit(clip)
            pp.join()
    sampler = ClockSampler(local) if rank == 0 else None   # NVML init happens here, before the barrier
    run(max(3, args.warmup))
    barrier()
    if sampler:
        sampler.reset()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    run(args.steps)
    stop.record()
    barrier()
    ms = start.elapsed_time(stop)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = T * world * args.steps / (ms / 1e3)

    # dominant kernel: time its launches alone (same stream, CUDA events) over the timed clip
    def time_kernel(fn, reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ab = algorithmic_bytes_per_frame()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
    sc_ms = time_kernel(lambda: ops.score_sc_tc(clip.y, BLOCK), max(3, args.steps))
    stages = {"score_pipeline": {"ms": stage_ms[0], "gbs": ab["score"] * T / stage_ms[0] / 1e6},
              "shrink": {"ms": stage_ms[1], "gbs": ab["shrink"] * T / stage_ms[1] / 1e6},
              "stretch": {"ms": stage_ms[2], "gbs": ab["stretch"] * T / stage_ms[2] / 1e6},
              "score_kernel_alone": {"ms": sc_ms, "gbs": ab["score"] * T / sc_ms / 1e6}}
    achieved = ab["score"] * T / sc_ms / 1e6
    roofline = {"bound": "hbm", "kernel": "score_umma_kernel<2> (elvis_score_sc_tc)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch over the 120-frame clip,
                # ncu --set full capture profiles/r1d_ncu_full_summary.csv (1.0038e9 + 0.0327e9)
                "traffic": 1.0364e9 if T == FRAMES else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ab["score"] * T,
                "whole_step": {"achieved": ab["total"] * T / ms_per_step / 1e6, "frac": ab["total"] * T / ms_per_step / 1e6 / peak,
                               "note": "all three stages, pipelined as timed"},
                "serial_step": {"ms": serial_ms, "achieved": ab["total"] * T / serial_ms / 1e6,
                                "frac": ab["total"] * T / serial_ms / 1e6 / peak, "note": "one clip at a time, one stream"},
                "stages": stages}

    # end to end: host buffers in, host buffers out, through the public host API
    e2e = None
    if not args.no_e2e:
        host = HostElvisV1(T, HEIGHT, WIDTH, BLOCK, SHRINK, ALPHA, BETA, dev, depth=2)
        i420 = torch.empty((T, HEIGHT * WIDTH * 3 // 2), dtype=torch.uint8, pin_memory=True)
        src = Yuv420.from_i420(i420, HEIGHT, WIDTH)
        for a, b in zip(src.planes, clip.planes):
            a.copy_(b)
        outs = [host.host_buffers(pinned=True) for _ in range(2)]
        torch.cuda.synchronize()
        for i in range(2):
            host.process(i420, *outs[i % 2])
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            host.process(i420, *outs[i % 2])
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        e2e = {"value": T * world * args.steps / sec, "unit": UNIT, "h2d_bytes_per_step": host.h2d_bytes * world,
               "d2h_bytes_per_step": host.d2h_bytes * world, "ms_per_step": sec / args.steps * 1e3,
               "api": "elvis_b200.pipeline.HostElvisV1.process (pinned host I420 in; masks + shrunk + stretched I420 out; "
                      "each rank scores its own clip on this leg)"}
        del host, outs, i420

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu:
            cb = cpu_arm(16, 2, 1)
            cb = {k_: cb[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
        launches_per_step = 2 + 2 + 1 + 1 + 1 + 1   # score(init+kernel) combine(init+kernel) normalize select shrink(YUV fused) stretch(YUV fused)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8 pixels, f32 DCT, f64 scores", "data": "synthetic",
                "config": config_dict(world, T), "roofline": roofline, "cpu_baseline": cb, "e2e": e2e,
                "clocks": clocks, "gpu_launches": launches_per_step * args.steps}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per GPU (default: the 120 of the headline config)")
    ap.add_argument("--serial", action="store_true", help="time one clip at a time instead of the two-stream pipeline")
    ap.add_argument("--depth", type=int, default=None, help="clips in flight in the stream pipeline (default: 3 on one GPU, 2 when sharded)")
    ap.add_argument("--move-ctas", type=int, default=3, help="shrink/stretch CTAs per SM while pipelined")
    ap.add_argument("--split-stretch", dest="split_stretch", action="store_true", default=None,
                    help="three pipeline stages (score | shrink | stretch); default on one GPU")
    ap.add_argument("--no-split-stretch", dest="split_stretch", action="store_false",
                    help="two pipeline stages (score | shrink+stretch); default when sharded: the NCCL kernels of the "
                         "halo exchange and the all-reduces need SM slots that three resident kernels do not leave")
    ap.add_argument("--stretch-ctas", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its
    # version banner to stdout when NCCL_DEBUG is set on the box), so everything but our own
    # line goes to stderr: fd 1 is pointed at fd 2 and the line is written to the saved fd.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
