#!/usr/bin/env python
"""Headline benchmark: 4K frames/s for score + shrink + stretch (BASELINE.json), plus the other
BASELINE.json configurations as selectable workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload v1|v2_dampen|v2_downsample|v2_blur|8k600]

Workloads (config.workload names the one a line was measured on)
  v1 (default)   BASELINE.json configs[1]: ELVIS v1 on synthetic planar YUV 4:2:0 clips, 3840x2160, 120
                 frames per GPU, 16x16 blocks, 50 % removal, alpha = beta = 0.5.  One clip takes the whole
                 hot path: SC/TC scoring (8x8 DCT) -> elvis combine / normalise -> per-row top-k mask ->
                 shrink -> stretch.  A STEP is a batch of --clips-per-step such clips (a stream of GOPs,
                 cycling over --distinct different clips resident in HBM), so that the K timed steps of
                 the driver's default invocation span seconds, not milliseconds.  With N > 1 (torchrun,
                 one rank per GPU) every clip is ONE 120*N-frame clip owned in consecutive 120-frame ranges:
                 a one-frame luma halo is exchanged with the neighbours and the two global min / max
                 normalisations are reduced across the ranks (weak scaling).
  v2_dampen      configs[2]: scoring + removability + per-block DCT dampening (strength map = removability).
  v2_downsample  configs[3]: adaptive downsample, 2-bit 1x/2x/4x/8x level map (levels + degrade + pack).
  v2_blur        configs[3], Gaussian-blur variant: rounds 0..10 (elvis.py:2171-2196).
  8k600          configs[4]: ONE 7680x4320 600-frame clip, frame ranges over the N ranks (strong scaling).
The default v1 line also carries short runs of the other four under "workloads" (skip: --no-extra).

value  = frames/s with the inputs resident in HBM (CUDA events on the launching streams, max over ranks)
e2e    = frames/s through the host API (pinned host I420 in, host outputs back; copies inside the region)
--impl reference = the CPU arm: the oracle port of the same workload on the host cores.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK, SHRINK, ALPHA, BETA = 16, 0.5, 0.5, 0.5
UNIT = "frames/s"
SPECS = {
    "v1": {"metric": "4K frames/s, score+shrink+stretch", "w": 3840, "h": 2160, "frames": 120, "scaling": "weak",
           "baseline_config": "configs[1]"},
    "v2_dampen": {"metric": "4K frames/s, v2 scoring + DCT dampening", "w": 3840, "h": 2160, "frames": 120, "scaling": "weak",
                  "baseline_config": "configs[2]"},
    "v2_downsample": {"metric": "4K frames/s, v2 adaptive downsample (2-bit level map)", "w": 3840, "h": 2160, "frames": 120,
                      "scaling": "weak", "baseline_config": "configs[3]"},
    "v2_blur": {"metric": "4K frames/s, v2 per-block Gaussian blur (rounds 0..10)", "w": 3840, "h": 2160, "frames": 120,
                "scaling": "weak", "baseline_config": "configs[3] (blur variant)"},
    "8k600": {"metric": "8K frames/s, score+shrink+stretch, 600-frame clip frame-sharded", "w": 7680, "h": 4320, "frames": 600,
              "scaling": "strong", "baseline_config": "configs[4]"},
}


def v1_bytes_per_frame(w, h, s=SHRINK):
    """score reads Y once; shrink reads+writes the kept blocks; stretch reads the kept
    blocks and writes the full frame (YUV 4:2:0 = 1.5 bytes/pixel)."""
    score = w * h
    shrink = 2 * (1 - s) * 1.5 * w * h
    stretch = (1 - s) * 1.5 * w * h + 1.5 * w * h
    return {"score": score, "shrink": shrink, "stretch": stretch, "total": score + shrink + stretch}


def v2_bytes_per_frame(w, h, with_scoring):
    """a v2 degradation reads and writes the full YUV 4:2:0 frame = 3 W H; + W H when the
    scoring pass runs in the same step."""
    return 3 * w * h + (w * h if with_scoring else 0)


def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        peaks = {}
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, read from the tracked ncu
    summary profiles/ncu_traffic.csv (columns: kernel, frames, dram_bytes_read, dram_bytes_write, source)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.csv")
    try:
        with open(path) as f:
            for row in csv.DictReader(f):
                if kernel_substr in row["kernel"]:
                    return {"bytes": float(row["dram_bytes_read"]) + float(row["dram_bytes_write"]), "frames": int(row["frames"]),
                            "source": "profiles/ncu_traffic.csv <- " + row["source"]}
    except (OSError, KeyError, ValueError):
        pass
    return None


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
        """Drop the samples taken so far (warm-up); called right before a timed region."""
        self.sm, self.bits = [], 0

    def snapshot(self):
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm, bits = list(self.sm), self.bits
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for n, b in self.REASONS.items() if bits & b), "samples": len(sm)}

    def stop(self):
        if self.thread:
            self._stop = True
            self.thread.join()


# ------------------------------------------------------------------------------ CPU arm
def cpu_v1(width, height, sample_frames, steps, warmup, keep=False):
    """Oracle port of the v1 path on the host cores, on the first `sample_frames` frames of the workload."""
    from elvis_b200.synth import synth_yuv420
    from oracle.cpu_baseline import CpuElvisV1
    cores = os.cpu_count() or 1
    clip = synth_yuv420(sample_frames, height, width, seed=1234, device="cpu")
    cpu = CpuElvisV1(clip.y.numpy(), clip.u.numpy(), clip.v.numpy(), BLOCK, SHRINK, ALPHA, BETA, workers=cores)
    try:
        for _ in range(warmup):
            cpu.step()
        times = [cpu.step() for _ in range(steps)]
        outputs = {k: v.copy() for k, v in cpu.outputs().items()} if keep else None
    finally:
        cpu.close()
    sec = sum(times) / len(times)
    res = {"value": sample_frames / sec, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"first {sample_frames} frames of the {width}x{height} workload per step, {steps} steps after {warmup} warm-up; "
                     f"oracle/cpu_baseline.py CpuElvisV1 (NumPy/SciPy port of spec scoring + elvis.py:1160-1220, 1387-1455, "
                     f"fork pool over {cores} cores)",
           "ms_per_step": sec * 1e3}
    return res, clip, outputs


def v2_map_from_scores(kind, scores):
    """Block map of a v2 workload from removability scores in [0, 1] (numpy float64)."""
    import numpy as np
    if kind == "blur":
        return np.round(scores * 10).astype(np.int32)                                  # elvis.py:2176
    if kind == "downsample":
        return np.minimum(np.round(scores * 4).astype(np.int32), 3)                    # elvis.py:2146, clamped to 2 bits
    return scores.astype(np.float32)                                                   # dampening strength


def cpu_v2(kind, width, height, sample_frames, steps, warmup, keep=False):
    """Oracle port of one v2 degradation (+ scoring for the dampening workload) on the host cores."""
    import numpy as np
    from elvis_b200.synth import synth_yuv420
    from oracle import ref_port as P
    from oracle import spec_scoring
    from oracle.cpu_baseline import CpuV2
    cores = os.cpu_count() or 1
    clip = synth_yuv420(sample_frames, height, width, seed=1234, device="cpu")
    y, u, v = clip.y.numpy(), clip.u.numpy(), clip.v.numpy()
    t0 = time.perf_counter()
    sc, tc = spec_scoring.sc_tc(y, BLOCK)
    scores = P.combine_removability(sc, tc, ALPHA, BETA)
    score_sec = time.perf_counter() - t0      # single-process; only the dampening workload's step includes it
    cpu = CpuV2(y, u, v, v2_map_from_scores(kind, scores), BLOCK, kind, workers=cores)
    try:
        for _ in range(warmup):
            cpu.step()
        times = [cpu.step() for _ in range(steps)]
        outputs = {k: v_.copy() for k, v_ in cpu.outputs().items()} if keep else None
    finally:
        cpu.close()
    sec = sum(times) / len(times) + (score_sec if kind == "dampen" else 0.0)
    res = {"value": sample_frames / sec, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"first {sample_frames} frames of the {width}x{height} workload per step, {steps} steps after {warmup} warm-up; "
                     f"oracle/cpu_baseline.py CpuV2({kind}) (restated cv2 fixed point / spec, fork pool over {cores} cores)"
                     + ("; scoring (single process) included" if kind == "dampen" else ""),
           "ms_per_step": sec * 1e3}
    return res, clip, scores, outputs


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    spec = SPECS[args.workload]
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    if args.workload in ("v1", "8k600"):
        sample = 16 if args.workload == "v1" else 4
        cb, _, _ = cpu_v1(spec["w"], spec["h"], sample, steps, warmup)
    else:
        kind = args.workload[3:]
        cb, _, _, _ = cpu_v2(kind, spec["w"], spec["h"], 2 if kind == "blur" else 4, steps, warmup)
    line = {"impl": "reference", "metric": spec["metric"], "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": spec["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, args.gpus, spec["frames"] if spec["scaling"] == "weak" else None, args),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def config_dict(workload, n_gpus, frames_per_gpu, args):
    spec = SPECS[workload]
    w, h = spec["w"], spec["h"]
    d = {"width": w, "height": h, "block_size": BLOCK, "dct_size": 8, "baseline_config": f"BASELINE.json {spec['baseline_config']}",
         "l2": "inputs larger than L2 (every clip >= 1.49 GB vs 126 MB L2), distinct clips in flight; no explicit flush"}
    if workload in ("v1", "8k600"):
        d["shrink_amount"] = SHRINK
    if workload == "v1":
        d["workload"] = (f"ELVIS v1 score+shrink+stretch, synthetic planar YUV420 {w}x{h}, 120 frames per GPU and clip, {BLOCK}x{BLOCK} "
                         f"blocks, 8x8 DCT, {int(SHRINK * 100)}% removal, alpha={ALPHA} beta={BETA} (BASELINE.json configs[1])")
        d.update(frames_per_gpu=frames_per_gpu, clips_per_step=args.clips_per_step, distinct_clips=args.distinct,
                 sharding=f"contiguous frame ranges x{n_gpus}, 1-frame luma halo")
    elif workload == "8k600":
        d["workload"] = (f"ELVIS v1 score+shrink+stretch on ONE synthetic planar YUV420 {w}x{h} clip of {spec['frames']} frames, "
                         f"{BLOCK}x{BLOCK} blocks, {int(SHRINK * 100)}% removal, frame ranges over {n_gpus} GPU(s) with a 1-frame luma "
                         f"halo (BASELINE.json configs[4])")
        d.update(frames_total=spec["frames"], sharding=f"sharding.frame_range x{n_gpus} (elvis.py:264-278)")
    else:
        what = {"v2_dampen": "SC/TC scoring + removability + per-block DCT dampening (strength map = removability)",
                "v2_downsample": "levels_from_scores + adaptive downsample with the 2-bit 1x/2x/4x/8x level map + 2-bit packing",
                "v2_blur": "levels_from_scores + per-block Gaussian blur, rounds 0..10 (elvis.py:2171-2196)"}[workload]
        d["workload"] = f"ELVIS v2 {what}, synthetic planar YUV420 {w}x{h}, 120 frames per GPU ({d['baseline_config']})"
        d.update(frames_per_gpu=frames_per_gpu, clips_per_step=1, sharding="independent frame ranges, no communication")
    return d


# ------------------------------------------------------------------------------ GPU arm: helpers
class Ctx:
    """Per-process state of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL's internal streams at high priority: its small kernels then take the first SM slot that
            # frees up instead of queueing behind the move kernels
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.peak, self.peak_src = load_peaks()
        self.sampler = ClockSampler(self.local) if self.rank == 0 else None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, run, steps, warmup):
        """W untimed steps, then exactly K steps between barriers, CUDA events on the current stream,
        max over ranks.  Returns (total ms, clocks sampled during the region)."""
        torch = self.torch
        for _ in range(warmup):
            run()
        self.barrier()
        if self.sampler:
            self.sampler.reset()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            run()
        stop.record()
        self.barrier()
        clocks = self.sampler.snapshot() if self.sampler else None
        return self.max_over_ranks(start.elapsed_time(stop)), clocks

    def time_kernel(self, fn, reps):
        """Average duration of `reps` back-to-back launches of one kernel, timed ALONE (CUDA events on its stream)
        after a short idle gap, so that it is a burst figure like the measured peak it is divided by -- not the tail of
        the power-capped sustained region that ran just before.  The SM clock during the launches is kept in
        self.kernel_clocks and reported beside the figure."""
        torch = self.torch
        torch.cuda.synchronize()
        time.sleep(0.3)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        if self.sampler:
            self.sampler.reset()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        self.kernel_clocks = self.sampler.snapshot() if self.sampler else None
        return a.elapsed_time(b) / reps


def roofline_dict(ctx, kernel, alg_bytes, ms, frames, traffic_key):
    achieved = alg_bytes / ms / 1e6
    tr = ncu_traffic(traffic_key)
    traffic = None
    if tr:      # scale the captured launch to this launch's frame count (traffic is linear in frames)
        traffic = tr["bytes"] * frames / tr["frames"]
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": ctx.peak, "unit": "GB/s", "frac": achieved / ctx.peak,
            "traffic": traffic, "traffic_source": tr["source"] if tr else None, "peak_source": ctx.peak_src,
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": ms,
            "launch_timing": "kernel alone, back-to-back launches after a 0.3 s idle gap (burst conditions, like the peak)",
            "launch_clocks": getattr(ctx, "kernel_clocks", None)}


# ------------------------------------------------------------------------------ GPU arm: v1 / 8k600
def measure_v1(ctx, workload, steps, warmup, with_e2e, with_cpu, with_parity):
    """The v1 path on `frames_local` frames per rank; returns the pieces of a JSON line."""
    torch, dist, args = ctx.torch, ctx.dist, ctx.args
    from elvis_b200 import ops, sharding
    from elvis_b200.pipeline import ElvisV1, ElvisV1Pipelined, HostElvisClient, HostElvisV1
    from elvis_b200.synth import synth_yuv420
    from elvis_b200.yuv import Yuv420
    spec = SPECS[workload]
    W, H = spec["w"], spec["h"]
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    strong = spec["scaling"] == "strong"
    if strong:
        sharding.check_shardable(spec["frames"], world)
        a, b = sharding.frame_range(spec["frames"], rank, world)
        T, offset, total = b - a, a, spec["frames"]
        clips_per_step, distinct = 1, 1
    else:
        T = args.frames
        offset, total = rank * T, T * world
        clips_per_step, distinct = args.clips_per_step, max(1, args.distinct)
    pipe = ElvisV1(BLOCK, SHRINK, ALPHA, BETA)
    by, bx = H // BLOCK, W // BLOCK
    k = int(SHRINK * bx)

    # N > 1: halo exchange and min / max reductions over peer memory (copy engines + mailbox all-reduce over NVLink,
    # elvis_b200/peer.py) when CUDA IPC works between the ranks, else over NCCL
    pg = None
    if world > 1 and args.transport != "nccl":
        from elvis_b200 import peer
        pg = peer.try_create(rank, world, dev)
        if pg is None and args.transport == "peer":
            raise RuntimeError("--transport peer: the peer-memory setup failed")
    transport_name = "peer memory: copy-engine halo copies + mailbox all-reduce over NVLink (elvis_b200.peer)" if pg else \
        ("NCCL send/recv + all_reduce (torch.distributed)" if world > 1 else "none (one GPU)")

    # `distinct` different global clips; rank r owns frames [offset, offset + T) of each; luma lives in a halo buffer
    halos, clips = [], []
    for j in range(distinct):
        halo = pg.halo_clip(T, H, W) if pg else sharding.HaloClip(T, H, W, dev)
        chroma = torch.empty((2, T, H // 2, W // 2), dtype=torch.uint8, device=dev)
        clip = Yuv420(halo.owned, chroma[0], chroma[1])
        synth_yuv420(T, H, W, seed=1234 + 17 * j, device=dev, out=clip, frame_offset=offset)
        halos.append(halo)
        clips.append(clip)
    shrunk = Yuv420.empty(T, H, (bx - k) * BLOCK, dev)
    full = Yuv420.empty(T, H, W, dev)

    def score_of(j):
        if world > 1:
            return sharding.sharded_removability(halos[j], total, BLOCK, ALPHA, BETA, rank, world, transport=pg)
        return pipe.score(clips[j])

    def serial_step(j=0, ev=None):
        if ev:
            ev[0].record()
        scores = score_of(j)
        if ev:
            ev[1].record()
        _, mask = pipe.shrink(clips[j], scores, shrunk)
        if ev:
            ev[2].record()
        pipe.stretch(shrunk, mask, full)
        if ev:
            ev[3].record()
        return scores, mask

    # (1) serial passes: one clip at a time on one stream -- per-stage breakdown and clip latency
    n_serial = max(3, min(steps, 10))
    for _ in range(3):
        serial_step()
    ctx.barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_serial)]
    for i in range(n_serial):
        serial_step(i % distinct, evs[i])
    ctx.barrier()
    stage_ms = [sum(e[j].elapsed_time(e[j + 1]) for e in evs) / n_serial for j in range(3)]
    serial_ms = sum(e[0].elapsed_time(e[3]) for e in evs) / n_serial

    # (2) the timed region.  4K: clips through the stream pipeline (every clip takes the full serial path);
    # 8K x 600: one clip at a time (a clip is tens of milliseconds of work, nothing to overlap)
    index_of = {id(c): j for j, c in enumerate(clips)}
    pipelined = not strong and not args.serial
    if pipelined:
        score_fn = comm_fn = None
        if world > 1:   # halo exchange ahead of time on the communication stream, reductions inside the score stage
            if pg:
                comm_fn = lambda c: pg.exchange_halo(halos[index_of[id(c)]])  # noqa: E731
            else:
                comm_fn = lambda c: sharding.exchange_halo(halos[index_of[id(c)]], rank, world)  # noqa: E731
            score_fn = lambda c, slot: sharding.sharded_removability(halos[index_of[id(c)]], total, BLOCK, ALPHA, BETA, rank,  # noqa: E731
                                                                     world, exchange=False, transport=pg)
        # three stages on one GPU and with the peer transport; two with NCCL (its SM-resident kernels wait for SM slots
        # next to three resident compute kernels -- measured in round 1)
        split = (world == 1 or pg is not None) if args.split_stretch is None else args.split_stretch
        depth = (3 if split else 2) if args.depth is None else args.depth
        pp = ElvisV1Pipelined(T, H, W, BLOCK, SHRINK, ALPHA, BETA, dev, depth=depth, score_fn=score_fn,
                              move_ctas_per_sm=args.move_ctas, comm_fn=comm_fn, split_stretch=split,
                              stretch_ctas_per_sm=args.stretch_ctas, prepared=not args.no_prepared)
        counter = [0]

        def run():
            for _ in range(clips_per_step):
                pp.submit(clips[counter[0] % distinct])
                counter[0] += 1
            pp.join()
        mode = f"{depth} clips in flight on {3 if split else 2} CUDA streams" + (" (+1 for the halo exchange)" if world > 1 else "") + \
               " (elvis_b200.pipeline.ElvisV1Pipelined); every clip takes the full serial path"
    else:
        counter = [0]

        def run():
            for _ in range(clips_per_step):
                serial_step(counter[0] % distinct)
                counter[0] += 1
        mode = "one clip at a time on one stream"
    ms, clocks = ctx.timed(run, steps, warmup)
    ms_per_step = ms / steps
    # host time spent enqueueing one clip, measured on an EMPTY launch queue (inside the timed region the host runs ahead
    # until the queue is full and then blocks, so its wall time there says nothing about its own cost): it only matters
    # when it approaches the device time per clip
    host_enqueue_ms = None
    if pipelined:
        ctx.barrier()
        n_probe = 16
        t_h = time.perf_counter()
        for i in range(n_probe):
            pp.submit(clips[i % distinct])
        host_enqueue_ms = (time.perf_counter() - t_h) / n_probe * 1e3
        pp.join()
        ctx.barrier()
    frames_per_step = total * clips_per_step
    value = frames_per_step * steps / (ms / 1e3)

    # (3) dominant kernel alone: the scoring kernel over this rank's frames (CUDA events, same stream)
    ab = v1_bytes_per_frame(W, H)
    sc_ms = ctx.time_kernel(lambda: ops.score_sc_tc(clips[0].y, BLOCK), max(3, min(steps, 20)))
    roof = roofline_dict(ctx, "score_umma_kernel<2> (elvis_score_sc_tc, tcgen05)", ab["score"] * T, sc_ms, T, "score_umma_kernel")
    per_clip_ms = ms_per_step / clips_per_step
    roof["whole_step"] = {"achieved": ab["total"] * T / per_clip_ms / 1e6, "frac": ab["total"] * T / per_clip_ms / 1e6 / ctx.peak,
                          "ms_per_clip": per_clip_ms, "host_enqueue_ms_per_clip": host_enqueue_ms, "note": "all three stages as timed (" + mode + "), per rank"}
    roof["serial_step"] = {"ms": serial_ms, "achieved": ab["total"] * T / serial_ms / 1e6,
                           "frac": ab["total"] * T / serial_ms / 1e6 / ctx.peak, "note": "one clip at a time, one stream"}
    roof["stages"] = {"score_pipeline": {"ms": stage_ms[0], "gbs": ab["score"] * T / stage_ms[0] / 1e6},
                      "shrink": {"ms": stage_ms[1], "gbs": ab["shrink"] * T / stage_ms[1] / 1e6},
                      "stretch": {"ms": stage_ms[2], "gbs": ab["stretch"] * T / stage_ms[2] / 1e6}}

    # (4) N > 1: the gathered sharded scores and masks must be bit-identical to one rank scoring the whole clip
    sharded_equal = None
    if world > 1:
        scores, mask = serial_step(0)
        sizes = [sharding.frame_range(total, r, world) if strong else (r * T, (r + 1) * T) for r in range(world)]
        tmax = max(b_ - a_ for a_, b_ in sizes)

        def gather(x):       # (T, ...) -> on rank 0 the (total, ...) concatenation
            pad = torch.zeros((tmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)
            pad[:x.shape[0]] = x
            out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, out, dst=0)
            if rank != 0:
                return None
            return torch.cat([o[:b_ - a_] for o, (a_, b_) in zip(out, sizes)])
        all_y, all_scores, all_mask = gather(clips[0].y), gather(scores), gather(mask)
        if rank == 0:
            single = pipe.score(Yuv420(all_y, None, None))
            single_mask = ops.select_rows(single, k, ops.REMOVE_HIGH)
            sharded_equal = bool(torch.equal(single, all_scores) and torch.equal(single_mask, all_mask))
            del single, single_mask
        del all_y, all_scores, all_mask
        ctx.barrier()

    # (5) end to end: host buffers in, host buffers out, through the public host API
    e2e = None
    if with_e2e:
        e2e = e2e_v1(ctx, clips[0], halos, T, H, W, total, frames_per_clip=total, strong=strong)

    # (6) CPU arm on a bounded sample + parity of the GPU path on exactly those frames
    cb = parity = None
    if rank == 0 and (with_cpu or with_parity):
        sample = 16 if not strong else 4
        cb_full, cpu_clip, cpu_out = cpu_v1(W, H, sample, 2, 1, keep=with_parity)
        cb = {k_: cb_full[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
        if with_parity:
            from oracle import verify
            gclip = Yuv420(*(p.to(dev) for p in cpu_clip.planes))
            g_scores, g_mask, g_sh, g_fu = pipe.run(gclip)
            torch.cuda.synchronize()
            gpu_out = {"scores": g_scores.cpu().numpy(), "mask": g_mask.cpu().numpy()}
            for tag, s_, f_ in (("y", g_sh.y, g_fu.y), ("u", g_sh.u, g_fu.u), ("v", g_sh.v, g_fu.v)):
                gpu_out["s" + tag], gpu_out["f" + tag] = s_.cpu().numpy(), f_.cpu().numpy()
            parity = verify.compare_v1(gpu_out, cpu_out, BLOCK, k)
            parity["checked_against"] = "oracle/cpu_baseline.py CpuElvisV1.outputs() on the CPU arm's frames"
    # score(init+kernel) combine(init+kernel) [normalize: its own launch in the sharded scorer and the serial path, else folded into] select
    # shrink(YUV fused) stretch(YUV fused)
    launches_per_clip = 2 + 2 + (1 if world > 1 or not pipelined else 0) + 1 + 1 + 1
    if pg:
        launches_per_clip += 2 + 4 + 2 + 2      # halo: 2 ack waits + 2 flag stores (the copies are copy-engine work), 2 arrival waits, 2 acks; 2 all-reduces
        pg.check()
    del clips, halos, shrunk, full
    if pipelined:
        del pp
    if pg:
        pg.close()
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": ms_per_step, "frames_per_step": frames_per_step, "frames_per_gpu": T, "clocks": clocks,
            "roofline": roof, "e2e": e2e, "cpu_baseline": cb, "parity_check": parity, "sharded_equals_single": sharded_equal,
            "gpu_launches": launches_per_clip * clips_per_step * steps, "timed_region_s": ms / 1e3, "mode": mode,
            "transport": transport_name}


def e2e_v1(ctx, clip, halos, T, H, W, total, frames_per_clip, strong):
    """frames/s through HostElvisV1 (composite: masks + shrunk + stretched out), plus the reference's
    server leg (masks bit-packed + shrunk out) and client leg (shrunk + masks in, stretched out)."""
    torch, args = ctx.torch, ctx.args
    from elvis_b200 import sharding
    from elvis_b200.pipeline import HostElvisClient, HostElvisV1
    from elvis_b200.yuv import Yuv420
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    light = T * H * W * 3 // 2 > 8e9       # a 30 GB clip: one buffer set, one stream, composite leg only (pinning takes long)
    n = 2 if light else max(2, args.e2e_clips)
    nbuf = 1 if light else 2
    score_fn = None
    if world > 1:        # the product path of a sharded job: halo exchange + global min / max across the ranks
        score_fn = lambda halo, index: sharding.sharded_removability(halo, total, BLOCK, ALPHA, BETA, rank, world)  # noqa: E731
    i420 = torch.empty((T, H * W * 3 // 2), dtype=torch.uint8, pin_memory=True)
    for a_, b_ in zip(Yuv420.from_i420(i420, H, W).planes, clip.planes):
        a_.copy_(b_)
    torch.cuda.synchronize()

    def leg(host, call, reps):
        for i in range(nbuf):
            call(i)
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(reps):
            call(i)
        torch.cuda.synchronize()
        sec = ctx.max_over_ranks(time.perf_counter() - t0)
        return {"value": frames_per_clip * reps / sec, "unit": UNIT, "ms_per_clip": sec / reps * 1e3,
                "h2d_bytes_per_clip": host.h2d_bytes * world, "d2h_bytes_per_clip": host.d2h_bytes * world}

    # composite (what the CPU arm computes too): clip in; masks + shrunk + stretched out
    host = HostElvisV1(T, H, W, BLOCK, SHRINK, ALPHA, BETA, dev, depth=nbuf, score_fn=score_fn)
    outs = [host.host_buffers(pinned=True) for _ in range(nbuf)]
    comp = leg(host, lambda i: host.process(i420, *outs[i % nbuf]), n)
    del host
    composite = {"value": comp["value"], "unit": UNIT, "h2d_bytes_per_step": comp["h2d_bytes_per_clip"] * n,
                 "d2h_bytes_per_step": comp["d2h_bytes_per_clip"] * n, "clips_per_step": n, "ms_per_step": comp["ms_per_clip"] * n,
                 "api": "elvis_b200.pipeline.HostElvisV1.process (pinned host I420 in; masks + shrunk + stretched I420 out)"
                        + ("; scoring through sharding.sharded_removability (halo exchange + global min/max over the ranks)" if world > 1 else "")}
    if light:
        del outs, i420
        torch.cuda.empty_cache()
        return composite
    # server leg: clip in; bit-packed masks + shrunk out (elvis.py:4389-4418)
    server = HostElvisV1(T, H, W, BLOCK, SHRINK, ALPHA, BETA, dev, depth=2, outputs=("mask", "shrunk"), pack_masks=True,
                         score_fn=score_fn)
    souts = [server.host_buffers(pinned=True) for _ in range(2)]
    srv = leg(server, lambda i: server.process(i420, *souts[i % 2]), n)
    del server
    # client leg: shrunk + bit-packed masks in; stretched out (elvis.py:4537-4557)
    client = HostElvisClient(T, H, W, BLOCK, SHRINK, dev, depth=2)
    full_out = [outs[0][1], outs[1][1]]
    cli = leg(client, lambda i: client.process(souts[0][0], souts[0][2], full_out[i % 2]), n)
    del client, outs, souts, full_out, i420
    torch.cuda.empty_cache()
    composite["server_leg"] = dict(srv, api="HostElvisV1(outputs=('mask','shrunk'), pack_masks=True): what the reference's server ships")
    composite["client_leg"] = dict(cli, api="HostElvisClient: shrunk I420 + packed masks in, stretched I420 out")
    return composite


# ------------------------------------------------------------------------------ GPU arm: v2 degradations
def measure_v2(ctx, workload, steps, warmup, with_e2e, with_cpu, with_parity):
    torch, args = ctx.torch, ctx.args
    from elvis_b200 import ops
    from elvis_b200.pipeline import ElvisV1, PresleyV2
    from elvis_b200.synth import synth_yuv420
    from elvis_b200.yuv import Yuv420
    kind = workload[3:]
    spec = SPECS[workload]
    W, H, T = spec["w"], spec["h"], args.frames
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    pipe, v2 = ElvisV1(BLOCK, SHRINK, ALPHA, BETA), PresleyV2(BLOCK)
    distinct = 2
    clips = [synth_yuv420(T, H, W, seed=1234 + 17 * j, device=dev, frame_offset=rank * T) for j in range(distinct)]
    outs = [Yuv420.empty(T, H, W, dev) for _ in range(distinct)]
    scores = [pipe.score(c) for c in clips]

    def step_on(clip, sc, out):
        """One pass of the workload over a resident clip -> (degraded clip, block map[, packed map])."""
        if kind == "dampen":
            s = pipe.score(clip)
            return v2.dampen(clip, s.float(), out), s
        if kind == "downsample":
            # elvis.py:2146 at bs 16 gives levels 0..4; the 2-bit map keeps 0..3: the degradation kernel and the packer
            # both clamp to their top level
            lv = ops.levels_from_scores(sc, ops.LEVELS_ROUND, 4)
            return v2.downsample_pow2(clip, lv, 3, out), ops.pack_levels_2bit(lv)
        rounds = ops.levels_from_scores(sc, ops.LEVELS_ROUND, 10)                # elvis.py:2176
        return v2.blur(clip, rounds, out), rounds
    counter = [0]

    def run():
        j = counter[0] % distinct
        counter[0] += 1
        step_on(clips[j], scores[j], outs[j])
    ms, clocks = ctx.timed(run, steps, warmup)
    ms_per_step = ms / steps
    value = T * world * steps / (ms / 1e3)

    # dominant kernel alone: the luma launch of the degradation (reads + writes the Y plane of every frame); the
    # power-of-two downsample handles Y, U and V in ONE launch (reads + writes the whole 4:2:0 frame)
    y_out = outs[0].y
    k_bytes = 2 * W * H * T
    if kind == "dampen":
        strength = scores[0].float()
        fn, name, key = (lambda: ops.dct_dampen(clips[0].y, strength, BLOCK, out=y_out)), \
            "dampen_packed_kernel (elvis_dct_dampen, luma launch; packed fp32 AAN butterflies)", "dampen_packed"
    elif kind == "downsample":
        lv = ops.levels_from_scores(scores[0], ops.LEVELS_ROUND, 4)
        fn, name, key = (lambda: v2.downsample_pow2(clips[0], lv, 3, outs[0])), \
            "downsample_pow2_yuv420_tma_kernel (elvis_degrade_downsample_pow2_yuv420, Y+U+V in one launch, TMA tiles)", "downsample_pow2_yuv420"
        k_bytes = 3 * W * H * T
    else:
        rounds = ops.levels_from_scores(scores[0], ops.LEVELS_ROUND, 10)
        fn, name, key = (lambda: ops.degrade_blur(clips[0].y, rounds, BLOCK, out=y_out)), \
            "blur_imma_tma_kernel<16> (elvis_degrade_blur, luma launch; mma.sync u8, TMA strips)", "blur_imma"
    k_ms = ctx.time_kernel(fn, max(3, min(steps, 20)))
    roof = roofline_dict(ctx, name, k_bytes, k_ms, T, key)
    alg = v2_bytes_per_frame(W, H, kind == "dampen") * T
    roof["whole_step"] = {"achieved": alg / ms_per_step / 1e6, "frac": alg / ms_per_step / 1e6 / ctx.peak,
                          "algorithmic_bytes": alg, "note": "the whole step as timed (Y, U and V launches + maps), per rank"}

    e2e = None
    if with_e2e:
        i420 = torch.empty((T, H * W * 3 // 2), dtype=torch.uint8, pin_memory=True)
        for a_, b_ in zip(Yuv420.from_i420(i420, H, W).planes, clips[0].planes):
            a_.copy_(b_)
        host_out = [torch.empty((T, H * W * 3 // 2), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        dbuf = [torch.empty((T, H * W * 3 // 2), dtype=torch.uint8, device=dev) for _ in range(2)]
        dout = [torch.empty((T, H * W * 3 // 2), dtype=torch.uint8, device=dev) for _ in range(2)]
        streams = [torch.cuda.Stream(dev) for _ in range(2)]
        map_host = [None, None]

        def call(i):
            j = i % 2
            with torch.cuda.stream(streams[j]):
                dbuf[j].copy_(i420, non_blocking=True)
                clip = Yuv420.from_i420(dbuf[j], H, W)
                sc = scores[0] if kind == "dampen" else pipe.score(clip)      # the host API scores the clip it was handed
                _, m = step_on(clip, sc, Yuv420.from_i420(dout[j], H, W))
                host_out[j].copy_(dout[j], non_blocking=True)
                if map_host[j] is None:
                    map_host[j] = torch.empty(m.shape, dtype=m.dtype, pin_memory=True)
                map_host[j].copy_(m, non_blocking=True)
        n = max(2, args.e2e_clips)
        for i in range(2):
            call(i)
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(n):
            call(i)
        torch.cuda.synchronize()
        sec = ctx.max_over_ranks(time.perf_counter() - t0)
        mbytes = map_host[0].numel() * map_host[0].element_size()
        e2e = {"value": T * world * n / sec, "unit": UNIT, "h2d_bytes_per_step": T * H * W * 3 // 2 * world * n,
               "d2h_bytes_per_step": (T * H * W * 3 // 2 + mbytes) * world * n, "clips_per_step": n, "ms_per_step": sec * 1e3,
               "api": "pinned host I420 in -> ElvisV1.score -> PresleyV2." + {"dampen": "dampen", "downsample": "downsample_pow2", "blur": "blur"}[kind]
                      + " -> degraded I420 + block map back to pinned host memory (two streams)"}
        del i420, host_out, dbuf, dout
    cb = parity = None
    if rank == 0 and (with_cpu or with_parity):
        sample = 2 if kind == "blur" else 4
        cb_full, cpu_clip, cpu_scores, cpu_out = cpu_v2(kind, W, H, sample, 1, 0, keep=with_parity)
        cb = {k_: cb_full[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
        if with_parity:
            from oracle import verify
            gclip = Yuv420(*(p.to(dev) for p in cpu_clip.planes))
            # the block map is derived from the ORACLE's scores on both sides, so that the pixel work is compared on
            # identical maps (score parity is the v1 check's business)
            sc = torch.from_numpy(cpu_scores).to(dev)
            if kind == "dampen":
                got = v2.dampen(gclip, sc.float())
            else:
                got, _ = step_on(gclip, sc, None)
            torch.cuda.synchronize()
            parity = verify.compare_planes({"y": got.y.cpu().numpy(), "u": got.u.cpu().numpy(), "v": got.v.cpu().numpy()}, cpu_out,
                                           tol=1 if kind == "dampen" else 0)
            parity["frames"] = sample
            parity["checked_against"] = f"oracle/cpu_baseline.py CpuV2({kind}).outputs() on the CPU arm's frames"
    launches = {"dampen": 6 + 3, "downsample": 1 + 1 + 1, "blur": 1 + 3}[kind]   # scoring (6) / levels, degrade launches, packer
    del clips, outs, scores
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": ms_per_step, "frames_per_step": T * world, "frames_per_gpu": T, "clocks": clocks,
            "roofline": roof, "e2e": e2e, "cpu_baseline": cb, "parity_check": parity, "sharded_equals_single": None,
            "gpu_launches": launches * steps, "timed_region_s": ms / 1e3, "mode": "one clip per step, two distinct clips alternating"}


def measure(ctx, workload, steps, warmup, with_e2e=True, with_cpu=True, with_parity=True):
    if workload in ("v1", "8k600"):
        return measure_v1(ctx, workload, steps, warmup, with_e2e, with_cpu, with_parity)
    return measure_v2(ctx, workload, steps, warmup, with_e2e, with_cpu, with_parity)


def run_ours(args):
    ctx = Ctx(args)
    spec = SPECS[args.workload]
    world = ctx.world
    warmup = max(3, args.warmup)
    m = measure(ctx, args.workload, args.steps, warmup, with_e2e=not args.no_e2e, with_cpu=(world == 1 and not args.no_cpu),
                with_parity=not args.no_parity)
    extras = {}
    if args.workload == "v1" and not args.no_extra:
        # short runs of the other BASELINE.json configurations, same process, same clock sampler
        for wl in ("v2_dampen", "v2_downsample", "v2_blur", "8k600"):
            heavy = wl == "8k600"
            r = measure(ctx, wl, 3 if heavy else 10, 3, with_e2e=not args.no_e2e and not heavy,
                        with_cpu=(world == 1 and not args.no_cpu), with_parity=not args.no_parity and not heavy)
            extras[wl] = {"metric": SPECS[wl]["metric"], "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                          "steps": 3 if heavy else 10, "warmup": 3, "scaling": SPECS[wl]["scaling"],
                          "config": config_dict(wl, world, r["frames_per_gpu"], args), "roofline": r["roofline"], "e2e": r["e2e"],
                          "cpu_baseline": r["cpu_baseline"], "parity_check": r["parity_check"],
                          "sharded_equals_single": r["sharded_equals_single"], "clocks": r["clocks"], "gpu_launches": r["gpu_launches"]}
            if r.get("transport") and world > 1:
                extras[wl]["config"]["transport"] = r["transport"]
    if ctx.sampler:
        ctx.sampler.stop()
    if ctx.rank == 0:
        cfg = config_dict(args.workload, world, m["frames_per_gpu"], args)
        cfg["frames_per_step"] = m["frames_per_step"]
        cfg["pipelining"] = m["mode"]
        if m.get("transport"):
            cfg["transport"] = m["transport"]
        line = {"metric": spec["metric"], "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": spec["scaling"],
                "vs_baseline": None, "dtype": "u8 pixels, f32 DCT, f64 scores", "data": "synthetic", "config": cfg,
                "roofline": m["roofline"], "cpu_baseline": m["cpu_baseline"], "e2e": m["e2e"], "clocks": m["clocks"],
                "gpu_launches": m["gpu_launches"], "timed_region_s": m["timed_region_s"], "parity_check": m["parity_check"],
                "sharded_equals_single": m["sharded_equals_single"]}
        if extras:
            line["workloads"] = extras
        emit(line)
    if world > 1:
        ctx.dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="v1", choices=sorted(SPECS))
    ap.add_argument("--frames", type=int, default=120, help="frames per GPU and clip (default: the 120 of the 4K configurations)")
    ap.add_argument("--clips-per-step", type=int, default=100,
                    help="v1: clips per step (a step is a batch of clips, so that K = 20 steps time >= 2 s of GPU work)")
    ap.add_argument("--distinct", type=int, default=4, help="v1: number of different clips resident in HBM that the steps cycle over")
    ap.add_argument("--e2e-clips", type=int, default=8, help="clips per end-to-end measurement")
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: halo exchange + min/max reductions over peer memory (default when CUDA IPC works) or NCCL")
    ap.add_argument("--serial", action="store_true", help="time one clip at a time instead of the stream pipeline")
    ap.add_argument("--depth", type=int, default=None, help="clips in flight in the stream pipeline (default: 3 on one GPU, 2 when sharded)")
    ap.add_argument("--move-ctas", type=int, default=3, help="shrink/stretch CTAs per SM while pipelined")
    ap.add_argument("--split-stretch", dest="split_stretch", action="store_true", default=None,
                    help="three pipeline stages (score | shrink | stretch); default on one GPU")
    ap.add_argument("--no-split-stretch", dest="split_stretch", action="store_false",
                    help="two pipeline stages (score | shrink+stretch); default when sharded")
    ap.add_argument("--stretch-ctas", type=int, default=None)
    ap.add_argument("--no-prepared", action="store_true", help="pipeline: go through the ops wrappers for every clip instead of replaying prepared calls")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="v1: skip the short runs of the other workloads")
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
