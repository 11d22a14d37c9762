"""Batched planar-YUV 4:2:0 pipelines: the B200-native shape of the hot path.

The reference processes one packed BGR frame per Python call; here a whole clip stays in HBM
as three planes (the layout of the raw yuv420p file the reference feeds to EVCA,
elvis.py:4324-4330) and every stage is one kernel launch per plane over all frames.  The
removal mask / level map computed from luma is applied to the co-located chroma blocks at
half the block size (SURVEY.md appendix "Layout")."""
from __future__ import annotations

import ctypes as C
import os

from typing import Optional, Tuple

import torch

from . import _lib, ops
from .elvis import blocks_to_remove
from .yuv import Yuv420


class _Prepared:
    """One C-ABI call with its ctypes arguments built once.  The wrappers of ops.py validate and convert their
    arguments on every call (20-40 us of Python each); a clip that flows through the stream pipeline makes ten of
    them, which is most of the host time per clip (0.6-0.7 ms against ~1 ms of device time at 4K x 120).  The
    pipeline therefore validates a (clip, slot) pair once through the ops wrappers' own checks and afterwards replays
    the raw calls; the objects in `keep` pin every buffer and struct the raw pointers refer to."""
    __slots__ = ("name", "fn", "args", "keep")

    def __init__(self, name: str, args: tuple, keep: tuple = ()):
        self.name, self.fn, self.args, self.keep = name, getattr(_lib.lib, name), args, keep

    def __call__(self) -> None:
        rc = self.fn(*self.args)
        if rc:
            _lib.raise_for(self.name, rc)


def _fused_move_ok(clip: Yuv420, bs: int) -> bool:
    """Geometry the one-launch Y+U+V move kernels accept (elvis_shrink_yuv420 / elvis_stretch_yuv420)."""
    if bs % 16:
        return False
    for i, p in enumerate(clip.planes):
        al = 16 if i == 0 else 8
        if p.dim() != 3 or p.stride(2) != 1 or p.data_ptr() % al or p.stride(0) % al or p.stride(1) % al:
            return False
    return True


def move_planes(src: Yuv420, dst: Yuv420, mask: torch.Tensor, bs: int, small_bx: int, stretch_: bool,
                ctas_per_sm: int = 0) -> None:
    """Shrink or stretch the three planes: one fused launch when the geometry allows, else per plane."""
    if small_bx > 0 and ops.move_yuv420(src.planes, dst.planes, mask, bs, small_bx, stretch_, ctas_per_sm):
        return
    for s, d, pb in ((src.y, dst.y, bs), (src.u, dst.u, bs // 2), (src.v, dst.v, bs // 2)):
        if stretch_:
            ops.stretch(s, mask, pb, out=d, ctas_per_sm=ctas_per_sm)
        else:
            ops.shrink(s, mask, pb, small_bx, out=d, ctas_per_sm=ctas_per_sm)


class ElvisV1:
    """score -> per-row top-k mask -> shrink, and stretch (elvis.py:4350-4394, 4550-4557),
    for a planar clip resident on the GPU."""

    def __init__(self, block_size: int = 16, shrink_amount: float = 0.5, alpha: float = 0.5, beta: float = 0.5, dct_size: int = 8):
        """dct_size: 8 (north_star's 8 x 8 tiles, the benchmarked mode) or block_size (one transform per
        block, what the reference's EVCA call asks for; elvis.reference_dct_size)."""
        if block_size % 2:
            raise ValueError("4:2:0 chroma needs an even block size")
        self.bs, self.shrink_amount, self.alpha, self.beta, self.dct_size = block_size, shrink_amount, alpha, beta, dct_size

    def grid(self, clip: Yuv420) -> Tuple[int, int, int]:
        h, w = clip.y.shape[1:]
        if h % self.bs or w % self.bs:
            raise ValueError("Image dimensions must be divisible by block_size.")
        bx = w // self.bs
        return h // self.bs, bx, blocks_to_remove(self.shrink_amount, bx)

    def score(self, clip: Yuv420, background: Optional[torch.Tensor] = None) -> torch.Tensor:
        sc, tc, norm = ops.score_sc_tc(clip.y, self.bs, dct_size=self.dct_size)
        r, mm = ops.combine_removability(sc, tc, norm, self.alpha, self.beta, background)
        return ops.normalize_(r, mm)

    def shrink(self, clip: Yuv420, scores: torch.Tensor, out: Optional[Yuv420] = None):
        by, bx, k = self.grid(clip)
        mask = ops.select_rows(scores, k, ops.REMOVE_HIGH)
        if out is None:
            out = Yuv420.empty(clip.y.shape[0], by * self.bs, (bx - k) * self.bs, clip.y.device)
        move_planes(clip, out, mask, self.bs, bx - k, stretch_=False)
        return out, mask

    def stretch(self, shrunk: Yuv420, mask: torch.Tensor, out: Optional[Yuv420] = None) -> Yuv420:
        t, by, bx = mask.shape
        if out is None:
            out = Yuv420.empty(t, by * self.bs, bx * self.bs, shrunk.y.device)
        move_planes(shrunk, out, mask, self.bs, shrunk.y.shape[2] // self.bs, stretch_=True)
        return out

    def run(self, clip: Yuv420, background: Optional[torch.Tensor] = None,
            shrunk_out: Optional[Yuv420] = None, stretched_out: Optional[Yuv420] = None):
        """One pass of the headline path: returns (scores, mask, shrunk, stretched)."""
        scores = self.score(clip, background)
        shrunk, mask = self.shrink(clip, scores, shrunk_out)
        stretched = self.stretch(shrunk, mask, stretched_out)
        return scores, mask, shrunk, stretched


class ElvisV1Pipelined:
    """Throughput mode for a stream of clips (GOPs): `depth` clips in flight on two CUDA streams
    (score | shrink + stretch) or, with split_stretch, three (score | shrink | stretch; measured on
    B200 with depth 3: 1.014 ms per 120-frame 4K clip instead of 1.068).

    Scoring is bound by instruction issue and uses ~1/4 of the HBM bandwidth; shrink and stretch
    are pure data movement and use almost no issue slots.  Run back to back they leave one of the
    two resources idle at any time, so the scoring of clip i+1 is overlapped with the
    shrink + stretch of clip i: `submit` enqueues the score stage on one stream and the move
    stage on the other, ordered by events; each of the `depth` slots owns its score, mask,
    shrunk and stretched buffers.  Every clip still takes exactly the serial path
    (score -> combine/normalise -> top-k -> shrink -> stretch); nothing is skipped or reused.
    """

    def __init__(self, n_frames: int, height: int, width: int, block_size: int = 16, shrink_amount: float = 0.5,
                 alpha: float = 0.5, beta: float = 0.5, device="cuda", depth: int = 2, score_fn=None,
                 move_ctas_per_sm: int = 3, comm_fn=None, split_stretch: bool = False, stretch_ctas_per_sm: int | None = None,
                 prepared: bool = True):
        """prepared: replay pre-built C-ABI calls for (clip, slot) pairs seen before instead of going through the
        ops wrappers every time (same kernels, same arguments, a fraction of the host time); only without score_fn."""
        self.pipe = ElvisV1(block_size, shrink_amount, alpha, beta)
        self.prepared = prepared
        self._programs: dict = {}
        d = torch.device(device)
        self._dev_index = d.index if d.index is not None else torch.cuda.current_device()
        self.dev = torch.device(device)
        # footprint of the shrink/stretch kernels per SM while they share it with the scoring
        # kernel: 3 x 256 threads keeps ~85 % of their stand-alone bandwidth and leaves the two
        # resident scoring CTAs their registers and issue slots (measured optimum on B200)
        self.move_ctas = move_ctas_per_sm
        self.T, self.H, self.W = n_frames, height, width
        by, bx = height // block_size, width // block_size
        self.k = blocks_to_remove(shrink_amount, bx)
        sw = (bx - self.k) * block_size
        # the move stage gets the higher stream priority: its short, register-light CTAs are placed
        # into the resources the long-running scoring CTAs leave free instead of queueing behind them
        prio = [int(x) for x in os.environ.get("ELVIS_PIPE_PRIO", "0,-1,-1").split(",")]
        self.s_score = torch.cuda.Stream(self.dev, priority=prio[0])
        self.s_move = torch.cuda.Stream(self.dev, priority=prio[1])
        # optional third stage: stretch on its own stream, so that the stretch of clip i overlaps
        # the shrink of clip i+1 as well (neither saturates DRAM alone)
        self.s_stretch = torch.cuda.Stream(self.dev, priority=prio[2]) if split_stretch else None
        self.stretch_ctas = move_ctas_per_sm if stretch_ctas_per_sm is None else stretch_ctas_per_sm
        self.score_fn = score_fn      # optional: clip, slot -> scores (e.g. the sharded scorer)
        # optional: clip -> None, communication the scorer depends on (the halo exchange of the
        # sharded path).  It runs on a third stream as soon as the clip is submitted, i.e. while
        # the previous clip is still being scored, and the score stage waits for it.
        self.comm_fn = comm_fn
        self.s_comm = torch.cuda.Stream(self.dev, priority=-1) if comm_fn is not None else None
        self.slots = []
        for _ in range(depth):
            self.slots.append({
                "sc": torch.empty((n_frames, by, bx), dtype=torch.float32, device=self.dev),
                "tc": torch.empty((n_frames, by, bx), dtype=torch.float32, device=self.dev),
                "norm": torch.empty(4, dtype=torch.float32, device=self.dev),
                "scores": torch.empty((n_frames, by, bx), dtype=torch.float64, device=self.dev),
                "smm": torch.empty(2, dtype=torch.float64, device=self.dev),
                "mask": torch.empty((n_frames, by, bx), dtype=torch.uint8, device=self.dev),
                "shrunk": Yuv420.empty(n_frames, height, sw, self.dev),
                "full": Yuv420.empty(n_frames, height, width, self.dev),
                "scored": torch.cuda.Event(), "shrunk_ev": torch.cuda.Event(), "done": torch.cuda.Event(), "ready": torch.cuda.Event(),
                "used": False,
            })
        self._next = 0

    # ------------------------------------------------------------------ prepared calls
    def _program(self, clip: Yuv420, index: int):
        """The five C-ABI calls of one clip through slot `index`, arguments converted once (None when the clip's
        geometry needs the generic per-plane path)."""
        key = (index,) + tuple((p.data_ptr(), p.stride(0), p.stride(1)) for p in clip.planes) + tuple(clip.y.shape)
        prog = self._programs.get(key)
        if prog is not None or key in self._programs:
            return prog
        slot, p = self.slots[index], self.pipe
        T, H, W = clip.y.shape
        by, bx = H // p.bs, W // p.bs
        small_bx = bx - self.k
        ok = ((T, H, W) == (self.T, self.H, self.W) and small_bx > 0 and _fused_move_ok(clip, p.bs)
              and _fused_move_ok(slot["shrunk"], p.bs) and _fused_move_ok(slot["full"], p.bs) and clip.y.is_cuda)
        if not ok:
            if len(self._programs) > 256:
                self._programs.clear()
            self._programs[key] = None
            return None
        vp, i32 = C.c_void_p, C.c_int32
        st = {"score": vp(self.s_score.cuda_stream), "move": vp(self.s_move.cuda_stream),
              "stretch": vp((self.s_stretch or self.s_move).cuda_stream)}
        y_plane = ops.plane_of(clip.y, "y")
        src = (_lib.Plane * 3)(*[ops.plane_of(t, "clip") for t in clip.planes])
        shr = (_lib.Plane * 3)(*[ops.plane_of(t, "shrunk") for t in slot["shrunk"].planes])
        ful = (_lib.Plane * 3)(*[ops.plane_of(t, "full") for t in slot["full"].planes])
        ptr = lambda t: vp(t.data_ptr())       # noqa: E731
        smooth = int(p.beta < 1 and T >= 2)
        prog = (
            _Prepared("elvis_score_sc_tc", (C.byref(y_plane), i32(T), vp(0), i32(p.bs), i32(p.dct_size), ptr(slot["sc"]), ptr(slot["tc"]),
                                            ptr(slot["norm"]), i32(0), i32(T), st["score"]), (y_plane, clip.y)),
            _Prepared("elvis_combine_removability", (ptr(slot["sc"]), ptr(slot["tc"]), i32(_lib.F32), ptr(slot["norm"]), i32(T), i32(by), i32(bx),
                                                     i32(0), i32(T), i32(1), i32(1), vp(0), C.c_double(p.alpha), C.c_double(p.beta), i32(smooth),
                                                     ptr(slot["scores"]), ptr(slot["smm"]), st["score"])),
            _Prepared("elvis_normalize_select_rows", (ptr(slot["scores"]), ptr(slot["smm"]), i32(T), i32(by), i32(bx), vp(0), i32(self.k),
                                                      i32(ops.REMOVE_HIGH), ptr(slot["mask"]), st["move"])),
            _Prepared("elvis_shrink_yuv420", (src, shr, i32(T), i32(p.bs), i32(by), i32(bx), i32(small_bx), ptr(slot["mask"]),
                                              i32(self.move_ctas), st["move"]), (src, shr, clip)),
            _Prepared("elvis_stretch_yuv420", (shr, ful, i32(T), i32(p.bs), i32(by), i32(bx), i32(small_bx), ptr(slot["mask"]),
                                               i32(self.stretch_ctas if self.s_stretch is not None else self.move_ctas), st["stretch"]), (shr, ful)),
        )
        if len(self._programs) > 256:
            self._programs.clear()
        self._programs[key] = prog
        return prog

    def _submit_prepared(self, slot: dict, prog) -> dict:
        score, combine, select, shrink, stretch = prog
        ready = slot["ready"]
        ready.record()                                        # inputs produced on the caller's stream
        self.s_score.wait_event(ready)
        if slot["used"]:
            self.s_score.wait_event(slot["done"])             # the slot's previous clip has left the move stage
        score()
        combine()
        slot["scored"].record(self.s_score)
        self.s_move.wait_event(slot["scored"])
        select()
        shrink()
        if self.s_stretch is None:
            stretch()
            slot["done"].record(self.s_move)
        else:
            slot["shrunk_ev"].record(self.s_move)
            self.s_stretch.wait_event(slot["shrunk_ev"])
            stretch()
            slot["done"].record(self.s_stretch)
        slot["used"] = True
        return slot

    def submit(self, clip: Yuv420) -> dict:
        """Enqueue one clip; returns its slot (dict with `scores`, `mask`, `shrunk`, `full`,
        and the `done` event).  The clip must stay unmodified until `done` fires."""
        index = self._next
        slot = self.slots[index]
        self._next = (self._next + 1) % len(self.slots)
        if self.prepared and self.score_fn is None and self.comm_fn is None and torch.cuda.current_device() == self._dev_index:
            prog = self._program(clip, index)
            if prog is not None:
                return self._submit_prepared(slot, prog)
        ready = torch.cuda.Event()
        ready.record()                                   # inputs produced on the caller's stream
        comm_done = None
        if self.comm_fn is not None:
            with torch.cuda.stream(self.s_comm):
                self.s_comm.wait_event(ready)
                self.comm_fn(clip)
                comm_done = torch.cuda.Event()
                comm_done.record()
        with torch.cuda.stream(self.s_score):
            self.s_score.wait_event(ready)
            if comm_done is not None:
                self.s_score.wait_event(comm_done)
            if slot["used"]:
                self.s_score.wait_event(slot["done"])    # the slot's previous clip has left the move stage
            if self.score_fn is not None:
                slot["scores"] = self.score_fn(clip, slot)
            else:
                p = self.pipe
                ops.score_sc_tc(clip.y, p.bs, out=(slot["sc"], slot["tc"], slot["norm"]))
                ops.combine_removability(slot["sc"], slot["tc"], slot["norm"], p.alpha, p.beta,
                                         out=(slot["scores"], slot["smm"]))
            slot["scored"].record()
        with torch.cuda.stream(self.s_move):
            self.s_move.wait_event(slot["scored"])
            # the final normalisation (elvis.py:1218) rides on the ranking pass unless the scorer already applied it
            ops.select_rows(slot["scores"], self.k, ops.REMOVE_HIGH, out=slot["mask"],
                            normalize_with=None if self.score_fn is not None else slot["smm"])
            sh, fu, bs = slot["shrunk"], slot["full"], self.pipe.bs
            move_planes(clip, sh, slot["mask"], bs, sh.y.shape[2] // bs, False, self.move_ctas)
            if self.s_stretch is None:
                move_planes(sh, fu, slot["mask"], bs, sh.y.shape[2] // bs, True, self.move_ctas)
                slot["done"].record()
            else:
                slot["shrunk_ev"].record()
        if self.s_stretch is not None:
            with torch.cuda.stream(self.s_stretch):
                self.s_stretch.wait_event(slot["shrunk_ev"])
                move_planes(sh, fu, slot["mask"], bs, sh.y.shape[2] // bs, True, self.stretch_ctas)
                slot["done"].record()
        slot["used"] = True
        return slot

    def join(self) -> None:
        """Make the caller's current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream()
        for slot in self.slots:
            if slot["used"]:
                cur.wait_event(slot["done"])


class PresleyV2:
    """v2 per-block degradations on a planar clip (luma block bs, chroma block bs/2)."""

    def __init__(self, block_size: int = 16):
        self.bs = block_size

    def _each(self, clip: Yuv420, fn, out: Optional[Yuv420]):
        if out is None:
            out = Yuv420.empty(clip.y.shape[0], clip.y.shape[1], clip.y.shape[2], clip.y.device)
        fn(clip.y, self.bs, out.y)
        fn(clip.u, self.bs // 2, out.u)
        fn(clip.v, self.bs // 2, out.v)
        return out

    def blur(self, clip: Yuv420, rounds: torch.Tensor, out: Optional[Yuv420] = None) -> Yuv420:
        return self._each(clip, lambda p, pb, o: ops.degrade_blur(p, rounds, pb, out=o), out)

    def downsample_pow2(self, clip: Yuv420, levels: torch.Tensor, max_level: int, out: Optional[Yuv420] = None) -> Yuv420:
        """level l reduces a block by 2**l per axis (elvis.py:2147-2163); chroma blocks use
        the same factor on their half-size block."""
        if out is None:
            out = Yuv420.empty(clip.y.shape[0], clip.y.shape[1], clip.y.shape[2], clip.y.device)
        if levels.shape[0] == clip.y.shape[0] and ops.degrade_yuv420("downsample_pow2", clip.planes, out.planes, levels, self.bs, max_level):
            return out            # Y, U and V in one launch

        def fn(p, pb, o):
            smalls = [max(1, pb >> l) for l in range(max_level + 1)]
            ops.degrade_downsample(p, levels, pb, smalls, out=o)
        return self._each(clip, fn, out)

    def dampen(self, clip: Yuv420, strength: torch.Tensor, out: Optional[Yuv420] = None) -> Yuv420:
        return self._each(clip, lambda p, pb, o: ops.dct_dampen(p, strength, pb, out=o), out)


class HostElvisV1:
    """The headline path for callers whose frames live in HOST memory (the end-to-end API): an
    I420 clip in; removal masks, shrunk I420 clip and stretched I420 clip out, all host tensors.
    Work is enqueued on one of `depth` private streams with its own device buffers, so consecutive
    calls overlap one clip's device->host copies with the next clip's host->device copy
    (full-duplex PCIe).  Pass pinned tensors to get async copies.

    outputs selects what is produced and copied back: the reference's SERVER keeps
    ("mask", "shrunk") -- it encodes the shrunk frames and ships the masks as a side channel
    (elvis.py:4389-4418) -- while the stretched clip is the CLIENT's product (HostElvisClient;
    elvis.py:4537-4557).  The default returns all three (the composite the benchmark's `e2e` times).
    pack_masks=True returns np.packbits-compatible bit-packed masks (the side-channel format,
    elvis.py:4412-4418), packed on the device.  score_fn(luma_with_halo_slots, slot_index) -> scores
    replaces the local scorer (the frame-sharded scorer of elvis_b200.sharding); the clip's luma then
    sits in slot["halo"], a HaloClip over the input buffer.
    """

    def __init__(self, n_frames: int, height: int, width: int, block_size: int = 16, shrink_amount: float = 0.5,
                 alpha: float = 0.5, beta: float = 0.5, device="cuda", depth: int = 2,
                 outputs: Tuple[str, ...] = ("mask", "shrunk", "stretched"), pack_masks: bool = False, score_fn=None):
        bad = set(outputs) - {"mask", "shrunk", "stretched"}
        if bad or not outputs:
            raise ValueError(f"outputs must be a non-empty subset of mask / shrunk / stretched, got {outputs}")
        self.pipe = ElvisV1(block_size, shrink_amount, alpha, beta)
        self.T, self.H, self.W = n_frames, height, width
        self.dev = torch.device(device)
        self.outputs, self.pack_masks, self.score_fn = tuple(outputs), pack_masks, score_fn
        by, bx = height // block_size, width // block_size
        self.by, self.bx = by, bx
        self.k = blocks_to_remove(shrink_amount, bx)
        self.sw = (bx - self.k) * block_size
        self.frame_bytes = height * width * 3 // 2
        self.slots = []
        from .sharding import HaloClip
        for _ in range(depth):
            # two spare frames around the clip: halo slots of the frame-sharded scorer (unused otherwise)
            buf = torch.empty((n_frames + 2, self.frame_bytes), dtype=torch.uint8, device=self.dev)
            slot = {"stream": torch.cuda.Stream(self.dev), "buf": buf, "in": buf[1:n_frames + 1],
                    "halo": HaloClip(n_frames, height, width, self.dev, buf=buf[:, :height * width].unflatten(1, (height, width))),
                    "shrunk": torch.empty((n_frames, height * self.sw * 3 // 2), dtype=torch.uint8, device=self.dev),
                    "mask": torch.empty((n_frames, by, bx), dtype=torch.uint8, device=self.dev)}
            if "stretched" in outputs:
                slot["full"] = torch.empty((n_frames, self.frame_bytes), dtype=torch.uint8, device=self.dev)
            if pack_masks:
                slot["bits"] = torch.empty(((n_frames * by * bx + 7) // 8,), dtype=torch.uint8, device=self.dev)
            self.slots.append(slot)
        self._next = 0

    @property
    def mask_bytes(self) -> int:
        n = self.T * self.by * self.bx
        return (n + 7) // 8 if self.pack_masks else n

    def host_buffers(self, pinned: bool = True):
        """Convenience: (shrunk_i420, stretched_i420, masks) host tensors of the right shapes (None for
        outputs that were not requested)."""
        mk = lambda *s: torch.empty(s, dtype=torch.uint8, pin_memory=pinned)   # noqa: E731
        return (mk(self.T, self.H * self.sw * 3 // 2) if "shrunk" in self.outputs else None,
                mk(self.T, self.frame_bytes) if "stretched" in self.outputs else None,
                (mk(self.mask_bytes) if self.pack_masks else mk(self.T, self.by, self.bx)) if "mask" in self.outputs else None)

    @property
    def h2d_bytes(self) -> int:
        return self.T * self.frame_bytes

    @property
    def d2h_bytes(self) -> int:
        n = 0
        if "shrunk" in self.outputs:
            n += self.T * self.H * self.sw * 3 // 2
        if "stretched" in self.outputs:
            n += self.T * self.frame_bytes
        if "mask" in self.outputs:
            n += self.mask_bytes
        return n

    def process(self, i420_host: torch.Tensor, shrunk_host: Optional[torch.Tensor] = None,
                stretched_host: Optional[torch.Tensor] = None, mask_host: Optional[torch.Tensor] = None) -> torch.cuda.Event:
        """Enqueue one clip; returns an event that fires when the requested outputs are in host
        memory.  Call .synchronize() on it (or on the device) before reading them."""
        slot = self.slots[self._next]
        index = self._next
        self._next = (self._next + 1) % len(self.slots)
        with torch.cuda.stream(slot["stream"]):
            slot["in"].copy_(i420_host, non_blocking=True)
            clip = Yuv420.from_i420(slot["in"], self.H, self.W)
            shrunk = Yuv420.from_i420(slot["shrunk"], self.H, self.sw)
            scores = self.score_fn(slot["halo"], index) if self.score_fn is not None else self.pipe.score(clip)
            mask = ops.select_rows(scores, self.k, ops.REMOVE_HIGH, out=slot["mask"])
            move_planes(clip, shrunk, mask, self.pipe.bs, self.bx - self.k, stretch_=False)
            if "stretched" in self.outputs:
                self.pipe.stretch(shrunk, mask, Yuv420.from_i420(slot["full"], self.H, self.W))
                stretched_host.copy_(slot["full"], non_blocking=True)
            if "mask" in self.outputs:
                mask_host.copy_(ops.pack_mask_bits(mask, out=slot["bits"]) if self.pack_masks else mask, non_blocking=True)
            if "shrunk" in self.outputs:
                shrunk_host.copy_(slot["shrunk"], non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        return done


class HostElvisClient:
    """The client side of the same flow for host buffers (elvis.py:4537-4557): shrunk I420 clip and the
    bit-packed mask side channel in, stretched I420 clip (black placeholders for the inpainter) out."""

    def __init__(self, n_frames: int, height: int, width: int, block_size: int = 16, shrink_amount: float = 0.5,
                 device="cuda", depth: int = 2, packed_masks: bool = True):
        self.bs, self.T, self.H, self.W = block_size, n_frames, height, width
        self.dev = torch.device(device)
        self.by, self.bx = height // block_size, width // block_size
        self.k = blocks_to_remove(shrink_amount, self.bx)
        self.sw = (self.bx - self.k) * block_size
        self.packed_masks = packed_masks
        n = n_frames * self.by * self.bx
        self.mask_bytes = (n + 7) // 8 if packed_masks else n
        self.slots = [{"stream": torch.cuda.Stream(self.dev),
                       "shrunk": torch.empty((n_frames, height * self.sw * 3 // 2), dtype=torch.uint8, device=self.dev),
                       "mask_in": torch.empty((self.mask_bytes,), dtype=torch.uint8, device=self.dev),
                       "full": torch.empty((n_frames, height * width * 3 // 2), dtype=torch.uint8, device=self.dev)}
                      for _ in range(depth)]
        self._next = 0

    @property
    def h2d_bytes(self) -> int:
        return self.T * self.H * self.sw * 3 // 2 + self.mask_bytes

    @property
    def d2h_bytes(self) -> int:
        return self.T * self.H * self.W * 3 // 2

    def process(self, shrunk_host: torch.Tensor, mask_host: torch.Tensor, stretched_host: torch.Tensor) -> torch.cuda.Event:
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        with torch.cuda.stream(slot["stream"]):
            slot["shrunk"].copy_(shrunk_host, non_blocking=True)
            slot["mask_in"].copy_(mask_host.reshape(-1), non_blocking=True)
            shape = (self.T, self.by, self.bx)
            mask = ops.unpack_mask_bits(slot["mask_in"], shape) if self.packed_masks else slot["mask_in"].view(shape)
            move_planes(Yuv420.from_i420(slot["shrunk"], self.H, self.sw), Yuv420.from_i420(slot["full"], self.H, self.W), mask,
                        self.bs, self.bx - self.k, stretch_=True)
            stretched_host.copy_(slot["full"], non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        return done
