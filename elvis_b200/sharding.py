"""Frame sharding across the GPUs of one box.

Frames are independent for shrink / stretch / degrade; scoring couples frame t with t-1
(TC) and t+1 (the blend uses TC[t+1], elvis.py:1180) and smoothing couples t with t-1
(elvis.py:1210-1213).  Each rank therefore owns a contiguous frame range (the split of the
reference's chunk_for_devices, elvis.py:264-278) and needs ONE luma frame of halo on each
side, exchanged with torch.distributed send/recv (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The elvis-mode global normalisations need two tiny all-reduces
({sc,tc} min/max and the final min/max); the utils per-frame mode needs none.

Everything here is device-agnostic plumbing; the arithmetic is injected (`kernels`), which is
elvis_b200.ops in production and an oracle-backed stand-in in tests/test_sharding_gloo.py.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: base = T // G frames each, the first T % G ranks get one extra."""
    base, extra = divmod(n_frames, world)
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def check_shardable(n_frames: int, world: int) -> None:
    """Every rank must own at least one frame (the same verdict on every rank, so nobody hangs)."""
    if n_frames < world:
        raise ValueError(f"{n_frames} frames cannot be sharded over {world} ranks: every rank needs at least one frame")


class HaloClip:
    """A rank's luma frames with one halo slot on each side, so that received halos land in
    place and the scoring kernel sees one contiguous extended clip (no concatenation copy).

        buf[0]            Y[a-1]   (valid iff rank > 0)
        buf[1 : n+1]      Y[a:b]   (owned)
        buf[n+1]          Y[b]     (valid iff rank < world-1)
    """

    def __init__(self, n_local: int, height: int, width: int, device, pin: bool = False, buf: Optional[torch.Tensor] = None):
        """buf: optional caller-owned (n_local + 2, H, W) uint8 view (frames may be strided, e.g. the luma
        of an I420 buffer; each frame must be contiguous so that it can be sent as one message)."""
        self.n = n_local
        if buf is not None and (tuple(buf.shape) != (n_local + 2, height, width) or not buf[0].is_contiguous()):
            raise ValueError("buf must be (n_local + 2, H, W) with contiguous frames")
        self.buf = buf if buf is not None else torch.empty((n_local + 2, height, width), dtype=torch.uint8, device=device)

    @property
    def owned(self) -> torch.Tensor:
        return self.buf[1:self.n + 1]

    def extended(self, rank: int, world: int) -> Tuple[torch.Tensor, int]:
        """(frames the kernels should see, index of the first owned frame inside them)."""
        lo = 0 if rank > 0 else 1
        hi = self.n + 2 if rank < world - 1 else self.n + 1
        return self.buf[lo:hi], 1 - lo


def exchange_halo(clip: HaloClip, rank: int, world: int, group=None) -> None:
    """Send my first/last owned frame to the previous/next rank and receive theirs into the
    halo slots.  One batched isend/irecv group => a single NCCL group call."""
    if world == 1:
        return
    if clip.n == 0:
        # frame_range gives a rank no frames when n_frames < world; its neighbours would still post
        # sends / receives to it and wait for ever
        raise ValueError("a rank without frames cannot take part in the halo exchange: shard over at most n_frames ranks")
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, clip.buf[1], rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, clip.buf[0], rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, clip.buf[clip.n], rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, clip.buf[clip.n + 1], rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


_signs: dict = {}


def allreduce_minmax(mm: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """In-place exact all-reduce of interleaved {min, max, min, max, ...}: one MAX reduction
    of {-min, max, ...}."""
    if world == 1:
        return mm
    key = (mm.device, mm.dtype, mm.numel())
    sign = _signs.get(key)
    if sign is None:
        sign = torch.ones_like(mm)
        sign[0::2] = -1
        _signs[key] = sign
    mm.mul_(sign)
    dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
    mm.mul_(sign)
    return mm


def sharded_removability(clip: HaloClip, n_frames_total: int, block_size: int, alpha: float, beta: float,
                         rank: int, world: int, background: Optional[torch.Tensor] = None, kernels=None,
                         group=None, exchange: bool = True, transport=None, dct_size: int = 8) -> torch.Tensor:
    """elvis-mode removability (elvis.py:1160-1220) of the owned frames -> (n, By, Bx) float64.
    dct_size: 8 (8 x 8 tiles) or block_size (one transform per block, the reference's EVCA call).
    background: optional uint8 (n+2, By, Bx) laid out like clip.buf (halo slots filled by the
    caller when it has the neighbours' masks; only slot 0 is ever read).  exchange=False: the
    caller has already run exchange_halo (e.g. ahead of time on a communication stream).
    transport: None = NCCL send/recv + all-reduce (torch.distributed); an elvis_b200.peer.PeerGroup =
    copy-engine peer copies and mailbox all-reduces over NVLink (clip must come from its halo_clip())."""
    if kernels is None:
        from . import ops as kernels
    check_shardable(n_frames_total, world)
    peer = transport if (transport is not None and world > 1) else None
    if peer is not None:
        if exchange:
            peer.exchange_halo(clip)
        peer.wait_halo(clip)
    elif exchange:
        exchange_halo(clip, rank, world, group)
    ext, first = clip.extended(rank, world)
    size_kw = {} if dct_size == 8 else {"dct_size": dct_size}
    sc, tc, norm = kernels.score_sc_tc(ext, block_size, minmax_range=(first, first + clip.n), **size_kw)
    if peer is not None:
        peer.release_halo(clip)          # stream ordered: the scoring kernel has read the halo slots
        reduce_ = peer.allreduce_minmax_
    else:
        reduce_ = lambda mm: allreduce_minmax(mm, world, group)      # noqa: E731
    reduce_(norm)
    bg = None
    if background is not None:
        lo = 1 - first
        bg = background[lo:lo + ext.shape[0]].contiguous()
    r, mm = kernels.combine_removability(sc, tc, norm, alpha, beta, bg, t_begin=first, t_count=clip.n,
                                         is_first=(rank == 0), is_last=(rank == world - 1),
                                         clip_frames=n_frames_total)
    reduce_(mm)
    return kernels.normalize_(r, mm)


def sharded_importance(clip: HaloClip, block_size: int, alpha: float, beta: float, rank: int, world: int,
                       foreground: Optional[torch.Tensor] = None, kernels=None, group=None, transport=None,
                       dct_size: int = 8) -> torch.Tensor:
    """utils-mode importance (utils.py:665-688) of the owned frames: per-frame normalisation,
    so the halo exchange is the only communication (transport: see sharded_removability)."""
    if kernels is None:
        from . import ops as kernels
    peer = transport if (transport is not None and world > 1) else None
    if peer is not None:
        peer.exchange_halo(clip)
        peer.wait_halo(clip)
    else:
        exchange_halo(clip, rank, world, group)
    ext, first = clip.extended(rank, world)
    sc, tc, _ = kernels.score_sc_tc(ext, block_size, **({} if dct_size == 8 else {"dct_size": dct_size}))
    if peer is not None:
        peer.release_halo(clip)
    fg = None
    if foreground is not None:
        lo = 1 - first
        fg = foreground[lo:lo + ext.shape[0]].to(sc.dtype).contiguous()
    return kernels.importance_scores(sc, tc, fg, alpha, beta, t_begin=first, t_count=clip.n,
                                     is_first=(rank == 0), is_last=(rank == world - 1))
