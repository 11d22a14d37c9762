"""Batched planar-YUV 4:2:0 pipelines: the B200-native shape of the hot path.

The reference processes one packed BGR frame per Python call; here a whole clip stays in HBM
as three planes (the layout of the raw yuv420p file the reference feeds to EVCA,
elvis.py:4324-4330) and every stage is one kernel launch per plane over all frames.  The
removal mask / level map computed from luma is applied to the co-located chroma blocks at
half the block size (SURVEY.md appendix "Layout")."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import ops
from .elvis import blocks_to_remove


@dataclass
class Yuv420:
    """Three CUDA uint8 planes: y (T, H, W), u and v (T, H/2, W/2)."""
    y: torch.Tensor
    u: torch.Tensor
    v: torch.Tensor

    @staticmethod
    def from_i420(buf: torch.Tensor, height: int, width: int) -> "Yuv420":
        """Views into a (T, H*W*3/2) I420 buffer (no copy)."""
        n, cw, ch = height * width, width // 2, height // 2
        return Yuv420(buf[:, :n].unflatten(1, (height, width)),
                      buf[:, n:n + cw * ch].unflatten(1, (ch, cw)),
                      buf[:, n + cw * ch:n + 2 * cw * ch].unflatten(1, (ch, cw)))

    @staticmethod
    def empty(n_frames: int, height: int, width: int, device="cuda") -> "Yuv420":
        buf = torch.empty((n_frames, height * width * 3 // 2), dtype=torch.uint8, device=device)
        return Yuv420.from_i420(buf, height, width)

    @property
    def planes(self):
        return (self.y, self.u, self.v)

    @property
    def nbytes(self) -> int:
        return sum(p.numel() for p in self.planes)


class ElvisV1:
    """score -> per-row top-k mask -> shrink, and stretch (elvis.py:4350-4394, 4550-4557),
    for a planar clip resident on the GPU."""

    def __init__(self, block_size: int = 16, shrink_amount: float = 0.5, alpha: float = 0.5, beta: float = 0.5):
        if block_size % 2:
            raise ValueError("4:2:0 chroma needs an even block size")
        self.bs, self.shrink_amount, self.alpha, self.beta = block_size, shrink_amount, alpha, beta

    def grid(self, clip: Yuv420) -> Tuple[int, int, int]:
        h, w = clip.y.shape[1:]
        if h % self.bs or w % self.bs:
            raise ValueError("Image dimensions must be divisible by block_size.")
        bx = w // self.bs
        return h // self.bs, bx, blocks_to_remove(self.shrink_amount, bx)

    def score(self, clip: Yuv420, background: Optional[torch.Tensor] = None) -> torch.Tensor:
        sc, tc, norm = ops.score_sc_tc(clip.y, self.bs)
        r, mm = ops.combine_removability(sc, tc, norm, self.alpha, self.beta, background)
        return ops.normalize_(r, mm)

    def shrink(self, clip: Yuv420, scores: torch.Tensor, out: Optional[Yuv420] = None):
        by, bx, k = self.grid(clip)
        mask = ops.select_rows(scores, k, ops.REMOVE_HIGH)
        if out is None:
            out = Yuv420.empty(clip.y.shape[0], by * self.bs, (bx - k) * self.bs, clip.y.device)
        ops.shrink(clip.y, mask, self.bs, bx - k, out=out.y)
        ops.shrink(clip.u, mask, self.bs // 2, bx - k, out=out.u)
        ops.shrink(clip.v, mask, self.bs // 2, bx - k, out=out.v)
        return out, mask

    def stretch(self, shrunk: Yuv420, mask: torch.Tensor, out: Optional[Yuv420] = None) -> Yuv420:
        t, by, bx = mask.shape
        if out is None:
            out = Yuv420.empty(t, by * self.bs, bx * self.bs, shrunk.y.device)
        ops.stretch(shrunk.y, mask, self.bs, out=out.y)
        ops.stretch(shrunk.u, mask, self.bs // 2, out=out.u)
        ops.stretch(shrunk.v, mask, self.bs // 2, out=out.v)
        return out

    def run(self, clip: Yuv420, background: Optional[torch.Tensor] = None,
            shrunk_out: Optional[Yuv420] = None, stretched_out: Optional[Yuv420] = None):
        """One pass of the headline path: returns (scores, mask, shrunk, stretched)."""
        scores = self.score(clip, background)
        shrunk, mask = self.shrink(clip, scores, shrunk_out)
        stretched = self.stretch(shrunk, mask, stretched_out)
        return scores, mask, shrunk, stretched


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
        def fn(p, pb, o):
            smalls = [max(1, pb >> l) for l in range(max_level + 1)]
            ops.degrade_downsample(p, levels, pb, smalls, out=o)
        return self._each(clip, fn, out)

    def dampen(self, clip: Yuv420, strength: torch.Tensor, out: Optional[Yuv420] = None) -> Yuv420:
        return self._each(clip, lambda p, pb, o: ops.dct_dampen(p, strength, pb, out=o), out)
