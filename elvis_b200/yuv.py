"""Planar YUV 4:2:0 clip container: three uint8 planes over torch storage.  Pure plumbing -- this
module does not load the CUDA library, so host-only tools (the synthetic clip generator, the CPU arm
of bench.py) can use it without mapping libelvis_b200.so."""
from __future__ import annotations

from dataclasses import dataclass

import torch


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
