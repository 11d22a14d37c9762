"""Seeded synthetic YUV 4:2:0 clips generated directly in device memory:
luma = smooth low-frequency field + band-limited texture whose amplitude varies per 64x64
region (so SC has spread and removal decisions are spatially coherent, as in real video) +
a global 2 px/frame pan and three moving rectangles (so TC != 0), clipped to [16, 235];
chroma = smooth fields with the same pan.  Plumbing only -- plain torch ops."""
from __future__ import annotations

import math

import torch

from .yuv import Yuv420


def synth_yuv420(n_frames: int, height: int, width: int, seed: int = 1234, device="cuda",
                 out: Yuv420 | None = None, frame_offset: int = 0) -> Yuv420:
    """frame_offset lets several ranks generate disjoint ranges of ONE global clip."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    total_pan = 2 * (frame_offset + n_frames) + 16
    wide = width + total_pan
    # fixed texture field and per-region amplitude, shared by all frames (the pan moves over it)
    tex = torch.randn((height, wide), generator=g, device=dev)
    tex = (tex + torch.roll(tex, 1, 0) + torch.roll(tex, 1, 1) + torch.roll(tex, (1, 1), (0, 1))) * 0.5
    amp = torch.rand(((height + 63) // 64, (wide + 63) // 64), generator=g, device=dev) ** 2 * 45.0
    amp = amp.repeat_interleave(64, 0).repeat_interleave(64, 1)[:height, :wide]
    yy = torch.arange(height, device=dev, dtype=torch.float32)[:, None]
    xx = torch.arange(wide, device=dev, dtype=torch.float32)[None, :]
    field = 118.0 + 55.0 * torch.sin(xx / 211.0) * torch.cos(yy / 157.0) + 18.0 * torch.sin((xx + yy) / 53.0) + tex * amp
    cu = 128.0 + 60.0 * torch.sin(xx[:, ::2] / 301.0) * torch.cos(yy[::2] / 173.0)
    cv = 128.0 + 60.0 * torch.cos(xx[:, ::2] / 257.0) * torch.sin(yy[::2] / 199.0)
    if out is None:
        out = Yuv420.empty(n_frames, height, width, dev)
    rect = max(16, min(height, width) // 8)
    for i in range(n_frames):
        t = frame_offset + i
        f = field[:, 2 * t:2 * t + width].clone()
        for r, (vx, vy, lum) in enumerate(((7, 3, 70.0), (-5, 4, -60.0), (3, -6, 45.0))):
            x0 = int((width * (r + 1) / 4 + vx * t) % max(1, width - rect))
            y0 = int((height * (r + 1) / 4 + vy * t) % max(1, height - rect))
            f[y0:y0 + rect, x0:x0 + rect] += lum
        out.y[i] = f.round().clamp_(16, 235).to(torch.uint8)
        out.u[i] = cu[:, t:t + width // 2].round().clamp_(16, 240).to(torch.uint8)
        out.v[i] = cv[:, t:t + width // 2].round().clamp_(16, 240).to(torch.uint8)
    return out
