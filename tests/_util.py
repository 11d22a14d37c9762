"""Seeded synthetic inputs shared by the parity tests."""
from __future__ import annotations

import numpy as np


def synth_luma(T: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    """Structured clip: smooth gradient + region-dependent texture, global 2 px/frame pan,
    a moving bright square; clipped to [16, 235].  (T, H, W) uint8."""
    rng = np.random.default_rng(seed)
    pad = 2 * T + 8
    tex = rng.normal(0, 1, (H, W + pad))
    amp = np.kron(rng.uniform(0, 40, ((H + 31) // 32, (W + pad + 31) // 32)), np.ones((32, 32)))[:H, :W + pad]
    yy, xx = np.mgrid[0:H, 0:W + pad]
    field = 110 + 50 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + tex * amp
    out = np.empty((T, H, W), np.uint8)
    for t in range(T):
        f = field[:, 2 * t:2 * t + W].copy()
        s = max(8, H // 4)
        x0, y0 = (5 * t) % max(1, W - s), (3 * t) % max(1, H - s)
        f[y0:y0 + s, x0:x0 + s] += 60
        out[t] = np.clip(np.rint(f), 16, 235).astype(np.uint8)
    return out


def synth_yuv420(T: int, H: int, W: int, seed: int = 0):
    rng = np.random.default_rng(seed + 1000)
    y = synth_luma(T, H, W, seed)
    u = rng.integers(0, 256, (T, H // 2, W // 2), dtype=np.uint8)
    v = rng.integers(0, 256, (T, H // 2, W // 2), dtype=np.uint8)
    return y, u, v


def random_scores(rng, shape, ties: str = "none") -> np.ndarray:
    s = rng.random(shape)
    if ties == "quantised":
        s = np.round(s * 6) / 6
    elif ties == "all":
        s = np.full(shape, 0.5)
    return s
