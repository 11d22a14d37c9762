"""TEST INFRASTRUCTURE ONLY -- normative spec of the v2 "DCT-coefficient dampening".

PARITY UNPINNED: the reference contains no implementation of this degradation; it is
named only in README.md:11 and README.md:44.  This file defines it and the CUDA kernel
is checked against it.

Definition (one 2-D uint8 plane, block size pb = bs for luma/packed channels, bs // 2 for
4:2:0 chroma; pb must be a multiple of 8):
  * every pb x pb block is tiled into 8x8 transform tiles; C = orthonormal DCT-II of the
    tile (float64 here, float32 on the device);
  * the block's strength s in [0, 1] (clamped) attenuates coefficient (u, v) by
        g(s, u, v) = 2 ** (-DAMPEN_OCTAVES * s * (u + v) / 14)
    (DC untouched; the highest frequency loses DAMPEN_OCTAVES = 4 octaves at s = 1);
  * out = clip(rint(IDCT(g * C)), 0, 255)  (round half to even).
s = 0 is the identity.  Rows/columns outside whole blocks are copied through.
"""
from __future__ import annotations

import numpy as np
from scipy.fft import dctn, idctn

DAMPEN_OCTAVES = 4.0


def gains(strength: np.ndarray) -> np.ndarray:
    """(...,) strengths -> (..., 8, 8) gains."""
    s = np.clip(np.asarray(strength, np.float64), 0.0, 1.0)
    u = np.arange(8, dtype=np.float64)
    f = (u[:, None] + u[None, :]) / 14.0
    return np.exp2(-DAMPEN_OCTAVES * s[..., None, None] * f)


def dampen_plane(plane: np.ndarray, strength: np.ndarray, pb: int, return_float: bool = False):
    """plane (H, W) uint8, strength (By, Bx) -> dampened plane (uint8, or the un-rounded
    float64 reconstruction when return_float)."""
    if pb % 8:
        raise ValueError("plane block size must be a multiple of 8")
    h, w = plane.shape
    by, bx = h // pb, w // pb
    r = pb // 8
    out = plane.astype(np.float64) if return_float else plane.copy()
    t = plane[:by * pb, :bx * pb].astype(np.float64)
    t = t.reshape(by * r, 8, bx * r, 8).swapaxes(1, 2)
    c = dctn(t, type=2, norm="ortho", axes=(2, 3))
    g = gains(np.repeat(np.repeat(np.asarray(strength)[:by, :bx], r, axis=0), r, axis=1))
    rec = idctn(c * g, type=2, norm="ortho", axes=(2, 3))
    rec = rec.swapaxes(1, 2).reshape(by * pb, bx * pb)
    if return_float:
        out[:by * pb, :bx * pb] = rec
        return out
    out[:by * pb, :bx * pb] = np.clip(np.rint(rec), 0, 255).astype(np.uint8)
    return out
