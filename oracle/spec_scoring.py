"""TEST INFRASTRUCTURE ONLY -- normative spec of the per-block SC/TC features.

PARITY UNPINNED: in the reference this arithmetic lives in the third-party package
`evca` (requirements.txt:60, `git+https://github.com/emanuele-artioli/EVCA.git`, no commit
pin), invoked at elvis.py:1014-1031 (`python -m evca.main -i raw.yuv -r WxH -b bs ...`)
and presley.py:202 (`analyze_frames(frames, EVCAConfig(block_size=bs))`).  The package is
not vendored, not installed in this image and the reference holds no test vectors for it,
so this file *defines* the features (VCA/EVCA family: DCT-energy texture + coefficient-
level temporal difference) and the CUDA kernel is checked against it, not against EVCA.

Definition (luma only, float64):
  * the (H, W) luma plane is cropped to whole bs x bs blocks; every block is tiled into
    N x N transform tiles, N = dct_size = 8;
  * C_t = orthonormal 2-D DCT-II of a tile of frame t (pixel values 0..255, no offset);
  * weight  w(u, v) = exp(|(u*v / N^2)^2 - 1|),  w(0, 0) = 0  (DC excluded);
  * SC[t, by, bx] = sum over the block's tiles, sum_{u,v} w(u,v) * |C_t(u,v)|              / bs^2
  * TC[t, by, bx] = sum over the block's tiles, sum_{u,v} w(u,v) * |C_t(u,v) - C_{t-1}(u,v)| / bs^2
    with TC[0] = 0 unless a `prev` halo frame (the frame before y[0]) is supplied.
"""
from __future__ import annotations

import numpy as np
from scipy.fft import dctn

DCT_SIZE = 8


def weights(n: int = DCT_SIZE) -> np.ndarray:
    u = np.arange(n, dtype=np.float64)
    w = np.exp(np.abs((np.outer(u, u) / (n * n)) ** 2 - 1.0))
    w[0, 0] = 0.0
    return w


def tile_coeffs(y: np.ndarray, bs: int, n: int = DCT_SIZE) -> np.ndarray:
    """(H, W) uint8 -> (By*bs/n, Bx*bs/n, n, n) float64 orthonormal DCT-II coefficients."""
    h, w = y.shape
    by, bx = h // bs, w // bs
    t = y[:by * bs, :bx * bs].astype(np.float64)
    t = t.reshape(by * bs // n, n, bx * bs // n, n).swapaxes(1, 2)
    return dctn(t, type=2, norm="ortho", axes=(2, 3))


def sc_tc(y: np.ndarray, bs: int, n: int = DCT_SIZE, prev: np.ndarray | None = None):
    """y: (T, H, W) uint8 luma.  Returns (SC, TC), each (T, By, Bx) float64."""
    if bs % n:
        raise ValueError("block_size must be a multiple of dct_size")
    T, h, w = y.shape
    by, bx = h // bs, w // bs
    r = bs // n
    wt = weights(n)
    sc = np.zeros((T, by, bx))
    tc = np.zeros((T, by, bx))
    last = tile_coeffs(prev, bs, n) if prev is not None else None
    for t in range(T):
        c = tile_coeffs(y[t], bs, n)
        e = (np.abs(c) * wt).sum(axis=(2, 3))
        sc[t] = e.reshape(by, r, bx, r).sum(axis=(1, 3)) / (bs * bs)
        if last is not None:
            d = (np.abs(c - last) * wt).sum(axis=(2, 3))
            tc[t] = d.reshape(by, r, bx, r).sum(axis=(1, 3)) / (bs * bs)
        last = c
    return sc, tc
