"""CPU check of the tensor-core formulation of the DCT dampening (dampen_hmma in dampen.cu): the gain
g(u, v) = q^(u+v) is separable, so dampening an 8 x 8 tile is X' = M X M^T with M = A^T diag(q^u) A
(A = orthonormal DCT-II) -- two small matrix products per tile with a per-block operator.  Emulates
mma.sync.m16n8k16 (f16 x f16 -> f32) with the PTX fragment layouts, the hi + lo f16 splits and the
operand chaining, against oracle/spec_dct_dampen.dampen_plane."""
import os
import sys

import numpy as np

sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", ".."), os.path.dirname(__file__)]
from oracle import spec_dct_dampen  # noqa: E402

F = np.float32


def dct_matrix():
    x = np.arange(8)
    A = np.cos((2 * x[None, :] + 1) * x[:, None] * np.pi / 16) * 0.5
    A[0] *= np.sqrt(0.5)
    return A


def rho(i):
    t, e = (i & 7) >> 1, i & 1
    return 4 * t + e + (2 if i >= 8 else 0)


def split16(x):
    hi = x.astype(np.float16)
    lo = (x.astype(F) - hi.astype(F)).astype(np.float16)
    return hi, lo


def hmma(Ah, B, C):
    """D = A B + C with f16 operands (exact products) and fp32 accumulation."""
    return (Ah.astype(np.float64) @ B.astype(np.float64)).astype(F) + C


def operator(s):
    """M(s) in fp32 the way the kernel builds it: sum_u q^u a_u a_u^T, q = 2^(-4 s / 14)."""
    A = dct_matrix().astype(F)
    q = F(np.exp2(F(-4.0) * F(s) / F(14.0)))
    M = np.zeros((8, 8), F)
    p = F(1.0)
    for u in range(8):
        M += p * np.outer(A[u], A[u]).astype(F)
        p = F(p * q)
    return M


def dampen_tile(tile, s):
    """16 x 16 uint8 tile = 2 x 2 DCT tiles sharing the strength s."""
    M = operator(s)
    Mbd = np.zeros((16, 16), F)
    Mbd[:8, :8] = M
    Mbd[8:, 8:] = M
    perm = np.array([rho(i) for i in range(16)])
    Mt = Mbd[np.ix_(perm, perm)]                       # relabelled operator: layout index -> pixel index
    Xt = (tile.astype(F) - F(128))[np.ix_(perm, perm)]  # what the threads hold, in layout coordinates (exact in f16)
    Mh, Ml = split16(Mt)
    # step 1: C1 = M X^T  (B operand = the held matrix read transposed)
    B1 = Xt.T.astype(np.float16)
    C1 = hmma(Ml, B1, hmma(Mh, B1, np.zeros((16, 16), F)))
    # step 2: Z = M C1^T with C1 split hi + lo, dropping lo x lo
    Yh, Yl = split16(C1.T)
    Z = hmma(Mh, Yl, hmma(Ml, Yh, hmma(Mh, Yh, np.zeros((16, 16), F))))
    out_layout = np.clip(np.rint(Z + F(128)), 0, 255).astype(np.uint8)
    out = np.zeros((16, 16), np.uint8)
    out[np.ix_(perm, perm)] = out_layout
    return out, Z + F(128)


def main():
    rng = np.random.default_rng(2)
    worst, off = 0.0, 0
    for it in range(300):
        kind = it % 3
        tile = rng.integers(0, 256, (16, 16), dtype=np.uint8) if kind == 0 else \
            ((rng.integers(0, 2, (16, 16)) * 255).astype(np.uint8) if kind == 1 else np.clip(rng.normal(128, 30, (16, 16)), 0, 255).astype(np.uint8))
        s = float(rng.random()) if it % 7 else float(it % 2)
        got, _ = dampen_tile(tile, s)
        ref_f = spec_dct_dampen.dampen_plane(tile, np.array([[s]]), 16, return_float=True)
        ref = spec_dct_dampen.dampen_plane(tile, np.array([[s]]), 16)
        off += int((got != ref).sum())
        worst = max(worst, float(np.abs(got.astype(np.float64) - np.clip(ref_f, 0, 255)).max()))
    print(f"pixels differing from the rounded spec: {off} of {300 * 256}; worst |u8 - float64 reconstruction| = {worst:.4f} (bar 0.5 + 0.0255)")
    assert worst <= 0.5 + 1e-4 * 255


if __name__ == "__main__":
    main()
