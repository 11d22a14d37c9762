"""CPU check of the tensor-core formulation of the per-block Gaussian blur (blur_imma in blur.cu),
transliterated: Z = G X G^T per round as three groups of mma.sync.m16n8k16 (u8 x u8 -> s32) whose
operands chain without data movement, against oracle/spec_cv.gaussian_blur5_rounds."""
import os
import sys

import numpy as np

sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", ".."), os.path.dirname(__file__)]
from oracle import spec_cv  # noqa: E402
from warp_emu import LANES, byte_perm, u32  # noqa: E402

G_, TQ = LANES >> 2, LANES & 3


def mma_u8(a0, a1, b0, c):
    """mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 with the PTX fragment layouts."""
    A = np.zeros((16, 16), np.int64)
    B = np.zeros((16, 8), np.int64)
    C = np.zeros((16, 8), np.int64)
    for lane in range(32):
        g, tq = lane >> 2, lane & 3
        for i in range(4):
            A[g, 4 * tq + i] = (int(a0[lane]) >> (8 * i)) & 0xFF
            A[g + 8, 4 * tq + i] = (int(a1[lane]) >> (8 * i)) & 0xFF
            B[4 * tq + i, g] = (int(b0[lane]) >> (8 * i)) & 0xFF
        C[g, 2 * tq], C[g, 2 * tq + 1] = np.int32(c[0][lane]), np.int32(c[1][lane])
        C[g + 8, 2 * tq], C[g + 8, 2 * tq + 1] = np.int32(c[2][lane]), np.int32(c[3][lane])
    D = A @ B + C
    out = [np.zeros(32, np.uint32) for _ in range(4)]
    for lane in range(32):
        g, tq = lane >> 2, lane & 3
        out[0][lane], out[1][lane] = D[g, 2 * tq] & 0xFFFFFFFF, D[g, 2 * tq + 1] & 0xFFFFFFFF
        out[2][lane], out[3][lane] = D[g + 8, 2 * tq] & 0xFFFFFFFF, D[g + 8, 2 * tq + 1] & 0xFFFFFFFF
    return out


def rho(i):
    """layout index -> pixel index inside the 16-wide tile (same map for rows and columns)."""
    t, e = (i & 7) >> 1, i & 1
    return 4 * t + e + (2 if i >= 8 else 0)


def gmat(PB):
    """16 x 16 operator of one blur pass along an axis of the tile: 5 taps, reflect-101 at the block edge."""
    taps = [14, 62, 104, 62, 14]
    G = np.zeros((16, 16), np.int64)
    for m in range(16):
        blk, ml = (m // PB) * PB, m % PB
        for d in range(-2, 3):
            G[m, blk + spec_cv._reflect101(ml + d, PB)] += taps[d + 2]
    return G


def blur_tile(tile, PB, rounds_q):
    """tile 16 x 16 uint8; rounds_q[qy][qx] rounds of every PB x PB block of the tile."""
    G = gmat(PB)
    r0 = np.array([rho(g) for g in G_])
    r1 = np.array([rho(g + 8) for g in G_])
    word = lambda rows: np.array([sum(int(tile[rows[l], 4 * TQ[l] + i]) << (8 * i) for i in range(4)) for l in range(32)], np.uint32)  # noqa: E731
    a0 = np.array([sum(int(G[r0[l], 4 * TQ[l] + i]) << (8 * i) for i in range(4)) for l in range(32)], np.uint32)
    a1 = np.array([sum(int(G[r1[l], 4 * TQ[l] + i]) << (8 * i) for i in range(4)) for l in range(32)], np.uint32)
    w0, w1 = word(r0), word(r1)
    qy, qx = (G_ >= (8 // (16 // PB) if PB == 8 else 99)).astype(int), (TQ >= 2).astype(int)
    if PB == 16:
        nr = np.full(32, rounds_q[0][0])
    else:
        nr = np.array([rounds_q[int(G_[l] >= 4)][int(TQ[l] >= 2)] for l in range(32)])
    zero = [np.zeros(32, np.uint32)] * 4
    init = [np.full(32, 128, np.uint32)] * 4
    for k in range(int(nr.max())):
        c1 = [mma_u8(a0, a1, w0, zero), mma_u8(a0, a1, w1, zero)]           # M1 = G X^T, n-tiles 0 and 1
        new = []
        for half in (0, 1):                                                   # layout rows g / g + 8 of M1 -> n-tile of step 2
            q = [c1[0][2 * half], c1[0][2 * half + 1], c1[1][2 * half], c1[1][2 * half + 1]]
            hi = byte_perm(byte_perm(q[0], q[1], 0x0051), byte_perm(q[2], q[3], 0x0051), 0x5410)
            lo = byte_perm(byte_perm(q[0], q[1], 0x0040), byte_perm(q[2], q[3], 0x0040), 0x5410)
            acc = mma_u8(a0, a1, hi, init)
            acc = [u32(x.astype(np.uint64) << np.uint64(8)) for x in acc]
            new.append(mma_u8(a0, a1, lo, acc))
        z0 = byte_perm(byte_perm(new[0][0], new[0][1], 0x0062), byte_perm(new[1][0], new[1][1], 0x0062), 0x5410)
        z1 = byte_perm(byte_perm(new[0][2], new[0][3], 0x0062), byte_perm(new[1][2], new[1][3], 0x0062), 0x5410)
        w0 = np.where(k < nr, z0, w0).astype(np.uint32)
        w1 = np.where(k < nr, z1, w1).astype(np.uint32)
    out = np.zeros((16, 16), np.uint8)
    for l in range(32):
        for i in range(4):
            out[r0[l], 4 * TQ[l] + i] = (int(w0[l]) >> (8 * i)) & 0xFF
            out[r1[l], 4 * TQ[l] + i] = (int(w1[l]) >> (8 * i)) & 0xFF
    return out


def main():
    rng = np.random.default_rng(1)
    for PB in (16, 8):
        bad = 0
        for it in range(40):
            tile = rng.integers(0, 256, (16, 16), dtype=np.uint8) if it % 3 else (rng.integers(0, 2, (16, 16)) * 255).astype(np.uint8)
            n = 16 // PB
            rq = rng.integers(0, 5, (n, n))
            got = blur_tile(tile, PB, rq)
            for qy in range(n):
                for qx in range(n):
                    blk = tile[qy * PB:(qy + 1) * PB, qx * PB:(qx + 1) * PB]
                    ref = spec_cv.gaussian_blur5_rounds(blk, int(rq[qy][qx]))
                    bad += not np.array_equal(ref, got[qy * PB:(qy + 1) * PB, qx * PB:(qx + 1) * PB])
        print(f"PB {PB}: mismatching blocks {bad}")
        assert bad == 0


if __name__ == "__main__":
    main()
