"""CPU check of elvis_b200/csrc/down_pow2.cuh (transliterated) against oracle/spec_cv.down_up."""
import os
import sys

import numpy as np

sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", ".."), os.path.dirname(__file__)]
from oracle import spec_cv  # noqa: E402
from warp_emu import LANES, add, byte_perm, mul, shfl, shfl_xor, u32  # noqa: E402


def pk_lo(x):
    return x & 0xFFFF


def pk_pair(lo, hi):
    return byte_perm(lo, hi, 0x5410)


def v_pair(s2l, a, ka, b, kb):
    mask = u32((0xFFFF >> s2l) * 0x00010001)
    x = (mul(a, ka) >> s2l) & mask
    y = (mul(b, kb) >> s2l) & mask
    return add(x, y, 0x00020002) >> 2


def rhe_pair(K, s):
    half = ((1 << (K - 1)) - 1) * 0x00010001
    mask = u32((0xFFFF >> K) * 0x00010001)
    odd = (s >> K) & 0x00010001
    return (add(s, half, odd) >> K) & mask


def down_up_pow2(PB, L, p0, p1, g, base):
    h, r = g & 1, g >> 1
    f = 1 << L
    e0 = add(p0 & 0x00FF00FF, (p0 >> 8) & 0x00FF00FF)
    e1 = add(p1 & 0x00FF00FF, (p1 >> 8) & 0x00FF00FF) if PB == 16 else np.zeros(32, np.uint32)
    if (PB == 16 and L == 4) or (PB == 8 and L == 3):
        s = add(e0 & 0xFFFF, e0 >> 16, e1 & 0xFFFF, e1 >> 16)
        m = 1
        while m < 2 * PB:
            s = add(s, shfl_xor(s, m))
            m <<= 1
        K = 2 * L
        v = add(s, (1 << (K - 1)) - 1, (s >> K) & 1) >> K
        p0 = mul(v, 0x01010101)
        return p0, p0.copy()
    E1 = np.zeros(32, np.uint32)
    if L == 1:
        e0 = add(e0, shfl_xor(e0, 2))
        if PB == 16:
            e1 = add(e1, shfl_xor(e1, 2))
        E0 = (add(e0, 0x00020002) >> 2) & 0x00FF00FF
        E1 = (add(e1, 0x00020002) >> 2) & 0x00FF00FF
    elif L == 2:
        if PB == 16:
            s = pk_pair(add(e0 & 0xFFFF, e0 >> 16), add(e1 & 0xFFFF, e1 >> 16))
        else:
            s = add(e0 & 0xFFFF, e0 >> 16)
        s = add(s, shfl_xor(s, 2))
        s = add(s, shfl_xor(s, 4))
        E0 = rhe_pair(4, s)
    else:
        s = add(e0 & 0xFFFF, e0 >> 16, e1 & 0xFFFF, e1 >> 16)
        for m in (2, 4, 8):
            s = add(s, shfl_xor(s, m))
        E0 = rhe_pair(6, s)
    cells = (PB // 2) >> L
    first = pk_lo(E0)
    last = (E1 >> 16) if cells == 4 else ((E0 >> 16) if cells == 2 else pk_lo(E0))
    nb = shfl_xor(np.where(h == 1, first, last).astype(np.uint32), 1)
    left = np.where(h == 1, nb, first).astype(np.uint32)
    right = np.where(h == 1, last, nb).astype(np.uint32)
    if PB == 16 and L == 1:
        T0, T1 = mul(E0, 3), mul(E1, 3)
        mid = byte_perm(E0, E1, 0x5432)
        t = [add(pk_pair(left, E0), T0), add(mid, T1), add(T0, mid), add(T1, byte_perm(E1, right, 0x5432))]
    elif PB == 16 and L == 2:
        A = pk_pair(left, E0)
        B = byte_perm(E0, right, 0x5432)
        t = [add(mul(A, 3), mul(E0, 5)), add(A, mul(E0, 7)), add(mul(E0, 7), B), add(mul(E0, 5), mul(B, 3))]
    elif PB == 16 and L == 3:
        LR, SS = pk_pair(left, right), mul(first, 0x00010001)
        t = [add(mul(LR, 7), mul(SS, 9)), add(mul(LR, 5), mul(SS, 11)), add(mul(LR, 3), mul(SS, 13)), add(LR, mul(SS, 15))]
    elif PB == 8 and L == 1:
        T0 = mul(E0, 3)
        t = [add(pk_pair(left, E0), T0), add(T0, byte_perm(E0, right, 0x5432))]
    else:
        LR, SS = pk_pair(left, right), mul(first, 0x00010001)
        t = [add(mul(LR, 3), mul(SS, 5)), add(LR, mul(SS, 7))]
    q = r & (f - 1)
    up = q < f // 2
    src = g + np.where(up, -2 * f, 2 * f)
    src = np.where((src < 0) | (src >= 2 * PB), g, src)
    c1 = (2 * r + 1 - f) & (2 * f - 1)
    k_other = np.where(up, 2 * f - c1, c1)
    k_own = 2 * f - k_other
    v = [v_pair(2 * L, ti, k_own, shfl(ti, base + src), k_other) for ti in t]
    if PB == 16 and L == 1:
        return byte_perm(v[0], v[2], 0x6240), byte_perm(v[1], v[3], 0x6240)
    if PB == 16 and L == 2:
        x, y = byte_perm(v[0], v[1], 0x6240), byte_perm(v[2], v[3], 0x6240)
        return byte_perm(x, y, 0x5410), byte_perm(x, y, 0x7632)
    if PB == 16:
        x, y = byte_perm(v[0], v[1], 0x2640), byte_perm(v[2], v[3], 0x2640)
        return byte_perm(x, y, 0x5410), byte_perm(x, y, 0x3276)
    if L == 1:
        return byte_perm(v[0], v[1], 0x6240), p1
    return byte_perm(v[0], v[1], 0x2640), p1


def run(PB, L, blocks):
    """blocks: list of PB x PB uint8 blocks filling the warp (1 for PB 16, 2 for PB 8)."""
    group = 2 * PB
    g = LANES % group
    base = LANES - g
    px = PB // 2
    p0, p1 = np.zeros(32, np.uint32), np.zeros(32, np.uint32)
    for lane in range(32):
        blk = blocks[lane // group]
        r, h = (lane % group) >> 1, lane & 1
        seg = blk[r, px * h:px * h + px].astype(np.uint32)
        p0[lane] = seg[0] | (seg[1] << 8) | (seg[2] << 16) | (seg[3] << 24)
        if PB == 16:
            p1[lane] = seg[4] | (seg[5] << 8) | (seg[6] << 16) | (seg[7] << 24)
    p0, p1 = down_up_pow2(PB, L, p0, p1, g, base)
    outs = [np.zeros((PB, PB), np.uint8) for _ in blocks]
    for lane in range(32):
        r, h = (lane % group) >> 1, lane & 1
        words = [p0[lane]] + ([p1[lane]] if PB == 16 else [])
        seg = [(w >> (8 * k)) & 0xFF for w in words for k in range(4)]
        outs[lane // group][r, px * h:px * h + px] = seg
    return outs


def main():
    rng = np.random.default_rng(0)
    for PB in (16, 8):
        for L in range(1, int(np.log2(PB)) + 1):
            bad = 0
            for it in range(300):
                nblk = 1 if PB == 16 else 2
                blocks = [rng.integers(0, 256, (PB, PB), dtype=np.uint8) if it % 3 else (rng.integers(0, 2, (PB, PB)) * 255).astype(np.uint8)
                          for _ in range(nblk)]
                got = run(PB, L, blocks)
                for b, o in zip(blocks, got):
                    bad += not np.array_equal(o, spec_cv.down_up(b, PB >> L))
            print(f"PB {PB} L {L}: mismatching blocks {bad}")
            assert bad == 0


if __name__ == "__main__":
    main()
