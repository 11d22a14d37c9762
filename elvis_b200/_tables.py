"""Host-side coefficient tables for elvis_degrade_downsample (layout documented in
csrc/downsample.cu).  The arithmetic follows OpenCV's resize setup code for 8-bit images:
the bilinear taps are float32 positions quantised to 11 bits, the fractional INTER_AREA
weights are float32 overlaps.  tests/test_tables.py compares these against the
independent restatement in oracle/spec_cv.py."""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

KIND_COPY, KIND_2X2, KIND_INT, KIND_FRAC = 0, 1, 2, 3


def level_stride(pb: int, lanczos: bool = False) -> int:
    return 8 + (pb + 1) + 4 * pb + (16 if lanczos else 8) * pb


def _lanczos4_taps(ssize: int, dsize: int):
    """cv2 INTER_LANCZOS4 (u8): (dsize, 8) clamped source indices and 11-bit integer weights."""
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    idx = np.zeros((dsize, 8), np.int32)
    coef = np.zeros((dsize, 8), np.int32)
    f32 = np.float32
    for d in range(dsize):
        pos = (d + 0.5) * (ssize / dsize) - 0.5
        base = math.floor(pos)
        x = f32(pos - base)
        w = np.zeros(8, np.float32)
        if x < np.finfo(np.float32).eps:
            w[3] = 1
        else:
            y0 = f32(-(x + 3) * f32(np.pi) * f32(0.25))
            s0, c0 = f32(np.sin(y0)), f32(np.cos(y0))
            total = f32(0)
            for i in range(8):
                y = f32(-(x + 3 - i) * f32(np.pi) * f32(0.25))
                w[i] = f32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
                total = f32(total + w[i])
            w = (w * (f32(1.0) / total)).astype(np.float32)
        coef[d] = np.clip(np.rint(w * f32(2048)), -32768, 32767).astype(np.int32)
        idx[d] = np.clip(np.arange(base - 3, base + 5), 0, ssize - 1)
    return idx, coef


def fast_stride(pb: int) -> int:
    """Words per level of the byte-weight tables used by downsample_fast_kernel:
    pb x {wx, wy} horizontal weight vectors, then pb x {i0, i1, b0, b1} vertical taps."""
    return 2 * pb + 4 * pb


def _linear_taps(ssize: int, dsize: int, horizontal: bool):
    d = np.arange(dsize, dtype=np.float64)
    pos = ((d + 0.5) * (ssize / dsize) - 0.5).astype(np.float32)
    base = np.floor(pos).astype(np.int64)
    frac = (pos - base.astype(np.float32)).astype(np.float32)
    if horizontal:   # taps AND weights collapse at the borders
        lo = base < 0
        hi = base >= ssize - 1
        frac = np.where(lo | hi, np.float32(0), frac)
        base = np.where(lo, 0, np.where(hi, ssize - 1, base))
        i0, i1 = base, np.minimum(base + 1, ssize - 1)
    else:            # only the row indices are clamped
        i0, i1 = np.clip(base, 0, ssize - 1), np.clip(base + 1, 0, ssize - 1)
    c0 = np.rint((np.float32(1) - frac) * np.float32(2048)).astype(np.int32)
    c1 = np.rint(frac * np.float32(2048)).astype(np.int32)
    return i0.astype(np.int32), i1.astype(np.int32), c0, c1


def _area_entries(ssize: int, dsize: int):
    """(start[dsize+1], src_index[], alpha float32[]) of the fractional area kernel."""
    scale = ssize / dsize
    start, src, alpha = [0], [], []
    for d in range(dsize):
        a, b = d * scale, d * scale + scale
        cell = min(scale, ssize - a)
        first, last = math.ceil(a), min(math.floor(b), ssize - 1)
        first = min(first, last)
        if first - a > 1e-3:
            src.append(first - 1)
            alpha.append((first - a) / cell)
        for s in range(first, last):
            src.append(s)
            alpha.append(1.0 / cell)
        if b - last > 1e-3:
            src.append(last)
            alpha.append(min(min(b - last, 1.0), cell) / cell)
        start.append(len(src))
    return np.array(start, np.int32), np.array(src, np.int32), np.array(alpha, np.float32)


@lru_cache(maxsize=64)
def build(pb: int, small_sizes: tuple, lanczos: bool = False) -> np.ndarray:
    """int32 blob with one entry per level; small_sizes[level] is the side the block is
    reduced to (== pb: the level copies the block).  lanczos=True: the up-sampling taps are the
    8-tap INTER_LANCZOS4 table instead of the two bilinear tap blocks (no fast part)."""
    stride = level_stride(pb, lanczos)
    blob = np.zeros((len(small_sizes), stride), np.int32)
    for lv, small in enumerate(small_sizes):
        small = int(small)
        e = blob[lv]
        e[0] = small
        if small >= pb:
            e[1] = KIND_COPY
            continue
        if pb % small == 0:
            f = pb // small
            e[1] = KIND_2X2 if f == 2 else KIND_INT
            e[2] = f
            e[3] = np.array([np.float32(1.0) / np.float32(f * f)], np.float32).view(np.int32)[0]
        else:
            e[1] = KIND_FRAC
            start, src, alpha = _area_entries(pb, small)
            assert len(src) <= 2 * pb
            e[4] = len(src)
            e[8:8 + small + 1] = start
            ent = e[8 + pb + 1: 8 + pb + 1 + 4 * pb].reshape(2 * pb, 2)
            ent[:len(src), 0] = src
            ent[:len(src), 1] = alpha.view(np.int32)
        off = 8 + pb + 1 + 4 * pb
        if lanczos:
            idx, coef = _lanczos4_taps(small, pb)
            e[off:off + 8 * pb] = idx.reshape(-1)
            e[off + 8 * pb:off + 16 * pb] = coef.reshape(-1)
            continue
        for horizontal in (True, False):
            i0, i1, c0, c1 = _linear_taps(small, pb, horizontal)
            e[off:off + pb] = i0
            e[off + pb:off + 2 * pb] = i1
            e[off + 2 * pb:off + 3 * pb] = c0
            e[off + 3 * pb:off + 4 * pb] = c1
            off += 4 * pb
    if lanczos:
        return blob.reshape(-1)
    # fast-path tables (planar 8/16-pixel blocks, power-of-two reductions): when every 11-bit
    # horizontal coefficient of a level is a multiple of 64 the two taps of an output pixel
    # become byte weights over the (<= 8 byte) source row, so that
    #     S[i0]*a0 + S[i1]*a1  ==  64 * (dp4a(row.lo, wx) + dp4a(row.hi, wy))
    fast = np.zeros((len(small_sizes), fast_stride(pb)), np.int32)
    for lv, small in enumerate(small_sizes):
        small = int(small)
        ok = small < pb and pb % small == 0 and (pb // small) & (pb // small - 1) == 0 and small <= 8 and pb in (8, 16)
        if small >= pb:
            ok = True          # copy level: nothing to tabulate
        if ok and small < pb:
            i0, i1, c0, c1 = _linear_taps(small, pb, True)
            ok = bool(np.all(c0 % 64 == 0) and np.all(c1 % 64 == 0))
            if ok:
                w = np.zeros((pb, 8), np.uint8)
                for d in range(pb):
                    w[d, i0[d]] += c0[d] // 64
                    w[d, i1[d]] += c1[d] // 64
                fast[lv, :2 * pb] = w.view(np.uint32).view(np.int32).reshape(-1)
                v = np.stack(_linear_taps(small, pb, False), axis=1).astype(np.int32)     # (pb, 4)
                fast[lv, 2 * pb:] = v.reshape(-1)
        blob[lv, 5] = 1 if ok else 0
    generic = blob.reshape(-1)
    pad = (-len(generic)) % 4                 # the fast part is read with 128-bit loads
    return np.concatenate([generic, np.zeros(pad, np.int32), fast.reshape(-1)])


def fast_levels_flag(pb: int, small_sizes: tuple) -> int:
    """The `fast_tables_ok` argument of elvis_degrade_downsample: 1 = every level has fast-path tables, 0 = none (or too
    many levels to describe), otherwise an even value with bit l + 1 set for every level l WITHOUT them."""
    blob = build(pb, tuple(small_sizes))
    st = level_stride(pb)
    if pb not in (8, 16):
        return 0
    fast = [int(blob[i * st + 5]) == 1 for i in range(len(small_sizes))]
    if all(fast):
        return 1
    if not any(fast) or len(fast) > 30:
        return 0
    return sum(1 << (i + 1) for i, f in enumerate(fast) if not f)


def all_fast(pb: int, small_sizes: tuple) -> bool:
    """True when every level of the blob has fast-path (byte-weight) tables."""
    blob = build(pb, tuple(small_sizes))
    st = level_stride(pb)
    return pb in (8, 16) and all(int(blob[i * st + 5]) == 1 for i in range(len(small_sizes)))


def gaussian_kernels(max_level: int) -> np.ndarray:
    """int32 (max_level + 1, 6 * max_level + 2): row L = {ksize, q[0..ksize)} -- cv2's 8.8
    fixed-point Gaussian for sigma = L and ksize = (0, 0) on 8-bit images: the exact Gaussian,
    normalised, times 256, rounded with error diffusion over the first half (centre included) and
    mirrored, so that it sums to 256.  Row 0 is unused."""
    stride = 6 * max_level + 2
    tab = np.zeros((max_level + 1, stride), np.int32)
    for level in range(1, max_level + 1):
        ksize = int(round(level * 6 + 1)) | 1
        r = ksize // 2
        x = np.arange(-r, r + 1, dtype=np.float64)
        k = np.exp(-(x * x) / (2.0 * level * level))
        k /= k.sum()
        carry = 0.0
        q = np.zeros(ksize, np.int64)
        for i in range(r + 1):
            v = k[i] * 256 + carry
            q[i] = math.floor(v + 0.5)
            carry = v - q[i]
        q[r + 1:] = q[:r][::-1]
        tab[level, 0] = ksize
        tab[level, 1:1 + ksize] = q
    return tab


@lru_cache(maxsize=64)
def area_f32_tables(ssize: int, dsize: int):
    """cv2's decimation table of a float INTER_AREA resize along one axis, as the three arrays
    elvis_resize_area_f32 takes (start[dsize + 1] int32, source index int32, weight float32)."""
    return _area_entries(ssize, dsize)


def area_f32_plan(sh: int, sw: int, dh: int, dw: int):
    """(int_scale_x, int_scale_y, simd_cols) of cv2.resize(float32, INTER_AREA): both 0 unless both
    ratios are integers, in which case cv2 takes its window-sum path; its 2 x 2 case computes the
    leading multiple of 4 destination columns with a 4-lane vector kernel that adds in a different
    order (oracle/spec_cv.py resize_area_f32, pinned against the cv2 of this image)."""
    sx, sy = sw / dw, sh / dh
    ix, iy = int(round(sx)), int(round(sy))
    eps = np.finfo(np.float64).eps
    if abs(sx - ix) < eps and abs(sy - iy) < eps:
        return ix, iy, (dw // 4) * 4 if (ix == 2 and iy == 2) else 0
    return 0, 0, 0


@lru_cache(maxsize=64)
def nearest_index(ssize: int, dsize: int) -> np.ndarray:
    """Source index per destination index of cv2.resize(INTER_NEAREST): cv2 multiplies by
    1 / (dsize / ssize), not by ssize / dsize -- the two differ in the last bit for some ratios."""
    ifx = 1.0 / (dsize / ssize)
    return np.minimum(np.floor(np.arange(dsize) * ifx).astype(np.int64), ssize - 1).astype(np.int32)


@lru_cache(maxsize=64)
def linear_float_index(ssize: int, dsize: int, fused: bool):
    """Source index (int32) and fraction (float64) per destination index of cv2.resize(float map,
    INTER_LINEAR): coordinate = (d + 0.5) * (ssize / dsize) - 0.5 in double -- one fused multiply-add on
    cv2's float64 path (`fused`; evaluated exactly with rationals here), a rounded product and a rounded
    difference on its float32 path -- index = floor, fraction = coordinate - index; left of the first /
    right of the last sample the fraction is 0."""
    from fractions import Fraction
    scale = ssize / dsize
    idx = np.empty(dsize, np.int32)
    frac = np.empty(dsize, np.float64)
    for d in range(dsize):
        fx = float(Fraction(2 * d + 1, 2) * Fraction(scale) - Fraction(1, 2)) if fused else (d + 0.5) * scale - 0.5
        s = int(np.floor(fx))
        f = fx - s
        if s < 0:
            s, f = 0, 0.0
        if s >= ssize - 1:
            s, f = ssize - 1, 0.0
        idx[d], frac[d] = s, f
    return idx, frac
