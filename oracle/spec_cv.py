"""TEST INFRASTRUCTURE ONLY -- integer restatement of the three OpenCV primitives the
reference's v2 degradations delegate to (parity pinned: checked bit-exact against the
real `cv2` of this image, 4.13.0, in tests/test_oracle.py; the reference pins
opencv-python 4.8.0.76 -- requirements.txt:42).

  * cv2.GaussianBlur(block, (5, 5), sigmaX=1.0)        elvis.py:2190, utils.py:1209, presley.py:989
  * cv2.resize(block, (s, s), interpolation=INTER_AREA)   elvis.py:2161, utils.py:1160, presley.py:982
  * cv2.resize(small, (bs, bs), interpolation=INTER_LINEAR) elvis.py:2163, utils.py:1161, presley.py:983

All functions take single-channel uint8 blocks shaped (..., h, w) -- leading axes are a
batch of independent blocks (cv2 treats channels independently for these three
operations) -- and return uint8.  The CUDA kernels in
elvis_b200/csrc/degrade.cu mirror these formulas; the coefficient tables they consume
are produced by elvis_b200/_tables.py with the same arithmetic as `linear_coeffs` /
`area_table` below (two independent implementations, compared in the tests).
"""
from __future__ import annotations

import numpy as np

GAUSS5_SIGMA1_Q8 = np.array([14, 62, 104, 62, 14], dtype=np.int64)  # sums to 256


def _reflect101(i: int, n: int) -> int:
    """cv2.BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba)."""
    if n == 1:
        return 0
    while i < 0 or i >= n:
        if i < 0:
            i = -i
        if i >= n:
            i = 2 * (n - 1) - i
    return i


def gaussian_blur5(block: np.ndarray) -> np.ndarray:
    """One round of the 5x5 sigma=1 blur on an isolated block (border at the block edge).

    Fixed-point separable filter: horizontal pass exact in 8.8, vertical pass exact in
    16.16, one final rounding (acc + 2^15) >> 16.
    """
    h, w = block.shape[-2:]
    src = block.astype(np.int64)
    cols = np.array([[_reflect101(x + d, w) for d in (-2, -1, 0, 1, 2)] for x in range(w)])
    rows = np.array([[_reflect101(y + d, h) for d in (-2, -1, 0, 1, 2)] for y in range(h)])
    hp = (src[..., :, cols] * GAUSS5_SIGMA1_Q8).sum(axis=-1)                     # (..., h, w)
    vp = (hp[..., rows, :] * GAUSS5_SIGMA1_Q8[:, None]).sum(axis=-2)
    return ((vp + 32768) >> 16).astype(np.uint8)


def gaussian_blur5_rounds(block: np.ndarray, rounds: int) -> np.ndarray:
    out = block
    for _ in range(int(rounds)):
        out = gaussian_blur5(out)
    return out


# ----------------------------------------------------------------------------- INTER_AREA
def area_table(ssize: int, dsize: int):
    """Per-axis overlap table of cv2's generic area resize: list of (si, di, alpha) with
    alpha a float32 weight.  scale = ssize/dsize (double); for each dst cell the source
    span [dx*scale, (dx+1)*scale) is split into a leading partial cell, whole cells and
    a trailing partial cell; weights are divided by min(scale, ssize - fsx1)."""
    scale = float(ssize) / float(dsize)
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = int(np.ceil(fsx1))
        sx2 = int(np.floor(fsx2))
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((sx1 - 1, dx, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((sx, dx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((sx2, dx, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area(block: np.ndarray, dsize: int) -> np.ndarray:
    """cv2.resize(block, (dsize, dsize), interpolation=cv2.INTER_AREA) for a square block."""
    ssize = block.shape[-1]
    assert block.shape[-2] == block.shape[-1]
    lead = block.shape[:-2]
    if dsize == ssize:
        return block.copy()
    if ssize % dsize == 0:
        f = ssize // dsize
        s = block.astype(np.int64).reshape(lead + (dsize, f, dsize, f)).sum(axis=(-3, -1))
        if f == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        # integer factor > 2: sum * float32(1/area), round half to even
        scale = np.float32(1.0) / np.float32(f * f)
        return np.rint(s.astype(np.float32) * scale).astype(np.uint8)
    # generic fractional scale: float32 accumulation in table order, rows then columns
    tab = area_table(ssize, dsize)
    buf = np.zeros(lead + (ssize, dsize), dtype=np.float32)
    srcf = block.astype(np.float32)
    for (si, di, a) in tab:
        buf[..., :, di] = buf[..., :, di] + srcf[..., :, si] * a
    out = np.zeros(lead + (dsize, dsize), dtype=np.float32)
    for (si, di, b) in tab:
        out[..., di, :] = out[..., di, :] + buf[..., si, :] * b
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------- INTER_LINEAR
def linear_coeffs(ssize: int, dsize: int, horizontal: bool):
    """Source index pairs and 11-bit coefficients of cv2's u8 bilinear resize.

    Returns (i0, i1, c0, c1): for dst index d, value = S[i0[d]]*c0[d] + S[i1[d]]*c1[d].
    Horizontal taps are clamped *with* their weights (f forced to 0 at the borders);
    vertical taps only have their row indices clamped (weights kept).
    """
    scale = float(ssize) / float(dsize)
    i0 = np.zeros(dsize, np.int32)
    i1 = np.zeros(dsize, np.int32)
    c0 = np.zeros(dsize, np.int32)
    c1 = np.zeros(dsize, np.int32)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if horizontal:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= ssize - 1:
                s, f = ssize - 1, np.float32(0)
            i0[d] = s
            i1[d] = min(s + 1, ssize - 1)
        else:
            i0[d] = min(max(s, 0), ssize - 1)
            i1[d] = min(max(s + 1, 0), ssize - 1)
        c0[d] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(2048)))
        c1[d] = int(np.rint(f * np.float32(2048)))
    return i0, i1, c0, c1


def resize_linear(small: np.ndarray, dsize: int) -> np.ndarray:
    """cv2.resize(small, (dsize, dsize), interpolation=cv2.INTER_LINEAR) for square u8."""
    ssize = small.shape[-1]
    assert small.shape[-2] == small.shape[-1]
    if ssize == dsize:
        return small.copy()
    s = small.astype(np.int64)
    hi0, hi1, a0, a1 = linear_coeffs(ssize, dsize, horizontal=True)
    vi0, vi1, b0, b1 = linear_coeffs(ssize, dsize, horizontal=False)
    rows = s[..., :, hi0] * a0 + s[..., :, hi1] * a1                     # (..., ssize, dsize) 19-bit
    r0 = rows[..., vi0, :] >> 4
    r1 = rows[..., vi1, :] >> 4
    out = (((b0[:, None] * r0) >> 16) + ((b1[:, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def down_up(block: np.ndarray, small: int) -> np.ndarray:
    """AREA down to small x small then LINEAR up to the block size (elvis.py:2160-2163)."""
    return resize_linear(resize_area(block, small), block.shape[-1])


# ------------------------------------------------- GaussianBlur(tile, (0, 0), sigma) + addWeighted
def gaussian_kernel_q8(sigma: float) -> np.ndarray:
    """cv2's 8.8 fixed-point Gaussian kernel for u8 images when ksize = (0, 0): ksize =
    round(6 sigma + 1) | 1; exact Gaussian normalised to 1, scaled by 256 and rounded with
    error diffusion over the first half (incl. the centre), then mirrored -- sums to 256.
    (sigma = 1 gives [1, 14, 62, 102, 62, 14, 1]; pinned against cv2 in tests/test_oracle.py.)"""
    ksize = int(round(sigma * 6 + 1)) | 1
    r = ksize // 2
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-x * x / (2.0 * sigma * sigma))
    k /= k.sum()
    q = np.zeros(ksize, np.int64)
    err = 0.0
    for i in range(r + 1):
        v = k[i] * 256 + err
        q[i] = int(np.floor(v + 0.5))
        err = v - q[i]
    q[r + 1:] = q[:r][::-1]
    return q


def gaussian_blur_sigma(tile: np.ndarray, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(tile, (0, 0), sigma) on an isolated 2-D u8 tile (REFLECT_101 at its edges,
    reflected as often as needed when the kernel is wider than the tile)."""
    q = gaussian_kernel_q8(sigma)
    r = len(q) // 2
    h, w = tile.shape[-2:]
    cols = np.array([[_reflect101(x + d, w) for d in range(-r, r + 1)] for x in range(w)])
    rows = np.array([[_reflect101(y + d, h) for d in range(-r, r + 1)] for y in range(h)])
    s = tile.astype(np.int64)
    hp = (s[..., :, cols] * q).sum(axis=-1)
    vp = (hp[..., rows, :] * q[:, None]).sum(axis=-2)
    return ((vp + 32768) >> 16).astype(np.uint8)


def unsharp(tile: np.ndarray, level: int) -> np.ndarray:
    """cv2.addWeighted(tile, 1 + a, GaussianBlur(tile, (0,0), max(1, level)), -a, 0), a = level/2
    (elvis.py:2852-2861, utils.py:1296-1300).  2*result = 2x + level*(x - blur) exactly, so the
    float rounding of cv2 is a round-half-to-even of an integer over 2, then saturation."""
    b = gaussian_blur_sigma(tile, max(1, int(level))).astype(np.int64)
    n = 2 * tile.astype(np.int64) + int(level) * (tile.astype(np.int64) - b)
    half = n >> 1                                   # floor(n / 2)
    out = half + ((n & 1) & (half & 1))             # odd n: round to the even neighbour
    return np.clip(out, 0, 255).astype(np.uint8)


# -------------------------------------------------------------------------- INTER_LANCZOS4
def lanczos4_taps(ssize: int, dsize: int):
    """Source indices (dsize, 8) and 11-bit integer coefficients (dsize, 8) of cv2's u8
    INTER_LANCZOS4 resize: float32 Lanczos-4 weights (cv2's sin/cos recurrence), normalised,
    times 2048, rounded to short; source indices clamped to the image."""
    s45 = 0.70710678118654752440084436210485
    cs = np.array([[1, 0], [-s45, -s45], [0, 1], [s45, -s45], [-1, 0], [s45, s45], [0, -1], [-s45, s45]], np.float64)
    scale = float(ssize) / float(dsize)
    idx = np.zeros((dsize, 8), np.int32)
    coef = np.zeros((dsize, 8), np.int32)
    for d in range(dsize):
        fx = (d + 0.5) * scale - 0.5
        sx = int(np.floor(fx))
        x = np.float32(fx - sx)
        c = np.zeros(8, np.float32)
        if x < np.finfo(np.float32).eps:
            c[3] = 1
        else:
            y0 = np.float32(-(x + 3) * np.float32(np.pi) * np.float32(0.25))
            s0, c0 = np.float32(np.sin(y0)), np.float32(np.cos(y0))
            total = np.float32(0)
            for i in range(8):
                y = np.float32(-(x + 3 - i) * np.float32(np.pi) * np.float32(0.25))
                c[i] = np.float32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
                total = np.float32(total + c[i])
            c = (c * (np.float32(1.0) / total)).astype(np.float32)
        coef[d] = np.clip(np.rint(c * np.float32(2048)), -32768, 32767).astype(np.int32)
        idx[d] = np.clip(np.arange(sx - 3, sx + 5), 0, ssize - 1)
    return idx, coef


def resize_lanczos4(small: np.ndarray, dsize: int) -> np.ndarray:
    """cv2.resize(small, (dsize, dsize), interpolation=cv2.INTER_LANCZOS4) for square u8 blocks
    (..., s, s): 8-tap horizontal pass in int32, 8-tap vertical pass, (acc + 2^21) >> 22, saturate."""
    ssize = small.shape[-1]
    idx, coef = lanczos4_taps(ssize, dsize)
    s = small.astype(np.int64)
    rows = (s[..., :, idx] * coef).sum(axis=-1)                                  # (..., ssize, dsize)
    out = (rows[..., idx, :] * coef[:, :, None]).sum(axis=-2)                    # (..., dsize, dsize)
    return np.clip((out + (1 << 21)) >> 22, 0, 255).astype(np.uint8)


# ---------------------------------------------------------------- float32 INTER_AREA, RGB -> I420 (8f rank 3)
def resize_area_f32(src: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """cv2.resize(float32 2-D map, (dw, dh), interpolation=cv2.INTER_AREA) when shrinking, operation
    by operation in float32 (cv2 imgproc resize.cpp: resizeAreaFast_ for integer ratios,
    ResizeArea_Invoker with the decimation table of `area_table` otherwise).  The 2 x 2 case has a
    4-lane vector kernel for the leading multiple of 4 columns that adds in a different order than
    the scalar tail; that width is a property of the cv2 build (128-bit universal intrinsics in the
    cv2 4.13 of this image, which the tests pin against)."""
    src = np.asarray(src, np.float32)
    sh, sw = src.shape
    if dh > sh or dw > sw:
        raise NotImplementedError("enlarging INTER_AREA uses cv2's bilinear kernel")
    sx, sy = sw / dw, sh / dh
    ix, iy = int(round(sx)), int(round(sy))
    eps = np.finfo(np.float64).eps
    out = np.zeros((dh, dw), np.float32)
    f32 = np.float32
    if abs(sx - ix) < eps and abs(sy - iy) < eps:
        scale = f32(1.0 / (ix * iy))
        simd_cols = (dw // 4) * 4 if (ix == 2 and iy == 2) else 0
        for y in range(dh):
            for x in range(dw):
                win = src[y * iy:(y + 1) * iy, x * ix:(x + 1) * ix]
                if x < simd_cols:
                    out[y, x] = f32(f32(f32(win[0, 0] + win[0, 1]) + f32(win[1, 0] + win[1, 1])) * f32(0.25))
                    continue
                vals = win.reshape(-1)
                s, k = f32(0), 0
                while k + 4 <= len(vals):
                    s = f32(s + f32(f32(f32(vals[k] + vals[k + 1]) + vals[k + 2]) + vals[k + 3]))
                    k += 4
                while k < len(vals):
                    s = f32(s + vals[k])
                    k += 1
                out[y, x] = f32(s * scale)
        return out
    xtab = [(d, s, f32(a)) for s, d, a in area_table(sw, dw)]
    ytab = [(d, s, f32(a)) for s, d, a in area_table(sh, dh)]
    first = np.ones(dh, bool)
    for dy, sy_, beta in ytab:
        buf = np.zeros(dw, np.float32)
        for dx, sx_, alpha in xtab:
            buf[dx] = f32(buf[dx] + f32(src[sy_, sx_] * alpha))
        term = (beta * buf).astype(np.float32)
        out[dy] = term if first[dy] else (out[dy] + term).astype(np.float32)
        first[dy] = False
    return out


def rgb_to_i420(frame: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, cv2.COLOR_RGB2YUV_I420) for an (H, W, 3) uint8 frame with even H, W:
    BT.601 limited range in 20-bit fixed point, chroma taken from the top-left pixel of each 2 x 2
    quad.  Returns the (H * 3 / 2, W) uint8 I420 image cv2 returns."""
    h, w = frame.shape[:2]
    if h % 2 or w % 2:
        raise ValueError("4:2:0 needs even frame dimensions")
    r, g, b = (frame[..., c].astype(np.int64) for c in range(3))
    half, sh = 1 << 19, 20
    y = (269484 * r + 528482 * g + 102760 * b + half + (16 << sh)) >> sh
    rs, gs, bs = r[::2, ::2], g[::2, ::2], b[::2, ::2]
    u = (-155188 * rs - 305135 * gs + 460324 * bs + half + (128 << sh)) >> sh
    v = (460324 * rs - 385875 * gs - 74448 * bs + half + (128 << sh)) >> sh
    return np.concatenate([y.reshape(-1), u.reshape(-1), v.reshape(-1)]).astype(np.uint8).reshape(h * 3 // 2, w)


# ----------------------------------------------------------------------------- float INTER_LINEAR, RGB2GRAY
def _fma(a: np.ndarray, b: float, c: np.ndarray, dtype) -> np.ndarray:
    """Elementwise fused a * b + c with ONE rounding to `dtype` (exact rational arithmetic)."""
    from fractions import Fraction
    out = np.empty(a.shape, dtype)
    fo, fa, fc, fb = out.reshape(-1), a.reshape(-1), c.reshape(-1), Fraction(float(b))
    for i in range(fa.size):
        v = Fraction(float(fa[i])) * fb + Fraction(float(fc[i]))
        if dtype == np.float64:
            fo[i] = float(v)
        else:                                  # round the exact value to float32 once (no double rounding)
            d = float(v)
            f = np.float32(d)
            if Fraction(float(f)) != v:        # d -> f may have rounded a half-way case the wrong way
                lo, hi = (np.nextafter(f, np.float32(-np.inf)), f) if Fraction(float(f)) > v else (f, np.nextafter(f, np.float32(np.inf)))
                dl, dh = v - Fraction(float(lo)), Fraction(float(hi)) - v
                if dl != dh:
                    f = lo if dl < dh else hi
                else:                          # tie: even mantissa
                    f = lo if (lo.view(np.uint32) & 1) == 0 else hi
            fo[i] = f
    return out


def linear_float_coords(ssize: int, dsize: int, fused: bool):
    """[(source index, fraction)] per destination index of cv2.resize(float map, INTER_LINEAR).
    Coordinate (d + 0.5) * (ssize / dsize) - 0.5 in double -- evaluated as ONE fused multiply-add on
    cv2's float64 path (`fused`), as a rounded product and a rounded difference on its float32 path
    (both pinned against cv2 4.13: the two differ where the product lands within an ulp of x.5) --
    then index = floor, fraction = coordinate - index, clamped with a zero fraction at the borders."""
    from fractions import Fraction
    scale = ssize / dsize
    out = []
    for d in range(dsize):
        fx = float(Fraction(2 * d + 1, 2) * Fraction(scale) - Fraction(1, 2)) if fused else (d + 0.5) * scale - 0.5
        s = int(np.floor(fx))
        f = fx - s
        if s < 0:
            s, f = 0, 0.0
        if s >= ssize - 1:
            s, f = ssize - 1, 0.0
        out.append((s, f))
    return out


def resize_linear_float(src: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=INTER_LINEAR) for a 2-D float32 / float64 map with at
    least two rows and two columns (utils.py:1127-1128, 1197-1198; elvis.py:2068-2073): fused lerp
    x0 + (x1 - x0) * f with the difference rounded first, horizontal pass then vertical pass; float32
    maps use float32 fractions.  (Single-row / single-column sources take another cv2 path: not restated.)"""
    dtype = src.dtype.type
    sh, sw = src.shape
    if sh < 2 or sw < 2:
        raise NotImplementedError("single-row / single-column source maps")
    fused = dtype == np.float64
    rows = np.empty((sh, dw), dtype)
    for x, (s, f) in enumerate(linear_float_coords(sw, dw, fused)):
        x0, x1 = src[:, s], src[:, min(s + 1, sw - 1)]
        rows[:, x] = _fma((x1 - x0).astype(dtype), float(dtype(f)), x0, dtype)
    out = np.empty((dh, dw), dtype)
    for y, (s, f) in enumerate(linear_float_coords(sh, dh, fused)):
        r0, r1 = rows[s], rows[min(s + 1, sh - 1)]
        out[y] = _fma((r1 - r0).astype(dtype), float(dtype(f)), r0, dtype)
    return out


def rgb_to_gray(frame: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, COLOR_RGB2GRAY) for uint8 (..., 3): 15-bit fixed point."""
    x = frame.astype(np.int64)
    return ((x[..., 0] * 9798 + x[..., 1] * 19235 + x[..., 2] * 3735 + (1 << 14)) >> 15).astype(np.uint8)
