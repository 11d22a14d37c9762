"""Device-level operators: CUDA uint8/float tensors in, CUDA tensors out, one C-ABI call
each, enqueued on torch's current stream.  torch is used for memory and streams only.

Clip layouts
  planar plane : (T, H, W) uint8           -- Y, U or V of a YUV clip
  packed clip  : (T, H, W, C) uint8        -- the reference's BGR/RGB frames (C = 3)
  block maps   : (T, By, Bx)               -- scores float64, masks uint8, levels int32
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import _lib, _tables
from ._lib import (F32, F64, LEVELS_INVERTED_BINS, LEVELS_INVERTED_ROUND, LEVELS_ROUND, REMOVE_HIGH,
                   REMOVE_LOW, Plane, call)

__all__ = ["score_sc_tc", "minmax", "combine_removability", "normalize_", "importance_scores", "select_rows",
           "shrink", "stretch", "move_yuv420", "degrade_yuv420", "levels_from_scores", "degrade_blur", "degrade_downsample", "dct_dampen",
           "restore_unsharp", "restore_lanczos", "temporal_blend_", "pack_mask_bits", "unpack_mask_bits", "pack_levels_2bit", "unpack_levels_2bit",
           "REMOVE_HIGH", "REMOVE_LOW", "LEVELS_ROUND", "LEVELS_INVERTED_ROUND", "LEVELS_INVERTED_BINS"]


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(fn):
    """Run an operator with the device of its first tensor argument current: the library launches
    on the CURRENT device and on the stream handed to it, so both must belong to the tensors' GPU
    (one process may drive several GPUs)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = None
        for a in args:                      # the first tensor argument (possibly inside a tuple of planes)
            if isinstance(a, (tuple, list)) and a:
                a = a[0]
            if isinstance(a, torch.Tensor):
                t = a
                break
        if t is not None and t.is_cuda and t.device.index != torch.cuda.current_device():
            with torch.cuda.device(t.device):
                return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _check_cuda(t: torch.Tensor, dtype, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (elvis_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def plane_of(clip: torch.Tensor, name: str = "clip") -> Plane:
    """struct elvis_plane of a (T, H, W) or (T, H, W, C) uint8 CUDA tensor; rows may be
    strided (a cropped view), pixels of a row must be dense."""
    _check_cuda(clip, torch.uint8, name)
    if clip.dim() == 3:
        ch, (sf, sr, sx) = 1, clip.stride()
        ok = sx == 1 or clip.shape[2] == 1
    elif clip.dim() == 4:
        ch = clip.shape[3]
        sf, sr, sx, sc = clip.stride()
        ok = (sc == 1 or ch == 1) and (sx == ch or clip.shape[2] == 1)
    else:
        raise ValueError(f"{name} must be (T, H, W) or (T, H, W, C)")
    if not ok:
        raise ValueError(f"{name}: pixels of a row must be contiguous")
    return Plane(clip.data_ptr(), sf, sr, clip.shape[1], clip.shape[2], ch, 0)


def _float_dtype(t: torch.Tensor, name: str) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise TypeError(f"{name} must be float32 or float64")


# ------------------------------------------------------------------------------ a1
def score_sc_tc(y: torch.Tensor, block_size: int, prev_halo: torch.Tensor | None = None, dct_size: int = 8,
                minmax_range: tuple | None = None, out: tuple | None = None):
    """SC/TC of a (T, H, W) luma clip -> (sc, tc, minmax): float32 (T, By, Bx) each and a
    float32[4] {sc_min, sc_max, tc_min, tc_max} over frames minmax_range (default: all)."""
    pl = plane_of(y, "y")
    if y.dim() != 3:
        raise ValueError("y must be (T, H, W)")
    T, H, W = y.shape
    by, bx = H // block_size, W // block_size
    if prev_halo is not None:
        _check_cuda(prev_halo, torch.uint8, "prev_halo")
        if prev_halo.shape != y.shape[1:] or prev_halo.stride() != y.stride()[1:]:
            prev_halo = _match_halo(prev_halo, y)
    if out is None:
        sc = torch.empty((T, by, bx), dtype=torch.float32, device=y.device)
        tc = torch.empty_like(sc)
        mm = torch.empty(4, dtype=torch.float32, device=y.device)
    else:   # caller-owned (T', By, Bx) float32 buffers with T' >= T, and a float32[4]
        sc, tc, mm = out[0][:T], out[1][:T], out[2]
        if sc.shape != (T, by, bx) or tc.shape != sc.shape or not (sc.is_contiguous() and tc.is_contiguous()):
            raise ValueError("out buffers must be contiguous (T, By, Bx) float32")
    lo, hi = (0, T) if minmax_range is None else minmax_range
    call("elvis_score_sc_tc", C.byref(pl), T, _ptr(prev_halo), block_size, dct_size, _ptr(sc), _ptr(tc), _ptr(mm),
         lo, hi, _stream())
    return sc, tc, mm


def _match_halo(halo: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    if halo.shape != y.shape[1:]:
        raise ValueError("prev_halo must have the shape of one luma frame")
    buf = torch.empty_strided(y.shape[1:], y.stride()[1:], dtype=torch.uint8, device=y.device)
    buf.copy_(halo)
    return buf


def minmax(x: torch.Tensor) -> torch.Tensor:
    if not x.is_contiguous():
        raise ValueError("minmax needs a contiguous tensor")
    out = torch.empty(2, dtype=torch.float64, device=x.device)
    call("elvis_minmax", _ptr(x), _float_dtype(x, "x"), x.numel(), _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------ a2
def combine_removability(sc: torch.Tensor, tc: torch.Tensor, norm: torch.Tensor, alpha: float, beta: float,
                         background: torch.Tensor | None = None, t_begin: int = 0, t_count: int | None = None,
                         is_first: bool = True, is_last: bool = True, clip_frames: int | None = None,
                         out: tuple | None = None):
    """Un-normalised smoothed removability of frames [t_begin, t_begin + t_count) plus its
    {min, max} (float64[2]).  clip_frames = length of the whole clip (for the reference's
    `shape[0] >= 2` smoothing condition); defaults to the local length."""
    dt = _float_dtype(sc, "sc")
    if tc.dtype != sc.dtype or norm.dtype != sc.dtype:
        raise TypeError("sc, tc and norm must share a dtype")
    if not (sc.is_contiguous() and tc.is_contiguous()) or sc.shape != tc.shape or sc.dim() != 3:
        raise ValueError("sc and tc must be contiguous (T, By, Bx)")
    t_ext, by, bx = sc.shape
    t_count = t_ext - t_begin if t_count is None else t_count
    clip_frames = t_ext if clip_frames is None else clip_frames
    if background is not None:
        _check_cuda(background, torch.uint8, "background")
        if background.shape != sc.shape or not background.is_contiguous():
            raise ValueError("background must be contiguous (T, By, Bx) uint8")
    if out is None:
        out = torch.empty((t_count, by, bx), dtype=torch.float64, device=sc.device)
        mm = torch.empty(2, dtype=torch.float64, device=sc.device)
    else:
        out, mm = out
        if out.shape != (t_count, by, bx) or not out.is_contiguous() or out.dtype != torch.float64:
            raise ValueError("out must be contiguous (t_count, By, Bx) float64")
    smooth = int(beta < 1 and clip_frames >= 2)
    call("elvis_combine_removability", _ptr(sc), _ptr(tc), dt, _ptr(norm), t_ext, by, bx, t_begin, t_count,
         int(is_first), int(is_last), _ptr(background), float(alpha), float(beta), smooth, _ptr(out), _ptr(mm), _stream())
    return out, mm


def normalize_(x: torch.Tensor, mm: torch.Tensor) -> torch.Tensor:
    _check_cuda(x, torch.float64, "x")
    _check_cuda(mm, torch.float64, "minmax")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous")
    call("elvis_normalize", _ptr(x), x.numel(), _ptr(mm), _stream())
    return x


# ------------------------------------------------------------------------------ a3
def importance_scores(sc: torch.Tensor, tc: torch.Tensor, foreground: torch.Tensor | None, alpha: float, beta: float,
                      t_begin: int = 0, t_count: int | None = None, is_first: bool = True, is_last: bool = True):
    dt = _float_dtype(sc, "sc")
    if tc.dtype != sc.dtype or (foreground is not None and foreground.dtype != sc.dtype):
        raise TypeError("sc, tc and foreground must share a dtype")
    if not (sc.is_contiguous() and tc.is_contiguous()) or sc.shape != tc.shape or sc.dim() != 3:
        raise ValueError("sc and tc must be contiguous (T, By, Bx)")
    if foreground is not None and (foreground.shape != sc.shape or not foreground.is_contiguous()):
        raise ValueError("foreground must be contiguous (T, By, Bx)")
    t_ext, by, bx = sc.shape
    t_count = t_ext - t_begin if t_count is None else t_count
    out = torch.empty((t_count, by, bx), dtype=torch.float64, device=sc.device)
    call("elvis_importance_scores", _ptr(sc), _ptr(tc), _ptr(foreground), dt, t_ext, by, bx, t_begin, t_count,
         int(is_first), int(is_last), float(alpha), float(beta), _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------ a4-a7
def select_rows(scores: torch.Tensor, k, polarity: int = REMOVE_HIGH, out: torch.Tensor | None = None,
                normalize_with: torch.Tensor | None = None) -> torch.Tensor:
    """(T, By, Bx) float64 -> uint8 mask, 1 = removed.  k: int, or int32 CUDA tensor (By,).
    normalize_with: {min, max} float64 on the device -- normalize_(scores, mm) folded into the same pass
    (the scores are normalised IN PLACE and ranked on the normalised values)."""
    _check_cuda(scores, torch.float64, "scores")
    if scores.dim() != 3 or not scores.is_contiguous():
        raise ValueError("scores must be contiguous (T, By, Bx)")
    T, by, bx = scores.shape
    mask = out if out is not None else torch.empty((T, by, bx), dtype=torch.uint8, device=scores.device)
    if mask.shape != (T, by, bx) or mask.dtype != torch.uint8 or not mask.is_contiguous():
        raise ValueError("out must be contiguous (T, By, Bx) uint8")
    if isinstance(k, torch.Tensor):
        _check_cuda(k, torch.int32, "k")
        if k.shape != (by,) or not k.is_contiguous():
            raise ValueError("per-row k must be a contiguous (By,) tensor")
        k_rows, k_all = _ptr(k), 0
    else:
        k_rows, k_all = _ptr(None), int(k)
    if normalize_with is not None:
        _check_cuda(normalize_with, torch.float64, "minmax")
        call("elvis_normalize_select_rows", _ptr(scores), _ptr(normalize_with), T, by, bx, k_rows, k_all, polarity, _ptr(mask), _stream())
    else:
        call("elvis_select_rows", _ptr(scores), T, by, bx, k_rows, k_all, polarity, _ptr(mask), _stream())
    return mask


def _mask_arg(mask: torch.Tensor) -> torch.Tensor:
    _check_cuda(mask, torch.uint8, "mask")
    if mask.dim() != 3 or not mask.is_contiguous():
        raise ValueError("mask must be contiguous (T, By, Bx) uint8")
    return mask


def shrink(clip: torch.Tensor, mask: torch.Tensor, block_px: int, out_bx: int, out: torch.Tensor | None = None,
           ctas_per_sm: int = 0) -> torch.Tensor:
    """Left-compact the blocks with mask == 0.  clip: (T, H, W[, C]); returns
    (T, By*block_px, out_bx*block_px[, C])."""
    mask = _mask_arg(mask)
    T, by, bx = mask.shape
    if clip.shape[0] != T:
        raise ValueError("clip and mask disagree on the frame count")
    if clip.shape[1] % block_px or clip.shape[2] % block_px:
        raise ValueError("Image dimensions must be divisible by block_size.")   # elvis.py:1376
    shape = (T, by * block_px, out_bx * block_px) + tuple(clip.shape[3:])
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=clip.device)
    elif tuple(out.shape) != shape:
        raise ValueError(f"out must be {shape}")
    if out_bx == 0 or out.numel() == 0:
        return out
    src, dst = plane_of(clip), plane_of(out, "out")
    call("elvis_shrink", C.byref(src), C.byref(dst), T, block_px, by, bx, out_bx, _ptr(mask), int(ctas_per_sm), _stream())
    return out


def stretch(shrunk: torch.Tensor, mask: torch.Tensor, block_px: int, out: torch.Tensor | None = None,
            ctas_per_sm: int = 0) -> torch.Tensor:
    """Inverse of shrink: (T, By*block_px, sbx*block_px[, C]) -> (T, By*block_px, Bx*block_px[, C])."""
    mask = _mask_arg(mask)
    T, by, bx = mask.shape
    if shrunk.shape[0] != T:
        raise ValueError("clip and mask disagree on the frame count")
    if shrunk.shape[1] % block_px or shrunk.shape[2] % block_px:
        raise ValueError("Image dimensions must be divisible by block_size.")
    sbx = shrunk.shape[2] // block_px
    shape = (T, by * block_px, bx * block_px) + tuple(shrunk.shape[3:])
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=shrunk.device)
    elif tuple(out.shape) != shape:
        raise ValueError(f"out must be {shape}")
    if sbx == 0:   # every block was removed: the canvas stays black
        return out.zero_()
    src, dst = plane_of(shrunk), plane_of(out, "out")
    call("elvis_stretch", C.byref(src), C.byref(dst), T, block_px, by, bx, sbx, _ptr(mask), int(ctas_per_sm), _stream())
    return out


def _yuv_planes(planes, name: str):
    arr = (Plane * 3)(*[plane_of(p, name) for p in planes])
    return arr


def move_yuv420(src, dst, mask: torch.Tensor, block_size: int, small_bx: int, stretch_: bool, ctas_per_sm: int = 0) -> bool:
    """Y, U and V of a planar 4:2:0 clip in one launch (src, dst: 3-tuples of (T, H, W) planes).
    Returns False when the geometry is not supported by the fused kernel (caller falls back to
    the per-plane operators)."""
    mask = _mask_arg(mask)
    T, by, bx = mask.shape
    if block_size % 16:
        return False
    s, d = _yuv_planes(src, "src"), _yuv_planes(dst, "dst")
    name = "elvis_stretch_yuv420" if stretch_ else "elvis_shrink_yuv420"
    rc = getattr(_lib.lib, name)(s, d, T, block_size, by, bx, small_bx, _ptr(mask), int(ctas_per_sm), _stream())
    if rc == _lib.ERR_UNSUPPORTED:
        return False
    if rc != 0:
        call(name, s, d, T, block_size, by, bx, small_bx, _ptr(mask), int(ctas_per_sm), _stream())   # raises
    return True


# ------------------------------------------------------------------------------ a8-a12, a14
def levels_from_scores(scores: torch.Tensor, rule: int, param: int) -> torch.Tensor:
    _check_cuda(scores, torch.float64, "scores")
    if not scores.is_contiguous():
        raise ValueError("scores must be contiguous")
    out = torch.empty(scores.shape, dtype=torch.int32, device=scores.device)
    call("elvis_levels_from_scores", _ptr(scores), scores.numel(), rule, int(param), _ptr(out), _stream())
    return out


def _degrade_args(clip: torch.Tensor, block_map: torch.Tensor, block_px: int, dtype, out):
    _check_cuda(block_map, dtype, "block map")
    if block_map.dim() != 3 or not block_map.is_contiguous() or block_map.shape[0] != clip.shape[0]:
        raise ValueError("block map must be contiguous (T, By, Bx)")
    T, by, bx = block_map.shape
    if by != clip.shape[1] // block_px or bx != clip.shape[2] // block_px:
        raise ValueError("block map does not match the clip's block grid")
    if out is None:
        out = torch.empty_like(clip, memory_format=torch.contiguous_format)
    return T, by, bx, out


def split_channels3(clip: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """(T, H, W, 3) uint8 packed frames -> (3, T, H, W) planes."""
    _check_cuda(clip, torch.uint8, "clip")
    if clip.dim() != 4 or clip.shape[3] != 3:
        raise ValueError("clip must be (T, H, W, 3)")
    T, H, W, _ = clip.shape
    if out is None:
        out = torch.empty((3, T, H, W), dtype=torch.uint8, device=clip.device)
    if clip.numel():
        planes = (Plane * 3)(*[plane_of(out[c], "planes") for c in range(3)])
        src = plane_of(clip)
        call("elvis_split_channels3", C.byref(src), planes, T, _stream())
    return out


def merge_channels3(planes: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """(3, T, H, W) uint8 planes -> (T, H, W, 3) packed frames."""
    _check_cuda(planes, torch.uint8, "planes")
    if planes.dim() != 4 or planes.shape[0] != 3:
        raise ValueError("planes must be (3, T, H, W)")
    _, T, H, W = planes.shape
    if out is None:
        out = torch.empty((T, H, W, 3), dtype=torch.uint8, device=planes.device)
    if planes.numel():
        pl = (Plane * 3)(*[plane_of(planes[c], "planes") for c in range(3)])
        dst = plane_of(out, "out")
        call("elvis_merge_channels3", pl, C.byref(dst), T, _stream())
    return out


def _per_channel(clip: torch.Tensor, out: torch.Tensor, block_px: int, by: int, bx: int, plane_op) -> bool:
    """Packed 3-channel clips with 8- / 16-pixel blocks: split into planes, run the planar (fast) kernel per channel,
    merge -- the channels are independent in every per-block degradation (planar Y/U/V == packed per channel).
    Partial blocks at the right / bottom edge ride along (the planar kernels copy them through)."""
    if clip.dim() != 4 or clip.shape[3] != 3 or block_px not in (8, 16) or not clip.numel():
        return False
    planes = split_channels3(clip)
    result = torch.empty_like(planes)
    for c in range(3):
        plane_op(planes[c], result[c])
    merge_channels3(result, out)
    return True


def degrade_blur(clip: torch.Tensor, rounds: torch.Tensor, block_px: int, out: torch.Tensor | None = None) -> torch.Tensor:
    T, by, bx, out = _degrade_args(clip, rounds, block_px, torch.int32, out)
    if _per_channel(clip, out, block_px, by, bx, lambda p, o: degrade_blur(p, rounds, block_px, out=o)):
        return out
    src, dst = plane_of(clip), plane_of(out, "out")
    call("elvis_degrade_blur", C.byref(src), C.byref(dst), T, block_px, by, bx, _ptr(rounds), _stream())
    return out


_table_cache: dict = {}


def _device_tables(block_px: int, small_sizes: tuple, device, lanczos: bool = False) -> torch.Tensor:
    key = (block_px, small_sizes, str(device), lanczos)
    t = _table_cache.get(key)
    if t is None:
        t = torch.from_numpy(_tables.build(block_px, small_sizes, lanczos).copy()).to(device)
        _table_cache[key] = t
    return t


def degrade_downsample(clip: torch.Tensor, levels: torch.Tensor, block_px: int, small_sizes,
                       out: torch.Tensor | None = None) -> torch.Tensor:
    """small_sizes[level] = side the block is reduced to before being scaled back."""
    T, by, bx, out = _degrade_args(clip, levels, block_px, torch.int32, out)
    small_sizes = tuple(int(s) for s in small_sizes)
    if _per_channel(clip, out, block_px, by, bx, lambda p, o: degrade_downsample(p, levels, block_px, small_sizes, out=o)):
        return out
    tab = _device_tables(block_px, small_sizes, clip.device)
    fast_ok = _tables.fast_levels_flag(block_px, small_sizes)
    src, dst = plane_of(clip), plane_of(out, "out")
    call("elvis_degrade_downsample", C.byref(src), C.byref(dst), T, block_px, by, bx, _ptr(levels), _ptr(tab),
         len(small_sizes), fast_ok, _stream())
    return out


def degrade_yuv420(kind: str, src, dst, block_map: torch.Tensor, block_size: int, param: int = 0) -> bool:
    """Y, U and V of a planar 4:2:0 clip in one launch (src, dst: 3-tuples of (T, H, W) planes).  kind:
    "downsample_pow2" (block_map int32 levels, param = max_level).  Returns False when the geometry is not
    supported by the fused kernel (the caller then uses the per-plane operators)."""
    T, by, bx = block_map.shape
    if block_size != 16 or not block_map.is_contiguous():
        return False
    s, d = _yuv_planes(src, "src"), _yuv_planes(dst, "dst")
    name = {"downsample_pow2": "elvis_degrade_downsample_pow2_yuv420"}[kind]
    _check_cuda(block_map, torch.int32, "block map")
    args = (s, d, T, block_size, by, bx, _ptr(block_map), int(param), _stream())
    rc = getattr(_lib.lib, name)(*args)
    if rc == _lib.ERR_UNSUPPORTED:
        return False
    if rc != 0:
        call(name, *args)   # raises
    return True


def dct_dampen(clip: torch.Tensor, strength: torch.Tensor, block_px: int, out: torch.Tensor | None = None) -> torch.Tensor:
    T, by, bx, out = _degrade_args(clip, strength, block_px, torch.float32, out)
    src, dst = plane_of(clip), plane_of(out, "out")
    call("elvis_dct_dampen", C.byref(src), C.byref(dst), T, block_px, by, bx, _ptr(strength), _stream())
    return out


# ------------------------------------------------------------------------------ 8f rank 1
_kernel_cache: dict = {}


def restore_unsharp(clip: torch.Tensor, levels: torch.Tensor, block_px: int, halo: int = 0, max_level: int | None = None,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    """Per-block unsharp mask (radius = level, amount = level / 2) -- the OpenCV client restorer.
    max_level: largest level in the map (computed with a device sync when None)."""
    T, by, bx, out = _degrade_args(clip, levels, block_px, torch.int32, out)
    if max_level is None:
        max_level = max(1, int(levels.max().item()))
    key = (max_level, str(clip.device))
    tab = _kernel_cache.get(key)
    if tab is None:
        tab = torch.from_numpy(_tables.gaussian_kernels(max_level)).to(clip.device)
        _kernel_cache[key] = tab
    src, dst = plane_of(clip), plane_of(out, "out")
    call("elvis_restore_unsharp", C.byref(src), C.byref(dst), T, block_px, by, bx, _ptr(levels), int(halo), _ptr(tab),
         max_level, tab.shape[1], _stream())
    return out


def restore_lanczos(clip: torch.Tensor, levels: torch.Tensor, block_px: int, small_sizes, out: torch.Tensor | None = None) -> torch.Tensor:
    """Per block: INTER_AREA down to small_sizes[level], INTER_LANCZOS4 back up (elvis.py:2773-2820)."""
    T, by, bx, out = _degrade_args(clip, levels, block_px, torch.int32, out)
    small_sizes = tuple(int(s) for s in small_sizes)
    tab = _device_tables(block_px, small_sizes, clip.device, lanczos=True)
    src, dst = plane_of(clip), plane_of(out, "out")
    call("elvis_restore_lanczos", C.byref(src), C.byref(dst), T, block_px, by, bx, _ptr(levels), _ptr(tab),
         len(small_sizes), _stream())
    return out


def temporal_blend_(clip: torch.Tensor, temporal_blend: float) -> torch.Tensor:
    """In place: frame[t] = uint8(tb * frame[t-1] + (1 - tb) * frame[t]) for t >= 1."""
    pl = plane_of(clip)
    call("elvis_temporal_blend", C.byref(pl), clip.shape[0], float(temporal_blend), _stream())
    return clip


# ------------------------------------------------------------------------------ a13
def pack_mask_bits(mask: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """np.packbits(mask) of a uint8 {0, 1} mask (MSB first, flat over all axes; elvis.py:4412-4418)."""
    _check_cuda(mask, torch.uint8, "mask")
    m = mask.contiguous()
    if out is None:
        out = torch.empty((m.numel() + 7) // 8, dtype=torch.uint8, device=m.device)
    elif out.numel() != (m.numel() + 7) // 8 or out.dtype != torch.uint8 or not out.is_contiguous():
        raise ValueError("out must hold ceil(n / 8) contiguous bytes")
    call("elvis_pack_mask_bits", _ptr(m), m.numel(), _ptr(out), _stream())
    return out


def unpack_mask_bits(packed: torch.Tensor, shape) -> torch.Tensor:
    _check_cuda(packed, torch.uint8, "packed")
    n = int(np.prod(shape))
    if packed.numel() * 8 < n:
        raise ValueError("packed buffer too small")
    out = torch.empty(tuple(shape), dtype=torch.uint8, device=packed.device)
    call("elvis_unpack_mask_bits", _ptr(packed.contiguous()), n, _ptr(out), _stream())
    return out


def pack_levels_2bit(levels: torch.Tensor) -> torch.Tensor:
    _check_cuda(levels, torch.int32, "levels")
    lv = levels.contiguous()
    bx = lv.shape[-1]
    rows = lv.numel() // bx
    out = torch.empty(tuple(lv.shape[:-1]) + ((bx + 3) // 4,), dtype=torch.uint8, device=lv.device)
    call("elvis_pack_levels_2bit", _ptr(lv), rows, bx, _ptr(out), _stream())
    return out


def unpack_levels_2bit(packed: torch.Tensor, bx: int) -> torch.Tensor:
    _check_cuda(packed, torch.uint8, "packed")
    p = packed.contiguous()
    if p.shape[-1] != (bx + 3) // 4:
        raise ValueError("packed rows must hold ceil(bx/4) bytes")
    rows = p.numel() // p.shape[-1]
    out = torch.empty(tuple(p.shape[:-1]) + (bx,), dtype=torch.int32, device=p.device)
    call("elvis_unpack_levels_2bit", _ptr(p), rows, bx, _ptr(out), _stream())
    return out


# ---------------------------------------------------------------- row+column shrink (8f rank 2)
def rowcol_dims(by: int, bx: int, target: int):
    """Pass structure of utils.py:790-836 / 893-948, which depends on the grid size and the
    removal target only -> (final_by, final_bx, blocks removed per pass)."""
    counts = []
    removed = 0
    while removed < target and by > 0 and bx > 0:
        n = min(by, target - removed)
        counts.append(n)
        removed += n
        if n == by:
            bx -= 1
        if removed >= target or bx <= 0:
            break
        n = min(bx, target - removed)
        counts.append(n)
        removed += n
        if n == bx:
            by -= 1
    return by, bx, counts


def rowcol_plan(importance: torch.Tensor, target: int):
    """importance (T, By, Bx) float64 -> (mask uint8 (T, By, Bx), position int32 (T, By, Bx),
    pass_indices int32 (T, P, max(By, Bx)), pass_counts int32 (T, P), meta int32 (T, 4))."""
    _check_cuda(importance, torch.float64, "importance")
    if importance.dim() != 3:
        raise ValueError("importance must be (T, By, Bx)")
    importance = importance.contiguous()
    T, by, bx = importance.shape
    dev = importance.device
    max_passes = by + bx + 2
    keys = torch.empty((T, by, bx), dtype=torch.int64, device=dev)
    pos = torch.empty((T, by, bx), dtype=torch.int32, device=dev)
    mask = torch.empty((T, by, bx), dtype=torch.uint8, device=dev)
    pidx = torch.zeros((T, max_passes, max(by, bx)), dtype=torch.int32, device=dev)
    pcnt = torch.zeros((T, max_passes), dtype=torch.int32, device=dev)
    meta = torch.zeros((T, 4), dtype=torch.int32, device=dev)
    call("elvis_rowcol_plan", _ptr(importance), T, by, bx, int(target), _ptr(keys), _ptr(pos), _ptr(mask), _ptr(pidx),
         _ptr(pcnt), max_passes, _ptr(meta), _stream())
    return mask, pos, pidx, pcnt, meta


def rowcol_expand(pass_indices: torch.Tensor, pass_counts: torch.Tensor, shrunk_by: int, shrunk_bx: int) -> torch.Tensor:
    """pass_indices (T, P, L) / pass_counts (T, P) int32 -> grid int32 (T, shrunk_by + #column
    passes, shrunk_bx + #row passes) of shrunk linear block indices, -1 = black block."""
    _check_cuda(pass_indices, torch.int32, "pass_indices")
    _check_cuda(pass_counts, torch.int32, "pass_counts")
    T, P, L = pass_indices.shape
    gh, gw = shrunk_by + P // 2, shrunk_bx + (P + 1) // 2
    grid = torch.empty((T, gh, gw), dtype=torch.int32, device=pass_indices.device)
    if grid.numel() == 0:
        return grid
    call("elvis_rowcol_expand", _ptr(pass_indices.contiguous()), _ptr(pass_counts.contiguous()), T, P, max(L, 1), shrunk_by,
         shrunk_bx, _ptr(grid), gh, gw, _stream())
    return grid


def invert_block_map(block_map: torch.Tensor, inverse_per_frame: int) -> torch.Tensor:
    """block_map (T, n) int32 of target indices -> (T, inverse_per_frame) int32: the last entry
    pointing at each target, -1 where none does."""
    _check_cuda(block_map, torch.int32, "block_map")
    block_map = block_map.contiguous()
    T, n = block_map.shape
    inv = torch.empty((T, inverse_per_frame), dtype=torch.int32, device=block_map.device)
    call("elvis_invert_block_map", _ptr(block_map), T, n, _ptr(inv), inverse_per_frame, _stream())
    return inv


def gather_blocks(clip: torch.Tensor, block_map: torch.Tensor, block_px: int, dst_by: int, dst_bx: int,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """out block (j, i) of frame t = clip block block_map[t, j, i] (linear over the clip's block
    grid) or zeros for negative entries.  block_map: (T, rows >= dst_by, pitch >= dst_bx) int32."""
    _check_cuda(block_map, torch.int32, "block_map")
    if block_map.dim() != 3 or block_map.stride(2) != 1 or block_map.stride(0) != block_map.shape[1] * block_map.stride(1):
        block_map = block_map.contiguous()
    T = block_map.shape[0]
    if clip.shape[0] != T:
        raise ValueError("clip and block_map disagree on the frame count")
    src_by, src_bx = clip.shape[1] // block_px, clip.shape[2] // block_px
    shape = (T, dst_by * block_px, dst_bx * block_px) + tuple(clip.shape[3:])
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=clip.device)
    elif tuple(out.shape) != shape:
        raise ValueError(f"out must be {shape}")
    if out.numel() == 0:
        return out
    if src_by == 0 or src_bx == 0:
        return out.zero_()
    src, dst = plane_of(clip), plane_of(out, "out")
    call("elvis_gather_blocks", C.byref(src), C.byref(dst), T, block_px, dst_by, dst_bx, src_by, src_bx, _ptr(block_map),
         block_map.shape[1], block_map.stride(1), _stream())
    return out


# ---------------------------------------------------------------- ROI side files, raw 4:2:0 (8f rank 3)
def roi_kvazaar(importance: torch.Tensor, base_qp: int, qp_range: int) -> torch.Tensor:
    """float64 importance (any shape) -> int8 delta QP of utils.py:1046-1052."""
    _check_cuda(importance, torch.float64, "importance")
    importance = importance.contiguous()
    out = torch.empty(importance.shape, dtype=torch.int8, device=importance.device)
    if out.numel():
        call("elvis_roi_kvazaar", _ptr(importance), importance.numel(), int(base_qp), int(qp_range), _ptr(out), _stream())
    return out


def roi_prepare_f32(x: torch.Tensor, mode: int) -> torch.Tensor:
    """mode 0: float32(x); mode 1: float32(clip(2 x - 1, -1, 1)) (float64 arithmetic)."""
    _check_cuda(x, torch.float64, "x")
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    if out.numel():
        call("elvis_roi_prepare_f32", _ptr(x), x.numel(), int(mode), _ptr(out), _stream())
    return out


def resize_area_f32(maps: torch.Tensor, dst_h: int, dst_w: int) -> torch.Tensor:
    """cv2.resize(map, (dst_w, dst_h), INTER_AREA) for every float32 map of (T, h, w); shrinking only."""
    _check_cuda(maps, torch.float32, "maps")
    if maps.dim() != 3:
        raise ValueError("maps must be (T, h, w)")
    maps = maps.contiguous()
    T, sh, sw = maps.shape
    if dst_h > sh or dst_w > sw or dst_h <= 0 or dst_w <= 0:
        raise NotImplementedError("INTER_AREA enlargement (cv2 switches to a bilinear kernel) is not supported")
    out = torch.empty((T, dst_h, dst_w), dtype=torch.float32, device=maps.device)
    if out.numel() == 0:
        return out
    ix, iy, simd = _tables.area_f32_plan(sh, sw, dst_h, dst_w)
    if ix:
        call("elvis_resize_area_f32", _ptr(maps), T, sh, sw, _ptr(out), dst_h, dst_w, None, None, None, None, None, None, ix, iy, simd,
             _stream())
        return out
    tabs = [torch.from_numpy(a).to(maps.device) for a in (*_tables.area_f32_tables(sw, dst_w), *_tables.area_f32_tables(sh, dst_h))]
    call("elvis_resize_area_f32", _ptr(maps), T, sh, sw, _ptr(out), dst_h, dst_w, *[_ptr(t) for t in tabs], 0, 0, 0, _stream())
    return out


def roi_svtav1_offsets(resized: torch.Tensor, base_crf: int, qp_range: int) -> torch.Tensor:
    """float32 importance on the 64-pixel grid -> int32 QP offsets of utils.py:1081-1088."""
    _check_cuda(resized, torch.float32, "resized")
    resized = resized.contiguous()
    out = torch.empty(resized.shape, dtype=torch.int32, device=resized.device)
    if out.numel():
        call("elvis_roi_svtav1_offsets", _ptr(resized), resized.numel(), int(base_crf), int(qp_range), _ptr(out), _stream())
    return out


def rgb_to_i420(frames: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """(T, H, W, 3) uint8 RGB -> (T, H*W*3/2) uint8 I420 (cv2.COLOR_RGB2YUV_I420, utils.py:460)."""
    _check_cuda(frames, torch.uint8, "frames")
    if frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError("frames must be (T, H, W, 3)")
    T, H, W, _ = frames.shape
    if H % 2 or W % 2:
        raise ValueError("4:2:0 needs even frame dimensions")
    if out is None:
        out = torch.empty((T, H * W * 3 // 2), dtype=torch.uint8, device=frames.device)
    elif tuple(out.shape) != (T, H * W * 3 // 2) or not out.is_contiguous():
        raise ValueError("out must be a contiguous (T, H*W*3/2) buffer")
    if out.numel() == 0:
        return out
    y = out[:, :H * W].view(T, H, W)
    u = out[:, H * W:H * W * 5 // 4].view(T, H // 2, W // 2)
    v = out[:, H * W * 5 // 4:].view(T, H // 2, W // 2)
    planes = [plane_of(frames, "frames"), plane_of(y, "y"), plane_of(u, "u"), plane_of(v, "v")]
    call("elvis_rgb_to_i420", *[C.byref(p) for p in planes], T, _stream())
    return out


# ---------------------------------------------------------------- pyramid pieces, level-map gray codec (8f rank 4)
def area_downscale(clip: torch.Tensor, factor: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """cv2.resize(frame, (W // factor, H // factor), INTER_AREA) of every uint8 frame of (T, H, W[, C])
    for an integer factor that divides H and W."""
    T, H, W = clip.shape[:3]
    if factor <= 0 or H % factor or W % factor:
        raise ValueError("factor must divide the frame dimensions")
    shape = (T, H // factor, W // factor) + tuple(clip.shape[3:])
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=clip.device)
    elif tuple(out.shape) != shape:
        raise ValueError(f"out must be {shape}")
    if out.numel():
        src, dst = plane_of(clip), plane_of(out, "out")
        call("elvis_area_downscale", C.byref(src), C.byref(dst), T, int(factor), _stream())
    return out


def merge_blocks_(dst: torch.Tensor, src: torch.Tensor, factors: torch.Tensor, threshold: int, block_px: int) -> torch.Tensor:
    """In place: dst block <- src block wherever factors (T, By, Bx) int32 <= threshold."""
    _check_cuda(factors, torch.int32, "factors")
    factors = factors.contiguous()
    T, by, bx = factors.shape
    if dst.shape != src.shape or dst.shape[0] != T:
        raise ValueError("dst, src and factors disagree")
    if dst.shape[1] != by * block_px or dst.shape[2] != bx * block_px:
        raise ValueError("Image dimensions must be divisible by block_size.")
    if dst.numel():
        s, d = plane_of(src, "src"), plane_of(dst, "dst")
        call("elvis_merge_blocks", C.byref(s), C.byref(d), T, int(block_px), by, bx, _ptr(factors), int(threshold), _stream())
    return dst


def levels_to_gray(maps: torch.Tensor, min_value: int, max_value: int) -> torch.Tensor:
    """int32 level maps -> uint8 gray frames, elvis.py:2200-2202."""
    _check_cuda(maps, torch.int32, "maps")
    maps = maps.contiguous()
    out = torch.empty(maps.shape, dtype=torch.uint8, device=maps.device)
    if out.numel():
        call("elvis_levels_to_gray", _ptr(maps), maps.numel(), int(min_value), int(max_value), _ptr(out), _stream())
    return out


def gray_to_levels(gray: torch.Tensor, min_value: float, max_value: float) -> torch.Tensor:
    """uint8 gray frames -> uint8 level maps, elvis.py:2238-2240."""
    _check_cuda(gray, torch.uint8, "gray")
    gray = gray.contiguous()
    out = torch.empty(gray.shape, dtype=torch.uint8, device=gray.device)
    if out.numel():
        call("elvis_gray_to_levels", _ptr(gray), gray.numel(), float(min_value), float(max_value), _ptr(out), _stream())
    return out


def resize_nearest(maps: torch.Tensor, dst_h: int, dst_w: int) -> torch.Tensor:
    """cv2.resize(map, (dst_w, dst_h), INTER_NEAREST) for every map of a dense (T, h, w) tensor with
    1-, 4- or 8-byte elements."""
    if not maps.is_cuda:
        raise TypeError("maps must be a CUDA tensor (elvis_b200 has no CPU path)")
    if maps.dim() != 3:
        raise ValueError("maps must be (T, h, w)")
    maps = maps.contiguous()
    T, sh, sw = maps.shape
    out = torch.empty((T, dst_h, dst_w), dtype=maps.dtype, device=maps.device)
    if out.numel() == 0:
        return out
    yi = torch.from_numpy(_tables.nearest_index(sh, dst_h)).to(maps.device)
    xi = torch.from_numpy(_tables.nearest_index(sw, dst_w)).to(maps.device)
    call("elvis_resize_nearest", _ptr(maps), maps.element_size(), T, sh, sw, _ptr(out), dst_h, dst_w, _ptr(yi), _ptr(xi), _stream())
    return out


def resize_linear_float(maps: torch.Tensor, dst_h: int, dst_w: int) -> torch.Tensor:
    """cv2.resize(map, (dst_w, dst_h), INTER_LINEAR) for every float32 / float64 map of a dense (T, h, w)
    tensor (utils.py:1127-1128; elvis.py:2068-2073).  Sources with a single row or column take another
    cv2 path and are not supported."""
    if not maps.is_cuda:
        raise TypeError("maps must be a CUDA tensor (elvis_b200 has no CPU path)")
    if maps.dim() != 3:
        raise ValueError("maps must be (T, h, w)")
    dt = _float_dtype(maps, "maps")
    maps = maps.contiguous()
    T, sh, sw = maps.shape
    if sh < 2 or sw < 2:
        raise NotImplementedError("INTER_LINEAR of a single-row / single-column map (cv2 takes another path) is not supported")
    out = torch.empty((T, dst_h, dst_w), dtype=maps.dtype, device=maps.device)
    if out.numel() == 0:
        return out
    fused = dt == F64
    tabs = [torch.from_numpy(a).to(maps.device) for a in (*_tables.linear_float_index(sh, dst_h, fused),
                                                           *_tables.linear_float_index(sw, dst_w, fused))]
    call("elvis_resize_linear_float", _ptr(maps), dt, T, sh, sw, _ptr(out), dst_h, dst_w, *[_ptr(t) for t in tabs], _stream())
    return out


def rgb_to_gray(frames: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """(T, H, W, 3) uint8 RGB -> (T, H, W) uint8 luma, cv2.COLOR_RGB2GRAY's fixed point."""
    _check_cuda(frames, torch.uint8, "frames")
    if frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError("frames must be (T, H, W, 3)")
    T, H, W, _ = frames.shape
    if out is None:
        out = torch.empty((T, H, W), dtype=torch.uint8, device=frames.device)
    elif tuple(out.shape) != (T, H, W):
        raise ValueError("out must be (T, H, W)")
    if out.numel():
        src, dst = plane_of(frames, "frames"), plane_of(out, "out")
        call("elvis_rgb_to_gray", C.byref(src), C.byref(dst), T, _stream())
    return out


def refill_map(mask: torch.Tensor, capacity: int) -> torch.Tensor:
    """(T, By, Bx) uint8 removal mask -> int32 (T, By, Bx): the row-major rank of every kept block among
    the kept blocks of its frame (the shrunk block it takes back, presley.py:806-819), -1 for removed
    blocks and for ranks >= capacity."""
    mask = _mask_arg(mask)
    T, by, bx = mask.shape
    out = torch.empty((T, by, bx), dtype=torch.int32, device=mask.device)
    if out.numel():
        call("elvis_refill_map", _ptr(mask), T, by * bx, int(capacity), _ptr(out), _stream())
    return out


# every public operator runs on the device of its first tensor argument
for _name in ("score_sc_tc", "minmax", "combine_removability", "normalize_", "importance_scores", "select_rows", "shrink",
              "stretch", "move_yuv420", "degrade_yuv420", "levels_from_scores", "degrade_blur", "degrade_downsample", "dct_dampen",
              "restore_unsharp", "restore_lanczos", "temporal_blend_", "pack_mask_bits", "unpack_mask_bits",
              "pack_levels_2bit", "unpack_levels_2bit", "rowcol_plan", "rowcol_expand", "invert_block_map", "gather_blocks",
              "roi_kvazaar", "roi_prepare_f32", "resize_area_f32", "roi_svtav1_offsets", "rgb_to_i420", "area_downscale",
              "merge_blocks_", "levels_to_gray", "gray_to_levels", "resize_nearest", "resize_linear_float", "rgb_to_gray",
              "refill_map"):
    globals()[_name] = _on_device(globals()[_name])
del _name
