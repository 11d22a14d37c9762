"""Drop-in mirrors of the hot-path functions of the reference's `elvis.py` (same names,
argument order, return tuples, dtypes and error behaviour; NumPy arrays in and out), executed
by the sm_100a kernels behind include/elvis_b200.h.  Each function cites the reference lines
it replaces.  Host<->device copies happen here; use elvis_b200.ops / elvis_b200.pipeline to
keep clips resident on the GPU.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np
import torch

from . import ops

_DEV = "cuda"


def _to_dev(a: np.ndarray, dtype=None) -> torch.Tensor:
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype))
    return torch.from_numpy(a).to(_DEV, non_blocking=False)


def _frames_to_dev(frames, dtype=np.uint8) -> torch.Tensor:
    """List of equally shaped host frames -> one (T, ...) device tensor, copied frame by frame: no np.stack of the whole
    clip on the host (a second full-size pageable buffer whose first touch costs more than the transfers)."""
    first = np.asarray(frames[0], dtype=dtype)
    out = torch.empty((len(frames),) + first.shape, dtype=torch.from_numpy(np.empty(0, dtype=first.dtype)).dtype, device=_DEV)
    for i, f in enumerate(frames):
        a = np.ascontiguousarray(np.asarray(f, dtype=dtype))
        if a.shape != first.shape:
            raise ValueError("all frames must share a shape")
        out[i].copy_(torch.from_numpy(a))
    return out


def _frames_to_host(clip: torch.Tensor) -> list:
    """(T, ...) device tensor -> list of host arrays, one download per frame."""
    return [clip[i].cpu().numpy() for i in range(clip.shape[0])]


def _packed_clip(image: np.ndarray) -> torch.Tensor:
    """(H, W, C) or (H, W) uint8 host image -> (1, H, W[, C]) device clip."""
    if image.dtype != np.uint8:
        raise TypeError("frames must be uint8")
    return _to_dev(image)[None]


def normalize_array(arr: np.ndarray) -> np.ndarray:
    """elvis.py:864-867."""
    x = _to_dev(arr, np.float64).reshape(-1)
    ops.normalize_(x, ops.minmax(x))
    return x.cpu().numpy().reshape(np.shape(arr))


# ---------------------------------------------------------------- scoring (elvis.py:968-1224)
def reference_dct_size(block_size: int, dct_size: int | None = None) -> int:
    """Transform size of the SC/TC features.  None = what the reference's calls ask EVCA for: the
    block size itself (`python -m evca.main ... -b block_size`, elvis.py:1022-1023;
    `EVCAConfig(block_size=bs)`, presley.py:202).  The planar pipelines and the benchmark pass 8
    explicitly (north_star: "8x8 DCT coefficients"; 16x16 blocks then sum four 8x8 tiles)."""
    return block_size if dct_size is None else int(dct_size)


def removability_from_luma(y: torch.Tensor, block_size: int, alpha: float = 0.5, smoothing_beta: float = 1,
                           background: torch.Tensor | None = None, dct_size: int | None = None) -> torch.Tensor:
    """Device-side body of calculate_removability_scores: (T, H, W) uint8 luma ->
    (T, By, Bx) float64 in [0, 1].  background: optional uint8 (T, By, Bx), non-zero =
    background block (elvis.py:1193).  dct_size: see reference_dct_size."""
    sc, tc, norm = ops.score_sc_tc(y, block_size, dct_size=reference_dct_size(block_size, dct_size))
    r, mm = ops.combine_removability(sc, tc, norm, alpha, smoothing_beta, background)
    return ops.normalize_(r, mm)


def removability_from_features(sc: np.ndarray, tc: np.ndarray, alpha: float = 0.5, smoothing_beta: float = 1,
                               background: np.ndarray | None = None) -> np.ndarray:
    """In-tree tail of calculate_removability_scores (elvis.py:1173-1220) on caller-supplied
    SC/TC (T, By, Bx) float64 -- e.g. EVCA's own CSVs.  Bit-exact against the reference."""
    scd, tcd = _to_dev(sc, np.float64), _to_dev(tc, np.float64)
    norm = torch.cat([ops.minmax(scd), ops.minmax(tcd)])
    bg = None if background is None else _to_dev(np.asarray(background) != 0, np.uint8)
    r, mm = ops.combine_removability(scd, tcd, norm, alpha, smoothing_beta, bg)
    return ops.normalize_(r, mm).cpu().numpy()


def calculate_removability_scores(raw_video_file: str, reference_frames_folder: str, width: int, height: int,
                                  block_size: int, alpha: float = 0.5, working_dir: str = ".",
                                  smoothing_beta: float = 1, *, dct_size: int | None = None) -> np.ndarray:
    """elvis.py:968-1224 (dct_size is this package's extra keyword, see reference_dct_size).  Reads the yuv420p file the reference hands to EVCA
    (elvis.py:1019), computes SC/TC on the GPU instead of the EVCA subprocess, and applies the
    in-tree combine.  Foreground masks are an external model's output (UFO, elvis.py:1109):
    if `<working_dir>/maps/ufo_masks/00001.png ...` exist they are applied exactly like the
    reference does (NEAREST resize to the block grid, x10 on zeros), otherwise skipped."""
    frame_bytes = width * height * 3 // 2
    n_file = os.path.getsize(raw_video_file) // frame_bytes
    frame_count = n_file
    if reference_frames_folder and os.path.isdir(reference_frames_folder):
        frame_count = min(n_file, len(os.listdir(reference_frames_folder)))      # elvis.py:978
    raw = np.memmap(raw_video_file, dtype=np.uint8, mode="r", shape=(n_file, frame_bytes))
    y = torch.from_numpy(np.ascontiguousarray(raw[:frame_count, :width * height])).to(_DEV)
    y = y.view(frame_count, height, width)
    by, bx = height // block_size, width // block_size
    background = None
    masks_dir = os.path.join(os.path.abspath(working_dir), "maps", "ufo_masks")
    if os.path.isdir(masks_dir):
        import cv2
        bg = np.zeros((frame_count, by, bx), np.uint8)
        for i in range(frame_count):
            path = os.path.join(masks_dir, f"{i + 1:05d}.png")
            if os.path.exists(path):
                m = cv2.imread(path, cv2.IMREAD_GRAYSCALE)      # file IO; the resize runs on the GPU
                small = ops.resize_nearest(_to_dev(m, np.uint8)[None], by, bx)           # elvis.py:1189-1193
                bg[i] = (small[0] == 0).cpu().numpy()
        background = _to_dev(bg)
    return removability_from_luma(y, block_size, alpha, smoothing_beta, background, dct_size).cpu().numpy()


# ---------------------------------------------------------------- block views (elvis.py:1369-1385, 1429-1434)
def split_image_into_blocks(image: np.ndarray, block_size: int) -> np.ndarray:
    """elvis.py:1369-1385 -- a strided host view, no device work."""
    h, w, c = image.shape
    if h % block_size != 0 or w % block_size != 0:
        raise ValueError("Image dimensions must be divisible by block_size.")
    return image.reshape(h // block_size, block_size, w // block_size, block_size, c).swapaxes(1, 2)


def combine_blocks_into_image(blocks: np.ndarray) -> np.ndarray:
    """elvis.py:1429-1434."""
    nby, nbx, bs, _, c = blocks.shape
    return blocks.swapaxes(1, 2).reshape(nby * bs, nbx * bs, c)


# ---------------------------------------------------------------- v1 shrink / stretch
def blocks_to_remove(shrink_amount: float, num_blocks_x: int) -> int:
    """elvis.py:1392-1396."""
    k = int(shrink_amount * num_blocks_x) if shrink_amount < 1.0 else int(shrink_amount)
    return min(k, num_blocks_x)


def apply_selective_removal(image: np.ndarray, frame_scores: np.ndarray, block_size: int,
                            shrink_amount: float) -> Tuple[np.ndarray, np.ndarray, List[List[int]]]:
    """elvis.py:1387-1427 -> (new_image, removal_mask int8 (By, Bx), removed columns per row)."""
    by, bx = frame_scores.shape
    if image.shape[0] % block_size or image.shape[1] % block_size:
        raise ValueError("Image dimensions must be divisible by block_size.")
    k = blocks_to_remove(shrink_amount, bx)
    scores = _to_dev(frame_scores, np.float64)[None]
    mask = ops.select_rows(scores, k, ops.REMOVE_HIGH)
    shrunk = ops.shrink(_packed_clip(image), mask, block_size, bx - k)
    mask_h = mask[0].cpu().numpy().astype(np.int8)
    coords = [np.flatnonzero(row).tolist() for row in mask_h]
    return shrunk[0].cpu().numpy(), mask_h, coords


def stretch_frame(shrunk_frame: np.ndarray, binary_mask: np.ndarray, block_size: int) -> np.ndarray:
    """elvis.py:1436-1455."""
    mask = _to_dev(np.asarray(binary_mask) != 0, np.uint8)[None]
    by, bx = mask.shape[1:]
    if shrunk_frame.shape[1] == 0:
        return np.zeros((by * block_size, bx * block_size) + shrunk_frame.shape[2:], shrunk_frame.dtype)
    return ops.stretch(_packed_clip(shrunk_frame), mask, block_size)[0].cpu().numpy()


# ---------------------------------------------------------------- v2 degradations
def filter_frame_downsample(image: np.ndarray, frame_scores: np.ndarray, block_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """elvis.py:2141-2169 -> (image, downsample_maps int32)."""
    if image.shape[0] % block_size or image.shape[1] % block_size:
        raise ValueError("Image dimensions must be divisible by block_size.")
    n_lv = int(np.log2(block_size))
    levels = ops.levels_from_scores(_to_dev(frame_scores, np.float64)[None], ops.LEVELS_ROUND, n_lv)
    # elvis.py:2147,2159: strength = float32(2**level); small = max(1, int(bs / strength))
    smalls = [block_size] + [max(1, int(block_size / np.float32(2.0 ** lv))) for lv in range(1, n_lv + 1)]
    out = ops.degrade_downsample(_packed_clip(image), levels, block_size, smalls)
    return out[0].cpu().numpy(), levels[0].cpu().numpy()


def filter_frame_gaussian(image: np.ndarray, frame_scores: np.ndarray, block_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """elvis.py:2171-2196 -> (image, blur_strengths int32)."""
    if image.shape[0] % block_size or image.shape[1] % block_size:
        raise ValueError("Image dimensions must be divisible by block_size.")
    rounds = ops.levels_from_scores(_to_dev(frame_scores, np.float64)[None], ops.LEVELS_ROUND, 10)
    out = ops.degrade_blur(_packed_clip(image), rounds, block_size)
    return out[0].cpu().numpy(), rounds[0].cpu().numpy()


# ---------------------------------------------------------------- OpenCV client restorer (8f rank 1)
def restore_blur_opencv_unsharp_mask(blurred_image: np.ndarray, blur_maps: np.ndarray, block_size: int) -> np.ndarray:
    """elvis.py:2822-2867 -- per-block unsharp mask, radius = blur rounds, amount = rounds / 2."""
    if blurred_image.shape[0] % block_size or blurred_image.shape[1] % block_size:
        raise ValueError("Image dimensions must be divisible by block_size.")
    levels = _to_dev(np.asarray(blur_maps), np.int32)[None]
    return ops.restore_unsharp(_packed_clip(blurred_image), levels, block_size)[0].cpu().numpy()


def restore_downsample_opencv_lanczos(downsampled_image: np.ndarray, downscale_maps: np.ndarray, block_size: int) -> np.ndarray:
    """elvis.py:2773-2820 -- per block with factor 2**level > 1: INTER_AREA down to
    max(1, block_size // factor), INTER_LANCZOS4 back up."""
    if downsampled_image.shape[0] % block_size or downsampled_image.shape[1] % block_size:
        raise ValueError("Image dimensions must be divisible by block_size.")
    maps = np.asarray(downscale_maps)
    top = max(0, int(maps.max())) if maps.size else 0
    if top == 0:                       # elvis.py:2793-2795: nothing was downsampled
        return downsampled_image
    smalls = [block_size] + [max(1, block_size // (2 ** lv)) for lv in range(1, top + 1)]
    levels = _to_dev(np.clip(maps, 0, top), np.int32)[None]
    return ops.restore_lanczos(_packed_clip(downsampled_image), levels, block_size, smalls)[0].cpu().numpy()


# ---------------------------------------------------------------- x265 per-block qpfile (8f rank 3)
def x265_ctu_size(block_size: int, width: int, height: int) -> int:
    """CTU size choice of encode_with_roi (elvis.py:2032-2053)."""
    valid = [16, 32, 64]
    largest = max(width, height)
    min_ctu = 64 if largest >= 4320 else 32 if largest >= 2160 else 16
    nearest = min(valid, key=lambda size: abs(size - block_size))
    if nearest < block_size:
        larger = [size for size in valid if size >= block_size]
        ctu = larger[0] if larger else valid[-1]
    else:
        ctu = nearest
    if ctu < min_ctu:
        ctu = [size for size in valid if size >= min_ctu][0]
    return ctu


def per_block_qp_maps(removability_scores: np.ndarray, block_size: int, width: int, height: int) -> Tuple[np.ndarray, int]:
    """Part 1 of encode_with_roi up to the aligned maps (elvis.py:2030-2074): scores in [0, 1] ->
    float32 QP offsets in [-1, 1] on the CTU grid (INTER_AREA when the CTU is at least a block, else
    INTER_LINEAR -- elvis.py:2068).  Returns (maps (T, rows, cols), ctu)."""
    import math
    T, by, bx = removability_scores.shape
    ctu = x265_ctu_size(block_size, width, height)
    cols, rows = math.ceil(width / ctu), math.ceil(height / ctu)
    qp = ops.roi_prepare_f32(_to_dev(removability_scores, np.float64), 1)
    if (rows, cols) != (by, bx):
        if ctu < block_size:            # CTU grid finer than the block grid: cv2's float32 bilinear resize
            qp = ops.resize_linear_float(qp, rows, cols)
        else:
            qp = ops.resize_area_f32(qp, rows, cols)
    return qp.cpu().numpy(), ctu


def write_per_block_qpfile(removability_scores: np.ndarray, block_size: int, width: int, height: int, qpfile_path: str) -> None:
    """The qpfile text of encode_with_roi (elvis.py:2076-2090): `frame P -1 bx,by,qp ...` per frame."""
    maps, _ = per_block_qp_maps(removability_scores, block_size, width, height)
    rows, cols = maps.shape[1:]
    with open(qpfile_path, "w") as f:
        for t in range(maps.shape[0]):
            parts = [f"{t} P -1"]
            parts.extend(f"{x},{y},{maps[t, y, x]:.4f}" for y in range(rows) for x in range(cols))
            f.write(" ".join(parts) + "\n")


# ---------------------------------------------------------------- adaptive pyramid reconstruction (8f rank 4)
def upscale_realesrgan_adaptive(downsampled_image: np.ndarray, downscale_maps: np.ndarray, block_size: int,
                                realesrgan_dir: str = None, *, upsample_fn=None) -> np.ndarray:
    """elvis.py:2522-2600.  The 2x upsampler is external (Real-ESRGAN in the reference): pass it as
    `upsample_fn` (image -> image of twice the size).  The pyramid around it -- INTER_AREA downscales of
    the input and the per-block restore of every stage -- runs on the GPU."""
    if upsample_fn is None:
        raise NotImplementedError("the Real-ESRGAN 2x upsampler is external to elvis_b200: pass upsample_fn")
    factors = np.power(2, np.asarray(downscale_maps)).astype(np.int32)
    max_factor = int(factors.max())
    height, width, _ = downsampled_image.shape
    if height % block_size or width % block_size:
        raise ValueError("Image dimensions must be divisible by block_size.")
    original = _packed_clip(downsampled_image)
    fac = _to_dev(factors)[None]
    current = ops.area_downscale(original, max_factor)[0].cpu().numpy() if max_factor > 1 else downsampled_image
    current_factor = max_factor // 2
    while current_factor >= 1:
        current = np.ascontiguousarray(upsample_fn(current))
        bs_now = block_size // current_factor
        if current.shape[0] % bs_now or current.shape[1] % bs_now:
            raise ValueError("Image dimensions must be divisible by block_size.")
        cur = _packed_clip(current)
        down = ops.area_downscale(original, current_factor) if current_factor > 1 else original
        if down.shape != cur.shape:
            raise ValueError("upsample_fn must double the image size")
        # blocks downsampled by <= the stage's factor come back from the (downscaled) input;
        # the others keep the upsampler's output and count as `current_factor` from now on
        current = ops.merge_blocks_(cur, down, fac, current_factor, bs_now)[0].cpu().numpy()
        fac = torch.clamp(fac, max=current_factor)
        current_factor //= 2
    return current


def strength_maps_to_gray(strength_maps: np.ndarray) -> np.ndarray:
    """The normalisation of encode_strength_maps (elvis.py:2200-2202): maps -> uint8 frames for the
    gray map video (PNG writing and the encode are external)."""
    maps = np.asarray(strength_maps)
    lo, hi = int(maps.min()), int(maps.max())
    if hi == lo:
        raise ValueError("constant strength maps cannot be normalised (the reference divides by zero)")
    return ops.levels_to_gray(_to_dev(maps, np.int32), lo, hi).cpu().numpy()


def gray_to_strength_maps(gray_frames: np.ndarray, min_val: float, max_val: float) -> np.ndarray:
    """The reconstruction of decode_strength_maps (elvis.py:2238-2243): decoded gray frames -> uint8
    maps; (min_val, max_val) = (0, 10) for blur rounds, (0, log2(block_size)) for downsample levels."""
    return ops.gray_to_levels(_to_dev(np.asarray(gray_frames), np.uint8), min_val, max_val).cpu().numpy()


# ---------------------------------------------------------------- side channels
def encode_strength_maps_to_npz(strength_maps: np.ndarray, output_path: str) -> None:
    """elvis.py:2247-2259 (uint8 maps, np.savez_compressed key `strength_maps`)."""
    if isinstance(strength_maps, list):
        strength_maps = np.stack(strength_maps, axis=0)
    if strength_maps.dtype != np.uint8:
        strength_maps = strength_maps.astype(np.uint8)
    np.savez_compressed(output_path, strength_maps=strength_maps)


def decode_strength_maps_from_npz(npz_path: str) -> np.ndarray:
    """elvis.py:2261-2272."""
    if not os.path.exists(npz_path):
        raise FileNotFoundError(f"Strength maps file not found: {npz_path}")
    return np.load(npz_path)["strength_maps"]


def pack_removal_masks(masks: np.ndarray) -> Tuple[np.ndarray, tuple]:
    """The mask side channel of run_elvis (elvis.py:4412-4418): np.packbits + shape."""
    m = _to_dev(np.asarray(masks) != 0, np.uint8)
    return ops.pack_mask_bits(m).cpu().numpy(), tuple(np.shape(masks))


def unpack_removal_masks(packed: np.ndarray, shape) -> np.ndarray:
    """elvis.py:4537-4539."""
    return ops.unpack_mask_bits(_to_dev(packed, np.uint8), shape).cpu().numpy()
