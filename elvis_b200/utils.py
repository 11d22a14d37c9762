"""Drop-in mirrors of the hot-path functions of the reference's `utils.py` (PRESLEY-era
API): importance scoring, row-only shrink/stretch and the adaptive degradations.  Same
signatures and return values as the reference; the work runs on the sm_100a kernels."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import ops
from .elvis import _frames_to_dev, _frames_to_host, _packed_clip, _to_dev


def calculate_importance_scores(frames, block_size: int, alpha: float, beta: float, complexities,
                                foreground_masks: np.ndarray) -> List[np.ndarray]:
    """utils.py:665-688.  complexities: object with .SC/.TC (T, By, Bx) -- EVCA's result or
    elvis_b200.presley.analyze_frames'.  `frames` and `block_size` are unused, as in the
    reference."""
    sc = _to_dev(np.asarray(complexities.SC), np.float64)
    tc = _to_dev(np.asarray(complexities.TC), np.float64)
    fg = _to_dev(np.asarray(foreground_masks), np.float64)
    imp = ops.importance_scores(sc, tc, fg, alpha, beta).cpu().numpy()
    return [imp[i] for i in range(len(imp))]


def row_only_plan(blocks_y: int, blocks_x: int, shrink_amount: float):
    """Pass structure of utils.py:711-735 -> (k per block row (By,) int32, output width in
    blocks).  A pass removes one block from every row in row order until
    int(By*Bx*shrink) are gone; the width drops once per pass, also after a partial one."""
    target = int(blocks_y * blocks_x * shrink_amount)
    passes = 0
    k = np.zeros(blocks_y, np.int32)
    removed = 0
    while removed < target and blocks_x - passes > 1:
        n = min(blocks_y, target - removed)
        k[:n] += 1
        removed += n
        passes += 1
    return k, blocks_x - passes


def shrink_frame_row_only(frame: np.ndarray, importance: np.ndarray, block_size: int,
                          shrink_amount: float) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:692-736 -> (shrunken, removal_mask bool, True = removed)."""
    h, w = frame.shape[:2]
    by, bx = h // block_size, w // block_size
    k, out_bx = row_only_plan(by, bx, shrink_amount)
    scores = _to_dev(np.asarray(importance, np.float64)[:by, :bx])[None]
    mask = ops.select_rows(scores, _to_dev(k), ops.REMOVE_LOW)
    clip = _packed_clip(frame)[:, :by * block_size, :bx * block_size]      # utils.py:707 crop
    shrunk = ops.shrink(clip, mask, block_size, out_bx)
    return shrunk[0].cpu().numpy(), mask[0].cpu().numpy().astype(bool)


def stretch_frame_row_only(shrunk_frame: np.ndarray, removal_mask: np.ndarray, block_size: int) -> np.ndarray:
    """utils.py:739-759."""
    mask = _to_dev(np.asarray(removal_mask) != 0, np.uint8)[None]
    by, bx = mask.shape[1:]
    sby, sbx = shrunk_frame.shape[0] // block_size, shrunk_frame.shape[1] // block_size
    if sbx == 0:
        return np.zeros((by * block_size, bx * block_size) + shrunk_frame.shape[2:], shrunk_frame.dtype)
    clip = _packed_clip(shrunk_frame)[:, :sby * block_size, :sbx * block_size]
    return ops.stretch(clip, mask, block_size)[0].cpu().numpy()


# ---------------------------------------------------------------- row+column shrink (8f rank 2)
def _rowcol_shrink(frame: np.ndarray, importance: np.ndarray, block_size: int, shrink_amount: float):
    h, w = frame.shape[:2]
    by, bx = h // block_size, w // block_size
    target = int(by * bx * shrink_amount)
    fby, fbx, counts = ops.rowcol_dims(by, bx, target)
    imp = _to_dev(np.asarray(importance, np.float64)[:by, :bx])[None]
    mask, pos, pidx, pcnt, _ = ops.rowcol_plan(imp, target)
    clip = _packed_clip(frame)[:, :by * block_size, :bx * block_size]
    shrunk = ops.gather_blocks(clip, pos, block_size, fby, fbx)
    return shrunk[0].cpu().numpy(), mask[0].cpu().numpy().astype(bool), pos[0, :fby, :fbx], pidx[0], counts, bx


def shrink_frame_position_map(frame: np.ndarray, importance: np.ndarray, block_size: int,
                              shrink_amount: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """utils.py:763-836 -> (shrunken, removal_mask bool, position_map (by', bx', 2) of
    (orig_y, orig_x))."""
    shrunk, mask, pos, _, _, bx = _rowcol_shrink(frame, importance, block_size, shrink_amount)
    lin = pos.cpu().numpy().astype(np.int64)
    return shrunk, mask, np.stack([lin // bx, lin % bx], axis=-1)


def stretch_frame_position_map(shrunk_frame: np.ndarray, removal_mask: np.ndarray, position_map: np.ndarray,
                               block_size: int) -> np.ndarray:
    """utils.py:839-858: every shrunk block back to position_map[j, i]; on duplicates the last one
    in row-major order wins, as in the reference's loop."""
    by, bx = np.shape(removal_mask)
    sby, sbx = shrunk_frame.shape[0] // block_size, shrunk_frame.shape[1] // block_size
    if shrunk_frame.shape[0] != sby * block_size or shrunk_frame.shape[1] != sbx * block_size:
        raise ValueError("cannot reshape shrunk frame into whole blocks")       # utils.py:846 reshape
    pm = np.asarray(position_map)[:sby, :sbx].astype(np.int64)
    if pm.size and (pm[..., 0].min() < -by or pm[..., 0].max() >= by or pm[..., 1].min() < -bx or pm[..., 1].max() >= bx):
        raise IndexError("position_map points outside the original block grid")
    lin = (pm[..., 0] % by) * bx + pm[..., 1] % bx if pm.size else np.zeros((sby, sbx), np.int64)
    if sby == 0 or sbx == 0:
        return np.zeros((by * block_size, bx * block_size) + shrunk_frame.shape[2:], shrunk_frame.dtype)
    inv = ops.invert_block_map(_to_dev(lin.reshape(1, -1), np.int32), by * bx).view(1, by, bx)
    return ops.gather_blocks(_packed_clip(shrunk_frame), inv, block_size, by, bx)[0].cpu().numpy()


def shrink_frame_removal_indices(frame: np.ndarray, importance: np.ndarray, block_size: int,
                                 shrink_amount: float) -> Tuple[np.ndarray, np.ndarray, list]:
    """utils.py:862-948 -> (shrunken, removal_mask bool, removal_indices: int32 arrays, row
    passes at even positions, column passes at odd ones)."""
    shrunk, mask, _, pidx, counts, _ = _rowcol_shrink(frame, importance, block_size, shrink_amount)
    pidx = pidx.cpu().numpy()
    return shrunk, mask, [pidx[i, :n].copy() for i, n in enumerate(counts)]


def stretch_frame_removal_indices(shrunk_frame: np.ndarray, removal_indices: list, orig_blocks_y: int, orig_blocks_x: int,
                                  block_size: int) -> np.ndarray:
    """utils.py:951-1018: replay the passes in reverse, black blocks at the recorded positions,
    crop to the original size."""
    sby, sbx = shrunk_frame.shape[0] // block_size, shrunk_frame.shape[1] // block_size
    if shrunk_frame.shape[0] != sby * block_size or shrunk_frame.shape[1] != sbx * block_size:
        raise ValueError("cannot reshape shrunk frame into whole blocks")       # utils.py:958 reshape
    P = len(removal_indices)
    L = max([len(a) for a in removal_indices] + [1])
    pidx = np.zeros((1, P, L), np.int32)
    pcnt = np.zeros((1, P), np.int32)
    for i, a in enumerate(removal_indices):
        pidx[0, i, :len(a)] = np.asarray(a, np.int32)
        pcnt[0, i] = len(a)
    gh, gw = sby + P // 2, sbx + (P + 1) // 2
    tail = shrunk_frame.shape[2:]
    if gh == 0 or gw == 0:
        return np.zeros((gh * block_size, gw * block_size) + tail, shrunk_frame.dtype)[:orig_blocks_y * block_size, :orig_blocks_x * block_size]
    if sby == 0 or sbx == 0:
        full = np.zeros((gh * block_size, gw * block_size) + tail, shrunk_frame.dtype)
    elif P == 0:
        full = shrunk_frame
    else:
        grid = ops.rowcol_expand(_to_dev(pidx), _to_dev(pcnt), sby, sbx)
        full = ops.gather_blocks(_packed_clip(shrunk_frame), grid, block_size, gh, gw)[0].cpu().numpy()
    return full[:orig_blocks_y * block_size, :orig_blocks_x * block_size]


def _block_importance(importance: np.ndarray, by: int, bx: int) -> np.ndarray:
    importance = np.asarray(importance)
    if importance.shape != (by, bx):   # utils.py:1127-1128: cv2.resize(map, (bx, by), INTER_LINEAR), on the GPU
        if importance.ndim != 2 or importance.dtype not in (np.float32, np.float64):
            raise TypeError("importance map must be a 2-D float32 / float64 array")
        importance = ops.resize_linear_float(_to_dev(importance)[None], by, bx)[0].cpu().numpy()
    return importance


def degrade_adaptive_downsample(frame: np.ndarray, importance: np.ndarray, block_size: int,
                                max_scale: int = 4) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:1101-1168 -> (frame, degradation_map int32 with values {0, 2, ..., max_scale})."""
    by, bx = frame.shape[0] // block_size, frame.shape[1] // block_size
    imp = _to_dev(_block_importance(importance, by, bx), np.float64)[None]
    levels = ops.levels_from_scores(imp, ops.LEVELS_INVERTED_BINS, max_scale)
    smalls = [block_size, block_size] + [max(1, block_size // s) for s in range(2, max_scale + 1)]
    out = ops.degrade_downsample(_packed_clip(frame), levels, block_size, smalls)
    return out[0].cpu().numpy(), levels[0].cpu().numpy()


def degrade_adaptive_blur(frame: np.ndarray, importance: np.ndarray, block_size: int,
                          max_rounds: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:1171-1217."""
    by, bx = frame.shape[0] // block_size, frame.shape[1] // block_size
    imp = _to_dev(_block_importance(importance, by, bx), np.float64)[None]
    rounds = ops.levels_from_scores(imp, ops.LEVELS_INVERTED_ROUND, max_rounds)
    out = ops.degrade_blur(_packed_clip(frame), rounds, block_size)
    return out[0].cpu().numpy(), rounds[0].cpu().numpy()


def restore_with_opencv_unsharp(frames: List[np.ndarray], degradation_maps: np.ndarray, block_size: int, halo: int = 0,
                                temporal_blend: float = 0.0, **kwargs) -> List[np.ndarray]:
    """utils.py:1320-1392: per-block unsharp mask driven by the level map, optional context halo
    and temporal blending; the whole clip in one batch."""
    if len(frames) == 0:
        return []
    clip = _frames_to_dev(frames)
    by, bx = clip.shape[1] // block_size, clip.shape[2] // block_size
    maps = np.zeros((len(frames), by, bx), np.int32)
    for i in range(min(len(frames), len(degradation_maps))):
        m = np.asarray(degradation_maps[i])
        if m.shape != (by, bx):      # utils.py:1343-1345: nearest resize of the map, via float32 as the reference does
            m = ops.resize_nearest(_to_dev(m.astype(np.float32))[None], by, bx)[0].cpu().numpy().astype(np.int32)
        maps[i] = m
    out = ops.restore_unsharp(clip, _to_dev(maps), block_size, halo=halo, max_level=max(1, int(maps.max())))
    if temporal_blend > 0:
        ops.temporal_blend_(out, temporal_blend)
    return _frames_to_host(out)


# utils.py:1253-1317: despite its name the reference's "lanczos" restorer runs the same unsharp mask
restore_with_opencv_lanczos = restore_with_opencv_unsharp


# ---------------------------------------------------------------- ROI side files and raw frames (8f rank 3)
def write_y4m(frames: List[np.ndarray], y4m_path: str, framerate: float) -> None:
    """utils.py:453-462: YUV4MPEG2 file, 4:2:0, frames converted like cv2.COLOR_RGB2YUV_I420."""
    height, width = frames[0].shape[:2]
    fps_num = int(round(framerate * 1000))
    clip = _frames_to_dev(frames)
    i420 = ops.rgb_to_i420(clip).cpu().numpy()
    with open(y4m_path, "wb") as f:
        f.write(f"YUV4MPEG2 W{width} H{height} F{fps_num}:1000 Ip A1:1 C420\n".encode())
        for t in range(len(frames)):
            f.write(b"FRAME\n")
            f.write(i420[t].tobytes())


def _same_shape_runs(maps):
    """Consecutive runs of equally shaped maps (the reference accepts a list of per-frame arrays)."""
    start = 0
    for i in range(1, len(maps) + 1):
        if i == len(maps) or np.shape(maps[i]) != np.shape(maps[start]):
            yield start, i
            start = i


def create_kvazaar_roi_file(importance_scores: List[np.ndarray], roi_path: str, base_qp: int, qp_range: int = 15) -> None:
    """utils.py:1026-1053: per frame `int32 w, h` then `int8 dqp[h][w]`."""
    with open(roi_path, "wb") as f:
        for a, b in _same_shape_runs(importance_scores):
            imp = _to_dev(np.stack([np.asarray(m) for m in importance_scores[a:b]]), np.float64)
            dqp = ops.roi_kvazaar(imp, base_qp, qp_range).cpu().numpy()
            h, w = dqp.shape[1:]
            for t in range(b - a):
                f.write(np.array([w, h], dtype=np.int32).tobytes())
                f.write(dqp[t].tobytes())


def create_svtav1_roi_file(importance_scores: List[np.ndarray], roi_path: str, base_crf: int, qp_range: int, width: int,
                           height: int) -> None:
    """utils.py:1056-1092: one text line per frame, QP offsets of the 64x64 superblocks in row order."""
    blocks_x, blocks_y = (width + 63) // 64, (height + 63) // 64
    with open(roi_path, "w") as f:
        for a, b in _same_shape_runs(importance_scores):
            imp = _to_dev(np.stack([np.asarray(m) for m in importance_scores[a:b]]), np.float64)
            resized = ops.resize_area_f32(ops.roi_prepare_f32(imp, 0), blocks_y, blocks_x)
            offsets = ops.roi_svtav1_offsets(resized, base_crf, qp_range).cpu().numpy()
            for t in range(b - a):
                f.write(f"{a + t} " + " ".join(map(str, offsets[t].flatten().astype(int))) + "\n")
