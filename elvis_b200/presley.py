"""Drop-in mirrors of the hot-path functions of the reference's `presley.py`: the batch
shrink/stretch wrappers (presley.py:761-827), the generic adaptive degradation
(presley.py:968-1039) and an `analyze_frames` stand-in for the external EVCA call
(presley.py:202).  The batch functions move the whole clip to the GPU once."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, List, Tuple

import numpy as np
import torch

from . import ops
from . import utils as _utils
from .elvis import _frames_to_dev, _frames_to_host, _to_dev, reference_dct_size


# ------------------------------------------------------------------ EVCA stand-in
@dataclass
class EVCAConfig:
    """`EVCAConfig(block_size=bs)` of presley.py:202.  dct_size None = one transform per block (the
    reference's call); 8 = north_star's 8 x 8 tiling."""
    block_size: int = 16
    dct_size: int | None = None


@dataclass
class Complexities:
    SC: np.ndarray
    TC: np.ndarray


def rgb_to_luma(frames: torch.Tensor) -> torch.Tensor:
    """(T, H, W, 3) uint8 RGB -> (T, H, W) uint8 luma on the GPU: cv2.COLOR_RGB2GRAY's 15-bit fixed
    point, (9798 R + 19235 G + 3735 B + 2^14) >> 15 (pinned against cv2 in tests/test_oracle.py)."""
    return ops.rgb_to_gray(frames)


def analyze_frames(frames: np.ndarray, config: EVCAConfig) -> Complexities:
    """Stand-in for `evca.analyze_frames(np.array(frames), EVCAConfig(block_size=bs))`
    (presley.py:202): per-block SC/TC of the clip's luma, float64 (T, By, Bx).  frames: (T, H, W)
    luma or (T, H, W, 3) RGB uint8."""
    f = _to_dev(np.asarray(frames), np.uint8)
    y = f if f.dim() == 3 else rgb_to_luma(f)
    sc, tc, _ = ops.score_sc_tc(y.contiguous(), config.block_size, dct_size=reference_dct_size(config.block_size, config.dct_size))
    return Complexities(sc.double().cpu().numpy(), tc.double().cpu().numpy())


calculate_importance_scores = _utils.calculate_importance_scores   # presley.py:129-152 == utils.py:665-688
shrink_frame_row_only = _utils.shrink_frame_row_only               # presley.py:713-757 == utils.py:692-736


# ------------------------------------------------------------------ batch shrink / stretch
def shrink_video_frames(frames: List[np.ndarray], importance_scores: List[np.ndarray], block_size: int,
                        shrink_amount: float, method: Callable = shrink_frame_row_only) -> Tuple[List[np.ndarray], List[Any]]:
    """presley.py:761-784.  With the row-only method the whole clip is processed in one
    batch on the GPU; any other callable is applied frame by frame like the reference."""
    if method is not shrink_frame_row_only or len(frames) == 0:
        out = [method(f, s, block_size, shrink_amount) for f, s in zip(frames, importance_scores)]
        return [o[0] for o in out], [o[1] for o in out]
    clip = _frames_to_dev(frames)
    h, w = clip.shape[1:3]
    by, bx = h // block_size, w // block_size
    k, out_bx = _utils.row_only_plan(by, bx, shrink_amount)
    scores = _to_dev(np.stack([np.asarray(s, np.float64)[:by, :bx] for s in importance_scores]))
    mask = ops.select_rows(scores, _to_dev(k), ops.REMOVE_LOW)
    shrunk = _frames_to_host(ops.shrink(clip[:, :by * block_size, :bx * block_size], mask, block_size, out_bx))
    mask_h = mask.cpu().numpy().astype(bool)
    return shrunk, [mask_h[i] for i in range(len(frames))]


def stretch_video_frames(shrunken_frames: List[np.ndarray], removal_masks: List[np.ndarray], block_size: int) -> List[np.ndarray]:
    """presley.py:787-827: the kept positions of a frame take the shrunk blocks back in ROW-MAJOR
    order over the whole frame (kept block i <- shrunk block (i // sbx, i % sbx), bounds-guarded).
    That equals the per-row refill of utils.stretch_frame_row_only only when every row kept the same
    number of blocks; after a partial last shrink pass (int(By*Bx*shrink) % By != 0) the rows differ
    and blocks wrap across rows, exactly as in the reference."""
    if len(shrunken_frames) == 0:
        return []
    masks = _to_dev(np.stack([np.asarray(m) != 0 for m in removal_masks]), np.uint8)
    clip = _frames_to_dev(shrunken_frames)
    by, bx = masks.shape[1:]
    sby, sbx = clip.shape[1] // block_size, clip.shape[2] // block_size
    if sbx == 0 or sby == 0:
        z = np.zeros((by * block_size, bx * block_size) + tuple(clip.shape[3:]), np.uint8)
        return [z.copy() for _ in shrunken_frames]
    refill = ops.refill_map(masks, sby * sbx)
    return _frames_to_host(ops.gather_blocks(clip[:, :sby * block_size, :sbx * block_size], refill, block_size, by, bx))


# ------------------------------------------------------------------ adaptive degradation
def generate_degradation_map(importance: np.ndarray, max_value: int) -> np.ndarray:
    """presley.py:968-975."""
    imp = _to_dev(np.asarray(importance), np.float64)
    return ops.levels_from_scores(imp, ops.LEVELS_INVERTED_ROUND, max_value).cpu().numpy()


def downscale_block(block: np.ndarray, scale: int) -> np.ndarray:
    """presley.py:978-983 (one block = a one-block frame)."""
    bs = block.shape[0]
    lv = torch.ones((1, 1, 1), dtype=torch.int32, device="cuda")
    out = ops.degrade_downsample(_to_dev(block)[None], lv, bs, [bs, max(1, bs // int(scale))])
    return out[0].cpu().numpy()


def blur_block(block: np.ndarray, rounds: int) -> np.ndarray:
    """presley.py:986-990."""
    bs = block.shape[0]
    r = torch.full((1, 1, 1), int(rounds), dtype=torch.int32, device="cuda")
    return ops.degrade_blur(_to_dev(block)[None], r, bs)[0].cpu().numpy()


def _degrade_clip(clip: torch.Tensor, levels: torch.Tensor, block_size: int, method: Callable) -> torch.Tensor:
    if method is downscale_block:
        top = max(1, int(levels.max().item()))
        smalls = [block_size] + [max(1, block_size // s) for s in range(1, top + 1)]
        return ops.degrade_downsample(clip, levels, block_size, smalls)
    if method is blur_block:
        return ops.degrade_blur(clip, levels, block_size)
    raise TypeError("method must be elvis_b200.presley.downscale_block or blur_block")


def degrade_frame(frame: np.ndarray, degradation_map: np.ndarray, block_size: int, method: Callable) -> np.ndarray:
    """presley.py:993-1013."""
    lv = _to_dev(np.asarray(degradation_map), np.int32)[None]
    return _degrade_clip(_to_dev(frame)[None], lv, block_size, method)[0].cpu().numpy()


def degrade_video_adaptive(frames: List[np.ndarray], importance_scores: List[np.ndarray], block_size: int,
                           max_value: int, method: Callable) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """presley.py:1016-1039, the whole clip in one batch."""
    if len(frames) == 0:
        return [], []
    clip = _frames_to_dev(frames)
    imp = _to_dev(np.stack(importance_scores), np.float64)
    levels = ops.levels_from_scores(imp, ops.LEVELS_INVERTED_ROUND, max_value)
    out = _frames_to_host(_degrade_clip(clip, levels, block_size, method))
    lv = levels.cpu().numpy()
    return out, [lv[i] for i in range(len(frames))]
