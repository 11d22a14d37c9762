"""TEST/BENCH INFRASTRUCTURE ONLY -- compares the outputs of the CUDA path with the oracle's on the
same clip and returns a verdict dict (used by tests/test_gpu_fullsize.py and by the parity check
bench.py runs after its timed region; never on a product path).

Bars: SC/TC-derived float64 scores within `score_atol` of the oracle's (the
scores are normalised to [0, 1]; the device computes the DCT in fp32, the oracle in float64); masks
bit-exact wherever the oracle's own decision margin -- the score gap between the last removed and
the first kept block of a row -- exceeds `tie_margin`; shrunk and stretched planes bit-exact against
the oracle applied to the mask the device produced, and against the oracle's own planes on every
frame whose mask is identical."""
from __future__ import annotations

import numpy as np

from . import ref_port as P


def row_margins(scores: np.ndarray, k: int) -> np.ndarray:
    """(T, By, Bx) removability scores -> (T, By) gap between the k-th and (k+1)-th highest score of
    every block row (inf when the row has no decision to make)."""
    bx = scores.shape[-1]
    if k <= 0 or k >= bx:
        return np.full(scores.shape[:-1], np.inf)
    s = np.sort(scores, axis=-1)
    return s[..., bx - k] - s[..., bx - k - 1]


def compare_v1(gpu: dict, cpu: dict, block_size: int, k: int, score_atol: float = 2e-6, tie_margin: float = 1e-5) -> dict:
    """gpu / cpu: {"scores" (T,By,Bx) f64, "mask" (T,By,Bx), "sy","su","sv" shrunk planes, "fy","fu","fv"
    stretched planes} as numpy arrays.  Returns counts and booleans; "ok" is the overall verdict."""
    gs, cs = np.asarray(gpu["scores"]), np.asarray(cpu["scores"])
    gm, cm = np.asarray(gpu["mask"]) != 0, np.asarray(cpu["mask"]) != 0
    out = {"frames": int(gs.shape[0]), "score_max_abs_err": float(np.abs(gs - cs).max()), "score_atol": score_atol}
    diff_rows = (gm != cm).any(axis=-1)                       # (T, By)
    margins = row_margins(cs, k)
    out["mask_blocks"] = int(gm.size)
    out["mask_mismatch_blocks"] = int((gm != cm).sum())
    out["mask_mismatch_rows"] = int(diff_rows.sum())
    out["mask_mismatch_rows_outside_ties"] = int((diff_rows & (margins > tie_margin)).sum())
    out["mask_equal"] = out["mask_mismatch_blocks"] == 0
    same = ~diff_rows.any(axis=-1)                            # frames whose whole mask agrees
    out["frames_with_identical_mask"] = int(same.sum())
    planes_ok_cpu, planes_ok_port = True, True
    for tag, pb in (("y", block_size), ("u", block_size // 2), ("v", block_size // 2)):
        gsh, gfu = np.asarray(gpu["s" + tag]), np.asarray(gpu["f" + tag])
        planes_ok_cpu &= bool(np.array_equal(gsh[same], np.asarray(cpu["s" + tag])[same]))
        planes_ok_cpu &= bool(np.array_equal(gfu[same], np.asarray(cpu["f" + tag])[same]))
        for t in np.flatnonzero(~same):                       # rare: oracle data movement on the device's mask
            rs = P.shrink_plane(np.asarray(cpu["in_" + tag])[t], gm[t].astype(np.uint8), pb) if "in_" + tag in cpu else None
            if rs is None:
                continue
            planes_ok_port &= bool(np.array_equal(gsh[t], rs))
            planes_ok_port &= bool(np.array_equal(gfu[t], P.stretch_plane(rs, gm[t].astype(np.uint8), pb)))
    out["shrunk_stretched_equal"] = bool(planes_ok_cpu and planes_ok_port)
    out["ok"] = bool(out["score_max_abs_err"] <= score_atol and out["mask_mismatch_rows_outside_ties"] == 0
                     and out["shrunk_stretched_equal"])
    return out


def compare_planes(gpu: dict, cpu: dict, tol: int = 0) -> dict:
    """v2 degradations: per-plane max |difference| (tol = 0: bit-exact)."""
    out = {"max_abs_diff": 0, "mismatch_pixels": 0}
    for tag in ("y", "u", "v"):
        d = np.abs(np.asarray(gpu[tag]).astype(np.int16) - np.asarray(cpu[tag]).astype(np.int16))
        out["max_abs_diff"] = max(out["max_abs_diff"], int(d.max()) if d.size else 0)
        out["mismatch_pixels"] += int((d > tol).sum())
    out["tolerance"] = tol
    out["ok"] = out["mismatch_pixels"] == 0
    return out
