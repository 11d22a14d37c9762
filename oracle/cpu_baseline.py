"""TEST/BENCH INFRASTRUCTURE ONLY -- the CPU arm of bench.py: the oracle port of the
headline path (SC/TC spec -> elvis combine -> per-row top-k -> shrink -> stretch) run on a
planar YUV 4:2:0 clip with every host core.

The reference runs these functions one frame at a time in a single Python process
(elvis.py:4389-4394, 4550-4557); to give the CPU its best showing the frames are fanned out
over a fork()ed process pool (inputs inherited copy-on-write, outputs written into shared
memory, so no pickling of frames).  The global normalisation forces two phases: all SC/TC
first, then mask + shrink + stretch."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import ref_port as P
from . import spec_scoring

_G: dict = {}


def _shared(shape, dtype=np.uint8):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = mp.RawArray("B", max(1, n))
    return np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def _score_chunk(ab):
    a, b = ab
    y = _G["y"]
    return spec_scoring.sc_tc(y[a:b], _G["bs"], prev=y[a - 1] if a > 0 else None)


def _move_chunk(ab):
    a, b = ab
    bs, k = _G["bs"], _G["k"]
    for t in range(a, b):
        mask = P.select_rows(_G["scores"][t], k, P.REMOVE_HIGH)
        _G["mask"][t] = mask
        for name, pb in (("y", bs), ("u", bs // 2), ("v", bs // 2)):
            s = P.shrink_plane(_G[name][t], mask, pb)
            _G["s" + name][t] = s
            _G["f" + name][t] = P.stretch_plane(s, mask, pb)
    return b - a


def _chunks(n, parts):
    parts = max(1, min(parts, n))
    edges = np.linspace(0, n, parts + 1).astype(int)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(parts) if edges[i + 1] > edges[i]]


class CpuElvisV1:
    def __init__(self, y, u, v, block_size=16, shrink_amount=0.5, alpha=0.5, beta=0.5, workers=None):
        self.workers = workers or os.cpu_count() or 1
        T, H, W = y.shape
        bx = W // block_size
        k = P.blocks_to_remove_elvis(shrink_amount, bx)
        self.alpha, self.beta, self.T = alpha, beta, T
        _G.clear()
        _G.update(y=y, u=u, v=v, bs=block_size, k=k)
        _G["scores"] = _shared((T, H // block_size, bx), np.float64)
        _G["mask"] = _shared((T, H // block_size, bx), np.uint8)
        for name, pl, pb in (("y", y, block_size), ("u", u, block_size // 2), ("v", v, block_size // 2)):
            _G["s" + name] = _shared((T, pl.shape[1], (bx - k) * pb))
            _G["f" + name] = _shared(pl.shape)
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None

    def step(self) -> float:
        """One pass over the clip; returns elapsed seconds."""
        t0 = time.perf_counter()
        mapper = self.pool.map if self.pool else lambda f, xs: list(map(f, xs))
        parts = mapper(_score_chunk, _chunks(self.T, self.workers))
        sc = np.concatenate([p[0] for p in parts])
        tc = np.concatenate([p[1] for p in parts])
        self.sc, self.tc = sc, tc
        _G["scores"][:] = P.combine_removability(sc, tc, self.alpha, self.beta)
        mapper(_move_chunk, _chunks(self.T, self.workers))
        return time.perf_counter() - t0

    def outputs(self):
        out = {k: _G[k] for k in ("scores", "mask", "sy", "su", "sv", "fy", "fu", "fv")}
        out.update(in_y=_G["y"], in_u=_G["u"], in_v=_G["v"], sc=self.sc, tc=self.tc)
        return out

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
            self.pool = None


# ------------------------------------------------------------------ v2 degradations (configs[2], configs[3])
def _v2_chunk(ab):
    a, b = ab
    bs, kind = _G["bs"], _G["kind"]
    for t in range(a, b):
        m = _G["map"][t]
        for name, pb in (("y", bs), ("u", bs // 2), ("v", bs // 2)):
            src = _G[name][t]
            if kind == "blur":
                out = P.blur_plane(src, m, pb)
            elif kind == "downsample":
                out = P.downsample_plane(src, np.where(m > 0, np.maximum(1, pb >> m), pb), pb)
            else:
                from . import spec_dct_dampen
                out = spec_dct_dampen.dampen_plane(src, m, pb)
            _G["o" + name][t] = out
    return b - a


class CpuV2:
    """One v2 degradation of a planar clip with every host core: kind = "blur" (rounds map int32,
    elvis.py:2171-2196), "downsample" (pow2 level map int32, elvis.py:2141-2169) or "dampen" (strength
    map float, oracle/spec_dct_dampen.py); chroma blocks at half the block size."""

    def __init__(self, y, u, v, block_map, block_size=16, kind="blur", workers=None):
        self.workers = workers or os.cpu_count() or 1
        self.T = y.shape[0]
        _G.clear()
        _G.update(y=y, u=u, v=v, bs=block_size, kind=kind, map=block_map)
        for name, pl in (("y", y), ("u", u), ("v", v)):
            _G["o" + name] = _shared(pl.shape)
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None

    def step(self) -> float:
        t0 = time.perf_counter()
        mapper = self.pool.map if self.pool else lambda f, xs: list(map(f, xs))
        mapper(_v2_chunk, _chunks(self.T, self.workers))
        return time.perf_counter() - t0

    def outputs(self):
        return {k: _G["o" + k] for k in ("y", "u", "v")}

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
            self.pool = None
