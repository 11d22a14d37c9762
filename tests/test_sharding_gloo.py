"""world_size-2 gloo test of the frame-sharding logic (halo exchange, min/max all-reduces,
extended-range bookkeeping) on CPU: the arithmetic is an oracle-backed stand-in for
elvis_b200.ops with the same call signatures, so only the distributed plumbing is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from elvis_b200 import sharding
from oracle import ref_port as P
from oracle import spec_scoring


class OracleKernels:
    """CPU stand-in with the signatures of elvis_b200.ops used by sharding.py."""

    @staticmethod
    def score_sc_tc(y, block_size, prev_halo=None, minmax_range=None):
        ynp = y.numpy()
        sc, tc = spec_scoring.sc_tc(ynp, block_size, prev=None if prev_halo is None else prev_halo.numpy())
        lo, hi = (0, len(ynp)) if minmax_range is None else minmax_range
        mm = np.array([sc[lo:hi].min(), sc[lo:hi].max(), tc[lo:hi].min(), tc[lo:hi].max()])
        return torch.from_numpy(sc), torch.from_numpy(tc), torch.from_numpy(mm)

    @staticmethod
    def combine_removability(sc, tc, norm, alpha, beta, background=None, t_begin=0, t_count=None, is_first=True,
                             is_last=True, clip_frames=None):
        sc, tc, n = sc.numpy(), tc.numpy(), norm.numpy()
        nz = lambda x, lo, hi: (x - lo) / (hi - lo) if hi > lo else x   # noqa: E731
        scn, tcn = nz(sc, n[0], n[1]), nz(tc, n[2], n[3])
        t_count = len(sc) - t_begin if t_count is None else t_count

        def r_at(t, last):
            r = scn[t] if last else alpha * scn[t] + (1 - alpha) * tcn[t + 1]
            if background is not None:
                r = np.where(background[t].numpy() != 0, r * 10.0, r)
            return r
        out = np.zeros((t_count,) + sc.shape[1:])
        smooth = beta < 1 and clip_frames >= 2
        for i in range(t_count):
            t = t_begin + i
            r = r_at(t, is_last and i == t_count - 1)
            if smooth and not (is_first and i == 0):
                r = beta * r + (1 - beta) * r_at(t - 1, False)
            out[i] = r
        return torch.from_numpy(out), torch.tensor([out.min(), out.max()], dtype=torch.float64)

    @staticmethod
    def normalize_(x, mm):
        lo, hi = float(mm[0]), float(mm[1])
        if hi > lo:
            x.copy_((x - lo) / (hi - lo))
        return x

    @staticmethod
    def importance_scores(sc, tc, fg, alpha, beta, t_begin=0, t_count=None, is_first=True, is_last=True):
        sc, tc = sc.numpy(), tc.numpy()
        t_count = len(sc) - t_begin if t_count is None else t_count

        def c_at(t, last):
            return sc[t] if last else alpha * sc[t] + (1 - alpha) * tc[t + 1]
        out = np.zeros((t_count,) + sc.shape[1:])
        for i in range(t_count):
            t = t_begin + i
            c = c_at(t, is_last and i == t_count - 1)
            if not (is_first and i == 0):
                c = beta * c + (1 - beta) * c_at(t - 1, False)
            if fg is not None:
                f = fg[t].numpy().astype(np.float64).copy()
                f[f < 0.5] = -1.0
                c = c * f
            out[i] = (c - c.min()) / (c.max() - c.min() + 1e-8)
        return torch.from_numpy(out)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, y_all, bg_all, fg_all, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, H, W = y_all.shape
        a, b = sharding.frame_range(T, rank, world)
        clip = sharding.HaloClip(b - a, H, W, "cpu")
        clip.owned.copy_(y_all[a:b])
        bg = torch.zeros((b - a + 2,) + bg_all.shape[1:], dtype=torch.uint8)
        bg[1:b - a + 1] = bg_all[a:b]
        if a > 0:
            bg[0] = bg_all[a - 1]
        r = sharding.sharded_removability(clip, T, 16, 0.3, 0.5, rank, world, background=bg, kernels=OracleKernels)
        fg = torch.ones((b - a + 2,) + fg_all.shape[1:], dtype=torch.float64)
        fg[1:b - a + 1] = fg_all[a:b]
        if a > 0:
            fg[0] = fg_all[a - 1]
        imp = sharding.sharded_importance(clip, 16, 0.3, 0.5, rank, world, foreground=fg, kernels=OracleKernels)
        results[rank] = (a, b, r.numpy().copy(), imp.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 7), (2, 2), (3, 8)])
def test_sharded_scores_equal_single_process(world, T):
    from _util import synth_luma
    y = synth_luma(T, 32, 48, seed=T)
    rng = np.random.default_rng(T)
    bg = rng.random((T, 2, 3)) > 0.6
    fg = rng.random((T, 2, 3))
    sc, tc = spec_scoring.sc_tc(y, 16)
    ref_r = P.combine_removability(sc, tc, 0.3, 0.5, bg)
    ref_i = P.importance_scores(sc, tc, 0.3, 0.5, fg)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), torch.from_numpy(y), torch.from_numpy(bg.astype(np.uint8)),
                            torch.from_numpy(fg), results), nprocs=world, join=True)
    got_r, got_i = np.zeros_like(ref_r), np.zeros_like(ref_i)
    for rank in range(world):
        a, b, r, imp = results[rank]
        assert (a, b) == sharding.frame_range(T, rank, world)
        got_r[a:b], got_i[a:b] = r, imp
    np.testing.assert_allclose(got_r, ref_r, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(got_i, ref_i, rtol=1e-12, atol=1e-15)


def test_frame_range_is_the_reference_split():
    """elvis.py:264-278: base = T // G, the first T % G ranks get one extra, contiguous."""
    for T in (1, 5, 8, 120, 601):
        for G in (1, 2, 3, 8):
            spans = [sharding.frame_range(T, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == T
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
