"""-m gpu, needs >= 2 GPUs (skipped on a one-GPU box): the frame-sharded scorer on real devices --
halo exchange and global min / max reductions over NCCL (NVLink) -- must give BIT-IDENTICAL scores
and masks to one GPU scoring the whole clip (elvis.py:264-278 split)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, y_all, bs, transport, results):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from elvis_b200 import ops, sharding
        T, H, W = y_all.shape
        a, b = sharding.frame_range(T, rank, world)
        clip = sharding.HaloClip(b - a, H, W, dev)
        clip.owned.copy_(y_all[a:b])
        k = (W // bs) // 2
        out = {}
        for rep in range(2):          # twice: the second pass reuses whatever state the transport keeps
            r = sharding.sharded_removability(clip, T, bs, 0.5, 0.5, rank, world)
            mask = ops.select_rows(r, k, ops.REMOVE_HIGH)
            imp = sharding.sharded_importance(clip, bs, 0.5, 0.5, rank, world)
            torch.cuda.synchronize()
            out[rep] = (r.cpu().numpy(), mask.cpu().numpy(), imp.cpu().numpy())
        assert all(np.array_equal(x, y_) for x, y_ in zip(out[0], out[1]))
        results[rank] = (a, b) + out[1]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("T,H,W,bs", [(11, 96, 160, 16), (30, 1088, 1920, 16), (9, 64, 256, 8)])
def test_sharded_equals_single_gpu(T, H, W, bs):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from _util import synth_luma
    from elvis_b200 import ops
    from elvis_b200.pipeline import ElvisV1
    from elvis_b200.yuv import Yuv420
    world = min(torch.cuda.device_count(), 4)
    y = synth_luma(T, H, W, seed=T)
    yd = torch.from_numpy(y).cuda()
    ref = ElvisV1(bs, 0.5, 0.5, 0.5).score(Yuv420(yd, None, None))
    ref_mask = ops.select_rows(ref, (W // bs) // 2, ops.REMOVE_HIGH)
    sc, tc, _ = ops.score_sc_tc(yd, bs)
    ref_imp = ops.importance_scores(sc, tc, None, 0.5, 0.5)
    ref, ref_mask, ref_imp = ref.cpu().numpy(), ref_mask.cpu().numpy(), ref_imp.cpu().numpy()
    results = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), torch.from_numpy(y), bs, "nccl", results), nprocs=world, join=True)
    for rank in range(world):
        a, b, r, m, imp = results[rank]
        assert np.array_equal(r, ref[a:b]), (rank, "scores")
        assert np.array_equal(m, ref_mask[a:b]), (rank, "mask")
        assert np.array_equal(imp, ref_imp[a:b]), (rank, "importance")
