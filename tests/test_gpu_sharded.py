"""-m gpu, needs >= 2 GPUs (skipped on a one-GPU box): the frame-sharded scorer on real devices --
halo exchange and global min / max reductions over NCCL, and over peer memory (copy-engine peer copies +
mailbox all-reduces, elvis_b200/peer.py) -- must give BIT-IDENTICAL scores and masks to one GPU scoring
the whole clip (elvis.py:264-278 split)."""
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
        from elvis_b200 import ops, peer, sharding
        T, H, W = y_all.shape
        a, b = sharding.frame_range(T, rank, world)
        pg = peer.PeerGroup(rank, world, dev) if transport == "peer" else None
        clip = pg.halo_clip(b - a, H, W) if pg else sharding.HaloClip(b - a, H, W, dev)
        k = (W // bs) // 2
        out = {}
        for rep in range(3):          # several passes: sequence numbers, acknowledgements and mailbox slots get reused
            # the first two passes run on different pixels: halos left over from them must not survive into the last
            clip.owned.copy_(y_all[a:b] if rep == 2 else torch.roll(y_all[a:b], 24, dims=2))
            r = sharding.sharded_removability(clip, T, bs, 0.5, 0.5, rank, world, transport=pg)
            mask = ops.select_rows(r, k, ops.REMOVE_HIGH)
            imp = sharding.sharded_importance(clip, bs, 0.5, 0.5, rank, world, transport=pg)
            torch.cuda.synchronize()
            out[rep] = (r.cpu().numpy(), mask.cpu().numpy(), imp.cpu().numpy())
        assert not np.array_equal(out[0][0], out[2][0])
        if pg:
            pg.check()
            pg.close()
        results[rank] = (a, b) + out[2]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport,impl", [("nccl", "umma"), ("peer", "umma"), ("nccl", "simt")])
@pytest.mark.parametrize("T,H,W,bs", [(11, 96, 160, 16), (30, 1088, 1920, 16), (9, 64, 256, 8)])
def test_sharded_equals_single_gpu(T, H, W, bs, transport, impl, monkeypatch):
    """Bit identity needs both sides to run the SAME scoring kernel, and that kernel to be independent of where a
    run of frames starts: the tcgen05 kernel computes every frame's coefficients afresh, so it is; the CUDA-core
    kernel (small / unaligned clips) carries a running fp32 sum from the start of its chunk, so its SC moves in the
    last bits with the sharding.  The tcgen05 kernel -- the one the sharded benchmark
    configurations run -- is checked bit for bit, the CUDA-core kernel within the scoring tolerance.  Spawned ranks
    inherit the environment."""
    import torch
    monkeypatch.setenv("ELVIS_SCORE_IMPL", impl)
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
    mp.spawn(_worker, args=(world, _free_port(), torch.from_numpy(y), bs, transport, results), nprocs=world, join=True)
    for rank in range(world):
        a, b, r, m, imp = results[rank]
        if impl == "umma":
            assert np.array_equal(r, ref[a:b]), (rank, "scores")
            assert np.array_equal(m, ref_mask[a:b]), (rank, "mask")
            assert np.array_equal(imp, ref_imp[a:b]), (rank, "importance")
        else:
            np.testing.assert_allclose(r, ref[a:b], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(imp, ref_imp[a:b], rtol=1e-4, atol=1e-6)
