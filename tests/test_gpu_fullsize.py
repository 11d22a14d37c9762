"""-m gpu: the hot path at BASELINE.json's full sizes (4K x 120 frames; an 8K slice), checked
through size-independent properties instead of the (slow) oracle:
  * every mask row removes exactly k blocks and they are the k best-scoring ones,
  * stretch(shrink(x)) reproduces x on kept blocks and is zero on removed ones,
  * scores are normalised to [0, 1] with both ends attained; repeated frames have TC == 0,
  * sharded scoring semantics: scoring a sub-range with a halo equals the slice of the full run,
  * pack/unpack side channels round-trip."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_dev():
    import torch
    assert torch.cuda.is_available()
    return torch, torch.device("cuda")


def _block_mask_to_pixels(torch, mask, pb):
    return mask.repeat_interleave(pb, dim=1).repeat_interleave(pb, dim=2)


@pytest.mark.parametrize("T,H,W", [(120, 2160, 3840), (12, 4320, 7680), (30, 1072, 1920)])
def test_v1_properties_full_size(torch_dev, T, H, W):
    torch, dev = torch_dev
    from elvis_b200 import ops
    from elvis_b200.pipeline import ElvisV1
    from elvis_b200.synth import synth_yuv420
    bs, amount = 16, 0.5
    clip = synth_yuv420(T, H, W, seed=99, device=dev)
    scores, mask, shrunk, full = ElvisV1(bs, amount, 0.5, 0.5).run(clip)
    by, bx = H // bs, W // bs
    k = int(amount * bx)
    assert scores.shape == (T, by, bx) and float(scores.min()) == 0.0 and float(scores.max()) == 1.0
    assert bool((mask.sum(dim=2) == k).all())
    # the removed set is the top-k of every row: no kept score exceeds a removed one
    removed_min = torch.where(mask.bool(), scores, torch.full_like(scores, 2.0)).amin(dim=2)
    kept_max = torch.where(mask.bool(), torch.full_like(scores, -1.0), scores).amax(dim=2)
    assert bool((kept_max <= removed_min).all())
    assert shrunk.y.shape == (T, H, W - k * bs) and shrunk.u.shape == (T, H // 2, (W - k * bs) // 2)
    for plane, src, pb in ((full.y, clip.y, bs), (full.u, clip.u, bs // 2), (full.v, clip.v, bs // 2)):
        for t0 in range(0, T, 20):          # chunked to bound temporary memory
            sl = slice(t0, min(T, t0 + 20))
            keep = _block_mask_to_pixels(torch, mask[sl] == 0, pb)
            assert torch.equal(plane[sl], src[sl] * keep)
    # kept blocks appear left-compacted and in order: shrink again from the stretched clip is the identity
    again, mask2 = ops.shrink(full.y, mask, bs, bx - k), mask
    assert torch.equal(again, shrunk.y)
    # side channels
    packed = ops.pack_mask_bits(mask)
    assert packed.numel() == (mask.numel() + 7) // 8 and torch.equal(ops.unpack_mask_bits(packed, mask.shape), mask)


def test_scoring_properties_full_size(torch_dev):
    torch, dev = torch_dev
    from elvis_b200 import ops
    from elvis_b200.synth import synth_yuv420
    T, H, W = 40, 2160, 3840
    y = synth_yuv420(T, H, W, seed=5, device=dev).y.clone()
    y[7] = y[6]
    y[21] = y[20]
    sc, tc, mm = ops.score_sc_tc(y, 16)
    assert bool((tc[0] == 0).all() and (tc[7] == 0).all() and (tc[21] == 0).all())
    assert bool((tc[1:7] > 0).any()) and bool((sc >= 0).all())
    assert torch.equal(sc[7], sc[6])
    assert mm.tolist() == [float(sc.min()), float(sc.max()), float(tc.min()), float(tc.max())]
    # a sub-range scored with its halo frame equals the slice of the full run (sharding invariant).
    # TC is identical (same transform of the same difference); SC is |running sum of coefficient
    # differences|, accumulated in fp32 from a different start frame: its error is ABSOLUTE (about
    # 3e-5 after 60 frames on a 0..90 scale, tools/dbg_sc_error.py), so the bound is rtol + atol
    # with atol = 1e-6 of the clip's SC range.
    sc_h, tc_h, _ = ops.score_sc_tc(y[13:29], 16, prev_halo=y[12])
    assert torch.equal(tc_h, tc[13:29])
    assert torch.allclose(sc_h, sc[13:29], rtol=1e-4, atol=1e-6 * float(sc.max()))
    # a flat clip has no AC energy
    flat = torch.full((3, 256, 512), 77, dtype=torch.uint8, device=dev)
    s0, t0, _ = ops.score_sc_tc(flat, 16)
    assert float(s0.abs().max()) < 1e-4 and float(t0.abs().max()) == 0.0


def test_v2_properties_full_size(torch_dev):
    torch, dev = torch_dev
    from elvis_b200 import ops
    from elvis_b200.pipeline import PresleyV2
    from elvis_b200.synth import synth_yuv420
    T, H, W, bs = 8, 2160, 3840, 16
    clip = synth_yuv420(T, H, W, seed=3, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    scores = torch.rand((T, H // bs, W // bs), generator=g, device=dev, dtype=torch.float64)
    v2 = PresleyV2(bs)
    zeros = torch.zeros(scores.shape, dtype=torch.int32, device=dev)
    for out in (v2.blur(clip, zeros), v2.downsample_pow2(clip, zeros, 3), v2.dampen(clip, zeros.float())):
        assert all(torch.equal(a, b) for a, b in zip(out.planes, clip.planes))        # level 0 is the identity
    rounds = ops.levels_from_scores(scores, ops.LEVELS_ROUND, 10)
    assert int(rounds.min()) >= 0 and int(rounds.max()) <= 10
    blurred = v2.blur(clip, rounds)
    untouched = _block_mask_to_pixels(torch, rounds == 0, bs)
    assert torch.equal(blurred.y[untouched], clip.y[untouched])
    # a blur never leaves the block's value range; the block mean moves by less than one grey level
    yb = blurred.y.view(T, H // bs, bs, W // bs, bs).float()
    y0 = clip.y.view(T, H // bs, bs, W // bs, bs).float()
    assert bool((yb.amax(dim=(2, 4)) <= y0.amax(dim=(2, 4))).all() and (yb.amin(dim=(2, 4)) >= y0.amin(dim=(2, 4))).all())
    levels = ops.levels_from_scores(scores, ops.LEVELS_ROUND, 4).clamp(max=3)
    down = v2.downsample_pow2(clip, levels, 3)
    yd = down.y.view(T, H // bs, bs, W // bs, bs).float()
    assert bool((yd.amax(dim=(2, 4)) <= y0.amax(dim=(2, 4))).all() and (yd.amin(dim=(2, 4)) >= y0.amin(dim=(2, 4))).all())
    packed = ops.pack_levels_2bit(levels)
    assert packed.shape == (T, H // bs, (W // bs + 3) // 4) and torch.equal(ops.unpack_levels_2bit(packed, W // bs), levels)
    # a 16x reduction (one value per block) makes every block constant, and constant blocks are fixed points
    lv4 = torch.full_like(levels, 4)
    once = v2.downsample_pow2(clip, lv4, 4)
    y1 = once.y.view(T, H // bs, bs, W // bs, bs)
    assert bool((y1.amax(dim=(2, 4)) == y1.amin(dim=(2, 4))).all())
    twice = v2.downsample_pow2(once, lv4, 4)
    assert torch.equal(once.y, twice.y) and torch.equal(v2.blur(once, rounds).y, once.y)


def test_rowcol_properties_full_size(torch_dev):
    """SURVEY 8f rank 2 at 4K (240 x 135 blocks of 16): a target met by whole passes removes exactly the
    blocks missing from the position map; stretching by position map restores every survivor and
    zeroes the rest; the recorded passes expand back to the original grid size."""
    torch, dev = torch_dev
    from elvis_b200 import ops
    bs, by, bx = 16, 135, 240
    gen = torch.Generator(device=dev).manual_seed(7)
    frame = torch.randint(1, 256, (1, by * bs, bx * bs, 3), dtype=torch.uint8, device=dev, generator=gen)
    imp = torch.rand((1, by, bx), dtype=torch.float64, device=dev, generator=gen)
    target = 20 * by + 20 * bx - 20 * 21 // 2 - 20 * 19 // 2       # 20 row passes and 20 column passes, all complete
    fby, fbx, counts = ops.rowcol_dims(by, bx, target)
    assert sum(counts) == target and (fby, fbx) == (by - 20, bx - 20) and len(counts) == 40
    mask, pos, pidx, pcnt, meta = ops.rowcol_plan(imp, target)
    assert meta[0].tolist() == [40, fby, fbx, target] and int(mask.sum()) == target
    kept = pos[0, :fby, :fbx].reshape(-1).long()
    assert kept.unique().numel() == kept.numel() == by * bx - target
    assert not bool(mask.reshape(-1)[kept].any())
    shrunk = ops.gather_blocks(frame, pos, bs, fby, fbx)
    inv = ops.invert_block_map(pos[:, :fby, :fbx].reshape(1, -1).contiguous(), by * bx).view(1, by, bx)
    assert bool(((inv >= 0) == (mask == 0)).all())
    full = ops.gather_blocks(shrunk, inv, bs, by, bx)
    keep_px = _block_mask_to_pixels(torch, mask == 0, bs)[..., None]
    assert torch.equal(full * keep_px, frame * keep_px) and not bool((full * (~keep_px)).any())
    grid = ops.rowcol_expand(pidx[:, :40].contiguous(), pcnt[:, :40].contiguous(), fby, fbx)
    assert grid.shape == (1, by, bx) and int((grid >= 0).sum()) == fby * fbx
    assert torch.equal(grid[grid >= 0].sort().values, torch.arange(fby * fbx, device=dev, dtype=torch.int32))


def test_v1_benchmark_config_vs_oracle_value_for_value(torch_dev):
    """BASELINE.json configs[1] -- 4K x 120 frames, 16x16 blocks, 50 % removal -- through the default
    path (tcgen05 scoring kernel, multi-chunk schedule) against the oracle on ALL 120 frames: SC/TC
    within rtol 1e-4 (+ 1e-6 of the SC range for flat blocks), scores within 2e-6 absolute, masks equal
    wherever the oracle's own decision margin exceeds 1e-5, shrunk and stretched planes bit for bit.
    The oracle runs on every host core through the fork pool of oracle/cpu_baseline.py (seconds)."""
    import numpy as np
    torch, dev = torch_dev
    from elvis_b200 import ops
    from elvis_b200.pipeline import ElvisV1
    from elvis_b200.synth import synth_yuv420
    from oracle import verify
    from oracle.cpu_baseline import CpuElvisV1
    T, H, W, bs, amount = 120, 2160, 3840, 16, 0.5
    clip = synth_yuv420(T, H, W, seed=1234, device=dev)
    sc, tc, _ = ops.score_sc_tc(clip.y, bs)
    scores, mask, shrunk, full = ElvisV1(bs, amount, 0.5, 0.5).run(clip)
    torch.cuda.synchronize()
    y, u, v = (p.cpu().numpy() for p in clip.planes)
    cpu = CpuElvisV1(y, u, v, bs, amount, 0.5, 0.5)
    try:
        cpu.step()
        ref = cpu.outputs()
        np.testing.assert_allclose(tc.cpu().numpy(), ref["tc"], rtol=1e-4, atol=1e-6 * float(ref["tc"].max()))
        np.testing.assert_allclose(sc.cpu().numpy(), ref["sc"], rtol=1e-4, atol=1e-6 * float(ref["sc"].max()))
        got = {"scores": scores.cpu().numpy(), "mask": mask.cpu().numpy()}
        for tag, s_, f_ in (("y", shrunk.y, full.y), ("u", shrunk.u, full.u), ("v", shrunk.v, full.v)):
            got["s" + tag], got["f" + tag] = s_.cpu().numpy(), f_.cpu().numpy()
        verdict = verify.compare_v1(got, ref, bs, int(amount * (W // bs)))
    finally:
        cpu.close()
    assert verdict["ok"], verdict
    assert verdict["frames_with_identical_mask"] >= T - 2, verdict      # near-ties are rare on this content
