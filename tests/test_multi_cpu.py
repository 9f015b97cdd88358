"""N>1 host logic on CPU: world_size-2 gloo processes exercise frame sharding, the ring halo exchange, the
band gather and the band algebra (halo sufficiency) -- the latter with the oracle standing in for the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereo_depth_b200.bands import HALO_POOLED, BandPlan, exchange_halos, gather_rows, shard_frames


def test_shard_frames_partitions_exactly():
    for n in (1, 7, 64, 65, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_frames(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_band_plan_geometry():
    p = BandPlan.make(2160, 3840, 2, 8, 0)
    assert (p.x0, p.x1, p.band_rows, p.halo_rows, p.local_rows) == (0, 135, 270, 24, 318)
    assert p.pooled_row_offset == -HALO_POOLED
    rows = p.global_rows_of_local_window()
    assert rows[0] == 2160 - 24 and rows[24] == 0 and rows[-1] == 270 + 23
    last = BandPlan.make(2160, 3840, 2, 8, 7)
    assert last.global_rows_of_local_window()[-1] == 23          # bottom halo wraps to the top of the image
    with pytest.raises(ValueError):
        BandPlan.make(2161, 3840, 2, 8, 0)
    with pytest.raises(ValueError):
        BandPlan.make(64, 64, 2, 8, 0)                           # bands shorter than the halo


def test_band_halo_follows_the_configuration():
    """halo = max(large_mbm_radius + ncc_patch_radius + 1, ceil((K + sad_patch_radius) / K)) pooled rows."""
    from stereo_depth_b200 import _native as N
    from stereo_depth_b200.bands import halo_pooled_rows
    c = N.default_config()
    assert halo_pooled_rows(c) == HALO_POOLED == 12
    c.large_mbm_radius, c.ncc_patch_radius = 14, 2
    assert halo_pooled_rows(c) == 17
    c = N.default_config()
    c.downscale_factor, c.sad_patch_radius, c.large_mbm_radius, c.mid_mbm_radius, c.small_mbm_radius = 1, 20, 3, 2, 1
    assert halo_pooled_rows(c) == 21          # the secondary-matching window dominates: ceil((1 + 20) / 1)
    p = BandPlan.make(2160, 3840, 2, 4, 1, halo=17)
    assert p.halo_rows == 34 and p.local_rows == 540 + 68 and p.pooled_row_offset == 270 - 17


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, H, W, K, D, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from stereo_depth_b200.synthetic import make_pair
        left, right, _ = make_pair(H, W, D, seed=21)
        plan = BandPlan.make(H, W, K, world, rank)
        r0, r1 = plan.x0 * K, plan.x1 * K
        lb = torch.from_numpy(left[:, r0:r1].copy())
        rb = torch.from_numpy(right[:, r0:r1].copy())
        lw = exchange_halos(lb, plan.halo_rows)
        rw = exchange_halos(rb, plan.halo_rows)
        rows = plan.global_rows_of_local_window()
        assert torch.equal(lw, torch.from_numpy(left[:, rows])), "halo exchange (left)"
        assert torch.equal(rw, torch.from_numpy(right[:, rows])), "halo exchange (right)"
        # band algebra: the pipeline up to the refined disparity on the local window reproduces the global rows
        kw = dict(downscale_factor=K, min_disparity=0, max_disparity=D - 1, width=W)
        glob = O.run(O.make_config(height=H, **kw), left, right, want=("wta", "refined"))
        loc = O.run(O.make_config(height=plan.local_rows, **kw), lw.numpy(), rw.numpy(), want=("wta", "refined"))
        h = plan.halo
        for st in ("wta", "refined"):
            # one extra row above (vertical fill) and below (horizontal fill) the band must be exact as well
            a = loc[st][h - 1:h + (plan.x1 - plan.x0) + 1]
            idx = [(plan.x0 - 1 + i) % (H // K) for i in range(a.shape[0])]
            assert np.array_equal(a.view(np.uint32), glob[st][idx].view(np.uint32)), st
        # gather of unequal bands
        mine = torch.from_numpy(glob["refined"][plan.x0:plan.x1].copy())
        sizes = [BandPlan.make(H, W, K, world, r).x1 - BandPlan.make(H, W, K, world, r).x0 for r in range(world)]
        full = gather_rows(mine, sizes)
        assert np.array_equal(full.numpy(), glob["refined"])
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H", [96, 100])     # 100 -> 50 pooled rows: unequal bands for world 2? (25/25) ; 96 -> 24/24
def test_world2_gloo_band_exchange_and_algebra(tmp_path, H):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), H, 128, 2, 16, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_world3_gloo_unequal_bands(tmp_path):
    world = 3   # 118/2 = 59 pooled rows -> bands 20/20/19
    mp.spawn(_worker, args=(world, _free_port(), 118, 96, 2, 16, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
