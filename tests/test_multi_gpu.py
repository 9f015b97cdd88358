"""GPU tests of the multi-GPU modes.  The single-GPU ones run everywhere; the NCCL ones need >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _plain(left, right, kw):
    import torch
    from stereo_depth_b200 import cuda_depth
    sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(**kw))
    return sm.compute_disparity_map(torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()).cpu().numpy().copy()


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("K", [1, 2])
def test_banded_world1_equals_plain(K, p2p):
    """World size 1: the band is the whole image and both halos wrap onto it -- must equal the normal path on
    EVERY cell (SAFE padding is truly circular), which checks sd_set_band / sd_compute_range and the fill rules."""
    import torch
    from stereo_depth_b200 import cuda_depth
    from stereo_depth_b200.bands import BandedStereoMatching
    from stereo_depth_b200.synthetic import make_pair
    H, W, D = 120, 256, 32
    left, right, _ = make_pair(H, W, D, seed=31)
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
    want = _plain(left, right, kw)
    sm = BandedStereoMatching(cuda_depth.StereoMatchingConfiguration(**kw), p2p=p2p)
    for _ in range(3):   # several frames: the flags are epochs, the published gray buffer alternates
        got = sm.compute(torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("p2p", [False, True])
def test_banded_world1_nondefault_radii(p2p):
    """The band halo is derived from the configuration (bands.halo_pooled_rows): radii larger than the defaults need
    more than 12 pooled halo rows.  large_mbm_radius 12 + cost radius 2 -> 15 rows; must equal the normal path."""
    import torch
    from stereo_depth_b200 import cuda_depth
    from stereo_depth_b200.bands import BandedStereoMatching, halo_pooled_rows
    from stereo_depth_b200.synthetic import make_pair
    H, W, D, K = 144, 256, 32, 2
    left, right, _ = make_pair(H, W, D, seed=37)
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1, ncc_patch_radius=2,
              sad_patch_radius=7, small_mbm_radius=2, mid_mbm_radius=5, large_mbm_radius=12)
    cfg = cuda_depth.StereoMatchingConfiguration(**kw)
    assert halo_pooled_rows(cfg._as_struct()) == 15
    want = _plain(left, right, kw)
    sm = BandedStereoMatching(cfg, p2p=p2p)
    assert sm.plan.halo == 15
    got = sm.compute(torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    sm.close()
    # a halo that is too small for the configuration is refused by the library
    from stereo_depth_b200 import _native as N
    local = N.SdConfig(*[getattr(cfg._as_struct(), f) for f in N.CONFIG_FIELDS])
    local.height = H + 2 * 24
    h = N.Handle(local, torch.cuda.current_device(), 1)
    with pytest.raises(RuntimeError, match="halo too small"):
        h.band_p2p_init(1, 0, [0, H], 24, N.SD_U8)
    h.close()


def _band_worker(rank, world, port, H, W, K, D, outdir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from stereo_depth_b200 import cuda_depth
        from stereo_depth_b200.bands import BandedStereoMatching, shard_frames
        from stereo_depth_b200.synthetic import make_batch, make_pair
        left, right, _ = make_pair(H, W, D, seed=33)
        kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
        sm = BandedStereoMatching(cuda_depth.StereoMatchingConfiguration(**kw))
        p = sm.plan
        lb = torch.from_numpy(left[:, p.x0 * K:p.x1 * K].copy()).cuda()
        rb = torch.from_numpy(right[:, p.x0 * K:p.x1 * K].copy()).cuda()
        band = sm.compute(lb, rb)
        full = sm.gather(band)
        torch.cuda.synchronize()
        np.save(os.path.join(outdir, f"band{rank}.npy"), full.cpu().numpy())
        # the same bands over peer memory (P2P stores + flags instead of NCCL), several frames back to back, the last
        # one with different content so stale halos / gray rows would show
        sp = BandedStereoMatching(cuda_depth.StereoMatchingConfiguration(**kw), p2p=True)
        left2, right2, _ = make_pair(H, W, D, seed=34)
        for _ in range(3):
            sp.compute(lb, rb)
        lb2 = torch.from_numpy(left2[:, p.x0 * K:p.x1 * K].copy()).cuda()
        rb2 = torch.from_numpy(right2[:, p.x0 * K:p.x1 * K].copy()).cuda()
        full2 = sp.gather(sp.compute(lb2, rb2))
        full1 = sp.gather(sp.compute(lb, rb))
        torch.cuda.synchronize()
        np.save(os.path.join(outdir, f"p2p{rank}.npy"), full1.cpu().numpy())
        np.save(os.path.join(outdir, f"p2p_b{rank}.npy"), full2.cpu().numpy())
        # uneven bands (pooled rows not divisible by the world size), float32 input, both transports
        Hu = H + 2 * K
        leftu, rightu, _ = make_pair(Hu, W, D, seed=35)
        kwu = dict(kw, height=Hu)
        for name, p2p in (("nccl", False), ("p2p", True)):
            su = BandedStereoMatching(cuda_depth.StereoMatchingConfiguration(**kwu), p2p=p2p)
            pu = su.plan
            lbu = torch.from_numpy(leftu[:, pu.x0 * K:pu.x1 * K].copy()).float().cuda()
            rbu = torch.from_numpy(rightu[:, pu.x0 * K:pu.x1 * K].copy()).float().cuda()
            fullu = su.gather(su.compute(lbu, rbu))
            torch.cuda.synchronize()
            np.save(os.path.join(outdir, f"uneven_{name}{rank}.npy"), fullu.cpu().numpy())
        # frame sharding: this rank's contiguous chunk of an 6-frame batch
        n = 6
        L, R = make_batch(n, 96, 160, 32, seed=5)
        a, b = shard_frames(n, world, rank)
        fm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=96, width=160, min_disparity=0,
                                                                              max_disparity=31))
        o = fm.compute_disparity_batch(torch.from_numpy(L[a:b]).cuda(), torch.from_numpy(R[a:b]).cuda())
        np.save(os.path.join(outdir, f"shard{rank}.npy"), o.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_row_bands_and_frame_shards_over_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    from stereo_depth_b200.synthetic import make_batch, make_pair
    H, W, K, D = 288, 512, 2, 64
    mp.spawn(_band_worker, args=(world, _free_port(), H, W, K, D, str(tmp_path)), nprocs=world, join=True)
    left, right, _ = make_pair(H, W, D, seed=33)
    want = _plain(left, right, dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1))
    left2, right2, _ = make_pair(H, W, D, seed=34)
    want2 = _plain(left2, right2, dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1))
    for r in range(world):
        got = np.load(tmp_path / f"band{r}.npy")
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"rank {r}"
        assert np.array_equal(np.load(tmp_path / f"p2p{r}.npy").view(np.uint32), want.view(np.uint32)), f"p2p rank {r}"
        assert np.array_equal(np.load(tmp_path / f"p2p_b{r}.npy").view(np.uint32), want2.view(np.uint32)), f"p2p (2nd scene) rank {r}"
    Hu = H + 2 * K
    leftu, rightu, _ = make_pair(Hu, W, D, seed=35)
    wantu = _plain(leftu, rightu, dict(height=Hu, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1))
    for r in range(world):
        for name in ("nccl", "p2p"):
            got = np.load(tmp_path / f"uneven_{name}{r}.npy")
            assert np.array_equal(got.view(np.uint32), wantu.view(np.uint32)), f"uneven {name} rank {r}"
    L, R = make_batch(6, 96, 160, 32, seed=5)
    import torch
    from stereo_depth_b200 import cuda_depth
    fm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=96, width=160, min_disparity=0,
                                                                          max_disparity=31))
    single = fm.compute_disparity_batch(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()).cpu().numpy()
    sharded = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])
    assert np.array_equal(sharded.view(np.uint32), single.view(np.uint32))
