#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE's own CUDA kernels on a B200.

Run on the GPU box (needs oracle/_ref built by oracle/build_ref.py in the CPU container):

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/.  The fixtures pin oracle/stereo_oracle.c
(tests/test_oracle_golden.py) and the sm_100a kernels (tests/test_gpu_parity.py) to the reference.

Each case stores the uint8 inputs and every intermediate the reference produces, obtained by calling
the reference launchers one by one through oracle/ref_stages.cc on tensors carved out of one
sentinel-filled pool (so the reference's out-of-bounds reads stay inside mapped memory and cells it
never writes are recognisable), plus the output of the unmodified public entry point
cuda_depth.StereoMatching.compute_disparity_map (torch_extension_module.cc:22-26).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from stereo_depth_b200.synthetic import make_pair  # noqa: E402

SENTINEL = -7777.0

# name -> (H, W, K, min_d, max_d, flavour)
CASES = {
    "g1_k2": (96, 160, 2, 0, 31, "dots"),
    "g2_k1_partial": (42, 100, 1, 0, 23, "dots"),
    "g3_k3": (90, 120, 3, 0, 29, "dots"),
    "g4_k2_flat": (64, 128, 2, 0, 31, "flat"),
    "g5_k2_mind": (64, 128, 2, 8, 39, "dots"),
    "g6_k2_big": (120, 200, 2, 0, 47, "dots"),
}


def load_ext(name):
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def inputs(H, W, D, flavour, seed):
    left, right, _ = make_pair(H, W, D, seed=seed)
    if flavour == "flat":
        # low-texture: constant patches (exact ties), a saturated patch and a smooth ramp
        rng = np.random.default_rng(seed + 99)
        for img in (left, right):
            img[:, : H // 3, : W // 3] = 128
            img[:, H // 2:, W // 2:] = 255
        ramp = (np.arange(W) * 255 // W).astype(np.uint8)
        left[:, H // 3: H // 2, :] = ramp
        right[:, H // 3: H // 2, :] = np.roll(ramp, -5)
        left[:, :8, W // 2:] = rng.integers(0, 2, (3, 8, W - W // 2), dtype=np.uint8) * 255
    return left, right


def run_reference(stages, cuda_depth, left_in, right_in, H, W, K, min_d, max_d, keep_volumes=True, api_needs_large_pool=False):
    """Every intermediate of the reference's own kernels for one pair ([3,H,W] uint8 or float32 numpy arrays):
    the launchers one by one on tensors carved out of one sentinel-filled pool, then the unmodified public entry
    point (`out_api`).  Also used live by tests/test_zz_reference_live.py at the BASELINE sizes."""
    Hd, Wd = (H + K - 1) // K, (W + K - 1) // K
    L = max_d // K - min_d // K + 1
    dev = torch.device("cuda")
    sizes = dict(gl=H * W, gr=H * W, pl=Hd * Wd, pr=Hd * Wd, cost=Hd * Wd * L, agg=Hd * Wd * L,
                 disp=Hd * Wd, out=H * W)
    guard = max(64 * Wd * L, 16 * W) + 4096
    total = sum(sizes.values()) + guard * (len(sizes) + 1)
    pool = torch.full((total,), SENTINEL, dtype=torch.float32, device=dev)
    views, off = {}, guard
    for k, n in sizes.items():
        views[k] = pool[off:off + n]
        off += n + guard
    gl, gr = views["gl"].view(H, W), views["gr"].view(H, W)
    pl, pr = views["pl"].view(Hd, Wd), views["pr"].view(Hd, Wd)
    cost, agg = views["cost"].view(Hd, Wd, L), views["agg"].view(Hd, Wd, L)
    disp, out = views["disp"].view(Hd, Wd), views["out"].view(H, W)

    left = torch.from_numpy(np.ascontiguousarray(left_in)).to(dev).float().contiguous()
    right = torch.from_numpy(np.ascontiguousarray(right_in)).to(dev).float().contiguous()
    res = {}
    stages.rgb_to_grayscale_inplace(left, gl)
    stages.rgb_to_grayscale_inplace(right, gr)
    stages.mean_pool_inplace(gl, pl, K)
    stages.mean_pool_inplace(gr, pr, K)
    stages.cost_volume(pl, pr, cost, 1, min_d // K, max_d // K)
    stages.aggregate(cost, agg, min_d // K, max_d // K, 1, 4, 10)
    stages.wta(agg, disp, min_d // K)
    torch.cuda.synchronize()
    res.update(gray_l=gl.cpu().numpy(), gray_r=gr.cpu().numpy(), pool_l=pl.cpu().numpy(),
               pool_r=pr.cpu().numpy(), wta=disp.cpu().numpy())
    if keep_volumes:
        res.update(cost=cost.cpu().numpy(), agg=agg.cpu().numpy())
    stages.secondary(gl, gr, agg, disp, 5, K)
    torch.cuda.synchronize()
    res["refined"] = disp.cpu().numpy()
    stages.upscale_vfill(gl, disp, out, K, 5)
    torch.cuda.synchronize()
    res["up"] = out.cpu().numpy()
    stages.hfill(gl, out, K, 5)
    torch.cuda.synchronize()
    res["out"] = out.cpu().numpy()

    # the unmodified public entry point, on the same inputs.  Its device_buffer tensors come from torch's caching
    # allocator: only allocations >= 1 MiB are carved from the one large cached segment below, smaller ones live in
    # 2 MiB small-pool blocks where the reference's out-of-bounds reads (up to H/3 rows past left_grayscaled) can leave
    # mapped memory.  Live tests therefore skip this leg for images whose gray planes are below 1 MiB.
    if api_needs_large_pool and H * W * 4 < (1 << 20) + 4096:
        res["out_api"] = None
        return res
    del views, gl, gr, pl, pr, cost, agg, disp, out, pool
    torch.cuda.empty_cache()
    big = torch.empty(max(256 << 20, 16 * total), dtype=torch.uint8, device=dev)
    del big  # one large cached segment: device_buffer's tensors are carved from mapped memory
    pre_guard = torch.zeros(64 << 20, dtype=torch.uint8, device=dev)  # secondary_matching reads rows BEFORE left_grayscaled
    cfg = cuda_depth.StereoMatchingConfiguration(height=H, width=W, downscale_factor=K,
                                                 min_disparity=min_d, max_disparity=max_d)
    sm = cuda_depth.StereoMatching(cfg)
    api = sm.compute_disparity_map(left, right)
    torch.cuda.synchronize()
    res["out_api"] = api.cpu().numpy().copy()
    del sm, api, pre_guard
    return res


def run_case(stages, cuda_depth, name, H, W, K, min_d, max_d, flavour, seed=4242):
    left_u8, right_u8 = inputs(H, W, max_d + 1, flavour, seed)
    res = dict(left=left_u8, right=right_u8,
               config=np.array([H, W, K, min_d, max_d, 1, 5, 5, 1, 4, 10], np.int32))
    res.update(run_reference(stages, cuda_depth, left_u8, right_u8, H, W, K, min_d, max_d))
    return res


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    stages, cuda_depth = load_ext("ref_stages"), load_ext("cuda_depth")
    for name, (H, W, K, mn, mx, flavour) in CASES.items():
        res = run_case(stages, cuda_depth, name, H, W, K, mn, mx, flavour)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **res)
        same = np.mean(res["out_api"] == res["out"])
        print(f"{name}: H={H} W={W} K={K} d=[{mn},{mx}] api==stages on {same:.4f} of cells", flush=True)


if __name__ == "__main__":
    main()
