#!/usr/bin/env python
"""Quick device-resident timing of the CUDA path (not the contract bench; see bench.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_b200 import cuda_depth  # noqa: E402
from stereo_depth_b200.synthetic import make_batch  # noqa: E402

cases = {"REFDEFAULT": (1080, 1920, 2, 263), "C4": (2160, 3840, 2, 256), "C1": (480, 640, 2, 64), "C3": (1080, 1920, 2, 128), "C2": (375, 1242, 1, 128), "C5": (720, 1280, 2, 128)}
names = [a for a in sys.argv[1:] if a in cases] or ["C3"]
variants = [a for a in sys.argv[1:] if a in ("generic", "fast", "ws")] or ["fast"]
nf = int(os.environ.get("NF", "8"))
for name in names:
    H, W, K, D = cases[name]
    if name == "REFDEFAULT":
        # scene disparities must lie inside [75, 262]: a D=128 scene (8..96) with the right view shifted by another 75 columns
        l, r = make_batch(2, H, W, 128)
        r = np.roll(r, -75, axis=3)
    else:
        l, r = make_batch(2, H, W, D)
    l = torch.from_numpy(np.concatenate([l] * (nf // 2))).cuda()
    r = torch.from_numpy(np.concatenate([r] * (nf // 2))).cuda()
    for variant in variants:
        mind = 75 if name == "REFDEFAULT" else 0   # the reference's default calibration (vmin=75, vmax=262): compat mode
        sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(
            height=H, width=W, downscale_factor=K, min_disparity=mind, max_disparity=D - 1), frames_per_launch=nf)
        sm.set_variant(variant)
        if os.environ.get("SCREEN") is not None and variant == "fast":
            sm.set_screen(os.environ["SCREEN"] == "1")
        out = sm.compute_disparity_batch(l, r)
        torch.cuda.synchronize()
        reps = 3 if variant == "generic" else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            sm.compute_disparity_batch(l, r, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * nf)
        sm.profile(True)
        sm.compute_disparity_batch(l, r, out=out)
        prof = {k: round(v[0] / nf, 4) for k, v in sm.profile_read().items()}
        sm.profile(False)
        Hd, Wd, L = sm.dims
        ops = 237.0 * Hd * Wd * L
        print(f"screen={sm.screen_active} evaluated_fraction={sm.screen_stats():.3f}", end="  ")
        print(f"{name} {variant}: {ms:.4f} ms/frame  {1000/ms:.1f} fps  (kernel-B algorithmic {ops/1e9:.2f} Gop -> "
              f"{ops/ms/1e9:.1f} Top/s if B were everything)  per-frame kernel ms: {prof}  B: {ops/prof['cost_agg_wta']/1e9:.1f} Top/s", flush=True)
