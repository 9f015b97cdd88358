#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel variant on shapes with overhanging tiles, odd L, K=1/2/3."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_b200 import cuda_depth  # noqa: E402
from stereo_depth_b200.synthetic import make_batch  # noqa: E402

CASES = [(75, 133, 2, 0, 30), (42, 100, 1, 0, 23), (90, 120, 3, 0, 29), (64, 128, 2, 8, 39), (136, 264, 2, 0, 63)]
for (H, W, K, mn, mx) in CASES:
    l, r = make_batch(3, H, W, mx + 1, seed=3)
    for variant in ("generic", "fast", "ws"):
        sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H, width=W, downscale_factor=K,
                                                                              min_disparity=mn, max_disparity=mx),
                                       frames_per_launch=2)
        sm.set_variant(variant)
        out = sm.compute_disparity_batch(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda())
        host = sm.compute_disparity_host(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory())
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), host), (H, W, variant)
        print(H, W, K, variant, float(out.mean()), flush=True)
print("sanitizer workload done")
