#!/usr/bin/env python
"""Short C3 run for ncu: a few chunks of 8 frames through the device path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_b200 import cuda_depth  # noqa: E402
from stereo_depth_b200.synthetic import make_batch  # noqa: E402
import numpy as np  # noqa: E402

H, W, K, D = 1080, 1920, 2, 128
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
l, r = make_batch(2, H, W, D)
l = torch.from_numpy(np.concatenate([l] * 4)).cuda()
r = torch.from_numpy(np.concatenate([r] * 4)).cuda()
sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H, width=W, downscale_factor=K,
                                                                      min_disparity=0, max_disparity=D - 1),
                               frames_per_launch=8)
out = None
for _ in range(reps):
    out = sm.compute_disparity_batch(l, r, out=out)
torch.cuda.synchronize()
print("done", float(out.mean()))
