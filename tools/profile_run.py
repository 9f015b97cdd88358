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

H, W, K, D = (2160, 3840, 2, 256) if os.environ.get("SD_SHAPE") == "C4" else (1080, 1920, 2, 128)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
l, r = make_batch(2, H, W, D)
nrep = 1 if H > 2000 else 4
l = torch.from_numpy(np.concatenate([l] * nrep)).cuda()
r = torch.from_numpy(np.concatenate([r] * nrep)).cuda()
sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H, width=W, downscale_factor=K,
                                                                      min_disparity=0, max_disparity=D - 1),
                               frames_per_launch=2 * nrep)
if os.environ.get("SD_VARIANT"):
    sm.set_variant(os.environ["SD_VARIANT"])
if os.environ.get("SD_SCREEN") is not None and sm.screen_active:
    sm.set_screen(os.environ["SD_SCREEN"] == "1")
out = None
for _ in range(reps):
    out = sm.compute_disparity_batch(l, r, out=out)
torch.cuda.synchronize()
print("done", float(out.mean()))
