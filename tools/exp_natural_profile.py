#!/usr/bin/env python
"""Per-kernel device times on the reference's natural pair (screen off / on, both disparity ranges)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_b200 import cuda_depth  # noqa: E402
from stereo_depth_b200.synthetic import load_natural_pair, make_batch  # noqa: E402

left, right, vmin, vmax = load_natural_pair()
H, W = left.shape[1:]
F = 16
ls = torch.from_numpy(np.stack([np.roll(left, 8 * i, axis=2) for i in range(F)])).cuda()
rs = torch.from_numpy(np.stack([np.roll(right, 8 * i, axis=2) for i in range(F)])).cuda()
dl, dr = make_batch(2, H, W, 128)
dl = torch.from_numpy(np.concatenate([dl] * (F // 2))).cuda()
dr = torch.from_numpy(np.concatenate([dr] * (F // 2))).cuda()
out = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
for name, (a, b), mn, mx in (("dots 0..127", (dl, dr), 0, 127), ("natural 0..127", (ls, rs), 0, 127), ("natural 75..262", (ls, rs), vmin, vmax)):
    for screen in (False, True):
        sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H, width=W, min_disparity=mn, max_disparity=mx),
                                       frames_per_launch=F)
        sm.set_variant("fast")
        sm.set_screen(screen)
        for _ in range(3):
            sm.compute_disparity_batch(a, b, out=out)
        sm.profile(True)
        sm.compute_disparity_batch(a, b, out=out)
        prof = {k: round(v[0] / F * 1e3, 1) for k, v in sm.profile_read_detail().items()}
        sm.profile(False)
        print(f"{name:16s} screen={screen} paused={sm.screen_paused} evaluated={sm.screen_stats():.3f} us/frame {prof} total {sum(prof.values()):.1f}", flush=True)
