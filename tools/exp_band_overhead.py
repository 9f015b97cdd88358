#!/usr/bin/env python
"""Experiment: cost of the row-band machinery itself.  World size 1 (the band is the whole frame, halos wrap onto it),
peer-memory and NCCL-free paths, against the plain single-frame call; C4 and a half-height frame."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_b200 import cuda_depth  # noqa: E402
from stereo_depth_b200.bands import BandedStereoMatching  # noqa: E402
from stereo_depth_b200.synthetic import make_pair  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for H in (2160, 1080, 320):
    W, K, D = 3840, 2, 256
    left, right, _ = make_pair(H, W, D, seed=1234)
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
    l, r = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    plain = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(**kw), frames_per_launch=1)
    t_plain = timed(lambda: plain.compute_disparity_map(l, r))
    plain.profile(True)
    plain.compute_disparity_map(l, r)
    prof = {k: round(v[0], 4) for k, v in plain.profile_read_detail().items()}
    plain.profile(False)
    sm = BandedStereoMatching(cuda_depth.StereoMatchingConfiguration(**kw), p2p=True)
    t_band = timed(lambda: sm.compute(l, r))
    print(f"H={H}: plain {t_plain:.4f} ms (variant {plain.active_variant}, screen {plain.screen_active}, split {plain.level_split(1)}) {prof}  "
          f"band world=1 p2p {t_band:.4f} ms (variant {sm.handle.active_variant}, screen {sm.handle.screen_active}, split {sm.handle.level_split})", flush=True)
    sm.close()
