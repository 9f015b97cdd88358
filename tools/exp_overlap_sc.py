#!/usr/bin/env python
"""Experiment: can secondary matching (kernel C: no shared memory, 128 registers) of one batch run UNDER the level screen
+ exact kernel (kernel 1) of another batch?  Two handles, two streams, sd_compute_range."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_b200 import cuda_depth  # noqa: E402
from stereo_depth_b200 import _native as N  # noqa: E402
from stereo_depth_b200.synthetic import make_batch  # noqa: E402

H, W, K, D = 1080, 1920, 2, 128
F = int(os.environ.get("F", "8"))
l, r = make_batch(4, H, W, D)
l = torch.from_numpy(np.concatenate([l] * (F // 4))).cuda()
r = torch.from_numpy(np.concatenate([r] * (F // 4))).cuda()
cfg = cuda_depth.StereoMatchingConfiguration(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
a, b = cuda_depth.StereoMatching(cfg, frames_per_launch=F), cuda_depth.StereoMatching(cfg, frames_per_launch=F)
out_a = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
out_b = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
for sm, o in ((a, out_a), (b, out_b)):
    sm.compute_disparity_batch(l, r, out=o)      # fills every stage of both handles
torch.cuda.synchronize()
sa = torch.cuda.Stream(priority=-1)
sb = torch.cuda.Stream(priority=0)


def rng(sm, o, k0, k1, stream):
    sm._handle.compute_range(l.data_ptr(), r.data_ptr(), N.SD_U8, F, o.data_ptr(), stream.cuda_stream, k0, k1)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
        torch.cuda.current_stream().wait_stream(sa)
        torch.cuda.current_stream().wait_stream(sb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def prep():
    sa.wait_stream(torch.cuda.current_stream())
    sb.wait_stream(torch.cuda.current_stream())


def only_k1():
    prep(); rng(a, out_a, 1, 1, sa)


def only_c():
    prep(); rng(b, out_b, 2, 2, sb)


def both():
    prep(); rng(a, out_a, 1, 1, sa); rng(b, out_b, 2, 2, sb)


t1, t2, t12 = timed(only_k1), timed(only_c), timed(both)
print(f"F={F}: kernel 1 (pad + screen + exact) {t1:.4f} ms, kernel C {t2:.4f} ms, both on two streams {t12:.4f} ms "
      f"(sum {t1 + t2:.4f}; hidden {100 * (t1 + t2 - t12) / t2:.0f} % of C)")
