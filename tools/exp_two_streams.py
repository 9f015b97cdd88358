#!/usr/bin/env python
"""Experiment: two handles on two streams, each half of the batch, vs one handle (kernel overlap across chunks)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_b200 import cuda_depth  # noqa: E402
from stereo_depth_b200.synthetic import make_batch  # noqa: E402

H, W, K, D = 1080, 1920, 2, 128
F = 60
l, r = make_batch(4, H, W, D)
l = torch.from_numpy(np.concatenate([l] * (F // 4))).cuda()
r = torch.from_numpy(np.concatenate([r] * (F // 4))).cuda()
cfg = cuda_depth.StereoMatchingConfiguration(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
fpl = int(os.environ.get("FPL", "15"))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


one = cuda_depth.StereoMatching(cfg, frames_per_launch=fpl)
out = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
t1 = timed(lambda: one.compute_disparity_batch(l, r, out=out))
print(f"one handle: {t1 / F * 1000:.1f} us/frame  {F / t1 * 1000:.0f} fps")

for nh in (2, 3):
    hs = [cuda_depth.StereoMatching(cfg, frames_per_launch=fpl) for _ in range(nh)]
    ss = [torch.cuda.Stream() for _ in range(nh)]
    chunks = [(i * fpl, min(F, (i + 1) * fpl)) for i in range((F + fpl - 1) // fpl)]

    def run():
        cur = torch.cuda.current_stream()
        for s in ss:
            s.wait_stream(cur)
        for i, (a, b) in enumerate(chunks):
            with torch.cuda.stream(ss[i % nh]):
                hs[i % nh].compute_disparity_batch(l[a:b], r[a:b], out=out[a:b])
        for s in ss:
            cur.wait_stream(s)

    t2 = timed(run)
    print(f"{nh} handles / streams, chunks round-robin: {t2 / F * 1000:.1f} us/frame  {F / t2 * 1000:.0f} fps")
