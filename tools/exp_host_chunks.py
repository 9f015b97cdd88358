#!/usr/bin/env python
"""Experiment: chunk schedule of the pipelined host path (sd_compute_host): SD_HOST_CHUNK x SD_HOST_EDGE."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, time, torch
sys.path.insert(0, %r)
from stereo_depth_b200 import backend, cuda_depth
from stereo_depth_b200.synthetic import make_batch
import numpy as np
H, W, D, F = 1080, 1920, 128, 64
l, r = make_batch(8, H, W, D)
lh = torch.from_numpy(np.concatenate([l] * 8)).pin_memory(); rh = torch.from_numpy(np.concatenate([r] * 8)).pin_memory()
out = torch.empty((F, H, W), dtype=torch.float32).pin_memory()
be = backend.CudaStereoMatchingBackend(cuda_depth.StereoMatchingConfiguration(height=H, width=W, min_disparity=0, max_disparity=D - 1))
for _ in range(3): be.process_batch(lh, rh, out=out)
best = 1e9
for rep in range(3):
    t = time.perf_counter()
    for _ in range(10): be.process_batch(lh, rh, out=out)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t)
print("%%.1f" %% (F * 10 / best))
''' % ROOT
for hc in (3, 4, 6, 8, 12):
    for edge in (1, 2):
        env = dict(os.environ, SD_HOST_CHUNK=str(hc), SD_HOST_EDGE=str(edge))
        out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
        print(f"SD_HOST_CHUNK={hc} SD_HOST_EDGE={edge}: {out.stdout.strip() or out.stderr[-300:]} frames/s end to end", flush=True)
