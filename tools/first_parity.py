#!/usr/bin/env python
"""GPU bring-up check: CUDA path vs CPU oracle, stage by stage, on a few seeded configs."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from parity_util import run_cuda_all_stages, mismatch, max_abs, agg3_from_volume  # noqa: E402
from stereo_depth_b200.synthetic import make_pair  # noqa: E402

CASES = [
    dict(height=96, width=160, downscale_factor=2, min_disparity=0, max_disparity=31),
    dict(height=42, width=100, downscale_factor=1, min_disparity=0, max_disparity=23),
    dict(height=90, width=120, downscale_factor=3, min_disparity=0, max_disparity=29),
    dict(height=64, width=128, downscale_factor=2, min_disparity=8, max_disparity=39),
    dict(height=480, width=640, downscale_factor=2, min_disparity=0, max_disparity=63),
]
variants = sys.argv[1:] or ["generic"]
for kw in CASES:
    H, W, D = kw["height"], kw["width"], kw["max_disparity"] + 1
    l, r, _ = make_pair(H, W, D, seed=777)
    cfg = O.make_config(**kw)
    # the kernels' default semantics: SAFE padding, reference-compat absolute index when min_disparity/K != 0
    mode = O.MODE_COMPAT if kw["min_disparity"] // kw["downscale_factor"] else O.MODE_SAFE
    ref = O.run(cfg, l, r, mode=mode, want=O.ALL_STAGES)
    ref["agg3"] = agg3_from_volume(ref["agg"], ref["wta"], kw["min_disparity"] // kw["downscale_factor"])
    for variant in variants:
        for dtype in ("u8", "f32"):
            try:
                t = time.time()
                got = run_cuda_all_stages(l, r, kw, variant=variant, dtype=dtype, volumes=(variant != "ws"))
                dt = time.time() - t
            except Exception as e:  # noqa: BLE001
                print(kw, variant, dtype, "FAILED:", e)
                continue
            line = [f"{H}x{W} K={kw['downscale_factor']} d=[{kw['min_disparity']},{kw['max_disparity']}] {variant}/{dtype} ({dt:.2f}s):"]
            for st in ("gray_l", "gray_r", "pool_l", "pool_r", "cost", "agg", "wta", "agg3", "refined", "out"):
                if st not in got:
                    continue
                line.append(f"{st}:{mismatch(got[st], ref[st])}/{max_abs(got[st], ref[st]):.3g}")
            print(" ".join(line), flush=True)
print("device:", torch.cuda.get_device_name(0))
