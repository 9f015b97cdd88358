#!/usr/bin/env python
"""Row-band mode benchmark (BASELINE config C4: one 3840x2160 frame, D=256, K=2, split over the GPUs of a box).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_bands.py

Every rank owns one band of the raw frame on its GPU; a step = halo ring exchange (NCCL send/recv over NVLink) +
gray/pool + all-gather of the left gray bands + fused matching + secondary + fill on the band.  Prints one JSON line
(frames/s of the whole frame, max over ranks) and checks the gathered result against a single-GPU run on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_b200 import cuda_depth  # noqa: E402
from stereo_depth_b200.bands import BandedStereoMatching  # noqa: E402
from stereo_depth_b200.synthetic import make_pair  # noqa: E402

H, W, K, D = 2160, 3840, 2, 256
steps, warmup = 20, 3
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
left, right, _ = make_pair(H, W, D, seed=1234)
kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
P2P = os.environ.get("SD_BANDS_P2P", "1") == "1"
sm = BandedStereoMatching(cuda_depth.StereoMatchingConfiguration(**kw), p2p=P2P, variant=os.environ.get("SD_BANDS_VARIANT"))
p = sm.plan
lb = torch.from_numpy(left[:, p.x0 * K:p.x1 * K].copy()).cuda()
rb = torch.from_numpy(right[:, p.x0 * K:p.x1 * K].copy()).cuda()
for _ in range(warmup):
    out = sm.compute(lb, rb)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = sm.compute(lb, rb)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
full = sm.gather(out)
ok = None
if rank == 0:
    plain = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(**kw), frames_per_launch=1)
    l, r = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    want = plain.compute_disparity_map(l, r)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(5):
        plain.compute_disparity_map(l, r)
    s1.record()
    torch.cuda.synchronize()
    ok = bool(torch.equal(full, want))
    print(json.dumps({"metric": "frames/s (single 3840x2160 frame, D=256, K=2, row bands)", "n_gpus": world,
                      "ms_per_frame": round(float(ms.item()), 4), "value": round(1000.0 / float(ms.item()), 2),
                      "single_gpu_ms_per_frame": round(s0.elapsed_time(s1) / 5, 4),
                      "fused_kernel": sm.handle.active_variant + ("+screen" if sm.handle.screen_active else ""), "band_rows": p.band_rows, "halo_rows": p.halo_rows, "bit_identical_to_single_gpu": ok,
                      "exchange": ("peer-memory stores of 2x24 raw rows per view + left gray bands read in place over NVLink (flags, no NCCL)"
                                   if P2P else "ring send/recv of 2x24 raw rows per view + all-gather of left gray bands (NCCL)")}), flush=True)
if world > 1:
    dist.destroy_process_group()
