#!/bin/bash
# round 2, call I: screened level split: whole suite + single-frame timings
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -12 gpurun_out/r2i_pytest.log
timeout 600 python tools/exp_band_overhead.py > gpurun_out/r2i_band.log 2>&1
cat gpurun_out/r2i_band.log
python - <<'PY' > gpurun_out/r2i_latency.log 2>&1
import torch, time, numpy as np
from stereo_depth_b200 import cuda_depth, backend
from stereo_depth_b200.synthetic import make_pair
for (H,W,D) in ((1080,1920,128),(720,1280,128),(480,640,64),(375,1242,128)):
    l,r,_=make_pair(H,W,D,seed=3)
    lt,rt=torch.from_numpy(l).cuda(),torch.from_numpy(r).cuda()
    K = 1 if H==375 else 2
    for split in (True, False):
        sm=cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H,width=W,downscale_factor=K,min_disparity=0,max_disparity=D-1),frames_per_launch=1)
        sm.set_level_split(split)
        for _ in range(3): sm.compute_disparity_map(lt,rt)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): sm.compute_disparity_map(lt,rt)
        e1.record(); torch.cuda.synchronize()
        print(f"{H}x{W} D={D} K={K} one frame per launch: split_on={split} S={sm.level_split(1)} variant={sm.active_variant} screen={sm.screen_active}: {e0.elapsed_time(e1)/20:.4f} ms/frame", flush=True)
PY
cat gpurun_out/r2i_latency.log
