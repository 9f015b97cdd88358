#!/bin/bash
# round 2, call D: level split tests + whole suite
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -30 gpurun_out/r2d_pytest.log
python - <<'PY' > gpurun_out/r2d_latency.log 2>&1
import torch, time, numpy as np
from stereo_depth_b200 import cuda_depth, backend
from stereo_depth_b200.synthetic import make_pair
for (H,W,D) in ((1080,1920,128),(720,1280,128),(480,640,64),(318,3840,256),(2160,3840,256)):
    l,r,_=make_pair(H,W,D,seed=3)
    lt,rt=torch.from_numpy(l).cuda(),torch.from_numpy(r).cuda()
    for split in (True, False):
        sm=cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H,width=W,downscale_factor=2,min_disparity=0,max_disparity=D-1),frames_per_launch=1)
        sm.set_level_split(split)
        for _ in range(3): sm.compute_disparity_map(lt,rt)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): sm.compute_disparity_map(lt,rt)
        e1.record(); torch.cuda.synchronize()
        print(f"{H}x{W} D={D} one frame per launch: split_on={split} S={sm.level_split(1)} variant={sm.active_variant} screen={sm.screen_active}: {e0.elapsed_time(e1)/20:.4f} ms/frame", flush=True)
PY
cat gpurun_out/r2d_latency.log
