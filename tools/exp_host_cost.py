import sys, time, torch
sys.path.insert(0, '.')
from stereo_depth_b200 import cuda_depth
from stereo_depth_b200.synthetic import make_pair
for (H,W,D) in ((1080,1920,128),(480,640,64)):
    l,r,_=make_pair(H,W,D,seed=3)
    lt,rt=torch.from_numpy(l).cuda(),torch.from_numpy(r).cuda()
    sm=cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H,width=W,min_disparity=0,max_disparity=D-1),frames_per_launch=1)
    for _ in range(5): sm.compute_disparity_map(lt,rt)
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(300): sm.compute_disparity_map(lt,rt)
    t_enq=time.perf_counter()-t0
    torch.cuda.synchronize()
    t_all=time.perf_counter()-t0
    print(f"{H}x{W}: host enqueue {t_enq/300*1e6:.1f} us/call, end-to-end {t_all/300*1e6:.1f} us/call, launches {sm.launches_per_call(1)}")
