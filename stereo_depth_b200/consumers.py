"""GPU versions of what the reference does with the disparity map after the hot path:
accuracy metrics (depth_estimation_pipeline_metrics.py:18-56) and the point cloud
(depth_estimation_pipeline_hooks.py:84-92, helpers/point_cloud_helpers.py:5-13)."""
import torch

from . import _native as N


def _check(rc, what):
    if rc != N.SD_OK:
        raise RuntimeError(f"{what} failed ({rc})")


def evaluate(disparity_estimate: torch.Tensor, disparity_gt: torch.Tensor, max_disparity: float, threshold: float = 3.0):
    """One fused pass: {'D1', 'Threshold_<n>', 'MAE', 'count'} over the mask 0 < gt <= max_disparity."""
    est = disparity_estimate.contiguous().float()
    gt = disparity_gt.to(est.device).contiguous().float()
    if est.shape != gt.shape or not est.is_cuda:
        raise RuntimeError("disparity_estimate and disparity_gt must be CUDA tensors of the same shape")
    out = torch.empty(4, dtype=torch.float64, device=est.device)
    stream = torch.cuda.current_stream(est.device).cuda_stream
    with torch.cuda.device(est.device):   # the launch must happen on the tensors' device, whatever is current
        _check(N.lib().sd_metrics(est.data_ptr(), gt.data_ptr(), est.numel(), float(max_disparity), float(threshold),
                                  out.data_ptr(), stream), "sd_metrics")
    cnt, d1, th, s = out.tolist()
    if cnt == 0:
        return {"D1": float("nan"), f"Threshold_{int(threshold)}": float("nan"), "MAE": float("nan"), "count": 0}
    return {"D1": d1 / cnt, f"Threshold_{int(threshold)}": th / cnt, "MAE": s / cnt, "count": int(cnt)}


def point_cloud(disparity_map: torch.Tensor, focal_length: float, baseline: float, invalid_disparity: float = -1.0):
    """[P,3] tensor of (column, row, depth) for every valid pixel, row-major like the reference's double loop."""
    d = disparity_map.contiguous().float()
    if d.dim() != 2 or not d.is_cuda:
        raise RuntimeError("disparity_map must be a 2-D CUDA tensor")
    H, W = d.shape
    nb = (H * W + 1023) // 1024
    xyz = torch.empty((H * W, 3), dtype=torch.float32, device=d.device)
    scratch = torch.empty(nb + 1, dtype=torch.int32, device=d.device)
    stream = torch.cuda.current_stream(d.device).cuda_stream
    with torch.cuda.device(d.device):
        _check(N.lib().sd_point_cloud(d.data_ptr(), H, W, float(baseline) * float(focal_length), float(invalid_disparity),
                                      xyz.data_ptr(), scratch.data_ptr(), stream), "sd_point_cloud")
    return xyz[: int(scratch[nb].item())]
