"""Synthetic random-dot stereo pairs (the workload BASELINE.json's configs are quoted on).

The matching convention is the reference's: a left pixel at column c corresponds to the right
pixel at column c - disparity (device_functions.cuh:68), so right[:, r, c] = left[:, r, c + g[r, c]].
Deterministic per (seed, frame index); numpy only (no GPU, no oracle).
"""
import numpy as np


def ground_truth(H, W, D, rng):
    """Piecewise-constant disparity in [D/16, 3D/4] with one horizontally slanted ramp."""
    g = np.full((H, W), max(1, D // 8), np.int32)
    lo, hi = max(1, D // 16), max(2, (3 * D) // 4)
    for _ in range(6):
        h = int(rng.integers(max(2, H // 8), max(3, H // 2)))
        w = int(rng.integers(max(2, W // 8), max(3, W // 2)))
        r0 = int(rng.integers(0, max(1, H - h)))
        c0 = int(rng.integers(0, max(1, W - w)))
        g[r0:r0 + h, c0:c0 + w] = int(rng.integers(lo, hi + 1))
    # slanted region: 1 px of disparity per 16 columns -> odd disparities, sub-pixel coverage
    h, w = max(2, H // 4), max(2, W // 3)
    r0 = int(rng.integers(0, max(1, H - h)))
    c0 = int(rng.integers(0, max(1, W - w)))
    ramp = lo + (np.arange(w, dtype=np.int32) // 16) % max(1, hi - lo)
    g[r0:r0 + h, c0:c0 + w] = ramp[None, :]
    return g


def make_pair(H, W, D, seed=1234, frame=0):
    """Returns (left u8 [3,H,W], right u8 [3,H,W], gt int32 [H,W])."""
    rng = np.random.default_rng(seed + frame)
    left = rng.integers(0, 256, (3, H, W), dtype=np.uint8)
    g = ground_truth(H, W, D, rng)
    cols = np.arange(W, dtype=np.int64)[None, :] + g
    valid = cols < W
    src = np.where(valid, cols, 0)
    right = np.take_along_axis(left, np.broadcast_to(src[None], (3, H, W)), axis=2)
    fresh = rng.integers(0, 256, (3, H, W), dtype=np.uint8)
    right = np.where(valid[None], right, fresh)
    noise = rng.integers(-2, 3, (3, H, W), dtype=np.int16)
    right = np.clip(right.astype(np.int16) + noise, 0, 255).astype(np.uint8)
    return left, right, g


def make_batch(n, H, W, D, seed=1234, first_frame=0):
    ls, rs = [], []
    for f in range(n):
        l, r, _ = make_pair(H, W, D, seed, first_frame + f)
        ls.append(l)
        rs.append(r)
    return np.stack(ls), np.stack(rs)


def natural_pair_path():
    import os
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "_ref", "natural_pair.npz")


def load_natural_pair():
    """The one natural stereo pair the reference ships (src/python/data/im0.png, im1.png, calib.txt: 1920x1080,
    vmin=75, vmax=262), as extracted by oracle/make_natural.py into the git-ignored data/_ref/.  Returns
    (left u8 [3,H,W], right u8 [3,H,W], vmin, vmax) or None when the file is not there."""
    import os
    p = natural_pair_path()
    if not os.path.exists(p):
        return None
    z = np.load(p)
    return z["left"], z["right"], int(z["vmin"]), int(z["vmax"])
