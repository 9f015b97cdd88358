"""CPU check of the bound behind the certified level screen (stereo_depth_b200/csrc/mbm_screen.cu).

The screen drops a disparity level when its APPROXIMATE aggregated cost is below 0.998 x the pixel's approximate
maximum, and the fused kernel then never evaluates it.  That is only sound if (header of mbm_screen.cu)
  (1) the reference's sequential fp32 chains stay within 3.7e-4 (relative) of the real-number value, and
  (2) any fp32 evaluation of the separable dissimilarity sums does too,
so that a dropped level cannot be the reference's arg-max.  This test restates both sides with numpy -- the real value
in float64, a separable float32 evaluation in a different order than either kernel -- and checks them against the
oracle's aggregated volume (the reference's exact chains) on textured, smooth and flat scenes: the observed deviations
must be far inside the bound, and the reference's arg-max must always survive the screen's test."""
import numpy as np
import pytest

from oracle import oracle as O

KEEP = np.float32(0.998)          # kKeep in mbm_screen.cu
EPS0 = 3.7e-4                     # bound on |A_ref / A_real - 1| and |A_screen / A_real - 1| used by the kernel's analysis


def _box(a, ri, rj):
    """Circular (SAFE padding = true modulo) box sum over rows -ri..ri, columns -rj..rj, in a's dtype."""
    rows = sum(np.roll(a, -i, axis=0) for i in range(-ri, ri + 1))
    return sum(np.roll(rows, -j, axis=1) for j in range(-rj, rj + 1))


def _scene(kind, rng, H, W):
    if kind == "dots":
        left = rng.integers(0, 256, (3, H, W)).astype(np.float32)
        right = np.roll(left, -6, axis=2) + rng.integers(-2, 3, (3, H, W))
    elif kind == "smooth":
        yy, xx = np.mgrid[0:H, 0:W]
        base = 120 + 60 * np.sin(xx / 17.0) * np.cos(yy / 11.0) + rng.normal(0, 0.5, (H, W))
        left = np.stack([base, base * 0.9 + 5, base * 0.8 + 11]).astype(np.float32)
        right = np.roll(left, -3, axis=2)
    else:  # constant images: every level ties exactly
        left = np.full((3, H, W), 77.25, np.float32)
        right = left.copy()
    return np.clip(left, 0, 255).astype(np.float32), np.clip(right, 0, 255).astype(np.float32)


@pytest.mark.parametrize("kind", ["dots", "smooth", "flat"])
def test_screen_bound_holds_against_the_reference_chains(kind):
    H, W, K, D = 96, 160, 2, 32
    rng = np.random.default_rng({"dots": 1, "smooth": 2, "flat": 3}[kind])
    left, right = _scene(kind, rng, H, W)
    cfg = O.make_config(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
    ref = O.run(cfg, left, right, want=("pool_l", "pool_r", "agg", "wta"))
    pl, pr, agg = ref["pool_l"], ref["pool_r"], ref["agg"]
    Hd, Wd, L = agg.shape
    assert pl.min() >= 0 and pl.max() <= 255 and pr.min() >= 0 and pr.max() <= 255   # the screen's precondition
    worst_ref = worst_scr = 0.0
    approx = np.empty((Hd, Wd, L), np.float32)
    for d in range(L):
        a = pl - np.roll(pr, d, axis=1)                       # fl32(l - r): same number in reference and screen
        tap = (np.float32(255.0) - np.abs(a)).astype(np.float32)   # the reference's tap, fl32(255 - |a|)
        cost64 = _box(tap.astype(np.float64), 1, 1)           # real-number sums of the reference's fp32 taps
        real = _box(cost64, 1, 10) * _box(cost64, 10, 1) * _box(cost64, 4, 4)
        # the screen: dissimilarities |a| summed separably in float32, similarities N*255 - sum at the end
        dis = _box(np.abs(a).astype(np.float32), 1, 1)
        h = np.float32(63 * 9 * 255.0) - _box(dis, 1, 10)
        v = np.float32(63 * 9 * 255.0) - _box(dis, 10, 1)
        c = np.float32(81 * 9 * 255.0) - _box(dis, 4, 4)
        approx[:, :, d] = (h * v) * c
        ok = real > 0
        worst_ref = max(worst_ref, float(np.max(np.abs(agg[:, :, d][ok] / real[ok] - 1.0))))
        worst_scr = max(worst_scr, float(np.max(np.abs(approx[:, :, d][ok] / real[ok] - 1.0))))
    assert worst_ref < EPS0 / 10, worst_ref      # measured ~1e-6: the analysis is conservative by two orders of magnitude
    assert worst_scr < EPS0 / 10, worst_scr
    # the certified property: the reference's arg-max (first maximum of the exact chains) passes the screen's test
    amax = approx.max(axis=2)
    d_ref = ref["wta"].astype(np.int64)
    kept = np.take_along_axis(approx, d_ref[..., None], axis=2)[..., 0] >= KEEP * amax
    assert kept.all(), int((~kept).sum())
    # and the test is selective on textured data (the whole point) while it keeps ties on flat data
    frac = float((approx >= (KEEP * amax)[..., None]).mean())
    if kind == "dots":
        assert frac < 0.1, frac
    if kind == "flat":
        assert frac > 0.5, frac
