"""Evidence hardening for the CUDA path (-m gpu):
  * a hypothesis property test over shapes, downscale factors, level counts, min_disparity, input dtype and every
    schedule of the fused kernel (generic / specialised / warp-specialised / screened), >= 200 examples, bit-exact
    against the oracle;
  * an adversarial scene family for the certified level screen: the relative gap between the two best levels is swept
    through [1e-4, 1e-2], i.e. across the screen's keep threshold (1 - kKeep = 2e-3) and clear threshold
    (kClear - 1 = 5e-3) of csrc/mbm_screen.cu;
  * guard-band canaries around every scratch allocation (SD_DEBUG_GUARDS=1, sd_check_guards) after ragged shapes, tiles
    that overhang the image, images smaller than a tile and the largest level counts -- stands in for compute-sanitizer,
    which this GPU pool does not offer.
"""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import oracle as O
from parity_util import mismatch
from stereo_depth_b200.synthetic import make_pair

pytestmark = pytest.mark.gpu


def _run(kw, left, right, variant, screen=None, stages=("wta", "refined")):
    import torch
    from stereo_depth_b200 import cuda_depth
    sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(**kw))
    sm.set_variant(variant)
    if screen is not None:
        sm.set_screen(screen)
    out = sm.compute_disparity_map(torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()).cpu().numpy().copy()
    res = {s: sm.stage(s).cpu().numpy() for s in stages}
    res["out"] = out
    res["evaluated_fraction"] = sm.screen_stats()
    res["handle"] = sm
    return res


@st.composite
def stereo_case(draw):
    K = draw(st.sampled_from([1, 2, 2, 3]))
    Hd, Wd = draw(st.integers(12, 76)), draw(st.integers(24, 150))
    H, W = Hd * K - draw(st.integers(0, K - 1)), Wd * K - draw(st.integers(0, K - 1))
    L = draw(st.integers(1, 48))
    min_ds = draw(st.sampled_from([0, 0, 0, 1, 5]))
    mn = min_ds * K + draw(st.integers(0, K - 1))
    mx = (min_ds + L - 1) * K + draw(st.integers(0, K - 1))
    if mx < mn:
        mx = mn
    variant = draw(st.sampled_from(["generic", "fast", "ws", "screened", "auto"]))
    flavour = draw(st.sampled_from(["dots_u8", "dots_f32", "float_fuzz", "smooth"]))
    seed = draw(st.integers(0, 2 ** 20))
    return dict(H=H, W=W, K=K, mn=mn, mx=mx, variant=variant, flavour=flavour, seed=seed)


def _images(c):
    H, W, seed = c["H"], c["W"], c["seed"]
    rng = np.random.default_rng(seed)
    if c["flavour"] in ("dots_u8", "dots_f32"):
        l, r, _ = make_pair(H, W, max(2, c["mx"] + 1), seed=seed)
        return (l, r) if c["flavour"] == "dots_u8" else (l.astype(np.float32), r.astype(np.float32))
    if c["flavour"] == "float_fuzz":      # fractional, negative and > 255 values, an exact-tie stripe
        l = (rng.random((3, H, W)) * 330 - 40).astype(np.float32)
        r = np.roll(l, -int(rng.integers(0, 9)), axis=2).copy()
        l[:, :, W // 2: W // 2 + 6] = l[:, :, W // 2: W // 2 + 1]
        return l, r
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)   # smooth: many near-ties between levels
    base = 120 + 70 * np.sin(xx / 19.0 + seed % 7) * np.cos(yy / 13.0) + rng.normal(0, 0.5, (H, W))
    l = np.clip(np.stack([base, base * 0.9 + 7, base * 0.8 + 13]), 0, 255).astype(np.float32)
    return l, np.roll(l, -int(rng.integers(1, 6)), axis=2).copy()


# SD_HYP_EXAMPLES=N runs a longer, randomly seeded soak (e.g. 3000) instead of the fixed 220 derandomised examples
@settings(max_examples=int(os.environ.get("SD_HYP_EXAMPLES", "220")), deadline=None,
          derandomize="SD_HYP_EXAMPLES" not in os.environ, database=None,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(stereo_case())
def test_property_every_schedule_equals_the_oracle(c):
    K, mn, mx = c["K"], c["mn"], c["mx"]
    kw = dict(height=c["H"], width=c["W"], downscale_factor=K, min_disparity=mn, max_disparity=mx)
    L = mx // K - mn // K + 1
    left, right = _images(c)
    variant, screen = c["variant"], None
    if variant == "ws" and L > 150:
        variant = "fast"
    if variant == "screened":
        variant, screen = "fast", (3 <= L <= 128) or None   # (min_disparity/K != 0 behind the screen: gather pass)
    mode = O.MODE_COMPAT if mn // K else O.MODE_SAFE
    ref = O.run(O.make_config(**kw), left, right, mode=mode, want=("wta", "refined", "out"))
    got = _run(kw, left, right, variant, screen)
    for s in ("wta", "refined", "out"):
        assert mismatch(got[s], ref[s]) == 0, (s, c)


# ---- adversarial near-threshold scenes for the level screen ----------------------------------------------------------------
def _two_peak_scene(H, W, period, amp, seed):
    """Left: horizontally periodic texture (period `period` full-res columns) => disparity levels d and d + period/K
    match equally well (an exact tie).  One of the two is then handicapped by `amp` gray levels on a sparse set of
    pixels of the right image, which moves the relative gap between the two best aggregated costs through the screen's
    thresholds as `amp` is swept."""
    rng = np.random.default_rng(seed)
    cell = rng.integers(30, 226, (3, H, period)).astype(np.float32)
    left = np.tile(cell, (1, 1, W // period + 2))[:, :, :W].copy()
    shift = 6
    right = np.roll(left, -shift, axis=2).copy()
    # break the periodicity slightly, only in the right image and only every 5th column: level `shift` keeps its perfect
    # score on the other columns, the alias shift + period loses `amp` per touched tap
    mask = (np.arange(W) // period) % 2 == 0
    right[:, :, mask & (np.arange(W) % 5 == 0)] += amp
    return left, np.clip(right, 0, 255).astype(np.float32)


def test_screen_on_adversarial_top2_gaps():
    """Relative top-2 gaps swept over [1e-4, 1e-2] around 1 - kKeep = 2e-3 and kClear - 1 = 5e-3: the screened result must
    equal the unscreened one and the oracle everywhere, and the sweep must really populate both sides of both thresholds."""
    H, W, K, D = 192, 640, 2, 96
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
    cfg = O.make_config(**kw)
    gaps_seen = []
    for i, amp in enumerate([0.02, 0.05, 0.1, 0.2, 0.35, 0.6, 1.0, 1.7, 3.0, 5.0, 9.0]):
        left, right = _two_peak_scene(H, W, 32, amp, seed=500 + i)
        ref = O.run(cfg, left, right, want=("agg", "wta", "refined", "out"))
        top2 = np.sort(ref["agg"], axis=2)[:, :, -2:]
        gap = 1.0 - top2[:, :, 0] / top2[:, :, 1]
        gaps_seen.append(gap[20:-20, 40:-40].ravel())
        on = _run(kw, left, right, "fast", True)
        off = _run(kw, left, right, "fast", False)
        assert on["handle"].screen_active
        for s in ("wta", "refined", "out"):
            assert mismatch(on[s], ref[s]) == 0, (s, amp)
            assert mismatch(off[s], ref[s]) == 0, (s, amp)
        assert 0.0 < on["evaluated_fraction"] <= 1.0
    g = np.concatenate(gaps_seen)
    for lo, hi in ((1e-4, 1e-3), (1e-3, 2e-3), (2e-3, 3e-3), (3e-3, 5e-3), (5e-3, 7e-3), (7e-3, 1e-2)):
        assert ((g >= lo) & (g < hi)).sum() > 200, (lo, hi, ((g >= lo) & (g < hi)).sum())


# ---- guard-band canaries -------------------------------------------------------------------------------------------------
GUARD_SHAPES = [
    # (H, W, K, min_d, max_d): ragged sizes, tile overhang, smaller than a tile, largest level counts, K = 1 / 3
    (75, 133, 2, 0, 30), (24, 44, 2, 0, 17), (136, 264, 2, 0, 63), (40, 48, 2, 0, 63), (72, 600, 2, 0, 511),
    (60, 520, 1, 0, 129), (91, 121, 3, 0, 29), (66, 70, 2, 3, 3), (64, 128, 2, 8, 39), (130, 258, 2, 0, 255),
]


@pytest.mark.parametrize("shape", GUARD_SHAPES)
def test_no_stray_stores_around_scratch(shape, monkeypatch):
    """Every scratch allocation sits between two 64 KB guard bands; after all schedules ran on awkward shapes not one
    guard byte may have changed (sd_check_guards), and the results still equal the oracle."""
    import torch
    from stereo_depth_b200 import cuda_depth
    monkeypatch.setenv("SD_DEBUG_GUARDS", "1")
    H, W, K, mn, mx = shape
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=mn, max_disparity=mx)
    L = mx // K - mn // K + 1
    left, right, _ = make_pair(H, W, mx + 1, seed=300 + H)
    mode = O.MODE_COMPAT if mn // K else O.MODE_SAFE
    ref = O.run(O.make_config(**kw), left, right, mode=mode, want=("out",))["out"]
    lt, rt = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    batch_l, batch_r = torch.stack([lt] * 3), torch.stack([rt] * 3)
    for variant, screen in (("generic", None), ("fast", False), ("fast", True), ("ws", None), ("auto", None)):
        if variant == "ws" and L > 150:
            continue
        sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(**kw), frames_per_launch=2)
        sm.set_variant(variant)
        if screen is not None:
            try:
                sm.set_screen(screen)
            except RuntimeError:      # the screen does not support this configuration (L < 3, L > 128)
                continue
        out = sm.compute_disparity_batch(batch_l, batch_r)          # 3 frames = 2 chunks (2 + 1)
        sm.compute_disparity_batch(batch_l.float(), batch_r.float())
        assert sm._handle.check_guards() == 0, (variant, screen, shape)
        assert mismatch(out[2].cpu().numpy(), ref) == 0, (variant, screen)
        for stg in ("gray_l", "pool_r", "wta", "agg3", "refined"):
            sm.stage(stg, frame=1)
        assert sm._handle.check_guards() == 0


def test_guard_check_is_not_vacuous(monkeypatch):
    """A deliberate 16-byte store right behind a scratch plane (through sd_stage_pointer + cudaMemset) must be counted."""
    import torch
    from cuda.bindings import runtime as cudart
    from stereo_depth_b200 import cuda_depth
    monkeypatch.setenv("SD_DEBUG_GUARDS", "1")
    H, W = 64, 96
    sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H, width=W, min_disparity=0, max_disparity=15),
                                   frames_per_launch=1)
    img = torch.randint(0, 256, (3, H, W), dtype=torch.uint8, device="cuda")
    sm.compute_disparity_map(img, img)
    assert sm._handle.check_guards() == 0
    ptr = sm._handle.stage_pointer("refined", 0)
    torch.cuda.synchronize()
    (err,) = cudart.cudaMemset(ptr + (H // 2) * (W // 2) * 4, 0x11, 16)
    assert int(err) == 0
    assert sm._handle.check_guards() == 16
    (err,) = cudart.cudaMemset(ptr - 8, 0x11, 8)
    assert int(err) == 0
    assert sm._handle.check_guards() == 24
    # without SD_DEBUG_GUARDS the check is refused, not silently green
    monkeypatch.delenv("SD_DEBUG_GUARDS")
    plain = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=H, width=W, min_disparity=0, max_disparity=15))
    with pytest.raises(RuntimeError, match="guard bands are off"):
        plain._handle.check_guards()


# ---- the screen kernel's OWN sums against the bound it relies on ----------------------------------------------------------
@pytest.mark.parametrize("K,D,kind", [(2, 64, "dots"), (2, 128, "smooth"), (2, 256, "dots"), (1, 40, "stripes"), (2, 128, "natural")])
def test_screen_kernel_sums_stay_inside_the_certified_bound(K, D, kind):
    """mbm_screen_kernel's approximate aggregated costs A' (dumped through sd_set_debug_screen) against the oracle's exact
    chains: wherever the bound applies (A >= T = 2^15 * 144600 * 185910, header of mbm_screen.cu) the relative deviation
    must stay below the 3.7e-4 the analysis grants -- observed ~1e-6 -- and the reference's arg-max must pass the
    screen's own keep test A'(d*) >= 0.998 * max A' at every pixel."""
    import torch
    from stereo_depth_b200 import cuda_depth
    from stereo_depth_b200.synthetic import load_natural_pair
    H, W = 150 * K, 232 * K
    rng = np.random.default_rng(41 + D)
    if kind == "dots":
        l, r, _ = make_pair(H, W, D, seed=77)
        l, r = l.astype(np.float32), r.astype(np.float32)
    elif kind == "natural":
        pair = load_natural_pair()
        if pair is None:
            pytest.skip("data/_ref/natural_pair.npz missing")
        l = np.ascontiguousarray(pair[0][:, 300:300 + H, 500:500 + W]).astype(np.float32)
        r = np.ascontiguousarray(pair[1][:, 300:300 + H, 500:500 + W]).astype(np.float32)
    else:
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        base = (120 + 60 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + rng.normal(0, 0.7, (H, W)) if kind == "smooth"
                else 128 + 100 * np.sign(np.sin(xx * (2 * np.pi / 12.0))) + rng.normal(0, 0.05, (H, W)))
        l = np.clip(np.stack([base, base * 0.9 + 5, base * 0.8 + 11]), 0, 255).astype(np.float32)
        r = np.roll(l, -5, axis=2).copy()
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
    ref = O.run(O.make_config(**kw), l, r, want=("agg", "wta", "out"))
    sm = cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(**kw), frames_per_launch=1)
    sm.set_variant("fast")
    sm.set_screen(True)
    sm.set_level_split(False)
    vol = sm.debug_screen(True)
    out = sm.compute_disparity_map(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()).cpu().numpy()
    approx = vol.cpu().numpy().astype(np.float64)
    sm.debug_screen(False)
    assert mismatch(out, ref["out"]) == 0
    exact = ref["agg"].astype(np.float64)
    T = 2.0 ** 15 * 144600.0 * 185910.0
    covered = exact >= T
    assert covered.mean() > 0.5, covered.mean()
    rel = np.abs(approx[covered] / exact[covered] - 1.0)
    assert rel.max() <= 3.7e-4, rel.max()
    assert rel.max() <= 2e-5, rel.max()          # what is actually observed, with margin: a regression guard
    # the reference's arg-max survives the screen's own test at every pixel
    best = ref["agg"].argmax(axis=2)
    a_best = np.take_along_axis(approx, best[..., None], axis=2)[..., 0]
    assert np.all(a_best >= np.float32(0.998) * approx.max(axis=2))
    print(f"{kind} K={K} D={D}: max |A'/A - 1| = {rel.max():.2e} over {covered.mean():.3f} of the cells")
