"""Direct parity, on the GPU box, against the REFERENCE'S OWN CUDA KERNELS run live (oracle/_ref/cuda_depth.so and
ref_stages.so, built from /root/reference by oracle/build_ref.py; unmodified algorithm, one-token torch-2.x patch):

  * BASELINE configs C1, C2, C3 at full size: reference kernels  vs  the sm_100a path through the C ABI  vs  the CPU
    oracle -- closes the chain "CUDA == oracle (full size)" + "oracle == reference (small fixtures)" with a direct
    "CUDA == reference" and "oracle == reference" at the sizes BASELINE.json names (stereo_matching.cc:22-43);
  * the one natural stereo pair the reference ships (src/python/data/im0.png, im1.png; extracted into the git-ignored
    data/_ref/ by oracle/make_natural.py) at the calib.txt defaults (75..262: reference-compat absolute-index path),
    at 0..255 (screened fast path), and four crops -- low texture and exact ties are where the summation order and the
    "first maximum wins" rule are observable.

Cells where the reference itself reads out of bounds / uninitialised memory are excluded by the oracle's taint masks
(bit 0 undefined, bit 1 deterministic but aliased); the masks cover < 7 % of the cells at these sizes and the fraction
is asserted.  This file sorts last on purpose: the reference's out-of-bounds reads are kept inside mapped memory by a
guard allocation, but if one ever faulted it must not take the other tests' CUDA context with it.
"""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as O
from parity_util import ROOT, mismatch, run_cuda_all_stages
from stereo_depth_b200.synthetic import load_natural_pair, make_pair

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HAVE_REF = all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("cuda_depth.so", "ref_stages.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (oracle/build_ref.py needs /root/reference)")

_mods = {}


def ref_modules():
    """(ref_stages, cuda_depth of the REFERENCE, run_reference from tests/golden/make_golden.py)"""
    if not _mods:
        import torch  # noqa: F401  (libtorch symbols first)

        def load(name, path):
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
        _mods["stages"] = load("ref_stages", os.path.join(REF_DIR, "ref_stages.so"))
        _mods["cuda_depth"] = load("cuda_depth", os.path.join(REF_DIR, "cuda_depth.so"))
        _mods["golden"] = load("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    return _mods["stages"], _mods["cuda_depth"], _mods["golden"].run_reference


def three_way(left, right, H, W, K, mn, mx, min_clean, variants=("auto", "fast")):
    """reference (live)  vs  oracle  vs  CUDA path.  Returns the clean fractions for reporting."""
    import torch
    stages, ref_cuda_depth, run_reference = ref_modules()
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=mn, max_disparity=mx)
    torch.cuda.empty_cache()
    ref = run_reference(stages, ref_cuda_depth, left, right, H, W, K, mn, mx, keep_volumes=False, api_needs_large_pool=True)
    assert ref["out_api"] is not None, "test images must have gray planes >= 1 MiB (see run_reference)"
    torch.cuda.empty_cache()
    cfg = O.make_config(**kw)
    want = ("gray_l", "gray_r", "pool_l", "pool_r", "wta", "refined", "out", "taint_agg", "taint_refined", "taint_out")
    orc = O.run(cfg, left, right, mode=O.MODE_REF, want=want)
    # ---- oracle (REF mode: aliased reads emulated) == reference wherever the reference is defined ------------------
    for st in ("gray_l", "gray_r", "pool_l", "pool_r"):
        assert mismatch(orc[st], ref[st]) == 0, st
    d_a, d_r, d_o = (orc["taint_agg"] & 1) == 0, (orc["taint_refined"] & 1) == 0, (orc["taint_out"] & 1) == 0
    assert mismatch(orc["wta"], ref["wta"], d_a) == 0
    assert mismatch(orc["refined"], ref["refined"], d_r) == 0
    assert mismatch(orc["out"], ref["out"], d_o) == 0
    assert mismatch(orc["out"], ref["out_api"], d_o) == 0
    # ---- CUDA path == reference on every untainted cell (bit 2, the absolute-index read for min_disparity != 0, is
    #      reproduced by the default compat mode) -------------------------------------------------------------------
    ok_a, ok_r, ok_o = (orc["taint_agg"] & 3) == 0, (orc["taint_refined"] & 3) == 0, (orc["taint_out"] & 3) == 0
    assert ok_a.mean() >= min_clean and ok_r.mean() >= min_clean, (ok_a.mean(), ok_r.mean())
    mode = O.MODE_COMPAT if mn // K else O.MODE_SAFE
    safe = O.run(cfg, left, right, mode=mode, want=("wta", "refined", "out"))
    for variant in variants:
        got = run_cuda_all_stages(left, right, kw, variant=variant, volumes=False,
                                  dtype="u8" if left.dtype == np.uint8 else "f32")
        for st in ("gray_l", "gray_r", "pool_l", "pool_r"):
            assert mismatch(got[st], ref[st]) == 0, (variant, st)
        assert mismatch(got["wta"], ref["wta"], ok_a) == 0, variant
        assert mismatch(got["refined"], ref["refined"], ok_r) == 0, variant
        assert mismatch(got["out"], ref["out"], ok_o) == 0, variant
        assert mismatch(got["out"], ref["out_api"], ok_o) == 0, variant
        # and == the oracle's SAFE definition on EVERY cell
        for st in ("wta", "refined", "out"):
            assert mismatch(got[st], safe[st]) == 0, (variant, st)
    return ok_a.mean(), ok_o.mean()


BASELINE_CASES = {
    # name: (H, W, K, D, seed, minimum clean fraction of the WTA plane)
    "C1": (480, 640, 2, 64, 11, 0.93),
    "C2": (375, 1242, 1, 128, 77, 0.96),
    "C3": (1080, 1920, 2, 128, 1234, 0.95),
}


@needs_ref
@pytest.mark.parametrize("name", sorted(BASELINE_CASES))
def test_live_reference_baseline_sizes(name):
    H, W, K, D, seed, min_clean = BASELINE_CASES[name]
    left, right, _ = make_pair(H, W, D, seed=seed)
    clean_wta, clean_out = three_way(left, right, H, W, K, 0, D - 1, min_clean)
    print(f"{name}: clean fraction wta {clean_wta:.4f}, out {clean_out:.4f}")


def natural():
    p = load_natural_pair()
    if p is None:
        pytest.skip("data/_ref/natural_pair.npz missing (oracle/make_natural.py needs /root/reference)")
    return p


@needs_ref
def test_live_reference_natural_pair_default_calibration():
    """im0/im1 at the reference's default configuration (stereo_matching_configuration.hh:6-10 = calib.txt:
    1920x1080, K=2, 75..262): min_disparity/K != 0, so this is the reference-compat (absolute-index) path."""
    left, right, vmin, vmax = natural()
    assert (vmin, vmax) == (75, 262) and left.shape == (3, 1080, 1920)
    three_way(left, right, 1080, 1920, 2, vmin, vmax, 0.95, variants=("auto",))


@needs_ref
def test_live_reference_natural_pair_from_zero():
    """Same pair with min_disparity = 0, max 255 (L = 128): the screened fast path on natural texture."""
    left, right, _, _ = natural()
    three_way(left, right, 1080, 1920, 2, 0, 255, 0.95, variants=("auto", "fast"))


CROPS = [
    # (row0, col0, H, W, K, min_d, max_d)   crops of the natural pair: low texture, ties, K = 1, 2, 3, min_disparity != 0
    # (every crop >= 262 400 pixels: the reference's buffers must come from the allocator's large pool, see run_reference)
    (0, 0, 400, 704, 2, 0, 127),
    (700, 1200, 380, 720, 2, 0, 191),
    (300, 600, 421, 641, 1, 0, 63),
    (540, 100, 360, 801, 3, 30, 150),
]


@needs_ref
@pytest.mark.parametrize("crop", CROPS)
def test_live_reference_natural_crops(crop):
    r0, c0, H, W, K, mn, mx = crop
    left, right, _, _ = natural()
    l = np.ascontiguousarray(left[:, r0:r0 + H, c0:c0 + W])
    r = np.ascontiguousarray(right[:, r0:r0 + H, c0:c0 + W])
    three_way(l, r, H, W, K, mn, mx, 0.80, variants=("auto", "fast", "generic"))


def test_c1_vs_oracle_every_stage():
    """BASELINE config C1 (640x480, D=64, K=2): every stage incl. the two volumes, all schedules, bit-exact vs the oracle."""
    H, W, K, D = 480, 640, 2, 64
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
    left, right, _ = make_pair(H, W, D, seed=11)
    ref = O.run(O.make_config(**kw), left, right, want=O.ALL_STAGES)
    for variant, volumes, screen in (("generic", True, None), ("fast", True, None), ("fast", False, True),
                                     ("ws", False, None), ("auto", False, None)):
        got = run_cuda_all_stages(left, right, kw, variant=variant, volumes=volumes, screen=screen)
        for st in ("gray_l", "gray_r", "pool_l", "pool_r", "cost", "agg", "wta", "refined", "out"):
            if st in got:
                assert mismatch(got[st], ref[st]) == 0, (variant, st)


# ---- the reference's UNMODIFIED backend adaptor running on the shim ---------------------------------
def _import_reference_backend(tmp_path):
    """Imports src/python/pipeline/depth/cuda_stereo_matching_backend.py byte for byte (from $REF_DIR when present, else
    from the bytes packed into data/_ref/reference_backend_src.npz by oracle/make_natural.py) with `cuda_depth` resolved to
    stereo_depth_b200.cuda_depth and `pipeline.depth.StereoMatching` to the shim's ABC."""
    import sys
    import types
    from stereo_depth_b200 import backend, cuda_depth
    ref_file = os.path.join(os.environ.get("REF_DIR", "/root/reference"), "src", "python", "pipeline", "depth",
                            "cuda_stereo_matching_backend.py")
    packed = os.path.join(ROOT, "data", "_ref", "reference_backend_src.npz")
    if os.path.exists(ref_file):
        src = open(ref_file, "rb").read()
    elif os.path.exists(packed):
        src = np.load(packed)["cuda_stereo_matching_backend"].tobytes()
    else:
        pytest.skip("reference backend source not available")
    path = tmp_path / "cuda_stereo_matching_backend.py"
    path.write_bytes(src)
    saved = {k: sys.modules.get(k) for k in ("cuda_depth", "pipeline", "pipeline.depth")}
    pkg, depth = types.ModuleType("pipeline"), types.ModuleType("pipeline.depth")
    pkg.__path__, depth.__path__ = [], []
    depth.StereoMatching = backend.StereoMatching
    pkg.depth = depth
    sys.modules.update({"cuda_depth": cuda_depth, "pipeline": pkg, "pipeline.depth": depth})
    try:
        spec = importlib.util.spec_from_file_location("reference_cuda_stereo_matching_backend", str(path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def test_unmodified_reference_backend_on_the_shim(tmp_path):
    """The reference's own CudaStereoMatchingBackend class, unmodified, with `import cuda_depth` resolved to the shim:
    same constructor call, same `.cuda().float().contiguous()` + compute_disparity_map path, bit-identical output."""
    import torch
    from stereo_depth_b200 import backend, cuda_depth
    mod = _import_reference_backend(tmp_path)
    H, W, K, D = 240, 416, 2, 64
    left, right, _ = make_pair(H, W, D, seed=21)
    kw = dict(height=H, width=W, downscale_factor=K, min_disparity=0, max_disparity=D - 1)
    ref_backend = mod.CudaStereoMatchingBackend(configuration=cuda_depth.StereoMatchingConfiguration(**kw))
    assert isinstance(ref_backend, backend.StereoMatching)
    got = ref_backend.process(torch.from_numpy(left), torch.from_numpy(right))   # uint8 CPU tensors, like a Camera yields
    assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == (H, W)
    got = got.cpu().numpy().copy()
    ours = backend.CudaStereoMatchingBackend(cuda_depth.StereoMatchingConfiguration(**kw))
    mine = ours.process(torch.from_numpy(left), torch.from_numpy(right)).cpu().numpy()
    want = O.run(O.make_config(**kw), left, right)["out"]
    assert mismatch(got, want) == 0 and mismatch(mine, want) == 0
    # default-constructed, like DepthEstimationPipeline would for a 1080x1980 camera (torch_extension_module.cc:10)
    assert mod.CudaStereoMatchingBackend.__init__.__defaults__[0]._key()[:2] == (1080, 1980)


def test_runner_with_the_reference_camera_protocol():
    """run_depth_estimation_pipeline[_evaluation] driven by an object with the reference's Camera / EvaluationCamera
    protocol (camera/camera.py:7-35), config derived by extract_config_from_camera (runner.py:12-25)."""
    import torch
    from stereo_depth_b200 import pipeline as P
    H, W, D = 128, 256, 48
    frames = [make_pair(H, W, D, seed=90, frame=f) for f in range(3)]

    class Cam(P.EvaluationCamera):
        def focal_length(self):
            return 700.0

        def baseline(self):
            return 0.54

        def get_image_shape(self):
            return (H, W)

        def get_disparity_boundaries(self):
            return (0, D - 1)

        def stream_image_pairs(self):
            for l, r, _ in frames:
                yield torch.from_numpy(l), torch.from_numpy(r)

        def stream_image_pairs_with_gt_disparity(self):
            for l, r, g in frames:
                yield torch.from_numpy(l), torch.from_numpy(r), torch.from_numpy(g.astype(np.float32))

    cam = Cam()
    cfg = P.extract_config_from_camera(cam)
    assert cfg.image_shape == (H, W) and (cfg.min_disparity, cfg.max_disparity) == (0, D - 1)
    pipe = P.DepthEstimationPipeline(cfg)
    seen = []

    class Keep:
        def on_pipeline_start(self):
            seen.append("start")

        def process(self, ctx):
            seen.append((ctx.frame_index, ctx.disparity_map.clone(), ctx.config is cfg))

        def on_pipeline_end(self):
            seen.append("end")

    P.run_depth_estimation_pipeline(cam, pipe, [Keep()])
    assert seen[0] == "start" and seen[-1] == "end" and [s[0] for s in seen[1:-1]] == [0, 1, 2]
    for (idx, disp, same_cfg), (l, r, _) in zip(seen[1:-1], frames):
        want = O.run(O.make_config(height=H, width=W, downscale_factor=2, min_disparity=0, max_disparity=D - 1), l, r)["out"]
        assert same_cfg and mismatch(disp.cpu().numpy(), want) == 0
    res = P.run_depth_estimation_pipeline_evaluation(cam, pipe, [P.D1Metric(), P.MAEMetric()], verbose=False)
    assert set(res) == {"D1", "MAE"} and res["MAE"] < 6.0   # small frames: occlusions and borders are a large share
    # a camera whose shape disagrees with the pipeline is refused with the reference's message (runner.py:22-25)
    other = P.DepthEstimationPipeline(P.DepthEstimationPipelineConfig(image_shape=(H, W + 2), min_disparity=0, max_disparity=D - 1))
    with pytest.raises(RuntimeError, match="Incompatible image shapes"):
        P.run_depth_estimation_pipeline(cam, other)
    with pytest.raises(RuntimeError, match="Incompatible image shapes"):
        P.run_depth_estimation_pipeline_evaluation(cam, other)
