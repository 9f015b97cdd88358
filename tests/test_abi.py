"""The C-ABI library loads on a CPU-only box and exports every symbol include/stereo_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from stereo_depth_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "stereo_b200.h")).read()
    declared = set(re.findall(r"\b(sd_[a-z_0-9]+)\s*\(", header))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    lib = N.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sd_abi_version() == 3


def test_config_defaults_match_reference_struct():
    c = N.default_config()
    got = [getattr(c, f) for f in N.CONFIG_FIELDS]
    # stereo_matching_configuration.hh:6-16
    assert got == [1080, 1920, 2, 75, 262, 1, 5, 5, 1, 4, 10]
    assert N.dims(c) == (540, 960, 262 // 2 - 75 // 2 + 1)


@pytest.mark.parametrize("kw,dims", [
    (dict(height=480, width=640, downscale_factor=2, min_disparity=0, max_disparity=63), (240, 320, 32)),
    (dict(height=375, width=1242, downscale_factor=1, min_disparity=0, max_disparity=127), (375, 1242, 128)),
    (dict(height=1080, width=1920, downscale_factor=2, min_disparity=0, max_disparity=127), (540, 960, 64)),
    (dict(height=2160, width=3840, downscale_factor=2, min_disparity=0, max_disparity=255), (1080, 1920, 128)),
    (dict(height=721, width=1281, downscale_factor=2, min_disparity=0, max_disparity=127), (361, 641, 64)),
])
def test_dims(kw, dims):
    c = N.default_config()
    for k, v in kw.items():
        setattr(c, k, v)
    assert N.dims(c) == dims


def test_invalid_configs_are_rejected_without_a_gpu():
    lib = N.lib()
    for field, val in (("height", 0), ("downscale_factor", 0), ("max_disparity", 10), ("small_mbm_radius", 11),
                       ("min_disparity", -1)):
        c = N.default_config()
        setattr(c, field, val)
        h = C.c_void_p()
        rc = lib.sd_create(C.byref(c), 0, 0, C.byref(h))
        assert rc == N.SD_ERR_BAD_ARG, (field, rc)
        assert lib.sd_last_error(h)
        lib.sd_destroy(h)


def test_null_arguments():
    lib = N.lib()
    assert lib.sd_create(None, 0, 0, None) == N.SD_ERR_BAD_ARG
    assert lib.sd_compute(None, None, None, 0, 1, None, None) == N.SD_ERR_BAD_ARG
    assert lib.sd_destroy(None) == N.SD_OK
    assert lib.sd_last_error(None) == b"null handle"


def test_python_shim_mirrors_reference_signature():
    from stereo_depth_b200 import cuda_depth, backend
    import inspect
    sig = inspect.signature(cuda_depth.StereoMatchingConfiguration.__init__)
    names = list(sig.parameters)[1:]
    assert names == list(N.CONFIG_FIELDS)
    defaults = [p.default for p in list(sig.parameters.values())[1:]]
    assert defaults == [1080, 1980, 2, 75, 262, 1, 5, 5, 1, 4, 10]  # pybind defaults, torch_extension_module.cc:9-19
    with pytest.raises(TypeError):
        cuda_depth.StereoMatchingConfiguration(height=1.5)
    assert issubclass(backend.CudaStereoMatchingBackend, backend.StereoMatching)
    assert list(inspect.signature(backend.StereoMatching.process).parameters) == ["self", "left_image", "right_image"]


def test_no_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stereo_depth_b200 import cuda_depth
    with pytest.raises(RuntimeError):
        cuda_depth.StereoMatching(cuda_depth.StereoMatchingConfiguration(height=64, width=64))


def test_unmodified_reference_backend_imports_against_the_shim(tmp_path):
    """src/python/pipeline/depth/cuda_stereo_matching_backend.py, byte for byte, resolves `import cuda_depth` and
    `from pipeline.depth import StereoMatching` against this package (no GPU needed to import it; running it is the
    -m gpu test tests/test_zz_reference_live.py::test_unmodified_reference_backend_on_the_shim)."""
    import torch
    from test_zz_reference_live import _import_reference_backend
    from stereo_depth_b200 import backend, cuda_depth
    mod = _import_reference_backend(tmp_path)
    assert issubclass(mod.CudaStereoMatchingBackend, backend.StereoMatching)
    assert mod.cuda_depth is cuda_depth
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            mod.CudaStereoMatchingBackend()


def test_runner_protocol_helpers_without_a_gpu():
    """extract_config_from_camera / validate_pipeline_config_wrt_camera (runner.py:12-25) are host logic."""
    from stereo_depth_b200 import pipeline as P

    class Cam:
        def get_image_shape(self):
            return (375, 1242)

        def get_disparity_boundaries(self):
            return (1, 64)

    cfg = P.extract_config_from_camera(Cam())
    assert cfg == P.DepthEstimationPipelineConfig(image_shape=(375, 1242), min_disparity=1, max_disparity=64)
    P.validate_pipeline_config_wrt_camera(cfg, Cam())
    with pytest.raises(RuntimeError, match=r"Pipeline expects: \(384, 1280\) but camera provides: \(375, 1242\)"):
        P.validate_pipeline_config_wrt_camera(P.DepthEstimationPipelineConfig(), Cam())
    assert issubclass(P.EvaluationCamera, P.Camera)
    with pytest.raises(TypeError):
        P.Camera()   # abstract, like the reference's
