"""Pins the CPU oracle to the REFERENCE: tests/golden/*.npz hold every intermediate the reference's own
CUDA kernels produced on a B200 (tests/golden/make_golden.py, through oracle/_ref).  The oracle must
reproduce them bit for bit wherever the reference is defined (taint bit 0 clear); in REF mode also on
the deterministic-but-aliased cells (taint bit 1)."""
import numpy as np
import pytest

from oracle import oracle as O
from parity_util import golden_cases, load_golden, mismatch, oracle_config_from_array

CASES = golden_cases()


def test_fixtures_present():
    assert len(CASES) >= 6, "golden fixtures missing: run tests/golden/make_golden.py on a GPU box"


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", [O.MODE_REF, O.MODE_SAFE, O.MODE_COMPAT])
def test_oracle_matches_reference(name, mode):
    g = load_golden(name)
    cfg = oracle_config_from_array(O, g["config"])
    res = O.run(cfg, g["left"], g["right"], mode=mode, want=O.ALL_STAGES)
    bad = {O.MODE_REF: 1, O.MODE_COMPAT: 3, O.MODE_SAFE: 7}[mode]   # bits: 1 undefined, 2 aliased read, 4 absolute-index read (min_d != 0)
    for st in ("gray_l", "gray_r", "pool_l", "pool_r", "cost"):
        assert mismatch(res[st], g[st]) == 0, st  # defined everywhere
    ok_a = (res["taint_agg"] & bad) == 0
    ok_r = (res["taint_refined"] & bad) == 0
    ok_o = (res["taint_out"] & bad) == 0
    assert ok_a.mean() > 0.5
    L = g["agg"].shape[2]
    assert mismatch(res["agg"], g["agg"], np.repeat(ok_a[..., None], L, axis=2)) == 0
    assert mismatch(res["wta"], g["wta"], ok_a) == 0
    assert mismatch(res["refined"], g["refined"], ok_r) == 0
    assert mismatch(res["out"], g["out"], ok_o) == 0
    # the unmodified public entry point (cuda_depth.StereoMatching.compute_disparity_map)
    assert mismatch(res["out"], g["out_api"], ok_o) == 0


@pytest.mark.parametrize("name", CASES)
def test_taint_model_is_not_vacuous(name):
    """Some tainted cells really differ from the reference (the masks are needed), and the
    untainted region is the large majority for realistic sizes."""
    g = load_golden(name)
    cfg = oracle_config_from_array(O, g["config"])
    res = O.run(cfg, g["left"], g["right"], mode=O.MODE_SAFE, want=("wta", "taint_agg"))
    tainted = res["taint_agg"] != 0
    assert mismatch(res["wta"], g["wta"], tainted) > 0


def test_vertical_fill_never_written_rows_are_tainted():
    g = load_golden("g1_k2")
    # the reference leaves row 1 (K=2) of the vertical-fill output untouched: still the sentinel
    assert np.all(g["up"][1, 0::2] == -7777.0)
    cfg = oracle_config_from_array(O, g["config"])
    res = O.run(cfg, g["left"], g["right"], want=("taint_out",))
    assert np.all(res["taint_out"][1] & 1)
