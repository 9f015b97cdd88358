"""TEST INFRASTRUCTURE -- ctypes wrapper around oracle/libstereo_oracle.so (stereo_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (stereo_depth_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libstereo_oracle.so")

MODE_SAFE = 0     # SAFE padding, relative index into the aggregated volume
MODE_COMPAT = 2   # SAFE padding + the reference's absolute-index read (differs only when min_disparity != 0)
MODE_REF = 3      # additionally emulates the reference's in-tensor aliased reads

CONFIG_FIELDS = ("height", "width", "downscale_factor", "min_disparity", "max_disparity",
                 "ncc_patch_radius", "sad_patch_radius", "threshold",
                 "small_mbm_radius", "mid_mbm_radius", "large_mbm_radius")
# defaults of the reference's POD (stereo_matching_configuration.hh:5-17)
CONFIG_DEFAULTS = dict(height=1080, width=1920, downscale_factor=2, min_disparity=75, max_disparity=262,
                       ncc_patch_radius=1, sad_patch_radius=5, threshold=5,
                       small_mbm_radius=1, mid_mbm_radius=4, large_mbm_radius=10)


class SoConfig(C.Structure):
    _fields_ = [(f, C.c_int32) for f in CONFIG_FIELDS]


_FP = C.POINTER(C.c_float)
_BP = C.POINTER(C.c_uint8)


class SoOutputs(C.Structure):
    _fields_ = [(n, _FP) for n in ("gray_l", "gray_r", "pool_l", "pool_r", "cost", "agg", "wta",
                                   "refined", "up", "out")] + \
               [(n, _BP) for n in ("taint_agg", "taint_refined", "taint_out")]


def build(force=False):
    src = os.path.join(HERE, "stereo_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libstereo_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.so_quad_peak.restype = C.c_float
        _lib.so_quad_peak.argtypes = [C.c_float] * 6
        _lib.so_run.restype = C.c_int
        _lib.so_num_threads.restype = C.c_int
    return _lib


def make_config(**kw):
    d = dict(CONFIG_DEFAULTS)
    for k, v in kw.items():
        if k not in d:
            raise KeyError(k)
        d[k] = int(v)
    return SoConfig(**d)


def dims(cfg):
    hd, wd, l = C.c_int32(), C.c_int32(), C.c_int32()
    lib().so_dims(C.byref(cfg), C.byref(hd), C.byref(wd), C.byref(l))
    return hd.value, wd.value, l.value


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_FP)


def _b(a):
    return a.ctypes.data_as(_BP)


def num_threads():
    return lib().so_num_threads()


def run(cfg, left, right, mode=MODE_SAFE, want=("out",)):
    """Whole pipeline on float32 (or uint8, cast like the reference's .float()) [3,H,W] images.

    `want` selects the intermediates to return; the big [Hd,Wd,L] volumes are only allocated on
    request.  Returns a dict of numpy arrays.
    """
    H, W = cfg.height, cfg.width
    Hd, Wd, L = dims(cfg)
    left, lp = _f(left)
    right, rp = _f(right)
    assert left.shape == (3, H, W) and right.shape == (3, H, W), (left.shape, (3, H, W))
    shapes = dict(gray_l=(H, W), gray_r=(H, W), pool_l=(Hd, Wd), pool_r=(Hd, Wd), cost=(Hd, Wd, L),
                  agg=(Hd, Wd, L), wta=(Hd, Wd), refined=(Hd, Wd), up=(H, W), out=(H, W))
    tshapes = dict(taint_agg=(Hd, Wd), taint_refined=(Hd, Wd), taint_out=(H, W))
    res, o = {}, SoOutputs()
    for n in want:
        if n in shapes:
            res[n] = np.empty(shapes[n], np.float32)
            setattr(o, n, res[n].ctypes.data_as(_FP))
        elif n in tshapes:
            res[n] = np.empty(tshapes[n], np.uint8)
            setattr(o, n, _b(res[n]))
        else:
            raise KeyError(n)
    rc = lib().so_run(C.byref(cfg), lp, rp, C.c_int(mode), C.byref(o))
    if rc != 0:
        raise ValueError(f"so_run failed: {rc}")
    return res


ALL_STAGES = ("gray_l", "gray_r", "pool_l", "pool_r", "cost", "agg", "wta", "refined", "up", "out",
              "taint_agg", "taint_refined", "taint_out")


# ---- single stages (unit tests) -----------------------------------------------------------

def gray(rgb):
    rgb, p = _f(rgb)
    _, H, W = rgb.shape
    out = np.empty((H, W), np.float32)
    lib().so_gray(p, C.c_int32(H), C.c_int32(W), out.ctypes.data_as(_FP))
    return out


def pool(g, K):
    g, p = _f(g)
    H, W = g.shape
    out = np.empty(((H + K - 1) // K, (W + K - 1) // K), np.float32)
    lib().so_pool(p, C.c_int32(H), C.c_int32(W), C.c_int32(K), out.ctypes.data_as(_FP))
    return out


def cost(pl, pr, L, min_d=0, radius=1):
    pl, a = _f(pl)
    pr, b = _f(pr)
    Hd, Wd = pl.shape
    out = np.empty((Hd, Wd, L), np.float32)
    lib().so_cost(a, b, C.c_int32(Hd), C.c_int32(Wd), C.c_int32(L), C.c_int32(min_d), C.c_int32(radius),
                  out.ctypes.data_as(_FP))
    return out


def aggregate(vol, rs=1, rm=4, rl=10, mode=MODE_SAFE):
    vol, p = _f(vol)
    Hd, Wd, L = vol.shape
    out = np.empty_like(vol)
    taint = np.empty((Hd, Wd), np.uint8)
    lib().so_aggregate(p, C.c_int32(Hd), C.c_int32(Wd), C.c_int32(L), C.c_int32(rs), C.c_int32(rm),
                       C.c_int32(rl), C.c_int(mode), out.ctypes.data_as(_FP), _b(taint))
    return out, taint


def wta(vol, min_d=0):
    vol, p = _f(vol)
    Hd, Wd, L = vol.shape
    out = np.empty((Hd, Wd), np.float32)
    lib().so_wta(p, C.c_int32(Hd), C.c_int32(Wd), C.c_int32(L), C.c_int32(min_d), out.ctypes.data_as(_FP))
    return out


def secondary(gl, gr, agg, disp, radius=5, K=2, min_d=0, mode=MODE_SAFE, taint_in=None):
    gl, a = _f(gl)
    gr, b = _f(gr)
    agg, c = _f(agg)
    H, W = gl.shape
    Hd, Wd, L = agg.shape
    d = np.array(disp, dtype=np.float32, copy=True, order="C")
    tin = _b(np.ascontiguousarray(taint_in, np.uint8)) if taint_in is not None else None
    tout = np.empty((Hd, Wd), np.uint8)
    lib().so_secondary(a, b, C.c_int32(H), C.c_int32(W), c, C.c_int32(Hd), C.c_int32(Wd), C.c_int32(L),
                       d.ctypes.data_as(_FP), C.c_int32(radius), C.c_int32(K), C.c_int32(min_d), C.c_int(mode),
                       tin, _b(tout))
    return d, tout


def vfill(gl, disp, K=2, threshold=5, taint_in=None):
    gl, a = _f(gl)
    disp, d = _f(disp)
    H, W = gl.shape
    Hd, Wd = disp.shape
    up = np.empty((H, W), np.float32)
    tin = _b(np.ascontiguousarray(taint_in, np.uint8)) if taint_in is not None else None
    tup = np.empty((H, W), np.uint8)
    lib().so_vfill(a, C.c_int32(H), C.c_int32(W), d, C.c_int32(Hd), C.c_int32(Wd), C.c_int32(K),
                   C.c_int32(threshold), up.ctypes.data_as(_FP), tin, _b(tup))
    return up, tup


def hfill(gl, up, K=2, threshold=5, taint_up=None):
    gl, a = _f(gl)
    up, u = _f(up)
    H, W = gl.shape
    out = np.empty((H, W), np.float32)
    tin = _b(np.ascontiguousarray(taint_up, np.uint8)) if taint_up is not None else None
    tout = np.empty((H, W), np.uint8)
    lib().so_hfill(a, C.c_int32(H), C.c_int32(W), u, C.c_int32(K), C.c_int32(threshold),
                   out.ctypes.data_as(_FP), tin, _b(tout))
    return out, tout


def quad_peak(x1, y1, x2, y2, x3, y3):
    return float(lib().so_quad_peak(*(C.c_float(float(v)) for v in (x1, y1, x2, y2, x3, y3))))
