#!/usr/bin/env python
"""TEST / BENCH INPUT DATA -- extracts the one natural stereo pair the reference ships
(src/python/data/im0.png, im1.png, calib.txt: a Middlebury-format 1920x1080 pair, vmin=75, vmax=262) into
data/_ref/natural_pair.npz as decoded uint8 CHW arrays, exactly what the reference's MiddleBuryStereoCamera
hands to the pipeline (camera/middlebury_stereo_camera.py:57-58, torchvision.io.read_image).

data/_ref/ is git-ignored (the images are the reference's, they are never committed) but travels to the GPU box with
gpurun like the built .so files; /root/reference does not exist there.  __graft_entry__.build() runs this when the
reference tree is present.  Consumers: tests/test_zz_reference_live.py (natural-image parity against the reference's own
kernels) and bench.py's `natural` leg.  Both skip when the file is absent.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "data", "_ref", "natural_pair.npz")
REF_DIR = os.environ.get("REF_DIR", "/root/reference")


BACKEND_OUT = os.path.join(ROOT, "data", "_ref", "reference_backend_src.npz")


def pack_reference_backend(force=False):
    """The reference's 17-line backend adaptor (src/python/pipeline/depth/cuda_stereo_matching_backend.py) as bytes inside an
    npz, so the GPU box (which has no /root/reference) can import it UNMODIFIED against the `cuda_depth` shim
    (tests/test_zz_reference_live.py::test_unmodified_reference_backend_on_the_shim).  Git-ignored like everything in
    data/_ref/; never a source file in this tree."""
    src = os.path.join(REF_DIR, "src", "python", "pipeline", "depth", "cuda_stereo_matching_backend.py")
    if (os.path.exists(BACKEND_OUT) and not force) or not os.path.exists(src):
        return
    os.makedirs(os.path.dirname(BACKEND_OUT), exist_ok=True)
    np.savez_compressed(BACKEND_OUT, cuda_stereo_matching_backend=np.frombuffer(open(src, "rb").read(), dtype=np.uint8))
    print(f"wrote {BACKEND_OUT}")


def main(force=False):
    pack_reference_backend(force)
    src = os.path.join(REF_DIR, "src", "python", "data")
    if os.path.exists(OUT) and not force:
        print("data/_ref/natural_pair.npz already there")
        return 0
    if not os.path.exists(os.path.join(src, "im0.png")):
        print(f"no natural pair under {src}; skipping")
        return 0
    from torchvision.io import read_image
    left = read_image(os.path.join(src, "im0.png")).numpy()
    right = read_image(os.path.join(src, "im1.png")).numpy()
    calib = {}
    for line in open(os.path.join(src, "calib.txt")):
        if "=" in line:
            k, v = line.strip().split("=", 1)
            calib[k] = v
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, left=left, right=right, vmin=np.int32(calib["vmin"]), vmax=np.int32(calib["vmax"]),
                        focal=np.float32(calib["cam0"].strip("[]").split()[0]), baseline=np.float32(calib["baseline"]))
    print(f"wrote {OUT}: {left.shape} {left.dtype}, vmin={calib['vmin']} vmax={calib['vmax']}")
    return 0


if __name__ == "__main__":
    sys.exit(main(force="--force" in sys.argv))
