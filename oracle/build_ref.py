#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- builds the *reference's own* CUDA stereo matcher into oracle/_ref/.

Recipe:
  1. copy  $REF_DIR/src/csrc/{depth,imageops}  to a temporary directory (never into this repo),
  2. apply the one-token patch the reference needs to compile against torch >= 2.x
     (`AT_DISPATCH_FLOATING_TYPES(x.type(), ...)` -> `x.scalar_type()`; 8 kernel files),
  3. compile the 8 kernel files for sm_100 with nvcc's default flags (what BuildExtension would
     use: -O3 on device code, FMA contraction on) and the 3 host files with g++,
  4. link  oracle/_ref/cuda_depth.so  (the reference's pybind module, unmodified API:
     src/csrc/depth/torch_extension_module.cc:6-27) and oracle/_ref/ref_stages.so (the same
     objects behind oracle/ref_stages.cc, one entry point per reference launcher).

Outputs go only into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
The reference cannot *run* in the CPU container (no GPU); it runs on the GPU box, where it is
the second oracle and the "reference CUDA kernels on one B200" timing arm.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_DIR = os.environ.get("REF_DIR", "/root/reference")

KERNELS = [
    "imageops/kernels/rgb_to_grayscale.cu",
    "imageops/kernels/mean_pool.cu",
    "depth/kernels/ncc_matching_cost_volume_construction.cu",
    "depth/kernels/multi_block_matching_cost_aggregation.cu",
    "depth/kernels/wta_disparity_selection.cu",
    "depth/kernels/secondary_matching.cu",
    "depth/kernels/upscale_disparity_vertical_fill.cu",
    "depth/kernels/horizontal_disparity_fill.cu",
]
HOST = [
    "depth/torch_extension_module.cc",
    "depth/stereo_matching.cc",
    "depth/buffer/device_buffer.cc",
]


def up_to_date():
    return all(os.path.exists(os.path.join(OUT, f)) for f in ("cuda_depth.so", "ref_stages.so"))


def main(force=False):
    if up_to_date() and not force:
        print("oracle/_ref already built")
        return 0
    if not os.path.isdir(os.path.join(REF_DIR, "src", "csrc", "depth")):
        print(f"reference tree not found at {REF_DIR}; keeping whatever is in oracle/_ref")
        return 0 if up_to_date() else 1
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="refbuild_")
    for sub in ("depth", "imageops"):
        shutil.copytree(os.path.join(REF_DIR, "src", "csrc", sub), os.path.join(tmp, sub))
    for k in KERNELS:
        p = os.path.join(tmp, k)
        src = open(p).read().replace(".type()", ".scalar_type()")
        open(p, "w").write(src)

    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", f"-I{tmp}"]
    common = ["-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
    nv = ["nvcc", "-c", "-std=c++17", "-O3", "-gencode", "arch=compute_100,code=sm_100",
          "--expt-relaxed-constexpr", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
          "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
          "-Xcompiler", "-fPIC", "-w"] + common + inc
    gx = ["g++", "-c", "-std=c++17", "-O2", "-fPIC", "-w", "-I/usr/local/cuda/include"] + common + inc

    jobs = []
    for k in KERNELS:
        o = os.path.join(tmp, k.replace("/", "_") + ".o")
        # the kernel objects are shared by both modules, so they must not bake in a module name
        jobs.append((nv + ["-DTORCH_EXTENSION_NAME=cuda_depth", os.path.join(tmp, k), "-o", o], o, "kern"))
    for h in HOST:
        o = os.path.join(tmp, h.replace("/", "_") + ".o")
        jobs.append((gx + ["-DTORCH_EXTENSION_NAME=cuda_depth", os.path.join(tmp, h), "-o", o], o, "depth"))
    o = os.path.join(tmp, "ref_stages.o")
    jobs.append((gx + ["-DTORCH_EXTENSION_NAME=ref_stages", os.path.join(HERE, "ref_stages.cc"), "-o", o], o, "stages"))

    def run(job):
        cmd, obj, _ = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("compile failed: " + " ".join(cmd[-3:]))
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))

    libdir = ce.library_paths()[0]
    link = ["-shared", f"-L{libdir}", "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda",
            "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{libdir}"]
    kern = [j[1] for j in jobs if j[2] == "kern"]
    depth = [j[1] for j in jobs if j[2] == "depth"]
    stages = [j[1] for j in jobs if j[2] == "stages"]
    subprocess.check_call(["g++", "-o", os.path.join(OUT, "cuda_depth.so")] + kern + depth + link)
    subprocess.check_call(["g++", "-o", os.path.join(OUT, "ref_stages.so")] + kern + stages + link)
    # SASS of the reference build: the FMA-contraction rules the oracle follows are read from it
    with open(os.path.join(OUT, "reference_sass.txt"), "w") as f:
        subprocess.run(["cuobjdump", "-sass", os.path.join(OUT, "cuda_depth.so")], stdout=f, stderr=subprocess.DEVNULL)
    shutil.rmtree(tmp, ignore_errors=True)
    print("built", os.listdir(OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main(force="--force" in sys.argv))
