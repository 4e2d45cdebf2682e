"""Build libmpb200.so (the sm_100a CUDA library behind include/mpb200.h) in-tree with nvcc.

Run as a script (`python motionplanning.jl_b200/build.py`) or call build_lib() -- this is what
__graft_entry__.build() does.  The .so lands next to the sources so it travels with the tree.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmpb200.so")
SOURCES = ["api.cu", "grid_rball.cu", "collide.cu", "lq.cu", "mc.cu", "brute_rball.cu", "tc_rball.cu", "order.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the reference rounds after every operation; parity kernels additionally use explicit
    # __dmul_rn/__dadd_rn intrinsics, this flag covers everything else
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build_lib(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "mpb200.h"))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libmpb200.so cannot be built")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
