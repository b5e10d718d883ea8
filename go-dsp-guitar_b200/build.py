"""Build libgdg.so (the CUDA product library) in-tree for sm_100a.

    python go-dsp-guitar_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgdg.so")
SOURCES = ["gdg_kernels.cu", "gdg_host.cpp"]
DEPS = ["gdg_stages.cuh", "gdg_stages2.cuh", "gdg_interp.cuh", "gdg_device.h", "gdg_backend.h",
        "../../include/gdg.h", "../../include/gdg_tables.h", "../../include/gdg_consts.h", "../../include/gdg_unit_params.def"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-o", OUT]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, f) for f in SOURCES]
    cmd += ["-lcudart"]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
