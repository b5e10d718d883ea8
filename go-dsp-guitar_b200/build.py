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
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".cu", ".cpp"))) + \
    sorted(os.path.join("..", "..", "include", f) for f in os.listdir(INCLUDE))        # every source and header: nothing stale ships
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
    cmd += os.environ.get("GDG_NVCC_EXTRA", "").split()           # development: -D overrides of the launch-shape constants
    cmd += [os.path.join(CSRC, f) for f in SOURCES]
    cmd += ["-lcudart"]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
