"""Build tests/emu/libgdg_emu.so: gdg_host.cpp + the host-emulated device side (tests only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "go-dsp-guitar_b200", "csrc")
OUT = os.path.join(HERE, "libgdg_emu.so")


def build(force=False):
    srcs = [os.path.join(HERE, "emu_backend.cpp"), os.path.join(CSRC, "gdg_host.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
        [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-ffp-contract=off",
           "-o", OUT] + srcs + ["-lm", "-lpthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
