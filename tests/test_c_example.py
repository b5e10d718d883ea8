"""examples/offline_batch.c drives the product through include/gdg.h alone (no Python, no torch): what a cgo wrapper does.
Without a GPU it must build, link and fail loudly; on the B200 it must run."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "go-dsp-guitar_b200")


def _build(tmp_path):
    exe = str(tmp_path / "offline_batch")
    cmd = ["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "offline_batch.c"),
           "-L" + LIBDIR, "-lgdg", "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_c_example_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked run")
    r = subprocess.run([exe, "2", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "24", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"24 channels x 5 blocks of 8192 samples @ 96000 Hz on (\S+): master energy ([0-9.]+), master-left level (-?\d+) dB", r.stdout)
    assert m and m.group(1).startswith("cuda") and float(m.group(2)) > 0.0 and -60 < int(m.group(3)) <= 30, r.stdout
