"""C++ host mirror (include/mmidx.hpp): header-only classes with the reference's names over the C ABI.
CPU: the check program compiles with -Wall -Wextra, links against libmmidx.so, reproduces RandomPermutation, raises the
reference's argument errors and fails loudly without a device.  GPU: the same program's round trip through Linear /
VladAggregator; it fails like any other parity test."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "multimedia-indexing_b200")


def _build(tmp_path):
    exe = str(tmp_path / "mirror_check")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "mirror_check.cpp"), "-L", PKG, "-lmmidx", f"-Wl,-rpath,{PKG}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_mirror_builds_links_and_reports_errors(tmp_path):
    import mmidx_b200 as M
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("the no-device error path needs a box without a GPU")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mirror_check ok" in r.stdout and "no CPU path" in r.stdout
    perm = [int(x) for x in r.stdout.splitlines()[0].split()[1:]]
    assert perm == [int(x) for x in M.random_permutation(1, 16)]  # java.util.Random + Collections.shuffle, both mirrors


@pytest.mark.gpu
def test_cpp_mirror_round_trip_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "mirror_check ok" in r.stdout, (r.stdout + r.stderr)[-1200:]
