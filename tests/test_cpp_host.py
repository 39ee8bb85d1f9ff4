"""The C++ host mirror (include/fastace_b200.hpp: fastace::BatchedEconomy + the batched decision-maker plugin
interfaces, same names as the reference's Economy / PersonDecisionMaker / FirmDecisionMaker) driven by a C++ test
program that reads like a test of the reference's own classes (tests/cpp/test_host_mirror.cpp)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def exe(native_lib, oracle):
    import __graft_entry__ as entry
    return entry.build_cpp_tests()


def test_builds_links_and_fails_loudly_without_a_gpu(exe):
    import torch
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libfastace_b200.so" in ldd and "liboracle.so" in ldd and "not found" not in ldd
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    res = subprocess.run([exe, "2", "1"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2 and "init failed" in res.stdout      # no CPU path: init() returns nullptr


@pytest.mark.gpu
def test_cpp_host_mirror_matches_oracle(exe):
    res = subprocess.run([exe, "24", "8"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().startswith("OK")
