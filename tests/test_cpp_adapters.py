"""The C++ host adapter classes (include/stitchb200.hpp) exercised by a C++ program shaped like the
reference's per-image loop (tests/cpp/test_adapters.cpp): every stage bit-compared with the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_adapters")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])
    assert os.path.exists(EXE)


def test_cpp_adapters_compile_and_refuse_to_run_without_a_device():
    import stitchingvideo_b200 as sv
    _build()
    if sv.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    r = subprocess.run([EXE, "--no-device"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_adapters_bit_exact_vs_oracle(gpu):
    _build()              # make: a no-op when the binary is newer than the headers it was compiled against
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_adapters: ok" in r.stdout
