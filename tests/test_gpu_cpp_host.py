"""Compiles tests/cpp/host_api_test.cpp against include/fhe_b200.hpp + libfhe_b200.so and runs it on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("degree,nmod", [(64, 3), (8192, 2)])
def test_cpp_host_api(tmp_path, degree, nmod):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = str(tmp_path / "host_api_test")
    lib_dir = os.path.join(ROOT, "fhe_rs_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"), "-o", exe,
                           "-L", lib_dir, "-lfhe_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe, str(degree), str(nmod)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
