"""The CMake layer (CMakeLists.txt beside cmake/modules/IkarusMacros.cmake:6-22 of the reference): configures with the
CUDA language for sm_100a and defines the library, wrapper and test targets.  Configure only -- the two-minute CUDA build
is exercised by `cmake --build` outside the CPU suite (and by ikarus_b200/build.py, which compiles the same source)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cmake") is None, reason="cmake not installed")
def test_cmake_configures_for_sm_100a(tmp_path):
    r = subprocess.run(["cmake", "-S", ROOT, "-B", str(tmp_path), "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc",
                        "-DCMAKE_CXX_COMPILER=/usr/bin/g++"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    cache = open(os.path.join(tmp_path, "CMakeCache.txt")).read()
    assert "CMAKE_CUDA_COMPILER:" in cache
    txt = open(os.path.join(ROOT, "CMakeLists.txt")).read()
    assert "CMAKE_CUDA_ARCHITECTURES 100a" in txt and "LANGUAGES CXX CUDA" in txt
    r = subprocess.run(["cmake", "--build", str(tmp_path), "--target", "help"], capture_output=True, text=True)
    for target in ("ikb200", "test_deviceflatassembler", "test_ikarus_branch"):
        assert target in r.stdout, target
