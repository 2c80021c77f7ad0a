"""The header-only C++ drop-in wrapper (include/ikarus_b200/deviceflatassembler.hh) over the C-ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_deviceflatassembler")


def _build():
    from ikarus_b200 import build

    build.build()
    libdir = os.path.join(ROOT, "ikarus_b200")
    src = os.path.join(ROOT, "tests", "cpp", "test_deviceflatassembler.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "ikarus_b200", "deviceflatassembler.hh")),
            os.path.getmtime(os.path.join(ROOT, "include", "ikarus_b200", "hosttypes.hh"))):
        subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-L", libdir,
                        "-likb200", f"-Wl,-rpath,{libdir}", "-o", EXE], check=True)
    return EXE


def test_cpp_wrapper_compiles_and_models_the_error_contract():
    exe = _build()
    r = subprocess.run([exe, "compile"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "compile-mode ok" in r.stdout


@pytest.mark.gpu
def test_cpp_wrapper_full_run_on_gpu():
    exe = _build()
    r = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "run-mode ok" in r.stdout
