"""The header-only C++ drop-in wrapper (include/ikarus_b200/deviceflatassembler.hh) over the C-ABI, in both of its
build modes: standalone value types, and the Ikarus/DUNE/Eigen branch compiled against the stand-in headers of
tests/cpp/stubs (none of those libraries exists in this image)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
HEADERS = [os.path.join(ROOT, "include", "ikarus_b200", f) for f in ("deviceflatassembler.hh", "hosttypes.hh")] + [
    os.path.join(ROOT, "include", "ikb200.h"), os.path.join(CPP, "flatassembler_concept.hh")]


def _build(name, extra_inc=()):
    from ikarus_b200 import build

    build.build()
    libdir = os.path.join(ROOT, "ikarus_b200")
    src, exe = os.path.join(CPP, name + ".cpp"), os.path.join(CPP, name)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(f) for f in [src] + HEADERS):
        inc = []
        for d in list(extra_inc) + [os.path.join(ROOT, "include")]:
            inc += ["-I", d]
        subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-Wall", "-Werror"] + inc + [src, "-L", libdir, "-likb200",
                                                                                          f"-Wl,-rpath,{libdir}", "-o", exe],
                       check=True)
    return exe


def test_cpp_wrapper_compiles_and_models_the_error_contract():
    exe = _build("test_deviceflatassembler")
    r = subprocess.run([exe, "compile"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "compile-mode ok" in r.stdout


def test_cpp_wrapper_ikarus_branch_compiles_against_the_stub_headers():
    """static_asserts inside: the restated Concepts::MatrixFlatAssembler, the typedefs of assembler/interface.hh:32-42,
    material detection by type, the AssemblerManipulator derivation pattern."""
    exe = _build("test_ikarus_branch", [os.path.join(CPP, "stubs")])
    r = subprocess.run([exe, "compile"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "compile-mode ok" in r.stdout


@pytest.mark.gpu
def test_cpp_wrapper_full_run_on_gpu():
    exe = _build("test_deviceflatassembler")
    r = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "run-mode ok" in r.stdout


@pytest.mark.gpu
def test_cpp_wrapper_ikarus_branch_run_on_gpu():
    exe = _build("test_ikarus_branch", [os.path.join(CPP, "stubs")])
    r = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "run-mode ok" in r.stdout
