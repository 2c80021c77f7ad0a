"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "ikb200.h")).read()
    return sorted(set(re.findall(r"^int (ikb_[a-z_0-9]+)\(", txt, flags=re.M)))


def test_library_builds_and_exports_every_declared_symbol():
    from ikarus_b200 import _capi, build

    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ikb200.h but not exported"
    assert sorted(_capi.SYMBOLS) == declared, "ctypes binding and header disagree"


def test_create_rejects_invalid_descriptions_without_gpu():
    from ikarus_b200 import _capi as capi

    lib = capi.load()
    h = ctypes.c_void_p()
    bad = capi.Desc(capi.IKB_ABI_VERSION, 4, 1, capi.STRAIN_GL, capi.MAT_NEOHOOKE, 0, 0, -1, 1.0, 1.0, 1, 24)
    assert lib.ikb_create(ctypes.byref(h), ctypes.byref(bad)) == capi.EINVAL
    # linear strain + NeoHooke is rejected statically by the reference as well
    bad = capi.Desc(capi.IKB_ABI_VERSION, 3, 1, capi.STRAIN_LINEAR, capi.MAT_NEOHOOKE, 0, 0, -1, 1.0, 1.0, 1, 24)
    assert lib.ikb_create(ctypes.byref(h), ctypes.byref(bad)) == capi.EINVAL
    # unsupported EAS variant -> Dune::NotImplemented equivalent
    bad = capi.Desc(capi.IKB_ABI_VERSION, 3, 1, capi.STRAIN_GL, capi.MAT_NEOHOOKE, 0, 11, -1, 1.0, 1.0, 1, 24)
    assert lib.ikb_create(ctypes.byref(h), ctypes.byref(bad)) == capi.ENOTIMPL
    bad = capi.Desc(99, 3, 1, capi.STRAIN_GL, capi.MAT_NEOHOOKE, 0, 0, -1, 1.0, 1.0, 1, 24)
    assert lib.ikb_create(ctypes.byref(h), ctypes.byref(bad)) == capi.EINVAL


def test_host_mirror_rejects_bad_skill_combinations():
    import ikarus_b200 as ik

    p = ik.toLamesFirstParameterAndShearModulus(emodul=100.0, nu=0.3)
    assert abs(p.lambda_ - 57.692307692307686) < 1e-12 and abs(p.mu - 38.46153846153846) < 1e-12
    with pytest.raises(TypeError):
        ik.linearElastic(ik.Materials.NeoHooke(p))
    with pytest.raises(TypeError):
        ik.nonLinearElastic(ik.Materials.LinearElasticity(p))
    basis = dict(dim=2, order=1, n_dof=8)
    X = np.zeros((1, 4, 2))
    ed = np.arange(8)[None]
    with pytest.raises(TypeError):
        ik.makeFE(basis, ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(p))), X, ed)  # 2D needs planeStrain
    dv = ik.DirichletValues(8)
    dv.fixDOFs(lambda f: f.__setitem__(slice(0, 2), True))
    assert dv.fixedDOFsize() == 2 and dv.isConstrained(1) and not dv.isConstrained(2)
