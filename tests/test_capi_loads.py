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


def test_host_load_sampling_volume_and_traction():
    """Host-side sampling of the reference's std::function loads (loads/volume.hh:86-106, loads/traction.hh:107-138):
    totals and nodal distribution for constant loads on a distorted-free box (no GPU needed)."""
    import ikarus_b200 as ik
    from ikarus_b200 import meshes

    cells, bbox = (3, 2, 2), (3.0, 1.0, 2.0)
    slab = meshes.structured_q1(cells, bbox)
    p = ik.toLamesFirstParameterAndShearModulus(emodul=100.0, nu=0.3)
    q = np.array([0.5, -1.0, 2.0])
    t = np.array([0.0, 3.0, -1.0])
    # faces on x = 3 (face id 1 of the elements with i = cells[0]-1)
    faces = [(e, 1) for e in range(slab.corner_coords.shape[0]) if abs(slab.corner_coords[e][:, 0].max() - 3.0) < 1e-12]
    assert len(faces) == 4
    sk = ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(p)), ik.volumeLoad(lambda x, lam: lam * q),
                   ik.neumannBoundaryLoad(faces, lambda x, lam: lam * t))
    fes = ik.makeFE(dict(dim=3, order=1, n_dof=slab.n_dof), sk, slab.corner_coords, slab.elem_dofs)
    f = fes.sample_external_load(2.0).reshape(-1, 3)
    total = f.sum(0)
    assert np.allclose(total, 2.0 * (q * 6.0 + t * 2.0), atol=1e-12)  # volume 6, face area 2
    # corner node (3,1,2): an eighth of one element volume and a quarter of one loaded face
    X = meshes.node_coords(cells, bbox)
    c = np.nonzero(np.all(np.abs(X - np.array(bbox)) < 1e-12, axis=1))[0][0]
    assert np.allclose(f[c], 2.0 * (q * 0.5 / 8.0 + t * 0.5 / 4.0), atol=1e-12)  # element volume 0.5, face area 0.5
    # 2D: traction on the top edge of a 2x2 grid, order-2 elements use the 2-point face rule
    import ikarus_oracle as o
    m2 = o.structured_mesh((2, 2), (2.0, 2.0), order=2)
    pm = ik.planeStrain(ik.Materials.LinearElasticity(p))
    faces2 = [(e, 3) for e in range(m2.n_elem) if abs(m2.corner_coords[e][:, 1].max() - 2.0) < 1e-12]
    fes2 = ik.makeFE(dict(dim=2, order=2, n_dof=m2.n_nodes * 2), ik.skills(ik.linearElastic(pm), ik.neumannBoundaryLoad(
        faces2, lambda x, lam: lam * np.array([x[0], 1.0]))), m2.corner_coords, m2.elem_dofs())
    f2 = fes2.sample_external_load(1.0).reshape(-1, 2)
    assert np.allclose(f2.sum(0), [2.0, 2.0], atol=1e-12)  # int_0^2 x dx = 2, int_0^2 1 dx = 2
    assert np.all(f2[m2.node_coords[:, 1] < 2.0 - 1e-9] == 0.0)
