"""SURVEY 8f-2 (planeStress part): Materials::planeStress (VanishingStress) in the Q1, Q2 and EAS element kernels, the
result kernel and the solvers, against the oracle and the reference's known answers."""
import numpy as np
import pytest

import ikarus_b200 as ik
import ikarus_oracle as o
from devproblems import device_assembler, entry_error
from golden_data import GOLDEN
from problems import PATCH_EXPECTED_D, distorted, fixed_distorted_quad, patch_test_mesh

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [
    # cells, order, strain, material, eas_m
    ((5, 4), 1, "gl", "neohooke", 0),
    ((5, 4), 1, "gl", "svk", 0),
    ((6, 3), 1, "linear", "linear", 0),
    ((3, 2), 2, "gl", "neohooke", 0),
    ((3, 2), 2, "gl", "svk", 0),
    ((3, 2), 2, "linear", "linear", 0),
    ((4, 3), 1, "gl", "neohooke", 4),
    ((4, 3), 1, "gl", "svk", 7),
    ((4, 3), 1, "linear", "linear", 5),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"Q{c[1]}-{c[2]}-{c[3]}-eas{c[4]}")
def test_plane_stress_assembly_matches_oracle(case):
    cells, order, strain, matk, m = case
    bbox = tuple(float(c) for c in cells)
    mesh = distorted(o.structured_mesh(cells, bbox, order=order), 0.12, 2)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material(matk, lam, mu, plane_stress=True, ps_tol=1e-10)
    kind = o.ElementKind(2, order, strain, m)
    flags = o.fix_nodes(mesh, o.boundary_nodes(o.structured_mesh(cells, bbox, order=order), 0, 0.0))
    rng = np.random.default_rng(6)
    n = flags.shape[0]
    fext = rng.uniform(-1, 1, n)
    d = 0.04 * rng.uniform(-1, 1, n)
    ref = o.FlatAssembler(mesh, kind, mat, flags, "interleaved", fext=fext)
    dev = device_assembler(mesh, kind, mat, flags, "interleaved", fext=fext)
    if m:
        alpha = 0.01 * rng.uniform(-1, 1, (mesh.n_elem, m))
        ref.alpha[:] = alpha
        dev.setInternalVariables(alpha)
    req = ik.FERequirements(d, 0.5)
    for mode, dbc in (("raw", ik.DBCOption.Raw), ("full", ik.DBCOption.Full), ("reduced", ik.DBCOption.Reduced)):
        if m and mode == "reduced":
            continue
        outer, inner = ref.pattern(mode)
        K = dev.matrix(req, ik.MatrixAffordance.stiffness, dbc)
        rows = np.repeat(np.arange(outer.shape[0] - 1), np.diff(outer))
        assert entry_error(K.data, ref.matrix_values(d, 0.5, mode), rows) <= TOL, mode
        R = dev.vector(req, ik.VectorAffordance.forces, dbc)
        Rref = ref.vector(d, 0.5, mode)
        assert np.abs(R - Rref).max(initial=0.0) <= TOL * np.abs(Rref).max(initial=1.0), mode
    if not m:
        E = dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
        Eref = ref.scalar(d, 0.5)
        assert abs(E - Eref) <= 1e-12 * max(1.0, abs(Eref))
    # stresses at local positions: reduced law (sigma_33 = 0 solution) and the zero-extended 3D law
    u = d[mesh.elem_dofs()].reshape(mesh.n_elem, -1, 2)
    a = ref.alpha if (m and strain == "gl") else None
    RT = ik.ResultTypes
    types = [(RT.linearStress, "native"), (RT.linearStressFull, "full")] if strain == "linear" else [
        (RT.PK2Stress, "native"), (RT.PK2StressFull, "full"), (RT.cauchyStress, "cauchy")]
    pts = np.array([[0.5, 0.5], [0.1, 0.9], [1.0, 0.0]])
    for rt, name in types:
        S = dev.calculateAt(rt, req, pts)
        for q, xi in enumerate(pts):
            Sref = o.stress_at(kind, mat, mesh.corner_coords, u, xi, alpha=a, result=name)
            assert np.abs(S[:, q] - Sref).max() <= 1e-11 * np.abs(Sref).max(), (rt, q)


def test_B2_single_element_eigenvalues_on_device():
    # tests/src/testnonlinearelasticity.hh:248-338: planeStress(SVK, nu = 0), fixed distorted Quad4
    mesh = fixed_distorted_quad()
    lam, mu = o.lame_from_E_nu(1000.0, 0.0)
    mat = o.Material("svk", lam, mu, plane_stress=True, ps_tol=1e-8)
    dev = device_assembler(mesh, o.ElementKind(2, 1, "gl"), mat, np.zeros(8, dtype=bool))
    gold = GOLDEN["plane_stress_single_element_eigenvalues"]
    req = ik.FERequirements(np.array(gold["d"], float), 0.0)
    K = dev.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Raw).toarray()
    exp = np.array(gold["abs_eigenvalues"], float)
    assert np.abs(np.sort(np.abs(np.linalg.eigvalsh(K))) - exp).max() < 1e-8


@pytest.mark.parametrize("dbc", [ik.DBCOption.Full, ik.DBCOption.Reduced], ids=["full", "reduced"])
def test_linear_patch_test_with_idbc_on_device(dbc):
    # tests/src/testinhomogeneousdbc.cpp:46-160: planeStress(LinearElasticity), inhomogeneous u_x at x = 0.24
    mesh = patch_test_mesh()
    dv = ik.DirichletValues(16, nodeCoords=mesh.node_coords)
    for i in (0, 1, 4):
        dv.setSingleDOF(i, True)
    dv.storeInhomogeneousBoundaryCondition(lambda x, l_: (0.001 * l_ if abs(x[0] - 0.24) < 1e-12 else 0.0 * l_, 0.0 * l_))
    assert dv.fixedDOFsize() == 5
    p = ik.toLamesFirstParameterAndShearModulus(emodul=1000.0, nu=0.25)
    fes = ik.makeFE(dict(dim=2, order=1, n_dof=16), ik.skills(ik.linearElastic(ik.planeStress(ik.Materials.LinearElasticity(p)))),
                    mesh.corner_coords, mesh.elem_dofs())
    asm = ik.makeSparseFlatAssembler(fes, dv)
    req = ik.FERequirements(np.zeros(16), 1.0)
    asm.bind(req, ik.elastoStatics, dbc)
    import scipy.sparse.linalg as spla
    K = asm.matrix()
    R = asm.vector() + asm.obtainForcesDueToIDBC()
    x = spla.spsolve(K.tocsc(), -R)
    d = req.globalSolution()
    d[:] = x if dbc == ik.DBCOption.Full else asm.createFullVector(x)
    inc = dv.evaluateInhomogeneousBoundaryCondition(1.0)
    d[inc != 0] = inc[inc != 0]
    big = np.abs(PATCH_EXPECTED_D) > 1e-10
    assert np.abs(d[big] - PATCH_EXPECTED_D[big]).max() < 1e-10
    sig = asm.calculateAt(ik.ResultTypes.linearStress, req, [0.5, 0.5])
    assert np.abs(sig[:, 0, 0] - GOLDEN["plane_stress_patch_test"]["sigma_xx"]).max() < 1e-10
