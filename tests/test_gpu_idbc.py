"""inhomogeneous Dirichlet values through the device path -- ikb_idbc_forces (K_raw * dd_D/dlambda on the
GPU), the NewtonRaphson sync branch and LoadControl -- against the oracle and the reference's elastic strip results."""
import numpy as np
import pytest

import ikarus_b200 as ik
import ikarus_oracle as o
from devproblems import device_assembler
from problems import ELASTIC_STRIP_EXPECTED, distorted, elastic_strip

pytestmark = pytest.mark.gpu
_MAT = {"svk": ik.Materials.StVenantKirchhoff, "neohooke": ik.Materials.NeoHooke}


def _device_strip(case, mode):
    mesh, kind, mat, flags, value, deriv, probe = elastic_strip(*case)
    n = flags.shape[0]
    dv = ik.DirichletValues(n, nodeCoords=mesh.node_coords)
    dv.container()[:] = flags
    dv.storeInhomogeneousBoundaryCondition(value)  # derivative by complex step
    p = ik.fe.LamesFirstParameterAndShearModulus(mat.lam, mat.mu)
    fes = ik.makeFE(dict(dim=2, order=case[1], n_dof=n), ik.skills(ik.nonLinearElastic(ik.planeStrain(_MAT[case[0]](p)))),
                    mesh.corner_coords, mesh.elem_dofs("interleaved"))
    return mesh, dv, ik.makeSparseFlatAssembler(fes, dv, mode=mode), probe


@pytest.mark.parametrize("dbc", [ik.DBCOption.Full, ik.DBCOption.Reduced], ids=["full", "reduced"])
@pytest.mark.parametrize("case", [("svk", 1), ("neohooke", 1), ("svk", 2), ("neohooke", 2)], ids=str)
def test_elastic_strip_reference_numbers(case, dbc):
    mesh, dv, asm, probe = _device_strip(case, "mirror")
    assert dv.fixedDOFsize() == 4 * (case[1] * 10 + 1)
    req = ik.FERequirements(np.zeros(asm.size()), 0.0)
    asm.bind(req, ik.elastoStatics, dbc)
    cfg = ik.NewtonRaphsonConfig(ik.NRSettings(tol=1e-8), ik.solvers.SparseDirectSolver(), ik.obtainForcesDueToIDBC)
    lc = ik.LoadControl(ik.NewtonRaphson(asm, cfg), ik.LoadControlConfig(1, 0.0, 1.0))
    info = lc.run(req)
    its, disp = ELASTIC_STRIP_EXPECTED[case]
    assert info.success and info.totalIterations == its
    d = req.globalSolution()
    assert abs(d[probe] - disp) < 1e-8 and req.parameter() == 1.0
    inc = dv.evaluateInhomogeneousBoundaryCondition(1.0)
    assert np.abs(d[inc != 0] - inc[inc != 0]).max() < 1e-8
    assert np.linalg.norm(asm.vector(req)) < 1e-8


def test_elastic_strip_resident_pcg():
    case = ("neohooke", 1)
    mesh, dv, asm, probe = _device_strip(case, "resident")
    req = ik.FERequirements(np.zeros(asm.size()), 0.0)
    asm.bind(req, ik.elastoStatics, ik.DBCOption.Full)
    cfg = ik.NewtonRaphsonConfig(ik.NRSettings(tol=1e-8), ik.DeviceLinearSolver(1e-13), ik.obtainForcesDueToIDBC)
    info = ik.LoadControl(ik.NewtonRaphson(asm, cfg), ik.LoadControlConfig(1, 0.0, 1.0)).run(req)
    its, disp = ELASTIC_STRIP_EXPECTED[case]
    assert info.success and info.totalIterations == its
    assert abs(req.globalSolution()[probe] - disp) < 1e-8


@pytest.mark.parametrize("dim", [2, 3])
def test_idbc_forces_match_oracle(dim):
    cells = (5, 4) if dim == 2 else (4, 3, 2)
    bbox = tuple(float(c) for c in cells)
    mesh = distorted(o.structured_mesh(cells, bbox, order=1), 0.1, 4)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material("neohooke", lam, mu, plane_strain=(dim == 2))
    kind = o.ElementKind(dim, 1, "gl")
    base = o.structured_mesh(cells, bbox, order=1)
    right = o.boundary_nodes(base, 0, bbox[0])
    flags = o.fix_nodes(mesh, o.boundary_nodes(base, 0, 0.0))
    on_right = np.zeros(mesh.n_nodes, dtype=bool)
    on_right[right] = True
    coords = mesh.node_coords
    idx = {tuple(np.round(c, 12)): i for i, c in enumerate(coords)}
    value = lambda x, l_: tuple((0.3 * (k + 1) * l_ + 0.1 * l_ * l_) if on_right[idx[tuple(np.round(x, 12))]] else 0.0 * l_
                                for k in range(dim))
    deriv = lambda x, l_: tuple((0.3 * (k + 1) + 0.2 * l_) if on_right[idx[tuple(np.round(x, 12))]] else 0.0
                                for k in range(dim))
    idbc = o.InhomogeneousDirichlet(mesh, [(value, deriv)])
    flags = idbc.flag(flags)
    ref = o.FlatAssembler(mesh, kind, mat, flags, "interleaved")
    n = flags.shape[0]
    dv = ik.DirichletValues(n, nodeCoords=coords)
    dv.container()[:] = o.fix_nodes(mesh, o.boundary_nodes(base, 0, 0.0))
    dv.storeInhomogeneousBoundaryCondition(value)
    assert np.array_equal(dv.container(), flags)
    assert np.abs(dv.evaluateInhomogeneousBoundaryConditionDerivative(1.0) - idbc.derivative(1.0)).max() < 1e-14
    p = ik.fe.LamesFirstParameterAndShearModulus(mat.lam, mat.mu)
    m = ik.Materials.NeoHooke(p)
    fes = ik.makeFE(dict(dim=dim, order=1, n_dof=n), ik.skills(ik.nonLinearElastic(ik.planeStrain(m) if dim == 2 else m)),
                    mesh.corner_coords, mesh.elem_dofs("interleaved"))
    dev = ik.makeSparseFlatAssembler(fes, dv)
    d = 0.02 * np.random.default_rng(8).uniform(-1, 1, n)
    req = ik.FERequirements(d, 0.6)
    for mode, dbc in (("full", ik.DBCOption.Full), ("reduced", ik.DBCOption.Reduced)):
        dev.bind(req, ik.elastoStatics, dbc)
        F = dev.obtainForcesDueToIDBC()
        Fref = o.idbc_forces(ref, d, 0.6, mode, idbc)
        assert F.shape == Fref.shape
        assert np.abs(F - Fref).max() <= 1e-12 * np.abs(Fref).max()
        if mode == "full":
            assert not F[flags].any()
