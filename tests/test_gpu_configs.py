"""GPU parity on scaled-down versions of the BASELINE.json configs (full sizes are bench / tools territory)."""
import numpy as np
import pytest

import ikarus_b200 as ik
import ikarus_oracle as o
from devproblems import device_assembler, entry_error

pytestmark = pytest.mark.gpu


def test_config0_cooks_membrane_one_reduced_linear_solve():
    """configs[0]: 2D Cook's membrane, Q1 plane-strain LinearElastic, one DBCOption::Reduced solve
    (x=48 xi, y=44 xi + eta (44 - 28 xi); E=1, nu=1/3; x=0 clamped; traction (0, 1/16) on x=48)."""
    nx = 16

    def cook(c):
        xi, eta = c[:, 0], c[:, 1]
        return np.stack([48.0 * xi, 44.0 * xi + eta * (44.0 - 28.0 * xi)], axis=-1)

    mesh = o.structured_mesh((nx, nx), (1.0, 1.0), mapping=cook)
    lam, mu = o.lame_from_E_nu(1.0, 1.0 / 3.0)
    mat = o.Material("linear", lam, mu, True)
    kind = o.ElementKind(2, 1, "linear")
    flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 0, 0.0))
    # consistent nodal load of the traction on the edge x = 48 (2-point Gauss per edge segment == exact for linear N)
    fext = np.zeros(mesh.n_nodes * 2)
    right = o.boundary_nodes(mesh, 0, 48.0)
    right = right[np.argsort(mesh.node_coords[right, 1])]
    for a, b in zip(right[:-1], right[1:]):
        L = np.linalg.norm(mesh.node_coords[b] - mesh.node_coords[a])
        for nd in (a, b):
            fext[2 * nd + 1] += 0.5 * L / 16.0
    ref = o.FlatAssembler(mesh, kind, mat, flags, fext=fext)
    dev = device_assembler(mesh, kind, mat, flags, fext=fext, mode="resident")
    d0 = np.zeros(ref.n)
    req = ik.FERequirements(d0, 1.0)
    dev.bind(req, ik.elastoStatics, ik.DBCOption.Reduced)
    R = dev.vector()
    K = dev.matrix()
    # reduced pattern and values against the oracle
    outer, inner = ref.pattern("reduced")
    Kh = K.to_scipy()
    assert np.array_equal(Kh.indptr, outer) and np.array_equal(Kh.indices, inner)
    rows = np.repeat(np.arange(outer.shape[0] - 1), np.diff(outer))
    assert entry_error(Kh.data, ref.matrix_values(d0, 1.0, "reduced"), rows) <= 1e-12
    ls = ik.DeviceLinearSolver(relTol=1e-14)
    dred = -ls(R, K)
    d = dev.createFullVector(dred)
    import scipy.sparse.linalg as spla
    dref = ref.create_full_vector(spla.spsolve(ref.matrix(d0, 1.0, "reduced").tocsc(), -ref.vector(d0, 1.0, "reduced")))
    assert np.abs(d - dref).max() <= 1e-9 * np.abs(dref).max()
    tip = np.nonzero(np.all(np.abs(mesh.node_coords - np.array([48.0, 60.0])) < 1e-9, axis=1))[0][0]
    assert d[2 * tip + 1] > 0  # the tip moves up under the upward shear load


def test_config2_hex27_svk_newton_with_device_pcg():
    """configs[2] scaled down: Q2 hexahedra (81-dof elements), StVenantKirchhoff, Newton with device CG."""
    mesh = o.structured_mesh((3, 2, 2), (1.5, 1.0, 1.0), order=2)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material("svk", lam, mu)
    kind = o.ElementKind(3, 2, "gl")
    flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 2, 0.0))
    n = mesh.n_nodes * 3
    fext = np.zeros(n)
    fext[0::3] = 2.0 / mesh.n_nodes  # lambda-proportional shear load
    ref = o.FlatAssembler(mesh, kind, mat, flags, fext=fext)
    dr, lamr, inf = o.load_control(ref, np.zeros(n), 3, 0.0, 30.0, tol=1e-8, dbc="full")
    dev = device_assembler(mesh, kind, mat, flags, fext=fext, mode="resident")
    req = ik.FERequirements(np.zeros(n), 0.0)
    dev.bind(req, ik.elastoStatics, ik.DBCOption.Full)
    nr = ik.NewtonRaphson(dev, ik.NewtonRaphsonConfig(ik.NRSettings(tol=1e-8), ik.DeviceLinearSolver(1e-13)))
    info = ik.LoadControl(nr, ik.LoadControlConfig(3, 0.0, 30.0)).run(req)
    assert info.success and inf["success"]
    assert [s.iterations for s in info.solverInfos] == inf["per_step"]
    assert np.abs(req.globalSolution() - dr).max() <= 1e-8
    assert np.abs(dr).max() > 1e-3  # genuinely nonlinear regime


def test_assembler_manipulator_style_callbacks():
    """AssemblerManipulator (assemblermanipulatorfuser.hh:242-385): host callbacks mutate the returned quantities;
    the reference's cantilever test applies its point load this way (tests/src/testcantileverbeam.hh:56-80)."""
    mesh = o.structured_mesh((4, 1, 1), (4.0, 1.0, 1.0))
    lam, mu = o.lame_from_E_nu(100.0, 0.3)
    mat = o.Material("neohooke", lam, mu)
    kind = o.ElementKind(3, 1, "gl")
    flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 0, 0.0))
    dev = device_assembler(mesh, kind, mat, flags)
    n = flags.shape[0]
    idx = n - 2
    calls = []

    def point_load(assembler, req, affordance, dbc, vec):
        calls.append((affordance, dbc))
        vec[idx] -= -req.parameter() * 1.0

    dev.bindVectorFunction(point_load)
    dev.bindMatrixFunction(lambda a, r, aff, dbc, K: K.__setitem__((0, 0), 7.0))
    dev.bindScalarFunction(lambda a, r, aff, val: val + 1.0)
    req = ik.FERequirements(np.zeros(n), 0.25)
    R = dev.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Full)
    assert R[idx] == 0.25 and calls == [(ik.VectorAffordance.forces, ik.DBCOption.Full)]
    K = dev.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Full)
    assert K[0, 0] == 7.0
    assert dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy) == 1.0  # zero energy at d = 0, plus callback


@pytest.mark.gpu
def test_loads_with_general_dependence_on_the_load_factor():
    """loads/volume.hh:67-106 and loads/traction.hh:70-138 evaluate f(x, lambda) at every call, whatever its dependence
    on lambda; the device mirror re-samples a non-proportional load whenever the load factor changes.  R and E against
    the oracle with the load vector of the respective lambda; a proportional extra nodal load rides along."""
    import ikarus_b200 as ik
    import ikarus_oracle as o

    mesh = o.structured_mesh((4, 3), (2.0, 1.5))
    lam_, mu_ = o.lame_from_E_nu(100.0, 0.3)
    mat = o.Material("neohooke", lam_, mu_, plane_strain=True)
    kind = o.ElementKind(2, 1, "gl")
    flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 0, 0.0))
    faces = [(e, 1) for e in range(mesh.n_elem) if abs(mesh.corner_coords[e][:, 0].max() - 2.0) < 1e-12]
    vol = lambda x, lam: np.array([0.0, -(lam * lam + 1.0) * (1.0 + x[0])])
    trac = lambda x, lam: np.array([np.sin(lam), 0.25 * x[1]])
    p = ik.fe.LamesFirstParameterAndShearModulus(mat.lam, mat.mu)
    sk = ik.skills(ik.nonLinearElastic(ik.planeStrain(ik.Materials.NeoHooke(p))), ik.volumeLoad(vol),
                   ik.neumannBoundaryLoad(faces, trac))
    n = mesh.n_nodes * 2
    fes = ik.makeFE(dict(dim=2, order=1, n_dof=n), sk, mesh.corner_coords, mesh.elem_dofs("interleaved"))
    dv = ik.DirichletValues(n)
    dv.container()[:] = flags
    dev = ik.makeSparseFlatAssembler(fes, dv)
    point = np.zeros(n)
    point[-1] = 0.3
    dev.setExternalLoad(point)  # lambda-proportional nodal load on top of the skills' loads
    d = 0.01 * np.random.default_rng(3).uniform(-1, 1, n)
    d[flags] = 0.0
    for lam in (0.7, 1.9, 0.0):
        fext = fes.sample_external_load(lam) + lam * point
        ref = o.FlatAssembler(mesh, kind, mat, flags, fext=fext)
        req = ik.FERequirements(d, lam)
        for mode, name in ((ik.DBCOption.Raw, "raw"), (ik.DBCOption.Full, "full"), (ik.DBCOption.Reduced, "reduced")):
            R = dev.vector(req, ik.VectorAffordance.forces, mode)
            Rr = ref.vector(d, 1.0, name)
            assert np.abs(R - Rr).max() <= 1e-12 * np.abs(Rr).max(), (lam, name)
        E = dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
        Er = ref.scalar(d, 1.0)
        assert abs(E - Er) <= 1e-12 * max(abs(Er), 1.0), lam
    # the load at lambda = 0 is not zero: this is what "not proportional" means
    assert np.abs(fes.sample_external_load(0.0)).max() > 0.1
