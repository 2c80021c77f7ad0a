"""GPU parity for the principal-stretch hyperelastic framework on the device (IKB_MAT_BLATZKO, FORM_PS in the
generalised-tangent kernel of ikb_elem_easdg.cuh): Materials::Hyperelastic<Deviatoric<BlatzKo>, Volumetric<VF0>>
(materials/hyperelastic/interface.hh:99-232, deviatoric/interface.hh:77-115, deviatoric/blatzko.hh:60-92) on the plain
Q1 element, under the strain enhancements E4..E21 and under both displacement-gradient enhancements -- against the
oracle, which is pinned on the reference's six Blatz-Ko cantilever known answers, and against those known answers
themselves through the device path."""
import numpy as np
import pytest

import ikarus_b200 as ik
import ikarus_oracle as o
from devproblems import device_assembler, entry_error
from problems import cantilever, distorted
from golden_data import GOLDEN

pytestmark = pytest.mark.gpu
RT = ik.ResultTypes
TOL = 1e-12      # relative to the row maximum (see tests/test_gpu_eas.py for the two norms)
TOL_8D = 5e-12   # SURVEY 8d norm; the condensed tangents subtract matrices of the size of K itself


def _row_scaled_error(dev, ref, rows):
    rowmax = np.zeros(rows.max() + 1)
    np.maximum.at(rowmax, rows, np.abs(ref))
    rm = np.where(rowmax[rows] == 0.0, 1.0, rowmax[rows])
    return float((np.abs(np.asarray(dev) - np.asarray(ref)) / rm).max(initial=0.0))


# (dim, m, enhancement)
CASES = [(3, 0, "strain"), (2, 0, "strain"), (3, 9, "strain"), (3, 21, "strain"), (2, 4, "strain"), (2, 5, "strain"),
         (2, 7, "strain"), (3, 9, "dg"), (3, 9, "dgt"), (2, 4, "dg"), (2, 4, "dgt")]


def _setup(dim, m, fn, seed=3):
    cells = (3, 2, 2) if dim == 3 else (4, 3)
    box = tuple(float(c) for c in cells)
    mesh = distorted(o.structured_mesh(cells, box), 0.15, seed + 2)
    mat = o.Material("blatzko", 0.0, 40.0, plane_strain=(dim == 2))
    kind = o.ElementKind(dim, 1, "gl", m, eas_function=fn)
    flags = o.fix_nodes(mesh, o.boundary_nodes(o.structured_mesh(cells, box), 0, 0.0))
    rng = np.random.default_rng(seed)
    n = flags.shape[0]
    d = 0.05 * rng.uniform(-1, 1, n)
    alpha = 0.01 * rng.uniform(-1, 1, (mesh.n_elem, m))
    return mesh, kind, mat, flags, d, alpha, rng


@pytest.mark.parametrize("dim,m,fn", CASES, ids=lambda v: str(v))
def test_blatzko_matrix_vector_energy_and_alpha_update(dim, m, fn):
    mesh, kind, mat, flags, d, alpha, rng = _setup(dim, m, fn)
    ref = o.FlatAssembler(mesh, kind, mat, flags)
    dev = device_assembler(mesh, kind, mat, flags)
    if m:
        ref.alpha = alpha.copy()
        dev.setInternalVariables(alpha)
    req = ik.FERequirements(d, 0.0)
    for mode, dbc in (("raw", ik.DBCOption.Raw), ("full", ik.DBCOption.Full), ("reduced", ik.DBCOption.Reduced)):
        K = dev.matrix(req, ik.MatrixAffordance.stiffness, dbc)
        outer, inner = ref.pattern(mode)
        assert np.array_equal(K.indptr, outer) and np.array_equal(K.indices, inner)
        rows = np.repeat(np.arange(outer.shape[0] - 1), np.diff(outer))
        Kref = ref.matrix_values(d, 0.0, mode)
        assert entry_error(K.data, Kref, rows) <= TOL_8D, mode
        assert _row_scaled_error(K.data, Kref, rows) <= TOL, mode
        R = dev.vector(req, ik.VectorAffordance.forces, dbc)
        Rr = ref.vector(d, 0.0, mode)
        assert np.abs(R - Rr).max() <= TOL_8D * np.abs(Rr).max(), mode
    Kd = dev.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Raw).toarray()
    assert np.array_equal(Kd, Kd.T)
    if m == 0:
        E = dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
        Er = ref.scalar(d, 0.0)
        assert abs(E - Er) <= 1e-12 * abs(Er)
    else:
        corr = 0.01 * rng.uniform(-1, 1, d.shape[0])
        dev.updateInternalVariables(req, corr)
        ref.update_eas(d, corr)
        assert np.abs(dev.internalVariables() - ref.alpha).max() <= 1e-11 * max(1.0, np.abs(ref.alpha).max())
        with pytest.raises(NotImplementedError):
            dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)


def test_blatzko_undeformed_state_takes_the_equal_stretch_branch():
    """At d = 0 all stretches are 1: the L_ikik moduli come from the limit 0.5 (L_iiii - L_iikk) that the reference takes
    under Dune::FloatCmp::eq(lambda_i, lambda_k, 1e-8) (deviatoric/interface.hh:98-106); R = 0 there."""
    mesh, kind, mat, flags, d, _, _ = _setup(3, 0, "strain")
    ref = o.FlatAssembler(mesh, kind, mat, flags)
    dev = device_assembler(mesh, kind, mat, flags)
    z = np.zeros_like(d)
    req = ik.FERequirements(z, 0.0)
    K = dev.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Raw)
    outer, _ = ref.pattern("raw")
    rows = np.repeat(np.arange(outer.shape[0] - 1), np.diff(outer))
    assert entry_error(K.data, ref.matrix_values(z, 0.0, "raw"), rows) <= TOL
    assert np.abs(dev.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Raw)).max() <= 1e-12


@pytest.mark.parametrize("dim,m,fn", [(3, 0, "strain"), (2, 0, "strain"), (3, 21, "strain"), (2, 4, "dgt"), (3, 9, "dg")],
                         ids=lambda v: str(v))
def test_blatzko_results_match_oracle(dim, m, fn):
    mesh, kind, mat, flags, d, alpha, rng = _setup(dim, m, fn, seed=6)
    n = d.shape[0]
    dev = device_assembler(mesh, kind, mat, np.zeros(n, dtype=bool))
    req = ik.FERequirements(d, 0.0)
    a = None
    if m:
        a = alpha
        dev.setInternalVariables(alpha)
    pts = np.vstack([np.full(dim, 0.5), rng.uniform(0, 1, (3, dim)), np.zeros(dim), np.ones(dim)])
    u = d[mesh.elem_dofs("interleaved")].reshape(mesh.n_elem, -1, dim)
    for rt, name in ((RT.PK2Stress, "native"), (RT.PK2StressFull, "full"), (RT.kirchhoffStress, "kirchhoff"),
                     (RT.cauchyStress, "cauchy")):
        S = dev.calculateAt(rt, req, pts)
        for q, xi in enumerate(pts):
            r = o.stress_at(kind, mat, mesh.corner_coords, u, xi, alpha=a, result=name)
            assert S[:, q].shape == r.shape
            assert np.abs(S[:, q] - r).max() <= 1e-11 * np.abs(r).max(), (rt, q)


@pytest.mark.parametrize(
    "dim,m,fn,iters,maxd",
    # tests/src/testcantileverbeamEAS.cpp: the Blatz-Ko rows (tests/golden/reference_known_answers.json)
    [(c["dim"], c["eas"], c["function"], c["newton_iterations"], c["max_abs_d"])
     for c in GOLDEN["cantilever_eas_blatzko"]["cases"]],
)
def test_reference_blatzko_cantilever_known_answers_on_device(dim, m, fn, iters, maxd):
    """The reference's own known answers for makeBlatzKo(40.0) (80 Newton iterations, max|d| to 1e-10) with the device
    assembler behind NewtonRaphson + LoadControl, as in tests/src/testcantileverbeam.hh:83-198."""
    mesh, kind, _, flags, fext = cantilever(dim, "neohooke", m)
    kind.eas_function = fn
    mat = o.Material("blatzko", 0.0, 40.0, plane_strain=(dim == 2))
    dev = device_assembler(mesh, kind, mat, flags, fext=fext)
    req = ik.FERequirements(np.zeros(flags.shape[0]), 0.0)
    dev.bind(req, ik.AffordanceCollection(vector=ik.VectorAffordance.forces, matrix=ik.MatrixAffordance.stiffness),
             ik.DBCOption.Full)
    nr = ik.NewtonRaphson(dev, ik.NewtonRaphsonConfig(ik.NRSettings(tol=1e-10)))
    info = ik.LoadControl(nr, ik.LoadControlConfig(20, 0.0, 1.0)).run(req)
    assert info.success and info.totalIterations == iters
    assert abs(np.abs(req.globalSolution()).max() - maxd) < 1e-10


def test_blatzko_descriptor_rules():
    """Served: Q1 family, 3D or planeStrain.  Quad9/Hex27 and planeStress answer NotImplemented."""
    lam, mu = 0.0, 40.0
    mesh = o.structured_mesh((2, 2), (1.0, 1.0), order=2)
    with pytest.raises(NotImplementedError):
        device_assembler(mesh, o.ElementKind(2, 2, "gl"), o.Material("blatzko", lam, mu, plane_strain=True),
                         np.zeros(mesh.n_nodes * 2, dtype=bool))
    mesh = o.structured_mesh((2, 2), (1.0, 1.0))
    with pytest.raises(NotImplementedError):
        device_assembler(mesh, o.ElementKind(2, 1, "gl"), o.Material("blatzko", lam, mu, plane_stress=True),
                         np.zeros(mesh.n_nodes * 2, dtype=bool))


# ------------------------------------------------------------------ every law of the framework (IKB_MAT_HYPERELASTIC)
_MU = 1000.0 / (2.0 * 1.25)  # the parameters of tests/src/testhyperelasticity.hh:231-249 (E = 1000, nu = 0.25)
_LAME = 1000.0 * 0.25 / (1.25 * 0.5)
_BULK = 1000.0 / (3.0 * 0.5)
_OG = ((2.0 * _MU / 3.0, _MU / 6.0, _MU / 6.0), (1.23, 0.59, 0.18))
LAWS = {
    "ogden_total+VF3": o.Hyper("ogden_total", _OG, vf=3, K=_LAME),
    "ogden_dev+VF2": o.Hyper("ogden_dev", _OG, vf=2, K=_BULK),
    "mooney_rivlin+VF1": o.Hyper("invariant", ((1, 0), (0, 1), (_MU / 2.0, _MU / 2.0)), vf=1, K=_BULK),
    "yeoh+VF5": o.Hyper("invariant", ((1, 2, 3), (0, 0, 0), (_MU / 2.0, _MU / 6.0, _MU / 3.0)), vf=5, K=_BULK),
    "arruda_boyce+VF6": o.Hyper("arrudaboyce", (_MU, 0.85), vf=6, K=_BULK),
    "gent+VF8": o.Hyper("gent", (_MU, 2.5), vf=8, K=_BULK),
    "ogden1_dev+VF4": o.Hyper("ogden_dev", ((_MU,), (2.0,)), vf=4, K=_BULK, beta=0.5),
    "mooney_rivlin+VF7": o.Hyper("invariant", ((1, 0), (0, 1), (_MU / 2.0, _MU / 2.0)), vf=7, K=_BULK, beta=0.5),
    "yeoh+VF9": o.Hyper("invariant", ((1, 2, 3), (0, 0, 0), (_MU / 2.0, _MU / 6.0, _MU / 3.0)), vf=9, K=_BULK),
    "gent+VF10": o.Hyper("gent", (_MU, 2.5), vf=10, K=_BULK, beta=0.4),
    "arruda_boyce+VF11": o.Hyper("arrudaboyce", (_MU, 0.85), vf=11, K=_BULK),
    "ogden_total_no_volumetric": o.Hyper("ogden_total", _OG),
    # the law of the reference's incompressible block (tests/src/testincompressibleblock.cpp:60); VF12 alone has a tangent
    # whose rows cancel to rounding noise, so it is paired with a deviatoric part here
    "ogden1_dev+VF12": o.Hyper("ogden_dev", ((_MU,), (2.0,)), vf=12, K=_BULK),
    "pure_volumetric_VF3": o.Hyper("none", (), vf=3, K=_BULK),
}


@pytest.mark.parametrize("dim,m,fn", [(3, 0, "strain"), (2, 0, "strain"), (3, 9, "strain"), (2, 4, "dgt")], ids=lambda v: str(v))
@pytest.mark.parametrize("law", sorted(LAWS))
def test_every_hyperelastic_law_matches_the_oracle(law, dim, m, fn):
    """Hyperelastic<Deviatoric<DF>, Volumetric<VF>> for every deviatoric function and VF1..VF12 through
    ikb_set_hyperelastic: K, R (and the energy of the plain element) and the PK2 stress against the oracle, whose laws
    reproduce the reference's material result tables (tests/test_hyperelastic_oracle.py)."""
    if law.startswith("pure_volumetric") and m:
        pytest.skip("a purely volumetric law leaves the enhanced block singular")
    mesh, kind, _, flags, d, alpha, rng = _setup(dim, m, fn)
    mat = o.Material("hyperelastic", 0.0, 0.0, plane_strain=(dim == 2), hyper=LAWS[law])
    ref = o.FlatAssembler(mesh, kind, mat, flags)
    dev = device_assembler(mesh, kind, mat, flags)
    if m:
        ref.alpha = alpha.copy()
        dev.setInternalVariables(alpha)
    req = ik.FERequirements(d, 0.0)
    for mode, dbc in (("raw", ik.DBCOption.Raw), ("reduced", ik.DBCOption.Reduced)):
        K = dev.matrix(req, ik.MatrixAffordance.stiffness, dbc)
        outer, _ = ref.pattern(mode)
        rows = np.repeat(np.arange(outer.shape[0] - 1), np.diff(outer))
        Kref = ref.matrix_values(d, 0.0, mode)
        assert entry_error(K.data, Kref, rows) <= TOL_8D, mode
        assert _row_scaled_error(K.data, Kref, rows) <= TOL, mode
        R = dev.vector(req, ik.VectorAffordance.forces, dbc)
        Rr = ref.vector(d, 0.0, mode)
        assert np.abs(R - Rr).max() <= TOL_8D * np.abs(Rr).max(), mode
    if m == 0:
        E = dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
        Er = ref.scalar(d, 0.0)
        assert abs(E - Er) <= 1e-11 * abs(Er)
    u = d[mesh.elem_dofs("interleaved")].reshape(mesh.n_elem, -1, dim)
    xi = np.full(dim, 0.3)
    S = dev.calculateAt(RT.PK2Stress, req, xi[None])
    r = o.stress_at(kind, mat, mesh.corner_coords, u, xi, alpha=alpha if m else None, result="native")
    assert np.abs(S[:, 0] - r).max() <= 1e-11 * np.abs(r).max()


def test_neohooke_recovery_on_the_device():
    """tests/src/testhyperelasticity.hh:139-196 through the assembler: Ogden<1, total>({mu}, {2}) with VF3 and Lame's
    first parameter IS the NeoHooke material -- the generalised-tangent kernel with the principal-stretch law against
    the factored NeoHooke kernels (tensor-core Hex8 kernel included), two independent device formulations."""
    for dim in (3, 2):
        mesh, kind, _, flags, d, _, _ = _setup(dim, 0, "strain")
        nh = device_assembler(mesh, kind, o.Material("neohooke", _LAME, _MU, plane_strain=(dim == 2)), flags)
        og = device_assembler(mesh, kind, o.Material("hyperelastic", 0.0, 0.0, plane_strain=(dim == 2),
                                                     hyper=o.Hyper("ogden_total", ((_MU,), (2.0,)), vf=3, K=_LAME)), flags)
        req = ik.FERequirements(d, 0.0)
        Ka = nh.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Full)
        Kb = og.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Full)
        rows = np.repeat(np.arange(Ka.indptr.shape[0] - 1), np.diff(Ka.indptr))
        assert entry_error(Kb.data, Ka.data, rows) <= TOL_8D
        Ra = nh.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Full)
        Rb = og.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Full)
        assert np.abs(Ra - Rb).max() <= TOL_8D * np.abs(Ra).max()
        Ea = nh.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
        Eb = og.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
        assert abs(Ea - Eb) <= 1e-11 * abs(Ea)


def test_hyperelastic_descriptor_rules():
    mesh = o.structured_mesh((2, 2, 2), (1.0, 1.0, 1.0))
    flags = np.zeros(mesh.n_nodes * 3, dtype=bool)
    with pytest.raises(ValueError):  # InvariantBasedT::checkExponents (invariantbased.hh:216-221)
        ik.Materials.makeInvariantBased((500.0, 500.0), (0, 0), (0, 1))
    with pytest.raises(NotImplementedError):  # more than three terms
        ik.Materials.makeOgden((1.0,) * 4, (2.0,) * 4)
    # Gent beyond its limiting stretch: Jm <= W1 - 3 throws in the reference (gent.hh:176-180)
    mat = o.Material("hyperelastic", 0.0, 0.0, hyper=o.Hyper("gent", (_MU, 1e-3)))
    dev = device_assembler(mesh, o.ElementKind(3, 1, "gl"), mat, flags)
    d = 0.2 * np.random.default_rng(0).uniform(-1, 1, flags.shape[0])
    with pytest.raises(Exception):
        dev.matrix(ik.FERequirements(d, 0.0), ik.MatrixAffordance.stiffness, ik.DBCOption.Raw)
