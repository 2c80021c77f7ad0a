"""calculateAt (stresses at local positions) on the device against the reference's vertex stress tables
(tests/src/resultcollection.hh:19-68, 150-163) and against the oracle for every element family."""
import numpy as np
import pytest

import ikarus_b200 as ik
import ikarus_oracle as o
from devproblems import device_assembler
from golden_data import GOLDEN
from problems import distorted

pytestmark = pytest.mark.gpu
RT = ik.ResultTypes


def _unit(dim):
    return o.structured_mesh((1,) * dim, (1.0,) * dim)


def test_A3_square_vertex_stress_tables():
    mesh = _unit(2)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material("linear", lam, mu, plane_strain=True)
    d = np.array([0, 0, 1, 1, 1, 1, 1, 1.0])
    verts = np.array([(0, 0), (1, 0), (0, 1), (1, 1)], float)
    gold = GOLDEN["square_vertex_stress"]
    exp = np.array(gold["no_eas"], float)
    flags = np.zeros(8, dtype=bool)
    dev = device_assembler(mesh, o.ElementKind(2, 1, "linear"), mat, flags)
    req = ik.FERequirements(d, 0.0)
    S = dev.calculateAt(RT.linearStress, req, verts)
    assert S.shape == (1, 4, 3) and np.allclose(S[0], exp, atol=1e-7)
    # linearStressFull: the 3D law behind plane strain adds sigma_zz = 1153.84615385 at vertex 0 (:53-68)
    Sf = dev.calculateAt(RT.linearStressFull, req, verts[:1])
    assert np.allclose(Sf[0, 0], [exp[0, 0], exp[0, 1], gold["sigma_zz_vertex0_full_3d_law"], 0, 0, exp[0, 2]], atol=1e-7)
    # with EAS(4): alpha = -D^-1 L d is recomputed from d (resultcollection.hh:39-51)
    exp4 = np.array(gold["eas4"], float)
    dev4 = device_assembler(mesh, o.ElementKind(2, 1, "linear", 4), mat, flags)
    S4 = dev4.calculateAt(RT.linearStress, req, verts)
    assert np.allclose(S4[0], exp4, atol=1e-7)
    with pytest.raises(Exception):
        dev.calculateAt(RT.PK2Stress, req, verts)  # not a result of the linear element


def test_A4_cube_vertex_stress_table():
    mesh = _unit(3)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    d = np.zeros(24)
    d[6:9] = 1.0
    exp = np.array(GOLDEN["cube_vertex_stress"]["values"], float)
    verts = np.array([[(v >> k) & 1 for k in range(3)] for v in range(8)], float)
    dev = device_assembler(mesh, o.ElementKind(3, 1, "linear"), o.Material("linear", lam, mu), np.zeros(24, dtype=bool))
    S = dev.calculateAt(RT.linearStress, ik.FERequirements(d, 0.0), verts)
    assert np.allclose(S[0], exp, atol=1e-7)


CASES = [
    # dim, cells, order, strain, material, eas_m
    (3, (3, 2, 2), 1, "gl", "neohooke", 0),
    (3, (2, 2, 2), 1, "gl", "svk", 0),
    (3, (2, 2, 1), 2, "gl", "neohooke", 0),
    (2, (4, 3), 1, "gl", "neohooke", 0),
    (2, (3, 2), 2, "gl", "svk", 0),
    (2, (3, 3), 2, "linear", "linear", 0),
    (3, (2, 2, 2), 1, "gl", "neohooke", 21),
    (3, (2, 2, 2), 1, "linear", "linear", 9),
    (2, (3, 3), 1, "gl", "svk", 7),
    (2, (3, 3), 1, "linear", "linear", 5),
    # displacement-gradient enhancements (H9 / H4): the stress at the ENHANCED displacement gradient with the stored
    # alpha (enhancedassumedstrains.hh:161-170); the seventh entry is the EAS function
    (3, (2, 2, 2), 1, "gl", "neohooke", 9, "dg"),
    (3, (2, 2, 2), 1, "gl", "svk", 9, "dgt"),
    (2, (3, 3), 1, "gl", "neohooke", 4, "dgt"),
    (2, (3, 3), 1, "gl", "svk", 4, "dg"),
]


@pytest.mark.parametrize("layout", ["interleaved", "lexicographic"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-Q{c[2]}-{c[3]}-{c[4]}-eas{c[5]}{c[6] if len(c) > 6 else ''}")
def test_results_match_oracle(case, layout):
    dim, cells, order, strain, matk, m = case[:6]
    fn = case[6] if len(case) > 6 else "strain"
    bbox = tuple(float(c) for c in cells)
    mesh = distorted(o.structured_mesh(cells, bbox, order=order), 0.12, 5)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material(matk, lam, mu, plane_strain=(dim == 2))
    kind = o.ElementKind(dim, order, strain, m, eas_function=fn)
    n = mesh.n_nodes * dim
    rng = np.random.default_rng(9)
    d = 0.03 * rng.uniform(-1, 1, n)
    dev = device_assembler(mesh, kind, mat, np.zeros(n, dtype=bool), layout)
    req = ik.FERequirements(d, 0.0)
    alpha = None
    if m and strain == "gl":
        alpha = 0.01 * rng.uniform(-1, 1, (mesh.n_elem, m))
        dev._check(dev._lib.ikb_eas_set_alpha(dev._h, alpha.ctypes.data))
    pts = np.vstack([np.full(dim, 0.5), rng.uniform(0, 1, (3, dim)), np.zeros(dim), np.ones(dim)])
    u = d[mesh.elem_dofs(layout)].reshape(mesh.n_elem, -1, dim)
    types = [(RT.linearStress, "native"), (RT.linearStressFull, "full")] if strain == "linear" else [
        (RT.PK2Stress, "native"), (RT.PK2StressFull, "full"), (RT.kirchhoffStress, "kirchhoff"), (RT.cauchyStress, "cauchy")]
    for rt, name in types:
        S = dev.calculateAt(rt, req, pts)
        for q, xi in enumerate(pts):
            ref = o.stress_at(kind, mat, mesh.corner_coords, u, xi, alpha=alpha, result=name)
            assert S[:, q].shape == ref.shape, (rt, S.shape, ref.shape)
            assert np.abs(S[:, q] - ref).max() <= 1e-11 * np.abs(ref).max(), (rt, q)


def test_result_function_mirror_reproduces_the_vertex_stress_table():
    """io/resultfunction.hh:57-157: ResultFunction::evaluate(comp, element, local) = fe.calculateAt<RT>(req, local)[comp]
    (through the user function).  On the reference's cube (tests/src/resultcollection.hh:150-163) every vertex value of
    the table comes out of evaluate(); vertexData() is what a vertex-data writer samples; a user function with its own
    ncomps()/name() replaces the component access; a new state of the bound requirement drops the cached tables."""
    mesh = _unit(3)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    d = np.zeros(24)
    d[6:9] = 1.0
    exp = np.array(GOLDEN["cube_vertex_stress"]["values"], float)
    verts = np.array([[(v >> k) & 1 for k in range(3)] for v in range(8)], float)
    dev = device_assembler(mesh, o.ElementKind(3, 1, "linear"), o.Material("linear", lam, mu), np.zeros(24, dtype=bool))
    req = ik.FERequirements(d, 0.0)
    with pytest.raises(Exception):
        ik.makeResultFunction(dev, RT.linearStress).evaluate(0, 0, verts[0])  # unbound assembler (resultfunction.hh:151)
    dev.bind(req, ik.elastoStatics, ik.DBCOption.Full)
    rf = ik.makeResultFunction(dev, RT.linearStress)
    assert rf.ncomps() == 6 and rf.name() == "linearStress"
    for v in range(8):
        for c in range(6):
            assert abs(rf.evaluate(c, 0, verts[v]) - exp[v, c]) <= 1e-7
    assert np.allclose(rf.vertexData()[0], exp, atol=1e-7)

    class VonMises:
        def __call__(self, s, pos, fe, comp):
            return np.sqrt(0.5 * ((s[0] - s[1])**2 + (s[1] - s[2])**2 + (s[2] - s[0])**2) + 3.0 * (s[3]**2 + s[4]**2 + s[5]**2))

        def ncomps(self):
            return 1

        def name(self):
            return "VonMises"

    vm = ik.makeResultFunction(dev, RT.linearStress, VonMises())
    assert vm.ncomps() == 1 and vm.name() == "VonMises"
    s = exp[3]
    ref = np.sqrt(0.5 * ((s[0] - s[1])**2 + (s[1] - s[2])**2 + (s[2] - s[0])**2) + 3.0 * (s[3]**2 + s[4]**2 + s[5]**2))
    assert abs(vm.evaluate(0, 0, verts[3]) - ref) <= 1e-7 * ref
    # the requirement is held by reference (assembler/interface.hh:210-222): changing d changes what evaluate() returns
    d[6:9] = 2.0
    assert abs(rf.evaluate(0, 0, verts[0]) - 2.0 * exp[0, 0]) <= 2e-7
