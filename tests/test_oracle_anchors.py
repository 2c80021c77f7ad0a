"""Pins the CPU oracle against the reference's own known-answer tests."""
import numpy as np
import pytest

import ikarus_oracle as o
from golden_data import GOLDEN
from problems import cantilever, distorted


@pytest.mark.parametrize(
    "dim,mat,m,iters,maxd",
    # tests/src/testcantileverbeamEAS.cpp:26-29, 64-67 (tests/golden/reference_known_answers.json)
    [(c["dim"], c["material"], c["eas"], c["newton_iterations"], c["max_abs_d"])
     for c in GOLDEN["cantilever_eas"]["cases"]],
)
def test_A1_A2_cantilever(dim, mat, m, iters, maxd):
    mesh, kind, material, flags, fext = cantilever(dim, mat, m)
    asm = o.FlatAssembler(mesh, kind, material, flags, fext=fext)
    d, lam, info = o.load_control(asm, np.zeros(asm.n), 20, 0.0, 1.0, tol=1e-10, dbc="full")
    assert info["success"]
    assert info["total_iterations"] == iters
    assert abs(np.abs(d).max() - maxd) < 1e-10
    assert abs(lam - 1.0) < 1e-10


@pytest.mark.parametrize(
    "dim,mat,m,fn,iters,maxd",
    # tests/src/testcantileverbeamEAS.cpp:34-57, 72-95: the displacement-gradient enhancements H4 / H9
    [(c["dim"], c["material"], c["eas"], c["function"], c["newton_iterations"], c["max_abs_d"])
     for c in GOLDEN["cantilever_eas_displacement_gradient"]["cases"]],
)
def test_A1_A2_cantilever_displacement_gradient(dim, mat, m, fn, iters, maxd):
    mesh, kind, material, flags, fext = cantilever(dim, mat, m)
    kind.eas_function = fn
    asm = o.FlatAssembler(mesh, kind, material, flags, fext=fext)
    d, lam, info = o.load_control(asm, np.zeros(asm.n), 20, 0.0, 1.0, tol=1e-10, dbc="full")
    assert info["success"] and info["total_iterations"] == iters
    assert abs(np.abs(d).max() - maxd) < 1e-10


@pytest.mark.parametrize(
    "dim,m,fn,iters,maxd",
    # tests/src/testcantileverbeamEAS.cpp: the Blatz-Ko rows (principal-stretch hyperelastic framework)
    [(c["dim"], c["eas"], c["function"], c["newton_iterations"], c["max_abs_d"])
     for c in GOLDEN["cantilever_eas_blatzko"]["cases"]],
)
def test_A1_A2_cantilever_blatzko(dim, m, fn, iters, maxd):
    """Materials::Hyperelastic<Deviatoric<BlatzKo>, Volumetric<VF0>> (materials/hyperelastic/interface.hh,
    deviatoric/interface.hh, deviatoric/blatzko.hh) under all three enhancement types."""
    mesh, kind, _, flags, fext = cantilever(dim, "neohooke", m)
    kind.eas_function = fn
    mat = o.Material("blatzko", 0.0, 40.0, plane_strain=(dim == 2))
    asm = o.FlatAssembler(mesh, kind, mat, flags, fext=fext)
    d, lam, info = o.load_control(asm, np.zeros(asm.n), 20, 0.0, 1.0, tol=1e-10, dbc="full")
    assert info["success"] and info["total_iterations"] == iters
    assert abs(np.abs(d).max() - maxd) < 1e-10


def test_blatzko_law_is_hyperelastic():
    """S = dpsi/dE and CC = dS/dE by central differences (Voigt, engineering shear), CC symmetric, stress-free reference
    state with the degenerate-eigenvalue branch of deviatoric/interface.hh:98-106."""
    mat = o.Material("blatzko", 0.0, 40.0)
    E = 0.1 * np.random.default_rng(0).uniform(-1, 1, (6, 6))
    psi, S, C = mat.evaluate(E)
    h = 1e-6
    for q in range(6):
        Ep, Em = E.copy(), E.copy()
        Ep[:, q] += h
        Em[:, q] -= h
        pp, Sp, _ = mat.evaluate(Ep)
        pm, Sm, _ = mat.evaluate(Em)
        assert np.abs((pp - pm) / (2 * h) - S[:, q]).max() <= 1e-7 * np.abs(S).max()
        assert np.abs((Sp - Sm) / (2 * h) - C[:, :, q]).max() <= 1e-7 * np.abs(C).max()
    assert np.abs(C - np.swapaxes(C, -1, -2)).max() <= 1e-12 * np.abs(C).max()
    p0, S0, C0 = mat.evaluate(np.zeros((1, 6)))
    assert np.abs(S0).max() == 0.0 and np.allclose(np.diag(C0[0]), [120, 120, 120, 40, 40, 40])


def test_displacement_gradient_tangent_is_the_derivative_of_the_condensed_residual():
    """K = dR/dd for the condensed system when alpha follows d (the static condensation eliminates alpha exactly to
    first order): finite differences on a distorted element with alpha != 0, both variants, 2D and 3D."""
    rng = np.random.default_rng(11)
    for dim, m in ((2, 4), (3, 9)):
        mesh = distorted(o.structured_mesh((1,) * dim, (1.0,) * dim), 0.15, 5)
        lam, mu = o.lame_from_E_nu(100.0, 0.3)
        mat = o.Material("neohooke", lam, mu, plane_strain=(dim == 2))
        X = mesh.corner_coords
        for fn in ("dg", "dgt"):
            kind = o.ElementKind(dim, 1, "gl", m, eas_function=fn)
            u = 0.05 * rng.uniform(-1, 1, (1, kind.nodes, dim))
            alpha = 0.02 * rng.uniform(-1, 1, (1, m))
            q = o.element_quantities(kind, mat, X, u, alpha)
            # generalised (u, alpha) tangent by differences of the UNcondensed residuals R + L^T.. is not exposed; check
            # the blocks instead: D = dRtilde/dalpha, L = dRtilde/du
            h = 1e-6
            Dfd = np.zeros((m, m))
            for p in range(m):
                ap, am = alpha.copy(), alpha.copy()
                ap[0, p] += h
                am[0, p] -= h
                Dfd[:, p] = (o.element_quantities(kind, mat, X, u, ap)["Rtilde"][0] -
                             o.element_quantities(kind, mat, X, u, am)["Rtilde"][0]) / (2 * h)
            assert np.abs(Dfd - q["D"][0]).max() <= 1e-6 * np.abs(q["D"][0]).max(), (dim, fn)
            Lfd = np.zeros((m, kind.ndof))
            for j in range(kind.ndof):
                up, um = u.copy().reshape(1, -1), u.copy().reshape(1, -1)
                up[0, j] += h
                um[0, j] -= h
                Lfd[:, j] = (o.element_quantities(kind, mat, X, up.reshape(u.shape), alpha)["Rtilde"][0] -
                             o.element_quantities(kind, mat, X, um.reshape(u.shape), alpha)["Rtilde"][0]) / (2 * h)
            assert np.abs(Lfd - q["L"][0]).max() <= 1e-6 * np.abs(q["L"][0]).max(), (dim, fn)
            assert np.abs(q["K"][0] - q["K"][0].T).max() == 0.0


def _unit_elem(dim):
    mesh = o.structured_mesh((1,) * dim, (1.0,) * dim)
    return mesh


def test_A3_square_vertex_stress():
    # tests/src/resultcollection.hh:19, 27-51 (plane strain, no EAS and EAS(4))
    mesh = _unit_elem(2)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material("linear", lam, mu, plane_strain=True)
    u = np.array([0, 0, 1, 1, 1, 1, 1, 1.0]).reshape(1, 4, 2)
    verts = [(0, 0), (1, 0), (0, 1), (1, 1)]
    exp = np.array(GOLDEN["square_vertex_stress"]["no_eas"], float)
    kind = o.ElementKind(2, 1, "linear")
    for v, e in zip(verts, exp):
        S = o.stress_at(kind, mat, mesh.corner_coords, u, np.array(v, float))[0]
        assert np.allclose(S, e, atol=1e-7)
    exp4 = np.array(GOLDEN["square_vertex_stress"]["eas4"], float)
    kind4 = o.ElementKind(2, 1, "linear", 4)
    for v, e in zip(verts, exp4):
        S = o.stress_at(kind4, mat, mesh.corner_coords, u, np.array(v, float))[0]
        assert np.allclose(S, e, atol=1e-7)


def test_A3_full_3d_sigma_zz():
    # resultcollection.hh:53-68: the underlying 3D law gives sigma_zz = 1153.84615385 at vertex 0
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat3 = o.Material("linear", lam, mu)
    mesh = _unit_elem(2)
    u = np.array([0, 0, 1, 1, 1, 1, 1, 1.0]).reshape(1, 4, 2)
    kind = o.ElementKind(2, 1, "linear")
    _, dN = o.shape_functions(2, 1, np.zeros(2))
    H = np.einsum("ac,aj->cj", u[0], dN)
    eps = 0.5 * (H + H.T)
    E6 = np.array([eps[0, 0], eps[1, 1], 0, 0, 0, 2 * eps[0, 1]])
    _, S, _ = mat3.evaluate(E6[None])
    g = GOLDEN["square_vertex_stress"]
    v0 = g["no_eas"][0]
    assert np.allclose(S[0], [v0[0], v0[1], g["sigma_zz_vertex0_full_3d_law"], 0, 0, v0[2]], atol=1e-7)


def test_A4_cube_vertex_stress():
    # tests/src/resultcollection.hh:20-21, 150-163
    mesh = _unit_elem(3)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material("linear", lam, mu)
    u = np.zeros(24)
    u[6:9] = 1.0
    u = u.reshape(1, 8, 3)
    exp = np.array(GOLDEN["cube_vertex_stress"]["values"], float)
    kind = o.ElementKind(3, 1, "linear")
    for v in range(8):
        xi = np.array([(v >> k) & 1 for k in range(3)], float)
        S = o.stress_at(kind, mat, mesh.corner_coords, u, xi)[0]
        assert np.allclose(S, exp[v], atol=1e-7), v


@pytest.mark.parametrize("dim", [2, 3])
def test_A5_K_nonlinear_at_zero_equals_linear(dim):
    # tests/src/testnonlinearelasticity.hh:194-246
    mesh = distorted(_unit_elem(dim), 0.2, 7)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    u = np.zeros((1, 2**dim, dim))
    Knl = o.element_quantities(o.ElementKind(dim, 1, "gl"), o.Material("svk", lam, mu, dim == 2), mesh.corner_coords, u)["K"]
    Kl = o.element_quantities(o.ElementKind(dim, 1, "linear"), o.Material("linear", lam, mu, dim == 2), mesh.corner_coords, u)["K"]
    assert np.allclose(Knl, Kl, rtol=0, atol=1e-8)


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("refine", [0, 1, 2])
def test_A6_assembler_invariants(order, refine):
    # tests/src/testassembler.cpp:122-201: YaspGrid 2x1 refined, Q1/Q2, Raw/Full/Reduced
    cells = (2 * 2**refine, 1 * 2**refine)
    mesh = o.structured_mesh(cells, (4.0, 2.0), order=order)
    lam, mu = o.lame_from_E_nu(100.0, 0.2)
    mat = o.Material("linear", lam, mu, plane_strain=True)
    kind = o.ElementKind(2, order, "linear")
    bn = np.unique(np.concatenate([o.boundary_nodes(mesh, a, v) for a, v in [(0, 0.0), (0, 4.0), (1, 0.0), (1, 2.0)]]))
    center = np.nonzero(np.all(np.abs(mesh.node_coords - np.array([2.0, 1.0])) < 1e-9, axis=1))[0]
    flags = o.fix_nodes(mesh, np.concatenate([bn, center]))
    # fixed-dof count formula of testassembler.cpp:145-152
    boundary_nodes = (2 * 2**refine + 1) * 2 + (2**refine + 1) * 2 - 4
    if order == 2:
        boundary_nodes *= 2
    center_node = 0 if (order == 1 and refine == 0) else 1
    assert center.size == center_node
    assert flags.sum() == 2 * (boundary_nodes + center_node)
    asm = o.FlatAssembler(mesh, kind, mat, flags)
    rng = np.random.default_rng(0)
    d = rng.uniform(-1, 1, asm.n) * 0.01
    Kraw = asm.matrix(d, 1.0, "raw").toarray()
    Kfull = asm.matrix(d, 1.0, "full").toarray()
    Kred = asm.matrix(d, 1.0, "reduced").toarray()
    for mode, Ks in (("raw", Kraw), ("full", Kfull), ("reduced", Kred)):
        assert np.allclose(Ks, asm.dense_matrix(d, 1.0, mode), rtol=0, atol=1e-15 * max(1.0, np.abs(Ks).max(initial=0.0)) * 50)
    fx = np.nonzero(flags)[0]
    assert np.all(Kfull[fx, fx] == 1.0)
    Kf2 = Kfull.copy()
    Kf2[fx, fx] = 0.0
    assert np.all(Kf2[fx, :] == 0) and np.all(Kf2[:, fx] == 0)
    fr = ~flags
    assert np.allclose(Kred, Kraw[np.ix_(fr, fr)], rtol=0, atol=1e-12)
    Rfull = asm.vector(d, 1.0, "full")
    Rraw = asm.vector(d, 1.0, "raw")
    Rred = asm.vector(d, 1.0, "reduced")
    assert np.all(Rfull[fx] == 0) and np.allclose(Rfull[fr], Rraw[fr]) and np.allclose(Rred, Rraw[fr])
    assert np.allclose(asm.create_reduced_vector(asm.create_full_vector(Rred)), Rred)
    # pattern of Full == pattern of Raw (simpleassemblers.inl:159-167 keeps the pattern)
    o1, i1 = asm.pattern("raw")
    assert o1[-1] == i1.shape[0]


@pytest.mark.parametrize("dim,m", [(2, 4), (2, 5), (2, 7), (3, 9), (3, 21)])
def test_A7_eas_properties(dim, m):
    # tests/src/testeas.hh:50-85, testnonlineareas.cpp:95-189
    mesh = distorted(_unit_elem(dim), 0.15, 11)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    kind = o.ElementKind(dim, 1, "gl", m)
    X = mesh.corner_coords
    # int M dV = 0
    pts, wts = kind.rule()
    Jt0, _, detJ0 = o._geometry(kind, X, np.full(dim, 0.5))
    T0inv = np.linalg.inv(o.transformation_matrix(Jt0) * detJ0[:, None, None])
    acc = 0
    for xi, w in zip(pts, wts):
        _, _, detJ = o._geometry(kind, X, xi)
        M = T0inv[0] @ o.eas_Mhat(dim, m, xi) / detJ[0]
        acc = acc + M * detJ[0] * w
    assert np.abs(acc).max() < 1e-13
    mat = o.Material("svk", lam, mu, dim == 2)
    u0 = np.zeros((1, 2**dim, dim))
    K = o.element_quantities(kind, mat, X, u0)["K"][0]
    assert np.allclose(K, K.T, rtol=0, atol=1e-9)
    ev = np.linalg.eigvalsh(K)
    nrb = 3 * dim - 3
    assert np.sum(np.abs(ev) < 1e-8 * np.abs(ev).max()) == nrb
    # eas(0) == plain element
    rng = np.random.default_rng(1)
    u = 0.05 * rng.uniform(-1, 1, u0.shape)
    q0 = o.element_quantities(o.ElementKind(dim, 1, "gl", 0), mat, X, u)
    matn = o.Material("neohooke", lam, mu, dim == 2)
    qn = o.element_quantities(o.ElementKind(dim, 1, "gl", 0), matn, X, u)
    assert np.allclose(qn["K"], np.swapaxes(qn["K"], 1, 2), atol=1e-9)
    assert np.isfinite(q0["E"]).all()


@pytest.mark.parametrize("dim,order,matk", [(3, 1, "neohooke"), (3, 1, "svk"), (2, 1, "neohooke"), (2, 1, "svk"),
                                             (3, 2, "svk"), (2, 2, "neohooke")])
def test_A8_fd_consistency(dim, order, matk):
    # R = dE/dd, K = dR/dd  (tests/src/testnonlinearelasticity.hh:340-369, checkfebyautodiff.hh)
    mesh = distorted(o.structured_mesh((1,) * dim, (1.0,) * dim, order=order), 0.1, 5)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material(matk, lam, mu, dim == 2)
    kind = o.ElementKind(dim, order, "gl")
    rng = np.random.default_rng(3)
    u = 0.01 * rng.uniform(-1, 1, (1, kind.nodes, dim))
    q = o.element_quantities(kind, mat, mesh.corner_coords, u)
    h = 1e-6
    nd = kind.ndof
    Rfd = np.zeros(nd)
    Kfd = np.zeros((nd, nd))
    for i in range(nd):
        up = u.copy().reshape(1, nd)
        um = up.copy()
        up[0, i] += h
        um[0, i] -= h
        qp = o.element_quantities(kind, mat, mesh.corner_coords, up.reshape(u.shape))
        qm = o.element_quantities(kind, mat, mesh.corner_coords, um.reshape(u.shape))
        Rfd[i] = (qp["E"][0] - qm["E"][0]) / (2 * h)
        Kfd[:, i] = (qp["R"][0] - qm["R"][0]) / (2 * h)
    assert np.allclose(q["R"][0], Rfd, rtol=1e-6, atol=1e-6 * np.abs(q["R"]).max())
    assert np.allclose(q["K"][0], Kfd, rtol=1e-6, atol=1e-6 * np.abs(q["K"]).max())


@pytest.mark.parametrize("dim,m,matk", [(3, 9, "neohooke"), (3, 21, "svk"), (2, 4, "neohooke"), (2, 7, "svk")])
def test_eas_condensed_tangent_fd(dim, m, matk):
    """Condensed K equals d/du of the condensed residual with alpha following the
    stationarity condition (the twin recipe of SURVEY.md §8c: K = Kuu - Kua Kaa^-1 Kau)."""
    mesh = distorted(_unit_elem(dim), 0.1, 9)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material(matk, lam, mu, dim == 2)
    kind = o.ElementKind(dim, 1, "gl", m)
    rng = np.random.default_rng(4)
    X = mesh.corner_coords
    u = 0.02 * rng.uniform(-1, 1, (1, kind.nodes, dim))

    def solve_alpha(uu):
        a = np.zeros((1, m))
        for _ in range(30):
            q = o.element_quantities(kind, mat, X, uu, a, want=())
            da = np.linalg.solve(q["D"], q["Rtilde"][..., None])[..., 0]
            a = a - da
            if np.abs(da).max() < 1e-14:
                break
        return a

    a = solve_alpha(u)
    q = o.element_quantities(kind, mat, X, u, a)
    assert np.abs(q["Rtilde"]).max() < 1e-9
    nd = kind.ndof
    h = 1e-6
    Kfd = np.zeros((nd, nd))
    for i in range(nd):
        up = u.copy().reshape(1, nd)
        um = up.copy()
        up[0, i] += h
        um[0, i] -= h
        qp = o.element_quantities(kind, mat, X, up.reshape(u.shape), solve_alpha(up.reshape(u.shape)))
        qm = o.element_quantities(kind, mat, X, um.reshape(u.shape), solve_alpha(um.reshape(u.shape)))
        Kfd[:, i] = (qp["R"][0] - qm["R"][0]) / (2 * h)
    assert np.allclose(q["K"][0], Kfd, rtol=1e-5, atol=1e-5 * np.abs(q["K"]).max())


def test_q1_numbering_matches_python_reference_test():
    # ikarus/python/test/linearelastictest.py:213-231: 3x3 grid, node (i,j) -> i + 4j; lexicographic dofs +16
    mesh = o.structured_mesh((3, 3), (1.0, 1.0))
    ij = np.rint(mesh.node_coords * 3).astype(int)
    assert np.all(ij[:, 0] + 4 * ij[:, 1] == np.arange(16))
    ed = mesh.elem_dofs("lexicographic")
    assert ed.max() == 31 and np.all(ed[0] == [0, 16, 1, 17, 4, 20, 5, 21])
    assert np.all(mesh.elem_dofs("interleaved")[0] == [0, 1, 2, 3, 8, 9, 10, 11])
