"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs."""
import numpy as np
import pytest

import ikarus_b200 as ik
import ikarus_oracle as o
from devproblems import device_assembler, entry_error
from problems import cantilever, distorted

pytestmark = pytest.mark.gpu
TOL = 1e-12  # BASELINE.json north_star: matrix and residual entries within 1e-12 relative in FP64

CASES = [
    # dim, cells, order, strain, material
    (3, (4, 3, 2), 1, "gl", "neohooke"),
    (3, (3, 2, 2), 1, "gl", "svk"),
    (3, (3, 2, 2), 1, "linear", "linear"),
    (2, (5, 4), 1, "gl", "neohooke"),
    (2, (5, 4), 1, "gl", "svk"),
    (2, (6, 3), 1, "linear", "linear"),
    # Q2 (Quad9 / Hex27, 81-dof elements)
    (3, (2, 2, 2), 2, "gl", "svk"),
    (3, (2, 2, 1), 2, "gl", "neohooke"),
    (3, (2, 1, 1), 2, "linear", "linear"),
    (2, (3, 2), 2, "gl", "neohooke"),
    (2, (3, 2), 2, "gl", "svk"),
    (2, (3, 2), 2, "linear", "linear"),
]


def _setup(dim, cells, order, strain, matk, layout="interleaved", distort=0.15, seed=0):
    bbox = tuple(float(c) for c in cells)
    mesh = distorted(o.structured_mesh(cells, bbox, order=order), distort, seed + 1) if distort else o.structured_mesh(
        cells, bbox, order=order)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material(matk, lam, mu, plane_strain=(dim == 2))
    kind = o.ElementKind(dim, order, strain)
    flags = o.fix_nodes(mesh, o.boundary_nodes(o.structured_mesh(cells, bbox, order=order), 0, 0.0), layout)
    # additionally fix single components to exercise per-component flags
    flags[-1] = True
    flags[len(flags) // 2] = True
    rng = np.random.default_rng(seed)
    fext = rng.uniform(-1, 1, flags.shape[0])
    d = 0.05 * rng.uniform(-1, 1, flags.shape[0])
    ref = o.FlatAssembler(mesh, kind, mat, flags, layout, fext=fext)
    dev = device_assembler(mesh, kind, mat, flags, layout, fext=fext)
    return mesh, ref, dev, d


@pytest.mark.parametrize("layout", ["interleaved", "lexicographic"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-Q{c[2]}-{c[3]}-{c[4]}")
def test_pattern_bit_exact_and_values(case, layout):
    mesh, ref, dev, d = _setup(*case, layout=layout)
    lam = 0.7
    req = ik.FERequirements(d, lam)
    for mode, dbc in (("raw", ik.DBCOption.Raw), ("full", ik.DBCOption.Full), ("reduced", ik.DBCOption.Reduced)):
        outer, inner = ref.pattern(mode)
        douter, dinner = dev.pattern(dbc)
        assert douter.dtype == np.int64 and dinner.dtype == np.int32
        assert np.array_equal(outer, douter), f"outer index differs ({mode})"
        assert np.array_equal(inner, dinner), f"inner index differs ({mode})"
        K = dev.matrix(req, ik.MatrixAffordance.stiffness, dbc)
        vals = ref.matrix_values(d, lam, mode)
        rows = np.repeat(np.arange(outer.shape[0] - 1), np.diff(outer))
        assert entry_error(K.data, vals, rows) <= TOL, mode
        R = dev.vector(req, ik.VectorAffordance.forces, dbc)
        Rref = ref.vector(d, lam, mode)
        assert R.shape == Rref.shape
        assert np.abs(R - Rref).max(initial=0.0) <= TOL * np.abs(Rref).max(initial=1.0), mode
        if mode == "full":
            fx = np.nonzero(ref.flags)[0]
            Kd = K.toarray()
            assert np.all(Kd[fx, fx] == 1.0) and np.all(R[fx] == 0.0)
            Kd[fx, fx] = 0.0
            assert np.all(Kd[fx, :] == 0.0) and np.all(Kd[:, fx] == 0.0)
    if case[3] != "eas":
        E = dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
        Eref = ref.scalar(d, lam)
        assert abs(E - Eref) <= 1e-12 * max(1.0, abs(Eref))
    # reduced-dof index maps bit exact
    assert [dev.constraintsBelow(i) for i in range(dev.size())] == list(ref.cb)
    assert dev.reducedSize() == ref.n_red


def test_constraints_below_on_device_bit_exact():
    import ctypes as C

    from ikarus_b200 import _capi as capi

    mesh, ref, dev, d = _setup(*CASES[0])
    out = np.zeros(dev.size(), dtype=np.int64)
    dev._check(dev._lib.ikb_get_constraints_below(dev._h, capi.ptr(out)))
    assert np.array_equal(out, ref.cb)


def test_element_linear_indices_match_reference_order():
    mesh, ref, dev, d = _setup(*CASES[0])
    outer, inner = ref.pattern("raw")
    nd = ref.kind.ndof
    for e in (0, mesh.n_elem // 2, mesh.n_elem - 1):
        # reference order: for c in dofs, for r in dofs (simpleassemblers.inl:253-266)
        pos = o.linear_indices(ref.elem_dofs[e:e + 1], outer, inner)[0].reshape(nd, nd).T.ravel()
        assert np.array_equal(dev.elementLinearIndices(e), pos)


@pytest.mark.parametrize("case", [CASES[3], CASES[5], CASES[0]], ids=lambda c: f"{c[0]}d-{c[4]}")
def test_sparse_equals_dense(case):
    # tests/src/testassembler.cpp:31-38, 122-183
    mesh, ref, dev, d = _setup(*case)
    dense = device_assembler(mesh, ref.kind, ref.mat, ref.flags, dense=True, fext=ref.fext)
    req = ik.FERequirements(d, 0.3)
    for dbc in (ik.DBCOption.Raw, ik.DBCOption.Full, ik.DBCOption.Reduced):
        Ks = dev.matrix(req, ik.MatrixAffordance.stiffness, dbc).toarray()
        Kd = dense.matrix(req, ik.MatrixAffordance.stiffness, dbc)
        assert Ks.shape == Kd.shape
        assert np.array_equal(Ks, Kd)  # same kernels, same summation order: bit identical
        assert np.array_equal(dev.vector(req, ik.VectorAffordance.forces, dbc),
                              dense.vector(req, ik.VectorAffordance.forces, dbc))
    assert dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy) == dense.scalar(
        req, ik.ScalarAffordance.mechanicalPotentialEnergy)


def test_run_to_run_determinism():
    mesh, ref, dev, d = _setup(*CASES[0])
    req = ik.FERequirements(d, 0.3)
    a = dev.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Full).data.copy()
    dev2 = device_assembler(mesh, ref.kind, ref.mat, ref.flags, fext=ref.fext)
    b = dev2.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Full).data
    assert np.array_equal(a, b)


def test_bind_and_unbound_errors():
    # assembler/interface.hh:150-193, tests/src/testassembler.cpp:289-311
    mesh, ref, dev, d = _setup(*CASES[3])
    with pytest.raises(ik.assembler.InvalidStateException):
        dev.requirement()
    with pytest.raises(ik.assembler.InvalidStateException):
        dev.matrix()
    req = ik.FERequirements(d, 0.1)
    dev.bind(req, ik.elastoStatics, ik.DBCOption.Reduced)
    assert dev.bound() and dev.requirement() is req and dev.dBCOption() == ik.DBCOption.Reduced
    R = dev.vector()
    assert R.shape[0] == dev.reducedSize()
    assert np.allclose(dev.createReducedVector(dev.createFullVector(R)), R)
    K = dev.matrix()
    assert K.shape == (dev.reducedSize(), dev.reducedSize())
    # requirement is held by reference: mutate in place and re-evaluate
    req.globalSolution()[:] = 0.0
    R0 = dev.vector()
    assert np.abs(R0 - ref.vector(np.zeros_like(d), 0.1, "reduced")).max() <= TOL * max(1.0, np.abs(R0).max())


def test_empty_reduced_system():
    # every node fixed: Reduced matrices are 0x0 (YaspGrid 2x1 Q1 of testassembler.cpp ref=0)
    mesh = o.structured_mesh((2, 1), (4.0, 2.0))
    lam, mu = o.lame_from_E_nu(100.0, 0.2)
    mat = o.Material("linear", lam, mu, True)
    kind = o.ElementKind(2, 1, "linear")
    flags = np.ones(mesh.n_nodes * 2, dtype=bool)
    dev = device_assembler(mesh, kind, mat, flags)
    req = ik.FERequirements(np.zeros(flags.shape[0]), 0.0)
    K = dev.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Reduced)
    assert K.shape == (0, 0)
    assert dev.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Reduced).shape == (0,)
    Kf = dev.matrix(req, ik.MatrixAffordance.stiffness, ik.DBCOption.Full).toarray()
    assert np.array_equal(Kf, np.eye(flags.shape[0]))


def test_material_failure_is_reported_not_aborted():
    mesh, ref, dev, d = _setup(3, (2, 2, 2), 1, "gl", "neohooke", distort=0.0)
    bad = np.zeros_like(d)
    X = mesh.node_coords
    bad[0::3] = -2.0 * X[:, 0]  # F = diag(-1,1,1): det C > 0 but inverted... use a collapse instead
    bad[0::3] = -1.0 * X[:, 0]  # F_11 = 0 -> det C = 0
    req = ik.FERequirements(bad, 0.0)
    with pytest.raises(ik.assembler.MaterialError):
        dev.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Raw)
    # the handle stays usable
    R = dev.vector(ik.FERequirements(np.zeros_like(d), 0.0), ik.VectorAffordance.forces, ik.DBCOption.Raw)
    assert np.isfinite(R).all()


@pytest.mark.parametrize("dbc", [ik.DBCOption.Full, ik.DBCOption.Reduced])
def test_spmv_and_pcg(dbc):
    mesh, ref, dev, d = _setup(3, (6, 4, 3), 1, "gl", "neohooke")
    dev = device_assembler(mesh, ref.kind, ref.mat, ref.flags, fext=ref.fext, mode="resident")
    req = ik.FERequirements(0.2 * d, 0.5)
    A = dev.matrix(req, ik.MatrixAffordance.stiffness, dbc)
    assert isinstance(A, ik.DeviceMatrix)
    Ah = A.to_scipy()
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, Ah.shape[0])
    y = A.matvec(x)
    yr = Ah @ x
    assert np.abs(y - yr).max() <= 1e-13 * np.abs(yr).max()
    b = rng.uniform(-1, 1, Ah.shape[0])
    if dbc == ik.DBCOption.Full:
        b[ref.flags] = 0.0
    ls = ik.DeviceLinearSolver(relTol=1e-13)
    sol = ls(b, A)
    import scipy.sparse.linalg as spla
    xr = spla.spsolve(Ah.tocsc(), b)
    assert ls.lastIterations > 0 and ls.lastRelRes <= 1e-13
    assert np.abs(sol - xr).max() <= 1e-9 * np.abs(xr).max()


@pytest.mark.parametrize("matk,dbc", [("neohooke", ik.DBCOption.Full), ("svk", ik.DBCOption.Reduced)])
def test_newton_loadcontrol_iteration_counts_match_oracle(matk, dbc):
    """Identical Newton iteration counts and load-displacement curve (north_star), displacement-based Hex8
    cantilever with the reference's point loads; device PCG vs the oracle's direct solver."""
    mesh, kind, mat, flags, fext = cantilever(3, matk, 0, cells=(10, 2, 2))
    ref = o.FlatAssembler(mesh, kind, mat, flags, fext=fext)
    mode = {ik.DBCOption.Full: "full", ik.DBCOption.Reduced: "reduced"}[dbc]
    dr, lamr, inf = o.load_control(ref, np.zeros(ref.n), 5, 0.0, 1.0, tol=1e-8, dbc=mode)
    dev = device_assembler(mesh, kind, mat, flags, fext=fext, mode="resident")
    req = ik.FERequirements(np.zeros(ref.n), 0.0)
    dev.bind(req, ik.elastoStatics, dbc)
    nr = ik.NewtonRaphson(dev, ik.NewtonRaphsonConfig(ik.NRSettings(tol=1e-8), ik.DeviceLinearSolver(1e-13)))
    lc = ik.LoadControl(nr, ik.LoadControlConfig(5, 0.0, 1.0))
    curve = []
    lc.listeners.append(lambda ls, r: curve.append((r.parameter(), float(np.abs(r.globalSolution()).max()))))
    info = lc.run(req)
    assert info.success and inf["success"]
    assert [s.iterations for s in info.solverInfos] == inf["per_step"]
    assert info.totalIterations == inf["total_iterations"]
    for (l1, m1), (l2, m2) in zip(curve, inf["curve"][1:]):
        assert abs(l1 - l2) < 1e-14 and abs(m1 - m2) <= 1e-8
    assert np.abs(req.globalSolution() - dr).max() <= 1e-8


def test_volume_load_sampling_matches_oracle_consistent_load():
    mesh, ref, dev, d = _setup(3, (3, 2, 2), 1, "gl", "svk")
    q = np.array([0.0, 0.0, -2.5])
    dev = device_assembler(mesh, ref.kind, ref.mat, ref.flags, volume=lambda x, lam: lam * q)
    # oracle consistent load: integrate N_a q over each element
    fext = np.zeros(ref.n)
    pts, wts = ref.kind.rule()
    for xi, w in zip(pts, wts):
        N, _ = o.shape_functions(3, 1, xi)
        _, _, detJ = o._geometry(ref.kind, mesh.corner_coords, xi)
        c = (N[None, :, None] * q[None, None, :] * (detJ * w)[:, None, None]).reshape(mesh.n_elem, -1)
        np.add.at(fext, ref.elem_dofs.ravel(), c.ravel())
    ref2 = o.FlatAssembler(mesh, ref.kind, ref.mat, ref.flags, fext=fext)
    req = ik.FERequirements(d, 1.7)
    R = dev.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Full)
    Rr = ref2.vector(d, 1.7, "full")
    assert np.abs(R - Rr).max() <= TOL * np.abs(Rr).max()
    E = dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
    assert abs(E - ref2.scalar(d, 1.7)) <= 1e-12 * abs(ref2.scalar(d, 1.7))


def test_pipelined_solution_upload_is_bit_identical():
    """ikb_set_solution_range on a large interleaved Q1 mesh copies d in pieces on the side stream and launches the
    element kernel chunk by chunk; K, R must be bit-identical to the plain ikb_set_solution path, also for sub-ranges
    and when another consumer of the solution (ikb_get_solution) comes between upload and assembly."""
    import ctypes as C
    from ikarus_b200 import _capi as capi
    cells = (40, 32, 32)  # 40960 elements: above the pipelining threshold
    mesh = distorted(o.structured_mesh(cells, (40.0, 32.0, 32.0), order=1), 0.1, 3)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material("neohooke", lam, mu)
    kind = o.ElementKind(3, 1, "gl")
    flags = o.fix_nodes(mesh, o.boundary_nodes(o.structured_mesh(cells, (40.0, 32.0, 32.0), order=1), 0, 0.0), "interleaved")
    dev = device_assembler(mesh, kind, mat, flags, "interleaved")
    rng = np.random.default_rng(5)
    n = flags.shape[0]
    d = 0.02 * rng.uniform(-1, 1, n)
    dev.bind(ik.FERequirements(d, 0.0), ik.elastoStatics, ik.DBCOption.Full)
    r0 = dev.vector().copy()
    v0 = dev.matrix().data.copy()
    lib, h = dev._lib, dev._h
    # whole range through the pipelined path, from a different starting state
    dev._check(lib.ikb_set_solution(h, capi.ptr(np.zeros(n))))
    dev._check(lib.ikb_assemble(h, 6, int(ik.DBCOption.Full)))
    dev._check(lib.ikb_set_solution_range(h, capi.ptr(d), 0, n))
    dev._check(lib.ikb_assemble(h, 6, int(ik.DBCOption.Full)))
    r1 = np.empty(n)
    v1 = np.empty_like(v0)
    dev._check(lib.ikb_get_vector(h, int(ik.DBCOption.Full), capi.ptr(r1)))
    dev._check(lib.ikb_get_matrix_values(h, int(ik.DBCOption.Full), capi.ptr(v1)))
    assert np.array_equal(r0, r1) and np.array_equal(v0, v1)
    # two sub-ranges (second one pipelined again), read back before assembling
    dev._check(lib.ikb_set_solution(h, capi.ptr(np.zeros(n))))
    cut = 3 * 20001
    dev._check(lib.ikb_set_solution_range(h, capi.ptr(d), 0, cut))
    dev._check(lib.ikb_set_solution_range(h, capi.ptr(d[cut:].copy()), cut, n - cut))
    back = np.empty(n)
    dev._check(lib.ikb_get_solution(h, capi.ptr(back)))
    assert np.array_equal(back, d)
    dev._check(lib.ikb_assemble(h, 6, int(ik.DBCOption.Full)))
    dev._check(lib.ikb_get_vector(h, int(ik.DBCOption.Full), capi.ptr(r1)))
    dev._check(lib.ikb_get_matrix_values(h, int(ik.DBCOption.Full), capi.ptr(v1)))
    assert np.array_equal(r0, r1) and np.array_equal(v0, v1)
    # Interleaved sweep (IKB_CHUNKS=n, opt-in): element kernel chunk by chunk, the rows completed by a chunk gathered on a
    # side stream beside the next chunk -- same kernels, same staged values: same bits, also behind a pipelined upload
    import os
    os.environ["IKB_CHUNKS"] = "5"
    try:
        chunked = device_assembler(mesh, kind, mat, flags, "interleaved")
    finally:
        del os.environ["IKB_CHUNKS"]
    lib2, h2 = chunked._lib, chunked._h
    for upload in (lib2.ikb_set_solution, None):
        chunked._check(lib2.ikb_set_solution(h2, capi.ptr(np.zeros(n))))
        chunked._check(lib2.ikb_assemble(h2, 6, int(ik.DBCOption.Full)))
        if upload:
            chunked._check(upload(h2, capi.ptr(d)))
        else:
            chunked._check(lib2.ikb_set_solution_range(h2, capi.ptr(d), 0, n))
        chunked._check(lib2.ikb_assemble(h2, 6, int(ik.DBCOption.Full)))
        chunked._check(lib2.ikb_get_vector(h2, int(ik.DBCOption.Full), capi.ptr(r1)))
        chunked._check(lib2.ikb_get_matrix_values(h2, int(ik.DBCOption.Full), capi.ptr(v1)))
        assert np.array_equal(r0, r1) and np.array_equal(v0, v1)
    n0 = chunked.launchCount()
    chunked._check(lib2.ikb_invalidate(h2))
    chunked._check(lib2.ikb_assemble(h2, 6, int(ik.DBCOption.Full)))
    assert chunked.launchCount() - n0 >= 7  # the chunked path really ran: 5 element launches, up to 5 + 1 gathers


UNSTRUCTURED = [
    # dim, cells, order, strain, material
    (3, (5, 4, 3), 1, "gl", "neohooke"),
    (2, (7, 6), 1, "gl", "svk"),
    (2, (4, 4), 2, "gl", "neohooke"),
    (3, (3, 2, 2), 2, "linear", "linear"),
]


@pytest.mark.parametrize("layout", ["interleaved", "lexicographic"])
@pytest.mark.parametrize("case", UNSTRUCTURED, ids=lambda c: f"{c[0]}d-Q{c[2]}-{c[4]}")
def test_unstructured_numbering_and_connectivity(case, layout):
    """Random node numbering, shuffled elements and an L-shaped domain with a hole: pattern bit exact and values
    within 1e-12 in all three Dirichlet modes."""
    from problems import unstructured
    dim, cells, order, strain, matk = case
    bbox = tuple(float(c) for c in cells)
    base = distorted(o.structured_mesh(cells, bbox, order=order), 0.1, 11)

    def drop(c):
        corner = all(c[k] > 0.5 * bbox[k] for k in range(dim))  # removes one corner block -> L / notched shape
        hole = all(abs(c[k] - 1.5) < 0.6 for k in range(dim))   # one interior cell
        return corner or hole

    mesh = unstructured(base, 7, drop)
    assert mesh.n_elem < base.n_elem
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material(matk, lam, mu, plane_strain=(dim == 2))
    kind = o.ElementKind(dim, order, strain)
    rng = np.random.default_rng(3)
    n = mesh.n_nodes * dim
    flags = rng.uniform(size=n) < 0.15
    fext = rng.uniform(-1, 1, n)
    d = 0.03 * rng.uniform(-1, 1, n)
    ref = o.FlatAssembler(mesh, kind, mat, flags, layout, fext=fext)
    dev = device_assembler(mesh, kind, mat, flags, layout, fext=fext)
    req = ik.FERequirements(d, 0.4)
    for mode, dbc in (("raw", ik.DBCOption.Raw), ("full", ik.DBCOption.Full), ("reduced", ik.DBCOption.Reduced)):
        outer, inner = ref.pattern(mode)
        douter, dinner = dev.pattern(dbc)
        assert np.array_equal(outer, douter) and np.array_equal(inner, dinner), mode
        K = dev.matrix(req, ik.MatrixAffordance.stiffness, dbc)
        rows = np.repeat(np.arange(outer.shape[0] - 1), np.diff(outer))
        assert entry_error(K.data, ref.matrix_values(d, 0.4, mode), rows) <= TOL, mode
        R = dev.vector(req, ik.VectorAffordance.forces, dbc)
        Rref = ref.vector(d, 0.4, mode)
        assert np.abs(R - Rref).max(initial=0.0) <= TOL * np.abs(Rref).max(initial=1.0), mode
    E = dev.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy)
    Eref = ref.scalar(d, 0.4)
    assert abs(E - Eref) <= 1e-12 * max(1.0, abs(Eref))
    assert [dev.constraintsBelow(i) for i in range(dev.size())] == list(ref.cb)


@pytest.mark.parametrize("case", [CASES[0], CASES[3], CASES[6], CASES[9]], ids=lambda c: f"{c[0]}d-Q{c[2]}")
@pytest.mark.parametrize("layout", ["interleaved", "lexicographic"])
def test_gather_variants_bit_identical(case, layout, monkeypatch):
    """The matrix gather has two kernels (pull: per-block contribution lists, default; tile: warp tile) and the pull
    kernel has rarely taken paths (contribution lists too long to stage, 64-bit staged offsets).  All of them add the same staged values in the same order, so every mode must give identical bits."""
    mesh, ref, dev, d = _setup(*case, layout=layout)
    req = ik.FERequirements(d, 0.4)
    modes = (ik.DBCOption.Raw, ik.DBCOption.Full, ik.DBCOption.Reduced)
    base = [dev.matrix(req, ik.MatrixAffordance.stiffness, m).data.copy() for m in modes]
    # (IKB_PULL_ASYNC=1: the row-pipelined cp.async form for Q1 kinds in Raw/Full mode)
    variants = [{"IKB_PULL_ASYNC": "1"}, {"IKB_GATHER": "tile"}, {"IKB_PULL_STAGE_MAX": "0"}, {"IKB_PULL_STAGE_MAX": "5"}, {"IKB_PULL_IDX64": "1"},
                {"IKB_PULL_IDX64": "1", "IKB_PULL_STAGE_MAX": "7"}]
    for env in variants:
        with monkeypatch.context() as mp:
            for k, v in env.items():
                mp.setenv(k, v)
            other = device_assembler(mesh, ref.kind, ref.mat, ref.flags, layout, fext=ref.fext)
        for m, a in zip(modes, base):
            b = other.matrix(req, ik.MatrixAffordance.stiffness, m).data
            assert np.array_equal(a, b), (env, m)
    # Pipelined sweep (IKB_FUSED=1; serves Hex8 NeoHooke/LinearElastic, other kinds fall through to the same path as
    # above): producer and consumer kernel side by side, K_e through a ring.  Same staged values added in the same order
    # => the same bits, also with a tiny ring (wrap-around with write-after-read guards) and over repeated sweeps
    # (one epoch of the completion counters per launch).
    for ring in (None, "0.01"):
        with monkeypatch.context() as mp:
            mp.setenv("IKB_FUSED", "1")
            if ring:
                mp.setenv("IKB_RING_MB", ring)
            sweep = device_assembler(mesh, ref.kind, ref.mat, ref.flags, layout, fext=ref.fext)
        for m, a in zip(modes, base):
            for _ in range(2):
                sweep.matrix(ik.FERequirements(d, 0.5), ik.MatrixAffordance.stiffness, m)
                assert np.array_equal(sweep.matrix(req, ik.MatrixAffordance.stiffness, m).data, a), (ring, m)
        R0 = dev.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Full)
        assert np.array_equal(sweep.vector(req, ik.VectorAffordance.forces, ik.DBCOption.Full), R0), ring


def test_pipelined_sweep_ring_guards_bit_identical(monkeypatch):
    """IKB_FUSED=1 on a mesh large enough for the ring to wrap (24x8x8 Hex8, 384 producer tickets in 6 completion groups):
    the smallest legal ring (IKB_SWEEP_MARGIN=0) forces write-after-read waits between the two kernels on every lap.
    Values must equal the back-to-back path bit for bit in all three Dirichlet modes, sweep after sweep."""
    mesh = o.structured_mesh((24, 8, 8), (3.0, 1.0, 1.0))
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material("neohooke", lam, mu)
    kind = o.ElementKind(3, 1, "gl")
    flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 0, 0.0))
    d = 0.01 * np.random.default_rng(7).uniform(-1, 1, flags.shape[0])
    d[flags] = 0.0
    req = ik.FERequirements(d, 0.0)
    dev = device_assembler(mesh, kind, mat, flags)
    with monkeypatch.context() as mp:
        mp.setenv("IKB_FUSED", "1")
        mp.setenv("IKB_RING_MB", "0.01")
        mp.setenv("IKB_SWEEP_MARGIN", "0")
        sweep = device_assembler(mesh, kind, mat, flags)
    for m in (ik.DBCOption.Raw, ik.DBCOption.Full, ik.DBCOption.Reduced):
        a = dev.matrix(req, ik.MatrixAffordance.stiffness, m).data
        ra = dev.vector(req, ik.VectorAffordance.forces, m)
        for _ in range(3):
            sweep.matrix(ik.FERequirements(d, 0.5), ik.MatrixAffordance.stiffness, m)
            assert np.array_equal(sweep.matrix(req, ik.MatrixAffordance.stiffness, m).data, a), m
            assert np.array_equal(sweep.vector(req, ik.VectorAffordance.forces, m), ra), m
    assert sweep.scalar(req, ik.ScalarAffordance.mechanicalPotentialEnergy) == dev.scalar(
        req, ik.ScalarAffordance.mechanicalPotentialEnergy)
