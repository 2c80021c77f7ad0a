"""planeStress (VanishingStress) in the oracle against the reference's known answers:
B2 single distorted Quad4 eigenvalues (tests/src/testnonlinearelasticity.hh:248-338),
the linear patch test with inhomogeneous Dirichlet values (tests/src/testinhomogeneousdbc.cpp:46-160).
(Anchor B1 of SURVEY 8c -- the 10x10 square under a volume load, tests/src/testnonlinearelasticity.hh:47-150 -- is NOT
reproduced: the restatement converges to energy -2.96032057 / max d 0.11292927 against the reference's -2.96051876 /
0.11293260, a 3e-5 relative difference in the load-stiffness ratio whose origin could not be determined without running
the reference; it is therefore not used as a pin.)"""
import numpy as np
import scipy.sparse.linalg as spla

import ikarus_oracle as o
from golden_data import GOLDEN
from problems import PATCH_EXPECTED_D, fixed_distorted_quad, patch_test_mesh


def test_B2_single_element_eigenvalues():
    mesh = fixed_distorted_quad()
    lam, mu = o.lame_from_E_nu(1000.0, 0.0)
    mat = o.Material("svk", lam, mu, plane_stress=True, ps_tol=1e-8)
    g = GOLDEN["plane_stress_single_element_eigenvalues"]
    d = np.array(g["d"], float)
    K = o.element_quantities(o.ElementKind(2, 1, "gl"), mat, mesh.corner_coords, d.reshape(1, 4, 2))["K"][0]
    ev = np.abs(np.linalg.eigvalsh(K))
    exp = np.array(g["abs_eigenvalues"], float)
    assert np.abs(np.sort(ev) - exp).max() < 1e-8


def test_plane_stress_laws_reduce_to_zero_normal_stress():
    rng = np.random.default_rng(0)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    Ev = 0.1 * rng.uniform(-1, 1, (5, 3))
    for kind in ("linear", "svk", "neohooke"):
        m = o.Material(kind, lam, mu, plane_stress=True, ps_tol=1e-12)
        psi, S, C = m.evaluate(Ev)
        E6 = m._reduce_stress(Ev)
        _, S6, _ = m._law3d(E6)
        assert np.abs(S6[:, 2:5]).max() < 1e-11 and np.allclose(S6[:, [0, 1, 5]], S)
        # condensed tangent = dS/dE of the reduced law (central differences)
        h = 1e-6
        for q in range(3):
            dE = np.zeros(3)
            dE[q] = h
            fd = (m.evaluate(Ev + dE)[1] - m.evaluate(Ev - dE)[1]) / (2 * h)
            assert np.abs(fd - C[:, :, q]).max() < 1e-5 * np.abs(C).max()


def test_linear_patch_test_with_inhomogeneous_dirichlet_values():
    mesh = patch_test_mesh()
    lam, mu = o.lame_from_E_nu(1000.0, 0.25)
    mat = o.Material("linear", lam, mu, plane_stress=True)
    kind = o.ElementKind(2, 1, "linear")
    flags = np.zeros(16, dtype=bool)
    flags[[0, 1, 4]] = True  # u at (0,0), u_x at (0,0.12)
    value = lambda x, l_: (0.001 * l_ if abs(x[0] - 0.24) < 1e-12 else 0.0 * l_, 0.0 * l_)
    deriv = lambda x, l_: (0.001 if abs(x[0] - 0.24) < 1e-12 else 0.0, 0.0)
    idbc = o.InhomogeneousDirichlet(mesh, [(value, deriv)])
    flags = idbc.flag(flags)
    assert flags.sum() == 5
    for dbc in ("full", "reduced"):
        asm = o.FlatAssembler(mesh, kind, mat, flags, "interleaved")
        d = np.zeros(16)
        R = asm.vector(d, 1.0, dbc) + o.idbc_forces(asm, d, 1.0, dbc, idbc)
        x = spla.spsolve(asm.matrix(d, 1.0, dbc).tocsc(), -R)
        d = x if dbc == "full" else asm.create_full_vector(x)
        d = idbc.sync(d, 1.0)
        big = np.abs(PATCH_EXPECTED_D) > 1e-10
        assert np.abs(d[big] - PATCH_EXPECTED_D[big]).max() < 1e-10
        u = d[mesh.elem_dofs()].reshape(5, 4, 2)
        sig = o.stress_at(kind, mat, mesh.corner_coords, u, np.array([0.5, 0.5]))
        assert np.abs(sig[:, 0] - GOLDEN["plane_stress_patch_test"]["sigma_xx"]).max() < 1e-10  # constant stress state


import pytest


@pytest.mark.xfail(strict=True, reason="anchor B1 is not reproduced by the restatement (3e-5 relative); see DESIGN.md section 8")
def test_B1_plane_stress_block_under_volume_load_is_not_reproduced():
    """tests/src/testnonlinearelasticity.hh:47-150 with createGrid<Grids::Yasp> (tests/src/testcommon.hh:70-79):
    unit square, 10 x 10 Quad4, planeStress(SVK(E = 1000, nu = 0.3), 1e-8), volume load (lambda, 2 lambda), y = 0
    clamped, lambda 0 -> 50.  Kept as a strict xfail so that the day the cause is found this test says so: none of
    load scale, E, nu alone maps the restatement's (energy, max d) onto the reference's pair, NR and TR reach the same
    minimum, and for SVK the stress reduction is exact after one Newton step, so the tolerance plays no role."""
    g = GOLDEN["not_reproduced"]["plane_stress_volume_load_B1"]
    mesh = o.structured_mesh((10, 10), (1.0, 1.0))
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material("svk", lam, mu, plane_stress=True, ps_tol=1e-8)
    kind = o.ElementKind(2, 1, "gl")
    flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 1, 0.0))
    fext = o.volume_load_vector(mesh, kind, lambda x: (1.0, 2.0))
    asm = o.FlatAssembler(mesh, kind, mat, flags, fext=fext)
    d, lamb, info = o.load_control(asm, np.zeros(asm.n), 5, 0.0, 50.0, tol=1e-11, dbc="full", max_iter=50)
    assert info["success"]
    # what the restatement gives is pinned too, so a change of the oracle shows up here
    assert abs(asm.scalar(d, lamb) - g["oracle_energy"]) < 1e-9 and abs(d.max() - g["oracle_max_d"]) < 1e-10
    assert abs(asm.scalar(d, lamb) - g["energy"]) < 1e-8 * abs(g["energy"])
    assert abs(d.max() - g["max_d"]) < 1e-12
