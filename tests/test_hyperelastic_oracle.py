"""The oracle's principal-stretch framework against the reference's own known answers: for every deviatoric function
(BlatzKo, Ogden total / deviatoric, Mooney-Rivlin, Yeoh, ArrudaBoyce, Gent) and five deformation states the energy, the
first derivatives and the 'second derivative' array of tests/src/materialresultcollection.hh, which
tests/src/testhyperelasticity.hh:198-228 checks to 1e-14; the NeoHooke recovery of :139-196; the volumetric functions'
derivatives (:100-137, there by automatic differentiation, here by complex step)."""
import json
import os

import numpy as np
import pytest

import ikarus_oracle as o

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "material_results.json")))
LAMBDA = G["lambda"]
MU = 1000.0 / (2.0 * 1.25)          # testMatPar(): E = 1000, nu = 0.25 (testhyperelasticity.hh:36-40)
LAME = 1000.0 * 0.25 / (1.25 * 0.5)
BULK = 1000.0 / (3.0 * 0.5)
OG_MU, OG_AL = (2.0 * MU / 3.0, MU / 6.0, MU / 6.0), (1.23, 0.59, 0.18)
LAWS = {  # testhyperelasticity.hh:238-249
    "BlatzKo": o.Hyper("blatzko", (MU,)),
    "OgdenTotal": o.Hyper("ogden_total", (OG_MU, OG_AL)),
    "OgdenDeviatoric": o.Hyper("ogden_dev", (OG_MU, OG_AL)),
    "MooneyRivlin": o.Hyper("invariant", ((1, 0), (0, 1), (MU / 2.0, MU / 2.0))),
    "Yeoh": o.Hyper("invariant", ((1, 2, 3), (0, 0, 0), (MU / 2.0, MU / 6.0, MU / 3.0))),
    "ArrudaBoyce": o.Hyper("arrudaboyce", (MU, 0.85)),
    "Gent": o.Hyper("gent", (MU, 2.5)),
}


def right_cauchy_green(state, lam=LAMBDA):
    """Deformations::rightCauchyGreen (testhyperelasticity.hh:42-98)."""
    F = {"Undeformed": np.eye(3), "Uniaxial": np.diag([lam, 1 / np.sqrt(lam), 1 / np.sqrt(lam)]),
         "Biaxial": np.diag([lam, lam, 1 / (lam * lam)]), "PureShear": np.diag([lam, 1.0, 1 / lam])}.get(state)
    return np.array(G["random_C"]) if F is None else F.T @ F


@pytest.mark.parametrize("state", ["Undeformed", "Uniaxial", "Biaxial", "PureShear", "Random"])
@pytest.mark.parametrize("name", sorted(LAWS))
def test_deviatoric_functions_reproduce_the_reference_tables(name, state):
    W, dW, d2S = LAWS[name].functions()
    lam = np.sqrt(np.linalg.eigvalsh(right_cauchy_green(state)))  # ascending, as Eigen::SelfAdjointEigenSolver
    ref = G["functions"][name][state]
    tol = 1e-13  # the reference checks 1e-14 on its own build; two ulp-level different pow() evaluations here
    scale = max(1.0, abs(ref["energy"]))
    assert abs(W(lam) - ref["energy"]) <= tol * scale
    f = np.array(ref["first"])
    assert np.abs(dW(lam) - f).max() <= tol * max(1.0, np.abs(f).max())
    s = np.array(ref["second"])
    assert np.abs(d2S(lam) - s).max() <= tol * max(1.0, np.abs(s).max())


def _voigt_strain(C):
    E = 0.5 * (C - np.eye(3))
    return o.to_voigt(E[None], strain=True)[0]


@pytest.mark.parametrize("state", ["Undeformed", "Uniaxial", "Biaxial", "PureShear", "Random"])
def test_neohooke_recovery(state):
    """testhyperelasticity.hh:139-196: Ogden<1, total>({mu}, {2}) + VF3 with Lame's first parameter IS NeoHooke, and
    Ogden<1, deviatoric>({mu}, {2}) IS InvariantBased<1>({mu / 2}, p = 1, q = 0) -- energy, stresses and tangent moduli,
    also for repeated principal stretches (the L_ikik limit)."""
    E6 = _voigt_strain(right_cauchy_green(state))[None]
    nh = o.Material("neohooke", LAME, MU)
    og = o.Material("hyperelastic", 0.0, 0.0, hyper=o.Hyper("ogden_total", ((MU,), (2.0,)), vf=3, K=LAME))
    ogd = o.Material("hyperelastic", 0.0, 0.0, hyper=o.Hyper("ogden_dev", ((MU,), (2.0,))))
    inv = o.Material("hyperelastic", 0.0, 0.0, hyper=o.Hyper("invariant", ((1,), (0,), (MU / 2.0,))))
    for a, b in ((nh, og), (inv, ogd)):
        pa, Sa, Ca = a.evaluate(E6)
        pb, Sb, Cb = b.evaluate(E6)
        assert abs(pa - pb).max() <= 1e-12 * max(1.0, abs(pa).max())
        assert np.abs(Sa - Sb).max() <= 1e-12 * max(1.0, np.abs(Sa).max())
        assert np.abs(Ca - Cb).max() <= 1e-11 * np.abs(Ca).max()


@pytest.mark.parametrize("vf,beta", [(1, 0), (2, 0), (3, 0), (4, 0.5), (5, 0), (6, 0), (7, 0.5), (8, 0), (9, 0), (10, 0.4),
                                     (11, 0), (12, 0)])
def test_volumetric_functions_are_consistent(vf, beta):
    """testhyperelasticity.hh:100-137 at J = sqrt(det testMatrix): U' and U'' are the derivatives of U (complex step for
    the first, central difference of U' for the second)."""
    U, dU, ddU = o._volumetric(vf, beta)
    J = np.sqrt(np.linalg.det(np.array(G["random_C"])))
    h = 1e-30
    assert abs(np.imag(U(J + 1j * h)) / h - dU(J)) <= 1e-14 * max(1.0, abs(dU(J)))
    assert abs(np.imag(dU(J + 1j * h)) / h - ddU(J)) <= 1e-14 * max(1.0, abs(ddU(J)))


@pytest.mark.parametrize("name", ["OgdenTotal", "OgdenDeviatoric", "MooneyRivlin", "Yeoh", "ArrudaBoyce", "Gent"])
def test_laws_with_volumetric_part_are_hyperelastic(name):
    """S = dpsi/dE and CC = dS/dE by central differences (Voigt, engineering shear) with a volumetric part attached."""
    law = LAWS[name]
    mat = o.Material("hyperelastic", 0.0, 0.0, hyper=o.Hyper(law.dev, law.params, vf=2, K=BULK))
    rng = np.random.default_rng(3)
    E6 = 0.08 * rng.uniform(-1, 1, 6)
    psi, S, C = mat.evaluate(E6[None])
    h = 1e-6
    Sfd = np.zeros(6)
    Cfd = np.zeros((6, 6))
    for q in range(6):
        dE = np.zeros(6)
        dE[q] = h
        pp, Sp, _ = mat.evaluate((E6 + dE)[None])
        pm, Sm_, _ = mat.evaluate((E6 - dE)[None])
        Sfd[q] = (pp[0] - pm[0]) / (2 * h)
        Cfd[:, q] = (Sp[0] - Sm_[0]) / (2 * h)
    assert np.abs(Sfd - S[0]).max() <= 1e-7 * np.abs(S[0]).max()
    assert np.abs(Cfd - C[0]).max() <= 1e-7 * np.abs(C[0]).max()
    assert np.abs(C[0] - C[0].T).max() <= 1e-12 * np.abs(C[0]).max()
