"""CPU oracle for the Ikarus global-assembly hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference algorithm for global FEM assembly
(K, R, E), the three Dirichlet modes and the Newton/LoadControl drivers.  It is the
*checker* for the CUDA path in ``ikarus_b200``; nothing in the product imports it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg
may import this module.

The reference itself (header-only C++ over DUNE + Eigen + dune-localfefunctions)
cannot be compiled or imported here (no DUNE/Eigen in the image), so
this oracle is a restatement.  It is pinned against the reference's own known-answer
tests (see ``tests/test_oracle_anchors.py``):

* A1/A2  cantilever Hex8+EAS21 / Quad4+EAS4, NeoHooke and SVK: 80 Newton iterations and
         max|d| to 1e-10                       (tests/src/testcantileverbeamEAS.cpp:20-32,61-70)
* A3/A4  vertex stress tables of the unit square / cube (tests/src/resultcollection.hh:19-51,150-163)
* A5     K_nonlinear(u=0) == K_linear          (tests/src/testnonlinearelasticity.hh:194-246)
* A6     assembler invariants Raw/Full/Reduced (tests/src/testassembler.cpp:122-201)
* A7     EAS: int M dV = 0, eas(0) == plain element, K symmetric (tests/src/testeas.hh:50-85)
* A8     R = dE/dd, K = dR/dd by finite differences (tests/src/testnonlinearelasticity.hh:340-369)

The element math deliberately follows the reference's *generic* formulation
(Voigt B-operator, 6x6 tangent, B^T C B + geometric stiffness) and not the factored
spatial form the CUDA kernels use, so the two are independent derivations.

All `file:line` citations are relative to /root/reference.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------------------
# Quadrature and shape functions (DUNE conventions)
# --------------------------------------------------------------------------------------


def gauss_legendre_01(n: int):
    """n-point Gauss-Legendre rule on [0,1] (DUNE QuadratureRules on the reference cube;
    restated in ikarus/utils/quadraturerulehelper.hh:29-48)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def tensor_rule(dim: int, n: int):
    """Tensor Gauss rule, first coordinate fastest. Point order only changes the
    summation order (<=1e-15 effects); the reference default is order 2*basisOrder
    (ikarus/finiteelements/mechanics/nonlinearelastic.hh:121-124)."""
    x, w = gauss_legendre_01(n)
    pts, wts = [], []
    for idx in itertools.product(range(n), repeat=dim):
        idx = idx[::-1]  # first coordinate fastest
        pts.append([x[i] for i in idx])
        wts.append(np.prod([w[i] for i in idx]))
    return np.array(pts), np.array(wts)


def _lagrange1d(order: int, xi: float):
    if order == 1:
        return np.array([1.0 - xi, xi]), np.array([-1.0, 1.0])
    if order == 2:
        v = np.array([2.0 * (xi - 0.5) * (xi - 1.0), 4.0 * xi * (1.0 - xi), 2.0 * xi * (xi - 0.5)])
        d = np.array([4.0 * xi - 3.0, 4.0 - 8.0 * xi, 4.0 * xi - 1.0])
        return v, d
    raise NotImplementedError(order)


def shape_functions(dim: int, order: int, xi):
    """Lagrange cube shape functions N[a] and reference gradients dN[a, k], node index
    lexicographic with x fastest (DUNE LagrangeCube local order)."""
    n1 = order + 1
    vals = [_lagrange1d(order, xi[k]) for k in range(dim)]
    nn = n1**dim
    N = np.empty(nn)
    dN = np.empty((nn, dim))
    for a in range(nn):
        idx = [(a // n1**k) % n1 for k in range(dim)]
        N[a] = np.prod([vals[k][0][idx[k]] for k in range(dim)])
        for k in range(dim):
            dN[a, k] = np.prod([vals[m][1][idx[m]] if m == k else vals[m][0][idx[m]] for m in range(dim)])
    return N, dN


# --------------------------------------------------------------------------------------
# Voigt helpers (ikarus/utils/tensorutils.hh:180-326)
# --------------------------------------------------------------------------------------

VOIGT3 = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]
VOIGT2 = [(0, 0), (1, 1), (0, 1)]


def voigt_pairs(dim):
    return VOIGT3 if dim == 3 else VOIGT2


def to_voigt(E, strain=True):
    """Batched matrix -> Voigt (tensorutils.hh:248-276); shear entries doubled for strains."""
    dim = E.shape[-1]
    f = 2.0 if strain else 1.0
    cols = [E[..., i, j] * (1.0 if i == j else f) for (i, j) in voigt_pairs(dim)]
    return np.stack(cols, axis=-1)


def from_voigt(v, strain=True):
    """Batched Voigt -> matrix (tensorutils.hh:293-326)."""
    s = v.shape[-1]
    dim = 3 if s == 6 else 2
    f = 0.5 if strain else 1.0
    E = np.zeros(v.shape[:-1] + (dim, dim))
    for q, (i, j) in enumerate(voigt_pairs(dim)):
        val = v[..., q] * (1.0 if i == j else f)
        E[..., i, j] = val
        E[..., j, i] = val
    return E


def tensor4_to_voigt(T):
    """(…,3,3,3,3) -> (…,6,6) (tensorutils.hh:219-227)."""
    out = np.zeros(T.shape[:-4] + (6, 6))
    for p, (i, j) in enumerate(VOIGT3):
        for q, (k, l) in enumerate(VOIGT3):
            out[..., p, q] = T[..., i, j, k, l]
    return out


# --------------------------------------------------------------------------------------
# Materials (ikarus/finiteelements/mechanics/materials/*)
# --------------------------------------------------------------------------------------


# deviatoric functions of the principal-stretch framework: each returns (W, dW/dlambda_i, dS[i][k]) of lambda[..., 3],
# dS as the reference's secondDerivativeImpl returns it (= Hess W - diag(W,i / lambda_i), the array the Deviatoric
# interface divides by lambda_i lambda_k, deviatoric/interface.hh:95-99).  The formulas follow the reference's
# statements term by term.
def _blatzko(mu):
    # deviatoric/blatzko.hh:60-92:  W = mu/2 (sum lambda_i^-2 + 2 J - 5)
    def W(lam):
        J = lam.prod(-1)
        return 0.5 * mu * ((1.0 / lam**2).sum(-1) + 2.0 * J - 5.0)

    def dW(lam):
        J = lam.prod(-1)
        return mu * (-1.0 / lam**3 + J[..., None] / lam)

    def d2S(lam):
        J = lam.prod(-1)
        dS = J[..., None, None] / (lam[..., :, None] * lam[..., None, :])
        diag = (1.0 / lam**2) * (1.0 / lam**2 - J[..., None]) + 3.0 / lam**4
        idx = np.arange(3)
        dS[..., idx, idx] = diag
        return mu * dS

    return W, dW, d2S


def _ogden(mus, alphas, deviatoric):
    """deviatoric/ogden.hh:98-190, total or deviatoric stretches (lambdaBar = J^(-1/3) lambda, materialhelpers.hh:150-156)."""
    mus, alphas = [float(m) for m in mus], [float(a) for a in alphas]

    def bar(lam):
        return lam * (lam.prod(-1) ** (-1.0 / 3.0))[..., None]

    def W(lam):
        e = 0.0
        if deviatoric:
            lb = bar(lam)
            for m, a in zip(mus, alphas):
                e = e + m / a * ((lb**a).sum(-1) - 3.0)
        else:
            logJ = np.log(lam.prod(-1))
            for m, a in zip(mus, alphas):
                e = e + m / a * ((lam**a).sum(-1) - 3.0) - m * logJ
        return e

    def dW(lam):
        if deviatoric:
            lb = bar(lam)
            dWb = sum(m * lb ** (a - 1.0) for m, a in zip(mus, alphas))
            sumLb = (lb * dWb).sum(-1)
            return (lb * dWb - (1.0 / 3.0) * sumLb[..., None]) / lam
        return sum(m * (lam**a - 1.0) for m, a in zip(mus, alphas)) / lam

    def d2S(lam):
        dS = np.zeros(lam.shape + (3,))
        idx = np.arange(3)
        if deviatoric:
            lb = bar(lam)
            dWl = dW(lam)
            for a_ in range(3):
                for b_ in range(3):
                    v = 0.0
                    for m, al in zip(mus, alphas):
                        psum = (lb**al).sum(-1)
                        if a_ == b_:
                            v = v + m * al * (1.0 / 3.0 * lb[..., a_] ** al + 1.0 / 9.0 * psum)
                        else:
                            v = v + m * al * (-(1.0 / 3.0) * (lb[..., a_] ** al + lb[..., b_] ** al) + 1.0 / 9.0 * psum)
                    v = v * (1.0 / (lam[..., a_] * lam[..., b_]))
                    if a_ == b_:
                        v = v - (2.0 / lam[..., a_]) * dWl[..., a_]
                    dS[..., a_, b_] = v
        else:
            for m, al in zip(mus, alphas):
                dS[..., idx, idx] += (-2.0 * m * (lam**al - 1.0) + m * lam**al * al) / lam**2
        return dS

    return W, dW, d2S


def _dev_invariants(lam):
    """deviatoric/deviatoricinvariants.hh:52-107 over materialhelpers.hh:158-170: W1 = I1 I3^(-1/3), W2 = I2 I3^(-2/3) and
    their first / second derivatives with respect to the principal stretches."""
    l2 = lam * lam
    I1 = l2.sum(-1)
    I2 = l2[..., 0] * l2[..., 1] + l2[..., 1] * l2[..., 2] + l2[..., 0] * l2[..., 2]
    I3 = l2.prod(-1)
    p13, p23 = I3 ** (1.0 / 3.0), I3 ** (2.0 / 3.0)
    W1 = I1 * I3 ** (-1.0 / 3.0)
    W2 = I2 * I3 ** (-2.0 / 3.0)
    d1 = 2.0 * (3.0 * l2 - I1[..., None]) / (3.0 * lam * p13[..., None])
    d2 = -2.0 * (3.0 * I3[..., None] / l2 - I2[..., None]) / (3.0 * lam * p23[..., None])
    dd1 = np.zeros(lam.shape + (3,))
    dd2 = np.zeros(lam.shape + (3,))
    for i in range(3):
        for j in range(3):
            if i == j:
                dd1[..., i, i] = (2.0 / 9.0) * (5.0 * I1 - 3.0 * l2[..., i]) / (l2[..., i] * p13)
                dd2[..., i, i] = (2.0 / 9.0) * ((15.0 * I3 / l2[..., i]) - I2) / (l2[..., i] * p23)
            else:
                dd1[..., i, j] = (4.0 / 9.0) * (I1 - 3.0 * (l2[..., i] + l2[..., j])) / (lam[..., i] * lam[..., j] * p13)
                dd2[..., i, j] = (-4.0 / 9.0) * (2.0 * I2 - 3.0 * l2[..., i] * l2[..., j]) / (lam[..., i] * lam[..., j] * p23)
    return W1, W2, d1, d2, dd1, dd2


def _pam(x, p, m):
    # InvariantBasedT::powerAndMultiply (invariantbased.hh:205-214)
    if m == 0:
        return 0.0 * x
    if p == 0:
        return m + 0.0 * x
    return x ** int(p) * m


def _invariant_based(pex, qex, c):
    """deviatoric/invariantbased.hh:84-188: W = sum_i c_i (W1 - 3)^p_i (W2 - 3)^q_i (Mooney-Rivlin: p = (1, 0), q = (0, 1);
    Yeoh: p = (1, 2, 3), q = 0; factory.hh:68-108)."""
    pex, qex, c = [int(v) for v in pex], [int(v) for v in qex], [float(v) for v in c]
    assert all(p or q for p, q in zip(pex, qex)), "the exponents p_i and q_i must not both be zero (invariantbased.hh:216-221)"

    def W(lam):
        W1, W2 = _dev_invariants(lam)[:2]
        W1, W2 = W1 - 3.0, W2 - 3.0
        return sum(m * W1**p * W2**q for m, p, q in zip(c, pex, qex))

    def dW(lam):
        W1, W2, d1, d2, _, _ = _dev_invariants(lam)
        W1, W2 = W1 - 3.0, W2 - 3.0
        out = np.zeros_like(lam)
        for m, p, q in zip(c, pex, qex):
            out += m * ((_pam(W1, p - 1, p) * W2**q)[..., None] * d1 + (W1**p * _pam(W2, q - 1, q))[..., None] * d2)
        return out

    def d2S(lam):
        W1, W2, d1, d2, dd1, dd2 = _dev_invariants(lam)
        W1, W2 = W1 - 3.0, W2 - 3.0
        dS = np.zeros(lam.shape + (3,))
        for m, p, q in zip(c, pex, qex):
            a1, a2 = _pam(W1, p - 1, p), _pam(W2, q - 1, q)
            b1, b2 = _pam(W1, p - 2, p * (p - 1)), _pam(W2, q - 2, q * (q - 1))
            for i in range(3):
                for j in range(3):
                    mix = d1[..., i] * d2[..., j] + d1[..., j] * d2[..., i]
                    f1 = (a2 * mix + W2**q * dd1[..., i, j]) * a1 * m
                    f2 = a2 * dd2[..., i, j] * W1**p * m
                    f3 = b1 * W2**q * d1[..., i] * d1[..., j] * m
                    f4 = b2 * W1**p * d2[..., i] * d2[..., j] * m
                    dS[..., i, j] += f1 + f2 + f3 + f4
                    if i == j:
                        dS[..., i, j] -= (1.0 / lam[..., i]) * m * (a1 * W2**q * d1[..., i] + W1**p * a2 * d2[..., i])
        return dS

    return W, dW, d2S


_AB_ALPHAS = (0.5, 1.0 / 20.0, 11.0 / 1050.0, 19.0 / 7000.0, 519.0 / 673750.0)


def _arruda_boyce(mu, lambdaM):
    """deviatoric/arrudaboyce.hh:88-160: five-term series in W1 with beta = 1 / lambdaM^2."""
    beta = 1.0 / lambdaM**2.0

    def W(lam):
        W1 = _dev_invariants(lam)[0]
        return mu * sum(a * beta**i * (W1 ** (i + 1) - 3.0 ** (i + 1)) for i, a in enumerate(_AB_ALPHAS))

    def dW(lam):
        W1, _, d1, _, _, _ = _dev_invariants(lam)
        return sum((mu * a * beta**j * W1**j * (j + 1))[..., None] * d1 for j, a in enumerate(_AB_ALPHAS))

    def d2S(lam):
        W1, _, d1, _, dd1, _ = _dev_invariants(lam)
        dyad = d1[..., :, None] * d1[..., None, :]
        dS = np.zeros(lam.shape + (3,))
        idx = np.arange(3)
        for p, a in enumerate(_AB_ALPHAS):
            f1 = mu * a * beta**p
            f2 = (W1**p)[..., None, None] * dd1 * (p + 1)
            f3 = (W1 ** (p - 1) * p * (p + 1))[..., None, None] * dyad if p else 0.0
            dS += f1 * (f2 + f3)
            dS[..., idx, idx] -= d1 / lam * f1 * (W1**p * (p + 1))[..., None]
        return dS

    return W, dW, d2S


def _gent(mu, Jm):
    """deviatoric/gent.hh:88-150: W = -mu/2 Jm ln(1 - (W1 - 3) / Jm)."""

    def check(W1):
        if np.any(Jm <= W1 - 3.0):
            raise FloatingPointError("The material parameter Jm should be greater than (W1 - 3)")

    def W(lam):
        W1 = _dev_invariants(lam)[0]
        check(W1)
        return -(mu / 2.0) * Jm * np.log(1.0 - ((W1 - 3.0) / Jm))

    def dW(lam):
        W1, _, d1, _, _, _ = _dev_invariants(lam)
        check(W1)
        return (mu * d1 * Jm) / (2.0 * (Jm - W1) + 6.0)[..., None]

    def d2S(lam):
        W1, _, d1, _, dd1, _ = _dev_invariants(lam)
        check(W1)
        f = 1.0 - ((W1 - 3.0) / Jm)
        dS = (mu / (2.0 * f))[..., None, None] * (dd1 + d1[..., :, None] * d1[..., None, :] / (f * Jm)[..., None, None])
        idx = np.arange(3)
        dS[..., idx, idx] -= (mu / (2.0 * lam * f[..., None])) * d1
        return dS

    return W, dW, d2S


def _no_deviatoric():
    # deviatoric/nodeviatoricfunction.hh (makePureVolumetric, factory.hh:161-165)
    return (lambda lam: 0.0 * lam[..., 0]), (lambda lam: 0.0 * lam), (lambda lam: np.zeros(lam.shape + (3,)))


def _volumetric(vf, beta):
    """volumetric/volumetricfunctions.hh:25-380: (U, U', U'') of J for VF0..VF12; VF4, VF7, VF10 take beta."""
    b = beta
    ln = np.log
    table = {
        0: (lambda J: 0.0 * J, lambda J: 0.0 * J, lambda J: 0.0 * J),
        1: (lambda J: 0.5 * (J - 1.0) ** 2, lambda J: J - 1.0, lambda J: 1.0 + 0.0 * J),
        2: (lambda J: 0.25 * ((J - 1.0) ** 2 + ln(J) ** 2), lambda J: 0.5 * (J - 1.0 + 1.0 / J * ln(J)),
            lambda J: 1.0 / (2.0 * J * J) * (1.0 + J * J - ln(J))),
        3: (lambda J: 0.5 * ln(J) ** 2, lambda J: 1.0 / J * ln(J), lambda J: 1.0 / J**2 * (1.0 - ln(J))),
        4: (lambda J: (1.0 / b**2) * ((1.0 / J**b) - 1.0 + b * ln(J)), lambda J: (1.0 / b) * ((1.0 / J) - (1.0 / (J ** (1.0 + b)))),
            lambda J: (1.0 / b) * ((1.0 / J ** (2.0 + b)) * (1.0 + b - J**b))),
        5: (lambda J: 0.25 * (J**2 - 1.0 - 2.0 * ln(J)), lambda J: 0.5 * (J - (1.0 / J)), lambda J: 0.5 * (1.0 + (1.0 / J**2))),
        6: (lambda J: J - ln(J) - 1.0, lambda J: 1.0 - (1.0 / J), lambda J: 1.0 / J**2),
        7: (lambda J: J**b * (b * ln(J) - 1.0) + 1.0, lambda J: b**2 * (1.0 / J ** (1.0 - b)) * ln(J),
            lambda J: b**2 * J ** (b - 2.0) * (1.0 + (b - 1.0) * ln(J))),
        8: (lambda J: J * ln(J) - J + 1.0, lambda J: ln(J), lambda J: 1.0 / J),
        9: (lambda J: (1.0 / 32.0) * (J**2 - J**-2.0) ** 2, lambda J: (1.0 / 8.0) * (J**3 - (1.0 / J**5)),
            lambda J: (1.0 / 8.0) * (5.0 * J**-6.0 + (3.0 * J**2))),
        10: (lambda J: (J / b) * (1.0 - (J**-b / (1.0 - b))) + (1.0 / (b - 1.0)), lambda J: (1.0 / b) * (1.0 - J**-b),
             lambda J: J ** (-1.0 - b)),
        11: (lambda J: (1.0 / 50.0) * (J**5.0 + J**-5.0 - 2.0), lambda J: (1.0 / 10.0) * (J**4.0 - J**-6.0),
             lambda J: (1.0 / 10.0) * (4.0 * J**3.0 + 6.0 * J**-7.0)),
        12: (lambda J: J - 1.0, lambda J: 1.0 + 0.0 * J, lambda J: 0.0 * J),
    }
    return table[int(vf)]


@dataclass
class Hyper:
    """A material of the principal-stretch framework, Hyperelastic<Deviatoric<DF>, Volumetric<VF>> as the factories build
    it (materials/hyperelastic/factory.hh:12-165).  dev / params:
      'blatzko' (mu,) | 'ogden_total' | 'ogden_dev' (mus, alphas) | 'invariant' (pex, qex, c) | 'arrudaboyce' (mu, lambdaM)
      | 'gent' (mu, Jm) | 'none' ();  vf = 0..12 with the penalty parameter K (Lame's first parameter for total, the bulk
    modulus for deviatoric stretches) and beta for VF4 / VF7 / VF10."""

    dev: str
    params: tuple
    vf: int = 0
    K: float = 0.0
    beta: float = 0.0

    def functions(self):
        d, p = self.dev, self.params
        if d == "blatzko":
            return _blatzko(float(p[0]))
        if d in ("ogden_total", "ogden_dev"):
            return _ogden(p[0], p[1], d == "ogden_dev")
        if d == "invariant":
            return _invariant_based(*p)
        if d == "arrudaboyce":
            return _arruda_boyce(float(p[0]), float(p[1]))
        if d == "gent":
            return _gent(float(p[0]), float(p[1]))
        if d == "none":
            return _no_deviatoric()
        raise NotImplementedError(d)


_DEVIATORIC = ("blatzko", "hyperelastic")


def lame_from_E_nu(E, nu):
    """physicshelper.hh:276-281."""
    lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    mu = E / (2.0 * (1.0 + nu))
    return lam, mu


@dataclass
class Material:
    """kind in {'linear', 'svk', 'neohooke'}; plane_strain wraps the 3D law in the
    VanishingStrain reduction (materials/vanishingstrain.hh:79-120)."""

    kind: str
    lam: float
    mu: float
    plane_strain: bool = False
    # planeStress(mat, tol) = VanishingStress<{2,2},{1,2},{0,2}> (materials/vanishingstress.hh:35-230, 252-260)
    plane_stress: bool = False
    ps_tol: float = 1e-12
    hyper: "Hyper | None" = None  # kind 'hyperelastic': the principal-stretch law; kind 'blatzko' = Hyper('blatzko', (mu,))

    # --- 3D law on Voigt GL strain (batched: E6[..., 6]) --------------------------------
    def _law3d(self, E6):
        lam, mu = self.lam, self.mu
        if self.kind in ("linear", "svk"):
            # svk.hh:77-164 ; linearelasticity.hh forwards to SVK
            C = np.zeros(E6.shape[:-1] + (6, 6))
            C[..., :3, :3] = lam
            for i in range(3):
                C[..., i, i] += 2.0 * mu
                C[..., 3 + i, 3 + i] = mu
            S = np.einsum("...pq,...q->...p", C, E6)
            tr = E6[..., :3].sum(-1)
            psi = 0.5 * lam * tr**2 + mu * ((E6[..., :3] ** 2).sum(-1) + 0.5 * (E6[..., 3:] ** 2).sum(-1))
            return psi, S, C
        if self.kind == "neohooke":
            # neohooke.hh:79-142 with C = 2E + I (strainconversions.hh:88-106)
            Cm = 2.0 * from_voigt(E6, strain=True) + np.eye(3)
            detC = np.linalg.det(Cm)
            if np.any(detC <= 0.0):  # materialhelpers.hh:120-126: Dune::FloatCmp::le(det, 0, 1e-10), relativeWeak == (det <= 0)
                raise FloatingPointError("Determinant of right Cauchy Green tensor C must be greater than zero")
            logJ = np.log(np.sqrt(detC))
            invC = np.linalg.inv(Cm)
            I = np.eye(3)
            trC = np.trace(Cm, axis1=-2, axis2=-1)
            psi = 0.5 * mu * (trC - 3.0 - 2.0 * logJ) + 0.5 * lam * logJ**2
            Sm = mu * (I - invC) + (lam * logJ)[..., None, None] * invC
            dy = np.einsum("...ij,...kl->...ijkl", invC, invC)
            ikjl = np.einsum("...ik,...jl->...ijkl", invC, invC)
            iljk = np.einsum("...il,...jk->...ijkl", invC, invC)
            T = lam * dy + (2.0 * (mu - lam * logJ))[..., None, None, None, None] * 0.5 * (ikjl + iljk)
            return psi, to_voigt(Sm, strain=False), tensor4_to_voigt(T)
        if self.kind in _DEVIATORIC:
            return self._hyperelastic(E6)
        raise NotImplementedError(self.kind)

    def _hyperelastic(self, E6):
        """Materials::Hyperelastic<Deviatoric<DF>, Volumetric<VF>> in principal stretches
        (materials/hyperelastic/interface.hh:99-232, deviatoric/interface.hh:77-115, materialhelpers.hh:138-164), input
        C = 2E + I:  lambda_i = sqrt(eig_i(C)), N = eigenvectors,
            S   = sum_i (W,i / lambda_i) N_i (x) N_i  +  J U'(J) C^-1
            L_iikk = dS(i,k) / (lambda_i lambda_k),   L_ikik = (S_i - S_k) / (lambda_i^2 - lambda_k^2)
                     [ 0.5 (L_iiii - L_iikk) when Dune::FloatCmp::eq(lambda_i, lambda_k, 1e-8) ]
            CC  = sum_ik L_iikk N_iN_i (x) N_kN_k + sum_{i != k} L_ikik N_iN_k (x) (N_iN_k + N_kN_i)
                  + J ((U' + J U'') C^-1 (x) C^-1 - 2 U' sym(C^-1 (.) C^-1))
        Here mu is the deviatoric parameter (makeBlatzKo(mu), factory.hh:34-39); the volumetric function is VF0 (none)."""
        law = self.hyper if self.kind == "hyperelastic" else Hyper("blatzko", (self.mu,))
        W, dW, d2S = law.functions()
        U, dU, ddU = _volumetric(law.vf, law.beta)
        Cm = 2.0 * from_voigt(E6, strain=True) + np.eye(3)
        ev, N = np.linalg.eigh(Cm)  # ascending, like Eigen::SelfAdjointEigenSolver
        lam = np.sqrt(ev)
        J = lam.prod(-1)
        if np.any(J <= 0.0) or np.any(ev <= 0.0):
            raise FloatingPointError("Determinant of right Cauchy Green tensor C must be greater than zero")
        psi = W(lam)
        Sp = dW(lam) / lam  # principal PK2 stresses
        dS = d2S(lam)       # [.., i, k]
        L1 = dS / (lam[..., :, None] * lam[..., None, :])  # L_iikk
        lam2 = lam * lam
        num = Sp[..., :, None] - Sp[..., None, :]
        den = lam2[..., :, None] - lam2[..., None, :]
        # Dune::FloatCmp::eq(a, b, 1e-8), relativeWeak: |a - b| <= 1e-8 * max(|a|, |b|)
        close = np.abs(lam[..., :, None] - lam[..., None, :]) <= 1e-8 * np.maximum(np.abs(lam[..., :, None]), np.abs(lam[..., None, :]))
        diagL = np.diagonal(L1, axis1=-2, axis2=-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            L2 = np.where(close, 0.5 * (diagL[..., :, None] - L1), num / np.where(den == 0.0, 1.0, den))
        Sm = np.einsum("...i,...ai,...bi->...ab", Sp, N, N)
        T = np.einsum("...ik,...ai,...bi,...ck,...dk->...abcd", L1, N, N, N, N)
        off = 1.0 - np.eye(3)
        T = T + np.einsum("...ik,ik,...ai,...bk,...ci,...dk->...abcd", L2, off, N, N, N, N)
        T = T + np.einsum("...ik,ik,...ai,...bk,...ck,...di->...abcd", L2, off, N, N, N, N)
        if law.vf:
            # interface.hh:103-157, 196-216: + K U(J);  S += J K U' C^-1;  CC += J ((U' + J U'') C^-1 (x) C^-1 - 2 U' sym(...))
            Up, Upp = law.K * dU(J), law.K * ddU(J)
            invC = np.linalg.inv(Cm)
            psi = psi + law.K * U(J)
            Sm = Sm + (J * Up)[..., None, None] * invC
            dy = np.einsum("...ij,...kl->...ijkl", invC, invC)
            ikjl = np.einsum("...ik,...jl->...ijkl", invC, invC)
            iljk = np.einsum("...il,...jk->...ijkl", invC, invC)
            T = T + J[..., None, None, None, None] * ((Up + J * Upp)[..., None, None, None, None] * dy
                                                        - (2.0 * Up)[..., None, None, None, None] * 0.5 * (ikjl + iljk))
        return psi, to_voigt(Sm, strain=False), tensor4_to_voigt(T)

    def _reduce_stress(self, Ev):
        """VanishingStress::reduceStress (vanishingstress.hh:150-199): Newton-Raphson (tol, at most 100 iterations,
        correction = -(A^-1 r)) on the DIAGONAL fixed strain component only -- E33, or C33 started from 1 for
        materials whose native measure is the right Cauchy-Green tensor (NeoHooke; derivative = moduli / 2) -- while
        the fixed shear strains stay zero.  Returns the full 6-Voigt GL strain of the solution."""
        lam, mu, tol = self.lam, self.mu, self.ps_tol
        E6 = np.zeros(Ev.shape[:-1] + (6,))
        E6[..., [0, 1, 5]] = Ev
        flat = E6.reshape(-1, 6)
        for row in flat:  # scalar Newton per point, exactly the reference's control flow
            if self.kind == "neohooke":
                C2 = np.array([[2 * row[0] + 1, row[5]], [row[5], 2 * row[1] + 1]])
                det2 = C2[0, 0] * C2[1, 1] - C2[0, 1] * C2[1, 0]
                c33 = 1.0
                for it in range(101):
                    detC = det2 * c33
                    if detC <= 0.0:  # same check, see above
                        raise FloatingPointError("Determinant of right Cauchy Green tensor C must be greater than zero")
                    lnJ = np.log(np.sqrt(detC))
                    f = mu * (1.0 - 1.0 / c33) + lam * lnJ / c33
                    if not (abs(f) > tol and it < 100):
                        break
                    df = (lam + 2.0 * (mu - lam * lnJ)) / (c33 * c33) / 2.0
                    c33 += -((1.0 / df) * f)
                if abs(f) > tol:
                    raise FloatingPointError("The stress reduction of the material was unsuccessful")
                row[2] = 0.5 * (c33 - 1.0)
            else:
                e33 = 0.0
                tr2 = row[0] + row[1]
                for it in range(101):
                    f = lam * (tr2 + e33) + 2.0 * mu * e33
                    if not (abs(f) > tol and it < 100):
                        break
                    e33 += -((1.0 / (lam + 2.0 * mu)) * f)
                if abs(f) > tol:
                    raise FloatingPointError("The stress reduction of the material was unsuccessful")
                row[2] = e33
        return flat.reshape(E6.shape)

    def evaluate(self, Ev):
        """Return (psi, S_voigt, C_voigt) for Voigt strain Ev[..., s], s = 3 (2D) or 6."""
        if Ev.shape[-1] == 6:
            assert not self.plane_strain and not self.plane_stress
            return self._law3d(Ev)
        free = [0, 1, 5]  # vanishingstrain.hh / vanishingstress.hh: fixed Voigt indices {2,3,4}
        if self.plane_stress:
            fixed = [2, 3, 4]
            E6 = self._reduce_stress(Ev)
            psi, S6, C6 = self._law3d(E6)
            # staticCondensation(C, fixedVoigtIndices) (utils/linearalgebrahelper.hh): K11 - K12 K22^-1 K21
            Cff = C6[..., free, :][..., :, free]
            Cfx = C6[..., free, :][..., :, fixed]
            Cxx = C6[..., fixed, :][..., :, fixed]
            Cred = Cff - Cfx @ np.linalg.solve(Cxx, np.swapaxes(Cfx, -1, -2))
            return psi, S6[..., free], Cred
        assert self.plane_strain, "2D elements need a reduced material (planeStrain or planeStress)"
        E6 = np.zeros(Ev.shape[:-1] + (6,))
        E6[..., free] = Ev
        psi, S6, C6 = self._law3d(E6)
        return psi, S6[..., free], C6[..., free, :][..., :, free]


# --------------------------------------------------------------------------------------
# EAS ansatz (strainenhancements/easvariants/linearandglstrains.hh:71-328)
# --------------------------------------------------------------------------------------

# (row in Voigt, polynomial as tuple of reference directions) per column
_EAS_TABLE = {
    (2, 4): [(0, (0,)), (1, (1,)), (2, (0,)), (2, (1,))],
    (2, 5): [(0, (0,)), (1, (1,)), (2, (0,)), (2, (1,)), (2, (0, 1))],
    (2, 7): [(0, (0,)), (1, (1,)), (2, (0,)), (2, (1,)), (0, (0, 1)), (1, (0, 1)), (2, (0, 1))],
    (3, 9): [(0, (0,)), (1, (1,)), (2, (2,)), (3, (1,)), (3, (2,)), (4, (0,)), (4, (2,)), (5, (0,)), (5, (1,))],
}
_EAS_TABLE[(3, 21)] = _EAS_TABLE[(3, 9)] + [
    (3, (0, 1)), (3, (0, 2)), (4, (0, 1)), (4, (1, 2)), (5, (0, 2)), (5, (1, 2)),
    (0, (0, 1)), (0, (0, 2)), (1, (0, 1)), (1, (1, 2)), (2, (0, 2)), (2, (1, 2)),
]


def eas_table(dim, m):
    if m == 0:
        return []
    if (dim, m) not in _EAS_TABLE:
        raise NotImplementedError(f"EAS variant E{m} in {dim}D")  # enhancedassumedstrains.hh:250-256
    return _EAS_TABLE[(dim, m)]


def eas_Mhat(dim, m, xi):
    s = dim * (dim + 1) // 2
    M = np.zeros((s, m))
    t = 2.0 * np.asarray(xi) - 1.0
    for j, (row, dirs) in enumerate(eas_table(dim, m)):
        M[row, j] = np.prod([t[k] for k in dirs])
    return M


def transformation_matrix(Jt):
    """Voigt transformation matrix from the transposed Jacobian (tensorutils.hh:408-465).
    Batched over leading dims of Jt[..., d, d] with Jt[i, j] = dx_j/dxi_i."""
    d = Jt.shape[-1]
    if d == 2:
        J11, J12, J21, J22 = Jt[..., 0, 0], Jt[..., 0, 1], Jt[..., 1, 0], Jt[..., 1, 1]
        rows = [[J11 * J11, J12 * J12, J11 * J12],
                [J21 * J21, J22 * J22, J21 * J22],
                [2 * J11 * J21, 2 * J12 * J22, J21 * J12 + J11 * J22]]
    else:
        J11, J12, J13 = Jt[..., 0, 0], Jt[..., 0, 1], Jt[..., 0, 2]
        J21, J22, J23 = Jt[..., 1, 0], Jt[..., 1, 1], Jt[..., 1, 2]
        J31, J32, J33 = Jt[..., 2, 0], Jt[..., 2, 1], Jt[..., 2, 2]
        rows = [
            [J11 * J11, J12 * J12, J13 * J13, J12 * J13, J11 * J13, J11 * J12],
            [J21 * J21, J22 * J22, J23 * J23, J22 * J23, J21 * J23, J21 * J22],
            [J31 * J31, J32 * J32, J33 * J33, J32 * J33, J31 * J33, J31 * J32],
            [2 * J21 * J31, 2 * J22 * J32, 2 * J23 * J33, J22 * J33 + J32 * J23, J31 * J23 + J21 * J33, J21 * J32 + J31 * J22],
            [2 * J11 * J31, 2 * J12 * J32, 2 * J13 * J33, J12 * J33 + J32 * J13, J11 * J33 + J31 * J13, J11 * J32 + J31 * J12],
            [2 * J11 * J21, 2 * J12 * J22, 2 * J13 * J23, J12 * J23 + J22 * J13, J11 * J23 + J21 * J13, J11 * J22 + J12 * J21],
        ]
    return np.stack([np.stack(r, axis=-1) for r in rows], axis=-2)


# --------------------------------------------------------------------------------------
# Element kernels, batched over elements
# --------------------------------------------------------------------------------------


@dataclass
class ElementKind:
    """dim, basis order (1 = Q1, 2 = Q2), strain ('linear' | 'gl'), EAS parameter count."""

    dim: int
    order: int = 1
    strain: str = "gl"
    eas_m: int = 0
    quad_n: int | None = None  # points per direction; default = order+1 (rule of order 2*order)
    # which measure the EAS parameters enhance (mechanics/strainenhancements/easfunctions/*.hh): "strain" = LinearStrain /
    # GreenLagrangeStrain (E4..E21), "dg" = DisplacementGradient, "dgt" = DisplacementGradientTransposed (H4 / H9)
    eas_function: str = "strain"

    @property
    def nodes(self):
        return (self.order + 1) ** self.dim

    @property
    def ndof(self):
        return self.nodes * self.dim

    @property
    def corners(self):
        return 2**self.dim

    def rule(self):
        return tensor_rule(self.dim, self.quad_n or (self.order + 1))


def _geometry(kind: ElementKind, X, xi):
    """Multilinear cube geometry from the 2^d corners X[e, c, d] (the *grid* geometry is
    used also for Q2 elements, nonlinearelastic.hh:117).  Returns Jt[e, i, j] = dx_j/dxi_i,
    its inverse and |det|."""
    _, dNg = shape_functions(kind.dim, 1, xi)
    Jt = np.einsum("ci,ecj->eij", dNg, X)
    detJ = np.abs(np.linalg.det(Jt))
    return Jt, np.linalg.inv(Jt), detJ


def _b_operator(kind: ElementKind, gradN, F):
    """B[e, a, s, d] = dE_voigt/dd_a (restated in the reference in
    strainenhancements/easfunctions/displacementgradient.hh:123-147).  For linear strains
    F is the identity (linearelastic.hh strainFunction)."""
    d = kind.dim
    ne, nn = gradN.shape[0], gradN.shape[1]
    s = d * (d + 1) // 2
    B = np.zeros((ne, nn, s, d))
    for q, (i, j) in enumerate(voigt_pairs(d)):
        if i == j:
            # N_a,i * g_i^T   with g_i = column i of F
            B[:, :, q, :] = gradN[:, :, i, None] * F[:, None, :, i]
        else:
            B[:, :, q, :] = gradN[:, :, j, None] * F[:, None, :, i] + gradN[:, :, i, None] * F[:, None, :, j]
    return B


def element_quantities(kind: ElementKind, mat: Material, X, u, alpha=None, want=("K", "R", "E")):
    """Per-element tangent K[e, ndof, ndof], residual R[e, ndof], energy E[e].

    Follows NonLinearElastic/LinearElastic::calculate{Matrix,Vector,Scalar}Impl
    (mechanics/nonlinearelastic.hh:376-430, mechanics/linearelastic.hh:354-405) and, with
    eas_m > 0, EnhancedAssumedStrains (mechanics/enhancedassumedstrains.hh:258-348,378-434).

    X: corner coordinates [e, 2^d, d]; u: nodal displacements [e, nodes, d];
    alpha: EAS parameters [e, m].
    Also returns the EAS blocks (D, L, Rtilde) when eas_m > 0.
    """
    d, nn, nd, m = kind.dim, kind.nodes, kind.ndof, kind.eas_m
    if m and kind.eas_function in ("dg", "dgt"):
        if kind.strain != "gl":
            raise NotImplementedError("EAS::DisplacementGradient enhances a nonlinear element")
        return _element_quantities_dg(kind, mat, X, u, np.zeros((X.shape[0], m)) if alpha is None else alpha, want)
    s = d * (d + 1) // 2
    ne = X.shape[0]
    pts, wts = kind.rule()
    K = np.zeros((ne, nn, d, nn, d))
    R = np.zeros((ne, nn, d))
    En = np.zeros(ne)
    if m:
        if alpha is None:
            alpha = np.zeros((ne, m))
        Dm = np.zeros((ne, m, m))
        Lm = np.zeros((ne, m, nn, d))
        Rt = np.zeros((ne, m))
        # EX ctor: T0InverseTransformed = (T(center) * detJ0)^-1   (easvariants/helperfunctions.hh:18-25)
        Jt0, _, detJ0 = _geometry(kind, X, np.full(d, 0.5))
        T0inv = np.linalg.inv(transformation_matrix(Jt0) * detJ0[:, None, None])
    I = np.eye(d)
    for xi, w in zip(pts, wts):
        _, dN = shape_functions(d, kind.order, xi)
        Jt, Jtinv, detJ = _geometry(kind, X, xi)
        gradN = np.einsum("eji,ai->eaj", Jtinv, dN)  # dN_a/dx_j
        H = np.einsum("eac,eaj->ecj", u, gradN)  # H[c, j] = du_c/dx_j
        if kind.strain == "gl":
            F = I + H
            Em = 0.5 * (H + np.swapaxes(H, -1, -2) + np.einsum("eki,ekj->eij", H, H))
        else:
            F = np.broadcast_to(I, H.shape)
            Em = 0.5 * (H + np.swapaxes(H, -1, -2))
        Ev = to_voigt(Em, strain=True)
        if m:
            Mh = eas_Mhat(d, m, xi)
            M = np.einsum("epq,qm->epm", T0inv, Mh) / detJ[:, None, None]  # linandglstrains.hh E*::operator()
            Ev = Ev + np.einsum("epm,em->ep", M, alpha)
        psi, S, C = mat.evaluate(Ev)
        if kind.strain == "linear":
            # LinearElastic: stress = C*eps, energy = 0.5 eps.C.eps (linearelastic.hh:158-177)
            S = np.einsum("epq,eq->ep", C, Ev)
            psi = 0.5 * np.einsum("ep,ep->e", Ev, S)
        B = _b_operator(kind, gradN, F)
        wd = w * detJ
        if "K" in want:
            K += np.einsum("eapi,epq,ebqk,e->eaibk", B, C, B, wd, optimize=True)
            if kind.strain == "gl":
                Sm = from_voigt(S, strain=False)
                kg = np.einsum("eai,eij,ebj,e->eab", gradN, Sm, gradN, wd, optimize=True)
                for c in range(d):
                    K[:, :, c, :, c] += kg
        if "R" in want or m:
            R += np.einsum("eapi,ep,e->eai", B, S, wd)
        if "E" in want:
            En += psi * wd
        if m:
            CM = np.einsum("epq,eqm->epm", C, M)
            Dm += np.einsum("epm,epn,e->emn", M, CM, wd)
            Lm += np.einsum("epm,eapi,e->emai", CM, B, wd)
            Rt += np.einsum("epm,ep,e->em", M, S, wd)
    K = K.reshape(ne, nd, nd)
    R = R.reshape(ne, nd)
    out = {}
    if m:
        Lf = Lm.reshape(ne, m, nd)
        Dinv = np.linalg.inv(Dm)
        # K.upper -= L^T D^-1 L ; lower = upper^T   (enhancedassumedstrains.hh:292-296)
        Kc = K - np.einsum("emi,emn,enj->eij", Lf, Dinv, Lf, optimize=True)
        iu = np.triu_indices(nd)
        Ks = np.zeros_like(Kc)
        Ks[:, iu[0], iu[1]] = Kc[:, iu[0], iu[1]]
        Ks = Ks + np.swapaxes(np.triu(Ks, 1), -1, -2)
        K = Ks
        R = R - np.einsum("emi,emn,en->ei", Lf, Dinv, Rt, optimize=True)  # :341-345
        out.update(D=Dm, L=Lf, Rtilde=Rt)
        En = np.full(ne, np.nan)  # EAS elements expose no potential (:302-310)
    out.update(K=K, R=R, E=En)
    return out


def eas_Hhat(dim, m, xi):
    """Reference-element ansatz of the displacement-gradient enhancement: H4 / H9
    (strainenhancements/easvariants/displacementgradient.hh:76-163): mode p = dim*i + j is (2 xi_j - 1) e_i (x) e_j."""
    if (dim, m) not in ((2, 4), (3, 9)):
        raise NotImplementedError(f"EAS variant H{m} in {dim}D")
    H = np.zeros((m, dim, dim))
    for i in range(dim):
        for j in range(dim):
            H[dim * i + j, i, j] = 2.0 * xi[j] - 1.0
    return H


def _element_quantities_dg(kind: ElementKind, mat: Material, X, u, alpha, want):
    """EnhancedAssumedStrains with EAS::DisplacementGradient (mechanics/enhancedassumedstrains.hh:258-434 over
    strainenhancements/easfunctions/displacementgradient.hh:40-282): the enhanced displacement gradient is
        H = H_c + sum_p alpha_p Htilde_p,  Htilde_p = (detJ0/detJ) J0^-T Hhat_p J0^-1     (helperfunctions.hh:27-36)
    and E = (H + H^T + H^T H)/2.  Every derivative the reference writes out is an instance of one rule: for a variation
    dF of the deformation gradient (e_c (x) grad N_a for a nodal dof, Htilde_p for an enhanced one)
        E,I  = sym(F^T dF_I)  (Voigt, engineering shear)        (:107-165, both wrtCoeff)
        E,IJ : S = tr(S dF_I^T dF_J)                               (:190-245, wrtCoeff 0, 1, 2)
    so K_uu, L, D and R, Rtilde are blocks of one generalised tangent / residual."""
    d, nn, nd, m = kind.dim, kind.nodes, kind.ndof, kind.eas_m
    transposed = kind.eas_function == "dgt"
    ne = X.shape[0]
    pts, wts = kind.rule()
    G = nd + m
    Kg = np.zeros((ne, G, G))
    Rg = np.zeros((ne, G))
    Jt0, Jt0inv, detJ0 = _geometry(kind, X, np.full(d, 0.5))
    I = np.eye(d)
    vp = voigt_pairs(d)
    if transposed:
        # compatible gradient and shape-function gradients at the element centre (centerPosition, :327-333)
        _, dN0 = shape_functions(d, kind.order, np.full(d, 0.5))
        gradN0 = np.einsum("eji,ai->eaj", Jt0inv, dN0)
        Fc0 = I + np.einsum("eac,eaj->ecj", u, gradN0)
    for xi, w in zip(pts, wts):
        _, dN = shape_functions(d, kind.order, xi)
        Jt, Jtinv, detJ = _geometry(kind, X, xi)
        gradN = np.einsum("eji,ai->eaj", Jtinv, dN)
        Hc = np.einsum("eac,eaj->ecj", u, gradN)
        Hh = eas_Hhat(d, m, xi)
        # jacobianInverseTransposed(center) maps reference to physical gradients: it is Jt0inv here
        Ht = np.einsum("e,eik,pkl,ejl->epij", detJ0 / detJ, Jt0inv, Hh, Jt0inv)
        Hsum = np.einsum("epij,ep->eij", Ht, alpha)
        gN = gradN
        if transposed:
            # H = H_c + F_c0 Htilde^T (displacementgradienttransposed.hh:343-360); a nodal variation then sees the
            # gradient g_a + Htilde g_a^0 (dNtilde, :133-141)
            H = Hc + np.einsum("eik,ejk->eij", Fc0, Hsum)
            gN = gradN + np.einsum("ejk,eak->eaj", Hsum, gradN0)
        else:
            H = Hc + Hsum
        F = I + H
        Em = 0.5 * (H + np.swapaxes(H, -1, -2) + np.einsum("eki,ekj->eij", H, H))
        psi, S, C = mat.evaluate(to_voigt(Em, strain=True))
        Sm = from_voigt(S, strain=False)
        # variations of F: nodal dofs (a, c) -> e_c (x) grad N_a, then the enhanced modes
        dF = np.zeros((ne, G, d, d))
        for a in range(nn):
            for c in range(d):
                dF[:, a * d + c, c, :] = gN[:, a, :]
        dF[:, nd:] = np.einsum("eik,epjk->epij", Fc0, Ht) if transposed else Ht
        if transposed:
            # the one non-vanishing second variation: d_alpha_p d_u(a,c) F = e_c (x) (Htilde_p g_a^0); its work with
            # P = F S is the dNXHtilde part of E,ad (:297-318)
            P = np.einsum("eik,ekj->eij", F, Sm)
            hg = np.einsum("epjk,eak->epaj", Ht, gradN0)
            mixed = np.einsum("ecj,epaj,e->epac", P, hg, w * detJ).reshape(ne, m, nd)
            Kg[:, nd:, :nd] += mixed
            Kg[:, :nd, nd:] += np.swapaxes(mixed, -1, -2)
        FtdF = np.einsum("eki,egkj->egij", F, dF)
        Bg = np.zeros((ne, G, len(vp)))
        for q, (i, j) in enumerate(vp):
            Bg[:, :, q] = FtdF[:, :, i, i] if i == j else FtdF[:, :, i, j] + FtdF[:, :, j, i]
        wd = w * detJ
        Kg += np.einsum("egp,epq,ehq,e->egh", Bg, C, Bg, wd, optimize=True)
        Kg += np.einsum("eij,egki,ehkj,e->egh", Sm, dF, dF, wd, optimize=True)
        Rg += np.einsum("egp,ep,e->eg", Bg, S, wd)
    K, R = Kg[:, :nd, :nd], Rg[:, :nd]
    Dm, Lf, Rt = Kg[:, nd:, nd:], Kg[:, nd:, :nd], Rg[:, nd:]
    Dinv = np.linalg.inv(Dm)
    Kc = K - np.einsum("emi,emn,enj->eij", Lf, Dinv, Lf, optimize=True)  # enhancedassumedstrains.hh:292-296
    iu = np.triu_indices(nd)
    Ks = np.zeros_like(Kc)
    Ks[:, iu[0], iu[1]] = Kc[:, iu[0], iu[1]]
    Ks = Ks + np.swapaxes(np.triu(Ks, 1), -1, -2)
    Rc = R - np.einsum("emi,emn,en->ei", Lf, Dinv, Rt, optimize=True)  # :341-345
    return dict(K=Ks, R=Rc, E=np.full(ne, np.nan), D=Dm.copy(), L=Lf.copy(), Rtilde=Rt.copy())


def eas_update_alpha(kind, mat, X, u, alpha, du):
    """alpha -= D^-1 (Rtilde + L * du_e), all at the OLD (u, alpha)
    (enhancedassumedstrains.hh:225-248; called on CORRECTION_UPDATED before x is updated,
    solver/nonlinearsolver/newtonraphson.hh:230-235)."""
    q = element_quantities(kind, mat, X, u, alpha, want=())
    rhs = q["Rtilde"] + np.einsum("emi,ei->em", q["L"], du.reshape(du.shape[0], -1))
    return alpha - np.linalg.solve(q["D"], rhs[..., None])[..., 0]


def stress_at(kind, mat, X, u, xi, alpha=None, result="native"):
    """calculateAt (nonlinearelastic.hh:237-271, linearelastic.hh:~200-240, enhancedassumedstrains.hh:127-187): stress in
    Voigt notation at local position xi for every element.  `result`:
      "native"    PK2Stress (GL kinematics) / linearStress (linear kinematics), s components
      "full"      PK2StressFull / linearStressFull: the underlying 3D law of a reduced (plane strain) material, 6 comps
      "kirchhoff" tau = F S F^T, "cauchy" tau / det F (transformStress, GL kinematics only)
    With EAS and linear strains alpha = -D^-1 L d as in EnhancedAssumedStrains::calculateAtImpl (:152-157)."""
    d = kind.dim
    _, dN = shape_functions(d, kind.order, xi)
    Jt, Jtinv, detJ = _geometry(kind, X, np.asarray(xi, float))
    gradN = np.einsum("eji,ai->eaj", Jtinv, dN)
    H = np.einsum("eac,eaj->ecj", u, gradN)
    if kind.strain == "gl":
        Em = 0.5 * (H + np.swapaxes(H, -1, -2) + np.einsum("eki,ekj->eij", H, H))
    else:
        Em = 0.5 * (H + np.swapaxes(H, -1, -2))
    Ev = to_voigt(Em)
    if kind.eas_m and kind.eas_function in ("dg", "dgt"):
        # EnhancedStrainFunction::computeDisplacementGradient with the stored alpha (enhancedassumedstrains.hh:161-170)
        if alpha is None:
            alpha = np.zeros((X.shape[0], kind.eas_m))
        _, Jt0inv, detJ0 = _geometry(kind, X, np.full(d, 0.5))
        Hsum = np.einsum("e,eik,pkl,ejl,ep->eij", detJ0 / detJ, Jt0inv, eas_Hhat(d, kind.eas_m, xi), Jt0inv, alpha)
        if kind.eas_function == "dgt":
            _, dN0 = shape_functions(d, kind.order, np.full(d, 0.5))
            Fc0 = np.eye(d) + np.einsum("eac,eaj->ecj", u, np.einsum("eji,ai->eaj", Jt0inv, dN0))
            H = H + np.einsum("eik,ejk->eij", Fc0, Hsum)
        else:
            H = H + Hsum
        Em = 0.5 * (H + np.swapaxes(H, -1, -2) + np.einsum("eki,ekj->eij", H, H))
        Ev = to_voigt(Em)
    elif kind.eas_m:
        if alpha is None:
            q = element_quantities(kind, mat, X, u, np.zeros((X.shape[0], kind.eas_m)), want=())
            alpha = -np.linalg.solve(q["D"], np.einsum("emi,ei->em", q["L"], u.reshape(u.shape[0], -1))[..., None])[..., 0]
        Jt0, _, detJ0 = _geometry(kind, X, np.full(d, 0.5))
        T0inv = np.linalg.inv(transformation_matrix(Jt0) * detJ0[:, None, None])
        M = np.einsum("epq,qm->epm", T0inv, eas_Mhat(d, kind.eas_m, xi)) / detJ[:, None, None]
        Ev = Ev + np.einsum("epm,em->ep", M, alpha)
    if result == "full" and d == 2:
        E6 = np.zeros(Ev.shape[:-1] + (6,))
        E6[..., [0, 1, 5]] = Ev
        psi, S, C = mat._law3d(E6)
        if kind.strain == "linear":
            S = np.einsum("epq,eq->ep", C, E6)
        return S
    psi, S, C = mat.evaluate(Ev)
    if kind.strain == "linear":
        S = np.einsum("epq,eq->ep", C, Ev)
    if result in ("kirchhoff", "cauchy"):
        if kind.strain != "gl":
            raise NotImplementedError("kirchhoffStress / cauchyStress need the nonlinear element")
        F = H + np.eye(d)
        tau = np.einsum("eik,ekl,ejl->eij", F, from_voigt(S, strain=False), F)
        if result == "cauchy":
            tau = tau / np.linalg.det(F)[:, None, None]
        return to_voigt(tau, strain=False)
    return S


# --------------------------------------------------------------------------------------
# Structured meshes with YaspGrid-compatible numbering
# --------------------------------------------------------------------------------------


@dataclass
class Mesh:
    dim: int
    order: int
    node_coords: np.ndarray  # [nNodes, dim]   Lagrange node positions
    elem_nodes: np.ndarray  # [nElem, nodes]  global node ids in DUNE local order
    corner_coords: np.ndarray  # [nElem, 2^d, dim]
    cells: tuple = ()
    extra: dict = field(default_factory=dict)

    @property
    def n_nodes(self):
        return self.node_coords.shape[0]

    @property
    def n_elem(self):
        return self.elem_nodes.shape[0]

    def elem_dofs(self, layout="interleaved"):
        """power<d>(lagrange<k>, FlatInterleaved) => dof = d*node + comp;
        FlatLexicographic => comp*nNodes + node (ikarus/python/test/linearelastictest.py:213-231).
        Element-local order is node-major, component-minor (finiteelements/fehelper.hh:145-153)."""
        d = self.dim
        if layout == "interleaved":
            return (self.elem_nodes[:, :, None] * d + np.arange(d)[None, None, :]).reshape(self.n_elem, -1)
        return (np.arange(d)[None, None, :] * self.n_nodes + self.elem_nodes[:, :, None]).reshape(self.n_elem, -1)


def structured_mesh(cells, bbox, order=1, mapping=None) -> Mesh:
    """YaspGrid(bbox, cells): axis-aligned cells, vertices and elements numbered
    lexicographically with x fastest; lagrange<1> global index = vertex index.
    For order 2 the nodes live on the (2n+1)^d lattice numbered lexicographically
    (NOT dune-functions' vertex/edge/face/cell ordering, which cannot be cross-checked
    here -- the device code consumes elemDofs and is numbering-agnostic)."""
    cells = tuple(int(c) for c in cells)
    dim = len(cells)
    npts = [order * c + 1 for c in cells]
    grids = [np.linspace(0.0, bbox[k], npts[k]) for k in range(dim)]
    # lexicographic, x fastest
    mg = np.meshgrid(*grids, indexing="ij")
    coords = np.stack([g.reshape(-1, order="F") for g in mg], axis=-1)
    stride = np.cumprod([1] + npts[:-1])
    eidx = np.stack([g.reshape(-1, order="F") for g in np.meshgrid(*[np.arange(c) for c in cells], indexing="ij")], axis=-1)
    n1 = order + 1
    loc = np.array([[(a // n1**k) % n1 for k in range(dim)] for a in range(n1**dim)])
    elem_nodes = ((eidx[:, None, :] * order + loc[None, :, :]) * stride[None, None, :]).sum(-1)
    cloc = np.array([[(a >> k) & 1 for k in range(dim)] for a in range(2**dim)]) * order
    corner_nodes = ((eidx[:, None, :] * order + cloc[None, :, :]) * stride[None, None, :]).sum(-1)
    if mapping is not None:
        coords = mapping(coords)
    return Mesh(dim, order, coords, elem_nodes.astype(np.int64), coords[corner_nodes], cells)


def boundary_nodes(mesh: Mesh, axis: int, value: float, tol=1e-8):
    return np.nonzero(np.abs(mesh.node_coords[:, axis] - value) < tol)[0]


def fix_nodes(mesh: Mesh, nodes, layout="interleaved", comps=None):
    """Dirichlet flags (utils/dirichletvalues.hh:111-131 fixBoundaryDOFs equivalent)."""
    d = mesh.dim
    flags = np.zeros(mesh.n_nodes * d, dtype=bool)
    comps = range(d) if comps is None else comps
    for c in comps:
        if layout == "interleaved":
            flags[np.asarray(nodes) * d + c] = True
        else:
            flags[c * mesh.n_nodes + np.asarray(nodes)] = True
    return flags


# --------------------------------------------------------------------------------------
# Flat assembler (ikarus/assembler/interface.hh, simpleassemblers.inl)
# --------------------------------------------------------------------------------------


def constraints_below(flags):
    """constraintsBelow_[i] = number of fixed dofs with index < i (assembler/interface.hh:51-62)."""
    f = np.asarray(flags, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(f)[:-1]])


def build_pattern(elem_dofs, n, flags=None):
    """Sparsity pattern as Eigen's setFromTriplets leaves it (compressed, inner indices
    sorted) -- simpleassemblers.inl:206-251.  The pattern is structurally symmetric, so
    the CSC arrays Eigen holds equal the CSR arrays returned here.  With flags: the
    Reduced pattern (rows/cols of fixed dofs dropped, index i - constraintsBelow(i)).
    Returns (outer[n+1] int64, inner[nnz] int32)."""
    ed = np.asarray(elem_dofs, dtype=np.int64)
    if flags is not None:
        cb = constraints_below(flags)
        keep = ~np.asarray(flags)[ed]
        red = ed - cb[ed]
        n = int(n - np.count_nonzero(flags))
    ne, nd = ed.shape
    rows_l, cols_l = [], []
    chunk = max(1, 2_000_000 // (nd * nd))
    keys = []
    for s in range(0, ne, chunk):
        e = ed[s:s + chunk]
        if flags is None:
            r = np.repeat(e, nd, axis=1).ravel()
            c = np.tile(e, (1, nd)).ravel()
        else:
            k = keep[s:s + chunk]
            rr = red[s:s + chunk]
            msk = (k[:, :, None] & k[:, None, :]).ravel()
            r = np.repeat(rr, nd, axis=1).ravel()[msk]
            c = np.tile(rr, (1, nd)).ravel()[msk]
        keys.append(np.unique(r * n + c))
    key = np.unique(np.concatenate(keys)) if keys else np.zeros(0, np.int64)
    rows = key // n
    inner = (key % n).astype(np.int32)
    outer = np.zeros(n + 1, dtype=np.int64)
    np.add.at(outer, rows + 1, 1)
    outer = np.cumsum(outer)
    return outer, inner


def linear_indices(elem_dofs, outer, inner):
    """Position of (row r, col c) in the value array (utils/eigensparseaddon.hh:19-31)."""
    ed = np.asarray(elem_dofs, dtype=np.int64)
    ne, nd = ed.shape
    r = np.repeat(ed, nd, axis=1)
    c = np.tile(ed, (1, nd))
    n = outer.shape[0] - 1
    key = outer[:-1].repeat(np.diff(outer)) * 0  # placeholder to keep dtype
    rowkey = np.repeat(np.arange(n, dtype=np.int64), np.diff(outer)) * n + inner
    pos = np.searchsorted(rowkey, (r * n + c).ravel())
    return pos.reshape(ne, nd * nd)


class FlatAssembler:
    """Restatement of Sparse/DenseFlatAssembler for one element kind + material on a mesh.

    matrix()/vector()/scalar() mirror assembler/interface.hh:300-462; DBC modes follow
    simpleassemblers.inl:59-204 (sparse) and :301-375 (dense).
    `fext` is the lambda-proportional external load vector (host-sampled volume/Neumann
    loads and point loads): R = F_int - lambda * fext.
    """

    def __init__(self, mesh: Mesh, kind: ElementKind, mat: Material, flags, layout="interleaved", fext=None):
        self.mesh, self.kind, self.mat = mesh, kind, mat
        self.flags = np.asarray(flags, dtype=bool)
        self.elem_dofs = mesh.elem_dofs(layout)
        self.n = mesh.n_nodes * mesh.dim
        self.cb = constraints_below(self.flags)
        self.n_red = self.n - int(self.flags.sum())
        self.fext = np.zeros(self.n) if fext is None else np.asarray(fext, float)
        self.alpha = np.zeros((mesh.n_elem, kind.eas_m)) if kind.eas_m else None
        self._pat = {}

    # -- maps ---------------------------------------------------------------------------
    def size(self):
        return self.n

    def reduced_size(self):
        return self.n_red

    def create_full_vector(self, red):
        full = np.zeros(self.n)
        full[~self.flags] = red
        return full

    def create_reduced_vector(self, full):
        return np.asarray(full)[~self.flags].copy()

    def pattern(self, dbc="raw"):
        key = "reduced" if dbc == "reduced" else "raw"
        if key not in self._pat:
            self._pat[key] = build_pattern(self.elem_dofs, self.n, self.flags if key == "reduced" else None)
        return self._pat[key]

    # -- element sweep ------------------------------------------------------------------
    def _local(self, d, want):
        u = d[self.elem_dofs].reshape(self.mesh.n_elem, self.kind.nodes, self.kind.dim)
        return element_quantities(self.kind, self.mat, self.mesh.corner_coords, u, self.alpha, want)

    def scalar(self, d, lam):
        q = self._local(np.asarray(d, float), ("E",))
        return float(q["E"].sum() - lam * self.fext @ d)

    def vector(self, d, lam, dbc="full"):
        q = self._local(np.asarray(d, float), ("R",))
        R = np.zeros(self.n)
        np.add.at(R, self.elem_dofs.ravel(), q["R"].ravel())
        R -= lam * self.fext
        if dbc == "raw":
            return R
        if dbc == "full":
            R[self.flags] = 0.0
            return R
        return R[~self.flags]

    def matrix_values(self, d, lam, dbc="full"):
        """CSR/CSC value array in the pattern of `pattern(dbc)`."""
        q = self._local(np.asarray(d, float), ("K",))
        Ke = q["K"]
        nd = self.kind.ndof
        if dbc == "reduced":
            outer, inner = self.pattern("reduced")
            n = self.n_red
            keep = ~self.flags[self.elem_dofs]
            red = self.elem_dofs - self.cb[self.elem_dofs]
            r = np.repeat(red, nd, axis=1)
            c = np.tile(red, (1, nd))
            msk = (keep[:, :, None] & keep[:, None, :]).reshape(self.mesh.n_elem, -1)
            rowkey = np.repeat(np.arange(n, dtype=np.int64), np.diff(outer)) * n + inner
            pos = np.searchsorted(rowkey, (r * n + c)[msk])
            vals = np.zeros(inner.shape[0])
            np.add.at(vals, pos, Ke.reshape(self.mesh.n_elem, -1)[msk])
            return vals
        outer, inner = self.pattern("raw")
        pos = linear_indices(self.elem_dofs, outer, inner)
        vals = np.zeros(inner.shape[0])
        np.add.at(vals, pos.ravel(), Ke.reshape(-1))
        if dbc == "full":
            rows = np.repeat(np.arange(self.n), np.diff(outer))
            kill = self.flags[rows] | self.flags[inner]
            vals[kill] = 0.0
            vals[(rows == inner) & self.flags[rows]] = 1.0
        return vals

    def matrix(self, d, lam, dbc="full"):
        import scipy.sparse as sp

        outer, inner = self.pattern(dbc)
        n = self.n_red if dbc == "reduced" else self.n
        return sp.csr_matrix((self.matrix_values(d, lam, dbc), inner, outer), shape=(n, n))

    def dense_matrix(self, d, lam, dbc="full"):
        """DenseFlatAssembler (simpleassemblers.inl:301-375)."""
        q = self._local(np.asarray(d, float), ("K",))
        A = np.zeros((self.n, self.n))
        for e in range(self.mesh.n_elem):
            idx = self.elem_dofs[e]
            A[np.ix_(idx, idx)] += q["K"][e]
        if dbc == "full":
            A[:, self.flags] = 0.0
            A[self.flags, :] = 0.0
            A[self.flags, self.flags] = 1.0
        elif dbc == "reduced":
            A = A[np.ix_(~self.flags, ~self.flags)]
        return A

    def update_eas(self, d, correction_full):
        if not self.kind.eas_m:
            return
        sh = (self.mesh.n_elem, self.kind.nodes, self.kind.dim)
        u = np.asarray(d)[self.elem_dofs].reshape(sh)
        du = np.asarray(correction_full)[self.elem_dofs].reshape(sh)
        self.alpha = eas_update_alpha(self.kind, self.mat, self.mesh.corner_coords, u, self.alpha, du)


# --------------------------------------------------------------------------------------
# NewtonRaphson + LoadControl (solver/nonlinearsolver/newtonraphson.hh:196-257,
# controlroutines/loadcontrol.inl:21-57)
# --------------------------------------------------------------------------------------


def volume_load_vector(mesh: Mesh, kind: ElementKind, f, layout="interleaved"):
    """VolumeLoad::calculateVectorImpl at lambda = 1 summed over the elements (mechanics/loads/volume.hh:86-106):
    fext_a = sum_gp N_a f(x_gp) detJ w with the element's own Gauss rule; the assembler then uses R -= lambda fext."""
    pts, wts = tensor_rule(kind.dim, kind.order + 1)
    X = mesh.corner_coords
    fext = np.zeros(mesh.n_nodes * kind.dim)
    dofs = mesh.elem_dofs(layout)
    for xi, w in zip(pts, wts):
        N, _ = shape_functions(kind.dim, kind.order, xi)
        Ng, _ = shape_functions(kind.dim, 1, xi)
        _, _, detJ = _geometry(kind, X, xi)
        xg = np.einsum("c,ecj->ej", Ng, X)
        fv = np.array([np.asarray(f(x), float) for x in xg])
        contrib = (N[None, :, None] * fv[:, None, :] * (detJ * w)[:, None, None]).reshape(mesh.n_elem, -1)
        np.add.at(fext, dofs.ravel(), contrib.ravel())
    return fext


class InhomogeneousDirichlet:
    """The inhomogeneous part of DirichletValues (utils/dirichletvalues.hh:214-281): nodal functions f(x, lambda)
    interpolated into the flat dof vector (Lagrange nodes => point evaluation), their lambda-derivatives, and the
    flag rule of setInhomogeneousBoundaryConditionFlag (:296-302: every dof whose value at `lambda0` is non-zero
    becomes constrained).  `fns` = list of (value(x, lam) -> dim-vector, derivative(x, lam) -> dim-vector)."""

    def __init__(self, mesh: Mesh, fns, layout="interleaved"):
        self.mesh, self.fns, self.layout = mesh, list(fns), layout

    def _interp(self, which, lam):
        m = self.mesh
        out = np.zeros(m.n_nodes * m.dim)
        for f in self.fns:
            vals = np.array([np.asarray(f[which](x, lam), float) for x in m.node_coords])  # [nNodes, dim]
            out += vals.reshape(-1) if self.layout == "interleaved" else vals.T.reshape(-1)
        return out

    def values(self, lam):
        return self._interp(0, lam)

    def derivative(self, lam):
        return self._interp(1, lam)

    def flag(self, flags, lam0=1.0):
        flags = np.array(flags, dtype=bool)
        flags[self.values(lam0) != 0.0] = True
        return flags

    def sync(self, d, lam):
        """Impl::updateFunctor, SyncFERequirements branch (nonlinearsolverfactory.hh:45-54)."""
        inc = self.values(lam)
        nz = inc != 0.0
        d[nz] = inc[nz]
        return d


def idbc_forces(asm: "FlatAssembler", d, lam, dbc, idbc: InhomogeneousDirichlet):
    """utils::obtainForcesDueToIDBC (utils/functionhelper.hh:170-185): K_raw * d(d_D)/d(lambda) at lambda = 1, zeroed at
    constrained dofs (Full) or reduced (otherwise)."""
    F = asm.matrix(d, lam, "raw") @ idbc.derivative(1.0)
    if dbc == "full":
        F[asm.flags] = 0.0
        return F
    return asm.create_reduced_vector(F)


def newton_raphson(asm: FlatAssembler, d, lam, tol=1e-8, max_iter=20, dbc="full", linear_solver=None, idbc=None,
                   step_size=0.0):
    import scipy.sparse.linalg as spla

    if linear_solver is None:
        linear_solver = lambda A, b: spla.spsolve(A.tocsc(), b)
    d = np.array(d, float)
    r = asm.vector(d, lam, dbc)
    A = asm.matrix(d, lam, dbc)
    if idbc is not None:
        r = r + idbc_forces(asm, d, lam, dbc, idbc) * step_size  # newtonraphson.hh:214-217
    rnorm = np.linalg.norm(r)
    it = 0
    while rnorm > tol and it < max_iter:
        corr = linear_solver(A, -r)
        full = corr if dbc != "reduced" else asm.create_full_vector(corr)
        asm.update_eas(d, full)  # CORRECTION_UPDATED before the solution update (:230-235)
        d = d + full
        if idbc is not None:
            d = idbc.sync(d, lam)  # x.syncParameterAndGlobalSolution (:237-238)
        r = asm.vector(d, lam, dbc)
        A = asm.matrix(d, lam, dbc)
        rnorm = np.linalg.norm(r)
        it += 1
    return d, dict(iterations=it, success=it != max_iter, residual_norm=rnorm)


def load_control(asm: FlatAssembler, d, load_steps, t_begin, t_end, lam0=0.0, **nr):
    """LoadControl::run: initial solve at the current lambda, then loadSteps increments (with inhomogeneous Dirichlet
    values the step size is handed to the solver and a step with 0 iterations still syncs d, loadcontrol.inl:44-46)."""
    step = (t_end - t_begin) / load_steps
    lam = lam0
    total = 0
    idbc = nr.get("idbc")
    d, info = newton_raphson(asm, d, lam, **nr)
    total += info["iterations"]
    curve = [(lam, float(np.abs(d).max()))]
    ok = info["success"]
    per_step = [info["iterations"]]
    for _ in range(load_steps):
        if not ok:
            break
        lam += step
        d, info = newton_raphson(asm, d, lam, **(dict(nr, step_size=step) if idbc is not None else nr))
        if idbc is not None and info["iterations"] == 0:
            d = idbc.sync(d, lam)
        total += info["iterations"]
        per_step.append(info["iterations"])
        ok = info["success"]
        curve.append((lam, float(np.abs(d).max())))
    return d, lam, dict(total_iterations=total, success=ok, curve=curve, per_step=per_step)


# ---------------------------------------------------------------------------------------------------------------
# Trust region with Steihaug-Toint truncated CG.
# Restates ikarus/linearalgebra/truncatedconjugategradient.hh:68-168 and
# ikarus/solver/nonlinearsolver/trustregion.hh:226-430, 481-547.  Pinned by tests/src/testtrustregion.cpp:124-190
# (11 outer iterations with the identity, 8 with the diagonal preconditioner, minimiser and energy to 1e-12).
TCG_STOP = ("negative curvature", "exceeded trust region", "reached target residual-kappa (linear)",
            "reached target residual-theta (superlinear)", "maximum inner iterations", "model increased")


def diagonal_preconditioner(A):
    """Eigen::DiagonalPreconditioner: 1/diag, 1 where the diagonal is zero."""
    dg = np.asarray(A.diagonal(), float)
    return np.where(dg != 0.0, 1.0 / np.where(dg != 0.0, dg, 1.0), 1.0)


def truncated_cg(A, b, x0, minv, Delta, kappa=0.1, theta=1.0, mininner=1, max_iters=None, tol=np.finfo(float).eps):
    """internal::truncated_conjugate_gradient (truncatedconjugategradient.hh:68-168).  `minv` is the inverse
    diagonal (ones for the identity preconditioner).  Returns x, info(iterations, stop, rel_error)."""
    n = b.shape[0]
    if max_iters is None:
        max_iters = 2 * n  # Eigen::IterativeSolverBase default
    x = np.array(x0, float)
    residual = b - A @ x
    rhs_norm = np.linalg.norm(b)
    tiny = np.finfo(float).tiny
    if rhs_norm <= tiny:
        return np.zeros(n), dict(iterations=0, stop=4, rel_error=0.0)
    threshold = max(tol * tol * rhs_norm * rhs_norm, tiny)
    res_norm = np.linalg.norm(residual)
    if res_norm * res_norm < threshold:
        return x, dict(iterations=0, stop=4, rel_error=res_norm / rhs_norm)
    e_Pd = 0.0
    e_Pe = float(x @ x)
    p = minv * residual
    stop = 4  # maximumInnerIterations
    abs_new = float(residual @ p)
    d_Pd = abs_new
    i = 1
    while i < max_iters:
        tmp = A @ p
        d_Hd = float(p @ tmp)
        # alpha is formed before the curvature test, exactly as in the reference (inf/nan for d_Hd == 0 is harmless:
        # the branch below is taken)
        with np.errstate(divide="ignore", invalid="ignore"):
            alpha = np.float64(abs_new) / np.float64(d_Hd)
            e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd
        if d_Hd <= 0 or e_Pe_new >= Delta * Delta:
            tau = (-e_Pd + np.sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd
            x = x + tau * p
            stop = 0 if d_Hd <= 0 else 1
            break
        e_Pe = e_Pe_new
        x = x + alpha * p
        residual = residual - alpha * tmp
        res_norm = np.linalg.norm(residual)
        if i >= mininner and res_norm <= rhs_norm * min(rhs_norm, kappa):
            stop = 2 if kappa < rhs_norm else 3
            break
        if res_norm < threshold:
            break
        z = minv * residual
        abs_old = abs_new
        abs_new = float(residual @ z)
        beta = abs_new / abs_old
        e_Pd = beta * (e_Pd + alpha * d_Pd)
        d_Pd = abs_new + beta * beta * d_Pd
        p = z + beta * p
        i += 1
    return x, dict(iterations=i, stop=stop, rel_error=res_norm / rhs_norm)


def trust_region(energy, gradient, hessian, x, update=None, precond="diagonal", min_iter=3, max_iter=1000,
                 grad_tol=1e-6, corr_tol=1e-6, rho_prime=0.01, rho_reg=1e6, Delta_bar=np.inf, Delta0=10.0,
                 on_correction=None):
    """TrustRegion::solve (trustregion.hh:226-430) without the random predictor.  `update(x, eta)` returns the new
    point (default x + eta); `on_correction(x, eta)` is the CORRECTION_UPDATED hook (EAS), called like the reference
    after the proposal was applied and before acceptance is decided (:396)."""
    if update is None:
        update = lambda xx, eta: xx + eta
    eps_ = 0.0001220703125
    x = np.array(x, float)
    e = energy(x)
    g = gradient(x)
    h = hessian(x)
    n = g.shape[0]
    stats = dict(energy=e, grad_norm=float(np.linalg.norm(g)), eta_norm=0.0, outer=0, inner_sum=0, rho=0.0)
    Delta = Delta0
    consecutive_rejected = 0
    history = []
    stop_reason = None
    while True:
        # stoppingCriterion (:481-527)
        if stats["grad_norm"] < grad_tol and stats["outer"] != 0:
            stop_reason = "gradient"
            break
        if stats["eta_norm"] < corr_tol and stats["outer"] != 0:
            stop_reason = "correction"
            break
        if stats["outer"] >= max_iter:
            stop_reason = "maxiter"
            break
        minv = diagonal_preconditioner(h) if precond == "diagonal" else np.ones(n)
        eta, inner = truncated_cg(h, -g, np.zeros(n), minv, Delta)
        stats["inner_sum"] += inner["iterations"]
        Heta = h @ eta
        stats["eta_norm"] = float(np.linalg.norm(eta))
        x = update(x, eta)
        e = energy(x)
        proposal = e
        rhonum = stats["energy"] - proposal
        rhoden = -float(eta @ (g + 0.5 * Heta))
        rho_r = max(1.0, abs(stats["energy"])) * eps_ * rho_reg
        rhonum += rho_r
        rhoden += rho_r
        model_decreased = rhoden > 0.0
        rho = rhonum / rhoden
        rho = -1.0 if rho < 0.0 else rho
        # Dune::FloatCmp::ge(energy - proposal, -1e-12): '>' or equal within the default relative-weak epsilon
        a_, b_ = stats["energy"] - proposal, -1e-12
        energy_decreased = a_ > b_ or abs(a_ - b_) <= 8 * np.finfo(float).eps * max(abs(a_), abs(b_))
        tr = "   "
        if rho < 1e-4 or not model_decreased or np.isnan(rho) or not energy_decreased:
            tr = "TR-"
            Delta /= 4.0
        elif rho > 0.99 and inner["stop"] in (0, 1):
            tr = "TR+"
            Delta = min(3.5 * Delta, Delta_bar)
        if model_decreased and rho > rho_prime and energy_decreased:
            accept = True
            consecutive_rejected = 0
        else:
            accept = False
            if consecutive_rejected >= 5:
                Delta /= 2
            else:
                Delta = min(Delta, stats["eta_norm"] / 2.0)
            consecutive_rejected += 1
        stats["outer"] += 1
        stats["rho"] = rho
        if on_correction is not None:
            on_correction(x, eta)
        history.append(dict(accept=accept, tr=tr, inner=inner["iterations"], stop=inner["stop"], rho=rho,
                            energy=stats["energy"], proposal=proposal, Delta=Delta, eta_norm=stats["eta_norm"]))
        if accept:
            stats["energy"] = proposal
        else:
            x = update(x, -eta)
        e = energy(x)
        g = gradient(x)
        h = hessian(x)
        stats["grad_norm"] = float(np.linalg.norm(g))
    success = stop_reason in ("gradient", "correction")
    return x, dict(iterations=stats["outer"], success=success, residual_norm=stats["grad_norm"], energy=e,
                   inner_iterations=stats["inner_sum"], stop=stop_reason, history=history)
