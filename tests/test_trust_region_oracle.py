"""The oracle's TrustRegion / truncated CG against the numbers the reference's own tests pin
(tests/src/testtrustregion.cpp:124-190): 11 outer iterations with the identity preconditioner, 8 with the diagonal
one, the minimiser and the energy."""
import numpy as np
import scipy.sparse as sp

import ikarus_oracle as o
from golden_data import GOLDEN


def _f3(x):
    return -10 * x[0] ** 2 + 10 * x[1] ** 2 + 4 * np.sin(x[0] * x[1]) - 2 * x[0] + x[0] ** 4


def _df3(x):
    c = np.cos(x[0] * x[1])
    return np.array([-20 * x[0] + 4 * c * x[1] - 2 + 4 * x[0] ** 3, 20 * x[1] + 4 * c * x[0]])


def _ddf3(x):
    s, c = np.sin(x[0] * x[1]), np.cos(x[0] * x[1])
    off = 4 * c - 4 * s * x[0] * x[1]
    return sp.csr_matrix(np.array([[-20 - 4 * s * x[1] ** 2 + 12 * x[0] ** 2, off], [off, 20 - 4 * s * x[0] ** 2]]))


def test_trust_region3_identity_and_diagonal_iteration_counts():
    g = GOLDEN["trust_region"]
    expected = np.array(g["minimiser"])
    for precond, iters in g["iterations"].items():
        x, info = o.trust_region(_f3, _df3, _ddf3, np.array([0.7, -3.3]), precond=precond, max_iter=30, grad_tol=1e-12,
                                 corr_tol=1e-12, Delta0=1)
        assert info["success"] and info["iterations"] == iters
        assert info["residual_norm"] < 1e-12
        assert np.abs(x - expected).max() < 1e-12
        assert abs(_f3(x) - g["energy"]) < 1e-12


def test_truncated_cg_stop_reasons():
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.normal(size=(12, 12)))
    A = sp.csr_matrix(Q @ np.diag(np.linspace(1, 50, 12)) @ Q.T)
    b = rng.normal(size=12)
    exact = np.linalg.solve(A.toarray(), b)
    # huge radius: stops on the kappa rule with a residual 10x smaller than |b| (kappa < |b|)
    x, info = o.truncated_cg(A, b, np.zeros(12), np.ones(12), 1e5)
    assert info["stop"] == 2 and np.linalg.norm(b - A @ x) <= 0.1 * np.linalg.norm(b)
    # tiny radius: lands exactly on the trust-region boundary
    x, info = o.truncated_cg(A, b, np.zeros(12), np.ones(12), 1e-3)
    assert info["stop"] == 1 and abs(np.linalg.norm(x) - 1e-3) < 1e-15
    # negative curvature: follows the direction to the boundary
    Aneg = sp.csr_matrix(Q @ np.diag(np.linspace(-3, 50, 12)) @ Q.T)
    x, info = o.truncated_cg(Aneg, Q[:, 0], np.zeros(12), np.ones(12), 2.0)
    assert info["stop"] == 0 and abs(np.linalg.norm(x) - 2.0) < 1e-12
    # kappa = 0 and mininner large: plain CG to the tolerance threshold
    x, info = o.truncated_cg(A, b, np.zeros(12), o.diagonal_preconditioner(A), 1e5, kappa=0.0, max_iters=200)
    assert np.abs(x - exact).max() < 1e-10
    # zero right-hand side
    x, info = o.truncated_cg(A, np.zeros(12), np.ones(12), np.ones(12), 1.0)
    assert info["iterations"] == 0 and not x.any()
