"""A torch-autograd twin of the oracle (the reference checks its elements against AutoDiffFE,
ikarus/finiteelements/autodifffe.hh:133-209 -- energy -> forces -> stiffness by automatic differentiation).

The twin knows ONLY the stored-energy densities, written here independently of oracle/ikarus_oracle.py in invariant /
principal-stretch form, and the kinematics E = 1/2 (H + H^T + H^T H) (+ M alpha for the strain enhancements).  torch
(float64, CPU) differentiates  W_e(u, alpha) = sum_gp w detJ psi  twice; the oracle's R, K -- and for the enhanced
elements its blocks D, L, Rtilde and the condensed tangent -- must be those derivatives to rounding.  Geometry constants
(shape-function derivatives at the Gauss points, Jacobians, the EAS ansatz) are not differentiated and are taken from
the oracle's helpers."""
import numpy as np
import pytest
import torch

import ikarus_oracle as o
from problems import distorted

torch.set_default_dtype(torch.float64)
E_MOD, NU = 1000.0, 0.3
LAM, MU = o.lame_from_E_nu(E_MOD, NU)
BULK = E_MOD / (3.0 * (1.0 - 2.0 * NU))


def _embed3(C, dim):
    """plane strain: the 3D law at zero out-of-plane strain (materials/vanishingstrain.hh) -- C33 = 1."""
    if dim == 3:
        return C
    C3 = torch.eye(3).repeat(C.shape[0], 1, 1)
    return torch.cat([torch.cat([C, C3[:, :2, 2:]], dim=2), C3[:, 2:, :]], dim=1)


def _stretches(C3):
    return torch.sqrt(torch.linalg.eigvalsh(C3))


def psi_svk(E3):
    tr = E3.diagonal(dim1=-2, dim2=-1).sum(-1)
    return 0.5 * LAM * tr**2 + MU * (E3 * E3).sum((-1, -2))


def psi_neohooke(E3):
    C = 2.0 * E3 + torch.eye(3)
    lnJ = 0.5 * torch.log(torch.linalg.det(C))
    return 0.5 * MU * (C.diagonal(dim1=-2, dim2=-1).sum(-1) - 3.0 - 2.0 * lnJ) + 0.5 * LAM * lnJ**2


def psi_blatzko(E3, mu=40.0):
    lam = _stretches(2.0 * E3 + torch.eye(3))
    return 0.5 * mu * ((1.0 / lam**2).sum(-1) + 2.0 * lam.prod(-1) - 5.0)


def psi_ogden_dev_vf3(E3, mus=(2.0 * MU / 3.0, MU / 6.0, MU / 6.0), alphas=(1.23, 0.59, 0.18)):
    lam = _stretches(2.0 * E3 + torch.eye(3))
    J = lam.prod(-1)
    lb = lam * J[:, None] ** (-1.0 / 3.0)
    return sum(m / a * ((lb**a).sum(-1) - 3.0) for m, a in zip(mus, alphas)) + BULK * 0.5 * torch.log(J) ** 2


def psi_mooney_rivlin_vf5(E3):
    C = 2.0 * E3 + torch.eye(3)
    I1 = C.diagonal(dim1=-2, dim2=-1).sum(-1)
    I2 = 0.5 * (I1**2 - (C @ C).diagonal(dim1=-2, dim2=-1).sum(-1))
    I3 = torch.linalg.det(C)
    J = torch.sqrt(I3)
    W1, W2 = I1 * I3 ** (-1.0 / 3.0), I2 * I3 ** (-2.0 / 3.0)
    return MU / 2.0 * (W1 - 3.0) + MU / 2.0 * (W2 - 3.0) + BULK * 0.25 * (J**2 - 1.0 - 2.0 * torch.log(J))


LAWS = {
    "svk": (psi_svk, lambda ps: o.Material("svk", LAM, MU, plane_strain=ps)),
    "neohooke": (psi_neohooke, lambda ps: o.Material("neohooke", LAM, MU, plane_strain=ps)),
    "blatzko": (psi_blatzko, lambda ps: o.Material("blatzko", 0.0, 40.0, plane_strain=ps)),
    "ogden_dev+VF3": (psi_ogden_dev_vf3, lambda ps: o.Material("hyperelastic", 0.0, 0.0, plane_strain=ps, hyper=o.Hyper(
        "ogden_dev", ((2.0 * MU / 3.0, MU / 6.0, MU / 6.0), (1.23, 0.59, 0.18)), vf=3, K=BULK))),
    "mooney_rivlin+VF5": (psi_mooney_rivlin_vf5, lambda ps: o.Material("hyperelastic", 0.0, 0.0, plane_strain=ps, hyper=o.Hyper(
        "invariant", ((1, 0), (0, 1), (MU / 2.0, MU / 2.0)), vf=5, K=BULK))),
}


def twin_energy(kind, psi, X, q, m):
    """W_e of every element for the generalised dofs q[e, ndof + m] = (u, alpha)."""
    d, nn = kind.dim, kind.nodes
    ne = X.shape[0]
    u = q[:, : nn * d].reshape(ne, nn, d)
    alpha = q[:, nn * d:]
    pts, wts = kind.rule()
    if m:
        Jt0, _, detJ0 = o._geometry(kind, X, np.full(d, 0.5))
        T0inv = torch.as_tensor(np.linalg.inv(o.transformation_matrix(Jt0) * detJ0[:, None, None]))
    W = torch.zeros(ne)
    for xi, w in zip(pts, wts):
        _, dN = o.shape_functions(d, kind.order, xi)
        _, Jtinv, detJ = o._geometry(kind, X, xi)
        gradN = torch.as_tensor(np.einsum("eji,ai->eaj", Jtinv, dN))
        H = torch.einsum("eac,eaj->ecj", u, gradN)
        Em = 0.5 * (H + H.transpose(-1, -2) + H.transpose(-1, -2) @ H)
        if m:
            M = torch.einsum("epq,qm->epm", T0inv, torch.as_tensor(o.eas_Mhat(d, m, xi))) / torch.as_tensor(detJ)[:, None, None]
            Ev = torch.einsum("epm,em->ep", M, alpha)  # Voigt strain, shear doubled
            add = torch.zeros_like(Em)
            for p, (i, j) in enumerate(o.voigt_pairs(d)):
                if i == j:
                    add[:, i, i] = add[:, i, i] + Ev[:, p]
                else:
                    add[:, i, j] = add[:, i, j] + 0.5 * Ev[:, p]
                    add[:, j, i] = add[:, j, i] + 0.5 * Ev[:, p]
            Em = Em + add
        E3 = 0.5 * (_embed3(2.0 * Em + torch.eye(d), d) - torch.eye(3))
        W = W + psi(E3) * torch.as_tensor(w * detJ)
    return W


def _derivatives(kind, psi, X, u, alpha):
    ne, m = X.shape[0], alpha.shape[1]
    q0 = torch.as_tensor(np.concatenate([u.reshape(ne, -1), alpha], axis=1)).clone().requires_grad_(True)
    W = twin_energy(kind, psi, X, q0, m)
    (g,) = torch.autograd.grad(W.sum(), q0, create_graph=True)
    n = q0.shape[1]
    Hs = torch.zeros(ne, n, n)
    for i in range(n):  # elements are independent: one backward pass per generalised dof gives column i of every element
        (col,) = torch.autograd.grad(g[:, i].sum(), q0, retain_graph=True)
        Hs[:, :, i] = col
    return W.detach().numpy(), g.detach().numpy(), Hs.numpy()


def _setup(dim, order, seed):
    cells = (2, 2, 1) if dim == 3 else (3, 2)
    if order == 2:
        cells = (2, 1, 1) if dim == 3 else (2, 2)
    mesh = distorted(o.structured_mesh(cells, tuple(float(c) for c in cells), order=order), 0.12, seed)
    rng = np.random.default_rng(seed)
    nn = (order + 1) ** dim
    u = 0.06 * rng.uniform(-1, 1, (mesh.n_elem, nn, dim))
    return mesh, u, rng


@pytest.mark.parametrize("dim,order", [(3, 1), (2, 1), (2, 2), (3, 2)])
@pytest.mark.parametrize("law", sorted(LAWS))
def test_forces_and_stiffness_are_the_derivatives_of_the_energy(law, dim, order):
    psi, make = LAWS[law]
    if order == 2 and law not in ("svk", "neohooke"):
        pytest.skip("the principal-stretch laws are served for Q1 elements")
    mesh, u, _ = _setup(dim, order, 11)
    kind = o.ElementKind(dim, order, "gl")
    q = o.element_quantities(kind, make(dim == 2), mesh.corner_coords, u)
    W, g, H = _derivatives(kind, psi, mesh.corner_coords, u, np.zeros((mesh.n_elem, 0)))
    assert np.abs(q["E"] - W).max() <= 1e-12 * np.abs(W).max()
    assert np.abs(q["R"] - g).max() <= 1e-11 * np.abs(g).max()
    assert np.abs(q["K"] - H).max() <= 1e-10 * np.abs(H).max()


@pytest.mark.parametrize("dim,m,law", [(3, 9, "neohooke"), (3, 21, "svk"), (2, 4, "svk"), (2, 7, "neohooke"), (3, 9, "blatzko"),
                                       (2, 5, "ogden_dev+VF3")])
def test_enhanced_strain_blocks_are_the_derivatives_of_the_enhanced_energy(dim, m, law):
    """EAS::GreenLagrangeStrain (easfunctions/greenlagrangestrain.hh:40-141): with E = E_c(u) + M(xi) alpha the blocks
    K_uu, L, D, R, Rtilde of enhancedassumedstrains.hh:258-348 are the second / first derivatives of W(u, alpha)."""
    psi, make = LAWS[law]
    mesh, u, rng = _setup(dim, 1, 5)
    alpha = 0.01 * rng.uniform(-1, 1, (mesh.n_elem, m))
    kind = o.ElementKind(dim, 1, "gl", m)
    q = o.element_quantities(kind, make(dim == 2), mesh.corner_coords, u, alpha)
    _, g, H = _derivatives(kind, psi, mesh.corner_coords, u, alpha)
    nd = kind.ndof
    Kuu, L, D = H[:, :nd, :nd], H[:, nd:, :nd], H[:, nd:, nd:]
    Ru, Rt = g[:, :nd], g[:, nd:]
    assert np.abs(q["D"] - D).max() <= 1e-10 * np.abs(D).max()
    assert np.abs(q["L"] - L).max() <= 1e-10 * np.abs(L).max()
    assert np.abs(q["Rtilde"] - Rt).max() <= 1e-10 * max(np.abs(Rt).max(), np.abs(Ru).max())
    Dinv = np.linalg.inv(D)
    Kc = Kuu - np.einsum("emi,emn,enj->eij", L, Dinv, L)
    Rc = Ru - np.einsum("emi,emn,en->ei", L, Dinv, Rt)
    assert np.abs(q["K"] - Kc).max() <= 5e-10 * np.abs(Kuu).max()
    assert np.abs(q["R"] - Rc).max() <= 5e-10 * np.abs(Ru).max()
