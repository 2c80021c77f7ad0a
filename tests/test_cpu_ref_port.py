"""The C port (oracle/cpu_ref.c, the CPU baseline of bench.py) against the numpy oracle."""
import numpy as np
import pytest

import cpu_ref
import ikarus_oracle as o
from problems import distorted


@pytest.mark.parametrize("dim,matk", [(3, "neohooke"), (3, "svk"), (3, "linear"), (2, "neohooke"), (2, "svk"), (2, "linear")])
def test_port_matches_numpy_oracle(dim, matk):
    cells = (3, 2, 2) if dim == 3 else (4, 3)
    mesh = distorted(o.structured_mesh(cells, tuple(float(c) for c in cells)), 0.15, 3)
    lam, mu = o.lame_from_E_nu(1000.0, 0.3)
    mat = o.Material(matk, lam, mu, dim == 2)
    kind = o.ElementKind(dim, 1, "linear" if matk == "linear" else "gl")
    flags = np.zeros(mesh.n_nodes * dim, dtype=bool)
    ref = o.FlatAssembler(mesh, kind, mat, flags)
    rng = np.random.default_rng(1)
    d = 0.05 * rng.uniform(-1, 1, ref.n)
    outer, inner = ref.pattern("raw")
    nd = kind.ndof
    lin = o.linear_indices(ref.elem_dofs, outer, inner)  # [e][r*nd+c]
    lin_cm = lin.reshape(-1, nd, nd).transpose(0, 2, 1).reshape(-1, nd * nd)  # reference order: for c, for r
    for nthreads in (1, 4):
        vals, R = cpu_ref.assemble(dim, matk, lam, mu, mesh.corner_coords, ref.elem_dofs, lin_cm, d, inner.shape[0],
                                   nthreads=nthreads)
        vr = ref.matrix_values(d, 0.0, "raw")
        Rr = ref.vector(d, 0.0, "raw")
        assert np.abs(vals - vr).max() <= 1e-12 * np.abs(vr).max()
        assert np.abs(R - Rr).max() <= 1e-12 * np.abs(Rr).max()
    if matk == "neohooke":
        # the "cpu_opt" baseline (factored tangent, fused K+R sweep): same values from a different derivation
        for nthreads in (1, 4):
            vo, Ro = cpu_ref.assemble_opt(dim, matk, lam, mu, mesh.corner_coords, ref.elem_dofs, lin_cm, d, inner.shape[0],
                                          nthreads=nthreads)
            assert np.abs(vo - vr).max() <= 1e-12 * np.abs(vr).max()
            assert np.abs(Ro - Rr).max() <= 1e-12 * np.abs(Rr).max()
    else:
        with pytest.raises(NotImplementedError):
            cpu_ref.assemble_opt(dim, matk, lam, mu, mesh.corner_coords, ref.elem_dofs, lin_cm, d, inner.shape[0])
    u = d[ref.elem_dofs[0]]
    K, Re = cpu_ref.element(dim, matk, lam, mu, mesh.corner_coords[0], u)
    q = o.element_quantities(kind, mat, mesh.corner_coords[:1], u.reshape(1, kind.nodes, dim))
    assert np.abs(K - q["K"][0]).max() <= 1e-12 * np.abs(K).max()
