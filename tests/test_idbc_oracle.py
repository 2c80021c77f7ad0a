"""Inhomogeneous Dirichlet path of the oracle (obtainForcesDueToIDBC, NewtonRaphson sync, LoadControl) against the
reference's elastic strip results (tests/src/testinhomogeneousdbc.cpp:22-33, testelasticstrip.hh)."""
import numpy as np
import pytest

import ikarus_oracle as o
from problems import ELASTIC_STRIP_EXPECTED, elastic_strip


@pytest.mark.parametrize("dbc", ["full", "reduced"])
@pytest.mark.parametrize("case", [("svk", 1), ("neohooke", 1), ("svk", 2), ("neohooke", 2)], ids=str)
def test_elastic_strip_load_control(case, dbc):
    if case[1] == 2 and dbc == "reduced":
        pytest.skip("Q2 is pinned in Full mode only (keeps the CPU suite short)")
    mesh, kind, mat, flags, value, deriv, probe = elastic_strip(*case)
    assert flags.sum() == 3 * (case[1] * 10 + 1)
    idbc = o.InhomogeneousDirichlet(mesh, [(value, deriv)])
    flags = idbc.flag(flags)
    assert flags.sum() == 4 * (case[1] * 10 + 1)  # testelasticstrip.hh:101-102
    asm = o.FlatAssembler(mesh, kind, mat, flags, "interleaved")
    d, lam, info = o.load_control(asm, np.zeros(flags.shape[0]), 1, 0.0, 1.0, tol=1e-8, dbc=dbc, idbc=idbc)
    its, disp = ELASTIC_STRIP_EXPECTED[case]
    assert info["success"] and info["total_iterations"] == its
    assert abs(d[probe] - disp) < 1e-8 and lam == 1.0
    inc = idbc.values(lam)
    assert np.abs(d[inc != 0] - inc[inc != 0]).max() < 1e-8
    assert np.linalg.norm(asm.vector(d, lam, dbc)) < 1e-8
