"""CPU tests of the host-side callers (ikarus_b200/solvers.py) against the oracle's own Newton/LoadControl,
using an oracle-backed stand-in for the assembler (no GPU, no libikb200 compute calls)."""
import numpy as np

import ikarus_oracle as o
from ikarus_b200.assembler import DBCOption
from ikarus_b200.solvers import LoadControl, LoadControlConfig, NewtonRaphson, NewtonRaphsonConfig, NRSettings
from problems import cantilever


class _Req:
    def __init__(self, d, lam):
        self._d, self._lam = np.array(d, float), float(lam)

    def globalSolution(self):
        return self._d

    def parameter(self):
        return self._lam

    def setParameter(self, lam):
        self._lam = float(lam)


class _OracleAssembler:
    """Duck-types what NewtonRaphson needs from a flat assembler (vector/matrix/dBCOption/createFullVector/...)."""

    def __init__(self, ref, dbc):
        self.ref, self._dbc, self.eas_calls = ref, dbc, 0

    def dBCOption(self):
        return self._dbc

    def reducedSize(self):
        return self.ref.reduced_size()

    def createFullVector(self, v):
        return self.ref.create_full_vector(v)

    def _mode(self):
        return {DBCOption.Full: "full", DBCOption.Reduced: "reduced", DBCOption.Raw: "raw"}[self._dbc]

    def vector(self, req):
        return self.ref.vector(req.globalSolution(), req.parameter(), self._mode())

    def matrix(self, req):
        return self.ref.matrix(req.globalSolution(), req.parameter(), self._mode())

    def updateInternalVariables(self, req, correction):
        self.eas_calls += 1
        self.ref.update_eas(req.globalSolution(), correction)


def _run(dbc, mat="neohooke", m=0):
    mesh, kind, material, flags, fext = cantilever(3, mat, m, cells=(6, 1, 1))
    ref = o.FlatAssembler(mesh, kind, material, flags, fext=fext)
    mode = "full" if dbc == DBCOption.Full else "reduced"
    dr, lamr, inf = o.load_control(o.FlatAssembler(mesh, kind, material, flags, fext=fext), np.zeros(ref.n), 4, 0.0, 1.0,
                                   tol=1e-9, dbc=mode)
    asm = _OracleAssembler(ref, dbc)
    req = _Req(np.zeros(ref.n), 0.0)
    nr = NewtonRaphson(asm, NewtonRaphsonConfig(NRSettings(tol=1e-9)))
    events = []
    nr.listeners.append(lambda msg, **kw: events.append(msg))
    info = LoadControl(nr, LoadControlConfig(4, 0.0, 1.0)).run(req)
    return info, inf, req, dr, asm, events


def test_newton_loadcontrol_mirror_reproduces_oracle_full():
    info, inf, req, dr, asm, events = _run(DBCOption.Full)
    assert info.success and info.totalIterations == inf["total_iterations"]
    assert [s.iterations for s in info.solverInfos] == inf["per_step"]
    assert np.abs(req.globalSolution() - dr).max() < 1e-12
    assert abs(req.parameter() - 1.0) < 1e-14
    # CORRECTION_UPDATED is notified before SOLUTION_CHANGED in every iteration (newtonraphson.hh:230-240)
    assert events[:2] == ["CORRECTION_UPDATED", "SOLUTION_CHANGED"] and len(events) == 2 * info.totalIterations


def test_newton_loadcontrol_mirror_reduced_expands_correction():
    info, inf, req, dr, asm, _ = _run(DBCOption.Reduced, "svk")
    assert info.success and info.totalIterations == inf["total_iterations"]
    assert np.abs(req.globalSolution() - dr).max() < 1e-12


def test_eas_update_is_called_once_per_iteration_with_full_correction():
    info, inf, req, dr, asm, _ = _run(DBCOption.Full, "neohooke", 9)
    assert info.success and asm.eas_calls == info.totalIterations == inf["total_iterations"]
    assert np.abs(req.globalSolution() - dr).max() < 1e-10


def test_newton_reports_failure_at_max_iterations():
    mesh, kind, material, flags, fext = cantilever(3, "neohooke", 0, cells=(4, 1, 1))
    ref = o.FlatAssembler(mesh, kind, material, flags, fext=fext)
    asm = _OracleAssembler(ref, DBCOption.Full)
    req = _Req(np.zeros(ref.n), 1.0)
    nr = NewtonRaphson(asm, NewtonRaphsonConfig(NRSettings(tol=1e-14, maxIter=2)))
    info = nr.solve(req)
    assert not info.success and info.iterations == 2  # newtonraphson.hh:251-252
    lc_info = LoadControl(nr, LoadControlConfig(3, 0.0, 1.0)).run(_Req(np.zeros(ref.n), 0.5))
    assert not lc_info.success  # LoadControl aborts the run (loadcontrol.inl:46-47)
