"""Callers of the hot path, restated on the host so the parity tests can drive the device assembler
exactly as the reference's NewtonRaphson / LoadControl do.  These are NOT part of the accelerated path;
the reference's own templates (solver/nonlinearsolver/newtonraphson.hh, controlroutines/loadcontrol.inl)
drive include/ikarus_b200/deviceflatassembler.hh unchanged when DUNE is available.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi as capi
from .assembler import DBCOption, DeviceMatrix


@dataclass
class NRSettings:
    """solver/nonlinearsolver/newtonraphson.hh:27-32"""
    tol: float = 1e-8
    maxIter: int = 20
    minIter: int = 0


@dataclass
class NewtonRaphsonConfig:
    parameters: NRSettings = field(default_factory=NRSettings)
    linearSolver: object = None


@dataclass
class NonLinearSolverInformation:
    success: bool = False
    residualNorm: float = 0.0
    correctionNorm: float = 0.0
    iterations: int = 0


@dataclass
class ControlInformation:
    success: bool = False
    totalIterations: int = 0
    solverInfos: list = field(default_factory=list)


class DeviceLinearSolver:
    """Callable LS `(rx, Ax) -> K^-1 rx` (the callable seam of NewtonRaphson, newtonraphson.hh:152, 221-226)
    running Jacobi-PCG on the GPU: the analogue of LinearSolver(SolverTypeTag::si_ConjugateGradient)
    (solver/linearsolver/linearsolver.cpp:23-24)."""

    def __init__(self, relTol=1e-13, maxIter=None):
        self.relTol, self.maxIter = relTol, maxIter
        self.lastIterations, self.lastRelRes, self.totalIterations = 0, 0.0, 0

    def __call__(self, rx, Ax):
        if not isinstance(Ax, DeviceMatrix):
            raise TypeError("DeviceLinearSolver needs the assembler in resident mode (matrix() -> DeviceMatrix)")
        asm = Ax.assembler
        n = Ax.shape[0]
        rhs = capi.as_f64(rx)
        x = np.empty(n)
        it, rel = C.c_int(), C.c_double()
        maxit = self.maxIter if self.maxIter is not None else 2 * n  # Eigen default
        asm._check(asm._lib.ikb_pcg_solve(asm._h, int(Ax.dbc), capi.ptr(rhs), capi.ptr(x), self.relTol, int(maxit),
                                          C.byref(it), C.byref(rel)))
        self.lastIterations, self.lastRelRes = it.value, rel.value
        self.totalIterations += it.value
        return x


class SparseDirectSolver:
    """Host direct solve for mirror mode (stands in for sd_UmfPackLU of the reference tests)."""

    def __call__(self, rx, Ax):
        import scipy.sparse.linalg as spla

        return spla.spsolve(Ax.tocsc(), rx)


class NewtonRaphson:
    """solver/nonlinearsolver/newtonraphson.hh:196-257 driving an assembler bound to (req, affordances, dbc)."""

    def __init__(self, assembler, config: NewtonRaphsonConfig = None):
        self.assembler = assembler
        cfg = config or NewtonRaphsonConfig()
        self.settings = cfg.parameters
        self.linearSolver = cfg.linearSolver or SparseDirectSolver()
        self.listeners = []

    def _notify(self, msg, **kw):
        for f in self.listeners:
            f(msg, **kw)

    def solve(self, req, stepSize=0.0):
        asm, s = self.assembler, self.settings
        dbc = asm.dBCOption()
        info = NonLinearSolverInformation(success=True)
        rx = asm.vector(req)
        Ax = asm.matrix(req)
        rNorm = float(np.linalg.norm(rx))
        info.residualNorm = rNorm
        it = 0
        while (rNorm > s.tol and it < s.maxIter) or it < s.minIter:
            correction = -np.asarray(self.linearSolver(rx, Ax))
            info.correctionNorm = float(np.linalg.norm(correction))
            # CORRECTION_UPDATED: EAS internal variables see the correction before x changes (:230-235)
            asm.updateInternalVariables(req, correction if dbc != DBCOption.Reduced else correction)
            self._notify("CORRECTION_UPDATED", correction=correction)
            # update functor of nonlinearsolverfactory.hh:33-56 (Reduced -> Full expansion)
            d = req.globalSolution()
            if dbc == DBCOption.Reduced and correction.shape[0] == asm.reducedSize():
                d += asm.createFullVector(correction)
            else:
                d += correction
            self._notify("SOLUTION_CHANGED")
            rx = asm.vector(req)
            Ax = asm.matrix(req)
            rNorm = float(np.linalg.norm(rx))
            info.residualNorm = rNorm
            it += 1
            info.iterations = it
        if it == s.maxIter:
            info.success = False
        info.iterations = it
        return info


@dataclass
class LoadControlConfig:
    """controlroutines/loadcontrol.hh:29-34"""
    loadSteps: int
    tbegin: float
    tEnd: float


class LoadControl:
    """controlroutines/loadcontrol.inl:21-57"""

    def __init__(self, nonLinearSolver, config: LoadControlConfig):
        self.nls = nonLinearSolver
        self.loadSteps = config.loadSteps
        self.stepSize = (config.tEnd - config.tbegin) / config.loadSteps
        self.listeners = []

    def run(self, req):
        info = ControlInformation()
        si = self.nls.solve(req)  # initial equilibrium check at the current lambda (:31)
        info.solverInfos.append(si)
        info.totalIterations += si.iterations
        if not si.success:
            return info
        for ls in range(self.loadSteps):
            req.setParameter(req.parameter() + self.stepSize)  # predictor (:56)
            si = self.nls.solve(req, self.stepSize)
            info.solverInfos.append(si)
            info.totalIterations += si.iterations
            for f in self.listeners:
                f(ls, req)
            if not si.success:
                return info
        info.success = True
        return info
