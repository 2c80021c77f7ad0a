"""Callers of the hot path, restated on the host so the parity tests can drive the device assembler
exactly as the reference's NewtonRaphson / LoadControl do.  These are NOT part of the accelerated path;
the reference's own templates (solver/nonlinearsolver/newtonraphson.hh, controlroutines/loadcontrol.inl)
drive include/ikarus_b200/deviceflatassembler.hh unchanged when DUNE is available.
"""
import ctypes as C
import enum
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
    # NonlinearSolverFactory::withIDBCForceFunction(assembler) (nonlinearsolverfactory.hh:88-104): callable
    # `assembler -> forces due to inhomogeneous Dirichlet values`; None = utils::IDBCForceDefault (no such forces)
    idbcForceFunction: object = None


def obtainForcesDueToIDBC(assembler):
    """utils/functionhelper.hh:170-185"""
    return assembler.obtainForcesDueToIDBC()


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
        self.idbcForceFunction = cfg.idbcForceFunction
        self.listeners = []

    def syncParameterAndGlobalSolution(self, req):
        """Impl::updateFunctor with SyncFERequirements (nonlinearsolverfactory.hh:45-54): prescribed values of the
        inhomogeneous Dirichlet functions at the current lambda overwrite the solution."""
        inc = self.assembler.dirichletValues().evaluateInhomogeneousBoundaryCondition(req.parameter())
        nz = inc != 0.0
        req.globalSolution()[nz] = inc[nz]

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
        if self.idbcForceFunction is not None:  # newtonraphson.hh:214-217
            rx = rx + self.idbcForceFunction(asm) * stepSize
            rNorm = float(np.linalg.norm(rx))
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
            if self.idbcForceFunction is not None:  # :237-238
                self.syncParameterAndGlobalSolution(req)
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
            if si.iterations == 0 and getattr(self.nls, "idbcForceFunction", None) is not None:
                self.nls.syncParameterAndGlobalSolution(req)  # loadcontrol.inl:44-46
            info.solverInfos.append(si)
            info.totalIterations += si.iterations
            for f in self.listeners:
                f(ls, req)
            if not si.success:
                return info
        info.success = True
        return info


# ------------------------------------------------------------------------------------------------------------------
# TrustRegion with the truncated CG on the device
class PreConditioner(enum.Enum):
    """solver/nonlinearsolver/trustregion.hh:31-36"""
    IncompleteCholesky = 0
    IdentityPreconditioner = 1
    DiagonalPreconditioner = 2


@dataclass
class TRSettings:
    """solver/nonlinearsolver/trustregion.hh:38-52"""
    verbosity: int = 5
    maxtime: float = float("inf")
    minIter: int = 3
    maxIter: int = 1000
    debug: int = 0
    grad_tol: float = 1e-6
    corr_tol: float = 1e-6
    rho_prime: float = 0.01
    useRand: bool = False
    rho_reg: float = 1e6
    Delta_bar: float = float("inf")
    Delta0: float = 10.0


TCG_STOP_REASON = ("negative curvature", "exceeded trust region", "reached target residual-kappa (linear)",
                   "reached target residual-theta (superlinear)", "maximum inner iterations", "model increased")


class DeviceTruncatedCG:
    """Eigen::TruncatedConjugateGradient (linearalgebra/truncatedconjugategradient.hh:170-282) on the resident matrix:
    `solveWithGuess(rhs, 0)` under the trust-region radius `Delta`.  Identity and Diagonal preconditioners."""

    def __init__(self, preConditioner=PreConditioner.DiagonalPreconditioner, kappa=0.1, theta=1.0, mininner=1,
                 maxIterations=None, tolerance=None):
        if preConditioner == PreConditioner.IncompleteCholesky:
            raise NotImplementedError("IncompleteCholesky is a sequential factorisation: use the Diagonal or Identity "
                                      "preconditioner on the device, or mirror mode with the reference's own tCG")
        self.precond = (capi.PRECOND_DIAGONAL if preConditioner == PreConditioner.DiagonalPreconditioner
                        else capi.PRECOND_IDENTITY)
        self.kappa, self.theta, self.mininner = kappa, theta, mininner
        self.maxIterations, self.tolerance = maxIterations, tolerance
        self.info = None

    def solve(self, Ax, rhs, Delta, download=True):
        if not isinstance(Ax, DeviceMatrix):
            raise TypeError("DeviceTruncatedCG needs the assembler in resident mode (matrix() -> DeviceMatrix)")
        asm = Ax.assembler
        n = Ax.shape[0]
        info = capi.TcgInfo(delta=float(Delta), kappa=self.kappa, theta=self.theta, mininner=self.mininner,
                            max_iters=int(self.maxIterations or 0), tol=float(self.tolerance or 0.0),
                            precond=self.precond)
        b = None if rhs is None else capi.as_f64(rhs)
        x = np.empty(n) if download else None
        asm._check(asm._lib.ikb_tcg_solve(asm._h, int(Ax.dbc), capi.ptr(b), capi.ptr(x), C.byref(info)))
        self.info = info
        return x


class TrustRegion:
    """solver/nonlinearsolver/trustregion.hh:226-430 on an assembler bound with
    AffordanceCollections.elastoStatics: energy = scalar(), gradient = vector(), Hessian = matrix() (resident),
    inner problem by DeviceTruncatedCG.  The random correction predictor (useRand) is not mirrored."""

    def __init__(self, assembler, settings: TRSettings = None, preConditioner=PreConditioner.DiagonalPreconditioner):
        self.assembler = assembler
        self.settings = settings or TRSettings()
        if self.settings.useRand:
            raise NotImplementedError("TRSettings.useRand")
        s = self.settings
        assert s.rho_prime < 0.25 and s.Delta_bar > 0 and 0 < s.Delta0 < s.Delta_bar  # setup() (:218-224)
        self.tcg = DeviceTruncatedCG(preConditioner)
        self.listeners = []
        self.history = []
        self.innerIterSum = 0

    def _notify(self, msg, **kw):
        for f in self.listeners:
            f(msg, **kw)

    def solve(self, req, stepSize=0.0):
        asm, s = self.assembler, self.settings
        eps_ = 0.0001220703125  # sqrt(sqrt(machine precision)) (:563)
        info = NonLinearSolverInformation()
        self.history, self.innerIterSum = [], 0
        self._notify("INIT")
        e = float(asm.scalar(req))
        g = np.array(asm.vector(req))
        h = asm.matrix(req)
        energy, gradNorm, etaNorm, outer = e, float(np.linalg.norm(g)), 0.0, 0
        Delta = s.Delta0
        rejected = 0
        stop = None
        d = req.globalSolution()
        dbc = asm.dBCOption()
        while True:
            if gradNorm < s.grad_tol and outer != 0:
                stop = "gradientNormTolReached"
                break
            if etaNorm < s.corr_tol and outer != 0:
                stop = "correctionNormTolReached"
                break
            if outer >= s.maxIter:
                stop = "maximumIterationsReached"
                break
            self._notify("ITERATION_STARTED")
            eta = self.tcg.solve(h, None, Delta)  # H eta = -g with the resident gradient
            ti = self.tcg.info
            self.innerIterSum += ti.iterations
            etaNorm = ti.eta_norm
            full = asm.createFullVector(eta) if dbc == DBCOption.Reduced and eta.shape[0] != d.shape[0] else eta
            d += full
            e = float(asm.scalar(req))
            proposal = e
            rhonum = energy - proposal
            rhoden = -(ti.g_dot_eta + 0.5 * ti.eta_h_eta)
            rhoReg = max(1.0, abs(energy)) * eps_ * s.rho_reg
            rhonum += rhoReg
            rhoden += rhoReg
            modelDecreased = rhoden > 0.0
            rho = rhonum / rhoden
            rho = -1.0 if rho < 0.0 else rho
            a_, b_ = energy - proposal, -1e-12  # Dune::FloatCmp::ge (:346)
            energyDecreased = a_ > b_ or abs(a_ - b_) <= 8 * np.finfo(float).eps * max(abs(a_), abs(b_))
            tr = "   "
            if rho < 1e-4 or not modelDecreased or np.isnan(rho) or not energyDecreased:
                tr = "TR-"
                Delta /= 4.0
            elif rho > 0.99 and ti.stop_reason in (0, 1):
                tr = "TR+"
                Delta = min(3.5 * Delta, s.Delta_bar)
            if modelDecreased and rho > s.rho_prime and energyDecreased:
                accept = True
                rejected = 0
            else:
                accept = False
                Delta = Delta / 2 if rejected >= 5 else min(Delta, etaNorm / 2.0)
                rejected += 1
            outer += 1
            info.correctionNorm, info.residualNorm = etaNorm, gradNorm
            self._notify("CORRECTION_UPDATED", correction=full)
            self.history.append(dict(accept=accept, tr=tr, inner=int(ti.iterations), stop=int(ti.stop_reason), rho=rho,
                                     energy=energy, proposal=proposal, Delta=Delta, eta_norm=etaNorm))
            if accept:
                energy = proposal
                self._notify("SOLUTION_CHANGED")
            else:
                d -= full
            e = float(asm.scalar(req))
            g = np.array(asm.vector(req))
            h = asm.matrix(req)
            gradNorm = float(np.linalg.norm(g))
            self._notify("ITERATION_ENDED")
        info.success = stop in ("correctionNormTolReached", "gradientNormTolReached")
        info.iterations = outer
        info.residualNorm = gradNorm
        self.stopReason = stop
        self.energy = e
        if info.success:
            self._notify("FINISHED_SUCESSFULLY")
        return info
