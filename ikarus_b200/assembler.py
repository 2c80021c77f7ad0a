"""Host-side mirror of the reference's flat assemblers over the C-ABI.

Models Concepts::MatrixFlatAssembler (ikarus/utils/concepts.hh:517-585): bind(), matrix(),
vector(), scalar(), size(), reducedSize(), constraintsBelow(), isConstrained(),
createFullVector(), createReducedVector() with the semantics of
ikarus/assembler/interface.hh:51-462 and simpleassemblers.inl.  All arithmetic runs in
libikb200.so on the GPU; there is no CPU path.
"""
import copy
import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np

from . import _capi as capi
from .fe import DirichletValues, FEContainer


class InvalidStateException(RuntimeError):
    """Dune::InvalidStateException (assembler/interface.hh:188-193, 229-256)."""


class NotImplementedInReference(NotImplementedError):
    """Dune::NotImplemented"""


class MaterialError(FloatingPointError):
    """The reference aborts on det C <= 0 (materials/materialhelpers.hh:120-126); here it is an exception."""


class ResultTypes(enum.IntEnum):
    """finiteelements/feresulttypes.hh (the stress results of the solid elements)"""
    linearStress = 0
    PK2Stress = 1
    linearStressFull = 2
    PK2StressFull = 3
    kirchhoffStress = 4
    cauchyStress = 5


class DBCOption(enum.IntEnum):
    """assembler/dirichletbcenforcement.hh"""
    Raw = capi.DBC_RAW
    Reduced = capi.DBC_REDUCED
    Full = capi.DBC_FULL


class ScalarAffordance(enum.IntEnum):
    noAffordance = 0
    mechanicalPotentialEnergy = 1


class VectorAffordance(enum.IntEnum):
    noAffordance = 0
    forces = 1


class MatrixAffordance(enum.IntEnum):
    noAffordance = 0
    stiffness = 1


@dataclass(frozen=True)
class AffordanceCollection:
    """finiteelements/ferequirements.hh:104-169"""
    scalar: ScalarAffordance = ScalarAffordance.noAffordance
    vector: VectorAffordance = VectorAffordance.noAffordance
    matrix: MatrixAffordance = MatrixAffordance.noAffordance

    def scalarAffordance(self):
        return self.scalar

    def vectorAffordance(self):
        return self.vector

    def matrixAffordance(self):
        return self.matrix


elastoStatics = AffordanceCollection(ScalarAffordance.mechanicalPotentialEnergy, VectorAffordance.forces,
                                     MatrixAffordance.stiffness)


class FERequirements:
    """finiteelements/ferequirements.hh:222-407: global solution vector + load factor."""

    def __init__(self, d=None, lam=0.0, n=None):
        self._d = np.zeros(n) if d is None else np.ascontiguousarray(d, dtype=np.float64)
        self._lam = float(lam)

    def globalSolution(self):
        return self._d

    def parameter(self):
        return self._lam

    def setParameter(self, lam):
        self._lam = float(lam)

    def insertGlobalSolution(self, d):
        self._d = np.ascontiguousarray(d, dtype=np.float64)

    def populated(self):
        return self._d is not None


class DeviceMatrix:
    """Handle to a matrix resident on the GPU (what matrix() returns in resident mode; the
    reference leaves the matrix type unconstrained, utils/concepts.hh:575-585)."""

    def __init__(self, assembler, dbc):
        self.assembler, self.dbc = assembler, DBCOption(dbc)

    @property
    def shape(self):
        n = self.assembler.reducedSize() if self.dbc == DBCOption.Reduced else self.assembler.size()
        return (n, n)

    def to_scipy(self):
        return self.assembler._download_matrix(self.dbc)

    def matvec(self, x):
        return self.assembler._spmv(self.dbc, x)


class _FlatAssemblerBase:
    def __init__(self, fes: FEContainer, dirichletValues: DirichletValues, device=-1, mode="mirror", rows=None):
        if mode not in ("mirror", "resident"):
            raise ValueError(mode)
        self._lib = capi.load()
        self._fes = fes
        self._dv = copy.copy(dirichletValues)  # the reference copies DirichletValues (:260): flags AND functions
        self._dv._flags = np.array(dirichletValues.container(), dtype=bool)
        if hasattr(dirichletValues, "_functions"):
            self._dv._functions = list(dirichletValues._functions)
        self._mode = mode
        flags = self._dv.container()
        self._n = int(fes.n_dof)
        self._cbelow = np.concatenate([[0], np.cumsum(flags.astype(np.int64))[:-1]])  # interface.hh:51-62
        self._nred = self._n - int(flags.sum())
        self._req = None
        self._aff = None
        self._dbc = None
        self._vec_cb, self._mat_cb, self._scal_cb = [], [], []
        mat = fes.solid.material
        desc = capi.Desc(capi.IKB_ABI_VERSION, fes.dim, fes.order, fes.solid.strain, mat.code, int(mat.reduced),
                         fes.numberOfInternalVariables(), device, mat.params.lambda_, mat.params.mu, len(fes), self._n,
                         float(mat.reduce_tol), fes.eas.function if fes.eas else 0, 0)
        self._h = C.c_void_p()
        rc = self._lib.ikb_create(C.byref(self._h), C.byref(desc))
        if rc == capi.ENOTIMPL:
            raise NotImplementedInReference("EAS is only supported for Q1 and H1 elements with m in {4,5,7} / {9,21} "
                                            "(displacement-gradient forms: nonlinear element, m = 4 / 9)")
        if rc != 0:
            raise ValueError(f"ikb_create failed with code {rc}")
        if mat.hyper is not None:
            dev, n, vfi, pex, qex, par, ex, K, beta = mat.hyper
            law = capi.Hyperelastic(dev, n, vfi, 0, (C.c_int32 * 3)(*pex), (C.c_int32 * 3)(*qex), (C.c_double * 3)(*par),
                                    (C.c_double * 3)(*ex), K, beta)
            self._check(self._lib.ikb_set_hyperelastic(self._h, C.byref(law)))
        self._check(self._lib.ikb_upload_mesh(self._h, capi.ptr(fes.corner_coords), capi.ptr(fes.elem_dofs)))
        self._flags_u8 = np.ascontiguousarray(flags, dtype=np.uint8)
        self._check(self._lib.ikb_upload_dirichlet(self._h, capi.ptr(self._flags_u8)))
        if rows is not None:
            # element-partitioned run: this handle owns the rows of nodes [rows[0], rows[1])
            self._check(self._lib.ikb_set_row_ownership(self._h, int(rows[0]), int(rows[1])))
        self._rows = rows
        self._check(self._lib.ikb_build_pattern(self._h))
        self._fext_lambda = None
        self._last_d = None
        self._user_fext, self._user_scales = None, True
        self._loads_proportional = True
        self._refresh_loads(1.0, force=True)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc == 0:
            return
        buf = C.create_string_buffer(512)
        self._lib.ikb_last_error(self._h, buf, 512)
        msg = buf.value.decode()
        if rc == capi.EMATERIAL:
            raise MaterialError(msg)
        if rc == capi.ENOTIMPL:
            raise NotImplementedInReference(msg)
        if rc == capi.ESTATE:
            raise InvalidStateException(msg)
        raise RuntimeError(f"libikb200 error {rc}: {msg}")

    def __del__(self):
        try:
            if self._h:
                self._lib.ikb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _refresh_loads(self, lam, force=False):
        """Volume and Neumann loads are host callbacks f(x, lambda) (loads/volume.hh:26, loads/traction.hh:29).  They are
        sampled on the host with the reference's rules (FEContainer.sample_external_load).  A load that is proportional
        to the load factor -- every reference test -- is sampled once at lambda = 1 and scaled on the device; any other
        dependence on lambda is re-sampled whenever the load factor changes and uploaded unscaled
        (R_e -= N f(x, lambda) detJ w, E_e -= u . f(x, lambda) detJ w: volume.hh:67-106, traction.hh:70-138)."""
        if not self._fes.loads:
            return
        if force:
            self._fext = self._fes.sample_external_load(1.0)
            f2 = self._fes.sample_external_load(2.0)
            f0 = self._fes.sample_external_load(0.0)
            scale = np.abs(self._fext).max(initial=0.0)
            self._loads_proportional = bool(np.allclose(f2, 2.0 * self._fext, rtol=1e-12, atol=1e-14 * scale)
                                            and np.allclose(f0, 0.0, rtol=0.0, atol=1e-14 * scale))
            self._fext_lambda = None
            if self._loads_proportional:
                self._check(self._lib.ikb_set_external_load(self._h, capi.ptr(self._fext), 1))
                return
        if not self._loads_proportional and lam != self._fext_lambda:
            f = self._fes.sample_external_load(lam)
            if self._user_fext is not None:
                f = f + (lam if self._user_scales else 1.0) * self._user_fext
            self._fext_now = f  # (kept alive until the copy has run)
            self._check(self._lib.ikb_set_external_load(self._h, capi.ptr(f), 0))
            self._check(self._lib.ikb_sync(self._h))
            self._fext_lambda = lam

    def setExternalLoad(self, fext, scalesWithLambda=True):
        """Extra nodal load vector, lambda-proportional by default (what the reference tests add through an
        AssemblerManipulator vector callback, tests/src/testcantileverbeam.hh:56-80)."""
        f = capi.as_f64(fext)
        self._user_fext, self._user_scales = f.copy(), bool(scalesWithLambda)
        if self._fes.loads and not self._loads_proportional:
            self._fext_lambda = None  # re-sampled together with the skills' loads at the next bind
            return
        if self._fes.loads:
            if not scalesWithLambda:
                raise NotImplementedInReference("a constant nodal load next to lambda-proportional load skills")
            f = f + self._fext
        self._fext_user = f
        self._check(self._lib.ikb_set_external_load(self._h, capi.ptr(f), int(scalesWithLambda)))

    def _push(self, req):
        d = req.globalSolution()
        if d.shape[0] != self._n:
            raise ValueError("The solution vector you passed has the wrong dimensions.")
        if self._last_d is None or not np.array_equal(self._last_d, d):
            self._check(self._lib.ikb_set_solution(self._h, capi.ptr(d)))
            self._check(self._lib.ikb_sync(self._h))
            self._last_d = d.copy()
        self._check(self._lib.ikb_set_parameter(self._h, req.parameter()))
        self._refresh_loads(req.parameter())

    # ------------------------------------------------------------------ FlatAssemblerBase
    def size(self):
        return self._n

    def reducedSize(self):
        return self._nred

    def isConstrained(self, i):
        return bool(self._dv.container()[i])

    def constraintsBelow(self, i):
        return int(self._cbelow[i])

    def estimateOfConnectivity(self):
        return len(self._fes) * 8

    def finiteElements(self):
        return self._fes

    def dirichletValues(self):
        return self._dv

    def createFullVector(self, reduced):
        reduced = np.asarray(reduced, float)
        assert reduced.shape[0] == self._nred, "The reduced vector you passed has the wrong dimensions."
        full = np.zeros(self._n)
        full[~self._dv.container()] = reduced
        return full

    def createReducedVector(self, full):
        full = np.asarray(full, float)
        assert full.shape[0] == self._n, "The full vector you passed has the wrong dimensions."
        return full[~self._dv.container()].copy()

    def calculateAt(self, resultType, req, local):
        """`fe.calculateAt<RT>(req, local)` of EVERY element at once (mechanics/nonlinearelastic.hh:237-271,
        linearelastic.hh, enhancedassumedstrains.hh:127-187).  `local`: one position or [npts, dim] positions in the
        reference element.  Returns [nElem, npts, ncomp] (Voigt), evaluated on the device."""
        self._push(req)
        loc = capi.as_f64(np.atleast_2d(local))
        dim = self._fes.dim
        if loc.shape[1] != dim:
            raise ValueError("local positions must have `dim` coordinates")
        full = resultType in (ResultTypes.linearStressFull, ResultTypes.PK2StressFull)
        ncomp = 6 if full else dim * (dim + 1) // 2
        out = np.empty((len(self._fes), loc.shape[0], ncomp))
        self._check(self._lib.ikb_calculate_at(self._h, int(resultType), capi.ptr(loc), loc.shape[0], capi.ptr(out)))
        return out

    def obtainForcesDueToIDBC(self):
        """utils::obtainForcesDueToIDBC (utils/functionhelper.hh:170-185) for the bound requirement and DBC option:
        K_raw * d(d_D)/d(lambda), zeroed at constrained dofs (Full) or reduced; the SpMV runs on the device."""
        dbc = self.dBCOption()
        self._push(self._req)
        inc = capi.as_f64(self._dv.evaluateInhomogeneousBoundaryConditionDerivative(1.0))
        out = np.empty(self._n if dbc == DBCOption.Full else self._nred)
        self._check(self._lib.ikb_idbc_forces(self._h, int(dbc), capi.ptr(inc), capi.ptr(out)))
        return out

    def bind(self, req=None, affordanceCollection=None, dbcOption=None):
        """interface.hh:150-180: stores a reference to the requirement (identity is tested in
        tests/src/testassembler.cpp:305-311) and copies of the enums."""
        if req is not None:
            self._req = req
        if affordanceCollection is not None:
            self._aff = affordanceCollection
        if dbcOption is not None:
            self._dbc = DBCOption(dbcOption)

    def bound(self):
        return self.boundToRequirement() and self.boundToAffordanceCollection() and self.boundToDBCOption()

    def boundToRequirement(self):
        return self._req is not None

    def boundToAffordanceCollection(self):
        return self._aff is not None

    def boundToDBCOption(self):
        return self._dbc is not None

    def requirement(self):
        if self._req is None:
            raise InvalidStateException("The requirement can only be obtained after binding")
        return self._req

    def affordanceCollection(self):
        if self._aff is None:
            raise InvalidStateException("The affordance can only be obtained after binding")
        return self._aff

    def dBCOption(self):
        if self._dbc is None:
            raise InvalidStateException("The dBCOption can only be obtained after binding")
        return self._dbc

    # ------------------------------------------------------------------ AssemblerManipulator hooks
    def bindVectorFunction(self, f):
        """assembler/assemblermanipulatorfuser.hh:242-385: f(assembler, req, affordance, dbc, vec) mutates
        the returned vector on the host."""
        self._vec_cb.append(f)

    def bindMatrixFunction(self, f):
        self._mat_cb.append(f)

    def bindScalarFunction(self, f):
        self._scal_cb.append(f)

    # ------------------------------------------------------------------ scalar / vector
    def _args(self, req, aff, dbc, kind):
        if req is None:
            req = self.requirement()
        if aff is None:
            c = self.affordanceCollection()
            aff = {"s": c.scalar, "v": c.vector, "m": c.matrix}[kind]
        if dbc is None and kind != "s":
            dbc = self.dBCOption()
        return req, aff, (None if dbc is None else DBCOption(dbc))

    def scalar(self, req=None, affordance=None):
        req, aff, _ = self._args(req, affordance, None, "s")
        if aff != ScalarAffordance.mechanicalPotentialEnergy:
            raise NotImplementedInReference(f"ScalarAffordance not implemented: {aff}")
        self._push(req)
        self._check(self._lib.ikb_assemble(self._h, capi.SCALAR, capi.DBC_RAW))
        e = C.c_double()
        self._check(self._lib.ikb_get_scalar(self._h, C.byref(e)))
        val = e.value
        for f in self._scal_cb:
            val = f(self, req, aff, val)
        return val

    def _assemble(self, req, what, dbc):
        self._push(req)
        self._check(self._lib.ikb_assemble(self._h, what, int(dbc)))

    def vector(self, req=None, affordance=None, dbcOption=None):
        req, aff, dbc = self._args(req, affordance, dbcOption, "v")
        if aff != VectorAffordance.forces:
            raise NotImplementedInReference(f"VectorAffordance not implemented: {aff}")
        # K and R come out of the same fused sweep; ask for both when a matrix is likely next
        what = capi.VECTOR | (capi.MATRIX if self._fuse_matrix(dbc) else 0)
        self._assemble(req, what, dbc)
        n = self._nred if dbc == DBCOption.Reduced else self._n
        out = np.empty(n)
        self._check(self._lib.ikb_get_vector(self._h, int(dbc), capi.ptr(out)))
        for f in self._vec_cb:
            f(self, req, aff, dbc, out)
        return out

    def vectorNorm(self, dbcOption=None):
        dbc = self.dBCOption() if dbcOption is None else DBCOption(dbcOption)
        v = C.c_double()
        self._check(self._lib.ikb_vector_norm(self._h, int(dbc), C.byref(v)))
        return v.value

    def _fuse_matrix(self, dbc):
        return False

    # ------------------------------------------------------------------ EAS state
    def updateInternalVariables(self, req, correction):
        """One call replaces the per-element CORRECTION_UPDATED listeners
        (enhancedassumedstrains.hh:350-359, controlroutinefactory.hh:43-46)."""
        if not self._fes.numberOfInternalVariables():
            return
        c = capi.as_f64(correction)
        if c.shape[0] != self._n:
            raise NotImplementedInReference(
                "Solution vector and correction vector should be of the same size. Check if DBCOption::Full is used. "
                f"The sizes are {self._n} and {c.shape[0]}")
        self._push(req)
        self._check(self._lib.ikb_eas_update(self._h, capi.ptr(c)))

    def internalVariables(self):
        m = self._fes.numberOfInternalVariables()
        a = np.zeros((len(self._fes), m))
        if m:
            self._check(self._lib.ikb_eas_get_alpha(self._h, capi.ptr(a)))
        return a

    def setInternalVariables(self, alpha):
        a = capi.as_f64(alpha)
        self._check(self._lib.ikb_eas_set_alpha(self._h, capi.ptr(a)))

    # ------------------------------------------------------------------ diagnostics
    def launchCount(self):
        n = C.c_int64()
        self._lib.ikb_launch_count(self._h, C.byref(n))
        return n.value

    def timePhase(self, phase, dbc, reps):
        ms = C.c_float()
        self._check(self._lib.ikb_time_phase(self._h, phase.encode(), int(dbc), reps, C.byref(ms)))
        return ms.value

    def stream(self):
        s = C.c_void_p()
        self._lib.ikb_stream(self._h, C.byref(s))
        return s.value


class SparseFlatAssembler(_FlatAssemblerBase):
    """SparseFlatAssembler (assembler/simpleassemblers.hh:106-177)."""

    def _fuse_matrix(self, dbc):
        return self._aff is not None and self._aff.matrix == MatrixAffordance.stiffness

    def pattern(self, dbcOption=DBCOption.Full):
        """(outerIndex, innerIndex) exactly as Eigen's compressed storage holds them."""
        dbc = DBCOption(dbcOption)
        rows, nnz = C.c_int64(), C.c_int64()
        self._check(self._lib.ikb_pattern_nnz(self._h, int(dbc), C.byref(rows), C.byref(nnz)))
        outer = np.zeros(rows.value + 1, dtype=np.int64)
        inner = np.zeros(nnz.value, dtype=np.int32)
        self._check(self._lib.ikb_get_pattern(self._h, int(dbc), capi.ptr(outer), capi.ptr(inner)))
        return outer, inner

    def elementLinearIndices(self, e):
        nd = self._fes.elem_dofs.shape[1]
        out = np.zeros(nd * nd, dtype=np.int64)
        self._check(self._lib.ikb_element_linear_indices(self._h, int(e), capi.ptr(out)))
        return out

    def _download_matrix(self, dbc):
        import scipy.sparse as sp

        outer, inner = self.pattern(dbc)
        vals = np.empty(inner.shape[0])
        self._check(self._lib.ikb_get_matrix_values(self._h, int(dbc), capi.ptr(vals)))
        n = outer.shape[0] - 1
        # Eigen holds this as CSC; pattern and values are symmetric, so the CSR view is identical
        return sp.csr_matrix((vals, inner, outer), shape=(n, n))

    def _spmv(self, dbc, x):
        x = capi.as_f64(x)
        y = np.empty(self._nred if dbc == DBCOption.Reduced else self._n)
        self._check(self._lib.ikb_spmv(self._h, int(dbc), capi.ptr(x), capi.ptr(y)))
        return y

    def matrix(self, req=None, affordance=None, dbcOption=None):
        req, aff, dbc = self._args(req, affordance, dbcOption, "m")
        if aff != MatrixAffordance.stiffness:
            raise NotImplementedInReference(f"MatrixAffordance not implemented: {aff}")
        self._assemble(req, capi.MATRIX | capi.VECTOR, dbc)
        if self._mode == "resident" and not self._mat_cb:
            return DeviceMatrix(self, dbc)
        A = self._download_matrix(dbc)
        for f in self._mat_cb:
            f(self, req, aff, dbc, A)
        return A


class DenseFlatAssembler(_FlatAssemblerBase):
    """DenseFlatAssembler (assembler/simpleassemblers.hh:188-231): dense column-major matrix."""

    def matrix(self, req=None, affordance=None, dbcOption=None):
        req, aff, dbc = self._args(req, affordance, dbcOption, "m")
        if aff != MatrixAffordance.stiffness:
            raise NotImplementedInReference(f"MatrixAffordance not implemented: {aff}")
        self._assemble(req, capi.MATRIX, dbc)
        n = self._nred if dbc == DBCOption.Reduced else self._n
        out = np.zeros((n, n), order="F")
        self._check(self._lib.ikb_get_dense_matrix(self._h, int(dbc), capi.ptr(out)))
        for f in self._mat_cb:
            f(self, req, aff, dbc, out)
        return out


def makeSparseFlatAssembler(fes, dirichletValues, **kw):
    """assembler/simpleassemblers.hh:174-177"""
    return SparseFlatAssembler(fes, dirichletValues, **kw)


def makeDenseFlatAssembler(fes, dirichletValues, **kw):
    """assembler/simpleassemblers.hh:228-231"""
    return DenseFlatAssembler(fes, dirichletValues, **kw)
