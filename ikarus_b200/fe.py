"""Host-side mirror of the reference's element description API (no arithmetic on the hot path here).

`makeFE(basis, skills(...))` + `fe.bind(element)` of the reference
(ikarus/finiteelements/fefactory.hh:67-72, febase.hh:114-117) become one `FEContainer`: the
walk over the grid that collects corner coordinates and `FEHelper::globalIndices` per element
(ikarus/finiteelements/fehelper.hh:194-197) is done once by the caller (or by
include/ikarus_b200/deviceflatassembler.hh when DUNE is present) and handed over as arrays.
"""
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import _capi as capi


# ---------------------------------------------------------------------------------- materials
@dataclass(frozen=True)
class LamesFirstParameterAndShearModulus:
    """ikarus/finiteelements/physicshelper.hh:53-57"""
    lambda_: float
    mu: float


def toLamesFirstParameterAndShearModulus(emodul: float, nu: float) -> LamesFirstParameterAndShearModulus:
    """ikarus/finiteelements/physicshelper.hh:276-281"""
    return LamesFirstParameterAndShearModulus(emodul * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)), emodul / (2.0 * (1.0 + nu)))


@dataclass(frozen=True)
class _Material:
    name: str
    code: int
    strain: int
    params: LamesFirstParameterAndShearModulus
    reduced: int = 0  # 0: 3D law, 1: planeStrain wrapper, 2: planeStress wrapper (capi IKB_REDUCE_*)
    reduce_tol: float = 1e-12
    hyper: tuple = None  # MAT_HYPERELASTIC: (deviatoric, n, volumetric, pex, qex, par, ex, K, beta) of ikb_hyperelastic

    def materialParameters(self):
        return self.params


class Materials:
    """Materials::LinearElasticity / StVenantKirchhoff / NeoHooke
    (mechanics/materials/linearelasticity.hh, svk.hh, hyperelastic/neohooke.hh)."""

    @staticmethod
    def LinearElasticity(p):
        return _Material("LinearElasticity", capi.MAT_LINEAR, capi.STRAIN_LINEAR, p)

    @staticmethod
    def StVenantKirchhoff(p):
        return _Material("StVenantKirchhoff", capi.MAT_SVK, capi.STRAIN_GL, p)

    @staticmethod
    def NeoHooke(p):
        return _Material("NeoHooke", capi.MAT_NEOHOOKE, capi.STRAIN_GL, p)

    @staticmethod
    def makeBlatzKo(mu):
        """Materials::makeBlatzKo(mu) = Hyperelastic<Deviatoric<BlatzKoT>, Volumetric<VF0T>>, the principal-stretch
        framework (materials/hyperelastic/factory.hh:34-39, interface.hh:99-232, deviatoric/blatzko.hh:60-92)."""
        return _Material("Hyperelastic (Deviatoric function: BlatzKo, Volumetric function: None)", capi.MAT_BLATZKO,
                         capi.STRAIN_GL, LamesFirstParameterAndShearModulus(0.0, float(mu)))


class VF:
    """A volumetric function VF0 .. VF12 (materials/hyperelastic/volumetric/volumetricfunctions.hh); beta for VF4/7/10."""

    def __init__(self, index: int, beta: float = 0.0):
        self.index, self.beta = int(index), float(beta)


def _pad3(v, cast=float):
    v = [cast(x) for x in v]
    if not 1 <= len(v) <= 3:
        raise NotImplementedError("1 to 3 terms are served on the device")
    return tuple(v + [cast(0)] * (3 - len(v)))


def _hyper(name, dev, n=0, par=(), ex=(), pex=(), qex=(), K=0.0, vf=None):
    vf = vf or VF(0)
    law = (dev, int(n), vf.index, _pad3(pex or [0], int), _pad3(qex or [0], int), _pad3(par or [0.0]), _pad3(ex or [0.0]),
           float(K), vf.beta)
    return _Material(name, capi.MAT_HYPERELASTIC, capi.STRAIN_GL, LamesFirstParameterAndShearModulus(0.0, 0.0), hyper=law)


def makeOgden(mu, alpha, K=0.0, vf=None, tag="total"):
    """Materials::makeOgden<n, tag>(mu, alpha, K, vf) (hyperelastic/factory.hh:17-41, deviatoric/ogden.hh);
    tag = 'total' | 'deviatoric' (PrincipalStretchTags)."""
    if len(mu) != len(alpha):
        raise ValueError("as many exponents as parameters")
    dev = capi.DEV_OGDEN_TOTAL if tag == "total" else capi.DEV_OGDEN_DEVIATORIC
    return _hyper(f"Hyperelastic (Ogden n = {len(mu)}, {tag})", dev, len(mu), par=mu, ex=alpha, K=K, vf=vf)


def makeInvariantBased(mu, pex, qex, K=0.0, vf=None):
    """Materials::makeInvariantBased<n>(mu, pex, qex, K, vf) (factory.hh:43-67, deviatoric/invariantbased.hh)."""
    if any(int(p) == 0 and int(q) == 0 for p, q in zip(pex, qex)):
        raise ValueError("the exponents p_i and q_i should not be zero at the same time")  # invariantbased.hh:216-221
    return _hyper(f"Hyperelastic (InvariantBased n = {len(mu)})", capi.DEV_INVARIANT_BASED, len(mu), par=mu, pex=pex,
                  qex=qex, K=K, vf=vf)


def makeMooneyRivlin(mu, K=0.0, vf=None):
    """Materials::makeMooneyRivlin (factory.hh:69-88): InvariantBased<2> with p = (1, 0), q = (0, 1)."""
    return makeInvariantBased(mu, (1, 0), (0, 1), K, vf)


def makeYeoh(mu, K=0.0, vf=None):
    """Materials::makeYeoh (factory.hh:90-109): InvariantBased<3> with p = (1, 2, 3), q = 0."""
    return makeInvariantBased(mu, (1, 2, 3), (0, 0, 0), K, vf)


def makeArrudaBoyce(mu, lambdaM, K=0.0, vf=None):
    """Materials::makeArrudaBoyce({mu, lambdaM}, K, vf) (factory.hh:111-131, deviatoric/arrudaboyce.hh)."""
    return _hyper("Hyperelastic (ArrudaBoyce)", capi.DEV_ARRUDA_BOYCE, par=(mu, lambdaM), K=K, vf=vf)


def makeGent(mu, Jm, K=0.0, vf=None):
    """Materials::makeGent({mu, Jm}, K, vf) (factory.hh:133-153, deviatoric/gent.hh)."""
    return _hyper("Hyperelastic (Gent)", capi.DEV_GENT, par=(mu, Jm), K=K, vf=vf)


def makePureVolumetric(vf, K):
    """Materials::makePureVolumetric(vf, K) (factory.hh:155-165)."""
    return _hyper("Hyperelastic (pure volumetric)", capi.DEV_NONE, K=K, vf=vf)


for _f in (makeOgden, makeInvariantBased, makeMooneyRivlin, makeYeoh, makeArrudaBoyce, makeGent, makePureVolumetric):
    setattr(Materials, _f.__name__, staticmethod(_f))
Materials.VF = VF


def planeStrain(mat: _Material) -> _Material:
    """Materials::planeStrain (mechanics/materials/vanishingstrain.hh:147-198)."""
    return _Material(mat.name, mat.code, mat.strain, mat.params, 1, hyper=mat.hyper)


def planeStress(mat: _Material, tol: float = 1e-12) -> _Material:
    """Materials::planeStress(mat, tol) = VanishingStress with S33 = S23 = S13 = 0
    (mechanics/materials/vanishingstress.hh:35-230, 252-260)."""
    return _Material(mat.name, mat.code, mat.strain, mat.params, 2, float(tol), hyper=mat.hyper)


# ------------------------------------------------------------------------------------- skills
@dataclass(frozen=True)
class _Solid:
    strain: int
    material: _Material


# EAS::LinearStrain / GreenLagrangeStrain / DisplacementGradient / DisplacementGradientTransposed
# (mechanics/strainenhancements/easfunctions/*.hh) -> IKB_EAS_*
_EAS_FUNCTIONS = {"GreenLagrangeStrain": capi.EAS_STRAIN, "LinearStrain": capi.EAS_STRAIN,
                  "DisplacementGradient": capi.EAS_DISPLACEMENT_GRADIENT,
                  "DisplacementGradientTransposed": capi.EAS_DISPLACEMENT_GRADIENT_TRANSPOSED}


@dataclass(frozen=True)
class _EAS:
    m: int
    enhanced: str = "GreenLagrangeStrain"

    @property
    def function(self):
        return _EAS_FUNCTIONS[self.enhanced]


@dataclass(frozen=True)
class _VolumeLoad:
    fn: Callable  # f(x[dim], lambda) -> force density [dim]


@dataclass(frozen=True)
class _NeumannLoad:
    fn: Callable  # t(x[dim], lambda) -> traction [dim]
    faces: tuple  # ((element index, face id), ...); DUNE cube faces: 2k -> xi_k = 0, 2k+1 -> xi_k = 1


def linearElastic(mat):
    """mechanics/linearelastic.hh linearElastic(mat): small strains; needs a linear-strain material."""
    if mat.strain != capi.STRAIN_LINEAR:
        raise TypeError("linearElastic needs a material with StrainTags::linear")
    return _Solid(capi.STRAIN_LINEAR, mat)


def nonLinearElastic(mat):
    """mechanics/nonlinearelastic.hh nonLinearElastic(mat): Green-Lagrange strains."""
    if mat.strain != capi.STRAIN_GL:
        raise TypeError("nonLinearElastic needs a material with StrainTags::greenLagrangian")
    return _Solid(capi.STRAIN_GL, mat)


def eas(numberOfInternalVariables=0, enhanced="GreenLagrangeStrain"):
    """mechanics/enhancedassumedstrains.hh eas<ES>(m) with ES in {LinearStrain, GreenLagrangeStrain, DisplacementGradient,
    DisplacementGradientTransposed} (the last two with m = dim*dim: H4 / H9)."""
    if enhanced not in _EAS_FUNCTIONS:
        raise NotImplementedError(f"unknown EAS enhancement {enhanced}")
    return _EAS(int(numberOfInternalVariables), enhanced)


def volumeLoad(fn):
    """mechanics/loads/volume.hh volumeLoad<dim>(f): f(x, lambda) sampled on the host."""
    return _VolumeLoad(fn)


def neumannBoundaryLoad(faces, fn):
    """mechanics/loads/traction.hh neumannBoundaryLoad(&patch, t): t(x, lambda) on the boundary faces of the patch,
    given here as (element, face) pairs; sampled on the host with the face rule of order = basis order (:122)."""
    return _NeumannLoad(fn, tuple((int(e), int(f)) for e, f in faces))


def skills(*s):
    """finiteelements/mixin.hh:354-357"""
    return tuple(s)


# ------------------------------------------------------------------- host shape functions (loads only)
def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def _lag1d(order, xi):
    if order == 1:
        return np.array([1.0 - xi, xi])
    return np.array([2.0 * (xi - 0.5) * (xi - 1.0), 4.0 * xi * (1.0 - xi), 2.0 * xi * (xi - 0.5)])


def _shape(dim, order, xi):
    n1 = order + 1
    v = [_lag1d(order, xi[k]) for k in range(dim)]
    return np.array([np.prod([v[k][(a // n1**k) % n1] for k in range(dim)]) for a in range(n1**dim)])


def _dshape_q1(dim, xi):
    out = np.zeros((2**dim, dim))
    for a in range(2**dim):
        for i in range(dim):
            t = 1.0 if (a >> i) & 1 else -1.0
            for k in range(dim):
                if k != i:
                    t *= xi[k] if (a >> k) & 1 else 1.0 - xi[k]
            out[a, i] = t
    return out


@dataclass
class FEContainer:
    """A bound container of identical finite elements: the `std::vector<FE>` the reference hands to
    makeSparseFlatAssembler (assembler/simpleassemblers.hh:174-177)."""
    dim: int
    order: int
    n_dof: int
    corner_coords: np.ndarray  # [nElem, 2^dim, dim]
    elem_dofs: np.ndarray  # [nElem, nodes*dim]
    solid: _Solid = None
    eas: Optional[_EAS] = None
    loads: tuple = ()
    extra: dict = field(default_factory=dict)

    def __len__(self):
        return self.corner_coords.shape[0]

    @property
    def nodes(self):
        return (self.order + 1) ** self.dim

    def numberOfInternalVariables(self):
        return self.eas.m if self.eas else 0

    def sample_external_load(self, lam=1.0):
        """R_e -= N_i f(x_gp, lambda) detJ w summed over elements (loads/volume.hh:86-106) -> fext[n_dof]."""
        fext = np.zeros(self.n_dof)
        for s in self.loads:
            if isinstance(s, _NeumannLoad):
                self._sample_traction(s, lam, fext)
        vl = [s for s in self.loads if isinstance(s, _VolumeLoad)]
        if not vl:
            return fext
        d, order = self.dim, self.order
        x1, w1 = _gauss01(order + 1)
        import itertools
        X = self.corner_coords
        for idx in itertools.product(range(order + 1), repeat=d):
            xi = np.array([x1[i] for i in idx])
            w = np.prod([w1[i] for i in idx])
            N = _shape(d, order, xi)
            Ng = _shape(d, 1, xi)
            dNg = _dshape_q1(d, xi)
            Jt = np.einsum("ci,ecj->eij", dNg, X)
            detJ = np.abs(np.linalg.det(Jt))
            xg = np.einsum("c,ecj->ej", Ng, X)
            f = np.zeros_like(xg)
            for s in vl:
                f += np.array([np.asarray(s.fn(p, lam), float) for p in xg])
            contrib = (N[None, :, None] * f[:, None, :] * (detJ * w)[:, None, None]).reshape(len(self), -1)
            np.add.at(fext, self.elem_dofs.ravel(), contrib.ravel())
        return fext


def _face_points(dim, order, face):
    """Quadrature points (in element reference coordinates), weights and the in-face directions of a cube face for the
    Gauss rule of polynomial order `order` (1 -> midpoint, 2 -> 2 points per direction)."""
    import itertools
    k, side = face // 2, face % 2
    x1, w1 = _gauss01(order // 2 + 1)
    dirs = [j for j in range(dim) if j != k]
    pts = []
    for idx in itertools.product(range(len(x1)), repeat=dim - 1):
        xi = np.zeros(dim)
        xi[k] = float(side)
        w = 1.0
        for j, i in zip(dirs, idx):
            xi[j] = x1[i]
            w *= w1[i]
        pts.append((xi, w))
    return pts, dirs


def _sample_traction_impl(self, load, lam, fext):
    """R_e -= N_i t(x, lambda) w detJ_face over the listed faces (loads/traction.hh:107-138)."""
    d, order = self.dim, self.order
    for e, face in load.faces:
        X = self.corner_coords[e]
        pts, dirs = _face_points(d, order, face)
        for xi, w in pts:
            N = _shape(d, order, xi)
            Ng = _shape(d, 1, xi)
            dNg = _dshape_q1(d, xi)
            Jt = np.einsum("ci,cj->ij", dNg, X)  # Jt[i][j] = dx_j/dxi_i
            if d == 2:
                area = np.linalg.norm(Jt[dirs[0]])
            else:
                area = np.linalg.norm(np.cross(Jt[dirs[0]], Jt[dirs[1]]))
            x = Ng @ X
            t = np.asarray(load.fn(x, lam), float)
            contrib = (N[:, None] * t[None, :] * (w * area)).reshape(-1)
            np.add.at(fext, self.elem_dofs[e], contrib)


FEContainer._sample_traction = _sample_traction_impl


def makeFE(basis, sk, corner_coords=None, elem_dofs=None):
    """makeFE(basisHandler, skills(...)) (finiteelements/fefactory.hh:67-72) followed by bind over all
    grid elements.  `basis` is a dict/obj with dim, order, n_dof."""
    dim, order, n_dof = basis["dim"], basis["order"], basis["n_dof"]
    solid = [s for s in sk if isinstance(s, _Solid)]
    if len(solid) != 1:
        raise TypeError("exactly one solid skill (linearElastic / nonLinearElastic) is required")
    e = [s for s in sk if isinstance(s, _EAS)]
    loads = tuple(s for s in sk if isinstance(s, (_VolumeLoad, _NeumannLoad)))
    mat = solid[0].material
    if dim == 2 and not mat.reduced:
        raise TypeError("2D elements need a reduced material (planeStrain or planeStress)")
    if dim == 3 and mat.reduced:
        raise TypeError("3D elements need a full 3D material")
    return FEContainer(dim, order, n_dof, np.ascontiguousarray(corner_coords, float),
                       np.ascontiguousarray(elem_dofs, np.int64), solid[0], e[0] if e else None, loads)


class DirichletValues:
    """utils/dirichletvalues.hh:73-311: the flags plus (with `nodeCoords` given) the inhomogeneous boundary functions.
    `nodeCoords` are the positions of the Lagrange nodes in global node numbering ([nNodes, dim]); interpolation of a
    nodal function into the power basis is point evaluation there."""

    def __init__(self, n_dof, nodeCoords=None, layout="interleaved"):
        self._flags = np.zeros(int(n_dof), dtype=bool)
        self._coords = None if nodeCoords is None else np.asarray(nodeCoords, float)
        self._layout = layout
        self._functions = []  # (value(x, lam), derivative(x, lam))

    def fixDOFs(self, f):
        f(self._flags)

    def setSingleDOF(self, i, flag=True):
        self._flags[i] = flag

    def fixIthDOF(self, i):
        self._flags[i] = True

    def isConstrained(self, i):
        return bool(self._flags[i])

    def fixedDOFsize(self):
        return int(self._flags.sum())

    def size(self):
        return self._flags.shape[0]

    def container(self):
        return self._flags

    # ---- inhomogeneous values (dirichletvalues.hh:214-302)
    def storeInhomogeneousBoundaryCondition(self, f, lambda_=1.0, derivative=None):
        """`f(globalCoord, lambda) -> dim values`.  The reference differentiates f in lambda with autodiff (:215-219);
        here `derivative(globalCoord, lambda)` may be given, otherwise a complex step is used (exact to rounding for
        functions built from arithmetic and numpy ufuncs)."""
        if self._coords is None:
            raise RuntimeError("DirichletValues needs nodeCoords to interpolate inhomogeneous boundary functions")
        if derivative is None:
            def derivative(x, lam, _f=f):
                return np.imag(np.asarray(_f(x, complex(lam, 1e-30)), dtype=complex)) / 1e-30
        self._functions.append((f, derivative))
        inc = self.evaluateInhomogeneousBoundaryCondition(lambda_)  # setInhomogeneousBoundaryConditionFlag (:296-302)
        self._flags[inc != 0.0] = True

    def _interpolate(self, which, lam):
        nn, dim = self._coords.shape
        out = np.zeros(self.size())
        for fn in self._functions:
            vals = np.array([np.real(np.asarray(fn[which](x, lam))).astype(float) for x in self._coords])
            out += vals.reshape(-1) if self._layout == "interleaved" else vals.T.reshape(-1)
        return out

    def evaluateInhomogeneousBoundaryCondition(self, lam):
        return self._interpolate(0, lam)

    def evaluateInhomogeneousBoundaryConditionDerivative(self, lam):
        return self._interpolate(1, lam)

    def hasInhomogeneousBoundaryConditions(self):
        return bool(self._functions)

    def setZeroAtConstrainedDofs(self, x):
        x[self._flags] = 0.0
