"""Builds the same problem for the oracle and for the device assembler."""
import numpy as np

import ikarus_b200 as ik
import ikarus_oracle as o

_MAT = {"linear": ik.Materials.LinearElasticity, "svk": ik.Materials.StVenantKirchhoff, "neohooke": ik.Materials.NeoHooke}


def _device_hyper(h):
    """oracle Hyper -> the factory call of the Python mirror (hyperelastic/factory.hh)."""
    M = ik.Materials
    vf = M.VF(h.vf, h.beta)
    d, p = h.dev, h.params
    if d == "blatzko":
        return M.makeBlatzKo(p[0])
    if d in ("ogden_total", "ogden_dev"):
        return M.makeOgden(p[0], p[1], h.K, vf, "total" if d == "ogden_total" else "deviatoric")
    if d == "invariant":
        return M.makeInvariantBased(p[2], p[0], p[1], h.K, vf)
    if d == "arrudaboyce":
        return M.makeArrudaBoyce(p[0], p[1], h.K, vf)
    if d == "gent":
        return M.makeGent(p[0], p[1], h.K, vf)
    return M.makePureVolumetric(vf, h.K)


def device_assembler(mesh, kind, mat, flags, layout="interleaved", fext=None, dense=False, mode="mirror", volume=None):
    p = ik.fe.LamesFirstParameterAndShearModulus(mat.lam, mat.mu)
    if mat.kind == "hyperelastic":
        m = _device_hyper(mat.hyper)
    else:
        m = ik.Materials.makeBlatzKo(mat.mu) if mat.kind == "blatzko" else _MAT[mat.kind](p)
    if mat.plane_strain:
        m = ik.planeStrain(m)
    if getattr(mat, "plane_stress", False):
        m = ik.planeStress(m, mat.ps_tol)
    solid = ik.linearElastic(m) if kind.strain == "linear" else ik.nonLinearElastic(m)
    sk = [solid]
    if kind.eas_m:
        fn = {"strain": "GreenLagrangeStrain" if kind.strain == "gl" else "LinearStrain", "dg": "DisplacementGradient",
              "dgt": "DisplacementGradientTransposed"}[getattr(kind, "eas_function", "strain")]
        sk.append(ik.eas(kind.eas_m, fn))
    if volume is not None:
        sk.append(ik.fe.volumeLoad(volume))
    n_dof = mesh.n_nodes * mesh.dim
    fes = ik.makeFE(dict(dim=mesh.dim, order=mesh.order, n_dof=n_dof), ik.skills(*sk), mesh.corner_coords,
                    mesh.elem_dofs(layout))
    dv = ik.DirichletValues(n_dof)
    dv.container()[:] = flags
    asm = (ik.makeDenseFlatAssembler if dense else ik.makeSparseFlatAssembler)(fes, dv, mode=mode)
    if fext is not None:
        asm.setExternalLoad(fext)
    return asm


def entry_error(dev, ref, rows=None):
    """SURVEY.md 8d parity norm: |dev-ref|_ij <= tol*max(|ref_ij|, 1e-3*max_row_i|ref|)."""
    dev, ref = np.asarray(dev), np.asarray(ref)
    if rows is None:
        scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max(initial=0.0))
    else:
        rowmax = np.zeros(rows.max() + 1)
        np.maximum.at(rowmax, rows, np.abs(ref))
        scale = np.maximum(np.abs(ref), 1e-3 * rowmax[rows])
    scale = np.where(scale == 0.0, 1.0, scale)
    return float((np.abs(dev - ref) / scale).max(initial=0.0))
