"""Times the assembly phases for the other BASELINE.json configs (C1, C3, C4, C5-sized slab) on one GPU.
Not the bench: numbers for DESIGN.md / profiles only."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ikarus_b200 as ik
from ikarus_b200 import meshes, _capi as capi
import ikarus_oracle as o

def lame(E, nu): return ik.toLamesFirstParameterAndShearModulus(emodul=E, nu=nu)

def run(name, cells, bbox, order, skills, dim, seed, amp, eas_m=0, flop=None, alpha_seed=None):
    t0 = time.time()
    if order == 1:
        slab = meshes.structured_q1(cells, bbox)
        cc, ed, n_dof, h = slab.corner_coords, slab.elem_dofs, slab.n_dof, slab.h
        flags = meshes.clamp_face_flags(cells, 0 if dim == 3 and name != "C3" else 0, 0)
    else:
        m = o.structured_mesh(cells, bbox, order=2)
        cc, ed, n_dof, h = m.corner_coords, m.elem_dofs(), m.n_nodes * dim, min(b / c for b, c in zip(bbox, cells))
        flags = o.fix_nodes(m, o.boundary_nodes(m, dim - 1, 0.0))
    fes = ik.makeFE(dict(dim=dim, order=order, n_dof=n_dof), skills, cc, ed)
    dv = ik.DirichletValues(n_dof); dv.container()[:] = flags
    t1 = time.time()
    asm = ik.SparseFlatAssembler(fes, dv, mode="resident")
    t2 = time.time()
    d = amp * h * np.random.default_rng(seed).uniform(-1, 1, n_dof); d[flags] = 0
    if eas_m and alpha_seed is not None:
        asm.setInternalVariables(0.01 * np.random.default_rng(alpha_seed).uniform(-1, 1, (len(fes), eas_m)))
    req = ik.FERequirements(d, 0.0); asm.bind(req, ik.elastoStatics, ik.DBCOption.Full)
    asm._assemble(req, capi.MATRIX | capi.VECTOR, ik.DBCOption.Full)
    asm._check(asm._lib.ikb_sync(asm._h))
    te = min(asm.timePhase("elements", ik.DBCOption.Full, 10) for _ in range(3))
    tg = min(asm.timePhase("gather", ik.DBCOption.Full, 10) for _ in range(3))
    ts = min(asm.timePhase("spmv", ik.DBCOption.Full, 10) for _ in range(3))
    ne = len(fes)
    import ctypes as C
    rows_c, nnz_c = C.c_int64(), C.c_int64()
    asm._check(asm._lib.ikb_pattern_nnz(asm._h, int(ik.DBCOption.Full), C.byref(rows_c), C.byref(nnz_c)))
    out = dict(config=name, elements=ne, dofs=n_dof, nnz=nnz_c.value, host_mesh_s=round(t1 - t0, 2), device_setup_s=round(t2 - t1, 2),
               elements_ms=te, gather_ms=tg, spmv_ms=ts, K_R_Melem_s=ne / (te + tg) / 1e3)
    if flop: out["canonical_tflops_elem_kernel"] = flop * ne / (te * 1e-3) / 1e12
    print(json.dumps(out), flush=True)
    del asm

which = sys.argv[1:] or ["C1", "C3", "C4", "C4b"]
if "C1" in which:
    p = ik.planeStrain(ik.Materials.LinearElasticity(lame(1.0, 1 / 3)))
    run("C1 Quad4 LE plane strain 64x64", (64, 64), (48.0, 60.0), 1, ik.skills(ik.linearElastic(p)), 2, 1, 0.0, flop=2236)
if "C3" in which:
    n = int(os.environ.get("C3_N", "64"))
    run(f"C3 Hex27 SVK {n}^3", (n, n, n), (1.0, 1.0, 1.0), 2, ik.skills(ik.nonLinearElastic(ik.Materials.StVenantKirchhoff(lame(1000.0, 0.3)))), 3, 43, 0.05, flop=1450764)
if "C4" in which:
    n = int(os.environ.get("C4_N", "96"))
    run(f"C4 Hex8+EAS21 NeoHooke nu=0.499 {n}^3", (n, n, n), (1.0, 1.0, 1.0), 1, ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(lame(1000.0, 0.499))), ik.eas(21)), 3, 44, 0.05, eas_m=21, flop=185961, alpha_seed=None)
if "C4b" in which:
    n = int(os.environ.get("C4_N", "96"))
    run(f"C4' Hex8+EAS9 NeoHooke nu=0.499 {n}^3", (n, n, n), (1.0, 1.0, 1.0), 1, ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(lame(1000.0, 0.499))), ik.eas(9)), 3, 44, 0.05, eas_m=9, flop=102045, alpha_seed=None)
if "C4dg" in which:
    # the displacement-gradient enhancements (H9) on the C4 mesh: first, plain formulation (ikb_elem_easdg.cuh)
    n = int(os.environ.get("C4_N", "96"))
    for fn in ("DisplacementGradient", "DisplacementGradientTransposed"):
        run(f"C4'' Hex8+H9 {fn} NeoHooke nu=0.499 {n}^3", (n, n, n), (1.0, 1.0, 1.0), 1, ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(lame(1000.0, 0.499))), ik.eas(9, fn)), 3, 44, 0.05, eas_m=9, flop=None, alpha_seed=None)
if "C5slab" in which:
    # one rank's share of C5 (256^3 over 8 GPUs): 256x256x32 elements
    run("C5 slab Hex8 NeoHooke 256x256x32", (256, 256, 32), (1.0, 1.0, 0.125), 1, ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(lame(1000.0, 0.3)))), 3, 46, 0.05, flop=59520)
if "C5" in which:
    # the full C5 mesh on ONE GPU: 16.8 M elements, 50.9 M dofs, 4.09e9 non-zeros (> 2^31: needs the 64-bit positions)
    run("C5 Hex8 NeoHooke 256^3 on one GPU", (256, 256, 256), (1.0, 1.0, 1.0), 1, ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(lame(1000.0, 0.3)))), 3, 46, 0.05, flop=59520)
