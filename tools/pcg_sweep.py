"""Times the resident Jacobi-PCG on the C2 matrix for several SpMV grid sizes (IKB_SPMV_BLOCKS)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ikarus_b200 as ik
from ikarus_b200 import meshes
cells = (128, 32, 32); H = 1 / 32
slab = meshes.structured_q1(cells, tuple(c * H for c in cells))
p = ik.toLamesFirstParameterAndShearModulus(emodul=1000.0, nu=0.3)
d = 0.05 * H * np.random.default_rng(42).uniform(-1, 1, slab.n_dof)
for blocks in [int(a) for a in sys.argv[1:]] or [592]:
    os.environ["IKB_SPMV_BLOCKS"] = str(blocks)
    fes = ik.makeFE(dict(dim=3, order=1, n_dof=slab.n_dof), ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(p))), slab.corner_coords, slab.elem_dofs)
    dv = ik.DirichletValues(slab.n_dof); dv.container()[:] = meshes.clamp_face_flags(cells, 0, 0)
    asm = ik.SparseFlatAssembler(fes, dv, mode="resident")
    asm.bind(ik.FERequirements(d, 0.0), ik.elastoStatics, ik.DBCOption.Full)
    A = asm.matrix(); R = asm.vector()
    ls = ik.DeviceLinearSolver(1e-8)
    for rep in range(3):
        t = time.perf_counter(); x = ls(-R, A); t = time.perf_counter() - t
    print(f"blocks {blocks}: {ls.lastIterations} its relres {ls.lastRelRes:.3e} {t*1e3:.1f} ms -> {t*1e6/ls.lastIterations:.1f} us/it", flush=True)
    del asm, A, fes
