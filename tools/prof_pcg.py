import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ikarus_b200 as ik
from ikarus_b200 import meshes
cells = (128, 32, 32); H = 1 / 32
slab = meshes.structured_q1(cells, tuple(c * H for c in cells))
p = ik.toLamesFirstParameterAndShearModulus(emodul=1000.0, nu=0.3)
fes = ik.makeFE(dict(dim=3, order=1, n_dof=slab.n_dof), ik.skills(ik.nonLinearElastic(ik.Materials.NeoHooke(p))), slab.corner_coords, slab.elem_dofs)
dv = ik.DirichletValues(slab.n_dof); dv.container()[:] = meshes.clamp_face_flags(cells, 0, 0)
asm = ik.SparseFlatAssembler(fes, dv, mode="resident")
d = 0.05 * H * np.random.default_rng(42).uniform(-1, 1, slab.n_dof)
req = ik.FERequirements(d, 0.0); asm.bind(req, ik.elastoStatics, ik.DBCOption.Full)
A = asm.matrix(); R = asm.vector()
ls = ik.DeviceLinearSolver(1e-8, maxIter=int(sys.argv[1]) if len(sys.argv) > 1 else 40)
x = ls(-R, A)
print("its", ls.lastIterations, ls.lastRelRes)
