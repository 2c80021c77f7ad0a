"""Scratch diagnostic: device vs oracle error of the displacement-gradient EAS kernel as a function of nu and h."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ikarus_b200 as ik
import ikarus_oracle as o
from devproblems import device_assembler

def run(nu, h, fn, ascale, m=9, eas_fn=None):
    pc = (6, 6, 3)
    mesh = o.structured_mesh(pc, tuple(c * h for c in pc))
    lam, mu = o.lame_from_E_nu(1000.0, nu)
    mat = o.Material("neohooke", lam, mu)
    kind = o.ElementKind(3, 1, "gl", m, eas_function=fn)
    flags = np.zeros(mesh.n_nodes * 3, dtype=bool)
    rng = np.random.default_rng(44)
    d = 0.05 * h * rng.uniform(-1, 1, flags.shape[0])
    hp = 2 if fn != "strain" else 8
    alpha = ascale * 0.01 * h**hp * rng.uniform(-1, 1, (mesh.n_elem, m))
    ref = o.FlatAssembler(mesh, kind, mat, flags); ref.alpha = alpha.copy()
    dev = device_assembler(mesh, kind, mat, flags); dev.setInternalVariables(alpha)
    K = dev.matrix(ik.FERequirements(d, 0.0), ik.MatrixAffordance.stiffness, ik.DBCOption.Raw)
    Kr = ref.matrix_values(d, 0.0, "raw")
    outer, _ = ref.pattern("raw")
    rows = np.repeat(np.arange(outer.shape[0] - 1), np.diff(outer))
    rowmax = np.zeros(rows.max() + 1); np.maximum.at(rowmax, rows, np.abs(Kr))
    return float((np.abs(K.data - Kr) / rowmax[rows]).max())

for fn, m in (("dg", 9), ("dgt", 9), ("strain", 21), ("strain", 9)):
    for nu in (0.3, 0.49, 0.499):
        print(fn, m, "nu", nu, "h=1/96 erow %.2e" % run(nu, 1 / 96, fn, 1.0, m), " alpha=0: %.2e" % run(nu, 1 / 96, fn, 0.0, m), " h=1: %.2e" % run(nu, 1.0, fn, 1.0, m), flush=True)
