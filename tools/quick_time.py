"""Scratch timing of the assembly phases on the C2 mesh (not the bench; numbers for tuning only)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ikarus_b200 as ik
import ikarus_oracle as o
from devproblems import device_assembler

cells = tuple(int(c) for c in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 32, 32)))
matk = sys.argv[4] if len(sys.argv) > 4 else "neohooke"
t0 = time.time()
mesh = o.structured_mesh(cells, (cells[0] / 32.0, cells[1] / 32.0, cells[2] / 32.0))
lam, mu = o.lame_from_E_nu(1000.0, 0.3)
strain = "linear" if matk == "linear" else "gl"
mat = o.Material(matk, lam, mu); kind = o.ElementKind(3, 1, strain)
flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 0, 0.0))
t1 = time.time()
dev = device_assembler(mesh, kind, mat, flags, mode="resident")
t2 = time.time()
rng = np.random.default_rng(42)
d = 0.05 / 32 * rng.uniform(-1, 1, flags.shape[0])
req = ik.FERequirements(d, 0.0)
dev.bind(req, ik.elastoStatics, ik.DBCOption.Full)
A = dev.matrix(); R = dev.vector()
print(f"mesh {cells} {matk}: {mesh.n_elem} elements, {flags.shape[0]} dofs; mesh gen {t1-t0:.2f}s setup {t2-t1:.2f}s |R|={np.linalg.norm(R):.6e}")
for ph in ("fused", "elements", "gather", "spmv", "dfma_peak", "dmma_peak"):
    if ph == "fused" and os.environ.get("IKB_FUSED", "0") != "1":
        continue
    for _ in range(2):
        ms = dev.timePhase(ph, ik.DBCOption.Full, 20 if not ph.endswith("_peak") else 3)
    if ph.endswith("_peak"):
        # dfma: 8 chains x 2048 FMAs per thread; dmma: 8 tiles x 2048 m8n8k4 (256 FMA) per warp
        fl = 148 * 16 * 256 * 2048 * 16 if ph == "dfma_peak" else 148 * 16 * 8 * 2048 * 8 * 512
        print(f"{ph}: {ms:.4f} ms  -> {fl/ms/1e9:.2f} TFLOP/s FP64")
    else:
        print(f"{ph}: {ms:.4f} ms  -> {mesh.n_elem/ms/1e3:.1f} Melem/s")
# whole K+R sweep through ikb_assemble (what bench.py times), wall clock around a synchronised batch
from ikarus_b200 import _capi as capi
lib, h = dev._lib, dev._h
for rep in range(2):
    lib.ikb_sync(h)
    t = time.time()
    for _ in range(50):
        lib.ikb_invalidate(h)
        dev._check(lib.ikb_assemble(h, capi.MATRIX | capi.VECTOR, capi.DBC_FULL))
    lib.ikb_sync(h)
    t = (time.time() - t) / 50
print(f"step: {t*1e3:.4f} ms  -> {mesh.n_elem/t/1e6:.1f} Melem/s (IKB_CHUNKS={os.environ.get('IKB_CHUNKS', 'default')})")
if os.environ.get("QT_NO_PCG"):
    sys.exit(0)
ls = ik.DeviceLinearSolver(1e-10)
t = time.time(); x = ls(-R, A); t = time.time() - t
print(f"pcg: {ls.lastIterations} its relres {ls.lastRelRes:.2e} in {t*1e3:.1f} ms -> {t*1e3/max(ls.lastIterations,1):.3f} ms/it")
