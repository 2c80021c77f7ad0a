#!/usr/bin/env python
"""bench.py -- FP64 K+R assembly throughput (Melem/s) and Newton-step time on B200.

A "step" is one K+R assembly sweep (element kernel + deterministic row gather, DBCOption::Full) over the workload
mesh, through the C-ABI of libikb200.so.

  --gpus 1   BASELINE.json configs[1] (C2): 3D cantilever Hex8 Q1 NeoHooke, YaspGrid 128x32x32, ~420k DOF.
             The same run also times configs[4] (C5, Hex8 NeoHooke 256^3) on the ONE GPU (`c5_single_gpu`), which
             is the same-workload reference for the strong-scaling lines below, and configs[3] (C4, Hex8 + EAS(21)
             NeoHooke nu = 0.499, 96^3; plus its E9 variant) as `c4_single_gpu` with the SURVEY 8d targets beside them.
  --gpus N   BASELINE.json configs[4] (C5): Hex8 Q1 NeoHooke 256x256x256 (~51M DOF, 4.09e9 nnz), the FIXED mesh cut
             into N z-slabs ("scaling": "strong"): every rank owns a contiguous row block, evaluates the elements
             touching it (one ghost element layer, no collective in assembly) and the Newton step runs the
             row-block Jacobi-PCG with halo exchange and all-reduced dot products (`newton_step`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|C2|C5]

--impl reference times the CPU port of the reference's assembly loops (oracle/cpu_ref.c; the reference itself needs
DUNE/Eigen and cannot be built here) on the host cores: all cores (count stated) on the C2 mesh per step at N=1 and on
a 131072-element sub-box of C5 at N>1, plus one single-thread figure (the reference is single-threaded).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EMOD, NU = 1000.0, 0.3
FLOP_PER_ELEM = 59520  # canonical Hex8 NeoHooke K+R flops per element
METRIC = "FP64 K+R assembly throughput (Hex8 NeoHooke, DBCOption::Full)"
UNIT = "Melem/s"

WORKLOADS = {
    # name: cells, bounding box, seed of the state, clamped face (axis, index), slab axis is always z
    "C2": dict(cells=(128, 32, 32), bbox=(4.0, 1.0, 1.0), seed=42, clamp=(0, 0),
               name="C2 cantilever Hex8 Q1 NeoHooke YaspGrid 128x32x32"),
    "C5": dict(cells=(256, 256, 256), bbox=(1.0, 1.0, 1.0), seed=46, clamp=(2, 0),
               name="C5 block Hex8 Q1 NeoHooke YaspGrid 256x256x256"),
}


def workload(name, world):
    """C2W (not a BASELINE config; --workload C2W): one C2 mesh per rank stacked in z (128x32x32N), the latency-bound
    end of the distributed PCG."""
    if name == "C2W":
        return dict(cells=(128, 32, 32 * world), bbox=(4.0, 1.0, 1.0 * world), seed=42, clamp=(0, 0),
                    name=f"C2 slabs: Hex8 Q1 NeoHooke YaspGrid 128x32x{32 * world} (one C2 mesh per rank)")
    return WORKLOADS[name]


def lame():
    return EMOD * NU / ((1 + NU) * (1 - 2 * NU)), EMOD / (2 * (1 + NU))


def synthetic_state(n_dof, h, seed):
    """d = 0.05*h*U(-1,1) (seed 42 for C2, 46 for C5)."""
    return 0.05 * h * np.random.default_rng(seed).uniform(-1.0, 1.0, n_dof)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.2)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for i, nme in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU baseline (oracle port)
_CPU_CACHE = {}


def _cpu_problem(wl, layers):
    """A `layers`-deep sub-box (in z) of the workload mesh with the workload's state: rows of the nodes below the top
    layer are complete, i.e. identical to the same rows of the full mesh."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ikarus_oracle as o

    W = WORKLOADS[wl]
    cells = (W["cells"][0], W["cells"][1], layers)
    key = (wl, layers)
    if key not in _CPU_CACHE:
        hz = W["bbox"][2] / W["cells"][2]
        mesh = o.structured_mesh(cells, (W["bbox"][0], W["bbox"][1], hz * layers))
        ed = mesh.elem_dofs()
        n = mesh.n_nodes * 3
        outer, inner = o.build_pattern(ed, n)
        lin = o.linear_indices(ed, outer, inner).reshape(-1, 24, 24).transpose(0, 2, 1).reshape(-1, 576)
        n_full = 3 * int(np.prod([c + 1 for c in W["cells"]]))
        d = synthetic_state(n_full, min(b / c for b, c in zip(W["bbox"], W["cells"])), W["seed"])[:n]
        _CPU_CACHE[key] = (mesh, ed, np.ascontiguousarray(lin), d, outer, inner)
    return _CPU_CACHE[key]


def cpu_baseline(wl, sample_elems, threads, keep=False, opt=False):
    """Times the C port of the reference's loops (oracle/cpu_ref.c): separate R and K sweeps, tangent per node pair,
    scatter through linear indices.  The one place bench.py executes oracle/ code: as the measured CPU baseline (and,
    with keep=True, as the checker of the device values on the sampled rows)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref

    W = WORKLOADS[wl]
    layers = max(1, min(W["cells"][2], int(round(sample_elems / (W["cells"][0] * W["cells"][1])))))
    mesh, ed, lin, d, outer, inner = _cpu_problem(wl, layers)
    lam, mu = lame()
    nnz = inner.shape[0]
    # opt: the "cpu_opt" baseline of SURVEY.md 8d -- the factored math of the device kernels, one fused K+R sweep, OpenMP
    fn = cpu_ref.assemble_opt if opt else cpu_ref.assemble
    fn(3, "neohooke", lam, mu, mesh.corner_coords[:256], ed[:256], lin[:256], d, nnz, nthreads=threads)
    t0 = time.perf_counter()
    vals, R = fn(3, "neohooke", lam, mu, mesh.corner_coords, ed, lin, d, nnz, nthreads=threads)
    dt = time.perf_counter() - t0
    out = {"value": mesh.n_elem / dt / 1e6, "unit": UNIT, "cores": int(threads), "kind": "port",
           "sample": f"{mesh.n_elem} elements ({mesh.cells[0]}x{mesh.cells[1]}x{layers} sub-box of {wl}"
                     f"{', the whole mesh' if layers == W['cells'][2] else ''}), "
                     f"{'fused K+R sweep with the factored tangent (cpu_opt)' if opt else 'K and R sweeps'}, {dt:.2f} s wall",
           "seconds": dt, "elements": int(mesh.n_elem)}
    if keep:
        out["_check"] = (vals, R, outer, layers)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload if args.workload != "auto" else ("C2" if args.gpus == 1 else "C5")
    threads = host_threads()
    per_step = 131072  # the whole C2 mesh; a 256x256x2 sub-box of C5
    runs = []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(wl, per_step, threads)
        if i >= args.warmup:
            runs.append(b)
    dt = float(np.mean([b["seconds"] for b in runs]))
    v = runs[0]["elements"] / dt / 1e6
    base = dict(runs[-1], value=v)
    base.pop("seconds"), base.pop("elements")
    one = cpu_baseline(wl, 8192, 1)  # the reference itself is single-threaded
    fast = cpu_baseline(wl, per_step, threads, opt=True)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                      "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic",
                      "config": {"workload": WORKLOADS[wl]["name"], "sample_elements_per_step": runs[0]["elements"],
                                 "threads": threads},
                      "cpu_baseline": base,
                      "cpu_baseline_1thread": {"value": one["value"], "unit": UNIT, "cores": 1, "kind": "port",
                                               "sample": one["sample"]},
                      "cpu_baseline_opt": {"value": fast["value"], "unit": UNIT, "cores": threads, "kind": "port",
                                           "sample": fast["sample"]},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
          file=JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------ our arm
def build_handle(wl, rank, world, local):
    import ikarus_b200 as ik
    from ikarus_b200 import meshes

    W = workload(wl, world)
    cells, bbox = W["cells"], W["bbox"]
    if world > 1:
        layers = cells[2] + 1
        lb, le = rank * layers // world, (rank + 1) * layers // world
        slab = meshes.structured_q1(cells, bbox, lb, le)
    else:
        slab = meshes.structured_q1(cells, bbox)
    lam, mu = lame()
    mat = ik.Materials.NeoHooke(ik.fe.LamesFirstParameterAndShearModulus(lam, mu))
    fes = ik.makeFE(dict(dim=3, order=1, n_dof=slab.n_dof), ik.skills(ik.nonLinearElastic(mat)), slab.corner_coords,
                    slab.elem_dofs)
    dv = ik.DirichletValues(slab.n_dof)
    dv.container()[:] = meshes.clamp_face_flags(cells, *W["clamp"])
    asm = ik.SparseFlatAssembler(fes, dv, device=local, mode="resident",
                                 rows=(slab.node_begin, slab.node_end) if world > 1 else None)
    need_lo, need_hi = int(slab.elem_dofs.min()), int(slab.elem_dofs.max()) + 1
    n_local_elems = len(fes)
    slab.corner_coords = slab.elem_dofs = None  # host copies are no longer needed
    return asm, slab, n_local_elems, need_lo, need_hi


def time_sweeps(asm, stream, steps, barrier, torch):
    """Exactly `steps` K+R sweeps of the current state, CUDA events on the launching stream."""
    from ikarus_b200 import _capi as capi

    lib, h = asm._lib, asm._h
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        lib.ikb_invalidate(h)
        asm._check(lib.ikb_assemble(h, capi.MATRIX | capi.VECTOR, capi.DBC_FULL))
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1)


def parity_check(asm, wl, check):
    """Device values of the benched state against the CPU port on the rows of the sampled sub-box (SURVEY.md 8d norm:
    |dev-ref| <= tol*max(|ref_ij|, 1e-3*max_row|ref|))."""
    from ikarus_b200 import _capi as capi

    vals_cpu, R_cpu, outer, layers = check
    W = WORKLOADS[wl]
    lib, h = asm._lib, asm._h
    n_rows = 3 * (W["cells"][0] + 1) * (W["cells"][1] + 1) * layers  # dofs of the nodes below the sub-box's top layer
    nnz_pref = int(outer[n_rows])
    asm._check(lib.ikb_assemble(h, capi.MATRIX | capi.VECTOR, capi.DBC_RAW))
    rows_c, nnz_c = C.c_int64(), C.c_int64()
    asm._check(lib.ikb_pattern_nnz(h, capi.DBC_RAW, C.byref(rows_c), C.byref(nnz_c)))
    vals = np.empty(nnz_c.value)
    R = np.empty(rows_c.value)
    asm._check(lib.ikb_get_matrix_values(h, capi.DBC_RAW, capi.ptr(vals)))
    asm._check(lib.ikb_get_vector(h, capi.DBC_RAW, capi.ptr(R)))
    asm._check(lib.ikb_sync(h))
    a, b = vals[:nnz_pref], vals_cpu[:nnz_pref]
    rows = np.repeat(np.arange(n_rows), np.diff(outer[: n_rows + 1]))
    rowmax = np.zeros(n_rows)
    np.maximum.at(rowmax, rows, np.abs(b))
    scale = np.maximum(np.abs(b), 1e-3 * rowmax[rows])
    scale[scale == 0.0] = 1.0
    errK = float((np.abs(a - b) / scale).max())
    errKrow = float((np.abs(a - b) / np.where(rowmax[rows] == 0.0, 1.0, rowmax[rows])).max())
    errR = float(np.abs(R[:n_rows] - R_cpu[:n_rows]).max() / np.abs(R_cpu[:n_rows]).max())
    # thresholds as in tests/test_gpu_baseline_sizes.py (the 8d norm of entries that are small by cancellation differs by
    # 2..6e-12 between any two double-precision evaluations; relative to the row maximum they agree to 1e-13)
    return {"K_max_err_8d_norm": errK, "K_max_err_rel_row_max": errKrow, "R_max_err": errR,
            "tol": {"K_8d_norm": 1e-11, "K_rel_row_max": 1e-13, "R": 1e-12},
            "ok": bool(errK <= 1e-11 and errKrow <= 1e-13 and errR <= 1e-12),
            "rows_compared": int(n_rows), "entries_compared": nnz_pref,
            "against": "oracle/cpu_ref.c on the benched state, DBCOption::Raw, SURVEY 8d norm"}


def newton_step(asm, barrier, distributed):
    """assemble K|R + Jacobi-PCG to 1e-8 + solution update, everything resident; wall clock around the device work."""
    from ikarus_b200 import _capi as capi

    lib, h = asm._lib, asm._h
    it, rel = C.c_int(), C.c_double()
    lib.ikb_invalidate(h)
    asm._check(lib.ikb_assemble(h, capi.MATRIX | capi.VECTOR, capi.DBC_FULL))
    asm._check(lib.ikb_pcg_solve(h, capi.DBC_FULL, None, None, 1e-8, 4, C.byref(it), C.byref(rel)))  # untimed warm-up
    barrier()
    t0 = time.perf_counter()
    lib.ikb_invalidate(h)
    asm._check(lib.ikb_assemble(h, capi.MATRIX | capi.VECTOR, capi.DBC_FULL))
    t1 = time.perf_counter()
    asm._check(lib.ikb_pcg_solve(h, capi.DBC_FULL, None, None, 1e-8, 50000, C.byref(it), C.byref(rel)))
    asm._check(lib.ikb_update_solution(h, capi.DBC_FULL, None))
    barrier()
    t2 = time.perf_counter()
    return {"ms": (t2 - t0) * 1e3, "pcg_iterations": it.value, "pcg_rel_tol": 1e-8, "pcg_rel_res": rel.value,
            "ms_per_pcg_iteration": (t2 - t1) * 1e3 / max(it.value, 1),
            "note": ("row-block Jacobi-PCG: halo of the search direction and partial dot products stored into the peers' "
                     "memory over NVLink from inside the three kernels of an iteration (NCCL send/recv + all-reduce if "
                     "CUDA IPC is unavailable), CUDA-graph batches"
                     if distributed else "Jacobi-PCG, SpMV and dot products on the device, no host sync per iteration")}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import ikarus_b200 as ik
    from ikarus_b200 import _capi as capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ikarus_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = args.workload if args.workload != "auto" else ("C2" if world == 1 else "C5")
    W = workload(wl, world)
    h_mesh = min(b / c for b, c in zip(W["bbox"], W["cells"]))
    t_setup = time.perf_counter()
    asm, slab, n_local, need_lo, need_hi = build_handle(wl, rank, world, local)
    lib, h = asm._lib, asm._h
    d_host = torch.empty(slab.n_dof, dtype=torch.float64).pin_memory()
    d_host.numpy()[:] = synthetic_state(slab.n_dof, h_mesh, W["seed"])
    n_rows = (slab.node_end - slab.node_begin) * 3
    r_host = torch.empty(n_rows, dtype=torch.float64).pin_memory()
    asm._check(lib.ikb_set_solution(h, C.c_void_p(d_host.data_ptr())))
    asm._check(lib.ikb_set_parameter(h, 0.0))
    d_ptr = d_host.data_ptr() + 8 * need_lo
    stream = torch.cuda.ExternalStream(asm.stream(), device=torch.device("cuda", local))
    WHAT, DBC = capi.MATRIX | capi.VECTOR, capi.DBC_FULL

    def barrier():
        asm._check(lib.ikb_sync(h))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        lib.ikb_invalidate(h)
        asm._check(lib.ikb_assemble(h, WHAT, DBC))
    barrier()
    t_setup = time.perf_counter() - t_setup
    launches0 = asm.launchCount()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timed region: exactly K steps
    ms_total = time_sweeps(asm, stream, args.steps, barrier, torch)
    launches = asm.launchCount() - launches0
    # ---- end-to-end: host d -> device, assemble, R -> host, every step (pinned buffers)
    for _ in range(2):
        asm._check(lib.ikb_set_solution_range(h, C.c_void_p(d_ptr), need_lo, need_hi - need_lo))
        asm._check(lib.ikb_assemble(h, WHAT, DBC))
        asm._check(lib.ikb_get_vector(h, DBC, C.c_void_p(r_host.data_ptr())))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        asm._check(lib.ikb_set_solution_range(h, C.c_void_p(d_ptr), need_lo, need_hi - need_lo))
        asm._check(lib.ikb_assemble(h, WHAT, DBC))
        asm._check(lib.ikb_get_vector(h, DBC, C.c_void_p(r_host.data_ptr())))  # waits for R only; K gather overlaps
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    # ---- mirror mode (N = 1): the same step with the matrix values fetched to the host as well, which is what an unchanged
    # host LinearSolver / TrustRegion consumes (ikb_get_matrix_values: the CSR value array in Eigen's order)
    e2e_mirror = None
    if world == 1 and not args.no_mirror:
        rows_c, nnz_c = C.c_int64(), C.c_int64()
        asm._check(lib.ikb_pattern_nnz(h, DBC, C.byref(rows_c), C.byref(nnz_c)))
        k_host = torch.empty(nnz_c.value, dtype=torch.float64).pin_memory()
        msteps = max(3, min(args.steps, 10))
        for i in range(2 + msteps):
            if i == 2:
                lib.ikb_sync(h)
                t0 = time.perf_counter()
            asm._check(lib.ikb_set_solution_range(h, C.c_void_p(d_ptr), need_lo, need_hi - need_lo))
            asm._check(lib.ikb_assemble(h, WHAT, DBC))
            asm._check(lib.ikb_get_vector(h, DBC, C.c_void_p(r_host.data_ptr())))
            asm._check(lib.ikb_get_matrix_values(h, DBC, C.c_void_p(k_host.data_ptr())))
        mirror_ms = (time.perf_counter() - t0) * 1e3 / msteps
        e2e_mirror = {"value": int(np.prod(W["cells"])) / mirror_ms / 1e3, "unit": UNIT, "ms_per_step": mirror_ms, "steps": msteps,
                      "d2h_bytes_per_step": 8 * (n_rows + nnz_c.value),
                      "note": "as e2e, plus the CSR values of K to a pinned host buffer every step (mirror mode: matrix() "
                              "returns a host Eigen::SparseMatrix); bound by the 8*nnz bytes over PCIe"}
        del k_host
    # ---- per-kernel durations (CUDA events on the handle's stream, inside the library)
    t_el = asm.timePhase("elements", DBC, 20)
    t_ga = asm.timePhase("gather", DBC, 20)
    if world > 1:
        t = torch.tensor([ms_total, e2e_ms, t_el, t_ga], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, t_el, t_ga = [float(x) for x in t.tolist()]
    # ---- Newton step
    newton = None
    if not args.no_newton:
        if world > 1:
            from ikarus_b200 import distributed as ikd

            ikd.init_communicator(asm, dist)
        newton = newton_step(asm, barrier, world > 1)
        if world > 1:
            t = torch.tensor([newton["ms"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            newton["ms"] = float(t.item())
        # the solution moved: back to the benched state for anything that follows
        asm._check(lib.ikb_set_solution(h, C.c_void_p(d_host.data_ptr())))
    n_elem_global = int(np.prod(W["cells"]))
    ms_step = ms_total / args.steps
    value = n_elem_global / ms_step / 1e3
    e2e_value = n_elem_global / (e2e_ms / args.steps) / 1e3

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (
            6650.0, "fallback (B200_PROFILING.md)")
        rows, nnz = C.c_int64(), C.c_int64()
        lib.ikb_pattern_nnz(h, DBC, C.byref(rows), C.byref(nnz))
        t_peak = asm.timePhase("dfma_peak", DBC, 3)
        fp64_peak = 148 * 16 * 256 * 2048 * 16 / (t_peak * 1e-3) / 1e12
        t_peak = asm.timePhase("dmma_peak", DBC, 3)
        dmma_peak = 148 * 16 * 8 * 2048 * 8 * 512 / (t_peak * 1e-3) / 1e12
        # dominant kernel: the row gather writes every CSR value and R entry once (8*nnz + 8*N)
        ga_bytes = 8.0 * nnz.value + 8.0 * rows.value
        el_bytes = 8.0 * (need_hi - need_lo) + 8.0 * 24 * n_local + 4.0 * 8 * n_local  # u, corner coordinates, connectivity
        if t_ga >= t_el:
            dom, t_dom, dom_bytes = "gather_pull_kernel", t_ga, ga_bytes
        else:
            dom, t_dom, dom_bytes = "elem_h8_mma_kernel", t_el, el_bytes
        traffic = None
        if world == 1 and wl == "C2":
            try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture of this workload
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(dom)
            except Exception:
                pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / (t_dom * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": dom_bytes / (t_dom * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": t_dom,
                    # the other roofline of the path (SURVEY.md 8d classifies C2..C5 as FP64-bound): canonical flops of
                    # the step against the measured FP64 peaks of this GPU (per rank)
                    "fp64": {"canonical_flop_per_elem": FLOP_PER_ELEM,
                             "achieved_tflops_step_per_gpu": FLOP_PER_ELEM * n_local / (ms_step * 1e-3) / 1e12,
                             "achieved_tflops_elem_kernel": FLOP_PER_ELEM * n_local / (t_el * 1e-3) / 1e12,
                             "peak_tflops_measured_dfma": fp64_peak, "peak_tflops_measured_dmma": dmma_peak,
                             "frac_step_of_dfma_peak": FLOP_PER_ELEM * n_local / (ms_step * 1e-3) / 1e12 / fp64_peak,
                             "note": "the kernels use a factored tangent (about 1/3 of the canonical flops) with the "
                                     "contraction on the FP64 tensor cores (DMMA); see DESIGN.md 4"}}
        whole_bytes = ga_bytes + el_bytes
        extra = {"kernels_ms": {"elem_h8_mma_kernel": t_el, "gather_pull_kernel(+gather_vec_kernel)": t_ga},
                 "step_hbm": {"algorithmic_bytes_per_gpu": whole_bytes,
                              "achieved_gbs_per_gpu": whole_bytes / (ms_step * 1e-3) / 1e9,
                              "frac_of_measured_peak": whole_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak},
                 "setup_s": t_setup}
        if newton is not None:
            extra["newton_step"] = newton
        if world == 1 and not args.no_newton:
            # TrustRegion inner solve (Steihaug-Toint tCG, diagonal preconditioner) on the same K, g
            lib.ikb_invalidate(h)
            asm._check(lib.ikb_assemble(h, WHAT, DBC))
            ti = capi.TcgInfo(delta=1e5, kappa=1e-6, theta=1.0, mininner=1, max_iters=4, tol=0.0,
                              precond=capi.PRECOND_DIAGONAL)
            asm._check(lib.ikb_tcg_solve(h, DBC, None, None, C.byref(ti)))  # untimed warm-up
            ti.max_iters = 0
            t0 = time.perf_counter()
            asm._check(lib.ikb_tcg_solve(h, DBC, None, None, C.byref(ti)))
            tcg_ms = (time.perf_counter() - t0) * 1e3
            extra["trust_region_inner"] = {"ms": tcg_ms, "iterations": int(ti.iterations),
                                           "stop_reason": int(ti.stop_reason),
                                           "us_per_iteration": 1e3 * tcg_ms / max(int(ti.iterations), 1),
                                           "rel_error": ti.rel_error}
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(wl, args.cpu_sample, host_threads(), keep=True)
            extra["parity"] = parity_check(asm, wl, cpu.pop("_check"))
            cpu.pop("seconds"), cpu.pop("elements")
        if world == 1 and wl == "C2" and not args.no_c5:
            del asm
            extra["c5_single_gpu"] = c5_on_one_gpu(local, torch)
        if world == 1 and wl == "C2" and not args.no_c4:
            extra["c4_single_gpu"] = c4_on_one_gpu()
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if (world > 1 and wl != "C2W") else "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": W["name"], "elements": n_elem_global, "dofs_global": slab.n_dof,
                          "material": "NeoHooke E=1000 nu=0.3", "dbc": "Full",
                          "state": f"d=0.05*h*U(-1,1) seed {W['seed']}",
                          "parallelism": f"{world} z-slabs of the fixed mesh, owner-computes, row-block CSR" if world > 1
                          else "single GPU",
                          "l2": "per-step working set per GPU (staged K_e + CSR values, >= 600 MB) exceeds the 126 MB "
                                "L2, no explicit flush"},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (need_hi - need_lo) * world,
                       "d2h_bytes_per_step": 8 * n_rows * world, "ms_per_step": e2e_ms / args.steps,
                       "note": "per rank: ikb_set_solution_range(host d, owned+ghost dofs) + ikb_assemble(K|R, Full) + "
                               "ikb_get_vector(host R, owned rows); byte counts are rank 0's times the rank count; "
                               "K stays resident for the device PCG"},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        if e2e_mirror is not None:
            out["e2e_mirror"] = e2e_mirror
        if cpu is not None:
            out["cpu_baseline"] = cpu
            fast = cpu_baseline(wl, args.cpu_sample, host_threads(), opt=True)
            out["cpu_baseline_opt"] = {k: fast[k] for k in ("value", "unit", "cores", "kind", "sample")}
        out.update(extra)
        print(json.dumps(out), file=JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def c5_on_one_gpu(local, torch):
    """configs[4] on ONE GPU: the same-workload single-GPU figure for the strong-scaling lines (K+R sweep and the
    cost of one PCG iteration; a full Jacobi-PCG solve of 51M dofs on one GPU is not run here)."""
    from ikarus_b200 import _capi as capi

    t0 = time.perf_counter()
    asm, slab, n_local, lo, hi = build_handle("C5", 0, 1, local)
    lib, h = asm._lib, asm._h
    W = WORKLOADS["C5"]
    d = synthetic_state(slab.n_dof, 1.0 / 256, W["seed"])
    asm._check(lib.ikb_set_solution(h, capi.ptr(d)))
    asm._check(lib.ikb_set_parameter(h, 0.0))
    stream = torch.cuda.ExternalStream(asm.stream(), device=torch.device("cuda", local))

    def barrier():
        asm._check(lib.ikb_sync(h))
        torch.cuda.synchronize()

    for _ in range(2):
        lib.ikb_invalidate(h)
        asm._check(lib.ikb_assemble(h, capi.MATRIX | capi.VECTOR, capi.DBC_FULL))
    barrier()
    setup = time.perf_counter() - t0
    steps = 5
    ms = time_sweeps(asm, stream, steps, barrier, torch) / steps
    it, rel = C.c_int(), C.c_double()
    asm._check(lib.ikb_pcg_solve(h, capi.DBC_FULL, None, None, 1e-8, 8, C.byref(it), C.byref(rel)))
    barrier()
    t1 = time.perf_counter()
    asm._check(lib.ikb_pcg_solve(h, capi.DBC_FULL, None, None, 1e-30, 96, C.byref(it), C.byref(rel)))
    barrier()
    pcg_ms = (time.perf_counter() - t1) * 1e3 / max(it.value, 1)
    rows, nnz = C.c_int64(), C.c_int64()
    lib.ikb_pattern_nnz(h, capi.DBC_FULL, C.byref(rows), C.byref(nnz))
    out = {"workload": W["name"], "elements": n_local, "dofs": slab.n_dof, "nnz": nnz.value, "value": n_local / ms / 1e3,
           "unit": UNIT, "ms_per_step": ms, "steps": steps, "pcg_ms_per_iteration": pcg_ms, "pcg_iterations_timed": it.value,
           "setup_s": setup}
    del asm
    return out


def c4_on_one_gpu():
    """configs[3] (C4: Hex8 + EAS(21) NeoHooke, nu = 0.499, 96^3) and its E9 variant on ONE GPU: the fused K+R sweep
    as element kernel + gathers, each phase timed with CUDA events on the handle's stream (best of 3 x 10 launches,
    the protocol of tools/config_times.py).  SURVEY 8d targets: 107 Melem/s (E21), 196 (E9)."""
    import ikarus_b200 as ik
    from ikarus_b200 import _capi as capi, meshes

    out = {}
    n = 96
    mat = ik.Materials.NeoHooke(ik.toLamesFirstParameterAndShearModulus(emodul=1000.0, nu=0.499))
    for key, m, target in (("E21", 21, 107.0), ("E9", 9, 196.0)):
        slab = meshes.structured_q1((n, n, n), (1.0, 1.0, 1.0))
        flags = meshes.clamp_face_flags((n, n, n), 0, 0)
        fes = ik.makeFE(dict(dim=3, order=1, n_dof=slab.n_dof), ik.skills(ik.nonLinearElastic(mat), ik.eas(m)),
                        slab.corner_coords, slab.elem_dofs)
        dv = ik.DirichletValues(slab.n_dof)
        dv.container()[:] = flags
        asm = ik.SparseFlatAssembler(fes, dv, mode="resident")
        d = 0.05 * slab.h * np.random.default_rng(44).uniform(-1, 1, slab.n_dof)
        d[flags] = 0
        req = ik.FERequirements(d, 0.0)
        asm.bind(req, ik.elastoStatics, ik.DBCOption.Full)
        asm._assemble(req, capi.MATRIX | capi.VECTOR, ik.DBCOption.Full)
        asm._check(asm._lib.ikb_sync(asm._h))
        te = min(asm.timePhase("elements", ik.DBCOption.Full, 10) for _ in range(3))
        tg = min(asm.timePhase("gather", ik.DBCOption.Full, 10) for _ in range(3))
        out[key] = {"workload": f"C4 Hex8+EAS{m} NeoHooke nu=0.499 {n}^3", "elements": len(fes), "dofs": slab.n_dof,
                    "elements_ms": te, "gather_ms": tg, "value": len(fes) / (te + tg) / 1e3, "unit": UNIT,
                    "survey_8d_target": target}
        del asm
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "C2", "C5", "C2W"])
    ap.add_argument("--no-mirror", action="store_true", help="skip the mirror-mode (K values to the host) e2e figure")
    ap.add_argument("--cpu-sample", type=int, default=32768, help="elements in the cpu_baseline sample of the GPU arm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-newton", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5-on-one-GPU figure of the N=1 run")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (Hex8 + EAS) figures of the N=1 run")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 10)  # each step is 1-2 s of CPU work
        args.warmup = min(args.warmup, 3)
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        wl = args.workload if args.workload != "auto" else ("C2" if world == 1 else "C5")
        if wl == "C5":
            args.steps = min(args.steps, 20)  # a C5 sweep takes 6-50 ms per rank
        run_ours(args)


# stdout carries exactly one JSON line: everything else any library prints (a box-wide NCCL_DEBUG=VERSION makes NCCL
# write its banner to fd 1) is sent to stderr by pointing fd 1 there and keeping a private copy for the result.
JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

if __name__ == "__main__":
    main()
