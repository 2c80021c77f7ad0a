#!/usr/bin/env python
"""bench.py -- FP64 K+R assembly throughput (Melem/s) and Newton-step time on B200.

A "step" is one fused K+R assembly sweep (element kernel + deterministic gather with DBCOption::Full)
over the workload mesh, through the C-ABI of libikb200.so.  Workload at N=1: BASELINE.json configs[1]
(3D cantilever Hex8 Q1 NeoHooke, YaspGrid 128x32x32, ~420k DOF).  At N>1 every rank owns a z-slab of
128x32x(32*N) (weak scaling, owner-computes with one ghost element layer, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

--impl reference times the CPU port of the reference's assembly loops (oracle/cpu_ref.c, all host
threads) on a bounded sample of the same workload; the reference itself cannot be built here.
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

CELLS = (128, 32, 32)
H = 1.0 / 32.0
EMOD, NU = 1000.0, 0.3
SEED = 42
FLOP_PER_ELEM = 59520  # canonical Hex8 NeoHooke K+R flops per element
METRIC = "FP64 K+R assembly throughput (Hex8 NeoHooke, DBCOption::Full)"
UNIT = "Melem/s"


def workload_name(n):
    return f"C2 cantilever Hex8 Q1 NeoHooke YaspGrid {CELLS[0]}x{CELLS[1]}x{CELLS[2] * n}"


def lame():
    return EMOD * NU / ((1 + NU) * (1 - 2 * NU)), EMOD / (2 * (1 + NU))


def synthetic_state(n_dof, h):
    """d = 0.05*h*U(-1,1), seed 42."""
    return 0.05 * h * np.random.default_rng(SEED).uniform(-1.0, 1.0, n_dof)


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
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
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
        time.sleep(0.15)
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


# ------------------------------------------------------------------------------------ CPU baseline
_CPU_CACHE = {}


def cpu_baseline(sample_elems, threads=None):
    """Times the C port of the reference's loops (oracle/cpu_ref.c): separate R and K sweeps, tangent per node
    pair, scatter through linear indices.  Bounded sample of the C2 workload (a 128x32xL sub-box)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref  # the one place bench.py executes oracle/ code: as the measured CPU baseline
    import ikarus_oracle as o

    layers = max(1, int(round(sample_elems / (CELLS[0] * CELLS[1]))))
    cells = (CELLS[0], CELLS[1], layers)
    if cells not in _CPU_CACHE:
        mesh = o.structured_mesh(cells, tuple(c * H for c in cells))
        ed = mesh.elem_dofs()
        n = mesh.n_nodes * 3
        outer, inner = o.build_pattern(ed, n)
        lin = o.linear_indices(ed, outer, inner).reshape(-1, 24, 24).transpose(0, 2, 1).reshape(-1, 576)
        _CPU_CACHE[cells] = (mesh, ed, np.ascontiguousarray(lin), synthetic_state(n, H), inner.shape[0])
    mesh, ed, lin, d, nnz = _CPU_CACHE[cells]
    lam, mu = lame()
    threads = threads or cpu_ref.max_threads()
    cpu_ref.assemble(3, "neohooke", lam, mu, mesh.corner_coords[:256], ed[:256], lin[:256], d, nnz,
                     nthreads=threads)  # warm-up
    t0 = time.perf_counter()
    cpu_ref.assemble(3, "neohooke", lam, mu, mesh.corner_coords, ed, lin, d, nnz, nthreads=threads)
    dt = time.perf_counter() - t0
    return {"value": mesh.n_elem / dt / 1e6, "unit": UNIT, "cores": int(threads), "kind": "port",
            "sample": f"{mesh.n_elem} elements ({cells[0]}x{cells[1]}x{cells[2]} sub-box of the workload), K and R "
                      f"sweeps, {dt:.2f} s wall", "seconds": dt, "elements": int(mesh.n_elem)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 16384  # elements per step: ~0.5-1 s of CPU work on 16 cores
    vals = []
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(per_step)
        if i >= args.warmup:
            vals.append(base)
    dt = float(np.mean([b["seconds"] for b in vals]))
    v = vals[0]["elements"] / dt / 1e6
    base = dict(vals[-1], value=v)
    base.pop("seconds"), base.pop("elements")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": workload_name(1), "sample_elements_per_step": vals[0]["elements"]},
                      "cpu_baseline": base,
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import ikarus_b200 as ik
    from ikarus_b200 import _capi as capi
    from ikarus_b200 import meshes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ikarus_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = world
    cells = (CELLS[0], CELLS[1], CELLS[2] * n)
    bbox = tuple(c * H for c in cells)
    # z-slab of node layers owned by this rank
    layers = cells[2] + 1
    lb = rank * layers // n
    le = (rank + 1) * layers // n
    slab = meshes.structured_q1(cells, bbox, lb, le) if n > 1 else meshes.structured_q1(cells, bbox)
    lam, mu = lame()
    mat = ik.Materials.NeoHooke(ik.fe.LamesFirstParameterAndShearModulus(lam, mu))
    fes = ik.makeFE(dict(dim=3, order=1, n_dof=slab.n_dof), ik.skills(ik.nonLinearElastic(mat)), slab.corner_coords,
                    slab.elem_dofs)
    dv = ik.DirichletValues(slab.n_dof)
    dv.container()[:] = meshes.clamp_face_flags(cells, 0, 0)
    asm = ik.SparseFlatAssembler(fes, dv, device=local, mode="resident",
                                 rows=(slab.node_begin, slab.node_end) if n > 1 else None)
    lib, h = asm._lib, asm._h
    d_host = torch.empty(slab.n_dof, dtype=torch.float64).pin_memory()
    d_host.numpy()[:] = synthetic_state(slab.n_dof, H)
    n_rows = (slab.node_end - slab.node_begin) * 3
    r_host = torch.empty(n_rows, dtype=torch.float64).pin_memory()
    asm._check(lib.ikb_set_solution(h, C.c_void_p(d_host.data_ptr())))
    asm._check(lib.ikb_set_parameter(h, 0.0))
    # dofs this rank needs every step: owned node layers plus one ghost layer on each side
    need_lo = int(slab.elem_dofs.min())
    need_hi = int(slab.elem_dofs.max()) + 1
    d_ptr = d_host.data_ptr() + 8 * need_lo
    stream = torch.cuda.ExternalStream(asm.stream(), device=torch.device("cuda", local))
    WHAT, DBC = capi.MATRIX | capi.VECTOR, capi.DBC_FULL

    def step():
        lib.ikb_invalidate(h)
        asm._check(lib.ikb_assemble(h, WHAT, DBC))

    def barrier():
        asm._check(lib.ikb_sync(h))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = asm.launchCount()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timed region: exactly K steps, CUDA events on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
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
    # ---- per-kernel durations (CUDA events on the handle's stream, inside the library)
    t_el = asm.timePhase("elements", DBC, 20)
    t_ga = asm.timePhase("gather", DBC, 20)
    if world > 1:
        t = torch.tensor([ms_total, e2e_ms, t_el, t_ga], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, t_el, t_ga = [float(x) for x in t.tolist()]
    # multi-GPU Newton step (all ranks take part: halo exchange + all-reduced CG scalars inside the library)
    dist_newton = None
    if world > 1 and not args.no_newton:
        from ikarus_b200 import distributed as ikd

        ikd.init_communicator(asm, dist)
        it, rel = C.c_int(), C.c_double()
        step()
        # untimed warm-up: the first collective calls set up the NCCL connections
        asm._check(lib.ikb_pcg_solve(h, DBC, None, None, 1e-8, 4, C.byref(it), C.byref(rel)))
        barrier()
        t0 = time.perf_counter()
        step()
        asm._check(lib.ikb_pcg_solve(h, DBC, None, None, 1e-8, 20000, C.byref(it), C.byref(rel)))
        asm._check(lib.ikb_update_solution(h, DBC, None))
        barrier()
        dist_newton = {"ms": (time.perf_counter() - t0) * 1e3, "pcg_iterations": it.value, "pcg_rel_tol": 1e-8,
                       "pcg_rel_res": rel.value, "note": "row-block PCG: NCCL halo exchange per SpMV, all-reduced dots"}
    n_elem_global = cells[0] * cells[1] * cells[2]
    ms_step = ms_total / args.steps
    value = n_elem_global / ms_step / 1e3
    e2e_value = n_elem_global / (e2e_ms / args.steps) / 1e3

    extra = {}
    cpu = None
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
        # algorithmic bytes of the dominant kernel (gather): CSR values and R written once
        ga_bytes = 8.0 * nnz.value + 8.0 * rows.value
        dom, t_dom, dom_bytes = ("gather_pull_kernel", t_ga, ga_bytes)
        if t_el > t_ga:
            # element kernel: u, corner coordinates, connectivity read once
            dom, t_dom = "elem_q1_kernel", t_el
            dom_bytes = 8.0 * slab.n_dof + 8.0 * 24 * len(fes) + 4.0 * 8 * len(fes)
        traffic = None
        try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(dom)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / (t_dom * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": dom_bytes / (t_dom * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": t_dom}
        t_peak = asm.timePhase("dfma_peak", DBC, 3)
        fp64_peak = 148 * 16 * 256 * 2048 * 16 / (t_peak * 1e-3) / 1e12
        whole_bytes = 8.0 * nnz.value + 8.0 * rows.value + 8.0 * slab.n_dof + 8.0 * 24 * len(fes) + 4.0 * 8 * len(fes)
        extra = {
            "kernels_ms": {"elem_q1_kernel": t_el, "gather_pull_kernel": t_ga},
            "fp64": {"canonical_flop_per_elem": FLOP_PER_ELEM,
                     "achieved_tflops_step": FLOP_PER_ELEM * len(fes) / (ms_step * 1e-3) / 1e12,
                     "achieved_tflops_elem_kernel": FLOP_PER_ELEM * len(fes) / (t_el * 1e-3) / 1e12,
                     "peak_tflops_measured_dfma": fp64_peak},
            "step_hbm": {"algorithmic_bytes": whole_bytes, "achieved_gbs": whole_bytes / (ms_step * 1e-3) / 1e9,
                         "frac_of_measured_peak": whole_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak},
        }
        # Newton-step time: assemble + Jacobi-PCG + update on the device
        if world == 1 and not args.no_newton:
            ls = ik.DeviceLinearSolver(relTol=1e-8, maxIter=20000)
            it, rel = C.c_int(), C.c_double()
            step()
            # TrustRegion inner solve (Steihaug-Toint tCG, diagonal preconditioner) on the same K, g
            from ikarus_b200 import _capi as capi
            ti = capi.TcgInfo(delta=1e5, kappa=1e-6, theta=1.0, mininner=1, max_iters=4, tol=0.0,
                              precond=capi.PRECOND_DIAGONAL)
            asm._check(lib.ikb_tcg_solve(h, DBC, None, None, C.byref(ti)))  # untimed warm-up
            ti.max_iters = 0
            t0 = time.perf_counter()
            asm._check(lib.ikb_tcg_solve(h, DBC, None, None, C.byref(ti)))
            tcg_ms = (time.perf_counter() - t0) * 1e3
            extra["trust_region_inner"] = {"ms": tcg_ms, "iterations": int(ti.iterations), "stop_reason": int(ti.stop_reason),
                                           "us_per_iteration": 1e3 * tcg_ms / max(int(ti.iterations), 1),
                                           "rel_error": ti.rel_error}
            asm._check(lib.ikb_pcg_solve(h, DBC, None, None, 1e-8, 4, C.byref(it), C.byref(rel)))  # untimed warm-up
            asm._check(lib.ikb_sync(h))
            t0 = time.perf_counter()
            step()
            asm._check(lib.ikb_pcg_solve(h, DBC, None, None, 1e-8, 20000, C.byref(it), C.byref(rel)))
            asm._check(lib.ikb_update_solution(h, DBC, None))
            asm._check(lib.ikb_sync(h))
            extra["newton_step"] = {"ms": (time.perf_counter() - t0) * 1e3, "pcg_iterations": it.value,
                                    "pcg_rel_tol": 1e-8, "pcg_rel_res": rel.value}
        if dist_newton is not None:
            extra["newton_step"] = dist_newton
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(args.cpu_sample)
            cpu.pop("seconds"), cpu.pop("elements")
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(n), "elements": n_elem_global, "dofs_global": slab.n_dof,
                          "material": "NeoHooke E=1000 nu=0.3", "dbc": "Full", "state": "d=0.05*h*U(-1,1) seed 42",
                          "parallelism": f"z-slab x{n}" if n > 1 else "single GPU",
                          "l2": "per-step working set (staged K_e 340 MB + CSR values 261 MB per GPU) exceeds the "
                                "126 MB L2, no explicit flush"},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (need_hi - need_lo) * n,
                       "d2h_bytes_per_step": 8 * n_rows * n, "ms_per_step": e2e_ms / args.steps,
                       "note": "per rank: ikb_set_solution_range(host d, owned+ghost dofs) + ikb_assemble(K|R, Full) + "
                               "ikb_get_vector(host R, owned rows); bytes are summed over ranks; "
                               "K stays resident for the device PCG"},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        if cpu is not None:
            out["cpu_baseline"] = cpu
        out.update(extra)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # keep stdout to the one JSON line: a box-wide NCCL_DEBUG=VERSION makes NCCL print its banner there
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=32768, help="elements in the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-newton", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 20)  # each step is ~1 s of CPU work on a bounded sample
        args.warmup = min(args.warmup, 3)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
