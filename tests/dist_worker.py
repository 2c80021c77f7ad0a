"""Worker for the multi-rank tests (launched with torch.distributed.run).

  mode "gloo": CPU, world_size 2 -- host logic only: slab partition + halo intervals reproduce what each rank needs.
  mode "nccl": one GPU per rank -- partitioned assembly is bit-identical to the single-GPU assembly on the owned rows,
               distributed PCG / Newton match the single-GPU run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch
import torch.distributed as dist

from ikarus_b200 import distributed as ikd
from ikarus_b200 import meshes

CELLS = (6, 4, 9)
BBOX = (3.0, 2.0, 4.5)


def slab_for(rank, world):
    layers = CELLS[2] + 1
    lb, le = ikd.slab_layers(layers, rank, world)
    return meshes.structured_q1(CELLS, BBOX, lb, le)


def need_range(slab):
    nodes = slab.elem_dofs[:, ::3] // 3
    return int(nodes.min()), int(nodes.max()) + 1


def run_gloo():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    slab = slab_for(rank, world)
    full = meshes.structured_q1(CELLS, BBOX)
    assert slab.n_dof == full.n_dof
    # every element is owned by exactly one rank; ghost layers overlap
    owned = torch.tensor([slab.n_owned_elems])
    dist.all_reduce(owned)
    assert int(owned) == full.corner_coords.shape[0]
    mine = (slab.node_begin, slab.node_end) + need_range(slab)
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    # owned ranges tile [0, nNodes) and the needed columns are covered by neighbours
    assert allr[0][0] == 0 and allr[-1][1] == slab.n_nodes
    for a, b in zip(allr[:-1], allr[1:]):
        assert a[1] == b[0]
    # emulate the halo exchange of a global-length vector with the library's interval logic
    truth = np.arange(slab.n_dof, dtype=np.float64) * 0.5 + 1.0
    v = np.full(slab.n_dof, np.nan)
    v[3 * mine[0]:3 * mine[1]] = truth[3 * mine[0]:3 * mine[1]]
    reqs = []
    bufs = []
    for s in range(world):
        if s == rank:
            continue
        sb, se, rb, re = ikd.halo_intervals(mine, allr[s])
        psb, pse, prb, pre = ikd.halo_intervals(allr[s], mine)
        assert (sb, se) == (prb, pre) and (rb, re) == (psb, pse)  # both sides agree
        if se > sb:
            reqs.append(dist.isend(torch.from_numpy(v[3 * sb:3 * se].copy()), s))
        if re > rb:
            buf = torch.empty(3 * (re - rb), dtype=torch.float64)
            reqs.append(dist.irecv(buf, s))
            bufs.append((rb, re, buf))
    for r in reqs:
        r.wait()
    for rb, re, buf in bufs:
        v[3 * rb:3 * re] = buf.numpy()
    lo, hi = mine[2], mine[3]
    assert np.array_equal(v[3 * lo:3 * hi], truth[3 * lo:3 * hi]), "halo exchange did not deliver every needed column"
    dist.barrier()
    if rank == 0:
        print("GLOO_OK")
    dist.destroy_process_group()


def run_nccl():
    import ctypes as C

    import ikarus_b200 as ik
    from ikarus_b200 import _capi as capi

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    E, nu = 1000.0, 0.3
    p = ik.toLamesFirstParameterAndShearModulus(emodul=E, nu=nu)
    eas_m = int(os.environ.get("IKB_TEST_EAS", "0"))
    sk = [ik.nonLinearElastic(ik.Materials.NeoHooke(p))] + ([ik.eas(eas_m)] if eas_m else [])
    flags = meshes.clamp_face_flags(CELLS, 0, 0)
    full = meshes.structured_q1(CELLS, BBOX)
    slab = slab_for(rank, world)
    n = full.n_dof
    fext = np.zeros(n)
    fext[2::3] = -0.02  # lambda-proportional nodal load in -z

    def make(mesh, rows):
        fes = ik.makeFE(dict(dim=3, order=1, n_dof=n), ik.skills(*sk), mesh.corner_coords, mesh.elem_dofs)
        dv = ik.DirichletValues(n)
        dv.container()[:] = flags
        a = ik.SparseFlatAssembler(fes, dv, device=local, mode="resident", rows=rows)
        a.setExternalLoad(fext)
        return a

    single = make(full, None)
    part = make(slab, (slab.node_begin, slab.node_end))
    ikd.init_communicator(part, dist)
    rng = np.random.default_rng(3)
    d = 0.01 * rng.uniform(-1, 1, n)
    d[flags] = 0.0
    lo, hi = 3 * slab.node_begin, 3 * slab.node_end
    for a in (single, part):
        a._check(a._lib.ikb_set_solution(a._h, capi.ptr(d)))
        a._check(a._lib.ikb_set_parameter(a._h, 0.4))
        a._check(a._lib.ikb_assemble(a._h, capi.MATRIX | capi.VECTOR | (0 if eas_m else capi.SCALAR), capi.DBC_FULL))
    # owned rows bit-identical to the single-GPU assembly
    outer, inner = single.pattern(ik.DBCOption.Full)
    vals = np.empty(inner.shape[0])
    single._check(single._lib.ikb_get_matrix_values(single._h, capi.DBC_FULL, capi.ptr(vals)))
    pouter, pinner = part.pattern(ik.DBCOption.Full)
    pvals = np.empty(pinner.shape[0])
    part._check(part._lib.ikb_get_matrix_values(part._h, capi.DBC_FULL, capi.ptr(pvals)))
    assert np.array_equal(pouter, outer[lo:hi + 1] - outer[lo]), "row-block pattern differs"
    assert np.array_equal(pinner, inner[outer[lo]:outer[hi]])
    assert np.array_equal(pvals, vals[outer[lo]:outer[hi]]), "row-block values differ from the single-GPU assembly"
    R = np.empty(n)
    single._check(single._lib.ikb_get_vector(single._h, capi.DBC_FULL, capi.ptr(R)))
    Rp = np.empty(hi - lo)
    part._check(part._lib.ikb_get_vector(part._h, capi.DBC_FULL, capi.ptr(Rp)))
    assert np.array_equal(Rp, R[lo:hi])
    nrm_s, nrm_p = C.c_double(), C.c_double()
    single._check(single._lib.ikb_vector_norm(single._h, capi.DBC_FULL, C.byref(nrm_s)))
    part._check(part._lib.ikb_vector_norm(part._h, capi.DBC_FULL, C.byref(nrm_p)))
    assert abs(nrm_s.value - nrm_p.value) <= 1e-13 * nrm_s.value
    if not eas_m:
        es, ep = C.c_double(), C.c_double()
        single._check(single._lib.ikb_get_scalar(single._h, C.byref(es)))
        part._check(part._lib.ikb_get_scalar(part._h, C.byref(ep)))
        assert abs(es.value - ep.value) <= 1e-12 * abs(es.value), (es.value, ep.value)
    # the two transports of the distributed PCG (NVLink peer memory: no NCCL inside the iteration; NCCL send/recv +
    # all-reduce) solve the same system: same iteration count up to the last-digit difference of the summation order
    # of the dot products, same solution
    part_nccl = make(slab, (slab.node_begin, slab.node_end))
    ikd.init_communicator(part_nccl, dist, peer_memory=False)
    sols, its = [], []
    for a in (part, part_nccl):
        a._check(a._lib.ikb_set_solution(a._h, capi.ptr(d)))
        a._check(a._lib.ikb_set_parameter(a._h, 0.4))
        a._check(a._lib.ikb_assemble(a._h, capi.MATRIX | capi.VECTOR, capi.DBC_FULL))
        x = np.zeros(hi - lo)
        it, rel = C.c_int(), C.c_double()
        for _ in range(2):  # twice: the second solve runs on a new epoch of the peer windows
            a._check(a._lib.ikb_pcg_solve(a._h, capi.DBC_FULL, None, capi.ptr(x), 1e-12, 5000, C.byref(it), C.byref(rel)))
        assert rel.value <= 1e-12 and it.value > 5
        sols.append(x.copy())
        its.append(it.value)
    assert abs(its[0] - its[1]) <= 2, its
    assert np.abs(sols[0] - sols[1]).max() <= 1e-9 * np.abs(sols[1]).max()
    del part_nccl
    # Newton on both: identical iteration counts, same solution
    z = np.zeros(n)
    for a in (single, part):
        a._check(a._lib.ikb_set_solution(a._h, capi.ptr(z)))
        if eas_m:
            a.setInternalVariables(np.zeros((len(a._fes), eas_m)))
    ns = ikd.DistributedNewton(single, tol=1e-9, pcg_tol=1e-13)
    npart = ikd.DistributedNewton(part, tol=1e-9, pcg_tol=1e-13)
    its_s = [ns.solve(lam)[0] for lam in (0.5, 1.0)]
    its_p = [npart.solve(lam)[0] for lam in (0.5, 1.0)]
    assert its_s == its_p and min(its_s) >= 2, (its_s, its_p)
    ds, dp = ns.solution(), npart.solution()
    lo_n, hi_n = need_range(slab)
    err = np.abs(ds[3 * lo_n:3 * hi_n] - dp[3 * lo_n:3 * hi_n]).max()
    assert err <= 1e-9 * np.abs(ds).max(), err
    dist.barrier()
    if rank == 0:
        print(f"NCCL_OK world={world} eas={eas_m} newton_its={its_p} pcg_its={npart.pcg_iterations} "
              f"max|d|={np.abs(ds).max():.6f} err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    {"gloo": run_gloo, "nccl": run_nccl}[sys.argv[1]]()
