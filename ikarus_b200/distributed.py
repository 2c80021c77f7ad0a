"""Element-partitioned multi-GPU driver: one process per GPU, `torch.distributed` only for the plumbing
(rendezvous, broadcasting the NCCL id); halo exchange and CG all-reduces run inside libikb200.so on its own
NCCL communicator.  The reference has no distributed code; this is new."""
import ctypes as C

import numpy as np

from . import _capi as capi
from .assembler import DBCOption, SparseFlatAssembler


def slab_layers(n_layers, rank, world):
    """Node layers [begin, end) of the slowest axis owned by `rank` (contiguous row block)."""
    return rank * n_layers // world, (rank + 1) * n_layers // world


def halo_intervals(mine, peer):
    """(sendBegin, sendEnd, recvBegin, recvEnd) in nodes; mine/peer = (ownB, ownE, needB, needE)."""
    lib = capi.load()
    a = np.asarray(mine, dtype=np.int64)
    b = np.asarray(peer, dtype=np.int64)
    out = np.zeros(4, dtype=np.int64)
    rc = lib.ikb_halo_intervals(capi.ptr(a), capi.ptr(b), capi.ptr(out))
    assert rc == 0
    return tuple(int(v) for v in out)


def init_communicator(asm: SparseFlatAssembler, dist, peer_memory=True):
    """Create the library's NCCL communicator: rank 0 makes the id, torch.distributed broadcasts it."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    buf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        rc = asm._lib.ikb_nccl_unique_id(capi.ptr(buf))
        if rc != 0:
            raise RuntimeError("ikb_nccl_unique_id failed (libnccl.so.2 not loadable)")
    t = torch.from_numpy(buf)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    buf = t.cpu().numpy().copy()
    asm._check(asm._lib.ikb_comm_init(asm._h, capi.ptr(buf), rank, world))
    if peer_memory and world <= 8:
        _open_peer_memory(asm, dist, rank, world)


def _open_peer_memory(asm, dist, rank, world):
    """All-gather the CUDA IPC handles of every rank's search-direction vector and control window, so that the PCG
    kernels can store halo entries and partial sums straight into the neighbours' memory over NVLink."""
    import torch

    mine = np.zeros(128, dtype=np.uint8)
    ok_local = asm._lib.ikb_comm_ipc_export(asm._h, capi.ptr(mine)) == 0
    dev = dist.get_backend() == "nccl"
    t = torch.from_numpy(mine)
    t = t.cuda() if dev else t
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    flag = torch.tensor([1 if ok_local else 0], dtype=torch.int32)
    flag = flag.cuda() if dev else flag
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return False
    allh = np.concatenate([g.cpu().numpy() for g in gathered]).astype(np.uint8)
    ok = C.c_int(0)
    asm._check(asm._lib.ikb_comm_ipc_import(asm._h, capi.ptr(allh), C.byref(ok)))
    flag = torch.tensor([ok.value], dtype=torch.int32)
    flag = flag.cuda() if dev else flag
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:  # some rank could not open a peer: everybody stays on the NCCL transport
        asm._check(asm._lib.ikb_comm_ipc_import(asm._h, None, C.byref(ok)))
        return False
    return True


class DistributedNewton:
    """Newton iteration with everything resident: assemble (owner-computes), all-reduced residual norm,
    distributed Jacobi-PCG, halo-exchanged correction, solution update.  Mirrors NewtonRaphson::solve
    (solver/nonlinearsolver/newtonraphson.hh:196-257) for DBCOption::Full."""

    def __init__(self, asm: SparseFlatAssembler, tol=1e-8, max_iter=20, pcg_tol=1e-12, pcg_max_it=100000):
        self.asm, self.tol, self.max_iter, self.pcg_tol, self.pcg_max_it = asm, tol, max_iter, pcg_tol, pcg_max_it
        self.pcg_iterations = []

    def _assemble(self):
        a = self.asm
        a._check(a._lib.ikb_assemble(a._h, capi.MATRIX | capi.VECTOR, capi.DBC_FULL))
        nrm = C.c_double()
        a._check(a._lib.ikb_vector_norm(a._h, capi.DBC_FULL, C.byref(nrm)))
        return nrm.value

    def solve(self, lam):
        a = self.asm
        a._check(a._lib.ikb_set_parameter(a._h, float(lam)))
        r = self._assemble()
        it = 0
        while r > self.tol and it < self.max_iter:
            n_it, rel = C.c_int(), C.c_double()
            a._check(a._lib.ikb_pcg_solve(a._h, capi.DBC_FULL, None, None, self.pcg_tol, self.pcg_max_it,
                                          C.byref(n_it), C.byref(rel)))
            self.pcg_iterations.append(n_it.value)
            if a._fes.numberOfInternalVariables():
                a._check(a._lib.ikb_eas_update(a._h, None))  # CORRECTION_UPDATED at the old state
            a._check(a._lib.ikb_update_solution(a._h, capi.DBC_FULL, None))
            r = self._assemble()
            it += 1
        return it, r

    def solution(self):
        a = self.asm
        d = np.empty(a.size())
        a._check(a._lib.ikb_get_solution(a._h, capi.ptr(d)))
        return d
