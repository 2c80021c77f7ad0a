// Row-block Jacobi-PCG over NVLink peer memory: the iteration contains no NCCL call and no host round trip.
//
// Replaces, for an element-partitioned run, the Eigen::ConjugateGradient + DiagonalPreconditioner the reference selects
// with SolverTypeTag::si_ConjugateGradient (ikarus/solver/linearsolver/linearsolver.cpp:23-24).  The reference has no
// distributed code; what is fused here is the communication the partitioning adds:
//
//   halo        The search direction p lives in a vector of GLOBAL length on every rank (owned rows + halo).  The
//               kernel that computes the new p stores the entries a neighbour needs straight into the neighbour's copy
//               (plain st.global on a pointer opened with cudaIpcOpenMemHandle: NVLink writes) and then raises a
//               sequence flag there.  No pack/unpack kernels, no ncclSend/ncclRecv.
//   overlap     The SpMV walks the rows that touch no halo column first; only when a warp reaches a boundary row does it
//               look at the flag (and reads those x entries past L1), so the halo transfer runs behind the interior rows.
//   reductions  p.q and (r.z, r.r): the last block of the producing kernel folds the block partials in a fixed order
//               and stores the rank's sum into a slot of EVERY rank's window, followed by a sequence flag; the
//               consuming kernel waits for all slots and adds them in rank order.  Every rank therefore sees the
//               same bits and takes the same decisions, and the result does not depend on arrival order.
//   graph       Three kernels per iteration, all arguments resident: batches are replayed from one CUDA graph; the
//               host reads the 48-byte state once per batch.
//
// Sequence numbers are (solve epoch << 32) | iteration+1, so the windows are never reset (a reset could erase a flag a
// faster peer has already raised) and stale values of an earlier solve can never satisfy a wait.
#pragma once
#include "ikb_pcg.cuh"

namespace ikb {

constexpr int PEER_MAXR = 8;
constexpr unsigned PEER_SPIN_LIMIT = 1u << 24;
// control window, in 8-byte words
constexpr int PW_HALO = 0;                      // [s]            halo of p for iteration k from rank s is in place
constexpr int PW_PQ = PW_HALO + PEER_MAXR;      // [par][s]       p.q of rank s
constexpr int PW_PQS = PW_PQ + 2 * PEER_MAXR;   // [par][s]       its sequence number
constexpr int PW_RZ = PW_PQS + 2 * PEER_MAXR;   // [par][s][2]    r.z, r.r of rank s
constexpr int PW_RZS = PW_RZ + 4 * PEER_MAXR;   // [par][s]
constexpr int PW_WORDS = PW_RZS + 2 * PEER_MAXR;

struct PeerComm {
  int rank, nranks;
  double* peerP[PEER_MAXR];                 // every rank's search-direction vector (global dof indexing); [rank] = mine
  unsigned long long* peerW[PEER_MAXR];     // every rank's control window
  int nSend;
  int sendPeer[PEER_MAXR];
  long long sendBegin[PEER_MAXR], sendEnd[PEER_MAXR];  // dof intervals of MY rows the peer reads
  int nRecv;
  int recvPeer[PEER_MAXR];                  // ranks whose rows my boundary rows read
};

struct PeerState {  // device resident, identical on every rank
  double rz[2];
  double rr, threshold;
  int iter, done, maxIter, pad;
  unsigned long long epoch;
  unsigned int arrive[4];  // arrival counters of the three kernels
  int timeout;
};

__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stReleaseSys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void stRelaxedSys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ bool peerWaitGe(const unsigned long long* w, unsigned long long target, PeerState* st) {
  unsigned spins = 0;
  while (ldAcquireSys(w) < target) {
    if (++spins > PEER_SPIN_LIMIT) {
      st->timeout = 1;
      return false;
    }
    __nanosleep(100);
  }
  return true;
}

// q = A p on the owned rows, interior rows first; p.q of this rank published to every rank
template <int D>
__global__ void __launch_bounds__(256, 6)
    peer_spmv_kernel(PatternView P, const double* __restrict__ vals, const double* x, double* __restrict__ y,
                     const double* __restrict__ pLocal, double* partial, PeerState* st, PeerComm C, unsigned long long* win,
                     int64_t firstInterior, int64_t endInterior) {
  if (st->done || st->iter >= st->maxIter) return;
  __shared__ double sh[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = st->iter;
  const unsigned long long seqBase = st->epoch << 32;
  const int64_t warpsTotal = (int64_t)gridDim.x * 8;
  const int64_t nInterior = endInterior - firstInterior;
  bool haloReady = (it == 0);  // the initial direction is exchanged before the loop
  double dot = 0.0;
  for (int64_t idx = (int64_t)blockIdx.x * 8 + warp; idx < P.nRowNodes; idx += warpsTotal) {
    int64_t g;
    bool boundary = false;
    if (idx < nInterior) {
      g = firstInterior + idx;
    } else {
      const int64_t j = idx - nInterior;
      g = j < firstInterior ? j : endInterior + (j - firstInterior);
      boundary = true;
      if (!haloReady) {
        if (lane < C.nRecv) peerWaitGe(win + PW_HALO + C.recvPeer[lane], seqBase | (unsigned long long)it, st);
        __syncwarp();
        haloReady = true;
      }
    }
    // halo entries are written by a peer while this kernel runs: boundary rows read x from L2 (a line fetched earlier
    // for an interior row may hold their old values)
    double s[D];
    if (boundary)
      spmvNodeRow<D, true>(P, vals, x, g, lane, s);
    else
      spmvNodeRow<D, false>(P, vals, x, g, lane, s);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const int64_t r = localRowOf(P, g, i);
        y[r] = s[i];
        dot = fma(pLocal[r], s[i], dot);
      }
    }
  }
  if (lane == 0) sh[warp] = dot;
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&st->arrive[0], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    const double total = blockFold(partial, gridDim.x, sh);
    if (threadIdx.x < C.nranks) {
      const int par = it & 1;
      stRelaxedSys(reinterpret_cast<double*>(C.peerW[threadIdx.x] + PW_PQ + par * PEER_MAXR + C.rank), total);
      __threadfence_system();
      stReleaseSys(C.peerW[threadIdx.x] + PW_PQS + par * PEER_MAXR + C.rank, seqBase | (unsigned long long)(it + 1));
    }
    if (threadIdx.x == 0) st->arrive[0] = 0u;
  }
}

// x += alpha p ; r -= alpha q ; z = dinv r ; (r.z, r.r) of this rank published to every rank
__global__ void __launch_bounds__(256, 4)
    peer_update_kernel(int64_t n, PeerState* st, PeerComm C, unsigned long long* win, const double* __restrict__ p,
                       const double* __restrict__ q, const double* __restrict__ dinv, double* x, double* r, double* z,
                       double* partial) {
  if (st->done || st->iter >= st->maxIter) return;
  __shared__ double sh[256];
  __shared__ double sh1[256];
  __shared__ double pqS;
  const int it = st->iter, par = it & 1;
  const unsigned long long seq = (st->epoch << 32) | (unsigned long long)(it + 1);
  // The grid is one resident wave (RED_BLOCKS = 4 x 148): every block first pulls its share of the vectors into
  // registers and only then looks for the peers' p.q, so the NVLink round trip hides behind the loads.
  constexpr int UPD = 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iFirst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double pv[UPD], qv[UPD], xv[UPD], rv[UPD], dv[UPD];
#pragma unroll
  for (int u = 0; u < UPD; ++u) {
    const int64_t i = iFirst + u * stride;
    const bool ok = i < n;
    pv[u] = ok ? p[i] : 0.0;
    qv[u] = ok ? q[i] : 0.0;
    xv[u] = ok ? x[i] : 0.0;
    rv[u] = ok ? r[i] : 0.0;
    dv[u] = ok ? dinv[i] : 0.0;
  }
  if (threadIdx.x < C.nranks) peerWaitGe(win + PW_PQS + par * PEER_MAXR + threadIdx.x, seq, st);
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int s = 0; s < C.nranks; ++s) t += __ldcg(reinterpret_cast<const double*>(win + PW_PQ + par * PEER_MAXR + s));
    pqS = t;
  }
  __syncthreads();
  const double pq = pqS;
  const double rz = st->rz[par];
  const double alpha = pq != 0.0 ? rz / pq : 0.0;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i0 = iFirst; i0 < n; i0 += UPD * stride) {
    if (i0 != iFirst) {
#pragma unroll
      for (int u = 0; u < UPD; ++u) {
        const int64_t i = i0 + u * stride;
        const bool ok = i < n;
        pv[u] = ok ? p[i] : 0.0;
        qv[u] = ok ? q[i] : 0.0;
        xv[u] = ok ? x[i] : 0.0;
        rv[u] = ok ? r[i] : 0.0;
        dv[u] = ok ? dinv[i] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < UPD; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) {
        x[i] = fma(alpha, pv[u], xv[u]);
        const double ri = fma(-alpha, qv[u], rv[u]);
        r[i] = ri;
        const double zi = dv[u] * ri;
        z[i] = zi;
        s0 = fma(ri, zi, s0);
        s1 = fma(ri, ri, s1);
      }
    }
  }
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, w);
    s1 += __shfl_down_sync(0xffffffffu, s1, w);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = s0;
    sh1[threadIdx.x >> 5] = s1;
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) {
      t0 += sh[w];
      t1 += sh1[w];
    }
    partial[blockIdx.x] = t0;
    partial[gridDim.x + blockIdx.x] = t1;
    __threadfence();
    last = atomicAdd(&st->arrive[1], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    const double a = blockFold(partial, gridDim.x, sh);
    const double b = blockFold(partial + gridDim.x, gridDim.x, sh);
    if (threadIdx.x < C.nranks) {
      double* dst = reinterpret_cast<double*>(C.peerW[threadIdx.x] + PW_RZ + (par * PEER_MAXR + C.rank) * 2);
      stRelaxedSys(dst, a);
      stRelaxedSys(dst + 1, b);
      __threadfence_system();
      stReleaseSys(C.peerW[threadIdx.x] + PW_RZS + par * PEER_MAXR + C.rank, seq);
    }
    if (threadIdx.x == 0) st->arrive[1] = 0u;
  }
}

// p = z + beta p on the owned rows (pGlob + off) AND in the neighbours' copies; the last block raises the halo flags
// and publishes the iteration's state
__global__ void __launch_bounds__(256)
    peer_direction_kernel(int64_t n, int64_t off, PeerState* st, PeerComm C, unsigned long long* win,
                          const double* __restrict__ z, double* pGlob) {
  if (st->done || st->iter >= st->maxIter) return;
  __shared__ double sums[2];
  const int it = st->iter, par = it & 1;
  const unsigned long long seq = (st->epoch << 32) | (unsigned long long)(it + 1);
  double* p = pGlob + off;
  constexpr int UPD = 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iFirst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double pv[UPD], zv[UPD];  // loaded before the wait, as in the update kernel
#pragma unroll
  for (int u = 0; u < UPD; ++u) {
    const int64_t i = iFirst + u * stride;
    pv[u] = i < n ? p[i] : 0.0;
    zv[u] = i < n ? z[i] : 0.0;
  }
  if (threadIdx.x < C.nranks) peerWaitGe(win + PW_RZS + par * PEER_MAXR + threadIdx.x, seq, st);
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int s = 0; s < C.nranks; ++s) {
      const double* src = reinterpret_cast<const double*>(win + PW_RZ + (par * PEER_MAXR + s) * 2);
      a += __ldcg(src);
      b += __ldcg(src + 1);
    }
    sums[0] = a;
    sums[1] = b;
  }
  __syncthreads();
  const double rzNew = sums[0], rr = sums[1];
  const double rz = st->rz[par];
  const double beta = rz != 0.0 ? rzNew / rz : 0.0;
  for (int64_t i0 = iFirst; i0 < n; i0 += UPD * stride) {
    if (i0 != iFirst) {
#pragma unroll
      for (int u = 0; u < UPD; ++u) {
        const int64_t i = i0 + u * stride;
        pv[u] = i < n ? p[i] : 0.0;
        zv[u] = i < n ? z[i] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < UPD; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) {
        const double v = fma(beta, pv[u], zv[u]);
        p[i] = v;
        // entries a neighbour reads go straight into its copy of p as well (the same thread, so the old value is
        // never read after it has been overwritten)
        const int64_t gi = off + i;
        for (int s = 0; s < C.nSend; ++s)
          if (gi >= C.sendBegin[s] && gi < C.sendEnd[s]) stRelaxedSys(C.peerP[C.sendPeer[s]] + gi, v);
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(&st->arrive[2], 1u);
    if (ticket == gridDim.x - 1) {
      st->arrive[2] = 0u;
      __threadfence_system();
      for (int s = 0; s < C.nSend; ++s) stReleaseSys(C.peerW[C.sendPeer[s]] + PW_HALO + C.rank, seq);
      st->rz[(it + 1) & 1] = rzNew;
      st->rr = rr;
      if (!(rr == rr))
        st->done = 2;
      else if (rr < st->threshold)
        st->done = 1;
      __threadfence();
      st->iter = it + 1;
    }
  }
}

__global__ void peer_init_kernel(PeerState* st, const double* rzDev, const double* bbDev, double relTol, int maxIter,
                                 unsigned long long epoch) {
  st->rz[0] = rzDev[0];
  st->rz[1] = 0.0;
  st->rr = bbDev[0];
  const double thr = relTol * relTol * bbDev[0];
  st->threshold = thr > 1e-300 ? thr : 1e-300;
  st->iter = 0;
  st->done = (bbDev[0] > 0.0 && bbDev[0] >= st->threshold) ? 0 : 1;
  st->maxIter = maxIter;
  st->epoch = epoch;
  st->arrive[0] = st->arrive[1] = st->arrive[2] = 0u;
  st->timeout = 0;
}

// rows [0, firstInterior) and [endInterior, nRowNodes) may read halo columns (exact for slab partitions,
// conservative otherwise)
__global__ void peer_boundary_rows_kernel(PatternView P, int64_t rowEnd, int32_t* firstInterior, int32_t* endInterior) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.nRowNodes) return;
  const int32_t b0 = P.nbrPtr[g], b1 = P.nbrPtr[g + 1];
  if (b1 == b0) return;
  if (P.nbrIdx[b0] < P.rowBegin) atomicMax(firstInterior, (int32_t)g + 1);
  if (P.nbrIdx[b1 - 1] >= rowEnd) atomicMin(endInterior, (int32_t)g);
}

}  // namespace ikb
