// Device SpMV and Jacobi-preconditioned conjugate gradients.
//
// Replaces the Eigen::ConjugateGradient + DiagonalPreconditioner pair the reference selects with
// SolverTypeTag::si_ConjugateGradient (ikarus/solver/linearsolver/linearsolver.cpp:23-24) and the
// SpMV inside it.  Dot products are two-stage fixed-shape reductions (deterministic); scalars stay
// on the device, the host reads one residual norm per iteration from pinned memory.
#pragma once
#include "ikb_gather.cuh"
#include "ikb_internal.cuh"

namespace ikb {

// Raw/Full matrix: one group of LANES lanes per scalar row, columns decoded from the node-block view.
template <int D, int LANES>
__global__ void __launch_bounds__(256) spmv_block_kernel(PatternView P, const double* __restrict__ vals,
                                                         const double* __restrict__ x, double* __restrict__ y) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid / LANES;  // local scalar row in (g, i) order
  const int lane = (int)(gid % LANES);
  const int64_t nRows = P.nRowNodes * D;
  double s = 0.0;
  int64_t outRow = 0;
  if (row < nRows) {
    const int64_t g = row / D;
    const int i = (int)(row % D);
    const int32_t b0 = P.nbrPtr[g];
    const int nnb = P.nbrPtr[g + 1] - b0;
    const int64_t start = rawRowStart(P, g, i, nnb);
    const int len = nnb * D;
    outRow = localRowOf(P, g, i);
    for (int j = lane; j < len; j += LANES) {
      int slot, k;
      if (P.layout == LAYOUT_INTERLEAVED) {
        slot = j / D;
        k = j - slot * D;
      } else {
        k = j / nnb;
        slot = j - k * nnb;
      }
      const int64_t gb = P.nbrIdx[b0 + slot];
      s = fma(vals[start + j], x[dofOf(P.layout, D, P.nNodes, gb, k)], s);
    }
  }
  // fixed-order butterfly inside the lane group
#pragma unroll
  for (int w = LANES / 2; w > 0; w >>= 1) s += __shfl_down_sync(0xffffffffu, s, w, LANES);
  if (row < nRows && lane == 0) y[outRow] = s;
}

// ---------------------------------------------------------------------------------------------------
// Sync-free PCG building blocks (single GPU).  One warp per node-row: the D scalar rows of a node share
// their column set, so x is gathered once per column and used for D rows; consecutive lanes read consecutive
// values of each row (coalesced).  The p.q partial sums are produced by the same kernel.  All folds of
// per-block partials are done redundantly by every block of the consumer kernel in the same fixed order,
// so no separate reduction kernels and no host round trip are needed per iteration; a device-side `done`
// flag turns the remaining launches of a batch into no-ops once |r|^2 < threshold.
struct CgState {
  double rz[2];      // r.z ping-pong by iteration parity
  double rr;         // |r|^2 of the last completed iteration
  double threshold;  // tol^2 |b|^2
  int iter;          // completed iterations
  int done;          // 1 converged, 2 NaN
  int maxIter;
  int pad;
};

__global__ void cg2_init_kernel(CgState* st, const double* rzDev, const double* bbDev, double relTol, int maxIter,
                                unsigned int* arrive) {
  st->rz[0] = rzDev[0];
  st->rz[1] = 0.0;
  st->rr = bbDev[0];
  const double thr = relTol * relTol * bbDev[0];
  st->threshold = thr > 1e-300 ? thr : 1e-300;
  st->iter = 0;
  st->done = (bbDev[0] > 0.0 && bbDev[0] >= st->threshold) ? 0 : 1;
  st->maxIter = maxIter;
  *arrive = 0u;
}

// One node-row of y = A x: the D scalar rows of node g share their column set.  Leaves the row sums in lane 0.
// COHERENT: x entries may have been written by another GPU while this kernel runs (halo of a row-block partition):
// read them from L2 (ld.global.cg) instead of through the non-coherent path.
template <int D, bool COHERENT>
__device__ __forceinline__ void spmvNodeRow(const PatternView& P, const double* __restrict__ vals, const double* x,
                                            int64_t g, int lane, double (&s)[D]) {
  const int32_t b0 = P.nbrPtr[g];
  const int nnb = P.nbrPtr[g + 1] - b0;
  const int len = nnb * D;
  int64_t start[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    s[i] = 0.0;
    start[i] = rawRowStart(P, g, i, nnb);
  }
  auto loadX = [&](int64_t dof) -> double { return COHERENT ? __ldcg(x + dof) : __ldg(x + dof); };
  if (nnb <= 32) {
    // Fast path (rows of up to 96 entries: every Q1 mesh).  The column nodes of the node-row come from ONE coalesced
    // load and are handed out by shuffle, and the lane passes are unrolled, so all matrix loads and x gathers of the
    // node-row are in flight together instead of one dependent index -> x chain per pass.  Same per-lane summation
    // order as the generic loop below.
    constexpr int U = 3;
    const int32_t myCol = lane < nnb ? __ldg(P.nbrIdx + b0 + lane) : 0;
    double av[U][D], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = lane + 32 * u;
      const bool valid = j < len;
#pragma unroll
      for (int i = 0; i < D; ++i) av[u][i] = valid ? __ldcs(vals + start[i] + j) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      xv[u] = 0.0;
      if (32 * u < len) {  // warp-uniform
        const int j = lane + 32 * u;
        const bool valid = j < len;
        int slot, k;
        if (P.layout == LAYOUT_INTERLEAVED) {
          slot = j / D;
          k = j - slot * D;
        } else {
          k = j / nnb;
          slot = j - k * nnb;
        }
        const int32_t col = __shfl_sync(0xffffffffu, myCol, valid ? slot : 0);
        if (valid) xv[u] = loadX(dofOf(P.layout, D, P.nNodes, col, k));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int i = 0; i < D; ++i) s[i] = fma(av[u][i], xv[u], s[i]);
  } else
    for (int j = lane; j < len; j += 32) {
      int slot, k;
      if (P.layout == LAYOUT_INTERLEAVED) {
        slot = j / D;
        k = j - slot * D;
      } else {
        k = j / nnb;
        slot = j - k * nnb;
      }
      // the matrix is streamed exactly once: keep it out of L1 (__ldcs) so the gathered x stays resident there
      double av[D];
#pragma unroll
      for (int i = 0; i < D; ++i) av[i] = __ldcs(vals + start[i] + j);
      const double xv = loadX(dofOf(P.layout, D, P.nNodes, __ldg(P.nbrIdx + b0 + slot), k));
#pragma unroll
      for (int i = 0; i < D; ++i) s[i] = fma(av[i], xv, s[i]);
    }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) s[i] += __shfl_down_sync(0xffffffffu, s[i], w);
}

template <int D>
__global__ void __launch_bounds__(256, 6) spmv_node_dot_kernel(PatternView P, const double* __restrict__ vals,
                                                            const double* __restrict__ x, double* __restrict__ y,
                                                            const double* __restrict__ pLocal, double* partial,
                                                            const CgState* st) {
  if (st && (st->done || st->iter >= st->maxIter)) return;
  __shared__ double sh[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warpsTotal = (int64_t)gridDim.x * 8;
  double dot = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * 8 + warp; g < P.nRowNodes; g += warpsTotal) {
    double s[D];
    spmvNodeRow<D, false>(P, vals, x, g, lane, s);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const int64_t r = localRowOf(P, g, i);
        y[r] = s[i];
        if (pLocal) dot = fma(pLocal[r], s[i], dot);
      }
    }
  }
  if (partial) {
    if (lane == 0) sh[warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += sh[w];
      partial[blockIdx.x] = t;
    }
  }
}

// deterministic fold of np partials by one whole block of 256 threads (every block computes the same value):
// strided per-thread sums, a shuffle tree per warp, then all threads add the 8 warp sums in the same order
__device__ __forceinline__ double blockFold(const double* __restrict__ partial, int np, double* sh) {
  double s = 0.0;
  for (int i0 = threadIdx.x; i0 < np; i0 += 4 * blockDim.x) {  // four independent loads in flight
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = i < np ? partial[i] : 0.0;
    }
    s += (v[0] + v[1]) + (v[2] + v[3]);
  }
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) s += __shfl_down_sync(0xffffffffu, s, w);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  double r = 0.0;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) r += sh[w];
  __syncthreads();
  return r;
}

// x += alpha p ; r -= alpha q ; z = dinv r ; block partials of r.z and r.r.  alpha = rz / fold(pq partials).
__global__ void __launch_bounds__(256, 4) cg2_update_kernel(int64_t n, const CgState* st, const double* __restrict__ pqPartial,
                                                         int npq, const double* __restrict__ p,
                                                         const double* __restrict__ q, const double* __restrict__ dinv,
                                                         double* x, double* r, double* z, double* partial) {
  if (st->done || st->iter >= st->maxIter) return;
  __shared__ double sh[256];
  __shared__ double sh1[256];
  const double pq = blockFold(pqPartial, npq, sh);
  const double rz = st->rz[st->iter & 1];
  const double alpha = pq != 0.0 ? rz / pq : 0.0;
  double s0 = 0.0, s1 = 0.0;
  // three grid strides are issued together so that the loads of a thread overlap (n is ~3 strides for the C2 mesh)
  constexpr int UPD = 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += UPD * stride) {
    double pv[UPD], qv[UPD], xv[UPD], rv[UPD], dv[UPD];
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
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) {
      t0 += sh[w];
      t1 += sh1[w];
    }
    partial[blockIdx.x] = t0;
    partial[gridDim.x + blockIdx.x] = t1;
  }
}

// p = z + beta p, beta = rz_new / rz ; the LAST block to finish publishes rz_new, rr, iter+1 and the done flag
// (all blocks must have read the old state first, hence the arrival counter).
__global__ void __launch_bounds__(256) cg2_direction_kernel(int64_t n, CgState* st, const double* __restrict__ partial,
                                                            int np, const double* __restrict__ z, double* p,
                                                            unsigned int* arrive) {
  if (st->done || st->iter >= st->maxIter) return;
  __shared__ double sh[256];
  const double rzNew = blockFold(partial, np, sh);
  const double rr = blockFold(partial + np, np, sh);
  const int it = st->iter;
  const double rz = st->rz[it & 1];
  const double beta = rz != 0.0 ? rzNew / rz : 0.0;
  constexpr int UPD = 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += UPD * stride) {
    double pv[UPD], zv[UPD];
#pragma unroll
    for (int u = 0; u < UPD; ++u) {
      const int64_t i = i0 + u * stride;
      pv[u] = i < n ? p[i] : 0.0;
      zv[u] = i < n ? z[i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < UPD; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) p[i] = fma(beta, pv[u], zv[u]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(arrive, 1u);
    if (ticket == gridDim.x - 1) {
      *arrive = 0u;
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

template <int LANES>
__global__ void __launch_bounds__(256) spmv_csr_kernel(const int64_t* __restrict__ outer,
                                                       const int32_t* __restrict__ inner,
                                                       const double* __restrict__ vals, int64_t rows,
                                                       const double* __restrict__ x, double* __restrict__ y) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid / LANES;
  const int lane = (int)(gid % LANES);
  double s = 0.0;
  if (row < rows)
    for (int64_t p = outer[row] + lane; p < outer[row + 1]; p += LANES) s = fma(vals[p], x[inner[p]], s);
#pragma unroll
  for (int w = LANES / 2; w > 0; w >>= 1) s += __shfl_down_sync(0xffffffffu, s, w, LANES);
  if (row < rows && lane == 0) y[row] = s;
}

template <int D>
__global__ void diag_inv_block_kernel(PatternView P, const double* __restrict__ vals, double* dinv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.nBlocks) return;
  const int64_t g = P.nbrRow[b];
  if (P.nbrIdx[b] != g + P.rowBegin) return;
  const int nnb = P.nbrPtr[g + 1] - P.nbrPtr[g];
  const int slot = (int)(b - P.nbrPtr[g]);
  for (int i = 0; i < D; ++i) {
    const double v = vals[rawRowStart(P, g, i, nnb) + rawEntryOffset(P, slot, i, nnb)];
    dinv[localRowOf(P, g, i)] = v != 0.0 ? 1.0 / v : 1.0;  // Eigen's DiagonalPreconditioner uses 1 for zeros
  }
}

__global__ void diag_inv_csr_kernel(const int64_t* __restrict__ outer, const int32_t* __restrict__ inner,
                                    const double* __restrict__ vals, int64_t rows, double* dinv) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double v = 0.0;
  for (int64_t p = outer[r]; p < outer[r + 1]; ++p)
    if (inner[p] == r) v = vals[p];
  dinv[r] = v != 0.0 ? 1.0 / v : 1.0;
}

// scalars on the device: [0]=rz, [1]=pq, [2]=rz_new, [3]=rr, [4]=bb
// x += alpha p ; r -= alpha q ; z = dinv*r ; partial sums of r.z and r.r
__global__ void __launch_bounds__(256) cg_update_kernel(int64_t n, const double* __restrict__ scal,
                                                        const double* __restrict__ p, const double* __restrict__ q,
                                                        const double* __restrict__ dinv, double* x, double* r,
                                                        double* z, double* partial) {
  __shared__ double sh0[256], sh1[256];
  const double pq = scal[1];
  const double alpha = pq != 0.0 ? scal[0] / pq : 0.0;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    const double zi = dinv[i] * ri;
    z[i] = zi;
    s0 = fma(ri, zi, s0);
    s1 = fma(ri, ri, s1);
  }
  sh0[threadIdx.x] = s0;
  sh1[threadIdx.x] = s1;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
      sh0[threadIdx.x] += sh0[threadIdx.x + w];
      sh1[threadIdx.x] += sh1[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = sh0[0];
    partial[gridDim.x + blockIdx.x] = sh1[0];
  }
}

// folds partial[0..np) -> scal[2] (rz_new), partial[np..2np) -> scal[3] (rr); mirrors rr to the host
__global__ void __launch_bounds__(256) cg_fold2_kernel(const double* __restrict__ partial, int np, double* scal,
                                                       double* hostMirror) {
  __shared__ double sh0[256], sh1[256];
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < np; i += 256) {
    s0 += partial[i];
    s1 += partial[np + i];
  }
  sh0[threadIdx.x] = s0;
  sh1[threadIdx.x] = s1;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
      sh0[threadIdx.x] += sh0[threadIdx.x + w];
      sh1[threadIdx.x] += sh1[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    scal[2] = sh0[0];
    scal[3] = sh1[0];
    if (hostMirror) hostMirror[0] = sh1[0];
  }
}

// p = z + beta p with beta = rz_new/rz ; then rz <- rz_new (done by thread 0 of block 0 AFTER use:
// every block reads scal[0], scal[2] before any write because the write happens in a later kernel)
__global__ void cg_direction_kernel(int64_t n, const double* __restrict__ scal, const double* __restrict__ z,
                                    double* p) {
  const double rz = scal[0];
  const double beta = rz != 0.0 ? scal[2] / rz : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], z[i]);
}
__global__ void cg_shift_kernel(double* scal) { scal[0] = scal[2]; }

__global__ void vec_scale_kernel(int64_t n, double a, const double* __restrict__ x, double* y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a * x[i];
}
__global__ void vec_mul_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a[i] * b[i];
}
__global__ void vec_axpy_kernel(int64_t n, double a, const double* __restrict__ x, double* y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fma(a, x[i], y[i]);
}
// Truncated CG (trust-region inner solve) building blocks; the scalar recurrences live on the host
// (linearalgebra/truncatedconjugategradient.hh:114-163), so alpha / beta arrive by value.
//   x += alpha p ; r -= alpha q ; z = minv r ; block partials of r.z and r.r
__global__ void __launch_bounds__(256) tcg_update_kernel(int64_t n, double alpha, const double* __restrict__ p,
                                                         const double* __restrict__ q, const double* __restrict__ minv,
                                                         double* x, double* r, double* z, double* partial) {
  __shared__ double sh[8], sh1[8];
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    const double zi = minv[i] * ri;
    z[i] = zi;
    s0 = fma(ri, zi, s0);
    s1 = fma(ri, ri, s1);
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
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) {
      t0 += sh[w];
      t1 += sh1[w];
    }
    partial[blockIdx.x] = t0;
    partial[gridDim.x + blockIdx.x] = t1;
  }
}
// p = z + beta p
// ---------------------------------------------------------------------------------------------------
// Sync-free Steihaug-Toint truncated CG (linearalgebra/truncatedconjugategradient.hh:68-168 of the reference) in the
// style of the cg2_* kernels: three kernels per iteration, every decision of the reference's loop (negative curvature,
// trust-region boundary, kappa/theta rule, residual threshold, iteration cap) taken on the device from sums that every
// block folds in the same order; the host reads the state once per batch of iterations.
struct TcgState {
  CgState cg;  // rz[0] = absNew (r.z), rr = |r|^2, threshold, iter = completed CG steps, done, maxIter (read by the SpMV)
  double e_Pd, e_Pe, d_Pd, alpha, rhsNorm, Delta, kappa;
  int stop, mininner, i, pad;  // i = the reference's loop counter (starts at 1)
  unsigned int arrive[2];
};

__global__ void tcg2_init_kernel(TcgState* st, const double* absNewDev, const double* bbDev, double tol, double Delta,
                                 double kappa, int mininner, long long maxIters, int stopInit) {
  const double tiny = 2.2250738585072014e-308;
  const double rhsNorm = sqrt(bbDev[1]);
  st->cg.rz[0] = absNewDev[0];
  st->cg.rz[1] = 0.0;
  st->cg.rr = bbDev[1];
  const double thr = tol * tol * rhsNorm * rhsNorm;
  st->cg.threshold = thr > tiny ? thr : tiny;
  st->cg.iter = 0;
  st->cg.maxIter = 0x7fffffff;
  st->e_Pd = 0.0;
  st->e_Pe = 0.0;
  st->d_Pd = absNewDev[0];
  st->alpha = 0.0;
  st->rhsNorm = rhsNorm;
  st->Delta = Delta;
  st->kappa = kappa;
  st->stop = stopInit;
  st->mininner = mininner;
  st->arrive[0] = st->arrive[1] = 0u;
  // x = 0 solves it (:87-99: zero iterations), or the loop condition 1 < maxIters fails at once
  const bool solved = rhsNorm <= tiny || rhsNorm * rhsNorm < st->cg.threshold;
  st->i = solved ? 0 : 1;
  st->cg.done = (solved || maxIters <= 1) ? 1 : 0;
}

// d_Hd = fold(p.q); either the boundary step x += tau p (negative curvature / trust region exceeded, :122-134) or
// x += alpha p, r -= alpha q, z = Minv r with the block partials of r.z and r.r
__global__ void __launch_bounds__(256, 4)
    tcg2_update_kernel(int64_t n, TcgState* st, const double* __restrict__ pqPartial, int npq, const double* __restrict__ p,
                       const double* __restrict__ q, const double* __restrict__ minv, double* x, double* r, double* z,
                       double* partial) {
  if (st->cg.done) return;
  __shared__ double sh[256];
  __shared__ double sh1[256];
  const double d_Hd = blockFold(pqPartial, npq, sh);
  const double absNew = st->cg.rz[0];
  const double alpha = absNew / d_Hd;
  const double e_Pd = st->e_Pd, e_Pe = st->e_Pe, d_Pd = st->d_Pd, Delta = st->Delta;
  const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
  const bool boundary = d_Hd <= 0 || e_Pe_new >= Delta * Delta;
  const double step = boundary ? (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd : alpha;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(step, p[i], x[i]);
    if (!boundary) {
      const double ri = fma(-alpha, q[i], r[i]);
      r[i] = ri;
      const double zi = minv[i] * ri;
      z[i] = zi;
      s0 = fma(ri, zi, s0);
      s1 = fma(ri, ri, s1);
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
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) {
      t0 += sh[w];
      t1 += sh1[w];
    }
    partial[blockIdx.x] = t0;
    partial[gridDim.x + blockIdx.x] = t1;
    __threadfence();
    if (atomicAdd(&st->arrive[0], 1u) == gridDim.x - 1) {  // everybody has read the old state
      st->arrive[0] = 0u;
      st->alpha = alpha;
      if (boundary) {
        st->stop = d_Hd <= 0 ? 0 /* NEGATIVE_CURVATURE */ : 1 /* EXCEEDED_TRUST_REGION */;
        __threadfence();
        st->cg.done = 1;
      } else {
        st->e_Pe = e_Pe_new;
      }
    }
  }
}

// stopping rules after the residual update (:142-151), then beta, the recurrences of e_Pd / d_Pd and p = z + beta p
__global__ void __launch_bounds__(256)
    tcg2_direction_kernel(int64_t n, TcgState* st, const double* __restrict__ partial, int np, long long maxIters,
                          const double* __restrict__ z, double* p) {
  if (st->cg.done) return;
  __shared__ double sh[256];
  const double absNewNext = blockFold(partial, np, sh);
  const double rr = blockFold(partial + np, np, sh);
  const double resNorm = sqrt(rr);
  const int i = st->i;
  const double rhsNorm = st->rhsNorm, kappa = st->kappa;
  int stopNow = -1;  // -1: continue
  if (!(resNorm == resNorm))
    stopNow = 100;  // NaN
  else if (i >= st->mininner && resNorm <= rhsNorm * fmin(rhsNorm, kappa))
    stopNow = kappa < rhsNorm ? 2 /* REACHED_KAPPA_LINEAR */ : 3 /* REACHED_THETA_SUPERLINEAR */;
  else if (resNorm < st->cg.threshold)  // the reference compares the NORM with the squared-norm threshold (:151)
    stopNow = st->stop;
  const double absOld = st->cg.rz[0];
  const double beta = absNewNext / absOld;
  if (stopNow < 0) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
      p[k] = fma(beta, p[k], z[k]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&st->arrive[1], 1u) == gridDim.x - 1) {
      st->arrive[1] = 0u;
      st->cg.rr = rr;
      if (stopNow >= 0) {
        st->stop = stopNow;
        __threadfence();
        st->cg.done = stopNow == 100 ? 2 : 1;
      } else {
        const double alpha = st->alpha, d_Pd = st->d_Pd;
        st->e_Pd = beta * (st->e_Pd + alpha * d_Pd);
        st->d_Pd = absNewNext + beta * beta * d_Pd;
        st->cg.rz[0] = absNewNext;
        st->cg.iter = st->cg.iter + 1;
        st->i = i + 1;
        if ((long long)(i + 1) >= maxIters) {
          __threadfence();
          st->cg.done = 1;  // loop condition i < maxIters (:113)
        }
      }
    }
  }
}

__global__ void tcg_direction_kernel(int64_t n, double beta, const double* __restrict__ z, double* p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], z[i]);
}
__global__ void vec_fill_kernel(int64_t n, double a, double* y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a;
}

// full <- reduced expansion / reduced <- full contraction (createFullVector / createReducedVector,
// assembler/simpleassemblers.inl:26-57)
__global__ void expand_reduced_kernel(int64_t n, const uint8_t* __restrict__ flags, const int32_t* __restrict__ cbelow,
                                      const double* __restrict__ red, double* full) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) full[i] = flags[i] ? 0.0 : red[i - cbelow[i]];
}
__global__ void contract_full_kernel(int64_t n, const uint8_t* __restrict__ flags, const int32_t* __restrict__ cbelow,
                                     const double* __restrict__ full, double* red) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !flags[i]) red[i - cbelow[i]] = full[i];
}

__global__ void zero_flagged_kernel(int64_t n, const uint8_t* __restrict__ flags, double* v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) v[i] = 0.0;
}

// FP64 FMA peak probe: 8 independent chains per thread
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c);
    a1 = fma(a1, b, c);
    a2 = fma(a2, b, c);
    a3 = fma(a3, b, c);
    a4 = fma(a4, b, c);
    a5 = fma(a5, b, c);
    a6 = fma(a6, b, c);
    a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace ikb
