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
