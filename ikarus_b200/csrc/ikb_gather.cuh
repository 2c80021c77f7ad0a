// Deterministic segmented gather of staged element blocks into the global matrix / vector,
// with the Dirichlet mode applied on the fly.
//
// Replaces the scatter loops of SparseFlatAssembler::assembleRawMatrixImpl / getMatrixImpl /
// getReducedMatrixImpl and VectorFlatAssembler::get{Raw,,Reduced}VectorImpl
// (ikarus/assembler/simpleassemblers.inl:59-204).  One thread per pattern block sums the
// contributions of its elements in ascending element order (no float atomics), so every run
// and every GPU count produces bit-identical values.
#pragma once
#include "ikb_internal.cuh"
#include "ikb_pattern.cuh"

namespace ikb {

struct GatherArgs {
  PatternView P;
  const int32_t* cptr;
  const uint32_t* csrc;
  const double* Kst;
  const double* Rst;
  const double* fext;   // may be null
  double fextScale;
  const uint8_t* flags; // may be null (no constraints)
  double* vals;         // output value array for mode dbc (null: skip matrix)
  double* vec;          // output vector for mode dbc (null: skip vector)
  int dbc;
  int npair;
  int nn;               // nodes per element
  // reduced mode
  const uint16_t* freeCnt;
  const uint16_t* freeTot;
  const int64_t* redRowStart;
  const int32_t* cbelow;
  int64_t redVecOffset;  // index of the first local free row in the reduced vector
};

template <int D>
__global__ void __launch_bounds__(256) gather_kernel(GatherArgs G) {
  constexpr int DD = D * D;
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= G.P.nBlocks) return;
  const PatternView& P = G.P;
  const int64_t g = P.nbrRow[b];
  const int64_t gb = P.nbrIdx[b];
  const int32_t c0 = G.cptr[b], c1 = G.cptr[b + 1];
  const bool diag = (g + P.rowBegin) == gb;

  bool rowFixed[D], colFixed[D];
  int64_t rowDof[D], colDof[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    rowDof[i] = dofOf(P.layout, D, P.nNodes, g + P.rowBegin, i);
    colDof[i] = dofOf(P.layout, D, P.nNodes, gb, i);
    rowFixed[i] = G.flags ? (G.flags[rowDof[i]] != 0) : false;
    colFixed[i] = G.flags ? (G.flags[colDof[i]] != 0) : false;
  }

  if (G.vals) {
    double acc[DD];
#pragma unroll
    for (int q = 0; q < DD; ++q) acc[q] = 0.0;
    for (int32_t c = c0; c < c1; ++c) {
      const uint32_t s = G.csrc[c];
      const double* p = G.Kst + (size_t)(s & SRC_MASK) * DD;
      if (s & SRC_TRANSPOSE) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int k = 0; k < D; ++k) acc[i * D + k] += p[k * D + i];
      } else {
#pragma unroll
        for (int q = 0; q < DD; ++q) acc[q] += p[q];
      }
    }
    const int nnb = P.nbrPtr[g + 1] - P.nbrPtr[g];
    const int slot = (int)(b - P.nbrPtr[g]);
    if (G.dbc == IKB_DBC_REDUCED) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if (rowFixed[i]) continue;
        const int64_t start = G.redRowStart[localRowOf(P, g, i)];
#pragma unroll
        for (int k = 0; k < D; ++k) {
          if (colFixed[k]) continue;
          G.vals[start + reducedRank(P, G.flags, G.freeCnt, G.freeTot, g, b, gb, k)] = acc[i * D + k];
        }
      }
    } else {
      const bool full = G.dbc == IKB_DBC_FULL;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const int64_t start = rawRowStart(P, g, i, nnb);
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double v = acc[i * D + k];
          // Full: zero constrained rows and columns, unit diagonal (simpleassemblers.inl:159-167)
          if (full && (rowFixed[i] || colFixed[k])) v = (diag && i == k) ? 1.0 : 0.0;
          G.vals[start + rawEntryOffset(P, slot, k, nnb)] = v;
        }
      }
    }
  }

  if (G.vec && diag) {
    // the diagonal block's contributions are exactly (element, la == lb) for every element
    // touching this node, in element order: reuse them for R
    double r[D];
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = 0.0;
    for (int32_t c = c0; c < c1; ++c) {
      const uint32_t s = G.csrc[c] & SRC_MASK;  // = e*npair + la   (k = 0)
      const int64_t e = s / G.npair;
      const int la = (int)(s - e * G.npair);
      const double* p = G.Rst + (size_t)e * (G.nn * D) + la * D;
#pragma unroll
      for (int i = 0; i < D; ++i) r[i] += p[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double v = r[i];
      if (G.fext) v -= G.fextScale * G.fext[rowDof[i]];
      if (G.dbc == IKB_DBC_REDUCED) {
        if (!rowFixed[i]) G.vec[rowDof[i] - G.cbelow[rowDof[i]] - G.redVecOffset] = v;
      } else {
        if (G.dbc == IKB_DBC_FULL && rowFixed[i]) v = 0.0;  // simpleassemblers.inl:90-92
        G.vec[localRowOf(P, g, i)] = v;
      }
    }
  }
}

// ------------------------------------------------------------------ deterministic reductions
// two-stage fixed-shape tree: partials[blockIdx] then a single block folds them in index order
template <int MODE>  // 0: sum x ; 1: sum x*y ; 2: sum x*x
__global__ void __launch_bounds__(256) reduce_stage1(const double* __restrict__ x, const double* __restrict__ y,
                                                     int64_t n, double* partial) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = x[i];
    if (MODE == 0)
      s += a;
    else if (MODE == 1)
      s = fma(a, y[i], s);
    else
      s = fma(a, a, s);
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(256) reduce_stage2(const double* __restrict__ partial, int np, double* out,
                                                     double scale, const double* addend) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < np; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = scale * sh[0] + (addend ? addend[0] : 0.0);
}

// element-partitioned runs: an element's energy counts on the rank that owns its first node
__global__ void mask_unowned_energy_kernel(const int32_t* __restrict__ elemNode0, int64_t nElem, int64_t rowBegin,
                                           int64_t rowEnd, double* Est) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nElem && (elemNode0[e] < rowBegin || elemNode0[e] >= rowEnd)) Est[e] = 0.0;
}

// dense mirror of a CSR matrix (DenseFlatAssembler, simpleassemblers.inl:301-375), column-major
__global__ void csr_to_dense_kernel(const int64_t* __restrict__ outer, const int32_t* __restrict__ inner,
                                    const double* __restrict__ vals, int64_t rows, double* dense) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  for (int64_t p = outer[r]; p < outer[r + 1]; ++p) dense[(size_t)inner[p] * rows + r] = vals[p];
}

}  // namespace ikb
