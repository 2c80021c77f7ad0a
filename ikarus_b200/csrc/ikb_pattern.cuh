// Sparsity pattern + deterministic gather map, built on the device.
//
// Replaces SparseFlatAssembler::createOccupationPattern / createReducedOccupationPattern /
// createLinearDOFsPerElement(Reduced) (ikarus/assembler/simpleassemblers.inl:206-299), which
// build a triplet list of every dof x dof pair of every element, let Eigen's setFromTriplets
// sort it, and then look up the value-array position of every K_e entry.
//
// Here: one (node-row, node-col) key per element node pair, one stable radix sort.  The sorted
// unique keys ARE the pattern (node-block form; the scalar CSR/CSC arrays Eigen would hold
// follow analytically, see PatternView), and the sorted payload IS the gather map: for every
// pattern block the list of (element, local pair) contributions in ascending element order --
// the order in which the reference's serial element loop adds them
// (simpleassemblers.inl:126-136), so the summation order per entry is the reference's.
#pragma once
#include <cub/cub.cuh>

#include "ikb_internal.cuh"

namespace ikb {

// code of the staged block holding K_e[la-rows, lb-cols] (maybe transposed)
__host__ __device__ inline uint32_t pairCode(int n, int npair, int64_t e, int la, int lb) {
  int k1 = lb - la;
  if (k1 < 0) k1 += n;
  const int half = n / 2;
  bool direct;
  if ((n & 1) == 0)
    direct = (k1 < half) || (k1 == half && la < half);
  else
    direct = (k1 <= half);
  uint32_t code;
  if (direct)
    code = (uint32_t)(e * npair + (int64_t)k1 * n + la);
  else
    code = (uint32_t)(e * npair + (int64_t)(n - k1) * n + lb) | SRC_TRANSPOSE;
  return code;
}

__global__ void gen_pairs_kernel(const int32_t* __restrict__ elemNode, int64_t nElem, int n, int npair, int64_t nNodes,
                                 int64_t rowBegin, int64_t rowEnd, uint64_t* keys, uint32_t* vals) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = nElem * n * n;
  if (idx >= total) return;
  // idx = (la*n + lb)*nElem + e  -> coalesced reads of elemNode
  const int64_t e = idx % nElem;
  const int pq = (int)(idx / nElem);
  const int la = pq / n, lb = pq % n;
  const int64_t ga = elemNode[(size_t)la * nElem + e];
  const int64_t gb = elemNode[(size_t)lb * nElem + e];
  const int64_t out = (e * n + la) * n + lb;  // element-major so the stable sort keeps element order
  if (ga < rowBegin || ga >= rowEnd) {
    keys[out] = ~0ull;
    vals[out] = 0;
  } else {
    keys[out] = (uint64_t)(ga - rowBegin) * (uint64_t)nNodes + (uint64_t)gb;
    vals[out] = pairCode(n, npair, e, la, lb);
  }
}

__global__ void decode_blocks_kernel(const uint64_t* __restrict__ ukeys, int64_t nBlocks, int64_t nNodes,
                                     int32_t* nbrIdx, int32_t* nbrRow) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nBlocks) return;
  const uint64_t k = ukeys[b];
  nbrRow[b] = (int32_t)(k / (uint64_t)nNodes);
  nbrIdx[b] = (int32_t)(k % (uint64_t)nNodes);
}

// nbrPtr[g] = first block whose row >= g
__global__ void row_ptr_kernel(const int32_t* __restrict__ nbrRow, int64_t nBlocks, int64_t nRowNodes,
                               int32_t* nbrPtr) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g > nRowNodes) return;
  int64_t lo = 0, hi = nBlocks;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (nbrRow[mid] < g)
      lo = mid + 1;
    else
      hi = mid;
  }
  nbrPtr[g] = (int32_t)lo;
}

// ---------------------------------------------------------------------------- node -> element adjacency
// The diagonal block (g,g) of row g lists exactly the (element, la) pairs touching node g, in element order.
__device__ __forceinline__ int32_t diagBlockOf(const PatternView& P, int64_t g) {
  int32_t lo = P.nbrPtr[g], hi = P.nbrPtr[g + 1];
  const int32_t target = (int32_t)(g + P.rowBegin);
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (P.nbrIdx[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

__global__ void adj_count_kernel(PatternView P, const int32_t* __restrict__ cptr, int32_t* adjCount, int32_t* rowLen) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.nRowNodes) return;
  const int32_t b = diagBlockOf(P, g);
  adjCount[g] = cptr[b + 1] - cptr[b];
  rowLen[g] = P.nbrPtr[g + 1] - P.nbrPtr[g];
}

__global__ void adj_fill_kernel(PatternView P, const int32_t* __restrict__ cptr, const uint32_t* __restrict__ csrc,
                                const int32_t* __restrict__ adjPtr, const int32_t* __restrict__ elemNode,
                                int64_t nElem, int n, int npair, uint32_t* adjCode, uint8_t* slotTab) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.nRowNodes) return;
  const int32_t b0 = P.nbrPtr[g], b1 = P.nbrPtr[g + 1];
  const int32_t bd = diagBlockOf(P, g);
  const int32_t c0 = cptr[bd];
  const int32_t cnt = cptr[bd + 1] - c0;
  const int32_t a0 = adjPtr[g];
  for (int32_t j = 0; j < cnt; ++j) {
    const uint32_t s = csrc[c0 + j] & SRC_MASK;  // diagonal pair: e*npair + la
    const int64_t e = s / (uint32_t)npair;
    const int la = (int)(s - e * npair);
    adjCode[a0 + j] = (uint32_t)(e * n + la);
    for (int lb = 0; lb < n; ++lb) {
      const int32_t nb = elemNode[(size_t)lb * nElem + e];
      int32_t l2 = b0, h2 = b1;
      while (l2 < h2) {
        const int32_t mid = (l2 + h2) >> 1;
        if (P.nbrIdx[mid] < nb)
          l2 = mid + 1;
        else
          h2 = mid;
      }
      slotTab[(size_t)(a0 + j) * n + lb] = (uint8_t)(l2 - b0);
    }
  }
}

// ---------------------------------------------------------------------------- reduced mode
__global__ void flags_to_int_kernel(const uint8_t* flags, int64_t n, int32_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = flags[i] ? 1 : 0;
}

// per node-row: running count of free columns per component before each slot
__global__ void free_counts_kernel(PatternView P, const uint8_t* __restrict__ flags, uint16_t* freeCnt,
                                   uint16_t* freeTot) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.nRowNodes) return;
  const int d = P.dim;
  int cnt[3] = {0, 0, 0};
  for (int32_t b = P.nbrPtr[g]; b < P.nbrPtr[g + 1]; ++b) {
    const int64_t gb = P.nbrIdx[b];
    for (int k = 0; k < d; ++k) {
      freeCnt[(size_t)b * d + k] = (uint16_t)cnt[k];
      if (!flags[dofOf(P.layout, d, P.nNodes, gb, k)]) cnt[k]++;
    }
  }
  for (int k = 0; k < d; ++k) freeTot[(size_t)g * d + k] = (uint16_t)cnt[k];
}

// local scalar row index in global row order: interleaved -> d*gLocal+i, lexicographic -> i*nRowNodes+gLocal
__host__ __device__ inline int64_t localRowOf(const PatternView& P, int64_t gLocal, int i) {
  return P.layout == LAYOUT_INTERLEAVED ? gLocal * P.dim + i : (int64_t)i * P.nRowNodes + gLocal;
}

__global__ void row_free_count_kernel(PatternView P, const uint8_t* __restrict__ flags,
                                      const uint16_t* __restrict__ freeTot, int64_t* rowCount) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P.nRowNodes * P.dim) return;
  const int64_t g = idx / P.dim;
  const int i = (int)(idx % P.dim);
  const int64_t r = dofOf(P.layout, P.dim, P.nNodes, g + P.rowBegin, i);
  int64_t c = 0;
  if (!flags[r])
    for (int k = 0; k < P.dim; ++k) c += freeTot[(size_t)g * P.dim + k];
  rowCount[localRowOf(P, g, i)] = c;
}

// rank of entry (slot,k) of block b within its reduced row
__device__ __forceinline__ int reducedRank(const PatternView& P, const uint8_t* flags, const uint16_t* freeCnt,
                                           const uint16_t* freeTot, int64_t gLocal, int64_t b, int64_t gb, int k) {
  const int d = P.dim;
  int rank = 0;
  if (P.layout == LAYOUT_INTERLEAVED) {
    for (int kk = 0; kk < d; ++kk) rank += freeCnt[(size_t)b * d + kk];
    for (int kk = 0; kk < k; ++kk) rank += flags[dofOf(P.layout, d, P.nNodes, gb, kk)] ? 0 : 1;
  } else {
    for (int kk = 0; kk < k; ++kk) rank += freeTot[(size_t)gLocal * d + kk];
    rank += freeCnt[(size_t)b * d + k];
  }
  return rank;
}

__global__ void reduced_inner_kernel(PatternView P, const uint8_t* __restrict__ flags,
                                     const uint16_t* __restrict__ freeCnt, const uint16_t* __restrict__ freeTot,
                                     const int64_t* __restrict__ redRowStart, const int32_t* __restrict__ cbelow,
                                     int32_t* redInner) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.nBlocks) return;
  const int d = P.dim;
  const int64_t g = P.nbrRow[b];
  const int64_t gb = P.nbrIdx[b];
  for (int i = 0; i < d; ++i) {
    const int64_t r = dofOf(P.layout, d, P.nNodes, g + P.rowBegin, i);
    if (flags[r]) continue;
    const int64_t start = redRowStart[localRowOf(P, g, i)];
    for (int k = 0; k < d; ++k) {
      const int64_t c = dofOf(P.layout, d, P.nNodes, gb, k);
      if (flags[c]) continue;
      redInner[start + reducedRank(P, flags, freeCnt, freeTot, g, b, gb, k)] = (int32_t)(c - cbelow[c]);
    }
  }
}

// compact redRowStart (all local rows) -> redOuter (free rows only, in row order)
__global__ void reduced_outer_kernel(const uint8_t* __restrict__ flags, const int32_t* __restrict__ cbelow,
                                     const int64_t* __restrict__ redRowStart, int64_t nRows, int64_t rowDofBegin,
                                     int64_t* redOuter, int64_t nnzRed, int64_t nRed) {
  const int64_t lr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lr == 0) redOuter[nRed] = nnzRed;
  if (lr >= nRows) return;
  const int64_t r = rowDofBegin + lr;  // single-GPU / interleaved slab: local rows are a contiguous dof range
  if (flags[r]) return;
  redOuter[r - cbelow[r] - (rowDofBegin - cbelow[rowDofBegin])] = redRowStart[lr];
}

// explicit raw pattern for download (outer/inner of the Raw/Full matrix)
__global__ void raw_pattern_kernel(PatternView P, int64_t* outer, int32_t* inner) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.nBlocks) return;
  const int d = P.dim;
  const int64_t g = P.nbrRow[b];
  const int64_t gb = P.nbrIdx[b];
  const int nnb = P.nbrPtr[g + 1] - P.nbrPtr[g];
  const int slot = (int)(b - P.nbrPtr[g]);
  for (int i = 0; i < d; ++i) {
    const int64_t start = rawRowStart(P, g, i, nnb);
    if (slot == 0) outer[localRowOf(P, g, i)] = start;
    for (int k = 0; k < d; ++k)
      inner[start + rawEntryOffset(P, slot, k, nnb)] = (int32_t)dofOf(P.layout, d, P.nNodes, gb, k);
  }
  if (b == 0) outer[P.nRowNodes * d] = (int64_t)d * d * P.nBlocks;
}

}  // namespace ikb
