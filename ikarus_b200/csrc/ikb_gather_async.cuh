// Row-pipelined pull gather for the Q1 kinds (Quad4, Hex8): staged K_e blocks travel through cp.async and shared memory.
//
// Same job, same contribution lists and the same summation order as gather_pull_kernel (ikb_gather.cuh; the scatter
// loops of SparseFlatAssembler::assembleRawMatrixImpl / getMatrixImpl, ikarus/assembler/simpleassemblers.inl:120-167),
// hence the same bits.  What changes is how the latency is paid.  ncu shows the register variant neither DRAM- nor
// L1-bound (DRAM 44 %; a staging layout that cut its L1 wavefronts by 29 % left the time unchanged): a row is a chain of
// dependent round trips -- nbrPtr -> cptr -> csrc -> nine groups of value loads, four 8-byte loads per lane at a time.
// Here a warp owns ROWS consecutive node-rows and runs them as a software pipeline:
//   * the row descriptors (first block, first code) of all its rows are fetched up front, one row per lane;
//   * while the staged blocks of row r -- all of them, 64 x 72 bytes for an interior Hex8 row -- are in flight to shared
//     memory (cp.async, one 8-byte piece per lane, written where the entry belongs after transposition; no registers hold
//     them), the codes
//     and per-block counts of row r+1 are loaded;
//   * the entries are then summed from shared memory, one lane per matrix entry, in ascending element order.
// One DRAM round trip per row instead of eleven, 4.6 KB outstanding per warp.  DBCOption::Full needs the flags of the
// row and of its column nodes only where one of them is constrained: rowSlow[] (one byte per row, recomputed when the
// flags or the pattern change) keeps those loads off the common path.  Rows that do not fit the buffers (more than 31
// pattern blocks or ASYNC_CAP contributions: unstructured patches) are served by pullRow of the register variant.
#pragma once
#include "ikb_gather.cuh"

namespace ikb {

constexpr int ASYNC_CAP = 64;    // staged blocks of a row held in shared memory (an interior structured Hex8 row has 64)
constexpr int ASYNC_WARPS = 4;   // warps per CTA
constexpr int ASYNC_ROWS = 8;    // consecutive node-rows per warp

template <int D>
struct AsyncSlot {
  static constexpr int BYTES = 8 * D * D;  // one staged block, row-major after the copy
};
__device__ __forceinline__ void cpAsync8(uint32_t dst, uint64_t src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// rowSlow[g] = 1 when the row node g or one of its column nodes carries a constrained dof
template <bool INTERLEAVED>
__global__ void row_slow_kernel(PatternView P, const uint8_t* __restrict__ flags, uint8_t* rowSlow) {
  constexpr int LAYOUT = INTERLEAVED ? LAYOUT_INTERLEAVED : LAYOUT_LEXICOGRAPHIC;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.nRowNodes) return;
  bool f = false;
  for (int32_t b = P.nbrPtr[g]; b < P.nbrPtr[g + 1]; ++b)
    for (int c = 0; c < P.dim; ++c) f |= flags[dofOf(LAYOUT, P.dim, P.nNodes, P.nbrIdx[b], c)] != 0;
  // (the row node is one of its own column nodes)
  rowSlow[g] = f ? 1 : 0;
}

template <int D, int DBC, bool INTERLEAVED, bool IDX32>
__global__ void __launch_bounds__(32 * ASYNC_WARPS, 9)
    gather_rows_async_kernel(GatherArgs G, const int32_t* __restrict__ cptr, const uint32_t* __restrict__ csrc,
                             const uint8_t* __restrict__ rowSlow) {
  static_assert(DBC != IKB_DBC_REDUCED, "the reduced mode is served by gather_pull_kernel");
  constexpr int DD = D * D, BPW = 32 / DD;
  constexpr int SLOT = AsyncSlot<D>::BYTES;
  constexpr int LAYOUT = INTERLEAVED ? LAYOUT_INTERLEAVED : LAYOUT_LEXICOGRAPHIC;
  constexpr unsigned FULLMASK = 0xffffffffu;
  __shared__ uint32_t codeBuf[ASYNC_WARPS][2][ASYNC_CAP];
  __shared__ __align__(8) unsigned char dataBuf[ASYNC_WARPS][ASYNC_CAP * SLOT];
  static_assert(ASYNC_CAP * SLOT >= PULL_CAP * 4, "the data buffer doubles as the code buffer of the fallback");
  const PatternView& P = G.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gFirst = ((int64_t)blockIdx.x * ASYNC_WARPS + warp) * ASYNC_ROWS;
  if (gFirst >= P.nRowNodes) return;
  const int nR = (int)min((int64_t)ASYNC_ROWS, P.nRowNodes - gFirst);
  unsigned char* dbuf = dataBuf[warp];
  const uint32_t dbase = (uint32_t)__cvta_generic_to_shared(dbuf);
  const uint64_t kst = reinterpret_cast<uint64_t>(G.Kst);

  // descriptors of the warp's rows, one row per lane (lane nR holds the end of the last row)
  const int32_t myB0 = P.nbrPtr[gFirst + min(lane, nR)];
  const int32_t myC0 = cptr[myB0];
  const int mySlow = (DBC != IKB_DBC_RAW && lane < nR) ? (int)rowSlow[gFirst + lane] : 0;

  const int q = lane / DD, v = lane - q * DD, i = v / D, k = v - i * D;
  const bool active = q < BPW;

  // contribution code (e*npair + p, bit 31 = transposed) -> 4-byte offset of the staged block | transposed
  // (64-bit form: 2*(e*npair+p) | transposed, the block at 8*DD bytes per index)
  auto toWord = [&](uint32_t c) -> uint32_t {
    return IDX32 ? (c & SRC_MASK) * (uint32_t)(2 * DD) + (c >> 31) : __funnelshift_l(c, c, 1);
  };

  uint32_t cw[2] = {0u, 0u};  // codes of the row being prefetched
  int32_t cpv = 0;            // its per-block code pointers (lane l: block l, lane nnb: end)
  bool ok = false;            // it fits the buffers (warp-uniform)
  auto loadRow = [&](int r) {
    const int32_t b0 = __shfl_sync(FULLMASK, myB0, r), b1 = __shfl_sync(FULLMASK, myB0, r + 1);
    const int32_t c0 = __shfl_sync(FULLMASK, myC0, r), c1 = __shfl_sync(FULLMASK, myC0, r + 1);
    const int nnb = b1 - b0, ncodes = c1 - c0;
    ok = nnb <= 31 && ncodes <= ASYNC_CAP;
    if (ok) {
      cpv = cptr[b0 + min(lane, nnb)];
      cw[0] = lane < ncodes ? csrc[c0 + lane] : 0u;
      cw[1] = lane + 32 < ncodes ? csrc[c0 + 32 + lane] : 0u;
    }
  };
  loadRow(0);

  // Copy pieces: one lane per matrix entry, BPW staged blocks per pass, 8 bytes each.  The piece lands where the entry
  // belongs AFTER transposition, so the data buffer holds every contribution as a plain row-major block at DD doubles
  // per slot and the summation below needs no per-contribution fix-up.
  const uint32_t srcEnt = 8u * (uint32_t)v;                         // entry v = (i,k) of the staged block ...
  const uint32_t dstEnt = 8u * (uint32_t)v, dstEntT = 8u * (uint32_t)(k * D + i);  // ... goes to (i,k), or to (k,i)
  const unsigned char* lbase = dbuf + 8 * v;

#pragma unroll 1
  for (int r = 0; r < nR; ++r) {
    const int64_t g = gFirst + r;
    const int32_t b0 = __shfl_sync(FULLMASK, myB0, r), c0 = __shfl_sync(FULLMASK, myC0, r);
    const int nnb = __shfl_sync(FULLMASK, myB0, r + 1) - b0;
    const int ncodes = __shfl_sync(FULLMASK, myC0, r + 1) - c0;
    const bool okCur = ok;
    const int32_t cpCur = cpv;
    uint32_t* cb = codeBuf[warp][r & 1];
    if (okCur) {
      static_assert(ASYNC_CAP == 64, "two codes per lane");
      cb[lane] = toWord(cw[0]);
      cb[lane + 32] = toWord(cw[1]);
      __syncwarp();
      if (active) {
        const uint64_t srcLane = kst + srcEnt;
        uint32_t dst = dbase + (uint32_t)q * SLOT;
        for (int c = q; c < ncodes; c += BPW, dst += BPW * SLOT) {
          const uint32_t w = cb[c];
          const uint64_t src = IDX32 ? srcLane + (uint64_t)(w & ~1u) * 4u : srcLane + (uint64_t)(w >> 1) * (uint64_t)(8 * DD);
          cpAsync8(dst + ((w & 1u) ? dstEntT : dstEnt), src);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // the next row's codes and counts travel while the copies fly
    if (r + 1 < nR) loadRow(r + 1);
    if (!okCur) {
      // (pullRow brings its own lane constants; the data buffer serves as its code buffer)
      pullRow<D, DBC, INTERLEAVED, IDX32, false, false>(G, cptr, csrc, g, reinterpret_cast<uint32_t*>(dbuf), lane);
      __syncwarp();
      continue;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    const int64_t gGlobal = g + P.rowBegin;
    const bool slow = (DBC != IKB_DBC_RAW) && __shfl_sync(FULLMASK, mySlow, r) != 0;
    bool rowFixed = false;
    if (slow) rowFixed = G.flags[dofOf(LAYOUT, D, P.nNodes, gGlobal, i)] != 0;
    double* dst = G.vals + (INTERLEAVED ? (int64_t)DD * b0 + (int64_t)i * D * nnb + k
                                        : (int64_t)i * D * P.nBlocks + (int64_t)D * b0 + (int64_t)k * nnb);
    constexpr int SSTRIDE = INTERLEAVED ? D : 1;  // doubles between the entries (i,k) of consecutive slots
    // per block: (offset of its first code in the row's list) | (number of codes) << 16
    const int32_t cnext = __shfl_down_sync(FULLMASK, cpCur, 1);
    const uint32_t pack = (lane < nnb) ? (uint32_t)(cpCur - c0) | ((uint32_t)(cnext - cpCur) << 16) : 0u;
    for (int s = 0; s < nnb; s += BPW) {
      const int sl = s + q;
      const uint32_t pk = __shfl_sync(FULLMASK, pack, sl & 31);
      const bool mine = active && sl < nnb;
      const int n = mine ? (int)(pk >> 16) : 0;
      const unsigned char* src = lbase + (pk & 0xffffu) * SLOT;
      double acc = 0.0;  // ascending element order, as the reference's serial loop adds them
      for (int j = 0; j < n; ++j) acc += *reinterpret_cast<const double*>(src + j * SLOT);
      if (mine) {
        if (slow) {
          // Full: zero constrained rows and columns, unit diagonal (simpleassemblers.inl:159-167)
          const int64_t gb = P.nbrIdx[b0 + sl];
          const bool colFixed = G.flags[dofOf(LAYOUT, D, P.nNodes, gb, k)] != 0;
          if (rowFixed || colFixed) acc = (gb == gGlobal && i == k) ? 1.0 : 0.0;
        }
        dst[(int64_t)SSTRIDE * sl] = acc;
      }
    }
    __syncwarp();  // the buffers are rewritten for the next row
  }
}

}  // namespace ikb
