// Deterministic segmented gather of staged element blocks into the global matrix / vector,
// with the Dirichlet mode applied on the fly.
//
// Replaces the scatter loops of SparseFlatAssembler::assembleRawMatrixImpl / getMatrixImpl /
// getReducedMatrixImpl and VectorFlatAssembler::get{Raw,,Reduced}VectorImpl
// (ikarus/assembler/simpleassemblers.inl:59-204).  One thread per pattern block sums the
// contributions of its elements in ascending element order (no float atomics), so every run
// and every GPU count produces bit-identical values.
#pragma once
#include <vector>
#include "ikb_internal.cuh"
#include "ikb_pattern.cuh"

namespace ikb {

struct GatherArgs {
  PatternView P;
  const int32_t* adjPtr;    // [nRowNodes+1]
  const uint32_t* adjCode;  // [nAdj] e*N + la, ascending element order per node
  const uint8_t* slotTab;   // [nAdj][N] slot of element node lb in the pattern row of the node
  const double* Kst;
  const double* Rst;
  const double* fext;   // may be null
  double fextScale;
  const uint8_t* flags; // may be null (no constraints)
  double* vals;         // output value array for mode dbc (null: skip matrix)
  double* vec;          // output vector for mode dbc (null: skip vector)
  int dbc;
  int npair;
  int maxOut;           // doubles of shared memory per warp
  const int16_t* offTab;  // [N][CHUNK] staged-K_e offsets, see gatherOffsetTable()
  // reduced mode
  const uint16_t* freeCnt;
  const uint16_t* freeTot;
  const int64_t* redRowStart;
  const int32_t* cbelow;
  int64_t redVecOffset;  // index of the first local free row in the reduced vector
  int64_t rowFirst = 0, rowEnd = -1;  // pull gather: local node-rows [rowFirst, rowEnd) (rowEnd < 0: all of them)
  int pullStageMax;      // pull gather: chunks with at most this many codes are staged in shared memory (<= PULL_CAP)
  // mirrored pull (Raw/Full): a row computes its blocks (g, g') with g' >= g (and those whose column node belongs to
  // another rank) and stores each off-diagonal one a second time, transposed, as block (g', g) of row g'
  const int32_t* rowDiag = nullptr;    // [nRowNodes] block index of the diagonal block of the row
  const int32_t* rowLowEnd = nullptr;  // [nRowNodes] end of the leading blocks whose column node is not owned
  const int32_t* mirrorBlk = nullptr;  // [nBlocks] block index of (g', g) for block (g, g') with g' owned
};

// One warp per node-row.  For every element touching the node (ascending element order, as the reference's
// serial loop visits them, ikarus/assembler/simpleassemblers.inl:126-136) the warp streams the N staged blocks
// K_e[la][0..N-1] -- consecutive lanes on consecutive doubles of a block -- and adds them into the node's
// D x (D*nnb) output tile in shared memory.  The positions written for one element are distinct, so there are
// no atomics and no ordering issue inside an element; a __syncwarp orders successive elements, which keeps the
// reference's summation order per entry.  The finished tile (contiguous in Eigen's value order for
// FlatInterleaved dofs) is written once, coalesced, with the Dirichlet mode applied.
// The inner loop is table driven: offTab[la][idx] is the offset of value idx of chunk la inside the
// symmetric-packed staged K_e (transposition folded in), slots come from one byte load per element node.
// RS = scalar rows handled per warp: D for Q1 (the whole node-row), 1 for Q2 (one scalar row per warp: the Hex27
// tile of a whole node-row is 9 KB and would limit the occupancy to 16 warps per SM).
template <int D, int N>
struct GatherCfg {
  static constexpr int RS = (N > 8) ? 1 : D;
  static constexpr int NU = D / RS;            // work units per node-row
  static constexpr int CHUNK = N * RS * D;     // values of one element that go into one unit
  static constexpr int NIT = (CHUNK + 31) / 32;
  static constexpr bool PERSIST = (N * CHUNK > 2048);
};

// offTab[la][idx]: offset of value idx = (lb, ii, k) of chunk la inside the symmetric-packed staged K_e for
// ii = row 0 of the unit, plus bit 14 = "stored directly" (then a further row i0 adds i0*D, else i0)
template <int D, int N>
inline std::vector<int16_t> gatherOffsetTable() {
  using Cfg = GatherCfg<D, N>;
  constexpr int DD = D * D, RS = Cfg::RS, CHUNK = Cfg::CHUNK, HALF = N / 2;
  std::vector<int16_t> tab((size_t)(N * CHUNK + 3) / 4 * 4, 0);  // padded to a multiple of 8 bytes
  for (int t = 0; t < N * CHUNK; ++t) {
    const int la = t / CHUNK, idx = t - la * CHUNK;
    const int lb = idx / (RS * D), q = idx - lb * (RS * D);
    const int i = q / D, k = q - i * D;
    int k1 = lb - la;
    if (k1 < 0) k1 += N;
    const bool direct = (N & 1) ? (k1 <= HALF) : (k1 < HALF || (k1 == HALF && la < HALF));
    const int off = direct ? (k1 * N + la) * DD + i * D + k : ((N - k1) * N + lb) * DD + k * D + i;
    tab[t] = (int16_t)(off | (direct ? 0x4000 : 0));
  }
  return tab;
}

template <int D, int N, int DBC, bool INTERLEAVED>
__global__ void __launch_bounds__(256) gather_kernel(GatherArgs G) {
  using Cfg = GatherCfg<D, N>;
  constexpr int DD = D * D;
  constexpr int RS = Cfg::RS, NU = Cfg::NU, CHUNK = Cfg::CHUNK, NIT = Cfg::NIT;
  constexpr int LAYOUT = INTERLEAVED ? LAYOUT_INTERLEAVED : LAYOUT_LEXICOGRAPHIC;
  extern __shared__ double gsm[];
  // offTab[la][idx]: offset of value idx = (lb, ii, k) of chunk la inside the symmetric-packed staged K_e for
  // ii = row 0 of the unit, plus bit 14 = "stored directly" (then a further row i0 adds i0*D, else i0)
  // (the table depends on (D, N) only: it is computed once on the host, gatherOffsetTable(), and copied per CTA)
  constexpr int TABQ = (N * CHUNK + 3) / 4;  // copied in 8-byte pieces (the device table is padded accordingly)
  __shared__ __align__(8) int16_t offTab[TABQ * 4];
  for (int t = threadIdx.x; t < TABQ; t += blockDim.x)
    reinterpret_cast<int2*>(offTab)[t] = reinterpret_cast<const int2*>(G.offTab)[t];
  __syncthreads();
  const PatternView& P = G.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* out = gsm + (size_t)warp * G.maxOut;
  // persistent CTAs for the large Q2 tables: the offset table is built once per CTA and reused for many units
  constexpr bool PERSIST = Cfg::PERSIST;
  const int64_t warpsTotal = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t nUnits = P.nRowNodes * NU;
  int64_t unit = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (unit >= nUnits) return;
  do {
  __syncwarp();
  const int64_t g = unit / NU;
  const int i0 = (int)(unit - g * NU) * RS;
  const int32_t b0 = P.nbrPtr[g];
  const int nnb = P.nbrPtr[g + 1] - b0;
  const int rowLen = D * nnb;
  // padded row stride of the tile: (i*stride + k) distinct mod 16 for i,k < D -> conflict-free accumulation
  const int rowStride = rowLen + ((D - (rowLen & 15)) + 16) % 16;
  const int32_t a0 = G.adjPtr[g], a1 = G.adjPtr[g + 1];

  if (G.vals) {
    for (int idx = lane; idx < RS * rowStride; idx += 32) out[idx] = 0.0;
    // element-independent lane constants: value idx = lane + 32*it of a chunk belongs to element node lbOf[it]
    int lbOf[NIT], outBase[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = lane + 32 * it;
      const int idc = idx < CHUNK ? idx : 0;
      const int lb = idc / (RS * D);
      const int q = idc - lb * (RS * D);
      const int i = q / D, k = q - i * D;
      lbOf[it] = lb;
      outBase[it] = i * rowStride + k;
    }
    __syncwarp();
    const int elemStride = G.npair * DD;
    // Two-stage software pipeline with explicit ping-pong buffers (no register copies): while element j is added
    // to the tile, the staged values and the slot bytes of element j+1 are already in flight.  The N slot bytes of
    // an element are fetched by lanes 0..N-1 with one load and distributed with shuffles.
    const int16_t* tabLane = offTab + lane;
    const double* __restrict__ Kst = G.Kst;
    const uint8_t* __restrict__ slotTab = G.slotTab;
    const uint32_t* __restrict__ adjCode = G.adjCode;
    auto fetch = [&](int32_t j, double (&v)[NIT], int& slotByte) {
      const uint32_t code = adjCode[j];
      const uint32_t e = code / (uint32_t)N;
      const uint32_t la = code - e * (uint32_t)N;
      const double* Ke = Kst + (size_t)e * (unsigned)elemStride;
      const int16_t* tab = tabLane + la * CHUNK;
      slotByte = lane < N ? (int)slotTab[(size_t)j * N + lane] : 0;
#pragma unroll
      for (int it = 0; it < NIT; ++it)
        if (it < NIT - 1 || lane + 32 * it < CHUNK) {
          const int tv = tab[32 * it];
          const int off = (RS == D) ? (tv & 0x3fff) : (tv & 0x3fff) + ((tv & 0x4000) ? i0 * D : i0);
          v[it] = Ke[off];
        }
    };
    auto accumulate = [&](const double (&v)[NIT], int slotByte) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int sl = __shfl_sync(0xffffffffu, slotByte, lbOf[it]);
        if (it < NIT - 1 || lane + 32 * it < CHUNK) out[outBase[it] + D * sl] += v[it];
      }
      __syncwarp();
    };
    double vA[NIT], vB[NIT];
    int sA = 0, sB = 0;
    int32_t j = a0;
    if (j < a1) fetch(j, vA, sA);
    while (j < a1) {
      if (j + 1 < a1) fetch(j + 1, vB, sB);
      accumulate(vA, sA);
      ++j;
      if (j >= a1) break;
      if (j + 1 < a1) fetch(j + 1, vA, sA);
      accumulate(vB, sB);
      ++j;
    }
    const int64_t gGlobal = g + P.rowBegin;
    bool rowFixed[RS];
    int64_t rowStart[RS];
#pragma unroll
    for (int ii = 0; ii < RS; ++ii) {
      const int i = i0 + ii;
      rowFixed[ii] = (DBC != IKB_DBC_RAW) ? (G.flags[dofOf(LAYOUT, D, P.nNodes, gGlobal, i)] != 0) : false;
      if (INTERLEAVED)
        rowStart[ii] = (int64_t)DD * b0 + (int64_t)i * rowLen;
      else
        rowStart[ii] = (int64_t)i * D * P.nBlocks + (int64_t)D * b0;
    }
    for (int idx2 = lane; idx2 < rowLen; idx2 += 32) {
      const int s = idx2 / D, k = idx2 - s * D;
      const int64_t gb = P.nbrIdx[b0 + s];
      const bool colFixed = (DBC != IKB_DBC_RAW) ? (G.flags[dofOf(LAYOUT, D, P.nNodes, gb, k)] != 0) : false;
      const int64_t offs = INTERLEAVED ? (int64_t)idx2 : (int64_t)k * nnb + s;
#pragma unroll
      for (int ii = 0; ii < RS; ++ii) {
        const int i = i0 + ii;
        double v = out[ii * rowStride + idx2];
        if (DBC == IKB_DBC_REDUCED) {
          if (!rowFixed[ii] && !colFixed)
            G.vals[G.redRowStart[localRowOf(P, g, i)] + reducedRank(P, G.flags, G.freeCnt, G.freeTot, g, b0 + s, gb, k)] = v;
        } else {
          // Full: zero constrained rows and columns, unit diagonal (simpleassemblers.inl:159-167)
          if (DBC == IKB_DBC_FULL && (rowFixed[ii] || colFixed)) v = (gb == gGlobal && i == k) ? 1.0 : 0.0;
          G.vals[rowStart[ii] + offs] = v;
        }
      }
    }
  }

    unit += warpsTotal;
  } while (PERSIST && unit < nUnits);
}

// Contribution-list gather ("pull"): one warp per node-row, one lane per matrix entry of a group of 32/(D*D) pattern
// blocks.  The stable sort of the pattern build leaves, for every pattern block, the list of staged K_e blocks that
// add into it in ascending element order (cptr/csrc: code = e*npair + p, bit 31 = read transposed) -- the order of
// the reference's serial element loop (ikarus/assembler/simpleassemblers.inl:126-136).  The D*D lanes of a block read
// one 8*D*D-byte staged block per contribution and add it in a register; no shared-memory read-modify-write, no
// atomics, values bit-identical to gather_kernel.  The codes of a chunk of blocks are staged in shared memory with
// coalesced loads so that the dependent chain per node-row is cptr -> codes -> values (three global latencies), and
// up to four contributions of a block are in flight per lane.
constexpr int PULL_CAP = 256;  // staged codes per warp
constexpr int PULL_WARPS_MAX = 4;  // warps per CTA (sizes the staging buffer: 4 KB per CTA, 64 KB per SM at 16 CTAs)

// staged values: read-only for the lifetime of the gather kernel -> non-coherent path; inside the pipelined sweep the
// staging ring is rewritten while the kernel runs -> L2 (ld.global.cg), the coherence point with the producing kernel
template <bool RING>
__device__ __forceinline__ double pullLoad(uint64_t a) {
  double v;
  if (RING)
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(a) : "memory");
  else
    asm("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(a));
  return v;
}
__device__ __forceinline__ void pullStore(uint64_t a, double v) {
  asm volatile("st.global.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory");
}

// one node-row of the pull gather; sm: PULL_CAP words of shared memory private to the warp
template <int D, int DBC, bool INTERLEAVED, bool IDX32, bool RING, bool MIRROR = false>
__device__ __forceinline__ void pullRow(const GatherArgs& G, const int32_t* __restrict__ cptr,
                                        const uint32_t* __restrict__ csrc, int64_t g, uint32_t* sm, int lane) {
  constexpr int DD = D * D;
  constexpr int BPW = 32 / DD;               // pattern blocks per warp pass
  constexpr int CH = (31 / BPW) * BPW;       // blocks per chunk (their cptr values + 1 fit one warp load)
  constexpr int LAYOUT = INTERLEAVED ? LAYOUT_INTERLEAVED : LAYOUT_LEXICOGRAPHIC;
  constexpr unsigned FULLMASK = 0xffffffffu;
  const PatternView& P = G.P;
  const int q = lane / DD, v = lane - q * DD, i = v / D, k = v - i * D;
  const bool active = q < BPW;
  const int32_t b0 = P.nbrPtr[g], b1 = P.nbrPtr[g + 1];
  const int nnb = b1 - b0;
  const int64_t gGlobal = g + P.rowBegin;
  const bool rowFixed = (DBC != IKB_DBC_RAW) ? (G.flags[dofOf(LAYOUT, D, P.nNodes, gGlobal, i)] != 0) : false;
  const bool anyRowFixed = (DBC != IKB_DBC_RAW) ? __any_sync(FULLMASK, rowFixed) : false;
  // Staged codes: w = 2*DD*(e*npair+p) + transposed = offset of the staged block in 4-byte units, flag in bit 0.  The
  // value of this lane sits at 4-byte unit w + 2*(i*D+k) (direct) or w - 1 + 2*(k*D+i) (transposed), i.e. at
  // laneBase + 4*(w + t*delta) with laneBase = Kst + 8*(i*D+k) and delta = 2*((k*D+i) - (i*D+k)) - 1: three integer
  // instructions per value.  IDX32 (checked by the host): w < 2^31.
  // (lane constants pass through an identity shuffle: ptxas would otherwise recompute them from threadIdx for every
  // group instead of keeping them in registers)
  uint64_t laneBase = reinterpret_cast<uint64_t>(G.Kst) + 8 * (i * D + k);
  int32_t delta = 2 * ((k * D + i) - (i * D + k)) - 1;
  laneBase = __shfl_sync(FULLMASK, (unsigned long long)laneBase, lane);
  delta = __shfl_sync(FULLMASK, delta, lane);
  auto stagedValue = [&](uint32_t w) -> double {
    if (IDX32) {
      uint64_t a;  // (spelled in PTX: the compiler turns t*delta into a compare and a select)
      asm("{\n\t.reg .b32 t, x;\n\tand.b32 t, %1, 1;\n\tmad.lo.s32 x, t, %2, %1;\n\tmad.wide.s32 %0, x, 4, %3;\n\t}"
          : "=l"(a)
          : "r"(w), "r"(delta), "l"(laneBase));
      return pullLoad<RING>(a);
    }
    // 64-bit offsets (staged codes are 2*(e*npair+p) + transposed here): laneBase + w*4*DD + t*(4*(delta+1) - 4*DD)
    uint64_t a;
    asm("{\n\t.reg .b32 t;\n\t.reg .b64 o;\n\tand.b32 t, %1, 1;\n\tmul.wide.s32 o, t, %2;\n\t"
        "mad.wide.u32 %0, %1, %3, %4;\n\tadd.s64 %0, %0, o;\n\t}"
        : "=l"(a)
        : "r"(w), "r"(4 * (delta + 1) - 4 * DD), "n"(4 * DD), "l"(laneBase));
    return pullLoad<RING>(a);
  };
  uint64_t dst = reinterpret_cast<uint64_t>(
      (DBC == IKB_DBC_REDUCED) ? G.vals
                               : G.vals + (INTERLEAVED ? (int64_t)DD * b0 + (int64_t)i * D * nnb + k
                                                       : (int64_t)i * D * P.nBlocks + (int64_t)D * b0 + (int64_t)k * nnb));
  dst = __shfl_sync(FULLMASK, (unsigned long long)dst, lane);
  constexpr int SSTRIDE = INTERLEAVED ? D : 1;  // doubles between the entries (i,k) of consecutive slots
  int64_t redStart = 0;
  if (DBC == IKB_DBC_REDUCED) redStart = G.redRowStart[localRowOf(P, g, i)];

  // K is symmetric and so is every staged K_e (the packed form keeps one of (a,b), (b,a)), so block (g', g) is bit for
  // bit the transpose of block (g, g'): with MIRROR only pass 0 (column node owned by another rank) and pass 1 (from the
  // diagonal block on) are gathered, halving the staged reads; the blocks in between arrive from their mirror rows.
  int32_t lowEnd = b1, diag = b1;
  if (MIRROR) {
    lowEnd = G.rowLowEnd[g];
    diag = G.rowDiag[g];
  }
#pragma unroll 1
  for (int pass = 0; pass < (MIRROR ? 2 : 1); ++pass) {
  const int32_t r0 = (MIRROR && pass == 1) ? diag : b0;
  const int32_t r1 = (MIRROR && pass == 0) ? lowEnd : b1;
  for (int32_t cb = r0; cb < r1; cb += CH) {
    const int nb = min(CH, r1 - cb);
    const int32_t cp = cptr[cb + min(lane, nb)];
    // where the transposed copy of block (cb + lane) goes: first entry of block (g', g) in row g', length of that row
    int64_t mBase = -1;
    int nnbM = 0;
    if (MIRROR && pass == 1 && lane < nb) {
      const int64_t gl = (int64_t)P.nbrIdx[cb + lane] - P.rowBegin;
      if (gl != g && gl >= 0 && gl < P.nRowNodes) {
        const int32_t mb = G.mirrorBlk[cb + lane];
        const int32_t p0 = P.nbrPtr[gl];
        nnbM = P.nbrPtr[gl + 1] - p0;
        mBase = INTERLEAVED ? (int64_t)DD * p0 + (int64_t)D * (mb - p0) : (int64_t)D * p0 + (mb - p0);
      }
    }
    // Dirichlet work only where the row or one of the chunk's column nodes is constrained (warp-uniform)
    bool slow = anyRowFixed;
    if (DBC != IKB_DBC_RAW && !slow) {
      bool f = false;
      if (lane < nb) {
        const int64_t gb = P.nbrIdx[cb + lane];
#pragma unroll
        for (int c = 0; c < D; ++c) f |= G.flags[dofOf(LAYOUT, D, P.nNodes, gb, c)] != 0;
      }
      slow = __any_sync(FULLMASK, f);
    }
    const int32_t cbase = __shfl_sync(FULLMASK, cp, 0);
    const int32_t ncodes = __shfl_sync(FULLMASK, cp, nb) - cbase;
    const bool staged = ncodes <= G.pullStageMax;
    __syncwarp();
    if (staged) {
      for (int t = lane; t < ncodes; t += 32) {
        const uint32_t c = csrc[cbase + t];
        sm[t] = IDX32 ? (c & SRC_MASK) * (uint32_t)(2 * DD) + (c >> 31) : __funnelshift_l(c, c, 1);
      }
      __syncwarp();
    }
    // per block of the chunk: (offset of its first code in the staged list) | (number of codes) << 16
    const int32_t cnext = __shfl_down_sync(FULLMASK, cp, 1);
    const uint32_t pack = (lane < nb) ? (uint32_t)(cp - cbase) | ((uint32_t)(cnext - cp) << 16) : 0u;
    const int slotBase = (int)(cb - b0);
    // writes the finished entry of pattern block (slot sg of the row) with the Dirichlet mode applied
    auto emit = [&](int sg, double val, int64_t mB, int nM) {
      if (DBC == IKB_DBC_RAW || !slow) {
        if (DBC == IKB_DBC_REDUCED) {
          const int32_t b = b0 + sg;
          G.vals[redStart + reducedRank(P, G.flags, G.freeCnt, G.freeTot, g, b, P.nbrIdx[b], k)] = val;
        } else {
          pullStore(dst + (uint64_t)(8 * SSTRIDE) * (uint32_t)sg, val);
          if (MIRROR && mB >= 0)  // entry (i,k) of (g,g') = entry (k,i) of (g',g)
            G.vals[mB + (INTERLEAVED ? (int64_t)k * D * nM + i : (int64_t)k * D * P.nBlocks + (int64_t)i * nM)] = val;
        }
      } else {
        const int32_t b = b0 + sg;
        const int64_t gb = P.nbrIdx[b];
        const bool colFixed = G.flags[dofOf(LAYOUT, D, P.nNodes, gb, k)] != 0;
        if (DBC == IKB_DBC_REDUCED) {
          if (!rowFixed && !colFixed)
            G.vals[redStart + reducedRank(P, G.flags, G.freeCnt, G.freeTot, g, b, gb, k)] = val;
        } else {
          // Full: zero constrained rows and columns, unit diagonal (simpleassemblers.inl:159-167)
          if (rowFixed || colFixed) val = (gb == gGlobal && i == k) ? 1.0 : 0.0;
          pullStore(dst + (uint64_t)(8 * SSTRIDE) * (uint32_t)sg, val);
          if (MIRROR && mB >= 0)  // the mirrored entry has row and column flags swapped: same rule, same value
            G.vals[mB + (INTERLEAVED ? (int64_t)k * D * nM + i : (int64_t)k * D * P.nBlocks + (int64_t)i * nM)] = val;
        }
      }
    };
    if (staged) {
      // One group = BPW blocks, one matrix entry per lane.  fetch() issues the loads of the first four contributions,
      // finish() adds them (plus any further batches of four) and writes the entry.
      struct Grp {
        double x[4];
        int n, s;
        uint32_t rel;
      };
      auto fetch = [&](int r, Grp& grp) {
        grp.s = r + q;  // block of this lane inside the chunk
        const uint32_t pk = __shfl_sync(FULLMASK, pack, grp.s);
        grp.n = (active && grp.s < nb) ? (int)(pk >> 16) : 0;
        grp.rel = pk & 0xffffu;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = (j < grp.n) ? sm[grp.rel + j] : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          grp.x[j] = 0.0;
          if (j < grp.n) grp.x[j] = stagedValue(w[j]);
        }
      };
      auto finish = [&](Grp& grp) {
        // lanes past their count add +0.0 (the sum starts from +0.0, so this changes no bit)
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += grp.x[j];
        const int nmax = __reduce_max_sync(FULLMASK, grp.n);
        for (int u = 4; u < nmax; u += 4) {
          uint32_t w[4];
          double x[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) w[j] = (u + j < grp.n) ? sm[grp.rel + u + j] : 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            x[j] = 0.0;
            if (u + j < grp.n) x[j] = stagedValue(w[j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) acc += x[j];
        }
        int64_t mB = -1;
        int nM = 0;
        if (MIRROR && pass == 1) {  // warp-uniform
          const int sl = grp.s < nb ? grp.s : 0;
          mB = __shfl_sync(FULLMASK, (long long)mBase, sl);
          nM = __shfl_sync(FULLMASK, nnbM, sl);
        }
        if (active && grp.s < nb) emit(slotBase + grp.s, acc, mB, nM);
      };
      // (issuing the loads of the next group before the adds of the current one was measured: the extra registers cost
      // more occupancy than the overlap gains -- 64 resident warps per SM hide the latency better)
      Grp grp;
      for (int r = 0; r < nb; r += BPW) {
        fetch(r, grp);
        finish(grp);
      }
    } else {
      // contribution list too long for the staging buffer: read the codes from global memory
      for (int r = 0; r < nb; r += BPW) {
        const int sl = r + q;
        const bool valid = active && sl < nb;
        const int32_t c0 = valid ? cptr[cb + sl] : 0;
        const int32_t c1 = valid ? cptr[cb + sl + 1] : 0;
        double acc = 0.0;
        for (int32_t c = c0; c < c1; ++c) {
          const uint32_t cc = csrc[c];
          acc += stagedValue(IDX32 ? (cc & SRC_MASK) * (uint32_t)(2 * DD) + (cc >> 31) : __funnelshift_l(cc, cc, 1));
        }
        int64_t mB = -1;
        int nM = 0;
        if (MIRROR && pass == 1) {
          const int sl2 = sl < nb ? sl : 0;
          mB = __shfl_sync(FULLMASK, (long long)mBase, sl2);
          nM = __shfl_sync(FULLMASK, nnbM, sl2);
        }
        if (valid) emit(slotBase + sl, acc, mB, nM);
      }
    }
  }
  }
}

// 32 registers per thread (16 CTAs of 4 warps per SM): the kernel is latency bound, 64 resident warps per SM measured
// 13 % faster than the 48 warps the unconstrained 40-register build reaches; 4 warps per CTA 2 % faster than 8.
template <int D, int DBC, bool INTERLEAVED, bool IDX32, bool MIRROR>
__global__ void __launch_bounds__(32 * PULL_WARPS_MAX, MIRROR ? 12 : 16)
    gather_pull_kernel(GatherArgs G, const int32_t* __restrict__ cptr, const uint32_t* __restrict__ csrc) {
  __shared__ uint32_t codeBuf[PULL_WARPS_MAX][PULL_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t g = G.rowFirst + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (g >= (G.rowEnd < 0 ? G.P.nRowNodes : G.rowEnd)) return;
  pullRow<D, DBC, INTERLEAVED, IDX32, false, MIRROR>(G, cptr, csrc, g, codeBuf[warp], lane);
}

// highest element touching each node-row (the adjacency lists ascend by element); -1 for rows without elements
__global__ void row_max_elem_kernel(int64_t nRowNodes, const int32_t* __restrict__ adjPtr, const uint32_t* __restrict__ adjCode,
                                    int n, int32_t* maxElem) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nRowNodes) return;
  const int32_t a0 = adjPtr[g], a1 = adjPtr[g + 1];
  maxElem[g] = a1 > a0 ? (int32_t)(adjCode[a1 - 1] / (uint32_t)n) : -1;
}

// one-time maps of the mirrored pull
__global__ void mirror_rows_kernel(PatternView P, int32_t* rowDiag, int32_t* rowLowEnd) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.nRowNodes) return;
  const int32_t b0 = P.nbrPtr[g], b1 = P.nbrPtr[g + 1];
  // first block whose column node is >= rowBegin (the lists ascend), and the diagonal block
  int32_t lo = b0, hi = b1;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (P.nbrIdx[mid] < P.rowBegin)
      lo = mid + 1;
    else
      hi = mid;
  }
  rowLowEnd[g] = lo;
  rowDiag[g] = diagBlockOf(P, g);
}
__global__ void mirror_blocks_kernel(PatternView P, int32_t* mirrorBlk) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.nBlocks) return;
  const int64_t gl = (int64_t)P.nbrIdx[b] - P.rowBegin;  // row of the mirrored block
  int32_t out = -1;
  if (gl >= 0 && gl < P.nRowNodes) {
    const int32_t target = (int32_t)(P.nbrRow[b] + P.rowBegin);
    int32_t lo = P.nbrPtr[gl], hi = P.nbrPtr[gl + 1];
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if (P.nbrIdx[mid] < target)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo < P.nbrPtr[gl + 1] && P.nbrIdx[lo] == target) out = lo;
  }
  mirrorBlk[b] = out;
}

// Residual gather: one thread per (node-row, component).  The node's (element, local node) adjacency is walked in
// ascending element order (VectorFlatAssembler::get*VectorImpl, ikarus/assembler/simpleassemblers.inl:59-118);
// external load and Dirichlet mode are applied on the fly.  A separate, tiny kernel so that the host can fetch R
// while the matrix gather is still running.
template <int D, int N, int DBC, bool INTERLEAVED>
__global__ void __launch_bounds__(256) gather_vec_kernel(GatherArgs G) {
  constexpr int LAYOUT = INTERLEAVED ? LAYOUT_INTERLEAVED : LAYOUT_LEXICOGRAPHIC;
  const PatternView& P = G.P;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nRowNodes * D) return;
  const int64_t g = t / D;
  const int i = (int)(t - g * D);
  const int32_t a0 = G.adjPtr[g], a1 = G.adjPtr[g + 1];
  double r = 0.0;
  for (int32_t j = a0; j < a1; ++j) {
    const uint32_t code = G.adjCode[j];
    const uint32_t e = code / (uint32_t)N;
    const int la = (int)(code - e * (uint32_t)N);
    r += G.Rst[(size_t)e * (N * D) + la * D + i];
  }
  const int64_t rowDof = dofOf(LAYOUT, D, P.nNodes, g + P.rowBegin, i);
  const bool rowFixed = (DBC != IKB_DBC_RAW) ? (G.flags[rowDof] != 0) : false;
  if (G.fext) r -= G.fextScale * G.fext[rowDof];
  if (DBC == IKB_DBC_REDUCED) {
    if (!rowFixed) G.vec[rowDof - G.cbelow[rowDof] - G.redVecOffset] = r;
  } else {
    if (DBC == IKB_DBC_FULL && rowFixed) r = 0.0;  // simpleassemblers.inl:90-92
    G.vec[localRowOf(P, g, i)] = r;
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
