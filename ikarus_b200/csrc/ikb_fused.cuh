// Pipelined K+R(+E) sweep for Hex8 meshes: the element kernel and the row assembly run AT THE SAME TIME as two
// persistent kernels, and the element matrices go from one to the other through an L2-resident ring instead of HBM.
//
// Replaces the element loop and the scatter of SparseFlatAssembler::assembleRawMatrixImpl / getMatrixImpl /
// getReducedMatrixImpl and VectorFlatAssembler::get*VectorImpl (ikarus/assembler/simpleassemblers.inl:59-204).  The
// back-to-back two-kernel path (all K_e staged in HBM, then gathered) moved 3.7x the algorithmic bytes and ran each
// kernel against its own limiter with the other's resources idle.
//
//   producer  sweep_elem_kernel   8 warps per SM, 128 registers.  Tickets of 4 consecutive elements in element order:
//             Gauss-point phase + DMMA contraction (h8_warp_elements); symmetric-packed K_e and R_e go to ring slot
//             e mod ringElems.  Completed tickets are counted per group of 64.
//   consumer  sweep_rows_kernel   32 warps per SM, 32 registers.  Tickets of 8 consecutive node-rows in row order: wait
//             until the element groups covering the rows' adjacent elements are complete, then the pull gather of
//             ikb_gather.cuh (contribution lists in ascending element order -- the order of the reference's serial
//             element loop, simpleassemblers.inl:126-136 -- one lane per matrix entry, register accumulation, no float
//             atomics) and the residual rows, with the Dirichlet mode applied.  Completed tickets are counted per group.
//
// The two kernels have different limiters (FP64/DMMA pipe vs. L1/LSU wavefronts and latency) and need different
// register budgets (128 x few warps vs. 32 x many warps), which is why they are two kernels sharing the SMs rather
// than one.  Both grids are sized so that every CTA of both is resident (2 + 8 CTAs of 128 threads per SM: 32 K + 32 K
// registers), so a waiting warp never keeps the warp it waits for off the machine.
//
// Ring: ringElems slots of 36 blocks x 72 B (+ 24 doubles of R_e).  For a mesh whose element numbering has a bounded
// front (structured or bandwidth-reduced numbering) the ring is a few tens of MB and stays in the 126 MB L2: neither the
// write of K_e nor its two reads reach HBM.  A mesh with an unbounded front gets one slot per element (plain staging,
// still overlapped).  Write-after-read guard: before a producer ticket overwrites a slot, the row tickets that read
// the slot's previous element must be complete; those only depend on OLDER producer tickets (checked when the ring
// size is chosen), so the oldest unfinished ticket of either kernel can always run.
// Counters only grow; an epoch per launch makes the targets, nothing is reset between launches.
#pragma once
#include <algorithm>
#include <vector>

#include "ikb_elem_h8mma.cuh"
#include "ikb_gather.cuh"

namespace ikb {

constexpr int SWEEP_RB = 8;     // node-rows per consumer ticket
constexpr int SWEEP_EG = 64;    // producer tickets per completion group
constexpr int SWEEP_RG = 32;    // consumer tickets per completion group
constexpr unsigned SWEEP_SPIN_LIMIT = 1u << 14;  // polls (0.5 .. 16 us apart) before a wait gives up (~0.25 s)

struct SweepCtl {                       // device-resident control block
  unsigned long long elemTicket;        // next producer ticket (monotone over launches)
  unsigned long long rowTicket;         // next consumer ticket
  unsigned long long rowWatermark;      // (epoch+1) << 32 | leading consumer groups known complete
  unsigned long long pad;
  unsigned long long t[4];              // diagnostics: first start / last end of producer and consumer warps (globaltimer, ns)
};

__device__ __forceinline__ unsigned long long globalTimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct SweepArgs {
  // producer
  ElemArgs E;
  H8Out O;
  const uint32_t* guard;       // [nElemTickets] consumer groups [0, guard[T]) must be complete before ticket T writes (null: no reuse)
  // consumer
  GatherArgs G;
  const int32_t* cptr;         // [nBlocks+1]
  const uint32_t* csrcRing;    // [nContrib] staged-block codes with the element taken modulo ringElems
  const uint32_t* adjRing;     // [nAdj] (e mod ringElems)*8 + la per (node-row, adjacent element)
  const uint2* rowWait;        // [nRowTickets] first / last producer group the ticket's rows depend on (first > last: none)
  // shared
  SweepCtl* ctl;
  unsigned* elemDone;          // [nElemGroups] completed producer tickets, cumulative over launches
  unsigned* rowDone;           // [nRowGroups]
  unsigned nElemTickets, nRowTickets;
  unsigned long long elemTicketBase, rowTicketBase;
  unsigned epoch;
  unsigned what;
  int timing;                  // IKB_SWEEP_DEBUG: record start/end times of the two kernels
};

__device__ __forceinline__ unsigned ldRelaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// waits until done[gi] has reached (epoch+1)*size(gi) for every group gi in [lo, hi]; false on timeout.
// The polls are relaxed L2 loads: what is ordered behind them is either my writes (guard) or reads that go to L2
// themselves (ld.global.cg in the gather), so no L1 invalidation is needed.
__device__ __forceinline__ bool sweepWait(const unsigned* done, unsigned lo, unsigned hi, unsigned nItems, int groupSize,
                                          unsigned epoch, int lane) {
  for (unsigned base = lo; base <= hi; base += 32) {
    const unsigned gi = base + lane;
    const bool mine = gi <= hi;
    unsigned target = 0;
    if (mine) {
      const unsigned first = gi * (unsigned)groupSize;
      const unsigned sz = min((unsigned)groupSize, nItems - first);
      target = (epoch + 1u) * sz;
    }
    unsigned spins = 0;
    for (;;) {
      bool ok = true;
      if (mine) ok = (int)(ldRelaxed(done + gi) - target) >= 0;
      if (__all_sync(0xffffffffu, ok)) break;
      if (++spins > SWEEP_SPIN_LIMIT) return false;
      // back off: thousands of waiting warps polling one counter would keep the completion signals queueing behind them
      __nanosleep(spins < 4 ? 500u : (spins < 12 ? 4000u : 16000u));
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------- producer
template <int FORM>
__global__ void __launch_bounds__(32 * H8Cfg::WARPS, 4) sweep_elem_kernel(SweepArgs S) {  // 128 registers
  extern __shared__ double smem[];
  constexpr unsigned FULLMASK = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wsm = smem + (size_t)warp * H8Cfg::WARP_DOUBLES;
  if (S.timing && lane == 0) atomicMin(&S.ctl->t[0], globalTimer());
  for (;;) {
    unsigned long long tk = 0;
    if (lane == 0) tk = atomicAdd(&S.ctl->elemTicket, 1ull);
    tk = __shfl_sync(FULLMASK, tk, 0) - S.elemTicketBase;
    if (tk >= S.nElemTickets) break;
    if (S.guard) {
      const unsigned need = S.guard[tk];
      if (need) {
        unsigned long long wm = 0;
        if (lane == 0) wm = *reinterpret_cast<volatile unsigned long long*>(&S.ctl->rowWatermark);
        wm = __shfl_sync(FULLMASK, wm, 0);
        const unsigned known = ((unsigned)(wm >> 32) == S.epoch + 1u) ? (unsigned)wm : 0u;
        if (known < need) {
          if (!sweepWait(S.rowDone, known, need - 1, S.nRowTickets, SWEEP_RG, S.epoch, lane)) atomicMin(S.E.errFlag, -2);
          if (lane == 0) atomicMax(&S.ctl->rowWatermark, ((unsigned long long)(S.epoch + 1u) << 32) | need);
        }
        __syncwarp();
      }
    }
    h8_warp_elements<FORM>(S.E, S.O, (int64_t)tk * 4, wsm);
    __threadfence();  // K_e / R_e of the four elements are visible before the ticket counts as complete
    __syncwarp();
    if (lane == 0) atomicAdd(S.elemDone + tk / SWEEP_EG, 1u);
  }
  if (S.timing && lane == 0) atomicMax(&S.ctl->t[1], globalTimer());
}

// ---------------------------------------------------------------------------------------------------- consumer
// MINB = 16: 32 registers (8 resident CTAs next to the producer's two), MINB = 12: 40 registers (6 CTAs)
template <int DBC, bool INTERLEAVED, int MINB>
__global__ void __launch_bounds__(32 * PULL_WARPS_MAX, MINB) sweep_rows_kernel(SweepArgs S) {
  constexpr int D = 3, N = 8;
  constexpr int LAYOUT = INTERLEAVED ? LAYOUT_INTERLEAVED : LAYOUT_LEXICOGRAPHIC;
  constexpr unsigned FULLMASK = 0xffffffffu;
  __shared__ uint32_t codeBuf[PULL_WARPS_MAX][PULL_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatherArgs& G = S.G;
  const PatternView& P = G.P;
  if (S.timing && lane == 0) atomicMin(&S.ctl->t[2], globalTimer());
  for (;;) {
    unsigned long long tk = 0;
    if (lane == 0) tk = atomicAdd(&S.ctl->rowTicket, 1ull);
    tk = __shfl_sync(FULLMASK, tk, 0) - S.rowTicketBase;
    if (tk >= S.nRowTickets) break;
    const uint2 w = S.rowWait[tk];
    if (w.x <= w.y) {
      if (!sweepWait(S.elemDone, w.x, w.y, S.nElemTickets, SWEEP_EG, S.epoch, lane)) atomicMin(S.E.errFlag, -2);
      __syncwarp();
    }
    const int64_t g0 = (int64_t)tk * SWEEP_RB;
    const int64_t g1 = min(g0 + (int64_t)SWEEP_RB, P.nRowNodes);
    if (S.what & IKB_VECTOR) {
      // residual rows: lane = (row of the ticket, component); ascending element order (simpleassemblers.inl:59-118)
      const int64_t g = g0 + lane / D;
      const int i = lane - (lane / D) * D;
      if (lane < SWEEP_RB * D && g < g1) {
        const int32_t a0 = G.adjPtr[g], a1 = G.adjPtr[g + 1];
        double r = 0.0;
        for (int32_t j = a0; j < a1; ++j) {
          const uint32_t code = S.adjRing[j];
          r += __ldcg(G.Rst + (size_t)(code >> 3) * (N * D) + (code & 7u) * D + i);
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
    }
    if (S.what & IKB_MATRIX) {
      for (int64_t g = g0; g < g1; ++g) pullRow<D, DBC, INTERLEAVED, true, true>(G, S.cptr, S.csrcRing, g, codeBuf[warp], lane);
    }
    // every lane's ring reads have returned (its stores depend on them) before the ticket counts as complete
    __syncwarp();
    if (lane == 0) atomicAdd(S.rowDone + tk / SWEEP_RG, 1u);
  }
  if (S.timing && lane == 0) atomicMax(&S.ctl->t[3], globalTimer());
}

// ---------------------------------------------------------------------------------------------------- one-time maps
// staged-block codes and adjacency codes with the element index taken modulo the ring size
__global__ void sweep_ring_codes_kernel(const uint32_t* __restrict__ csrc, int64_t nContrib, int npair, uint32_t ringElems,
                                        uint32_t* csrcRing) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nContrib) return;
  const uint32_t s = csrc[c];
  const uint32_t code = s & SRC_MASK;
  const uint32_t e = code / (uint32_t)npair;
  const uint32_t p = code - e * (uint32_t)npair;
  csrcRing[c] = ((e % ringElems) * (uint32_t)npair + p) | (s & SRC_TRANSPOSE);
}
__global__ void sweep_ring_adj_kernel(const uint32_t* __restrict__ adjCode, int64_t nAdj, uint32_t ringElems, uint32_t* adjRing) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nAdj) return;
  const uint32_t code = adjCode[j];
  adjRing[j] = ((code >> 3) % ringElems) * 8u + (code & 7u);
}

// per consumer ticket: first and last producer ticket among the elements adjacent to its rows
__global__ void sweep_row_range_kernel(int64_t nRowNodes, const int32_t* __restrict__ adjPtr,
                                       const uint32_t* __restrict__ adjCode, int32_t* lo, int32_t* hi) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t g0 = t * SWEEP_RB;
  if (g0 >= nRowNodes) return;
  const int64_t g1 = min(g0 + (int64_t)SWEEP_RB, nRowNodes);
  int32_t mn = INT32_MAX, mx = -1;
  for (int64_t g = g0; g < g1; ++g) {
    const int32_t a0 = adjPtr[g], a1 = adjPtr[g + 1];
    if (a1 > a0) {
      mn = min(mn, (int32_t)(adjCode[a0] >> 5));  // element / 4 (adjacency lists ascend in the element index)
      mx = max(mx, (int32_t)(adjCode[a1 - 1] >> 5));
    }
  }
  lo[t] = mn;
  hi[t] = mx;
}

// per producer ticket: the highest owned node-row any of its elements touches (-1: none)
__global__ void sweep_ticket_rows_kernel(const int32_t* __restrict__ elemNode, int64_t nElem, int64_t rowBegin, int64_t rowEnd,
                                         int32_t* maxRow) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t * 4 >= nElem) return;
  int32_t mx = -1;
  for (int64_t e = t * 4; e < min(t * 4 + 4, nElem); ++e)
    for (int a = 0; a < 8; ++a) {
      const int64_t g = (int64_t)elemNode[(size_t)a * nElem + e] - rowBegin;
      if (g >= 0 && g < rowEnd - rowBegin) mx = max(mx, (int32_t)g);
    }
  maxRow[t] = mx;
}

// Host: write-after-read guards for a ring of ringElems (a multiple of 4) slots.  Producer ticket T overwrites the slots
// of ticket T - ringElems/4, whose elements are read by the consumer tickets up to the one holding its highest row:
// those tickets -- rounded up to whole completion groups -- must be complete first.  Returns false when one of them
// depends on a producer ticket that is not at least `margin` tickets older than T (margin 0 = bare deadlock freedom).
inline bool sweepBuildGuards(int64_t nElemTickets, int64_t nRowTickets, const std::vector<int32_t>& rowHi,
                             const std::vector<int32_t>& ticketMaxRow, int64_t ringElems, int64_t margin,
                             std::vector<uint32_t>& guard) {
  guard.assign((size_t)nElemTickets, 0u);
  const int64_t back = ringElems / 4;
  // newest producer ticket the consumer groups [0, k] depend on
  const int64_t nGroups = (nRowTickets + SWEEP_RG - 1) / SWEEP_RG;
  std::vector<int32_t> pre((size_t)nGroups);
  int32_t run = -1;
  for (int64_t k = 0; k < nGroups; ++k) {
    for (int64_t t = k * SWEEP_RG; t < std::min<int64_t>(nRowTickets, (k + 1) * SWEEP_RG); ++t)
      if (rowHi[(size_t)t] >= 0)  // the consumer waits for the whole producer GROUP holding its last ticket
        run = std::max<int32_t>(run, (int32_t)std::min<int64_t>(nElemTickets - 1, ((int64_t)rowHi[(size_t)t] / SWEEP_EG + 1) * SWEEP_EG - 1));
    pre[(size_t)k] = run;
  }
  int32_t runRow = -1;  // highest row touched by the tickets up to T - back
  for (int64_t T = back; T < nElemTickets; ++T) {
    runRow = std::max(runRow, ticketMaxRow[(size_t)(T - back)]);
    if (runRow < 0) continue;
    const int64_t needTickets = (int64_t)runRow / SWEEP_RB + 1;
    const int64_t needGroups = (needTickets + SWEEP_RG - 1) / SWEEP_RG;
    if ((int64_t)pre[(size_t)needGroups - 1] >= T - margin) return false;
    guard[(size_t)T] = (uint32_t)needGroups;
  }
  return true;
}

}  // namespace ikb
