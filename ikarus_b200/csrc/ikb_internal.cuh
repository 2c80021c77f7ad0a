// Internal state of a device flat assembler handle (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/ikb200.h"

namespace ikb {

// FORM_PS: the principal-stretch hyperelastic laws (ikb_material_ps.cuh), served by the generalised-tangent kernel
enum Form { FORM_LE = 0, FORM_SVK = 1, FORM_NH = 2, FORM_PS = 3 };
enum Layout { LAYOUT_INTERLEAVED = 0, LAYOUT_LEXICOGRAPHIC = 1 };

// The law, as ikb_set_hyperelastic receives it (include/ikb200.h: ikb_hyperelastic)
struct PsLaw {
  int dev = 1;  // IKB_DEV_*: 0 none, 1 BlatzKo, 2 Ogden (total stretches), 3 Ogden (deviatoric), 4 InvariantBased, 5 ArrudaBoyce, 6 Gent
  int n = 0;    // terms of Ogden / InvariantBased (<= 3)
  int vf = 0;   // VF0 .. VF12
  int pex[3] = {0, 0, 0}, qex[3] = {0, 0, 0};
  double par[3] = {0, 0, 0};  // Ogden: mu_i; InvariantBased: c_i; BlatzKo: mu; ArrudaBoyce: mu, lambdaM; Gent: mu, Jm
  double ex[3] = {0, 0, 0};   // Ogden: alpha_i
  double K = 0.0, beta = 0.0; // penalty parameter and beta (VF4, VF7, VF10) of the volumetric function
};


constexpr uint32_t SRC_TRANSPOSE = 0x80000000u;
constexpr uint32_t SRC_MASK = 0x7fffffffu;
// doubles per staged K_e block
__host__ __device__ constexpr int blockStride(int dim) { return dim * dim; }

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// Node-block view of the sparsity pattern.  Scalar CSR positions are analytic in it:
//   interleaved   : row d*ga+i starts at d*d*nbrPtr[ga] + i*d*nnb(ga); entry (slot,k) at +d*slot+k
//   lexicographic : row i*nNodes+ga starts at i*d*nBlocks + d*nbrPtr[ga]; entry (k,slot) at +k*nnb(ga)+slot
struct PatternView {
  int dim;
  int layout;
  int64_t nNodes;     // global node count (columns)
  int64_t rowBegin;   // first owned node-row
  int64_t nRowNodes;  // owned node-rows
  int64_t nBlocks;
  const int32_t* nbrPtr;  // [nRowNodes+1]
  const int32_t* nbrIdx;  // [nBlocks] global neighbour node, sorted per row
  const int32_t* nbrRow;  // [nBlocks] local row-node of the block
};

__host__ __device__ inline int64_t rawRowStart(const PatternView& P, int64_t gaLocal, int i, int nnb) {
  const int d = P.dim;
  if (P.layout == LAYOUT_INTERLEAVED) return (int64_t)d * d * P.nbrPtr[gaLocal] + (int64_t)i * d * nnb;
  return (int64_t)i * d * P.nBlocks + (int64_t)d * P.nbrPtr[gaLocal];
}
__host__ __device__ inline int64_t rawEntryOffset(const PatternView& P, int slot, int k, int nnb) {
  if (P.layout == LAYOUT_INTERLEAVED) return (int64_t)P.dim * slot + k;
  return (int64_t)k * nnb + slot;
}
__host__ __device__ inline int64_t dofOf(int layout, int dim, int64_t nNodes, int64_t node, int c) {
  return layout == LAYOUT_INTERLEAVED ? node * dim + c : (int64_t)c * nNodes + node;
}

struct Handle {
  ikb_desc desc{};
  int dim = 0, order = 0, nn = 0, nd = 0, nc = 0, npair = 0, form = 0, easM = 0;
  int easFunction = 0;  // IKB_EAS_*
  PsLaw ps;             // FORM_PS: the principal-stretch law
  bool psSet = false;
  int layout = 0;
  int64_t nElem = 0, nDof = 0, nNodes = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream: lets the host fetch R while the matrix gather still runs
  cudaEvent_t evVec = nullptr;     // recorded after the residual gather
  // Pipelined host->device solution upload (ikb_set_solution_range on Q1 interleaved meshes): the element list is cut
  // into SOL_CHUNKS chunks; chunk c needs dofs < chunkDofEnd[c] only, so d is copied in pieces on the side stream and
  // the element kernel of chunk c starts as soon as piece c has landed.
  static constexpr int SOL_CHUNKS = 4;
  std::vector<int64_t> chunkElemEnd, chunkDofEnd;
  cudaEvent_t evPiece[SOL_CHUNKS] = {nullptr, nullptr, nullptr, nullptr};
  bool piecesPending = false;
  cudaEvent_t evUse = nullptr;   // recorded on the main stream after the latest kernel/copy touching U
  bool solutionShared = false;   // U was handed out through ikb_device_ptr: fall back to ordering behind the whole stream
  cudaEvent_t evFork = nullptr;    // main stream -> side stream fork point (element kernel done)
  std::string lastError;
  int64_t launches = 0;

  // mesh
  DevBuf<double> X;         // [nc*dim][nElem]  corner coordinates, element fastest
  DevBuf<int32_t> elemNode; // [nn][nElem]      global node ids, element fastest
  DevBuf<double> Lap;       // [npair][nElem] displacement-independent Laplacian table of Q1 elements
  bool meshUploaded = false;

  // row ownership (multi-GPU); defaults to all nodes
  int64_t rowBegin = 0, rowEnd = -1;

  // state
  DevBuf<double> U, Fext, Corr;
  double lambda = 0.0;
  int fextScales = 1;
  bool hasFext = false;
  DevBuf<uint8_t> flags;
  bool hasFlags = false;
  uint64_t stateVersion = 1;  // bumped on every change of d / lambda / alpha / fext

  // pattern (node-block) + gather map
  bool patternBuilt = false;
  int64_t nBlocks = 0;
  DevBuf<int32_t> nbrPtr, nbrIdx, nbrRow, cptr;
  DevBuf<uint32_t> csrc;
  // node -> adjacent (element, local node) list in ascending element order, and for each of them the slot of
  // every element node in the node's pattern row (gather map of the warp-per-node gather)
  DevBuf<int32_t> adjPtr;       // [nRowNodes+1]
  DevBuf<uint32_t> adjCode;     // [nAdj]  e*n + la
  DevBuf<uint8_t> slotTab;      // [nAdj][n]
  int maxNbr = 0;               // longest pattern row in nodes
  int pullStageMax = 1 << 30;   // clamped to PULL_CAP at launch (IKB_PULL_STAGE_MAX, test hook)
  int pullWarps = 4;            // warps per CTA of the pull gather (IKB_PULL_WARPS, tuning; 4 measured 2 % faster than 8)
  bool pullMirror = false;      // IKB_PULL_MIRROR=1: gather the upper block triangle only and store every block twice (bit-identical; measured slower: 0.205 vs 0.184 ms on C2)
  DevBuf<int32_t> rowDiag, rowLowEnd, mirrorBlk;
  bool pullIdx64 = false;       // force the 64-bit offset path of the pull gather (IKB_PULL_IDX64, test hook)
  bool elemMma = true;          // Hex8 NeoHooke/LinearElastic: tangent contraction by DMMA (IKB_ELEM=fma: FMA kernel)
  int h8MinBlocks = 4;          // register budget of the DMMA kernel as resident CTAs per SM (IKB_H8_MINB, tuning)
  bool pullAsync = false;       // IKB_PULL_ASYNC=1: row-pipelined cp.async gather for Q1 kinds, Raw/Full (ikb_gather_async.cuh; measured slower)
  DevBuf<uint8_t> rowSlow;      // [nRowNodes] row or one of its column nodes constrained
  bool rowSlowValid = false;
  // Interleaved sweep (Q1 kinds, pull gather): the element list is cut into sweepChunks chunks; the rows that are
  // complete after chunk c (a prefix of the row list) are gathered on a side stream while chunk c+1 is evaluated, so
  // the staged K_e of a chunk is read back while it still sits in L2.  Opt-in (IKB_CHUNKS=n): measured slower than the two
  // back-to-back kernels on C2 (0.340 ms with 4 chunks, 0.347 with 8, 0.326 back to back).
  int sweepChunks = 0;
  std::vector<int64_t> sweepElemEnd, sweepRowEnd;  // per chunk: end of its element range, end of the complete row prefix
  bool sweepChunksBuilt = false;
  cudaEvent_t evChunk[32] = {};
  bool gatherPull = true;       // matrix gather through the per-block contribution lists (IKB_GATHER=tile: warp tile gather)
  // reduced-mode structures
  bool reducedBuilt = false;
  int64_t nRed = 0, nnzRed = 0;
  DevBuf<int32_t> cbelow;       // constraintsBelow [nDof]
  DevBuf<uint16_t> freeCnt;     // [nBlocks][dim]
  DevBuf<uint16_t> freeTot;     // [nRowNodes][dim]
  DevBuf<int64_t> redRowStart;  // [nDofLocalRows+1] indexed by local scalar row (row order), exclusive scan
  DevBuf<int32_t> redInner;     // [nnzRed]
  DevBuf<int64_t> redOuter;     // [nRed+1]

  // staging + results
  DevBuf<double> Kst, Rst, Est;
  DevBuf<double> vals[3];      // indexed by IKB_DBC_*
  DevBuf<double> vec[3];
  uint64_t valsVersion[3] = {0, 0, 0}, vecVersion[3] = {0, 0, 0}, energyVersion = 0;
  uint64_t stagedVersion = 0;
  unsigned stagedWhat = 0;
  double energy = 0.0;
  DevBuf<int16_t> gatherTab;   // gather offset table of this handle's (dim, nodes per element)
  DevBuf<double> scratch;      // reductions
  int spmvBlocks = 888;        // grid of the SpMV inside PCG: 6 resident blocks x 148 SMs (IKB_SPMV_BLOCKS overrides, for tuning)
  DevBuf<int32_t> errFlag;     // first failing element (material abort), INT_MAX if none

  // pipelined sweep (ikb_fused.cuh): Hex8, NeoHooke / LinearElastic
  bool fusedEnabled = false;    // IKB_FUSED=1 selects the pipelined sweep (measured slower than back-to-back)
  bool fusedTried = false, fusedOk = false;
  double fusedRingMB = 32.0;    // ring budget (IKB_RING_MB); grown until no producer ticket has to wait, cut to the mesh
  int64_t ringElems = 0;
  unsigned nElemTickets = 0, nRowTickets = 0, fusedEpoch = 0;
  unsigned long long elemTicketBase = 0, rowTicketBase = 0;
  int sweepElemGrid = 0, sweepRowGrid = 0;
  int sweepElemCtas = 2, sweepRowCtas = 8;  // resident CTAs per SM of the two persistent kernels (IKB_SWEEP_CTAS=e,r)
  int sweepMargin = -1;         // IKB_SWEEP_MARGIN (test hook): producer tickets assumed in flight when sizing the ring
  bool sweepDebug = false;      // IKB_SWEEP_DEBUG: kernel start/end stamps and a state dump when a wait gives up
  bool fusedGuards = false;     // the ring is smaller than the mesh: producer tickets carry write-after-read guards
  DevBuf<double> ring, rring;   // [ringElems][36*9], [ringElems][24]
  DevBuf<uint32_t> csrcRing, adjRing, sweepGuard;
  DevBuf<uint32_t> sweepRowWait;  // uint2 per consumer ticket
  DevBuf<unsigned long long> sweepCtl;
  DevBuf<unsigned> sweepDone;   // producer groups, then consumer groups
  int64_t nElemGroups = 0, nRowGroups = 0;
  cudaStream_t stream3 = nullptr;  // the consumer kernel runs beside the producer
  cudaEvent_t evSweepFork = nullptr, evSweepJoin = nullptr;

  // EAS
  DevBuf<double> alpha;        // [nElem][m]
  DevBuf<double> T0inv;        // [S*S][nElem] (T(center) detJ0)^-1 per element

  // PCG work
  DevBuf<double> cgR, cgZ, cgP, cgQ, cgX, cgDinv, cgB;
  DevBuf<double> cgScal;       // device scalars
  double* hostScal = nullptr;  // pinned
  DevBuf<uint8_t> cgState;     // CgState + arrival counter of the sync-free PCG
  cudaGraphExec_t cgGraph = nullptr;  // one captured batch of PCG iterations
  const void* cgGraphKey[3] = {nullptr, nullptr, nullptr};
  DevBuf<uint8_t> tcgState;           // TcgState of the sync-free truncated CG
  cudaGraphExec_t tcgGraph = nullptr;
  const void* tcgGraphKey[3] = {nullptr, nullptr, nullptr};
  long long tcgGraphMaxIters = -1;

  // NCCL (element-partitioned runs)
  void* comm = nullptr;
  int rank = 0, nranks = 1;
  int64_t colBegin = 0, colEnd = 0;          // node range touched by the uploaded elements (owned + ghost)
  std::vector<int64_t> peerRanges;           // [nranks][4] = ownBegin, ownEnd, needBegin, needEnd (nodes)
  DevBuf<double> cgPglob;                    // search direction, global length (owned + halo filled)
  // peer-memory transport of the distributed PCG (ikb_pcg_peer.cuh)
  bool peerReady = false;
  bool peerEnabled = true;                   // IKB_PCG_PEER=0: NCCL transport
  DevBuf<unsigned long long> peerWin;        // my control window (exported by IPC)
  std::vector<void*> peerOpened;             // pointers opened with cudaIpcOpenMemHandle (closed on destroy)
  void* peerCommHost = nullptr;              // PeerComm (host copy, passed by value to the kernels)
  unsigned long long peerEpoch = 0;
  int64_t peerFirstInterior = -1, peerEndInterior = -1;
  DevBuf<uint8_t> peerState;
  cudaGraphExec_t peerGraph = nullptr;
  const void* peerGraphKey[2] = {nullptr, nullptr};

  PatternView view() const {
    PatternView P;
    P.dim = dim;
    P.layout = layout;
    P.nNodes = nNodes;
    P.rowBegin = rowBegin;
    P.nRowNodes = rowEnd - rowBegin;
    P.nBlocks = nBlocks;
    P.nbrPtr = nbrPtr.p;
    P.nbrIdx = nbrIdx.p;
    P.nbrRow = nbrRow.p;
    return P;
  }
  int64_t nnzRaw() const { return (int64_t)dim * dim * nBlocks; }
  int64_t nRowsLocal() const { return (rowEnd - rowBegin) * dim; }
};

inline int fail(Handle* h, int code, const std::string& msg) {
  if (h) h->lastError = msg;
  return code;
}

#define IKB_CUDA(h, expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ikb::fail(h, IKB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)

#define IKB_LAUNCH_CHECK(h)                                                                       \
  do {                                                                                            \
    (h)->launches++;                                                                              \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess)                                                                        \
      return ikb::fail(h, IKB_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(_e));    \
  } while (0)

inline unsigned gridFor(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

}  // namespace ikb
