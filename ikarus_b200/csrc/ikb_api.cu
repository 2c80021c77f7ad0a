// C-ABI entry points of libikb200.so (see include/ikb200.h).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>

#include "ikb_dist.cuh"
#include "ikb_elem_eas.cuh"
#include "ikb_elem_easdg.cuh"
#include "ikb_elem_h8mma.cuh"
#include "ikb_elem_q1.cuh"
#include "ikb_elem_q2.cuh"
#include "ikb_fused.cuh"
#include "ikb_gather.cuh"
#include "ikb_gather_async.cuh"
#include "ikb_internal.cuh"
#include "ikb_pattern.cuh"
#include "ikb_pcg.cuh"
#include "ikb_pcg_peer.cuh"
#include "ikb_results.cuh"

using namespace ikb;

namespace {

constexpr int RED_BLOCKS = 592;  // 4 x 148 SMs; fixed so reductions are reproducible
constexpr int MAX_SPMV_BLOCKS = 2368;  // upper bound of the PCG SpMV grid (p.q partials live in scratch[0, MAX))

Handle* H(ikb_handle h) { return reinterpret_cast<Handle*>(h); }

int checkHandle(Handle* h) {
  if (!h) return IKB_EINVAL;
  cudaSetDevice(h->device);
  return IKB_OK;
}

int dbcValid(int dbc) { return dbc == IKB_DBC_RAW || dbc == IKB_DBC_REDUCED || dbc == IKB_DBC_FULL; }

template <int D, int M>
cudaError_t launchEasForm(Handle* h, const EasArgs& EA) {
  if (h->form == FORM_LE) return launchElemEas<D, FORM_LE, M>(EA, h->stream);
  if (h->form == FORM_SVK) return launchElemEas<D, FORM_SVK, M>(EA, h->stream);
  return launchElemEas<D, FORM_NH, M>(EA, h->stream);
}

// displacement-gradient enhancement (ikb_elem_easdg.cuh)
template <int D>
cudaError_t launchEasDg(Handle* h, const EasArgs& EA) {
  const bool tr = h->easFunction == IKB_EAS_DISPLACEMENT_GRADIENT_TRANSPOSED;
  if (h->form == FORM_SVK) return tr ? launchElemEasDg<D, FORM_SVK, ENH_DGT>(EA, h->stream) : launchElemEasDg<D, FORM_SVK, ENH_DG>(EA, h->stream);
  if (h->form == FORM_NH) return tr ? launchElemEasDg<D, FORM_NH, ENH_DGT>(EA, h->stream) : launchElemEasDg<D, FORM_NH, ENH_DG>(EA, h->stream);
  if (h->form == FORM_PS) return tr ? launchElemEasDg<D, FORM_PS, ENH_DGT>(EA, h->stream) : launchElemEasDg<D, FORM_PS, ENH_DG>(EA, h->stream);
  return cudaErrorInvalidValue;
}

// principal-stretch laws: the plain element (m = 0) and the strain enhancements run through the generalised-tangent
// kernel as well (ikb_elem_easdg.cuh, ENH_STRAIN)
cudaError_t launchPsStrain(Handle* h, const EasArgs& EA) {
  if (h->dim == 2) {
    if (h->easM == 0) return launchElemEasDg<2, FORM_PS, ENH_STRAIN, 0>(EA, h->stream);
    if (h->easM == 4) return launchElemEasDg<2, FORM_PS, ENH_STRAIN, 4>(EA, h->stream);
    if (h->easM == 5) return launchElemEasDg<2, FORM_PS, ENH_STRAIN, 5>(EA, h->stream);
    if (h->easM == 7) return launchElemEasDg<2, FORM_PS, ENH_STRAIN, 7>(EA, h->stream);
  } else {
    if (h->easM == 0) return launchElemEasDg<3, FORM_PS, ENH_STRAIN, 0>(EA, h->stream);
    if (h->easM == 9) return launchElemEasDg<3, FORM_PS, ENH_STRAIN, 9>(EA, h->stream);
    if (h->easM == 21) return launchElemEasDg<3, FORM_PS, ENH_STRAIN, 21>(EA, h->stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launchEas(Handle* h, const EasArgs& EA) {
  if (h->easFunction != IKB_EAS_STRAIN) return h->dim == 2 ? launchEasDg<2>(h, EA) : launchEasDg<3>(h, EA);
  if (h->form == FORM_PS) return launchPsStrain(h, EA);
  if (h->dim == 2) {
    if (h->easM == 4) return launchEasForm<2, 4>(h, EA);
    if (h->easM == 5) return launchEasForm<2, 5>(h, EA);
    if (h->easM == 7) return launchEasForm<2, 7>(h, EA);
  } else {
    if (h->easM == 9) return launchEasForm<3, 9>(h, EA);
    if (h->easM == 21) return launchEasForm<3, 21>(h, EA);
  }
  return cudaErrorInvalidValue;
}

void markSolutionUse(Handle* h) { cudaEventRecord(h->evUse, h->stream); }

// Orders the main stream behind a pipelined solution upload that no element kernel has consumed yet.
void joinSolution(Handle* h) {
  if (!h->piecesPending) return;
  cudaStreamWaitEvent(h->stream, h->evPiece[Handle::SOL_CHUNKS - 1], 0);
  h->piecesPending = false;
}

// the Q1 element kernel of the handle's kind over the element range described by A
cudaError_t launchQ1Kernel(Handle* h, const ElemArgs& A) {
  cudaError_t e = cudaErrorInvalidValue;
  if (h->dim == 3 && h->elemMma && A.Lap && h->form != FORM_SVK) {
    // Hex8 with the contraction on the FP64 tensor cores (ikb_elem_h8mma.cuh); IKB_ELEM=fma selects the FMA kernel
    if (h->form == FORM_LE) e = launchElemH8Mma<FORM_LE>(A, h->stream, h->h8MinBlocks);
    if (h->form == FORM_NH) e = launchElemH8Mma<FORM_NH>(A, h->stream, h->h8MinBlocks);
  } else if (h->dim == 3) {
    if (h->form == FORM_LE) e = launchElemQ1<3, FORM_LE>(A, h->stream);
    if (h->form == FORM_SVK) e = launchElemQ1<3, FORM_SVK>(A, h->stream);
    if (h->form == FORM_NH) e = launchElemQ1<3, FORM_NH>(A, h->stream);
  } else {
    if (h->form == FORM_LE) e = launchElemQ1<2, FORM_LE>(A, h->stream);
    if (h->form == FORM_SVK) e = launchElemQ1<2, FORM_SVK>(A, h->stream);
    if (h->form == FORM_NH) e = launchElemQ1<2, FORM_NH>(A, h->stream);
  }
  return e;
}

int launchElements(Handle* h, unsigned what, const double* dU = nullptr, const double* Uoverride = nullptr,
                   double* alphaOverride = nullptr) {
  ElemArgs A;
  A.X = h->X.p;
  A.elemNode = h->elemNode.p;
  A.U = h->U.p;
  A.Kst = h->Kst.p;
  A.Rst = h->Rst.p;
  A.Est = h->Est.p;
  A.errFlag = h->errFlag.p;
  A.Lap = h->Lap.p;
  A.nElem = h->nElem;
  A.elemBegin = 0;
  A.elemCount = h->nElem;
  A.nNodes = h->nNodes;
  A.layout = h->layout;
  A.lambda = h->desc.lambda;
  A.mu = h->desc.mu;
  A.what = what;
  A.planeStress = h->desc.plane_strain == IKB_REDUCE_PLANE_STRESS;
  A.psTol = h->desc.reduce_tol > 0.0 ? h->desc.reduce_tol : 1e-12;
  cudaError_t e = cudaErrorInvalidValue;
  if (h->order == 1 && h->easM == 0 && h->form != FORM_PS) {
    // a pipelined solution upload is consumed chunk by chunk: chunk c waits for piece c of d only
    const int nch = h->piecesPending ? Handle::SOL_CHUNKS : 1;
    for (int c = 0; c < nch; ++c) {
      if (h->piecesPending) {
        cudaStreamWaitEvent(h->stream, h->evPiece[c], 0);
        const int64_t b = c ? h->chunkElemEnd[c - 1] : 0;
        A.elemBegin = b;
        A.elemCount = h->chunkElemEnd[c] - b;
        A.X = h->X.p + b;
        A.elemNode = h->elemNode.p + b;
        A.Lap = h->Lap.p ? h->Lap.p + b : nullptr;
        A.Kst = h->Kst.p + (size_t)b * h->npair * h->dim * h->dim;
        A.Rst = h->Rst.p + (size_t)b * h->nd;
        A.Est = h->Est.p + b;
        if (c) h->launches++;
      }
      e = launchQ1Kernel(h, A);
      if (e != cudaSuccess) break;
    }
    h->piecesPending = false;
  } else if (h->order == 1) {
    EasArgs EA;
    EA.E = A;
    EA.T0inv = h->T0inv.p;
    EA.alpha = alphaOverride ? alphaOverride : h->alpha.p;
    if (Uoverride) EA.E.U = Uoverride;
    EA.dU = dU;
    EA.updateMode = dU ? 1 : 0;
    EA.ps = h->ps;
    e = launchEas(h, EA);
  } else if (h->order == 2 && h->easM == 0) {
    if (h->dim == 3) {
      if (h->form == FORM_LE) e = launchElemQ2<3, FORM_LE>(A, h->stream);
      if (h->form == FORM_SVK) e = launchElemQ2<3, FORM_SVK>(A, h->stream);
      if (h->form == FORM_NH) e = launchElemQ2<3, FORM_NH>(A, h->stream);
    } else {
      if (h->form == FORM_LE) e = launchElemQ2<2, FORM_LE>(A, h->stream);
      if (h->form == FORM_SVK) e = launchElemQ2<2, FORM_SVK>(A, h->stream);
      if (h->form == FORM_NH) e = launchElemQ2<2, FORM_NH>(A, h->stream);
    }
  } else {
    return fail(h, IKB_ENOTIMPL, "EAS is only supported for Q1 and H1 elements");
  }
  h->launches++;
  markSolutionUse(h);
  if (e != cudaSuccess) return fail(h, IKB_ECUDA, std::string("element kernel: ") + cudaGetErrorString(e));
  return IKB_OK;
}

// staging arrays of the two-kernel path, allocated for what is actually staged (the fused sweep needs Est only)
int ensureStaging(Handle* h, unsigned what) {
  const size_t dd = (size_t)h->dim * h->dim;
  if ((what & IKB_MATRIX) && !h->Kst.p) IKB_CUDA(h, h->Kst.alloc((size_t)h->nElem * h->npair * dd));
  if ((what & IKB_VECTOR) && !h->Rst.p) IKB_CUDA(h, h->Rst.alloc((size_t)h->nElem * h->nd));
  if ((what & IKB_SCALAR) && !h->Est.p) IKB_CUDA(h, h->Est.alloc((size_t)h->nElem));
  return IKB_OK;
}

int ensureSolution(Handle* h) {
  if (!h->U.p) {
    IKB_CUDA(h, h->U.alloc((size_t)h->nDof));
    IKB_CUDA(h, cudaMemsetAsync(h->U.p, 0, h->U.bytes(), h->stream));
    markSolutionUse(h);
  }
  return IKB_OK;
}

int ensureReduced(Handle* h) {
  if (h->reducedBuilt) return IKB_OK;
  if (!h->hasFlags) return fail(h, IKB_ESTATE, "Reduced mode needs ikb_upload_dirichlet first");
  if (!h->patternBuilt) return fail(h, IKB_ESTATE, "pattern not built");
  if (h->rowBegin != 0 || h->rowEnd != h->nNodes)
    return fail(h, IKB_ENOTIMPL, "DBCOption::Reduced is only available on an unpartitioned handle");
  const PatternView P = h->view();
  const int64_t n = h->nDof;
  const int tpb = 256;
  // constraintsBelow = exclusive prefix sum of the flags (assembler/interface.hh:51-62)
  DevBuf<int32_t> fi;
  IKB_CUDA(h, fi.alloc((size_t)n + 1));
  IKB_CUDA(h, cudaMemsetAsync(fi.p, 0, fi.bytes(), h->stream));
  flags_to_int_kernel<<<gridFor(n, tpb), tpb, 0, h->stream>>>(h->flags.p, n, fi.p);
  IKB_LAUNCH_CHECK(h);
  IKB_CUDA(h, h->cbelow.alloc((size_t)n + 1));
  size_t tmpBytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, fi.p, h->cbelow.p, n + 1, h->stream);
  DevBuf<uint8_t> tmp;
  IKB_CUDA(h, tmp.alloc(tmpBytes));
  IKB_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, fi.p, h->cbelow.p, n + 1, h->stream));
  h->launches++;
  int32_t nFixed = 0;
  IKB_CUDA(h, cudaMemcpyAsync(&nFixed, h->cbelow.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->nRed = n - nFixed;

  IKB_CUDA(h, h->freeCnt.alloc((size_t)h->nBlocks * h->dim));
  IKB_CUDA(h, h->freeTot.alloc((size_t)P.nRowNodes * h->dim));
  free_counts_kernel<<<gridFor(P.nRowNodes, tpb), tpb, 0, h->stream>>>(P, h->flags.p, h->freeCnt.p, h->freeTot.p);
  IKB_LAUNCH_CHECK(h);
  const int64_t nRows = P.nRowNodes * h->dim;
  DevBuf<int64_t> rowCount;
  IKB_CUDA(h, rowCount.alloc((size_t)nRows + 1));
  IKB_CUDA(h, cudaMemsetAsync(rowCount.p, 0, rowCount.bytes(), h->stream));
  row_free_count_kernel<<<gridFor(nRows, tpb), tpb, 0, h->stream>>>(P, h->flags.p, h->freeTot.p, rowCount.p);
  IKB_LAUNCH_CHECK(h);
  IKB_CUDA(h, h->redRowStart.alloc((size_t)nRows + 1));
  tmpBytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, rowCount.p, h->redRowStart.p, nRows + 1, h->stream);
  IKB_CUDA(h, tmp.alloc(tmpBytes));
  IKB_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, rowCount.p, h->redRowStart.p, nRows + 1, h->stream));
  h->launches++;
  IKB_CUDA(h, cudaMemcpyAsync(&h->nnzRed, h->redRowStart.p + nRows, sizeof(int64_t), cudaMemcpyDeviceToHost,
                              h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  IKB_CUDA(h, h->redInner.alloc((size_t)std::max<int64_t>(h->nnzRed, 1)));
  IKB_CUDA(h, h->redOuter.alloc((size_t)h->nRed + 1));
  reduced_inner_kernel<<<gridFor(h->nBlocks, tpb), tpb, 0, h->stream>>>(P, h->flags.p, h->freeCnt.p, h->freeTot.p,
                                                                        h->redRowStart.p, h->cbelow.p, h->redInner.p);
  IKB_LAUNCH_CHECK(h);
  reduced_outer_kernel<<<gridFor(std::max<int64_t>(nRows, 1), tpb), tpb, 0, h->stream>>>(
      h->flags.p, h->cbelow.p, h->redRowStart.p, nRows, 0, h->redOuter.p, h->nnzRed, h->nRed);
  IKB_LAUNCH_CHECK(h);
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  tmp.release();
  fi.release();
  rowCount.release();
  h->reducedBuilt = true;
  return IKB_OK;
}

int64_t rowsOf(Handle* h, int dbc) { return dbc == IKB_DBC_REDUCED ? h->nRed : h->nRowsLocal(); }
int64_t nnzOf(Handle* h, int dbc) { return dbc == IKB_DBC_REDUCED ? h->nnzRed : h->nnzRaw(); }

GatherArgs makeGatherArgs(Handle* h, unsigned what, int dbc) {
  GatherArgs G;
  G.P = h->view();
  G.adjPtr = h->adjPtr.p;
  G.adjCode = h->adjCode.p;
  G.slotTab = h->slotTab.p;
  G.Kst = h->Kst.p;
  G.Rst = h->Rst.p;
  G.fext = h->hasFext ? h->Fext.p : nullptr;
  G.fextScale = h->fextScales ? h->lambda : 1.0;
  G.flags = h->hasFlags ? h->flags.p : nullptr;
  G.vals = (what & IKB_MATRIX) ? h->vals[dbc].p : nullptr;
  G.vec = (what & IKB_VECTOR) ? h->vec[dbc].p : nullptr;
  G.dbc = dbc;
  G.npair = h->npair;
  // tile rows are padded to a stride == D (mod 16) doubles; Q2 kinds gather one scalar row per warp (GatherCfg)
  const int rs = h->nn > 8 ? 1 : h->dim;
  G.maxOut = rs * (h->dim * std::max(h->maxNbr, 1) + 16);
  G.offTab = nullptr;
  G.freeCnt = h->freeCnt.p;
  G.freeTot = h->freeTot.p;
  G.redRowStart = h->redRowStart.p;
  G.cbelow = h->cbelow.p;
  G.redVecOffset = 0;
  G.pullStageMax = std::min(h->pullStageMax, PULL_CAP);
  return G;
}

ElemArgs makeElemArgs(Handle* h, unsigned what) {
  ElemArgs A;
  A.X = h->X.p;
  A.elemNode = h->elemNode.p;
  A.U = h->U.p;
  A.Kst = h->Kst.p;
  A.Rst = h->Rst.p;
  A.Est = h->Est.p;
  A.errFlag = h->errFlag.p;
  A.Lap = h->Lap.p;
  A.nElem = h->nElem;
  A.elemBegin = 0;
  A.elemCount = h->nElem;
  A.nNodes = h->nNodes;
  A.layout = h->layout;
  A.lambda = h->desc.lambda;
  A.mu = h->desc.mu;
  A.what = what;
  A.planeStress = h->desc.plane_strain == IKB_REDUCE_PLANE_STRESS;
  A.psTol = h->desc.reduce_tol > 0.0 ? h->desc.reduce_tol : 1e-12;
  return A;
}

// One-time maps, ring and guards of the pipelined sweep (after the pattern build).  Leaves fusedOk false when the mesh or
// element kind is outside what the sweep covers; the back-to-back two-kernel path serves those.
int ensureFused(Handle* h) {
  if (h->fusedTried) return IKB_OK;
  h->fusedTried = true;
  h->fusedOk = false;
  if (!h->fusedEnabled || h->dim != 3 || h->nn != 8 || h->easM != 0 || h->form == FORM_SVK || !h->Lap.p) return IKB_OK;
  if (!h->gatherPull || !h->cptr.p || !h->csrc.p || !h->adjPtr.p || h->nBlocks == 0 || h->nElem == 0) return IKB_OK;
  const int64_t nRowNodes = h->rowEnd - h->rowBegin;
  if (nRowNodes == 0) return IKB_OK;
  const int tpb = 256;
  int32_t nAdj = 0, nContrib = 0;
  IKB_CUDA(h, cudaMemcpy(&nAdj, h->adjPtr.p + nRowNodes, sizeof(int32_t), cudaMemcpyDeviceToHost));
  IKB_CUDA(h, cudaMemcpy(&nContrib, h->cptr.p + h->nBlocks, sizeof(int32_t), cudaMemcpyDeviceToHost));
  h->nElemTickets = (unsigned)((h->nElem + 3) / 4);
  h->nRowTickets = (unsigned)((nRowNodes + SWEEP_RB - 1) / SWEEP_RB);
  h->nElemGroups = (h->nElemTickets + SWEEP_EG - 1) / SWEEP_EG;
  h->nRowGroups = (h->nRowTickets + SWEEP_RG - 1) / SWEEP_RG;
  // dependency data: per consumer ticket the producer-ticket range, per producer ticket its highest row
  DevBuf<int32_t> dLo, dHi, dMaxRow;
  IKB_CUDA(h, dLo.alloc(h->nRowTickets));
  IKB_CUDA(h, dHi.alloc(h->nRowTickets));
  IKB_CUDA(h, dMaxRow.alloc(h->nElemTickets));
  sweep_row_range_kernel<<<gridFor(h->nRowTickets, tpb), tpb, 0, h->stream>>>(nRowNodes, h->adjPtr.p, h->adjCode.p, dLo.p, dHi.p);
  IKB_LAUNCH_CHECK(h);
  sweep_ticket_rows_kernel<<<gridFor(h->nElemTickets, tpb), tpb, 0, h->stream>>>(h->elemNode.p, h->nElem, h->rowBegin, h->rowEnd,
                                                                               dMaxRow.p);
  IKB_LAUNCH_CHECK(h);
  std::vector<int32_t> rowLo(h->nRowTickets), rowHi(h->nRowTickets), maxRow(h->nElemTickets);
  IKB_CUDA(h, cudaMemcpyAsync(rowLo.data(), dLo.p, dLo.bytes(), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaMemcpyAsync(rowHi.data(), dHi.p, dHi.bytes(), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaMemcpyAsync(maxRow.data(), dMaxRow.p, dMaxRow.bytes(), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  std::vector<uint32_t> wait((size_t)h->nRowTickets * 2);
  for (size_t t = 0; t < rowLo.size(); ++t) {
    if (rowHi[t] < 0) {
      wait[2 * t] = 1;  // rows without elements: nothing to wait for
      wait[2 * t + 1] = 0;
    } else {
      wait[2 * t] = (uint32_t)(rowLo[t] / SWEEP_EG);
      wait[2 * t + 1] = (uint32_t)(rowHi[t] / SWEEP_EG);
    }
  }
  // Ring size (a multiple of 4 elements = whole producer tickets): starts at the budget and grows until no producer
  // ticket would have to wait for a consumer ticket that depends on one of the `margin` newest producer tickets --
  // the tickets that can be in flight when every resident producer warp holds one, times two for uneven progress.
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const int64_t margin = h->sweepMargin >= 0 ? h->sweepMargin : 2 * (int64_t)sms * h->sweepElemCtas * H8Cfg::WARPS;
  const int64_t elemsPadded = (h->nElem + 3) / 4 * 4;
  int64_t ringElems = std::min<int64_t>(elemsPadded, std::max<int64_t>(8, (int64_t)(h->fusedRingMB * 1048576.0 / (348.0 * 8.0)) / 4 * 4));
  std::vector<uint32_t> guard;
  while (ringElems < elemsPadded &&
         !sweepBuildGuards(h->nElemTickets, h->nRowTickets, rowHi, maxRow, ringElems, margin, guard))
    ringElems = std::min<int64_t>(elemsPadded, (ringElems + ringElems / 4 + 3) / 4 * 4);
  h->ringElems = ringElems;
  h->fusedGuards = ringElems < elemsPadded;
  IKB_CUDA(h, h->ring.alloc((size_t)ringElems * 36 * 9));
  IKB_CUDA(h, h->rring.alloc((size_t)ringElems * 24));
  if (h->fusedGuards) {
    IKB_CUDA(h, h->sweepGuard.alloc(guard.size()));
    IKB_CUDA(h, cudaMemcpy(h->sweepGuard.p, guard.data(), h->sweepGuard.bytes(), cudaMemcpyHostToDevice));
  }
  IKB_CUDA(h, h->sweepRowWait.alloc(wait.size()));
  IKB_CUDA(h, cudaMemcpy(h->sweepRowWait.p, wait.data(), h->sweepRowWait.bytes(), cudaMemcpyHostToDevice));
  IKB_CUDA(h, h->csrcRing.alloc((size_t)std::max(nContrib, 1)));
  IKB_CUDA(h, h->adjRing.alloc((size_t)std::max(nAdj, 1)));
  sweep_ring_codes_kernel<<<gridFor(nContrib, tpb), tpb, 0, h->stream>>>(h->csrc.p, nContrib, h->npair, (uint32_t)ringElems,
                                                                        h->csrcRing.p);
  IKB_LAUNCH_CHECK(h);
  sweep_ring_adj_kernel<<<gridFor(nAdj, tpb), tpb, 0, h->stream>>>(h->adjCode.p, nAdj, (uint32_t)ringElems, h->adjRing.p);
  IKB_LAUNCH_CHECK(h);
  IKB_CUDA(h, h->sweepDone.alloc((size_t)(h->nElemGroups + h->nRowGroups)));
  IKB_CUDA(h, cudaMemsetAsync(h->sweepDone.p, 0, h->sweepDone.bytes(), h->stream));
  IKB_CUDA(h, h->sweepCtl.alloc(8));
  IKB_CUDA(h, cudaMemsetAsync(h->sweepCtl.p, 0, h->sweepCtl.bytes(), h->stream));
  h->fusedEpoch = 0;
  h->elemTicketBase = h->rowTicketBase = 0;
  if (!h->stream3) {
    int prioLo = 0, prioHi = 0;
    cudaDeviceGetStreamPriorityRange(&prioLo, &prioHi);
    IKB_CUDA(h, cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, prioLo));
    IKB_CUDA(h, cudaEventCreateWithFlags(&h->evSweepFork, cudaEventDisableTiming));
    IKB_CUDA(h, cudaEventCreateWithFlags(&h->evSweepJoin, cudaEventDisableTiming));
  }
  // Load both kernels NOW: with lazy module loading the first launch of a kernel synchronises with running work, and a
  // consumer kernel that cannot start until the producer has left would leave the producer waiting for it forever.
  {
    cudaFuncAttributes fa;
    cudaError_t le = cudaSuccess;
    auto touch = [&](const void* f) {
      if (le == cudaSuccess) le = cudaFuncGetAttributes(&fa, f);
    };
    touch((const void*)sweep_elem_kernel<FORM_LE>);
    touch((const void*)sweep_elem_kernel<FORM_NH>);
#define IKB_TOUCH_ROWS(MODE)                                         \
  touch((const void*)sweep_rows_kernel<MODE, true, 16>);             \
  touch((const void*)sweep_rows_kernel<MODE, false, 16>);            \
  touch((const void*)sweep_rows_kernel<MODE, true, 12>);             \
  touch((const void*)sweep_rows_kernel<MODE, false, 12>);
    IKB_TOUCH_ROWS(IKB_DBC_RAW)
    IKB_TOUCH_ROWS(IKB_DBC_FULL)
    IKB_TOUCH_ROWS(IKB_DBC_REDUCED)
#undef IKB_TOUCH_ROWS
    if (le != cudaSuccess) return fail(h, IKB_ECUDA, std::string("pipelined sweep: ") + cudaGetErrorString(le));
  }
  h->sweepElemGrid = sms * h->sweepElemCtas;
  h->sweepRowGrid = sms * h->sweepRowCtas;
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->fusedOk = true;
  return IKB_OK;
}

// K, R and E of the current state: producer and consumer kernel side by side (Hex8 pipelined sweep)
int launchFused(Handle* h, unsigned what, int dbc) {
  SweepArgs S;
  S.E = makeElemArgs(h, what);
  S.O.Kst = h->ring.p;
  S.O.Rst = h->rring.p;
  S.O.ringElems = h->ringElems;
  S.guard = h->fusedGuards ? h->sweepGuard.p : nullptr;
  S.G = makeGatherArgs(h, what, dbc);
  S.G.Kst = h->ring.p;
  S.G.Rst = h->rring.p;
  S.cptr = h->cptr.p;
  S.csrcRing = h->csrcRing.p;
  S.adjRing = h->adjRing.p;
  S.rowWait = reinterpret_cast<const uint2*>(h->sweepRowWait.p);
  S.ctl = reinterpret_cast<SweepCtl*>(h->sweepCtl.p);
  S.elemDone = h->sweepDone.p;
  S.rowDone = h->sweepDone.p + h->nElemGroups;
  S.nElemTickets = h->nElemTickets;
  S.nRowTickets = h->nRowTickets;
  S.elemTicketBase = h->elemTicketBase;
  S.rowTicketBase = h->rowTicketBase;
  S.epoch = h->fusedEpoch;
  S.what = what;
  S.timing = h->sweepDebug ? 1 : 0;
  if (h->sweepDebug) {
    const unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull};
    cudaMemcpyAsync(h->sweepCtl.p + 4, init, sizeof(init), cudaMemcpyHostToDevice, h->stream);
  }
  joinSolution(h);
  // fork: the consumer kernel on its own stream, behind everything queued on the main stream so far
  cudaEventRecord(h->evSweepFork, h->stream);
  cudaStreamWaitEvent(h->stream3, h->evSweepFork, 0);
  // Both kernels ask for the largest shared-memory carve-out: an SM only changes its carve-out when it is idle, so a
  // consumer CTA whose shared memory did not fit next to the producer's configuration would wait for the producer to
  // leave the SM -- and the producer may be waiting for the consumer.
  const void* ek = h->form == FORM_LE ? (const void*)sweep_elem_kernel<FORM_LE> : (const void*)sweep_elem_kernel<FORM_NH>;
  cudaError_t e = cudaFuncSetAttribute(ek, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H8Cfg::SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ek, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) {
    if (h->form == FORM_LE)
      sweep_elem_kernel<FORM_LE><<<h->sweepElemGrid, 32 * H8Cfg::WARPS, H8Cfg::SMEM, h->stream>>>(S);
    else
      sweep_elem_kernel<FORM_NH><<<h->sweepElemGrid, 32 * H8Cfg::WARPS, H8Cfg::SMEM, h->stream>>>(S);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && (what & (IKB_MATRIX | IKB_VECTOR))) {
    const int tpbR = 32 * PULL_WARPS_MAX;
#define IKB_SWEEP_ROWS2(MODE, IL)                                                               \
  if (h->sweepRowCtas > 6) {                                                                    \
    cudaFuncSetAttribute(sweep_rows_kernel<MODE, IL, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                         cudaSharedmemCarveoutMaxShared);                                       \
    sweep_rows_kernel<MODE, IL, 16><<<h->sweepRowGrid, tpbR, 0, h->stream3>>>(S);              \
  } else {                                                                                      \
    cudaFuncSetAttribute(sweep_rows_kernel<MODE, IL, 12>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                         cudaSharedmemCarveoutMaxShared);                                       \
    sweep_rows_kernel<MODE, IL, 12><<<h->sweepRowGrid, tpbR, 0, h->stream3>>>(S);              \
  }
#define IKB_SWEEP_ROWS(MODE)                                                                    \
  if (h->layout == LAYOUT_INTERLEAVED) {                                                        \
    IKB_SWEEP_ROWS2(MODE, true)                                                                 \
  } else {                                                                                      \
    IKB_SWEEP_ROWS2(MODE, false)                                                                \
  }
    if (dbc == IKB_DBC_RAW) {
      IKB_SWEEP_ROWS(IKB_DBC_RAW)
    } else if (dbc == IKB_DBC_FULL) {
      IKB_SWEEP_ROWS(IKB_DBC_FULL)
    } else {
      IKB_SWEEP_ROWS(IKB_DBC_REDUCED)
    }
#undef IKB_SWEEP_ROWS2
#undef IKB_SWEEP_ROWS
    e = cudaGetLastError();
    h->launches++;
    h->rowTicketBase += (unsigned long long)h->nRowTickets + (unsigned long long)h->sweepRowGrid * PULL_WARPS_MAX;
  }
  // join
  cudaEventRecord(h->evSweepJoin, h->stream3);
  cudaStreamWaitEvent(h->stream, h->evSweepJoin, 0);
  if (e != cudaSuccess) return fail(h, IKB_ECUDA, std::string("pipelined sweep: ") + cudaGetErrorString(e));
  h->launches++;
  h->elemTicketBase += (unsigned long long)h->nElemTickets + (unsigned long long)h->sweepElemGrid * H8Cfg::WARPS;
  h->fusedEpoch++;
  markSolutionUse(h);
  if (what & IKB_VECTOR) cudaEventRecord(h->evVec, h->stream);
  if (h->sweepDebug) {
    unsigned long long t[4];
    cudaStreamSynchronize(h->stream);
    cudaMemcpy(t, h->sweepCtl.p + 4, sizeof(t), cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[ikb sweep] producer %.1f us, consumer %.1f us, consumer starts %.1f us after producer, total %.1f us\n",
                 (t[1] - t[0]) * 1e-3, (t[3] - t[2]) * 1e-3, ((double)t[2] - (double)t[0]) * 1e-3,
                 (std::max(t[1], t[3]) - std::min(t[0], t[2])) * 1e-3);
  }
  return IKB_OK;
}

// row-pipelined cp.async gather (Q1 kinds, Raw / Full)
template <int DIM, int MODE, bool IL>
void launchRowsAsync(Handle* h, const GatherArgs& G, bool idx32, int64_t nRowNodes) {
  if constexpr (MODE != IKB_DBC_REDUCED) {
    const unsigned grid = gridFor(nRowNodes, ASYNC_WARPS * ASYNC_ROWS);
    if (idx32)
      gather_rows_async_kernel<DIM, MODE, IL, true><<<grid, ASYNC_WARPS * 32, 0, h->stream>>>(G, h->cptr.p, h->csrc.p, h->rowSlow.p);
    else
      gather_rows_async_kernel<DIM, MODE, IL, false><<<grid, ASYNC_WARPS * 32, 0, h->stream>>>(G, h->cptr.p, h->csrc.p, h->rowSlow.p);
  }
}

// rowFirst/rowEnd/st: a row range of the matrix on another stream (interleaved sweep; pull gather only)
int launchGather(Handle* h, unsigned what, int dbc, int64_t rowFirst = 0, int64_t rowEnd = -1, cudaStream_t st = nullptr) {
  GatherArgs G = makeGatherArgs(h, what, dbc);
  G.rowFirst = rowFirst;
  G.rowEnd = rowEnd;
  cudaStream_t gst = st ? st : h->stream;
  const int64_t nRange = (rowEnd < 0 ? (h->rowEnd - h->rowBegin) : rowEnd) - rowFirst;
  if (nRange <= 0 && !(what & IKB_VECTOR)) return IKB_OK;
  const int rs = h->nn > 8 ? 1 : h->dim;
  const int64_t nRowNodes = h->rowEnd - h->rowBegin;
  if (h->nBlocks == 0 || nRowNodes == 0) return IKB_OK;
  if (!h->gatherTab.p) {
    std::vector<int16_t> tab;
    if (h->dim == 3 && h->nn == 8) tab = gatherOffsetTable<3, 8>();
    if (h->dim == 3 && h->nn == 27) tab = gatherOffsetTable<3, 27>();
    if (h->dim == 2 && h->nn == 4) tab = gatherOffsetTable<2, 4>();
    if (h->dim == 2 && h->nn == 9) tab = gatherOffsetTable<2, 9>();
    IKB_CUDA(h, h->gatherTab.alloc(tab.size()));
    IKB_CUDA(h, cudaMemcpyAsync(h->gatherTab.p, tab.data(), tab.size() * sizeof(int16_t), cudaMemcpyHostToDevice, h->stream));
    IKB_CUDA(h, cudaStreamSynchronize(h->stream));  // tab is a local
  }
  G.offTab = h->gatherTab.p;
  const int warps = 8;
  const int pw = h->pullWarps;  // warps per CTA of the pull gather
  const size_t smem = (size_t)warps * G.maxOut * sizeof(double);
  // one warp per work unit (node-row, or scalar row for Q2)
  const int64_t wantBlocks = gridFor(nRowNodes * (h->dim / rs), warps);
  cudaError_t e = cudaSuccess;
  const unsigned vecGrid = gridFor(nRowNodes * h->dim, 256);
  // pull gather: signed 32-bit staged offsets (in 4-byte units) when the whole staged K_e buffer is addressable that way
  const bool idx32 = !h->pullIdx64 && (double)h->nElem * h->npair * h->dim * h->dim * 2.0 < 2147483000.0;
  // mirrored pull: one-time maps (diagonal block and leading foreign blocks per row, mirror block per block)
  const bool mirror = h->pullMirror && h->gatherPull && h->csrc.p && G.vals && dbc != IKB_DBC_REDUCED;
  if (mirror && !h->mirrorBlk.p) {
    IKB_CUDA(h, h->rowDiag.alloc((size_t)nRowNodes));
    IKB_CUDA(h, h->rowLowEnd.alloc((size_t)nRowNodes));
    IKB_CUDA(h, h->mirrorBlk.alloc((size_t)h->nBlocks));
    mirror_rows_kernel<<<gridFor(nRowNodes, 256), 256, 0, h->stream>>>(G.P, h->rowDiag.p, h->rowLowEnd.p);
    IKB_LAUNCH_CHECK(h);
    mirror_blocks_kernel<<<gridFor(h->nBlocks, 256), 256, 0, h->stream>>>(G.P, h->mirrorBlk.p);
    IKB_LAUNCH_CHECK(h);
  }
  // row-pipelined cp.async gather: Q1 kinds, Raw / Full
  const bool rowsAsync = h->pullAsync && h->gatherPull && h->csrc.p && G.vals && !mirror && h->nn <= 8 && dbc != IKB_DBC_REDUCED;
  if (rowsAsync && dbc == IKB_DBC_FULL && !h->rowSlowValid) {
    if (!h->rowSlow.p) IKB_CUDA(h, h->rowSlow.alloc((size_t)nRowNodes));
    if (h->layout == LAYOUT_INTERLEAVED)
      row_slow_kernel<true><<<gridFor(nRowNodes, 256), 256, 0, h->stream>>>(G.P, G.flags, h->rowSlow.p);
    else
      row_slow_kernel<false><<<gridFor(nRowNodes, 256), 256, 0, h->stream>>>(G.P, G.flags, h->rowSlow.p);
    IKB_LAUNCH_CHECK(h);
    h->rowSlowValid = true;
  }
  G.rowDiag = h->rowDiag.p;
  G.rowLowEnd = h->rowLowEnd.p;
  G.mirrorBlk = h->mirrorBlk.p;
#define IKB_GATHER3(DIM, NN, MODE, IL)                                                                              \
  {                                                                                                                  \
    if (G.vec) {                                                                                                     \
      /* with a matrix gather following, the residual kernel forks onto the side stream and runs beside it */       \
      cudaStream_t vs = G.vals ? h->stream2 : h->stream;                                                             \
      if (G.vals) {                                                                                                  \
        cudaEventRecord(h->evFork, h->stream);                                                                       \
        cudaStreamWaitEvent(h->stream2, h->evFork, 0);                                                               \
      }                                                                                                              \
      gather_vec_kernel<DIM, NN, MODE, IL><<<vecGrid, 256, 0, vs>>>(G);                                              \
      h->launches++;                                                                                                 \
      cudaEventRecord(h->evVec, vs);                                                                                 \
    }                                                                                                                \
    if (G.vals && h->gatherPull && h->csrc.p) {                                                                      \
      constexpr bool canMirror = MODE != IKB_DBC_REDUCED;                                                            \
      if (rowsAsync && NN <= 8)                                                                                      \
        launchRowsAsync<DIM, MODE, IL>(h, G, idx32, nRowNodes);                                                      \
      else if (canMirror && mirror && idx32)                                                                         \
        gather_pull_kernel<DIM, MODE, IL, true, canMirror><<<gridFor(nRange, pw), pw * 32, 0, gst>>>(      \
            G, h->cptr.p, h->csrc.p);                                                                                \
      else if (canMirror && mirror)                                                                                  \
        gather_pull_kernel<DIM, MODE, IL, false, canMirror><<<gridFor(nRange, pw), pw * 32, 0, gst>>>(     \
            G, h->cptr.p, h->csrc.p);                                                                                \
      else if (idx32)                                                                                                \
        gather_pull_kernel<DIM, MODE, IL, true, false><<<gridFor(nRange, pw), pw * 32, 0, gst>>>(          \
            G, h->cptr.p, h->csrc.p);                                                                                \
      else                                                                                                           \
        gather_pull_kernel<DIM, MODE, IL, false, false><<<gridFor(nRange, pw), pw * 32, 0, gst>>>(         \
            G, h->cptr.p, h->csrc.p);                                                                                \
      if (G.vec) cudaStreamWaitEvent(h->stream, h->evVec, 0); /* join */                                             \
    } else if (G.vals) {                                                                                             \
      e = cudaFuncSetAttribute(gather_kernel<DIM, NN, MODE, IL>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                               (int)smem);                                                                           \
      int occ = 1;                                                                                                   \
      if (e == cudaSuccess)                                                                                          \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gather_kernel<DIM, NN, MODE, IL>, warps * 32, smem); \
      const bool persist = GatherCfg<DIM, NN>::PERSIST;                                                              \
      const unsigned grid = (unsigned)(persist ? std::min<int64_t>(wantBlocks, (int64_t)148 * std::max(occ, 1) * 4) \
                                               : wantBlocks);                                                        \
      if (e == cudaSuccess) gather_kernel<DIM, NN, MODE, IL><<<grid, warps * 32, smem, h->stream>>>(G);              \
      if (G.vec) cudaStreamWaitEvent(h->stream, h->evVec, 0); /* join */                                             \
    }                                                                                                                \
  }
#define IKB_GATHER2(DIM, NN, MODE)                    \
  if (h->layout == LAYOUT_INTERLEAVED)                \
    IKB_GATHER3(DIM, NN, MODE, true)                  \
  else                                                \
    IKB_GATHER3(DIM, NN, MODE, false)
#define IKB_GATHER(DIM, NN)                           \
  if (h->dim == DIM && h->nn == NN) {                 \
    if (dbc == IKB_DBC_RAW) {                         \
      IKB_GATHER2(DIM, NN, IKB_DBC_RAW)               \
    } else if (dbc == IKB_DBC_FULL) {                 \
      IKB_GATHER2(DIM, NN, IKB_DBC_FULL)              \
    } else {                                          \
      IKB_GATHER2(DIM, NN, IKB_DBC_REDUCED)           \
    }                                                 \
  }
  IKB_GATHER(3, 8)
  IKB_GATHER(3, 27)
  IKB_GATHER(2, 4)
  IKB_GATHER(2, 9)
#undef IKB_GATHER
#undef IKB_GATHER2
#undef IKB_GATHER3
  if (e != cudaSuccess) return fail(h, IKB_ECUDA, std::string("gather launch: ") + cudaGetErrorString(e));
  if (!G.vals) h->launches--;  // only the residual kernel was launched (already counted)
  IKB_LAUNCH_CHECK(h);
  return IKB_OK;
}

// The captured solver graphs hold raw pointers and sizes of the pattern they were captured for; the caches are keyed on
// pointer values, and a freed-and-reallocated array can come back at the same address: drop them with the pattern.
void dropSolverGraphs(Handle* h) {
  if (h->cgGraph) cudaGraphExecDestroy(h->cgGraph);
  if (h->tcgGraph) cudaGraphExecDestroy(h->tcgGraph);
  if (h->peerGraph) cudaGraphExecDestroy(h->peerGraph);
  h->cgGraph = h->tcgGraph = h->peerGraph = nullptr;
  h->cgGraphKey[0] = h->tcgGraphKey[0] = h->peerGraphKey[0] = nullptr;
}

// Chunk tables of the interleaved sweep: element ranges of equal size (whole CTAs of the element kernels) and, per
// chunk, the end of the prefix of node-rows all of whose elements lie in chunks 0..c.
int ensureSweepChunks(Handle* h) {
  if (h->sweepChunksBuilt) return IKB_OK;
  h->sweepChunksBuilt = true;
  h->sweepElemEnd.clear();
  h->sweepRowEnd.clear();
  const int64_t nRowNodes = h->rowEnd - h->rowBegin;
  const int K = std::min(h->sweepChunks, 32);
  if (K < 2 || h->order != 1 || h->easM != 0 || h->form == FORM_PS || !h->gatherPull || h->pullMirror || h->pullAsync || h->fusedEnabled ||
      !h->csrc.p || !h->adjPtr.p || nRowNodes == 0 || h->nElem < (int64_t)K * 4096)
    return IKB_OK;
  DevBuf<int32_t> dMax;
  IKB_CUDA(h, dMax.alloc((size_t)nRowNodes));
  row_max_elem_kernel<<<gridFor(nRowNodes, 256), 256, 0, h->stream>>>(nRowNodes, h->adjPtr.p, h->adjCode.p, h->nn, dMax.p);
  IKB_LAUNCH_CHECK(h);
  std::vector<int32_t> maxElem((size_t)nRowNodes);
  IKB_CUDA(h, cudaMemcpyAsync(maxElem.data(), dMax.p, dMax.bytes(), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  int64_t row = 0;
  for (int c = 0; c < K; ++c) {
    const int64_t end = c + 1 == K ? h->nElem : std::min<int64_t>(h->nElem, (h->nElem * (c + 1) / K + 255) / 256 * 256);
    while (row < nRowNodes && maxElem[(size_t)row] < end) ++row;
    h->sweepElemEnd.push_back(end);
    h->sweepRowEnd.push_back(c + 1 == K ? nRowNodes : row);
  }
  for (int c = 0; c < K; ++c)
    if (!h->evChunk[c]) IKB_CUDA(h, cudaEventCreateWithFlags(&h->evChunk[c], cudaEventDisableTiming));
  if (!h->stream3) {
    int prioLo = 0, prioHi = 0;
    cudaDeviceGetStreamPriorityRange(&prioLo, &prioHi);
    IKB_CUDA(h, cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, prioLo));
    IKB_CUDA(h, cudaEventCreateWithFlags(&h->evSweepFork, cudaEventDisableTiming));
    IKB_CUDA(h, cudaEventCreateWithFlags(&h->evSweepJoin, cudaEventDisableTiming));
  }
  return IKB_OK;
}

// Interleaved sweep: element kernel of chunk c on the main stream, matrix gather of the rows completed by chunk c on
// the side stream (beside the element kernel of chunk c+1); the residual gather follows the last chunk as usual.
// Same kernels, same staged values, same summation order as the back-to-back path: only the launch shape differs.
int launchInterleaved(Handle* h, unsigned stageWhat, unsigned gatherWhat, int dbc) {
  ElemArgs A = makeElemArgs(h, stageWhat);
  const int K = (int)h->sweepElemEnd.size();
  const int nsol = (int)h->chunkElemEnd.size();
  int64_t rowDone = 0;
  for (int c = 0; c < K; ++c) {
    const int64_t b = c ? h->sweepElemEnd[c - 1] : 0, end = h->sweepElemEnd[c];
    if (h->piecesPending) {
      // a pipelined solution upload: this chunk needs the piece that holds the dofs of its last element
      int sc = 0;
      while (sc + 1 < nsol && h->chunkElemEnd[sc] < end) ++sc;
      cudaStreamWaitEvent(h->stream, h->evPiece[sc], 0);
    }
    A.elemBegin = b;
    A.elemCount = end - b;
    A.X = h->X.p + b;
    A.elemNode = h->elemNode.p + b;
    A.Lap = h->Lap.p ? h->Lap.p + b : nullptr;
    A.Kst = h->Kst.p + (size_t)b * h->npair * h->dim * h->dim;
    A.Rst = h->Rst.p + (size_t)b * h->nd;
    A.Est = h->Est.p + b;
    const cudaError_t e = launchQ1Kernel(h, A);
    h->launches++;
    if (e != cudaSuccess) return fail(h, IKB_ECUDA, std::string("element kernel: ") + cudaGetErrorString(e));
    if (h->sweepRowEnd[c] > rowDone) {
      cudaEventRecord(h->evChunk[c], h->stream);
      cudaStreamWaitEvent(h->stream3, h->evChunk[c], 0);
      const int rc = launchGather(h, IKB_MATRIX, dbc, rowDone, h->sweepRowEnd[c], h->stream3);
      if (rc) return rc;
      rowDone = h->sweepRowEnd[c];
    }
  }
  h->piecesPending = false;
  markSolutionUse(h);
  if (gatherWhat & IKB_VECTOR) {
    const int rc = launchGather(h, IKB_VECTOR, dbc);
    if (rc) return rc;
  }
  cudaEventRecord(h->evSweepJoin, h->stream3);
  cudaStreamWaitEvent(h->stream, h->evSweepJoin, 0);
  return IKB_OK;
}

int checkMaterialError(Handle* h) {
  int32_t flag = INT_MAX;
  IKB_CUDA(h, cudaMemcpyAsync(&flag, h->errFlag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (flag != INT_MAX) {
    const int32_t reset = INT_MAX;
    cudaMemcpyAsync(h->errFlag.p, &reset, sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
    if (flag == -2) {
      if (std::getenv("IKB_SWEEP_DEBUG") && h->sweepDone.p) {
        // where the two kernels stood when a wait gave up
        cudaStreamSynchronize(h->stream);
        std::vector<unsigned> done((size_t)(h->nElemGroups + h->nRowGroups));
        unsigned long long ctl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpy(done.data(), h->sweepDone.p, done.size() * sizeof(unsigned), cudaMemcpyDeviceToHost);
        cudaMemcpy(ctl, h->sweepCtl.p, sizeof(ctl), cudaMemcpyDeviceToHost);
        int64_t fe = -1, fr = -1;
        for (int64_t k = 0; k < h->nElemGroups && fe < 0; ++k) {
          const unsigned sz = (unsigned)std::min<int64_t>(SWEEP_EG, h->nElemTickets - k * SWEEP_EG);
          if (done[(size_t)k] != h->fusedEpoch * sz) fe = k;
        }
        for (int64_t k = 0; k < h->nRowGroups && fr < 0; ++k) {
          const unsigned sz = (unsigned)std::min<int64_t>(SWEEP_RG, h->nRowTickets - k * SWEEP_RG);
          if (done[(size_t)(h->nElemGroups + k)] != h->fusedEpoch * sz) fr = k;
        }
        std::fprintf(stderr,
                     "[ikb sweep] epoch %u ring %lld elems (guards %d) tickets e %u r %u | claimed e %llu (base %llu) r %llu (base %llu) "
                     "watermark %llx | first incomplete producer group %lld of %lld (count %u), consumer group %lld of %lld (count %u)\n",
                     h->fusedEpoch, (long long)h->ringElems, (int)h->fusedGuards, h->nElemTickets, h->nRowTickets, ctl[0],
                     h->elemTicketBase, ctl[1], h->rowTicketBase, ctl[2], (long long)fe, (long long)h->nElemGroups,
                     fe >= 0 ? done[(size_t)fe] : 0u, (long long)fr, (long long)h->nRowGroups,
                     fr >= 0 ? done[(size_t)(h->nElemGroups + fr)] : 0u);
      }
      // a dependency wait of the fused sweep ran into its spin limit (should not happen): results are invalid; start
      // the counters afresh and serve this handle by the two-kernel path from now on
      if (h->sweepDone.p) cudaMemsetAsync(h->sweepDone.p, 0, h->sweepDone.bytes(), h->stream);
      if (h->sweepCtl.p) cudaMemsetAsync(h->sweepCtl.p, 0, h->sweepCtl.bytes(), h->stream);
      h->fusedEpoch = 0;
      h->elemTicketBase = h->rowTicketBase = 0;
      h->fusedOk = false;
      h->stateVersion++;
      return fail(h, IKB_ECUDA, "fused sweep: dependency wait timed out");
    }
    // the reference aborts here (materials/materialhelpers.hh:120-126)
    return fail(h, IKB_EMATERIAL,
                "Determinant of right Cauchy Green tensor C must be greater than zero (first failing element " +
                    std::to_string(flag) + ")");
  }
  return IKB_OK;
}

int deviceDot(Handle* h, int mode, const double* x, const double* y, int64_t n, double* outDev, double scale,
              const double* addend) {
  double* partial = h->scratch.p;
  if (mode == 0)
    reduce_stage1<0><<<RED_BLOCKS, 256, 0, h->stream>>>(x, y, n, partial);
  else if (mode == 1)
    reduce_stage1<1><<<RED_BLOCKS, 256, 0, h->stream>>>(x, y, n, partial);
  else
    reduce_stage1<2><<<RED_BLOCKS, 256, 0, h->stream>>>(x, y, n, partial);
  IKB_LAUNCH_CHECK(h);
  reduce_stage2<<<1, 256, 0, h->stream>>>(partial, RED_BLOCKS, outDev, scale, addend);
  IKB_LAUNCH_CHECK(h);
  return IKB_OK;
}

int launchSpmv(Handle* h, int dbc, const double* x, double* y) {
  const int tpb = 256;
  if (dbc == IKB_DBC_REDUCED) {
    if (h->nRed == 0) return IKB_OK;
    spmv_csr_kernel<8><<<gridFor(h->nRed * 8, tpb), tpb, 0, h->stream>>>(h->redOuter.p, h->redInner.p,
                                                                         h->vals[dbc].p, h->nRed, x, y);
  } else {
    const PatternView P = h->view();
    const int64_t rows = P.nRowNodes * h->dim;
    if (rows == 0) return IKB_OK;
    const unsigned grid = (unsigned)std::min<int64_t>((P.nRowNodes + 7) / 8, 148 * 16);
    if (h->dim == 3)
      spmv_node_dot_kernel<3><<<grid, tpb, 0, h->stream>>>(P, h->vals[dbc].p, x, y, nullptr, nullptr, nullptr);
    else
      spmv_node_dot_kernel<2><<<grid, tpb, 0, h->stream>>>(P, h->vals[dbc].p, x, y, nullptr, nullptr, nullptr);
  }
  IKB_LAUNCH_CHECK(h);
  return IKB_OK;
}

int haloExchange(Handle* h, double* v) {
  if (!h->comm || h->nranks == 1) return IKB_OK;
  const int D = h->dim;
  const int64_t* me = h->peerRanges.data() + 4 * h->rank;
  int rc = nccl().groupStart();
  for (int s = 0; s < h->nranks && rc == 0; ++s) {
    if (s == h->rank) continue;
    const int64_t* pr = h->peerRanges.data() + 4 * s;
    int64_t sb, se, rb, re;
    haloIntervals(me[0], me[1], me[2], me[3], pr[0], pr[1], pr[2], pr[3], sb, se, rb, re);
    if (se > sb) rc = nccl().send(v + D * sb, (size_t)D * (se - sb), NCCL_FLOAT64, s, h->comm, h->stream);
    if (rc == 0 && re > rb) rc = nccl().recv(v + D * rb, (size_t)D * (re - rb), NCCL_FLOAT64, s, h->comm, h->stream);
  }
  const int rc2 = nccl().groupEnd();
  if (rc != 0 || rc2 != 0) return fail(h, IKB_ENCCL, std::string("halo exchange: ") + nccl().errorString(rc ? rc : rc2));
  h->launches++;
  return IKB_OK;
}

int allReduceSum(Handle* h, double* devScalars, int count) {
  if (!h->comm || h->nranks == 1) return IKB_OK;
  const int rc = nccl().allReduce(devScalars, devScalars, (size_t)count, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
  if (rc != 0) return fail(h, IKB_ENCCL, std::string("ncclAllReduce: ") + nccl().errorString(rc));
  h->launches++;
  return IKB_OK;
}

// Jacobi-PCG over row blocks: halo exchange of the search direction before every SpMV, one all-reduce
// for p.q and one for (r.z, r.r) per iteration.  Full mode only.
int distPcg(Handle* h, const double* rhsHost, double* xHost, double relTol, int maxIt, int* itersOut,
            double* relResOut) {
  const int dbc = IKB_DBC_FULL;
  const int64_t n = h->nRowsLocal();
  const int64_t off = h->rowBegin * h->dim;
  const int tpb = 256;
  for (auto* b : {&h->cgR, &h->cgZ, &h->cgQ, &h->cgX, &h->cgDinv})
    if (b->n < (size_t)std::max<int64_t>(n, 1)) IKB_CUDA(h, b->alloc((size_t)std::max<int64_t>(n, 1)));
  if (h->cgPglob.n < (size_t)h->nDof) {
    IKB_CUDA(h, h->cgPglob.alloc((size_t)h->nDof));
    IKB_CUDA(h, cudaMemsetAsync(h->cgPglob.p, 0, h->cgPglob.bytes(), h->stream));
  }
  if (h->Corr.n < (size_t)h->nDof) IKB_CUDA(h, h->Corr.alloc((size_t)h->nDof));
  IKB_CUDA(h, cudaMemsetAsync(h->Corr.p, 0, h->Corr.bytes(), h->stream));
  double* p = h->cgPglob.p + off;
  int rc;
  {
    // Agreement before the first data collective: a rank whose state is not ready must not leave on its own -- the
    // others would wait for it in the all-reduce below for ever.  Every rank contributes a flag and all of them return.
    double bad = (h->valsVersion[dbc] != h->stateVersion || (!rhsHost && h->vecVersion[dbc] != h->stateVersion)) ? 1.0 : 0.0;
    IKB_CUDA(h, cudaMemcpyAsync(h->cgScal.p + 5, &bad, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if ((rc = allReduceSum(h, h->cgScal.p + 5, 1))) return rc;
    IKB_CUDA(h, cudaMemcpyAsync(&bad, h->cgScal.p + 5, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    IKB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (bad > 0.0) return fail(h, IKB_ESTATE, "matrix or resident residual not assembled for the current state (on this or another rank)");
  }
  if (rhsHost) {
    IKB_CUDA(h, cudaMemcpyAsync(h->cgR.p, rhsHost, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  } else {
    vec_scale_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, -1.0, h->vec[dbc].p, h->cgR.p);
    IKB_LAUNCH_CHECK(h);
  }
  IKB_CUDA(h, cudaMemsetAsync(h->cgX.p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(double), h->stream));
  if (h->nBlocks) {
    if (h->dim == 3)
      diag_inv_block_kernel<3><<<gridFor(h->nBlocks, tpb), tpb, 0, h->stream>>>(h->view(), h->vals[dbc].p, h->cgDinv.p);
    else
      diag_inv_block_kernel<2><<<gridFor(h->nBlocks, tpb), tpb, 0, h->stream>>>(h->view(), h->vals[dbc].p, h->cgDinv.p);
    IKB_LAUNCH_CHECK(h);
  }
  double* scal = h->cgScal.p;
  vec_mul_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, h->cgDinv.p, h->cgR.p, h->cgZ.p);
  IKB_LAUNCH_CHECK(h);
  IKB_CUDA(h, cudaMemcpyAsync(p, h->cgZ.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if ((rc = deviceDot(h, 1, h->cgR.p, h->cgZ.p, n, scal + 0, 1.0, nullptr))) return rc;
  if ((rc = deviceDot(h, 2, h->cgR.p, nullptr, n, scal + 4, 1.0, nullptr))) return rc;
  if ((rc = allReduceSum(h, scal + 0, 1))) return rc;
  if ((rc = allReduceSum(h, scal + 4, 1))) return rc;
  double bb = 0.0;
  IKB_CUDA(h, cudaMemcpyAsync(&bb, scal + 4, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  int it = 0;
  double rr = bb;
  const double threshold = std::max(relTol * relTol * bb, 1e-300);
  (void)threshold;
  if (bb > 0.0 && h->peerReady) {
    // Peer-memory transport (ikb_pcg_peer.cuh): halo stores and partial sums go straight into the neighbours' memory
    // from inside the three kernels of an iteration; batches are replayed from a CUDA graph.
    if ((rc = haloExchange(h, h->cgPglob.p))) return rc;  // initial direction (NCCL, once)
    if (!h->peerState.p) IKB_CUDA(h, h->peerState.alloc(256));
    PeerState* st = reinterpret_cast<PeerState*>(h->peerState.p);
    const PeerComm& PC = *reinterpret_cast<const PeerComm*>(h->peerCommHost);
    h->peerEpoch++;
    peer_init_kernel<<<1, 1, 0, h->stream>>>(st, scal + 0, scal + 4, relTol, maxIt, h->peerEpoch);
    IKB_LAUNCH_CHECK(h);
    const PatternView P = h->view();
    if (h->peerFirstInterior < 0) {
      DevBuf<int32_t> fe;
      IKB_CUDA(h, fe.alloc(2));
      const int32_t init[2] = {0, (int32_t)P.nRowNodes};
      IKB_CUDA(h, cudaMemcpyAsync(fe.p, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
      peer_boundary_rows_kernel<<<gridFor(P.nRowNodes, tpb), tpb, 0, h->stream>>>(P, h->rowEnd, fe.p, fe.p + 1);
      IKB_LAUNCH_CHECK(h);
      int32_t out[2];
      IKB_CUDA(h, cudaMemcpyAsync(out, fe.p, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
      IKB_CUDA(h, cudaStreamSynchronize(h->stream));
      if (out[0] > out[1]) out[0] = out[1] = 0;  // every row may touch the halo: no interior phase
      h->peerFirstInterior = out[0];
      h->peerEndInterior = out[1];
      fe.release();
    }
    double* pqPartial = h->scratch.p;
    double* rzPartial = h->scratch.p + MAX_SPMV_BLOCKS;
    PeerState* hs = reinterpret_cast<PeerState*>(h->hostScal);
    const int batch = 32;
    auto enqueueBatch = [&]() {
      for (int b = 0; b < batch; ++b) {
        if (h->dim == 3)
          peer_spmv_kernel<3><<<h->spmvBlocks, tpb, 0, h->stream>>>(P, h->vals[dbc].p, h->cgPglob.p, h->cgQ.p, p, pqPartial,
                                                                     st, PC, h->peerWin.p, h->peerFirstInterior,
                                                                     h->peerEndInterior);
        else
          peer_spmv_kernel<2><<<h->spmvBlocks, tpb, 0, h->stream>>>(P, h->vals[dbc].p, h->cgPglob.p, h->cgQ.p, p, pqPartial,
                                                                     st, PC, h->peerWin.p, h->peerFirstInterior,
                                                                     h->peerEndInterior);
        peer_update_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, st, PC, h->peerWin.p, p, h->cgQ.p, h->cgDinv.p, h->cgX.p,
                                                              h->cgR.p, h->cgZ.p, rzPartial);
        peer_direction_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, off, st, PC, h->peerWin.p, h->cgZ.p, h->cgPglob.p);
      }
    };
    if (!h->peerGraph || h->peerGraphKey[0] != h->vals[dbc].p || h->peerGraphKey[1] != h->cgPglob.p) {
      if (h->peerGraph) cudaGraphExecDestroy(h->peerGraph);
      h->peerGraph = nullptr;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        enqueueBatch();
        if (cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && graph) {
          if (cudaGraphInstantiate(&h->peerGraph, graph, 0) != cudaSuccess) h->peerGraph = nullptr;
          cudaGraphDestroy(graph);
        }
      }
      cudaGetLastError();
      h->peerGraphKey[0] = h->vals[dbc].p;
      h->peerGraphKey[1] = h->cgPglob.p;
    }
    static_assert(sizeof(PeerState) <= 16 * sizeof(double), "hostScal holds the state");
    while (true) {
      if (h->peerGraph) {
        IKB_CUDA(h, cudaGraphLaunch(h->peerGraph, h->stream));
      } else {
        enqueueBatch();
      }
      h->launches += 3 * batch;
      IKB_CUDA(h, cudaGetLastError());
      IKB_CUDA(h, cudaMemcpyAsync(hs, st, sizeof(PeerState), cudaMemcpyDeviceToHost, h->stream));
      IKB_CUDA(h, cudaStreamSynchronize(h->stream));
      if (hs->timeout) return fail(h, IKB_ENCCL, "distributed PCG: a peer did not arrive (peer-memory wait timed out)");
      if (hs->done || hs->iter >= maxIt) break;
    }
    it = hs->iter;
    rr = hs->rr;
    if (hs->done == 2) return fail(h, IKB_ECUDA, "PCG produced NaN (matrix not positive definite?)");
  } else if (bb > 0.0) {
    // No host round trip per iteration: convergence is decided on the device from the all-reduced |r|^2 (identical
    // on every rank, so all ranks stop at the same iteration); NCCL calls are stream ordered.  The host looks at the
    // 48-byte state once per batch; iterations enqueued after convergence are no-ops apart from the (harmless)
    // collectives.
    if (!h->cgState.p) IKB_CUDA(h, h->cgState.alloc(128));
    CgState* st = reinterpret_cast<CgState*>(h->cgState.p);
    unsigned int* arrive = reinterpret_cast<unsigned int*>(h->cgState.p + 96);
    cg2_init_kernel<<<1, 1, 0, h->stream>>>(st, scal + 0, scal + 4, relTol, maxIt, arrive);
    IKB_LAUNCH_CHECK(h);
    const PatternView P = h->view();
    double* pqPartial = h->scratch.p;
    double* rzPartial = h->scratch.p + MAX_SPMV_BLOCKS;
    CgState* hs = reinterpret_cast<CgState*>(h->hostScal);
    const int batch = 16;
    auto enqueueBatch = [&]() -> int {
      for (int b = 0; b < batch; ++b) {
        int r2;
        if ((r2 = haloExchange(h, h->cgPglob.p))) return r2;
        if (h->dim == 3)
          spmv_node_dot_kernel<3><<<h->spmvBlocks, tpb, 0, h->stream>>>(P, h->vals[dbc].p, h->cgPglob.p, h->cgQ.p, p,
                                                                        pqPartial, st);
        else
          spmv_node_dot_kernel<2><<<h->spmvBlocks, tpb, 0, h->stream>>>(P, h->vals[dbc].p, h->cgPglob.p, h->cgQ.p, p,
                                                                        pqPartial, st);
        reduce_stage2<<<1, tpb, 0, h->stream>>>(pqPartial, h->spmvBlocks, scal + 1, 1.0, nullptr);
        if ((r2 = allReduceSum(h, scal + 1, 1))) return r2;
        // the all-reduced scalars are passed as one-entry "partial" arrays
        cg2_update_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, st, scal + 1, 1, p, h->cgQ.p, h->cgDinv.p, h->cgX.p,
                                                             h->cgR.p, h->cgZ.p, rzPartial);
        cg_fold2_kernel<<<1, tpb, 0, h->stream>>>(rzPartial, RED_BLOCKS, scal, nullptr);
        if ((r2 = allReduceSum(h, scal + 2, 2))) return r2;
        cg2_direction_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, st, scal + 2, 1, h->cgZ.p, p, arrive);
      }
      return IKB_OK;
    };
    // (capturing the NCCL operations of a batch into a CUDA graph hung on the 2-GPU box and is not used)
    while (true) {
      if ((rc = enqueueBatch())) return rc;
      h->launches += 8 * batch;
      IKB_CUDA(h, cudaGetLastError());
      IKB_CUDA(h, cudaMemcpyAsync(hs, st, sizeof(CgState), cudaMemcpyDeviceToHost, h->stream));
      IKB_CUDA(h, cudaStreamSynchronize(h->stream));
      if (hs->done || hs->iter >= maxIt) break;
    }
    it = hs->iter;
    rr = hs->rr;
    if (hs->done == 2) return fail(h, IKB_ECUDA, "PCG produced NaN (matrix not positive definite?)");
  }
  if (itersOut) *itersOut = it;
  if (relResOut) *relResOut = bb > 0.0 ? std::sqrt(rr / bb) : 0.0;
  IKB_CUDA(h, cudaMemcpyAsync(h->Corr.p + off, h->cgX.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if ((rc = haloExchange(h, h->Corr.p))) return rc;
  if (xHost) IKB_CUDA(h, cudaMemcpyAsync(xHost, h->cgX.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  return IKB_OK;
}

}  // namespace

extern "C" {

int ikb_create(ikb_handle* out, const ikb_desc* desc) {
  if (!out || !desc) return IKB_EINVAL;
  *out = nullptr;
  if (desc->abi_version != IKB_ABI_VERSION) return IKB_EINVAL;
  if (desc->dim != 2 && desc->dim != 3) return IKB_EINVAL;
  if (desc->order != 1 && desc->order != 2) return IKB_EINVAL;
  if (desc->n_elem < 0 || desc->n_dof <= 0 || desc->n_dof % desc->dim != 0) return IKB_EINVAL;
  if (desc->n_dof >= (int64_t)INT32_MAX) return IKB_EINVAL;
  int form;
  if (desc->strain == IKB_STRAIN_LINEAR && desc->material == IKB_MAT_LINEAR_ELASTICITY)
    form = FORM_LE;
  else if (desc->strain == IKB_STRAIN_GREEN_LAGRANGE && desc->material == IKB_MAT_SVK)
    form = FORM_SVK;
  else if (desc->strain == IKB_STRAIN_GREEN_LAGRANGE && desc->material == IKB_MAT_NEOHOOKE)
    form = FORM_NH;
  else if (desc->strain == IKB_STRAIN_GREEN_LAGRANGE &&
           (desc->material == IKB_MAT_BLATZKO || desc->material == IKB_MAT_HYPERELASTIC))
    form = FORM_PS;
  else
    return IKB_EINVAL;  // the reference statically rejects these strain/material pairings too
  // principal-stretch laws: Q1 family (plain, strain- and displacement-gradient-enhanced), 3D or plane strain
  if (form == FORM_PS && (desc->order != 1 || desc->plane_strain == IKB_REDUCE_PLANE_STRESS)) return IKB_ENOTIMPL;
  if (desc->dim == 2 && desc->plane_strain != IKB_REDUCE_PLANE_STRAIN && desc->plane_strain != IKB_REDUCE_PLANE_STRESS)
    return IKB_EINVAL;  // 2D needs a reduced material
  if (desc->dim == 3 && desc->plane_strain) return IKB_EINVAL;
  const int m = desc->eas_m;
  const bool easOk = m == 0 || (desc->order == 1 && ((desc->dim == 2 && (m == 4 || m == 5 || m == 7)) ||
                                                     (desc->dim == 3 && (m == 9 || m == 21))));
  if (!easOk) return IKB_ENOTIMPL;  // Dune::NotImplemented in the reference (enhancedassumedstrains.hh:250-256)
  if (desc->eas_function < IKB_EAS_STRAIN || desc->eas_function > IKB_EAS_DISPLACEMENT_GRADIENT_TRANSPOSED) return IKB_EINVAL;
  if (m && desc->eas_function != IKB_EAS_STRAIN) {
    // H4 / H9 on the nonlinear element (enhancedassumedstrains.hh:85-90; easvariants.hh); plane stress is served for
    // the strain enhancements only
    if (form == FORM_LE) return IKB_EINVAL;
    if (m != desc->dim * desc->dim || desc->plane_strain == IKB_REDUCE_PLANE_STRESS) return IKB_ENOTIMPL;
  }

  int dev = desc->device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return IKB_ECUDA;
  if (cudaSetDevice(dev) != cudaSuccess) return IKB_ECUDA;
  Handle* h = new Handle();
  h->desc = *desc;
  h->device = dev;
  h->dim = desc->dim;
  h->order = desc->order;
  h->nn = 1;
  for (int k = 0; k < h->dim; ++k) h->nn *= (h->order + 1);
  h->nd = h->nn * h->dim;
  h->nc = 1 << h->dim;
  h->npair = h->nn * (h->nn + 1) / 2;
  h->form = form;
  if (desc->material == IKB_MAT_BLATZKO) {
    h->ps.dev = IKB_DEV_BLATZKO;
    h->ps.par[0] = desc->mu;
    h->psSet = true;
  }
  h->easM = m;
  h->easFunction = m ? desc->eas_function : IKB_EAS_STRAIN;
  h->nElem = desc->n_elem;
  h->nDof = desc->n_dof;
  h->nNodes = desc->n_dof / desc->dim;
  h->rowBegin = 0;
  h->rowEnd = h->nNodes;
  if ((double)h->nElem * h->npair >= 2147483647.0) {
    delete h;
    return IKB_EINVAL;
  }
  if (const char* gm = std::getenv("IKB_GATHER")) h->gatherPull = std::string(gm) != "tile";
  if (const char* em = std::getenv("IKB_ELEM")) h->elemMma = std::string(em) != "fma";
  if (const char* mb = std::getenv("IKB_H8_MINB")) h->h8MinBlocks = std::atoi(mb);
  if (const char* fu = std::getenv("IKB_FUSED")) h->fusedEnabled = std::atoi(fu) != 0;
  h->sweepDebug = std::getenv("IKB_SWEEP_DEBUG") != nullptr;
  if (const char* pp = std::getenv("IKB_PCG_PEER")) h->peerEnabled = std::atoi(pp) != 0;
  if (const char* sm = std::getenv("IKB_SWEEP_MARGIN")) h->sweepMargin = std::atoi(sm);  // test hook: 0 = smallest legal ring
  if (const char* sc = std::getenv("IKB_SWEEP_CTAS")) {
    int a = 0, b = 0;
    if (std::sscanf(sc, "%d,%d", &a, &b) == 2 && a >= 1 && b >= 1) {
      h->sweepElemCtas = a;
      h->sweepRowCtas = b;
    }
  }
  if (const char* rm = std::getenv("IKB_RING_MB")) h->fusedRingMB = std::max(0.01, std::atof(rm));
  // test hooks for the rarely taken paths of the pull gather (long contribution lists, > 2^31 staged offsets)
  if (const char* sm = std::getenv("IKB_PULL_STAGE_MAX")) h->pullStageMax = std::max(std::atoi(sm), -1);
  if (const char* i64 = std::getenv("IKB_PULL_IDX64")) h->pullIdx64 = std::atoi(i64) != 0;
  if (const char* pm = std::getenv("IKB_PULL_MIRROR")) h->pullMirror = std::atoi(pm) != 0;
  if (const char* w = std::getenv("IKB_PULL_WARPS")) h->pullWarps = std::min(std::max(std::atoi(w), 1), PULL_WARPS_MAX);
  if (const char* pa = std::getenv("IKB_PULL_ASYNC")) h->pullAsync = std::atoi(pa) != 0;
  if (const char* ch = std::getenv("IKB_CHUNKS")) h->sweepChunks = std::atoi(ch);
  if (const char* sb = std::getenv("IKB_SPMV_BLOCKS")) h->spmvBlocks = std::min(std::max(std::atoi(sb), 1), MAX_SPMV_BLOCKS);
  int prioLo = 0, prioHi = 0;
  cudaDeviceGetStreamPriorityRange(&prioLo, &prioHi);  // the side stream outranks the main one
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prioHi) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->evVec, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->evUse, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->evPiece[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->evPiece[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->evPiece[2], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->evPiece[3], cudaEventDisableTiming) != cudaSuccess ||
      h->errFlag.alloc(1) != cudaSuccess || h->scratch.alloc(MAX_SPMV_BLOCKS + 2 * RED_BLOCKS + 16) != cudaSuccess ||
      h->cgScal.alloc(16) != cudaSuccess || cudaMallocHost(reinterpret_cast<void**>(&h->hostScal), 16 * sizeof(double)) != cudaSuccess) {
    delete h;
    return IKB_ECUDA;
  }
  const int32_t reset = INT_MAX;
  cudaMemcpy(h->errFlag.p, &reset, sizeof(int32_t), cudaMemcpyHostToDevice);
  cudaMemset(h->cgScal.p, 0, h->cgScal.bytes());
  *out = reinterpret_cast<ikb_handle>(h);
  return IKB_OK;
}

int ikb_destroy(ikb_handle hh) {
  Handle* h = H(hh);
  if (!h) return IKB_EINVAL;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->Lap.release();
  for (auto* b : {&h->X, &h->U, &h->Fext, &h->Corr, &h->Kst, &h->Rst, &h->Est, &h->scratch, &h->alpha, &h->cgR, &h->cgZ,
                  &h->cgP, &h->cgQ, &h->cgX, &h->cgDinv, &h->cgB, &h->cgScal})
    b->release();
  for (int i = 0; i < 3; ++i) {
    h->vals[i].release();
    h->vec[i].release();
  }
  h->elemNode.release();
  h->flags.release();
  h->nbrPtr.release();
  h->nbrIdx.release();
  h->nbrRow.release();
  h->cptr.release();
  h->csrc.release();
  h->adjPtr.release();
  h->adjCode.release();
  h->slotTab.release();
  h->rowSlow.release();
  h->rowSlowValid = false;
  h->sweepChunksBuilt = false;
  dropSolverGraphs(h);
  h->cbelow.release();
  h->freeCnt.release();
  h->freeTot.release();
  h->redRowStart.release();
  h->redInner.release();
  h->redOuter.release();
  h->errFlag.release();
  for (void* pp : h->peerOpened) cudaIpcCloseMemHandle(pp);
  if (h->peerGraph) cudaGraphExecDestroy(h->peerGraph);
  delete reinterpret_cast<PeerComm*>(h->peerCommHost);
  h->peerWin.release();
  h->peerState.release();
  if (h->comm) nccl().commDestroy(h->comm);
  h->cgPglob.release();
  h->cgState.release();
  h->rowDiag.release();
  h->rowLowEnd.release();
  h->mirrorBlk.release();
  h->ring.release();
  h->rring.release();
  h->csrcRing.release();
  h->adjRing.release();
  h->sweepGuard.release();
  h->sweepRowWait.release();
  h->sweepCtl.release();
  h->sweepDone.release();
  for (auto& ev : h->evChunk)
    if (ev) cudaEventDestroy(ev);
  if (h->stream3) cudaStreamDestroy(h->stream3);
  if (h->evSweepFork) cudaEventDestroy(h->evSweepFork);
  if (h->evSweepJoin) cudaEventDestroy(h->evSweepJoin);
  h->gatherTab.release();
  if (h->cgGraph) cudaGraphExecDestroy(h->cgGraph);
  if (h->tcgGraph) cudaGraphExecDestroy(h->tcgGraph);
  h->tcgState.release();
  h->T0inv.release();
  if (h->hostScal) cudaFreeHost(h->hostScal);
  if (h->evVec) cudaEventDestroy(h->evVec);
  if (h->evFork) cudaEventDestroy(h->evFork);
  if (h->evUse) cudaEventDestroy(h->evUse);
  for (cudaEvent_t ev : h->evPiece)
    if (ev) cudaEventDestroy(ev);
  cudaStreamDestroy(h->stream);
  cudaStreamDestroy(h->stream2);
  delete h;
  return IKB_OK;
}

int ikb_last_error(ikb_handle hh, char* buf, size_t len) {
  Handle* h = H(hh);
  if (!h || !buf || len == 0) return IKB_EINVAL;
  std::strncpy(buf, h->lastError.c_str(), len - 1);
  buf[len - 1] = 0;
  return IKB_OK;
}

int ikb_upload_mesh(ikb_handle hh, const double* corner, const int64_t* elemDofs) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!corner || !elemDofs) return fail(h, IKB_EINVAL, "null mesh arrays");
  const int D = h->dim, nn = h->nn, nc = h->nc, nd = h->nd;
  const int64_t ne = h->nElem, nNodes = h->nNodes;
  if (ne == 0) {
    h->meshUploaded = true;
    return IKB_OK;
  }
  // detect the flat power-basis layout from the first node, then verify every node
  int layout = -1;
  {
    const int64_t* d0 = elemDofs;
    bool inter = d0[0] % D == 0, lex = d0[0] < nNodes;
    for (int c = 1; c < D; ++c) {
      inter = inter && d0[c] == d0[0] + c;
      lex = lex && d0[c] == d0[0] + (int64_t)c * nNodes;
    }
    if (inter)
      layout = LAYOUT_INTERLEAVED;
    else if (lex)
      layout = LAYOUT_LEXICOGRAPHIC;
  }
  if (layout < 0)
    return fail(h, IKB_ENOTIMPL, "element dofs are neither FlatInterleaved nor FlatLexicographic per Lagrange node");
  std::vector<int32_t> en((size_t)nn * ne);
  bool ok = true;
  for (int64_t e = 0; e < ne && ok; ++e)
    for (int a = 0; a < nn; ++a) {
      const int64_t* d = elemDofs + (size_t)e * nd + (size_t)a * D;
      int64_t node;
      if (layout == LAYOUT_INTERLEAVED) {
        node = d[0] / D;
        for (int c = 0; c < D; ++c) ok = ok && d[c] == node * D + c;
      } else {
        node = d[0];
        for (int c = 0; c < D; ++c) ok = ok && d[c] == node + (int64_t)c * nNodes;
      }
      ok = ok && node >= 0 && node < nNodes;
      en[(size_t)a * ne + e] = (int32_t)node;
    }
  if (!ok) return fail(h, IKB_EINVAL, "inconsistent element dof indices");
  {
    int32_t mn = INT32_MAX, mx = -1;
    for (int32_t v : en) {
      mn = std::min(mn, v);
      mx = std::max(mx, v);
    }
    h->colBegin = mn;
    h->colEnd = (int64_t)mx + 1;
  }
  // chunk table for the pipelined solution upload (Q1, interleaved dofs, enough elements to be worth it)
  h->chunkElemEnd.clear();
  h->chunkDofEnd.clear();
  h->piecesPending = false;
  if (h->order == 1 && h->easM == 0 && h->form != FORM_PS && layout == LAYOUT_INTERLEAVED && ne >= (int64_t)Handle::SOL_CHUNKS * 8192) {
    int32_t runMax = -1;
    int64_t e = 0;
    for (int c = 0; c < Handle::SOL_CHUNKS; ++c) {
      int64_t end = c + 1 == Handle::SOL_CHUNKS ? ne : std::min<int64_t>(ne, ((ne * (c + 1) / Handle::SOL_CHUNKS) + 255) / 256 * 256);
      for (; e < end; ++e)
        for (int a = 0; a < nn; ++a) runMax = std::max(runMax, en[(size_t)a * ne + e]);
      h->chunkElemEnd.push_back(end);
      h->chunkDofEnd.push_back(((int64_t)runMax + 1) * D);
    }
  }
  // Corner coordinates are kept RELATIVE to corner 0 of their element: every kernel only forms the Jacobian
  // sum_c dN_c x_c (the shape-function derivatives sum to zero, so the shift changes nothing mathematically), and with
  // absolute coordinates that sum cancels |x|-sized terms down to h-sized ones -- a relative error of eps*|x|/h in J
  // (6e-14 on the 256^3 unit cube).  The subtraction itself is correctly rounded, error <= eps*h.
  std::vector<double> xs((size_t)nc * D * ne);
  for (int64_t e = 0; e < ne; ++e)
    for (int q = 0; q < nc * D; ++q)
      xs[(size_t)q * ne + e] = corner[(size_t)e * nc * D + q] - corner[(size_t)e * nc * D + (q % D)];
  h->layout = layout;
  IKB_CUDA(h, h->elemNode.alloc(en.size()));
  IKB_CUDA(h, h->X.alloc(xs.size()));
  IKB_CUDA(h, cudaMemcpyAsync(h->elemNode.p, en.data(), h->elemNode.bytes(), cudaMemcpyHostToDevice, h->stream));
  IKB_CUDA(h, cudaMemcpyAsync(h->X.p, xs.data(), h->X.bytes(), cudaMemcpyHostToDevice, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->meshUploaded = true;
  h->patternBuilt = false;
  h->reducedBuilt = false;
  h->rowSlowValid = false;
  h->sweepChunksBuilt = false;
  dropSolverGraphs(h);
  h->fusedTried = h->fusedOk = false;
  h->stateVersion++;
  h->Lap.release();
  if (h->order == 1 && h->easM == 0 && h->form != FORM_SVK && h->form != FORM_PS) {
    IKB_CUDA(h, h->Lap.alloc((size_t)h->npair * ne));
    if (D == 3)
      lap_q1_kernel<3><<<gridFor(ne, 128), 128, 0, h->stream>>>(h->X.p, ne, h->Lap.p);
    else
      lap_q1_kernel<2><<<gridFor(ne, 128), 128, 0, h->stream>>>(h->X.p, ne, h->Lap.p);
    IKB_LAUNCH_CHECK(h);
  }
  if (h->easM) {
    IKB_CUDA(h, h->alpha.alloc((size_t)ne * h->easM));
    IKB_CUDA(h, cudaMemsetAsync(h->alpha.p, 0, h->alpha.bytes(), h->stream));  // initializeState (:373-376)
    // EX ctor: T0InverseTransformed = (transformationMatrix(geo, center) * detJ0)^-1
    // (easvariants/helperfunctions.hh:18-25, utils/tensorutils.hh:408-465); one-time, on the host
    const int S = D * (D + 1) / 2;
    std::vector<double> t0((size_t)S * S * ne);
    for (int64_t e = 0; e < ne; ++e) {
      double J[3][3] = {{0}};
      for (int c = 0; c < nc; ++c)
        for (int i = 0; i < D; ++i) {
          double dn = ((c >> i) & 1) ? 1.0 : -1.0;
          for (int k = 0; k < D; ++k)
            if (k != i) dn *= 0.5;
          for (int k = 0; k < D; ++k) J[i][k] += dn * (corner[(size_t)e * nc * D + c * D + k] - corner[(size_t)e * nc * D + k]);
        }
      double T[6][6], Ti[6][6];
      double det;
      if (D == 2) {
        det = std::fabs(J[0][0] * J[1][1] - J[0][1] * J[1][0]);
        const double J11 = J[0][0], J12 = J[0][1], J21 = J[1][0], J22 = J[1][1];
        const double t[3][3] = {{J11 * J11, J12 * J12, J11 * J12},
                                {J21 * J21, J22 * J22, J21 * J22},
                                {2 * J11 * J21, 2 * J12 * J22, J21 * J12 + J11 * J22}};
        for (int p = 0; p < 3; ++p)
          for (int q = 0; q < 3; ++q) T[p][q] = t[p][q] * det;
      } else {
        const double J11 = J[0][0], J12 = J[0][1], J13 = J[0][2], J21 = J[1][0], J22 = J[1][1], J23 = J[1][2],
                     J31 = J[2][0], J32 = J[2][1], J33 = J[2][2];
        det = std::fabs(J11 * (J22 * J33 - J23 * J32) - J12 * (J21 * J33 - J23 * J31) + J13 * (J21 * J32 - J22 * J31));
        const double t[6][6] = {
            {J11 * J11, J12 * J12, J13 * J13, J12 * J13, J11 * J13, J11 * J12},
            {J21 * J21, J22 * J22, J23 * J23, J22 * J23, J21 * J23, J21 * J22},
            {J31 * J31, J32 * J32, J33 * J33, J32 * J33, J31 * J33, J31 * J32},
            {2 * J21 * J31, 2 * J22 * J32, 2 * J23 * J33, J22 * J33 + J32 * J23, J31 * J23 + J21 * J33,
             J21 * J32 + J31 * J22},
            {2 * J11 * J31, 2 * J12 * J32, 2 * J13 * J33, J12 * J33 + J32 * J13, J11 * J33 + J31 * J13,
             J11 * J32 + J31 * J12},
            {2 * J11 * J21, 2 * J12 * J22, 2 * J13 * J23, J12 * J23 + J22 * J13, J11 * J23 + J21 * J13,
             J11 * J22 + J12 * J21}};
        for (int p = 0; p < 6; ++p)
          for (int q = 0; q < 6; ++q) T[p][q] = t[p][q] * det;
      }
      // Gauss-Jordan with partial pivoting
      for (int p = 0; p < S; ++p)
        for (int q = 0; q < S; ++q) Ti[p][q] = p == q ? 1.0 : 0.0;
      for (int col = 0; col < S; ++col) {
        int piv = col;
        for (int r = col + 1; r < S; ++r)
          if (std::fabs(T[r][col]) > std::fabs(T[piv][col])) piv = r;
        if (T[piv][col] == 0.0) return fail(h, IKB_EINVAL, "singular EAS transformation matrix (degenerate element)");
        if (piv != col)
          for (int q = 0; q < S; ++q) {
            std::swap(T[piv][q], T[col][q]);
            std::swap(Ti[piv][q], Ti[col][q]);
          }
        const double ip = 1.0 / T[col][col];
        for (int q = 0; q < S; ++q) {
          T[col][q] *= ip;
          Ti[col][q] *= ip;
        }
        for (int r = 0; r < S; ++r) {
          if (r == col) continue;
          const double f = T[r][col];
          if (f == 0.0) continue;
          for (int q = 0; q < S; ++q) {
            T[r][q] -= f * T[col][q];
            Ti[r][q] -= f * Ti[col][q];
          }
        }
      }
      for (int p = 0; p < S; ++p)
        for (int q = 0; q < S; ++q) t0[(size_t)(p * S + q) * ne + e] = Ti[p][q];
    }
    IKB_CUDA(h, h->T0inv.alloc(t0.size()));
    IKB_CUDA(h, cudaMemcpy(h->T0inv.p, t0.data(), h->T0inv.bytes(), cudaMemcpyHostToDevice));
  }
  return IKB_OK;
}

int ikb_upload_dirichlet(ikb_handle hh, const uint8_t* flags) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!flags) return fail(h, IKB_EINVAL, "null flags");
  IKB_CUDA(h, h->flags.alloc((size_t)h->nDof));
  IKB_CUDA(h, cudaMemcpyAsync(h->flags.p, flags, h->flags.bytes(), cudaMemcpyHostToDevice, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->hasFlags = true;
  h->reducedBuilt = false;
  h->rowSlowValid = false;
  h->sweepChunksBuilt = false;
  dropSolverGraphs(h);
  h->vals[IKB_DBC_REDUCED].release();
  h->vec[IKB_DBC_REDUCED].release();
  h->stateVersion++;
  return IKB_OK;
}

int ikb_set_row_ownership(ikb_handle hh, int64_t nodeBegin, int64_t nodeEnd) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (nodeBegin < 0 || nodeEnd > h->nNodes || nodeBegin > nodeEnd) return fail(h, IKB_EINVAL, "bad node range");
  if (h->layout == LAYOUT_LEXICOGRAPHIC && h->meshUploaded && !(nodeBegin == 0 && nodeEnd == h->nNodes))
    return fail(h, IKB_ENOTIMPL, "row partitioning needs FlatInterleaved dofs");
  h->rowBegin = nodeBegin;
  h->rowEnd = nodeEnd;
  h->patternBuilt = false;
  h->reducedBuilt = false;
  h->rowSlowValid = false;
  h->sweepChunksBuilt = false;
  dropSolverGraphs(h);
  return IKB_OK;
}

int ikb_build_pattern(ikb_handle hh) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!h->meshUploaded) return fail(h, IKB_ESTATE, "upload the mesh first");
  const int nn = h->nn;
  const int64_t total = h->nElem * nn * nn;
  const int tpb = 256;
  const int64_t nRowNodes = h->rowEnd - h->rowBegin;
  h->nBlocks = 0;
  IKB_CUDA(h, h->nbrPtr.alloc((size_t)nRowNodes + 1));
  if (total == 0) {
    IKB_CUDA(h, cudaMemsetAsync(h->nbrPtr.p, 0, h->nbrPtr.bytes(), h->stream));
    h->patternBuilt = true;
    return IKB_OK;
  }
  DevBuf<uint64_t> keys, keysOut, ukeys;
  DevBuf<uint32_t> vals;
  DevBuf<int32_t> counts;
  DevBuf<int64_t> nRuns;
  DevBuf<uint8_t> tmp;
  IKB_CUDA(h, keys.alloc((size_t)total));
  IKB_CUDA(h, keysOut.alloc((size_t)total));
  IKB_CUDA(h, vals.alloc((size_t)total));
  IKB_CUDA(h, h->csrc.alloc((size_t)total));
  gen_pairs_kernel<<<gridFor(total, tpb), tpb, 0, h->stream>>>(h->elemNode.p, h->nElem, nn, h->npair, h->nNodes,
                                                              h->rowBegin, h->rowEnd, keys.p, vals.p);
  IKB_LAUNCH_CHECK(h);
  // key < nRowNodes*nNodes (or the all-ones sentinel of non-owned rows)
  int endBit = 64;
  if (h->rowBegin == 0 && h->rowEnd == h->nNodes) {
    const double maxKey = (double)nRowNodes * (double)h->nNodes;
    endBit = std::min(64, (int)std::ceil(std::log2(maxKey + 1.0)) + 1);
  }
  size_t tmpBytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys.p, keysOut.p, vals.p, h->csrc.p, total, 0, endBit, h->stream);
  IKB_CUDA(h, tmp.alloc(tmpBytes));
  IKB_CUDA(h, cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keys.p, keysOut.p, vals.p, h->csrc.p, total, 0, endBit,
                                              h->stream));
  h->launches++;
  vals.release();
  // run-length encode the sorted keys: unique keys = pattern blocks, counts = contributions per block
  IKB_CUDA(h, nRuns.alloc(1));
  IKB_CUDA(h, counts.alloc((size_t)total + 1));
  ukeys.p = keys.p;  // reuse the unsorted-key buffer for the unique keys
  tmpBytes = 0;
  cub::DeviceRunLengthEncode::Encode(nullptr, tmpBytes, keysOut.p, ukeys.p, counts.p, nRuns.p, total, h->stream);
  IKB_CUDA(h, tmp.alloc(tmpBytes));
  IKB_CUDA(h, cub::DeviceRunLengthEncode::Encode(tmp.p, tmpBytes, keysOut.p, ukeys.p, counts.p, nRuns.p, total,
                                                 h->stream));
  h->launches++;
  int64_t runs = 0;
  IKB_CUDA(h, cudaMemcpyAsync(&runs, nRuns.p, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  // drop the sentinel run of non-owned rows (it sorts last)
  uint64_t lastKey = 0;
  IKB_CUDA(h, cudaMemcpy(&lastKey, ukeys.p + (runs - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (lastKey == ~0ull) runs -= 1;
  if (runs >= (int64_t)INT32_MAX) {
    ukeys.p = nullptr;
    return fail(h, IKB_EINVAL, "too many pattern blocks for 32-bit block indices");
  }
  h->nBlocks = runs;
  IKB_CUDA(h, h->cptr.alloc((size_t)runs + 1));
  IKB_CUDA(h, h->nbrIdx.alloc((size_t)std::max<int64_t>(runs, 1)));
  IKB_CUDA(h, h->nbrRow.alloc((size_t)std::max<int64_t>(runs, 1)));
  tmpBytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, counts.p, h->cptr.p, runs + 1, h->stream);
  IKB_CUDA(h, tmp.alloc(tmpBytes));
  IKB_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, counts.p, h->cptr.p, runs + 1, h->stream));
  h->launches++;
  if (runs > 0) {
    decode_blocks_kernel<<<gridFor(runs, tpb), tpb, 0, h->stream>>>(ukeys.p, runs, h->nNodes, h->nbrIdx.p, h->nbrRow.p);
    IKB_LAUNCH_CHECK(h);
  }
  row_ptr_kernel<<<gridFor(nRowNodes + 1, tpb), tpb, 0, h->stream>>>(h->nbrRow.p, runs, nRowNodes, h->nbrPtr.p);
  IKB_LAUNCH_CHECK(h);
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  ukeys.p = nullptr;  // aliased keys.p
  keys.release();
  keysOut.release();
  nRuns.release();
  counts.release();
  // node -> (element, local node) adjacency in element order + slot tables: the map of the warp-per-node gather
  {
    const PatternView P = h->view();
    DevBuf<int32_t> adjCount, rowLen, maxLen;
    IKB_CUDA(h, adjCount.alloc((size_t)nRowNodes + 1));
    IKB_CUDA(h, rowLen.alloc((size_t)std::max<int64_t>(nRowNodes, 1)));
    IKB_CUDA(h, maxLen.alloc(1));
    IKB_CUDA(h, cudaMemsetAsync(adjCount.p, 0, adjCount.bytes(), h->stream));
    adj_count_kernel<<<gridFor(nRowNodes, tpb), tpb, 0, h->stream>>>(P, h->cptr.p, adjCount.p, rowLen.p);
    IKB_LAUNCH_CHECK(h);
    IKB_CUDA(h, h->adjPtr.alloc((size_t)nRowNodes + 1));
    tmpBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, adjCount.p, h->adjPtr.p, nRowNodes + 1, h->stream);
    IKB_CUDA(h, tmp.alloc(tmpBytes));
    IKB_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, adjCount.p, h->adjPtr.p, nRowNodes + 1, h->stream));
    h->launches++;
    tmpBytes = 0;
    cub::DeviceReduce::Max(nullptr, tmpBytes, rowLen.p, maxLen.p, nRowNodes, h->stream);
    IKB_CUDA(h, tmp.alloc(tmpBytes));
    IKB_CUDA(h, cub::DeviceReduce::Max(tmp.p, tmpBytes, rowLen.p, maxLen.p, nRowNodes, h->stream));
    h->launches++;
    int32_t nAdj = 0, mx = 0;
    IKB_CUDA(h, cudaMemcpyAsync(&nAdj, h->adjPtr.p + nRowNodes, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    IKB_CUDA(h, cudaMemcpyAsync(&mx, maxLen.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    IKB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->maxNbr = mx;
    if (mx > 255) return fail(h, IKB_ENOTIMPL, "pattern rows with more than 255 neighbour nodes are not supported");
    IKB_CUDA(h, h->adjCode.alloc((size_t)std::max(nAdj, 1)));
    IKB_CUDA(h, h->slotTab.alloc((size_t)std::max(nAdj, 1) * nn));
    adj_fill_kernel<<<gridFor(nRowNodes, tpb), tpb, 0, h->stream>>>(P, h->cptr.p, h->csrc.p, h->adjPtr.p, h->elemNode.p,
                                                                    h->nElem, nn, h->npair, h->adjCode.p, h->slotTab.p);
    IKB_LAUNCH_CHECK(h);
    IKB_CUDA(h, cudaStreamSynchronize(h->stream));
    adjCount.release();
    rowLen.release();
    maxLen.release();
    // the tile gather only needs the adjacency derived from the per-block contribution lists; the pull gather reads them
    if (!h->gatherPull) {
      h->csrc.release();
      h->cptr.release();
    }
  }
  tmp.release();
  h->patternBuilt = true;
  h->reducedBuilt = false;
  h->rowSlowValid = false;
  h->sweepChunksBuilt = false;
  dropSolverGraphs(h);
  h->rowDiag.release();
  h->rowLowEnd.release();
  h->mirrorBlk.release();
  h->fusedTried = h->fusedOk = false;
  for (int i = 0; i < 3; ++i) {
    h->vals[i].release();
    h->vec[i].release();
    h->valsVersion[i] = h->vecVersion[i] = 0;
  }
  return IKB_OK;
}

int ikb_pattern_nnz(ikb_handle hh, int dbc, int64_t* rows, int64_t* nnz) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc)) return fail(h, IKB_EINVAL, "bad dbc");
  if (!h->patternBuilt) return fail(h, IKB_ESTATE, "pattern not built");
  if (dbc == IKB_DBC_REDUCED) {
    int rc = ensureReduced(h);
    if (rc) return rc;
  }
  if (rows) *rows = rowsOf(h, dbc);
  if (nnz) *nnz = nnzOf(h, dbc);
  return IKB_OK;
}

int ikb_get_pattern(ikb_handle hh, int dbc, int64_t* outer, int32_t* inner) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc) || !outer || !inner) return fail(h, IKB_EINVAL, "bad arguments");
  if (!h->patternBuilt) return fail(h, IKB_ESTATE, "pattern not built");
  if (dbc == IKB_DBC_REDUCED) {
    int rc = ensureReduced(h);
    if (rc) return rc;
    IKB_CUDA(h, cudaMemcpy(outer, h->redOuter.p, (size_t)(h->nRed + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (h->nnzRed)
      IKB_CUDA(h, cudaMemcpy(inner, h->redInner.p, (size_t)h->nnzRed * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return IKB_OK;
  }
  const PatternView P = h->view();
  const int64_t rows = h->nRowsLocal(), nnz = h->nnzRaw();
  DevBuf<int64_t> o;
  DevBuf<int32_t> in;
  IKB_CUDA(h, o.alloc((size_t)rows + 1));
  IKB_CUDA(h, in.alloc((size_t)std::max<int64_t>(nnz, 1)));
  IKB_CUDA(h, cudaMemsetAsync(o.p, 0, o.bytes(), h->stream));
  if (h->nBlocks) {
    raw_pattern_kernel<<<gridFor(h->nBlocks, 256), 256, 0, h->stream>>>(P, o.p, in.p);
    IKB_LAUNCH_CHECK(h);
  }
  IKB_CUDA(h, cudaMemcpyAsync(outer, o.p, o.bytes(), cudaMemcpyDeviceToHost, h->stream));
  if (nnz) IKB_CUDA(h, cudaMemcpyAsync(inner, in.p, (size_t)nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  o.release();
  in.release();
  return IKB_OK;
}

int ikb_get_constraints_below(ikb_handle hh, int64_t* out) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!out) return fail(h, IKB_EINVAL, "null out");
  int rc = ensureReduced(h);
  if (rc) return rc;
  std::vector<int32_t> tmp((size_t)h->nDof);
  IKB_CUDA(h, cudaMemcpy(tmp.data(), h->cbelow.p, tmp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < tmp.size(); ++i) out[i] = tmp[i];
  return IKB_OK;
}

int ikb_element_linear_indices(ikb_handle hh, int64_t elem, int64_t* out) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!h->patternBuilt || !out || elem < 0 || elem >= h->nElem) return fail(h, IKB_EINVAL, "bad arguments");
  // host-side lookup in the node-block view (diagnostic path, not hot)
  const int D = h->dim, nn = h->nn;
  const int64_t nRowNodes = h->rowEnd - h->rowBegin;
  std::vector<int32_t> nodes(nn);
  for (int a = 0; a < nn; ++a)
    IKB_CUDA(h, cudaMemcpy(&nodes[a], h->elemNode.p + (size_t)a * h->nElem + elem, sizeof(int32_t),
                           cudaMemcpyDeviceToHost));
  std::vector<int32_t> ptr((size_t)nRowNodes + 1);
  IKB_CUDA(h, cudaMemcpy(ptr.data(), h->nbrPtr.p, ptr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  PatternView P = h->view();
  P.nbrPtr = ptr.data();
  int64_t q = 0;
  for (int cb = 0; cb < nn; ++cb)
    for (int ck = 0; ck < D; ++ck)
      for (int ra = 0; ra < nn; ++ra)
        for (int ri = 0; ri < D; ++ri, ++q) {
          const int64_t g = nodes[ra] - h->rowBegin;
          if (g < 0 || g >= nRowNodes) {
            out[q] = -1;
            continue;
          }
          const int nnb = ptr[g + 1] - ptr[g];
          std::vector<int32_t> nb(nnb);
          IKB_CUDA(h, cudaMemcpy(nb.data(), h->nbrIdx.p + ptr[g], nnb * sizeof(int32_t), cudaMemcpyDeviceToHost));
          const int slot = (int)(std::lower_bound(nb.begin(), nb.end(), nodes[cb]) - nb.begin());
          out[q] = rawRowStart(P, g, ri, nnb) + rawEntryOffset(P, slot, ck, nnb);
        }
  return IKB_OK;
}

int ikb_set_solution(ikb_handle hh, const double* d) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!d) return fail(h, IKB_EINVAL, "null solution");
  if (!h->U.p) IKB_CUDA(h, h->U.alloc((size_t)h->nDof));
  joinSolution(h);
  IKB_CUDA(h, cudaMemcpyAsync(h->U.p, d, h->U.bytes(), cudaMemcpyHostToDevice, h->stream));
  markSolutionUse(h);
  h->stateVersion++;
  return IKB_OK;
}

int ikb_set_solution_range(ikb_handle hh, const double* d, int64_t dofBegin, int64_t count) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!d || dofBegin < 0 || count < 0 || dofBegin + count > h->nDof) return fail(h, IKB_EINVAL, "bad dof range");
  int rc = ensureSolution(h);
  if (rc) return rc;
  if (count && !h->chunkDofEnd.empty() && count >= (int64_t)1 << 16) {
    // pieces on the side stream, one event each; launchElements() waits per chunk
    // The copy only has to wait for the latest consumer of U on the main stream (evUse), not for work queued behind it
    // (a matrix gather still running does not read U), so back-to-back steps overlap upload and gather.
    joinSolution(h);
    if (h->solutionShared) markSolutionUse(h);
    IKB_CUDA(h, cudaStreamWaitEvent(h->stream2, h->evUse, 0));
    const int64_t end = dofBegin + count;
    int64_t lo = dofBegin;
    for (int c = 0; c < Handle::SOL_CHUNKS; ++c) {
      const int64_t hi = c + 1 == Handle::SOL_CHUNKS ? end : std::min(end, std::max(lo, h->chunkDofEnd[c]));
      if (hi > lo)
        IKB_CUDA(h, cudaMemcpyAsync(h->U.p + lo, d + (lo - dofBegin), (size_t)(hi - lo) * sizeof(double),
                                    cudaMemcpyHostToDevice, h->stream2));
      IKB_CUDA(h, cudaEventRecord(h->evPiece[c], h->stream2));
      lo = hi;
    }
    h->piecesPending = true;
  } else if (count) {
    joinSolution(h);
    IKB_CUDA(h, cudaMemcpyAsync(h->U.p + dofBegin, d, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    markSolutionUse(h);
  }
  h->stateVersion++;
  return IKB_OK;
}

int ikb_get_solution(ikb_handle hh, double* d) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  int rc = ensureSolution(h);
  if (rc) return rc;
  joinSolution(h);
  IKB_CUDA(h, cudaMemcpyAsync(d, h->U.p, h->U.bytes(), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  return IKB_OK;
}

int ikb_set_parameter(ikb_handle hh, double lambda) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (lambda != h->lambda) h->stateVersion++;
  h->lambda = lambda;
  return IKB_OK;
}

int ikb_set_external_load(ikb_handle hh, const double* fext, int scales) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!fext) {
    h->hasFext = false;
    h->stateVersion++;
    return IKB_OK;
  }
  if (!h->Fext.p) IKB_CUDA(h, h->Fext.alloc((size_t)h->nDof));
  IKB_CUDA(h, cudaMemcpyAsync(h->Fext.p, fext, h->Fext.bytes(), cudaMemcpyHostToDevice, h->stream));
  h->hasFext = true;
  h->fextScales = scales ? 1 : 0;
  h->stateVersion++;
  return IKB_OK;
}

int ikb_assemble(ikb_handle hh, unsigned what, int dbc) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc) || (what & ~7u) || what == 0) return fail(h, IKB_EINVAL, "bad affordance/dbc");
  if (!h->meshUploaded || !h->patternBuilt) return fail(h, IKB_ESTATE, "mesh/pattern missing");
  if (dbc != IKB_DBC_RAW && !h->hasFlags) return fail(h, IKB_ESTATE, "Dirichlet flags missing");
  if (h->form == FORM_PS && !h->psSet) return fail(h, IKB_ESTATE, "ikb_set_hyperelastic has not been called");
  if ((what & IKB_SCALAR) && h->easM)
    return fail(h, IKB_ENOTIMPL,
                "EAS element do not support any scalar calculations, i.e. they are not derivable from a potential");
  int rc;
  if ((rc = ensureSolution(h))) return rc;
  if (dbc == IKB_DBC_REDUCED && (rc = ensureReduced(h))) return rc;
  if ((rc = ensureFused(h))) return rc;

  // which outputs are stale for the current (d, lambda, alpha)?
  unsigned needGather = 0;
  if ((what & IKB_MATRIX) && h->valsVersion[dbc] != h->stateVersion) needGather |= IKB_MATRIX;
  if ((what & IKB_VECTOR) && h->vecVersion[dbc] != h->stateVersion) needGather |= IKB_VECTOR;
  const bool needEnergy = (what & IKB_SCALAR) && h->energyVersion != h->stateVersion;
  unsigned needStage = needGather | (needEnergy ? IKB_SCALAR : 0);
  if (h->fusedOk && needGather) {
    // Hex8 fused sweep: elements and row assembly in one launch, nothing staged in HBM (ikb_fused.cuh)
    const bool stagedE = h->stagedVersion == h->stateVersion && (h->stagedWhat & IKB_SCALAR);
    const unsigned fw = needGather | ((needEnergy && !stagedE) ? IKB_SCALAR : 0);
    if ((rc = ensureStaging(h, fw & IKB_SCALAR))) return rc;
    if ((needGather & IKB_MATRIX) && !h->vals[dbc].p)
      IKB_CUDA(h, h->vals[dbc].alloc((size_t)std::max<int64_t>(nnzOf(h, dbc), 1)));
    if ((needGather & IKB_VECTOR) && !h->vec[dbc].p)
      IKB_CUDA(h, h->vec[dbc].alloc((size_t)std::max<int64_t>(rowsOf(h, dbc), 1)));
    if ((rc = launchFused(h, fw, dbc))) return rc;
    if (needGather & IKB_MATRIX) h->valsVersion[dbc] = h->stateVersion;
    if (needGather & IKB_VECTOR) h->vecVersion[dbc] = h->stateVersion;
    if (fw & IKB_SCALAR) {
      h->stagedWhat = (h->stagedVersion == h->stateVersion ? h->stagedWhat : 0) | IKB_SCALAR;
      h->stagedVersion = h->stateVersion;
    }
    needGather = 0;
    needStage = 0;
  }
  if (h->stagedVersion == h->stateVersion) needStage &= ~h->stagedWhat;
  if ((needStage & IKB_MATRIX) && (needGather & IKB_MATRIX)) {
    if ((rc = ensureSweepChunks(h))) return rc;
    if (!h->sweepElemEnd.empty()) {
      if ((rc = ensureStaging(h, needStage))) return rc;
      if ((needGather & IKB_MATRIX) && !h->vals[dbc].p)
        IKB_CUDA(h, h->vals[dbc].alloc((size_t)std::max<int64_t>(nnzOf(h, dbc), 1)));
      if ((needGather & IKB_VECTOR) && !h->vec[dbc].p)
        IKB_CUDA(h, h->vec[dbc].alloc((size_t)std::max<int64_t>(rowsOf(h, dbc), 1)));
      if ((rc = launchInterleaved(h, needStage, needGather, dbc))) return rc;
      h->stagedWhat = (h->stagedVersion == h->stateVersion ? h->stagedWhat : 0) | needStage;
      h->stagedVersion = h->stateVersion;
      if (needGather & IKB_MATRIX) h->valsVersion[dbc] = h->stateVersion;
      if (needGather & IKB_VECTOR) h->vecVersion[dbc] = h->stateVersion;
      needStage = 0;
      needGather = 0;
    }
  }
  if (needStage) {
    if ((rc = ensureStaging(h, needStage))) return rc;
    const unsigned stageWhat = needStage;
    if ((rc = launchElements(h, stageWhat))) return rc;
    h->stagedWhat = (h->stagedVersion == h->stateVersion ? h->stagedWhat : 0) | stageWhat;
    h->stagedVersion = h->stateVersion;
  }
  if (needGather) {
    if ((needGather & IKB_MATRIX) && !h->vals[dbc].p)
      IKB_CUDA(h, h->vals[dbc].alloc((size_t)std::max<int64_t>(nnzOf(h, dbc), 1)));
    if ((needGather & IKB_VECTOR) && !h->vec[dbc].p)
      IKB_CUDA(h, h->vec[dbc].alloc((size_t)std::max<int64_t>(rowsOf(h, dbc), 1)));
    if ((rc = launchGather(h, needGather, dbc))) return rc;
    if (needGather & IKB_MATRIX) h->valsVersion[dbc] = h->stateVersion;
    if (needGather & IKB_VECTOR) h->vecVersion[dbc] = h->stateVersion;
  }
  if (needEnergy) {
    double* eDev = h->cgScal.p + 8;
    double* lDev = h->cgScal.p + 9;
    if (h->hasFext) {
      // E -= s * fext . d   (loads/volume.hh:67-84, traction.hh:70-105)
      if ((rc = deviceDot(h, 1, h->Fext.p, h->U.p, h->nDof, lDev, -(h->fextScales ? h->lambda : 1.0), nullptr)))
        return rc;
      markSolutionUse(h);
    }
    const bool partitioned = h->rowBegin != 0 || h->rowEnd != h->nNodes;
    if (partitioned && h->nElem) {
      mask_unowned_energy_kernel<<<gridFor(h->nElem, 256), 256, 0, h->stream>>>(h->elemNode.p, h->nElem, h->rowBegin,
                                                                                h->rowEnd, h->Est.p);
      IKB_LAUNCH_CHECK(h);
      h->stagedWhat &= ~IKB_SCALAR;  // Est was modified in place
    }
    // with a communicator every rank holds the full d and fext, so the load term is added once, after the reduction
    if ((rc = deviceDot(h, 0, h->Est.p, nullptr, h->nElem, eDev, 1.0, (h->hasFext && !h->comm) ? lDev : nullptr)))
      return rc;
    if (h->comm) {
      if ((rc = allReduceSum(h, eDev, 1))) return rc;
      if (h->hasFext) {
        vec_axpy_kernel<<<1, 1, 0, h->stream>>>(1, 1.0, lDev, eDev);
        IKB_LAUNCH_CHECK(h);
      }
    }
    h->energyVersion = h->stateVersion;
  }
  return IKB_OK;
}

int ikb_invalidate(ikb_handle hh) {
  Handle* h = H(hh);
  if (!h) return IKB_EINVAL;
  h->stateVersion++;
  return IKB_OK;
}

int ikb_get_vector(ikb_handle hh, int dbc, double* out) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc) || !out) return fail(h, IKB_EINVAL, "bad arguments");
  if (h->vecVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "vector not assembled for the current state");
  const int64_t n = rowsOf(h, dbc);
  // R is complete once the residual gather has run (evVec); it is copied on the side stream so that a matrix gather
  // still running on the main stream is not waited for.  The material error flag is final by then as well.
  IKB_CUDA(h, cudaStreamWaitEvent(h->stream2, h->evVec, 0));
  int32_t flag = INT_MAX;
  if (n) IKB_CUDA(h, cudaMemcpyAsync(out, h->vec[dbc].p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream2));
  IKB_CUDA(h, cudaMemcpyAsync(&flag, h->errFlag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream2));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream2));
  if (flag != INT_MAX) return checkMaterialError(h);
  return IKB_OK;
}

int ikb_get_scalar(ikb_handle hh, double* e) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!e) return fail(h, IKB_EINVAL, "null out");
  if (h->energyVersion != h->stateVersion) return fail(h, IKB_ESTATE, "scalar not assembled for the current state");
  IKB_CUDA(h, cudaMemcpyAsync(e, h->cgScal.p + 8, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return checkMaterialError(h);
}

int ikb_get_matrix_values(ikb_handle hh, int dbc, double* out) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc) || !out) return fail(h, IKB_EINVAL, "bad arguments");
  if (h->valsVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "matrix not assembled for the current state");
  const int64_t nnz = nnzOf(h, dbc);
  if (nnz)
    IKB_CUDA(h, cudaMemcpyAsync(out, h->vals[dbc].p, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return checkMaterialError(h);
}

int ikb_get_dense_matrix(ikb_handle hh, int dbc, double* out) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc) || !out) return fail(h, IKB_EINVAL, "bad arguments");
  if (h->valsVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "matrix not assembled for the current state");
  const int64_t rows = rowsOf(h, dbc);
  if (rows == 0) return IKB_OK;
  if ((double)rows * rows * 8.0 > 8e9) return fail(h, IKB_EINVAL, "dense matrix too large");
  DevBuf<double> dense;
  IKB_CUDA(h, dense.alloc((size_t)rows * rows));
  IKB_CUDA(h, cudaMemsetAsync(dense.p, 0, dense.bytes(), h->stream));
  DevBuf<int64_t> o;
  DevBuf<int32_t> in;
  const int64_t* outer;
  const int32_t* inner;
  if (dbc == IKB_DBC_REDUCED) {
    outer = h->redOuter.p;
    inner = h->redInner.p;
  } else {
    IKB_CUDA(h, o.alloc((size_t)rows + 1));
    IKB_CUDA(h, in.alloc((size_t)h->nnzRaw()));
    raw_pattern_kernel<<<gridFor(h->nBlocks, 256), 256, 0, h->stream>>>(h->view(), o.p, in.p);
    IKB_LAUNCH_CHECK(h);
    outer = o.p;
    inner = in.p;
  }
  csr_to_dense_kernel<<<gridFor(rows, 128), 128, 0, h->stream>>>(outer, inner, h->vals[dbc].p, rows, dense.p);
  IKB_LAUNCH_CHECK(h);
  IKB_CUDA(h, cudaMemcpyAsync(out, dense.p, dense.bytes(), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  dense.release();
  o.release();
  in.release();
  return IKB_OK;
}

int ikb_vector_norm(ikb_handle hh, int dbc, double* norm) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc) || !norm) return fail(h, IKB_EINVAL, "bad arguments");
  if (h->vecVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "vector not assembled for the current state");
  const int64_t n = rowsOf(h, dbc);
  int rc = deviceDot(h, 2, h->vec[dbc].p, nullptr, n, h->cgScal.p + 10, 1.0, nullptr);
  if (rc) return rc;
  if ((rc = allReduceSum(h, h->cgScal.p + 10, 1))) return rc;
  double s = 0.0;
  IKB_CUDA(h, cudaMemcpyAsync(&s, h->cgScal.p + 10, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  rc = checkMaterialError(h);
  *norm = std::sqrt(s);
  return rc;
}

int ikb_eas_update(ikb_handle hh, const double* correction) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!h->easM) return IKB_OK;
  if (!h->meshUploaded) return fail(h, IKB_ESTATE, "mesh missing");
  int rc;
  if ((rc = ensureSolution(h))) return rc;
  if ((rc = ensureStaging(h, 0))) return rc;  // update mode stages nothing
  if (h->Corr.n < (size_t)h->nDof) {
    if (!correction) return fail(h, IKB_ESTATE, "no resident correction");
    IKB_CUDA(h, h->Corr.alloc((size_t)h->nDof));
  }
  if (correction)
    IKB_CUDA(h, cudaMemcpyAsync(h->Corr.p, correction, (size_t)h->nDof * sizeof(double), cudaMemcpyHostToDevice,
                                h->stream));
  // D, L, Rtilde at the OLD (d, alpha): must run before ikb_update_solution / ikb_set_solution of the new d
  if ((rc = launchElements(h, 0, h->Corr.p))) return rc;
  h->stateVersion++;
  return IKB_OK;
}
int ikb_calculate_at(ikb_handle hh, int resultType, const double* local, int nPoints, double* out) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!local || !out || nPoints <= 0) return fail(h, IKB_EINVAL, "bad arguments");
  if (resultType < IKB_RESULT_LINEAR_STRESS || resultType > IKB_RESULT_CAUCHY_STRESS)
    return fail(h, IKB_EINVAL, "unknown result type");
  if (!h->meshUploaded) return fail(h, IKB_ESTATE, "mesh missing");
  const bool linearType = resultType == IKB_RESULT_LINEAR_STRESS || resultType == IKB_RESULT_LINEAR_STRESS_FULL;
  if (linearType != (h->form == FORM_LE))  // supportsResultType (linearelastic.hh / nonlinearelastic.hh)
    return fail(h, IKB_ENOTIMPL, "The requested result type is not supported by this element");
  int rc;
  if ((rc = ensureSolution(h))) return rc;
  joinSolution(h);
  const int S = h->dim * (h->dim + 1) / 2;
  const bool full = resultType == IKB_RESULT_LINEAR_STRESS_FULL || resultType == IKB_RESULT_PK2_STRESS_FULL;
  const int ncomp = full ? 6 : S;
  const int64_t nOut = h->nElem * (int64_t)nPoints * ncomp;
  if (nOut == 0) return IKB_OK;
  DevBuf<double> dLocal, dOut, aTmp, uZero;
  IKB_CUDA(h, dLocal.alloc((size_t)nPoints * h->dim));
  IKB_CUDA(h, dOut.alloc((size_t)nOut));
  IKB_CUDA(h, cudaMemcpyAsync(dLocal.p, local, dLocal.bytes(), cudaMemcpyHostToDevice, h->stream));
  const double* alpha = h->alpha.p;
  if (h->easM && h->form == FORM_LE) {
    // linear strains: alpha = -D^-1 L d recomputed from the current d (enhancedassumedstrains.hh:152-157).  The
    // update kernel evaluates alpha - D^-1 (Rt + L du) at the state it is given; at (d = 0, alpha = 0) Rt vanishes,
    // so du = d yields exactly -D^-1 L d (D, L do not depend on the state for the linear element).
    if ((rc = ensureStaging(h, 0))) return rc;  // update mode stages nothing
    IKB_CUDA(h, aTmp.alloc((size_t)h->nElem * h->easM));
    IKB_CUDA(h, uZero.alloc((size_t)h->nDof));
    IKB_CUDA(h, cudaMemsetAsync(aTmp.p, 0, aTmp.bytes(), h->stream));
    IKB_CUDA(h, cudaMemsetAsync(uZero.p, 0, uZero.bytes(), h->stream));
    if ((rc = launchElements(h, 0, h->U.p, uZero.p, aTmp.p))) return rc;
    alpha = aTmp.p;
  }
  ResultArgs A;
  A.X = h->X.p;
  A.elemNode = h->elemNode.p;
  A.U = h->U.p;
  A.T0inv = h->T0inv.p;
  A.alpha = alpha;
  A.local = dLocal.p;
  A.out = dOut.p;
  A.errFlag = h->errFlag.p;
  A.nElem = h->nElem;
  A.nNodes = h->nNodes;
  A.layout = h->layout;
  A.npts = nPoints;
  A.form = h->form;
  A.ps = h->ps;
  A.planeStrain = h->desc.plane_strain;
  A.easM = h->easM;
  A.easFunction = h->easFunction;
  A.resultType = resultType;
  A.ncomp = ncomp;
  A.lambda = h->desc.lambda;
  A.mu = h->desc.mu;
  A.psTol = h->desc.reduce_tol > 0.0 ? h->desc.reduce_tol : 1e-12;
  const unsigned grid = gridFor(h->nElem * (int64_t)nPoints, 128);
  if (h->dim == 3 && h->order == 1)
    result_at_kernel<3, 1><<<grid, 128, 0, h->stream>>>(A);
  else if (h->dim == 3)
    result_at_kernel<3, 2><<<grid, 128, 0, h->stream>>>(A);
  else if (h->order == 1)
    result_at_kernel<2, 1><<<grid, 128, 0, h->stream>>>(A);
  else
    result_at_kernel<2, 2><<<grid, 128, 0, h->stream>>>(A);
  IKB_LAUNCH_CHECK(h);
  markSolutionUse(h);
  IKB_CUDA(h, cudaMemcpyAsync(out, dOut.p, dOut.bytes(), cudaMemcpyDeviceToHost, h->stream));
  return checkMaterialError(h);
}

int ikb_eas_get_alpha(ikb_handle hh, double* alpha) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!h->easM || !alpha) return fail(h, IKB_EINVAL, "no EAS state");
  IKB_CUDA(h, cudaMemcpyAsync(alpha, h->alpha.p, h->alpha.bytes(), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  return IKB_OK;
}
int ikb_set_hyperelastic(ikb_handle hh, const ikb_hyperelastic* law) {
  Handle* h = H(hh);
  if (checkHandle(h) || !law) return IKB_EINVAL;
  if (h->form != FORM_PS || h->desc.material != IKB_MAT_HYPERELASTIC)
    return fail(h, IKB_EINVAL, "the handle was not created with IKB_MAT_HYPERELASTIC");
  if (law->deviatoric < IKB_DEV_NONE || law->deviatoric > IKB_DEV_GENT || law->volumetric < 0 || law->volumetric > 12)
    return fail(h, IKB_EINVAL, "unknown deviatoric or volumetric function");
  const bool terms = law->deviatoric == IKB_DEV_OGDEN_TOTAL || law->deviatoric == IKB_DEV_OGDEN_DEVIATORIC ||
                     law->deviatoric == IKB_DEV_INVARIANT_BASED;
  if (terms && (law->n < 1 || law->n > 3)) return fail(h, IKB_ENOTIMPL, "1 to 3 terms are served");
  if (law->deviatoric == IKB_DEV_NONE && law->volumetric == 0) return fail(h, IKB_EINVAL, "a law without any part");
  PsLaw ps;
  ps.dev = law->deviatoric;
  ps.n = terms ? law->n : 0;
  ps.vf = law->volumetric;
  for (int i = 0; i < 3; ++i) {
    ps.par[i] = law->par[i];
    ps.ex[i] = law->ex[i];
    ps.pex[i] = law->pex[i];
    ps.qex[i] = law->qex[i];
    if (law->deviatoric == IKB_DEV_INVARIANT_BASED && i < law->n) {
      // InvariantBasedT::checkExponents (invariantbased.hh:216-221); exponents are std::size_t there
      if (law->pex[i] < 0 || law->qex[i] < 0 || (law->pex[i] == 0 && law->qex[i] == 0))
        return fail(h, IKB_EINVAL, "the exponents p_i and q_i should not be zero at the same time");
    }
    if ((law->deviatoric == IKB_DEV_OGDEN_TOTAL || law->deviatoric == IKB_DEV_OGDEN_DEVIATORIC) && i < law->n &&
        law->ex[i] == 0.0)
      return fail(h, IKB_EINVAL, "an Ogden exponent is zero");
  }
  ps.K = law->K;
  ps.beta = law->beta;
  if ((ps.vf == 4 || ps.vf == 7 || ps.vf == 10) && ps.beta == 0.0) return fail(h, IKB_EINVAL, "this volumetric function needs beta");
  h->ps = ps;
  h->psSet = true;
  h->stateVersion++;
  return IKB_OK;
}

int ikb_eas_set_alpha(ikb_handle hh, const double* alpha) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!h->easM || !alpha) return fail(h, IKB_EINVAL, "no EAS state");
  IKB_CUDA(h, cudaMemcpyAsync(h->alpha.p, alpha, h->alpha.bytes(), cudaMemcpyHostToDevice, h->stream));
  h->stateVersion++;
  return IKB_OK;
}

int ikb_spmv(ikb_handle hh, int dbc, const double* x, double* y) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc) || !x || !y) return fail(h, IKB_EINVAL, "bad arguments");
  if (h->valsVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "matrix not assembled for the current state");
  const int64_t n = rowsOf(h, dbc);
  if (n == 0) return IKB_OK;
  const int64_t nx = dbc == IKB_DBC_REDUCED ? h->nRed : h->nDof;
  if (h->cgP.n < (size_t)nx) IKB_CUDA(h, h->cgP.alloc((size_t)nx));
  if (h->cgQ.n < (size_t)n) IKB_CUDA(h, h->cgQ.alloc((size_t)n));
  IKB_CUDA(h, cudaMemcpyAsync(h->cgP.p, x, (size_t)nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int rc = launchSpmv(h, dbc, h->cgP.p, h->cgQ.p);
  if (rc) return rc;
  IKB_CUDA(h, cudaMemcpyAsync(y, h->cgQ.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  return IKB_OK;
}

int ikb_idbc_forces(ikb_handle hh, int dbc, const double* dInc, double* out) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc) || !dInc || !out) return fail(h, IKB_EINVAL, "bad arguments");
  if (!h->hasFlags) return fail(h, IKB_ESTATE, "Dirichlet flags missing");
  if (h->rowBegin != 0 || h->rowEnd != h->nNodes || h->comm)
    return fail(h, IKB_ENOTIMPL, "inhomogeneous Dirichlet forces on a partitioned handle");
  int rc;
  if ((rc = ikb_assemble(hh, IKB_MATRIX, IKB_DBC_RAW))) return rc;  // assembler->matrix(DBCOption::Raw) (:172)
  const int64_t n = h->nDof;
  if (n == 0) return IKB_OK;
  if (h->cgP.n < (size_t)n) IKB_CUDA(h, h->cgP.alloc((size_t)n));
  if (h->cgQ.n < (size_t)n) IKB_CUDA(h, h->cgQ.alloc((size_t)n));
  IKB_CUDA(h, cudaMemcpyAsync(h->cgP.p, dInc, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if ((rc = launchSpmv(h, IKB_DBC_RAW, h->cgP.p, h->cgQ.p))) return rc;
  const int tpb = 256;
  if (dbc == IKB_DBC_FULL) {
    zero_flagged_kernel<<<gridFor(n, tpb), tpb, 0, h->stream>>>(n, h->flags.p, h->cgQ.p);  // setZeroAtConstrainedDofs
    IKB_LAUNCH_CHECK(h);
    IKB_CUDA(h, cudaMemcpyAsync(out, h->cgQ.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  } else {
    if ((rc = ensureReduced(h))) return rc;
    if (h->cgZ.n < (size_t)std::max<int64_t>(h->nRed, 1)) IKB_CUDA(h, h->cgZ.alloc((size_t)std::max<int64_t>(h->nRed, 1)));
    contract_full_kernel<<<gridFor(n, tpb), tpb, 0, h->stream>>>(n, h->flags.p, h->cbelow.p, h->cgQ.p, h->cgZ.p);
    IKB_LAUNCH_CHECK(h);
    if (h->nRed)
      IKB_CUDA(h, cudaMemcpyAsync(out, h->cgZ.p, (size_t)h->nRed * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  return IKB_OK;
}

int ikb_pcg_solve(ikb_handle hh, int dbc, const double* rhs, double* x, double relTol, int maxIt, int* itersOut,
                  double* relResOut) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (dbc != IKB_DBC_FULL && dbc != IKB_DBC_REDUCED) return fail(h, IKB_EINVAL, "PCG needs the Full or Reduced matrix");
  if (h->rowBegin != 0 || h->rowEnd != h->nNodes || h->comm) {
    if (!h->comm) return fail(h, IKB_ESTATE, "partitioned handle: call ikb_comm_init before solving");
    if (dbc != IKB_DBC_FULL) return fail(h, IKB_ENOTIMPL, "distributed PCG supports DBCOption::Full");
    // (the state of this rank is checked inside, together with the other ranks')
    return distPcg(h, rhs, x, relTol, maxIt, itersOut, relResOut);
  }
  if (h->valsVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "matrix not assembled for the current state");
  const int64_t n = rowsOf(h, dbc);
  if (itersOut) *itersOut = 0;
  if (relResOut) *relResOut = 0.0;
  if (n == 0) return IKB_OK;
  for (auto* b : {&h->cgR, &h->cgZ, &h->cgP, &h->cgQ, &h->cgX, &h->cgDinv})
    if (b->n < (size_t)n) IKB_CUDA(h, b->alloc((size_t)n));
  if (h->Corr.n < (size_t)h->nDof) IKB_CUDA(h, h->Corr.alloc((size_t)h->nDof));
  const int tpb = 256;
  int rc;
  // r = b (x0 = 0)
  if (rhs) {
    IKB_CUDA(h, cudaMemcpyAsync(h->cgR.p, rhs, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  } else {
    if (h->vecVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "resident residual not assembled");
    vec_scale_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, -1.0, h->vec[dbc].p, h->cgR.p);
    IKB_LAUNCH_CHECK(h);
  }
  IKB_CUDA(h, cudaMemsetAsync(h->cgX.p, 0, (size_t)n * sizeof(double), h->stream));
  if (dbc == IKB_DBC_REDUCED) {
    diag_inv_csr_kernel<<<gridFor(n, tpb), tpb, 0, h->stream>>>(h->redOuter.p, h->redInner.p, h->vals[dbc].p, n,
                                                                h->cgDinv.p);
  } else if (h->dim == 3) {
    diag_inv_block_kernel<3><<<gridFor(h->nBlocks, tpb), tpb, 0, h->stream>>>(h->view(), h->vals[dbc].p, h->cgDinv.p);
  } else {
    diag_inv_block_kernel<2><<<gridFor(h->nBlocks, tpb), tpb, 0, h->stream>>>(h->view(), h->vals[dbc].p, h->cgDinv.p);
  }
  IKB_LAUNCH_CHECK(h);
  double* scal = h->cgScal.p;
  // z = Dinv r ; p = z ; rz = r.z ; bb = r.r
  vec_mul_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, h->cgDinv.p, h->cgR.p, h->cgZ.p);
  IKB_LAUNCH_CHECK(h);
  IKB_CUDA(h, cudaMemcpyAsync(h->cgP.p, h->cgZ.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if ((rc = deviceDot(h, 1, h->cgR.p, h->cgZ.p, n, scal + 0, 1.0, nullptr))) return rc;
  if ((rc = deviceDot(h, 2, h->cgR.p, nullptr, n, scal + 4, 1.0, nullptr))) return rc;
  double bb = 0.0;
  IKB_CUDA(h, cudaMemcpyAsync(&bb, scal + 4, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  int it = 0;
  double rr = bb;
  // Eigen's criterion: stop when |r|^2 < tol^2 |b|^2 (ConjugateGradient.h), x = 0 for b = 0
  const double threshold = std::max(relTol * relTol * bb, 1e-300);
  if (bb > 0.0 && dbc == IKB_DBC_FULL) {
    // sync-free path: 3 kernels per iteration, convergence decided on the device, the host looks at the state
    // once per batch of iterations
    if (!h->cgState.p) IKB_CUDA(h, h->cgState.alloc(128));
    CgState* st = reinterpret_cast<CgState*>(h->cgState.p);
    unsigned int* arrive = reinterpret_cast<unsigned int*>(h->cgState.p + 96);
    cg2_init_kernel<<<1, 1, 0, h->stream>>>(st, scal + 0, scal + 4, relTol, maxIt, arrive);
    IKB_LAUNCH_CHECK(h);
    const PatternView P = h->view();
    double* pqPartial = h->scratch.p;
    double* rzPartial = h->scratch.p + MAX_SPMV_BLOCKS;
    CgState* hs = reinterpret_cast<CgState*>(h->hostScal);
    const int batch = 32;
    auto enqueueBatch = [&]() {
      for (int b = 0; b < batch; ++b) {
        if (h->dim == 3)
          spmv_node_dot_kernel<3><<<h->spmvBlocks, tpb, 0, h->stream>>>(P, h->vals[dbc].p, h->cgP.p, h->cgQ.p, h->cgP.p,
                                                                        pqPartial, st);
        else
          spmv_node_dot_kernel<2><<<h->spmvBlocks, tpb, 0, h->stream>>>(P, h->vals[dbc].p, h->cgP.p, h->cgQ.p, h->cgP.p,
                                                                        pqPartial, st);
        cg2_update_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, st, pqPartial, h->spmvBlocks, h->cgP.p, h->cgQ.p,
                                                             h->cgDinv.p, h->cgX.p, h->cgR.p, h->cgZ.p, rzPartial);
        cg2_direction_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, st, rzPartial, RED_BLOCKS, h->cgZ.p, h->cgP.p, arrive);
      }
    };
    // a batch is captured once into a CUDA graph (all arguments are resident pointers) and replayed
    if (!h->cgGraph || h->cgGraphKey[0] != h->vals[dbc].p || h->cgGraphKey[1] != h->cgP.p ||
        h->cgGraphKey[2] != h->nbrIdx.p) {
      if (h->cgGraph) cudaGraphExecDestroy(h->cgGraph);
      h->cgGraph = nullptr;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        enqueueBatch();
        if (cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && graph) {
          if (cudaGraphInstantiate(&h->cgGraph, graph, 0) != cudaSuccess) h->cgGraph = nullptr;
          cudaGraphDestroy(graph);
        }
      }
      cudaGetLastError();
      h->cgGraphKey[0] = h->vals[dbc].p;
      h->cgGraphKey[1] = h->cgP.p;
      h->cgGraphKey[2] = h->nbrIdx.p;
    }
    while (true) {
      if (h->cgGraph) {
        IKB_CUDA(h, cudaGraphLaunch(h->cgGraph, h->stream));
      } else {
        enqueueBatch();
      }
      h->launches += 3 * batch;
      IKB_CUDA(h, cudaGetLastError());
      IKB_CUDA(h, cudaMemcpyAsync(hs, st, sizeof(CgState), cudaMemcpyDeviceToHost, h->stream));
      IKB_CUDA(h, cudaStreamSynchronize(h->stream));
      if (hs->done || hs->iter >= maxIt) break;
    }
    it = hs->iter;
    rr = hs->rr;
    if (hs->done == 2) return fail(h, IKB_ECUDA, "PCG produced NaN (matrix not positive definite?)");
  } else if (bb > 0.0) {
    while (it < maxIt && rr >= threshold) {
      if ((rc = launchSpmv(h, dbc, h->cgP.p, h->cgQ.p))) return rc;
      if ((rc = deviceDot(h, 1, h->cgP.p, h->cgQ.p, n, scal + 1, 1.0, nullptr))) return rc;
      cg_update_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, scal, h->cgP.p, h->cgQ.p, h->cgDinv.p, h->cgX.p, h->cgR.p,
                                                          h->cgZ.p, h->scratch.p + MAX_SPMV_BLOCKS);
      IKB_LAUNCH_CHECK(h);
      cg_fold2_kernel<<<1, tpb, 0, h->stream>>>(h->scratch.p + MAX_SPMV_BLOCKS, RED_BLOCKS, scal, nullptr);
      IKB_LAUNCH_CHECK(h);
      IKB_CUDA(h, cudaMemcpyAsync(h->hostScal, scal + 3, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      cg_direction_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, scal, h->cgZ.p, h->cgP.p);
      IKB_LAUNCH_CHECK(h);
      cg_shift_kernel<<<1, 1, 0, h->stream>>>(scal);
      IKB_LAUNCH_CHECK(h);
      IKB_CUDA(h, cudaStreamSynchronize(h->stream));
      rr = h->hostScal[0];
      ++it;
      if (!(rr == rr)) return fail(h, IKB_ECUDA, "PCG produced NaN (matrix not positive definite?)");
    }
  }
  if (itersOut) *itersOut = it;
  if (relResOut) *relResOut = bb > 0.0 ? std::sqrt(rr / bb) : 0.0;
  // keep the correction resident at full size for ikb_update_solution / ikb_eas_update
  if (dbc == IKB_DBC_REDUCED) {
    expand_reduced_kernel<<<gridFor(h->nDof, tpb), tpb, 0, h->stream>>>(h->nDof, h->flags.p, h->cbelow.p, h->cgX.p,
                                                                        h->Corr.p);
    IKB_LAUNCH_CHECK(h);
  } else {
    IKB_CUDA(h, cudaMemcpyAsync(h->Corr.p, h->cgX.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  if (x) IKB_CUDA(h, cudaMemcpyAsync(x, h->cgX.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  return IKB_OK;
}

int ikb_tcg_solve(ikb_handle hh, int dbc, const double* rhs, double* x, ikb_tcg_info* info) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!info) return fail(h, IKB_EINVAL, "null tcg info");
  if (dbc != IKB_DBC_FULL && dbc != IKB_DBC_REDUCED) return fail(h, IKB_EINVAL, "tCG needs the Full or Reduced matrix");
  if (info->precond != IKB_PRECOND_IDENTITY && info->precond != IKB_PRECOND_DIAGONAL)
    return fail(h, IKB_ENOTIMPL, "tCG preconditioner: only Identity and Diagonal run on the device");
  if (!(info->delta > 0.0)) return fail(h, IKB_EINVAL, "trust-region radius must be positive");
  if (h->valsVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "matrix not assembled for the current state");
  if (h->rowBegin != 0 || h->rowEnd != h->nNodes || h->comm)
    return fail(h, IKB_ENOTIMPL, "tCG on a partitioned handle");
  const int64_t n = rowsOf(h, dbc);
  info->iterations = 0;
  info->stop_reason = IKB_TCG_MAXIMUM_INNER_ITERATIONS;
  info->rel_error = info->eta_norm = info->g_dot_eta = info->eta_h_eta = 0.0;
  if (n == 0) return IKB_OK;
  for (auto* b : {&h->cgR, &h->cgZ, &h->cgP, &h->cgQ, &h->cgX, &h->cgDinv, &h->cgB})
    if (b->n < (size_t)n) IKB_CUDA(h, b->alloc((size_t)n));
  if (h->Corr.n < (size_t)h->nDof) IKB_CUDA(h, h->Corr.alloc((size_t)h->nDof));
  const int tpb = 256;
  int rc;
  double* scal = h->cgScal.p;
  double hostS[2];
  auto fetch = [&](const double* dev, int cnt) -> int {
    IKB_CUDA(h, cudaMemcpyAsync(hostS, dev, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    IKB_CUDA(h, cudaStreamSynchronize(h->stream));
    return IKB_OK;
  };
  // b (kept for g.eta), residual = b - H*0 = b, x = 0
  if (rhs) {
    IKB_CUDA(h, cudaMemcpyAsync(h->cgB.p, rhs, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  } else {
    if (h->vecVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "resident gradient not assembled");
    vec_scale_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, -1.0, h->vec[dbc].p, h->cgB.p);
    IKB_LAUNCH_CHECK(h);
  }
  IKB_CUDA(h, cudaMemcpyAsync(h->cgR.p, h->cgB.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  IKB_CUDA(h, cudaMemsetAsync(h->cgX.p, 0, (size_t)n * sizeof(double), h->stream));
  if (info->precond == IKB_PRECOND_DIAGONAL) {
    if (dbc == IKB_DBC_REDUCED) {
      if ((rc = ensureReduced(h))) return rc;
      diag_inv_csr_kernel<<<gridFor(n, tpb), tpb, 0, h->stream>>>(h->redOuter.p, h->redInner.p, h->vals[dbc].p, n,
                                                                  h->cgDinv.p);
    } else if (h->dim == 3) {
      diag_inv_block_kernel<3><<<gridFor(h->nBlocks, tpb), tpb, 0, h->stream>>>(h->view(), h->vals[dbc].p, h->cgDinv.p);
    } else {
      diag_inv_block_kernel<2><<<gridFor(h->nBlocks, tpb), tpb, 0, h->stream>>>(h->view(), h->vals[dbc].p, h->cgDinv.p);
    }
  } else {
    vec_fill_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, 1.0, h->cgDinv.p);
  }
  IKB_LAUNCH_CHECK(h);
  // p = Minv r ; absNew = r.p ; |b|^2
  vec_mul_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, h->cgDinv.p, h->cgR.p, h->cgP.p);
  IKB_LAUNCH_CHECK(h);
  if ((rc = deviceDot(h, 1, h->cgR.p, h->cgP.p, n, scal + 0, 1.0, nullptr))) return rc;
  if ((rc = deviceDot(h, 2, h->cgR.p, nullptr, n, scal + 1, 1.0, nullptr))) return rc;
  if ((rc = fetch(scal, 2))) return rc;
  double absNew = hostS[0];
  const double rhsNorm = std::sqrt(hostS[1]);
  const double tiny = std::numeric_limits<double>::min();
  const double tol = info->tol > 0.0 ? info->tol : std::numeric_limits<double>::epsilon();
  const int64_t maxIters = info->max_iters > 0 ? info->max_iters : 2 * n;
  const double Delta = info->delta;
  double resNorm = rhsNorm;
  int64_t i = 0;
  int stop = IKB_TCG_MAXIMUM_INNER_ITERATIONS;
  bool solved = rhsNorm <= tiny;  // x = 0 (:87-92)
  const double threshold = std::max(tol * tol * rhsNorm * rhsNorm, tiny);
  if (!solved && resNorm * resNorm < threshold) solved = true;  // :94-99
  if (dbc == IKB_DBC_FULL && !solved) {
    // Sync-free path (tcg2_* kernels): the whole loop of the reference, stopping rules included, runs on the device;
    // batches of iterations are replayed from a CUDA graph and the host reads the state once per batch.
    if (!h->tcgState.p) IKB_CUDA(h, h->tcgState.alloc(256));
    TcgState* st = reinterpret_cast<TcgState*>(h->tcgState.p);
    tcg2_init_kernel<<<1, 1, 0, h->stream>>>(st, scal + 0, scal + 0, tol, Delta, info->kappa, info->mininner,
                                             (long long)maxIters, IKB_TCG_MAXIMUM_INNER_ITERATIONS);
    IKB_LAUNCH_CHECK(h);
    const PatternView P = h->view();
    double* pqPartial = h->scratch.p;
    double* rzPartial = h->scratch.p + MAX_SPMV_BLOCKS;
    const int batch = 16;
    const long long maxItersLL = (long long)maxIters;
    auto enqueueBatch = [&]() {
      for (int b = 0; b < batch; ++b) {
        if (h->dim == 3)
          spmv_node_dot_kernel<3><<<h->spmvBlocks, tpb, 0, h->stream>>>(P, h->vals[dbc].p, h->cgP.p, h->cgQ.p, h->cgP.p,
                                                                        pqPartial, &st->cg);
        else
          spmv_node_dot_kernel<2><<<h->spmvBlocks, tpb, 0, h->stream>>>(P, h->vals[dbc].p, h->cgP.p, h->cgQ.p, h->cgP.p,
                                                                        pqPartial, &st->cg);
        tcg2_update_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, st, pqPartial, h->spmvBlocks, h->cgP.p, h->cgQ.p,
                                                              h->cgDinv.p, h->cgX.p, h->cgR.p, h->cgZ.p, rzPartial);
        tcg2_direction_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, st, rzPartial, RED_BLOCKS, maxItersLL, h->cgZ.p,
                                                                 h->cgP.p);
      }
    };
    if (!h->tcgGraph || h->tcgGraphKey[0] != h->vals[dbc].p || h->tcgGraphKey[1] != h->cgP.p ||
        h->tcgGraphKey[2] != h->nbrIdx.p || h->tcgGraphMaxIters != maxItersLL) {
      if (h->tcgGraph) cudaGraphExecDestroy(h->tcgGraph);
      h->tcgGraph = nullptr;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        enqueueBatch();
        if (cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && graph) {
          if (cudaGraphInstantiate(&h->tcgGraph, graph, 0) != cudaSuccess) h->tcgGraph = nullptr;
          cudaGraphDestroy(graph);
        }
      }
      cudaGetLastError();
      h->tcgGraphKey[0] = h->vals[dbc].p;
      h->tcgGraphKey[1] = h->cgP.p;
      h->tcgGraphKey[2] = h->nbrIdx.p;
      h->tcgGraphMaxIters = maxItersLL;
    }
    TcgState hsT;
    while (true) {
      if (h->tcgGraph) {
        IKB_CUDA(h, cudaGraphLaunch(h->tcgGraph, h->stream));
      } else {
        enqueueBatch();
      }
      h->launches += 3 * batch;
      IKB_CUDA(h, cudaGetLastError());
      IKB_CUDA(h, cudaMemcpyAsync(&hsT, st, sizeof(TcgState), cudaMemcpyDeviceToHost, h->stream));
      IKB_CUDA(h, cudaStreamSynchronize(h->stream));
      if (hsT.cg.done) break;
    }
    if (hsT.cg.done == 2) return fail(h, IKB_ECUDA, "tCG produced NaN");
    i = hsT.i;
    stop = hsT.stop;
    resNorm = std::sqrt(hsT.cg.rr);
  } else if (!solved) {
    double e_Pd = 0.0, e_Pe = 0.0, d_Pd = absNew;
    double* partial = h->scratch.p + MAX_SPMV_BLOCKS;
    i = 1;
    while (i < maxIters) {
      if ((rc = launchSpmv(h, dbc, h->cgP.p, h->cgQ.p))) return rc;
      if ((rc = deviceDot(h, 1, h->cgP.p, h->cgQ.p, n, scal + 0, 1.0, nullptr))) return rc;
      if ((rc = fetch(scal, 1))) return rc;
      const double d_Hd = hostS[0];
      const double alpha = absNew / d_Hd;
      const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
      if (d_Hd <= 0 || e_Pe_new >= Delta * Delta) {  // negative curvature or trust region exceeded (:122-134)
        const double tau = (-e_Pd + std::sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
        vec_axpy_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, tau, h->cgP.p, h->cgX.p);
        IKB_LAUNCH_CHECK(h);
        stop = d_Hd <= 0 ? IKB_TCG_NEGATIVE_CURVATURE : IKB_TCG_EXCEEDED_TRUST_REGION;
        break;
      }
      e_Pe = e_Pe_new;
      tcg_update_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, alpha, h->cgP.p, h->cgQ.p, h->cgDinv.p, h->cgX.p, h->cgR.p,
                                                          h->cgZ.p, partial);
      IKB_LAUNCH_CHECK(h);
      cg_fold2_kernel<<<1, tpb, 0, h->stream>>>(partial, RED_BLOCKS, scal, nullptr);  // scal[2] = r.z, scal[3] = r.r
      IKB_LAUNCH_CHECK(h);
      if ((rc = fetch(scal + 2, 2))) return rc;
      resNorm = std::sqrt(hostS[1]);
      if (!(resNorm == resNorm)) return fail(h, IKB_ECUDA, "tCG produced NaN");
      if (i >= info->mininner && resNorm <= rhsNorm * std::min(rhsNorm, info->kappa)) {  // :142-150
        stop = info->kappa < rhsNorm ? IKB_TCG_REACHED_KAPPA_LINEAR : IKB_TCG_REACHED_THETA_SUPERLINEAR;
        break;
      }
      // the reference compares the NORM with the squared-norm threshold here (:151); kept as is
      if (resNorm < threshold) break;
      const double absOld = absNew;
      absNew = hostS[0];
      const double beta = absNew / absOld;
      e_Pd = beta * (e_Pd + alpha * d_Pd);
      d_Pd = absNew + beta * beta * d_Pd;
      tcg_direction_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(n, beta, h->cgZ.p, h->cgP.p);
      IKB_LAUNCH_CHECK(h);
      ++i;
    }
  }
  info->iterations = i;
  info->stop_reason = stop;
  info->rel_error = rhsNorm > tiny ? resNorm / rhsNorm : 0.0;
  // model terms: |eta|, g.eta = -b.eta, eta.H eta
  if ((rc = launchSpmv(h, dbc, h->cgX.p, h->cgQ.p))) return rc;
  if ((rc = deviceDot(h, 2, h->cgX.p, nullptr, n, scal + 0, 1.0, nullptr))) return rc;
  if ((rc = deviceDot(h, 1, h->cgB.p, h->cgX.p, n, scal + 1, 1.0, nullptr))) return rc;
  if ((rc = fetch(scal, 2))) return rc;
  info->eta_norm = std::sqrt(hostS[0]);
  info->g_dot_eta = -hostS[1];
  if ((rc = deviceDot(h, 1, h->cgX.p, h->cgQ.p, n, scal + 0, 1.0, nullptr))) return rc;
  if ((rc = fetch(scal, 1))) return rc;
  info->eta_h_eta = hostS[0];
  if (dbc == IKB_DBC_REDUCED) {
    expand_reduced_kernel<<<gridFor(h->nDof, tpb), tpb, 0, h->stream>>>(h->nDof, h->flags.p, h->cbelow.p, h->cgX.p,
                                                                        h->Corr.p);
    IKB_LAUNCH_CHECK(h);
  } else {
    IKB_CUDA(h, cudaMemcpyAsync(h->Corr.p, h->cgX.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  if (x) IKB_CUDA(h, cudaMemcpyAsync(x, h->cgX.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  return IKB_OK;
}

int ikb_update_solution(ikb_handle hh, int dbc, const double* correction) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!dbcValid(dbc)) return fail(h, IKB_EINVAL, "bad dbc");
  int rc = ensureSolution(h);
  if (rc) return rc;
  const int tpb = 256;
  if (h->Corr.n < (size_t)h->nDof) {
    if (!correction) return fail(h, IKB_ESTATE, "no resident correction");
    IKB_CUDA(h, h->Corr.alloc((size_t)h->nDof));
  }
  if (correction) {
    if (dbc == IKB_DBC_REDUCED) {
      if ((rc = ensureReduced(h))) return rc;
      if (h->cgX.n < (size_t)h->nRed) IKB_CUDA(h, h->cgX.alloc((size_t)std::max<int64_t>(h->nRed, 1)));
      IKB_CUDA(h, cudaMemcpyAsync(h->cgX.p, correction, (size_t)h->nRed * sizeof(double), cudaMemcpyHostToDevice,
                                  h->stream));
      expand_reduced_kernel<<<gridFor(h->nDof, tpb), tpb, 0, h->stream>>>(h->nDof, h->flags.p, h->cbelow.p, h->cgX.p,
                                                                          h->Corr.p);
      IKB_LAUNCH_CHECK(h);
    } else {
      IKB_CUDA(h, cudaMemcpyAsync(h->Corr.p, correction, (size_t)h->nDof * sizeof(double), cudaMemcpyHostToDevice,
                                  h->stream));
    }
  }
  joinSolution(h);
  vec_axpy_kernel<<<RED_BLOCKS, tpb, 0, h->stream>>>(h->nDof, 1.0, h->Corr.p, h->U.p);
  IKB_LAUNCH_CHECK(h);
  markSolutionUse(h);
  h->stateVersion++;
  return IKB_OK;
}

int ikb_nccl_unique_id(void* id128) {
  if (!id128) return IKB_EINVAL;
  std::string err;
  if (!nccl().load(err)) return IKB_ENCCL;
  NcclId id;
  if (nccl().getUniqueId(&id) != 0) return IKB_ENCCL;
  std::memcpy(id128, id.internal, 128);
  return IKB_OK;
}

int ikb_comm_init(ikb_handle hh, const void* id128, int rank, int nranks) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!id128 || rank < 0 || rank >= nranks) return fail(h, IKB_EINVAL, "bad communicator arguments");
  if (!h->meshUploaded) return fail(h, IKB_ESTATE, "upload the mesh and set the row ownership first");
  if (h->layout != LAYOUT_INTERLEAVED) return fail(h, IKB_ENOTIMPL, "row partitioning needs FlatInterleaved dofs");
  std::string err;
  if (!nccl().load(err)) return fail(h, IKB_ENCCL, err);
  NcclId id;
  std::memcpy(id.internal, id128, 128);
  int rc = nccl().commInitRank(&h->comm, nranks, id, rank);
  if (rc != 0) return fail(h, IKB_ENCCL, std::string("ncclCommInitRank: ") + nccl().errorString(rc));
  h->rank = rank;
  h->nranks = nranks;
  // everybody learns everybody's owned rows and touched columns
  DevBuf<int64_t> mine, all;
  IKB_CUDA(h, mine.alloc(4));
  IKB_CUDA(h, all.alloc((size_t)4 * nranks));
  const int64_t r4[4] = {h->rowBegin, h->rowEnd, h->colBegin, h->colEnd};
  IKB_CUDA(h, cudaMemcpyAsync(mine.p, r4, sizeof(r4), cudaMemcpyHostToDevice, h->stream));
  rc = nccl().allGather(mine.p, all.p, 4, NCCL_INT64, h->comm, h->stream);
  if (rc != 0) return fail(h, IKB_ENCCL, std::string("ncclAllGather: ") + nccl().errorString(rc));
  h->peerRanges.assign((size_t)4 * nranks, 0);
  IKB_CUDA(h, cudaMemcpyAsync(h->peerRanges.data(), all.p, all.bytes(), cudaMemcpyDeviceToHost, h->stream));
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  mine.release();
  all.release();
  return IKB_OK;
}

int ikb_comm_ipc_export(ikb_handle hh, void* handles128) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!handles128) return fail(h, IKB_EINVAL, "null handle buffer");
  if (!h->comm) return fail(h, IKB_ESTATE, "ikb_comm_init first");
  if (!h->peerEnabled || h->nranks > PEER_MAXR) return fail(h, IKB_ENOTIMPL, "peer-memory transport disabled or too many ranks");
  if (h->cgPglob.n < (size_t)h->nDof) {
    IKB_CUDA(h, h->cgPglob.alloc((size_t)h->nDof));
    IKB_CUDA(h, cudaMemsetAsync(h->cgPglob.p, 0, h->cgPglob.bytes(), h->stream));
  }
  if (!h->peerWin.p) {
    IKB_CUDA(h, h->peerWin.alloc(256));
    IKB_CUDA(h, cudaMemsetAsync(h->peerWin.p, 0, h->peerWin.bytes(), h->stream));
  }
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaIpcMemHandle_t hp, hw;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  IKB_CUDA(h, cudaIpcGetMemHandle(&hp, h->cgPglob.p));
  IKB_CUDA(h, cudaIpcGetMemHandle(&hw, h->peerWin.p));
  std::memcpy(handles128, &hp, 64);
  std::memcpy(static_cast<char*>(handles128) + 64, &hw, 64);
  return IKB_OK;
}

static void closePeers(Handle* h) {
  for (void* p : h->peerOpened) cudaIpcCloseMemHandle(p);
  h->peerOpened.clear();
  h->peerReady = false;
  if (h->peerGraph) cudaGraphExecDestroy(h->peerGraph);
  h->peerGraph = nullptr;
}

int ikb_comm_ipc_import(ikb_handle hh, const void* all, int* ok) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (ok) *ok = 0;
  closePeers(h);
  if (!all) return IKB_OK;  // back to the NCCL transport
  if (!h->comm || !h->peerWin.p || h->nranks > PEER_MAXR) return fail(h, IKB_ESTATE, "ikb_comm_ipc_export first");
  if (!h->peerCommHost) h->peerCommHost = new PeerComm();
  PeerComm& PC = *reinterpret_cast<PeerComm*>(h->peerCommHost);
  std::memset(&PC, 0, sizeof(PC));
  PC.rank = h->rank;
  PC.nranks = h->nranks;
  const char* rec = static_cast<const char*>(all);
  for (int s = 0; s < h->nranks; ++s) {
    if (s == h->rank) {
      PC.peerP[s] = h->cgPglob.p;
      PC.peerW[s] = h->peerWin.p;
      continue;
    }
    cudaIpcMemHandle_t hp, hw;
    std::memcpy(&hp, rec + (size_t)s * 128, 64);
    std::memcpy(&hw, rec + (size_t)s * 128 + 64, 64);
    void *pp = nullptr, *pw = nullptr;
    if (cudaIpcOpenMemHandle(&pp, hp, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&pw, hw, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      if (pp) cudaIpcCloseMemHandle(pp);
      closePeers(h);
      return IKB_OK;  // *ok stays 0: the caller falls back to NCCL on every rank
    }
    h->peerOpened.push_back(pp);
    h->peerOpened.push_back(pw);
    PC.peerP[s] = static_cast<double*>(pp);
    PC.peerW[s] = static_cast<unsigned long long*>(pw);
  }
  // who reads which of my rows, and whose rows my boundary rows read
  const int D = h->dim;
  const int64_t* me = h->peerRanges.data() + 4 * h->rank;
  for (int s = 0; s < h->nranks; ++s) {
    if (s == h->rank) continue;
    const int64_t* pr = h->peerRanges.data() + 4 * s;
    int64_t sb, se, rb, re;
    haloIntervals(me[0], me[1], me[2], me[3], pr[0], pr[1], pr[2], pr[3], sb, se, rb, re);
    if (se > sb) {
      PC.sendPeer[PC.nSend] = s;
      PC.sendBegin[PC.nSend] = (long long)D * sb;
      PC.sendEnd[PC.nSend] = (long long)D * se;
      PC.nSend++;
    }
    if (re > rb) PC.recvPeer[PC.nRecv++] = s;
  }
  h->peerFirstInterior = h->peerEndInterior = -1;
  h->peerReady = true;
  if (ok) *ok = 1;
  return IKB_OK;
}

int ikb_halo_intervals(const int64_t* mine4, const int64_t* peer4, int64_t* out4) {
  if (!mine4 || !peer4 || !out4) return IKB_EINVAL;
  haloIntervals(mine4[0], mine4[1], mine4[2], mine4[3], peer4[0], peer4[1], peer4[2], peer4[3], out4[0], out4[1], out4[2],
                out4[3]);
  return IKB_OK;
}

int ikb_halo_exchange(ikb_handle hh, const char* what) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!what) return fail(h, IKB_EINVAL, "null array name");
  const std::string w(what);
  double* v = nullptr;
  if (w == "solution") {
    joinSolution(h);
    v = h->U.p;
  } else if (w == "correction")
    v = h->Corr.p;
  else
    return fail(h, IKB_EINVAL, "unknown array");
  if (!v) return fail(h, IKB_ESTATE, "array not allocated");
  const int rc = haloExchange(h, v);
  if (v == h->U.p) markSolutionUse(h);
  return rc;
}

int ikb_stream(ikb_handle hh, void** s) {
  Handle* h = H(hh);
  if (!h || !s) return IKB_EINVAL;
  *s = reinterpret_cast<void*>(h->stream);
  return IKB_OK;
}
int ikb_sync(ikb_handle hh) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  IKB_CUDA(h, cudaStreamSynchronize(h->stream2));
  return checkMaterialError(h);
}
int ikb_launch_count(ikb_handle hh, int64_t* n) {
  Handle* h = H(hh);
  if (!h || !n) return IKB_EINVAL;
  *n = h->launches;
  return IKB_OK;
}
int ikb_device_ptr(ikb_handle hh, const char* what, int dbc, void** ptr) {
  Handle* h = H(hh);
  if (!h || !what || !ptr || !dbcValid(dbc)) return IKB_EINVAL;
  const std::string w(what);
  if (w == "solution") {
    joinSolution(h);
    h->solutionShared = true;  // the caller may use U on the main stream behind our back
    *ptr = h->U.p;
  } else if (w == "residual")
    *ptr = h->vec[dbc].p;
  else if (w == "values")
    *ptr = h->vals[dbc].p;
  else if (w == "correction")
    *ptr = h->Corr.p;
  else
    return fail(h, IKB_EINVAL, "unknown array");
  return IKB_OK;
}

int ikb_time_phase(ikb_handle hh, const char* phase, int dbc, int reps, float* ms) {
  Handle* h = H(hh);
  if (checkHandle(h)) return IKB_EINVAL;
  if (!phase || !ms || reps <= 0 || !dbcValid(dbc)) return fail(h, IKB_EINVAL, "bad arguments");
  const std::string p(phase);
  cudaEvent_t e0, e1;
  IKB_CUDA(h, cudaEventCreate(&e0));
  IKB_CUDA(h, cudaEventCreate(&e1));
  int rc = IKB_OK;
  DevBuf<double> tmp;
  if (p == "dfma_peak" || p == "dmma_peak") IKB_CUDA(h, tmp.alloc((size_t)148 * 16 * 256));
  if (p == "elements" || p == "gather" || p == "fused") {
    if ((rc = ikb_assemble(hh, IKB_MATRIX | IKB_VECTOR, dbc))) return rc;
    if (p == "fused" && !h->fusedOk) return fail(h, IKB_ESTATE, "the fused sweep does not serve this handle");
    if (p != "fused" && (rc = ensureStaging(h, IKB_MATRIX | IKB_VECTOR))) return rc;
    if (p == "gather" && h->fusedOk && (rc = launchElements(h, IKB_MATRIX | IKB_VECTOR))) return rc;
  } else if (p == "spmv") {
    if (h->valsVersion[dbc] != h->stateVersion) return fail(h, IKB_ESTATE, "assemble first");
    const int64_t n = std::max(rowsOf(h, dbc), h->nDof);
    if (h->cgP.n < (size_t)n) IKB_CUDA(h, h->cgP.alloc((size_t)n));
    if (h->cgQ.n < (size_t)n) IKB_CUDA(h, h->cgQ.alloc((size_t)n));
    IKB_CUDA(h, cudaMemsetAsync(h->cgP.p, 0, h->cgP.bytes(), h->stream));
  }
  IKB_CUDA(h, cudaStreamSynchronize(h->stream));
  IKB_CUDA(h, cudaEventRecord(e0, h->stream));
  for (int r = 0; r < reps && rc == IKB_OK; ++r) {
    if (p == "elements")
      rc = launchElements(h, IKB_MATRIX | IKB_VECTOR);
    else if (p == "gather")
      rc = launchGather(h, IKB_MATRIX | IKB_VECTOR, dbc);
    else if (p == "fused")
      rc = launchFused(h, IKB_MATRIX | IKB_VECTOR, dbc);
    else if (p == "spmv")
      rc = launchSpmv(h, dbc, h->cgP.p, h->cgQ.p);
    else if (p == "dfma_peak") {
      dfma_peak_kernel<<<148 * 16, 256, 0, h->stream>>>(tmp.p, 2048);
      h->launches++;
    } else if (p == "dmma_peak") {
      // 8 tiles x 2048 DMMAs per warp, 256 FMA each
      dmma_peak_kernel<<<148 * 16, 256, 0, h->stream>>>(tmp.p, 2048);
      h->launches++;
    } else
      rc = fail(h, IKB_EINVAL, "unknown phase");
  }
  IKB_CUDA(h, cudaEventRecord(e1, h->stream));
  IKB_CUDA(h, cudaEventSynchronize(e1));
  float t = 0.f;
  IKB_CUDA(h, cudaEventElapsedTime(&t, e0, e1));
  *ms = t / reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  tmp.release();
  return rc;
}

}  // extern "C"
