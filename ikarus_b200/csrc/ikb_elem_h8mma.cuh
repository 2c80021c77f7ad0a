// Hex8 element kernel with the tangent contraction on the FP64 tensor cores (DMMA m8n8k4).
//
// Same reference functions, same inputs and the same staged outputs as elem_q1_kernel<3, FORM, true>
// (NonLinearElastic / LinearElastic ::calculateMatrixImpl / calculateVectorImpl / calculateScalarImpl,
// ikarus/finiteelements/mechanics/nonlinearelastic.hh:376-430, linearelastic.hh:354-405, NeoHooke
// materials/hyperelastic/neohooke.hh:79-142); only the work decomposition of the contraction differs.
//
// With the factored tangent of ikb_elem_q1.cuh,  K_e[(a,i),(b,j)] = sum_g c1_g m_a[i] m_b[j] + c2_g m_b[i] m_a[j]
// (+ mu lap_ab delta_ij from the table), write V_c = [m_a[c] at Gauss point g] as an 8 x 8 (node x Gauss point)
// matrix per component c.  For a fixed component pair (i,j) the 8 x 8 matrix over NODE pairs is
//   K^{ij} = V_i diag(c1) V_j^T + V_j diag(c2) V_i^T        = [V_i | V_j] (8 x 16) . [c1 V_j^T ; c2 V_i^T] (16 x 8)
// i.e. four m8n8k4 DMMAs.  The A operand of lane l is V_c[l>>2][4h + (l&3)], the B operand is the SAME register scaled
// by c1 / c2 of that Gauss point, so a lane needs 6 + 4 doubles per element from shared memory instead of the ~32
// loads per Gauss point of the FMA formulation (which ncu showed bound by shared-memory wavefronts, not by the FP64
// pipe).  Only the six pairs i <= j are contracted (24 DMMAs); K^{ji} is the node-transpose of K^{ij} and comes from
// the owning lane by shuffle.  Then lane (a = l>>2, q = l&3) holds the complete 3x3 blocks of the node pairs (a, 2q)
// and (a, 2q+1).
//
// Phase 1 (kinematics + material per Gauss point) keeps the thread-per-Gauss-point layout: a warp evaluates four
// elements at once, lane = 8*el + g.  R_e is reduce-scattered over the 8 Gauss-point lanes of an element with
// shuffles (fixed tree => deterministic), E_e likewise.
#pragma once
#include "ikb_elem_q1.cuh"

namespace ikb {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

struct H8Cfg {
  static constexpr int GPS = 28;                // doubles per Gauss-point record: V (24), c1, c2, 2 pad
  static constexpr int HS = 4 * GPS + 2;        // stride between the two Gauss-point halves (== 2 mod 16)
  static constexpr int ES = 2 * HS + 1;         // element stride (odd): phase-1 stores and phase-2 loads conflict-free
  static constexpr int REC = 4 * ES + (4 * ES) % 2;  // records of the warp's four elements (kept 16-byte aligned)
  // packed K_e for the coalesced write-out: rows k = 0..4 of the pair table (p = k*8 + a), row stride 74 doubles
  // instead of 72 so that the 8-byte stores of a half-warp (lanes (a,q): k = (2q+r-a)&7) hit 16 distinct banks
  static constexpr int KROW = 74;
  static constexpr int KBUF = 4 * KROW + 36;
  static constexpr int WARP_DOUBLES = REC + KBUF;
  static constexpr int WARPS = 4;               // warps per CTA (no block-level synchronisation anywhere)
  static constexpr size_t SMEM = (size_t)WARPS * WARP_DOUBLES * 8;
};

// Where the element results go: symmetric-packed K_e (36 blocks) and R_e of element e at slot e of the staging arrays,
// or at slot e mod ringElems when the staging arrays are a ring (pipelined sweep, ikb_fused.cuh).
struct H8Out {
  double* Kst = nullptr;
  double* Rst = nullptr;
  int64_t ringElems = 0;  // 0: one slot per element
};

// One warp, four elements e0..e0+3 (those >= elemCount are skipped).  wsm: WARP_DOUBLES doubles of shared memory
// private to the warp.
template <int FORM>
__device__ __forceinline__ void h8_warp_elements(const ElemArgs& A, const H8Out& O, int64_t e0, double* wsm) {
  constexpr int D = 3, N = 8;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int el = lane >> 3, t = lane & 7;
  const int64_t eMine = e0 + el;
  const bool active = eMine < A.elemCount;
  const int64_t e = active ? eMine : A.elemCount - 1;  // clamped: inactive lanes still take part in the shuffles

  // ------------------------------------------------------------------ phase 1: Gauss point t of element e
  {
    // the Laplacian table entries are consumed after the DMMAs: pull them into L2 now
    if (active && (A.what & IKB_MATRIX)) {
#pragma unroll
      for (int k = 0; k <= 4; ++k)
        if (!(k == 4 && t >= 4)) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.Lap + (size_t)(k * N + t) * A.nElem + e));
    }
    const double lo = 0.5 - 0.28867513459481287, hi = 0.5 + 0.28867513459481287;
    double xi[D], om[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      xi[k] = ((t >> k) & 1) ? hi : lo;
      om[k] = 1.0 - xi[k];
    }
    double dN[N][D];
#pragma unroll
    for (int c = 0; c < N; ++c) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double v = ((c >> i) & 1) ? 1.0 : -1.0;
#pragma unroll
        for (int k = 0; k < D; ++k)
          if (k != i) v *= ((c >> k) & 1) ? xi[k] : om[k];
        dN[c][i] = v;
      }
    }
    double Jt[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) Jt[i][k] = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c) {
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double x = __ldg(A.X + (size_t)(c * D + k) * A.nElem + e);
#pragma unroll
        for (int i = 0; i < D; ++i) Jt[i][k] = fma(dN[c][i], x, Jt[i][k]);
      }
    }
    double Ji[D][D];
    const double detJ = fabs(invSmall<D>(Jt, Ji));
    const double w = detJ * 0.125;

    double g[N][D];
    double H[D][D];
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int j = 0; j < D; ++j) H[c][j] = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s = fma(Ji[j][i], dN[a][i], s);
        g[a][j] = s;
      }
      const int64_t node = __ldg(A.elemNode + (size_t)a * A.nElem + e);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double u = __ldg(A.U + dofOf(A.layout, D, A.nNodes, node, c));
#pragma unroll
        for (int j = 0; j < D; ++j) H[c][j] = fma(u, g[a][j], H[c][j]);
      }
    }

    const double lam = A.lambda, mu = A.mu;
    double wP[D][D], Am[D][D];
    double c1, c2, psiw;
    if constexpr (FORM == FORM_LE) {
      // eps = sym(H); sigma = lam tr(eps) I + 2 mu eps   (linearelastic.hh:158-177)
      const double tr = H[0][0] + H[1][1] + H[2][2];
      double psi = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const double eps = 0.5 * (H[i][j] + H[j][i]);
          const double sig = 2.0 * mu * eps + (i == j ? lam * tr : 0.0);
          wP[i][j] = w * sig;
          psi = fma(eps, sig, psi);
          Am[i][j] = (i == j) ? 1.0 : 0.0;
        }
      c1 = lam * w;
      c2 = mu * w;
      psiw = 0.5 * psi * w;
    } else {
      // FORM_NH  (neohooke.hh:79-142 with C = 2E + I)
      double F[D][D], Cm[D][D], Ci[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) F[i][j] = H[i][j] + (i == j ? 1.0 : 0.0);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) s = fma(F[k][i], F[k][j], s);
          Cm[i][j] = s;
        }
      const double detC = invSmall<D>(Cm, Ci);
      // the reference aborts at Dune::FloatCmp::le(detC, 0, 1e-10) (materials/materialhelpers.hh:120-126); FloatCmp's
      // default style is relativeWeak, |a-b| <= eps*max(|a|,|b|), which against b = 0 holds only for a == 0: detC <= 0
      if (active && !(detC > 0.0))
        atomicMin(A.errFlag, (int32_t)(e + A.elemBegin < 0x7fffffff ? e + A.elemBegin : 0x7ffffffe));
      const double lnJ = 0.5 * log(detC);
      const double mup = mu - lam * lnJ;
      const double trC = Cm[0][0] + Cm[1][1] + Cm[2][2];
      psiw = (0.5 * mu * (trC - 3.0 - 2.0 * lnJ) + 0.5 * lam * lnJ * lnJ) * w;
      double Sm[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          Sm[i][j] = (i == j ? mu : 0.0) - mup * Ci[i][j];
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) s = fma(F[i][k], Ci[k][j], s);
          Am[i][j] = s;  // F C^-1 = F^-T
        }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) s = fma(F[i][k], Sm[k][j], s);
          wP[i][j] = w * s;
        }
      c1 = lam * w;
      c2 = mup * w;
    }

    if (A.what & IKB_MATRIX) {
      // record of Gauss point t: V[c][a] = m_a[c], then c1, c2
      double* gp = wsm + el * H8Cfg::ES + (t >> 2) * H8Cfg::HS + (t & 3) * H8Cfg::GPS;
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double s;
          if constexpr (FORM == FORM_LE) {
            s = g[a][i];
          } else {
            s = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s = fma(Am[i][j], g[a][j], s);
          }
          gp[i * N + a] = s;
        }
      gp[24] = c1;
      gp[25] = c2;
    }

    if (A.what & IKB_SCALAR) {
      double s = psiw;
      s += __shfl_xor_sync(FULL, s, 4);
      s += __shfl_xor_sync(FULL, s, 2);
      s += __shfl_xor_sync(FULL, s, 1);
      if (active && t == 0) A.Est[e] = s;
    }

    if (A.what & IKB_VECTOR) {
      // R_a += wP g_a at this Gauss point, then reduce-scatter over the element's 8 lanes: lane t ends with node t
      double r[N][D];
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) s = fma(wP[i][j], g[a][j], s);
          r[a][i] = s;
        }
      const bool h4 = t & 4, h2 = t & 2, h1 = t & 1;
      double s1[4][D], s2[2][D], s3[D];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const double send = h4 ? r[a][i] : r[a + 4][i];
          const double keep = h4 ? r[a + 4][i] : r[a][i];
          s1[a][i] = keep + __shfl_xor_sync(FULL, send, 4);
        }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const double send = h2 ? s1[a][i] : s1[a + 2][i];
          const double keep = h2 ? s1[a + 2][i] : s1[a][i];
          s2[a][i] = keep + __shfl_xor_sync(FULL, send, 2);
        }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double send = h1 ? s2[0][i] : s2[1][i];
        const double keep = h1 ? s2[1][i] : s2[0][i];
        s3[i] = keep + __shfl_xor_sync(FULL, send, 1);
      }
      if (active) {
        const int64_t slot = O.ringElems ? e % O.ringElems : e;
#pragma unroll
        for (int i = 0; i < D; ++i) O.Rst[(size_t)slot * (N * D) + t * D + i] = s3[i];
      }
    }
  }
  if (!(A.what & IKB_MATRIX)) return;
  __syncwarp();

  // ------------------------------------------------------------------ phase 2: one element at a time on the tensor cores
  const int a = lane >> 2, q = lane & 3;
  double* kbuf = wsm + H8Cfg::REC;
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    const int64_t ej = e0 + j;
    if (ej >= A.elemCount) break;  // warp-uniform
    const double* rec = wsm + j * H8Cfg::ES + q * H8Cfg::GPS;
    double v[D][2], vb1[D][2], vb2[D][2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double c1 = rec[h * H8Cfg::HS + 24], c2 = rec[h * H8Cfg::HS + 25];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        v[c][h] = rec[h * H8Cfg::HS + c * N + a];
        vb1[c][h] = c1 * v[c][h];
        vb2[c][h] = c2 * v[c][h];
      }
    }
    double acc[D][D][2];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int jj = 0; jj < D; ++jj) acc[i][jj][0] = acc[i][jj][1] = 0.0;
    // component pairs i <= j only (24 DMMAs): K^{ji} = (K^{ij})^T as 8 x 8 node matrices, i.e. K^{ji}[a][b] = K^{ij}[b][a],
    // which lane (b, a>>1) holds in accumulator register a&1 -- fetched with shuffles below
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int jj = i; jj < D; ++jj) {
          dmma884(acc[i][jj][0], acc[i][jj][1], v[i][h], vb1[jj][h]);
          dmma884(acc[i][jj][0], acc[i][jj][1], v[jj][h], vb2[i][h]);
        }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int srcLane = 4 * (2 * q + r) + (a >> 1);
#pragma unroll
      for (int i = 1; i < D; ++i)
#pragma unroll
        for (int jj = 0; jj < i; ++jj) {
          const double t0 = __shfl_sync(0xffffffffu, acc[jj][i][0], srcLane);
          const double t1 = __shfl_sync(0xffffffffu, acc[jj][i][1], srcLane);
          acc[i][jj][r] = (a & 1) ? t1 : t0;
        }
    }
    // lane (a,q) now holds the blocks (a, 2q) and (a, 2q+1)
    __syncwarp();  // kbuf of the previous element has been copied out
    {
      // the packed K_e keeps pair p = k*8 + a' for block (a', (a'+k) mod 8), k <= 4 (k == 4: a' < 4 only)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * q + r;
        const int k = (b - a) & 7;
        if (k < 4 || (k == 4 && a < 4)) {
          const double lap = A.mu * __ldg(A.Lap + (size_t)(k * N + a) * A.nElem + ej);
          double* dst = kbuf + k * H8Cfg::KROW + a * 9;
#pragma unroll
          for (int i = 0; i < D; ++i)
#pragma unroll
            for (int jj = 0; jj < D; ++jj) {
              // (the diagonal block is exactly symmetric by construction: its lower entries are the shuffled upper ones)
              double x = acc[i][jj][r];
              if (i == jj) x += lap;
              dst[i * D + jj] = x;
            }
        }
      }
      __syncwarp();
      const double2* src2 = reinterpret_cast<const double2*>(kbuf);
      double2* dst2 = reinterpret_cast<double2*>(O.Kst + (size_t)(O.ringElems ? ej % O.ringElems : ej) * 36 * 9);
#pragma unroll
      for (int it = 0; it < 6; ++it) {
        const int idx2 = lane + 32 * it;  // 162 double2: rows k < 4 hold 36, row 4 holds 18
        if (idx2 < 162) {
          const int k = idx2 / 36;
          dst2[idx2] = src2[idx2 + k];  // row stride 37 double2 in kbuf
        }
      }
    }
  }
}

template <int FORM, int MINB>
__global__ void __launch_bounds__(32 * H8Cfg::WARPS, MINB) elem_h8_mma_kernel(ElemArgs A) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5;
  double* wsm = smem + (size_t)warp * H8Cfg::WARP_DOUBLES;
  const int64_t e0 = ((int64_t)blockIdx.x * H8Cfg::WARPS + warp) * 4;
  if (e0 >= A.elemCount) return;
  H8Out O;
  O.Kst = A.Kst;
  O.Rst = A.Rst;
  h8_warp_elements<FORM>(A, O, e0, wsm);
}

// minBlocks = resident CTAs per SM the register allocation is held to: 4 -> 128 registers, 5 -> 96, 6 -> 80 (spills)
template <int FORM, int MINB>
cudaError_t launchElemH8MmaImpl(const ElemArgs& A, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(elem_h8_mma_kernel<FORM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)H8Cfg::SMEM);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((A.elemCount + 4 * H8Cfg::WARPS - 1) / (4 * H8Cfg::WARPS));
  if (grid == 0) return cudaSuccess;
  elem_h8_mma_kernel<FORM, MINB><<<grid, 32 * H8Cfg::WARPS, H8Cfg::SMEM, st>>>(A);
  return cudaGetLastError();
}
template <int FORM>
cudaError_t launchElemH8Mma(const ElemArgs& A, cudaStream_t st, int minBlocks) {
  if (minBlocks == 4) return launchElemH8MmaImpl<FORM, 4>(A, st);
  if (minBlocks == 6) return launchElemH8MmaImpl<FORM, 6>(A, st);
  return launchElemH8MmaImpl<FORM, 5>(A, st);
}

// FP64 tensor-core peak probe: 8 independent m8n8k4 accumulator tiles per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = threadIdx.x * 1e-9 + k;
  const double a = 1.0000001, b = 1e-7 * (threadIdx.x & 3);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) dmma884(c[k][0], c[k][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace ikb
