// Fused K_e / R_e / E_e kernel for displacement-based Q1 elements (Quad4, Hex8).
//
// Replaces the per-element work of NonLinearElastic / LinearElastic
// ::calculateMatrixImpl / calculateVectorImpl / calculateScalarImpl
// (ikarus/finiteelements/mechanics/nonlinearelastic.hh:376-430,
//  ikarus/finiteelements/mechanics/linearelastic.hh:354-405) and the material calls below
// them (materials/svk.hh:77-164, materials/hyperelastic/neohooke.hh:79-142,
// materials/vanishingstrain.hh:79-120 for plane strain).
//
// Work decomposition: N = 2^D threads per element (N nodes == N Gauss points).
//   phase 1  thread g evaluates Gauss point g: Jacobian, physical shape gradients g_a,
//            deformation gradient F, strain, stress, and the per-node vectors the pair
//            formula needs; results go to a shared-memory record of the element.
//   phase 2  thread a owns the node pairs (a, (a+k) mod N), k = 0..N/2, accumulates their
//            DxD blocks over all Gauss points in registers, then writes the symmetric-packed
//            K_e, its rows of R_e, and (thread 0) E_e to the staging arrays.
//
// The reference contracts B_a^T C B_b with a 6x6 Voigt tangent for every node pair.  All three
// in-scope laws have the isotropic structure  CC = l' X(x)X + 2 m' sym(X(.)X)  (X = I for
// LinearElasticity/SVK, X = C^-1 for NeoHooke), so with  m_a = F X g_a :
//   B_a^T C B_b + (g_a.S g_b) I = l' m_a m_b^T + m' m_b m_a^T + m' (g_a.X g_b) F X F^T + (g_a.S g_b) I
// which for displacement-based NeoHooke (F X F^T = I, S = mu I - m' C^-1) collapses to
//   l m_a m_b^T + m' m_b m_a^T + mu (g_a.g_b) I.
// Same result to rounding, ~4x fewer flops than the Voigt contraction.
#pragma once
#include "ikb_internal.cuh"

namespace ikb {

struct ElemArgs {
  const double* X;         // [nc*D][nElem]
  const int32_t* elemNode; // [N][nElem]
  const double* U;         // [nDof]
  double* Kst;             // [nElem][NPAIR][D*D]
  double* Rst;             // [nElem][N*D]
  double* Est;             // [nElem]
  const double* Lap;       // [NPAIR][nElem] sum_g w_g grad N_a . grad N_b (reference geometry only), may be null
  int32_t* errFlag;
  int64_t nElem;
  // Q1 kernels can run an element sub-range: the launcher shifts every per-element pointer by elemBegin and the kernel
  // sees elements [0, elemCount) with the SoA stride nElem unchanged (elemBegin is only used to report error ids).
  int64_t elemBegin, elemCount;
  int64_t nNodes;
  int layout;
  double lambda, mu;
  unsigned what;
  int planeStress;  // 2D: Materials::planeStress instead of planeStrain
  double psTol;     // tolerance of the stress reduction
};

// VanishingStress::reduceStress for planeStress (materials/vanishingstress.hh:150-199): Newton-Raphson with tolerance
// tol on |S33|, at most 100 iterations, correction = -(A^-1 r), on the out-of-plane normal strain only (the fixed
// shear strains stay zero).  Linear / St.Venant-Kirchhoff law: unknown E33 started from 0 (derivative lambda + 2 mu).
__device__ __forceinline__ bool reduceE33(double lam, double mu, double tol, double tr2, double& e33) {
  e33 = 0.0;
  double f;
  for (int it = 0;; ++it) {
    f = lam * (tr2 + e33) + 2.0 * mu * e33;
    if (!(fabs(f) > tol && it < 100)) break;
    e33 += -((1.0 / (lam + 2.0 * mu)) * f);
  }
  return !(fabs(f) > tol);
}
// NeoHooke works on the right Cauchy-Green tensor: unknown C33 started from 1 (initUnknownStrains :137-148), residual
// S33 = mu (1 - 1/C33) + lambda ln J / C33, derivative = moduli_3333 / 2 with moduli_3333 = (lambda + 2 mu') / C33^2.
// det2 = det of the in-plane C.  Also returns ln J of the solution.
__device__ __forceinline__ bool reduceC33(double lam, double mu, double tol, double det2, double& c33, double& lnJ) {
  c33 = 1.0;
  lnJ = 0.0;
  double f;
  for (int it = 0;; ++it) {
    const double detC = det2 * c33;
    if (!(detC > 0.0)) return false;  // checkPositiveOrAbort: relativeWeak compare against 0 == (detC <= 0)
    lnJ = log(sqrt(detC));
    f = mu * (1.0 - 1.0 / c33) + lam * lnJ / c33;
    if (!(fabs(f) > tol && it < 100)) break;
    const double df = (lam + 2.0 * (mu - lam * lnJ)) / (c33 * c33) / 2.0;
    c33 += -((1.0 / df) * f);
  }
  return !(fabs(f) > tol);
}
// lambda of the condensed in-plane tangent: staticCondensation over the fixed Voigt indices {2,3,4}
// (utils/linearalgebrahelper.hh:521-535) leaves C_red = lambda_red X(x)X + 2 mu' I_X with
// lambda_red = lambda - lambda^2 / (lambda + 2 mu')   (mu' = mu for the linear laws).
__device__ __forceinline__ double condensedLambda(double lam, double mup) { return lam - lam * lam / (lam + 2.0 * mup); }

template <int D>
__device__ __forceinline__ double invSmall(const double (&A)[D][D], double (&Ai)[D][D]) {
  if constexpr (D == 2) {
    const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    const double id = 1.0 / det;
    Ai[0][0] = A[1][1] * id;
    Ai[0][1] = -A[0][1] * id;
    Ai[1][0] = -A[1][0] * id;
    Ai[1][1] = A[0][0] * id;
    return det;
  } else {
    const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    const double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    const double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    const double id = 1.0 / det;
    Ai[0][0] = c00 * id;
    Ai[1][0] = c01 * id;
    Ai[2][0] = c02 * id;
    Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
    Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
    Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
    Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
    return det;
  }
}

template <int D, int FORM>
struct Q1Cfg {
  static constexpr int N = 1 << D;
  static constexpr int DD = D * D;
  static constexpr int SYM = D * (D + 1) / 2;
  static constexpr int NV = (FORM == FORM_LE) ? 1 : 2;  // per-node vectors in the record: m (and g)
  static constexpr int NPAIR = N * (N + 1) / 2;
  static constexpr int KMAX = N / 2;                    // pair offsets k = 0..KMAX
  // per-Gauss-point scalars
  static constexpr int O_C1 = 0, O_C2 = 1;
  static constexpr int O_C3 = 2;                                          // NH only
  static constexpr int O_A2 = 2, O_WS = 2 + SYM;                          // SVK only
  static constexpr int O_WP = (FORM == FORM_LE) ? 2 : (FORM == FORM_NH ? 3 : 2 + 2 * SYM);
  static constexpr int O_PSI = O_WP + DD;
  static constexpr int NS = O_PSI + 1;
  static constexpr int VEC = NV * D * N;
  static constexpr int GPS0 = VEC + NS;
  static constexpr int GPS = GPS0 + ((5 - GPS0 % 4) % 4);   // == 1 (mod 4): conflict-free phase-1 stores
  static constexpr int BS = blockStride(D);
  static constexpr int S0 = (N * GPS > NPAIR * BS) ? N * GPS : NPAIR * BS;  // also holds the packed K_e for write-out
  static constexpr int S = S0 + ((N % 16) - (S0 % 16) + 16) % 16;  // == N (mod 16): conflict-free phase-2 loads
  static constexpr int SMEM_BUDGET = 110 * 1024;
  static constexpr int EPW = 32 / N;                               // elements per warp
  static constexpr int EPC0 = SMEM_BUDGET / (S * 8);
  static constexpr int EPC1 = (EPC0 / EPW) * EPW;
  static constexpr int EPC = EPC1 > 256 / N ? 256 / N : EPC1;      // elements per CTA (<= 256 threads)
  static constexpr int TPB = EPC * N;
  static constexpr size_t SMEM = (size_t)EPC * S * 8;
};

// index of the symmetric entry (i,j) in row-major upper packing
template <int D>
__device__ __forceinline__ constexpr int symIdx(int i, int j) {
  if (i > j) {
    const int t = i;
    i = j;
    j = t;
  }
  return i * D - i * (i - 1) / 2 + (j - i);
}

template <int D, int FORM, bool USELAP>
__global__ void __launch_bounds__(Q1Cfg<D, FORM>::TPB, (Q1Cfg<D, FORM>::SMEM > 56 * 1024) ? 2 : 4)
    elem_q1_kernel(ElemArgs A) {
  using C = Q1Cfg<D, FORM>;
  constexpr int N = C::N, DD = C::DD;
  extern __shared__ double smem[];

  const int tid = threadIdx.x;
  const int el = tid / N;     // element within CTA
  const int t = tid % N;      // Gauss point (phase 1) / row node (phase 2)
  const int64_t e = (int64_t)blockIdx.x * C::EPC + el;
  const bool active = e < A.elemCount;
  double* rec = smem + (size_t)el * C::S;
  const unsigned grpMask = (N == 32) ? 0xffffffffu : (((1u << N) - 1u) << ((threadIdx.x & 31u) / N * N));

  // ------------------------------------------------------------------ phase 1
  if (active) {
    if constexpr (USELAP) {
      // the Laplacian table entries of this thread (node t) are needed only after the Gauss loop of phase 2: pull them
      // into L2 now (no registers held) instead of paying a DRAM latency at the very end
#pragma unroll
      for (int k = 0; k <= C::KMAX; ++k)
        if (!(k == C::KMAX && t >= N / 2))
          asm volatile("prefetch.global.L2 [%0];" ::"l"(A.Lap + (size_t)(k * N + t) * A.nElem + e));
    }
    const double lo = 0.5 - 0.28867513459481287, hi = 0.5 + 0.28867513459481287;
    double xi[D], om[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      xi[k] = ((t >> k) & 1) ? hi : lo;
      om[k] = 1.0 - xi[k];
    }
    // reference-cell shape gradients at this Gauss point
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
    // Jt[i][k] = dx_k/dxi_i
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
    double w = detJ;
#pragma unroll
    for (int k = 0; k < D; ++k) w *= 0.5;

    double* gp = rec + t * C::GPS;
    double* vM = gp;                            // m_a  [c][node]
    double* vG = gp + (C::NV - 1) * D * N;      // g_a  [c][node] (aliases m for LE)
    double* sc = gp + C::VEC;

    // physical gradients g_a[j] = sum_i Ji[j][i] dN[a][i]; displacement gradient H[c][j]
    double H[D][D];
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int j = 0; j < D; ++j) H[c][j] = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a) {
      double g[D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s = fma(Ji[j][i], dN[a][i], s);
        g[j] = s;
        vG[j * N + a] = s;
      }
      const int64_t node = __ldg(A.elemNode + (size_t)a * A.nElem + e);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double u = __ldg(A.U + dofOf(A.layout, D, A.nNodes, node, c));
#pragma unroll
        for (int j = 0; j < D; ++j) H[c][j] = fma(u, g[j], H[c][j]);
      }
    }

    const double lam = A.lambda, mu = A.mu;
    if constexpr (FORM == FORM_LE) {
      // eps = sym(H); sigma = lam tr(eps) I + 2 mu eps   (linearelastic.hh:158-177, svk.hh)
      double tr = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) tr += H[i][i];
      double lamR = lam;
      if constexpr (D == 2) {
        if (A.planeStress) {  // sigma_33 = 0: eps_33 joins the trace, condensed lambda (psi = sigma:eps/2 is unchanged)
          double e33;
          if (!reduceE33(lam, mu, A.psTol, tr, e33))
            atomicMin(A.errFlag, (int32_t)(e + A.elemBegin < 0x7fffffff ? e + A.elemBegin : 0x7ffffffe));
          tr += e33;
          lamR = condensedLambda(lam, mu);
        }
      }
      double psi = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const double eps = 0.5 * (H[i][j] + H[j][i]);
          const double sig = 2.0 * mu * eps + (i == j ? lam * tr : 0.0);
          sc[C::O_WP + i * D + j] = w * sig;
          psi = fma(eps, sig, psi);
        }
      sc[C::O_C1] = lamR * w;
      sc[C::O_C2] = mu * w;
      sc[C::O_PSI] = 0.5 * psi * w;
    } else {
      double F[D][D], Cm[D][D];
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
      double Sm[D][D];
      double Am[D][D];  // m_a = Am g_a
      double psi;
      if constexpr (FORM == FORM_SVK) {
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) tr += 0.5 * (Cm[i][i] - 1.0);
        double ee = 0.0, lamR = lam;
        if constexpr (D == 2) {
          if (A.planeStress) {
            double e33;
            if (!reduceE33(lam, mu, A.psTol, tr, e33))
              atomicMin(A.errFlag, (int32_t)(e + A.elemBegin < 0x7fffffff ? e + A.elemBegin : 0x7ffffffe));
            tr += e33;
            ee = e33 * e33;
            lamR = condensedLambda(lam, mu);
          }
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double E = 0.5 * (Cm[i][j] - (i == j ? 1.0 : 0.0));
            Sm[i][j] = 2.0 * mu * E + (i == j ? lam * tr : 0.0);
            ee = fma(E, E, ee);
            Am[i][j] = F[i][j];
          }
        psi = 0.5 * lam * tr * tr + mu * ee;
        sc[C::O_C1] = lamR * w;
        sc[C::O_C2] = mu * w;
        // A2 = mu w F F^T (sym), wS = w S (sym)
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = i; j < D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(F[i][k], F[j][k], s);
            sc[C::O_A2 + symIdx<D>(i, j)] = mu * w * s;
            sc[C::O_WS + symIdx<D>(i, j)] = w * Sm[i][j];
          }
      } else {  // FORM_NH  (neohooke.hh:79-142 with C = 2E + I)
        double Ci[D][D];
        const double detC = invSmall<D>(Cm, Ci);
        if (!(detC > 0.0)) atomicMin(A.errFlag, (int32_t)(e + A.elemBegin < 0x7fffffff ? e + A.elemBegin : 0x7ffffffe));
        double lnJ = 0.5 * log(detC);
        double c33 = 1.0;  // plane strain: C_33 = 1
        bool planeStress = false;
        if constexpr (D == 2) {
          if (A.planeStress) {
            planeStress = true;
            if (!reduceC33(lam, mu, A.psTol, detC, c33, lnJ))
              atomicMin(A.errFlag, (int32_t)(e + A.elemBegin < 0x7fffffff ? e + A.elemBegin : 0x7ffffffe));
          }
        }
        const double mup = mu - lam * lnJ;
        double trC = (D == 2) ? c33 : 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) trC += Cm[i][i];
        psi = 0.5 * mu * (trC - 3.0 - 2.0 * lnJ) + 0.5 * lam * lnJ * lnJ;
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
        sc[C::O_C1] = (planeStress ? condensedLambda(lam, mup) : lam) * w;
        sc[C::O_C2] = mup * w;
        sc[C::O_C3] = mu * w;
      }
      sc[C::O_PSI] = psi * w;
      // wP = w F S
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) s = fma(F[i][k], Sm[k][j], s);
          sc[C::O_WP + i * D + j] = w * s;
        }
      // m_a = Am g_a
#pragma unroll
      for (int a = 0; a < N; ++a) {
        double g[D];
#pragma unroll
        for (int j = 0; j < D; ++j) g[j] = vG[j * N + a];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) s = fma(Am[i][j], g[j], s);
          vM[i * N + a] = s;
        }
      }
    }
  }
  __syncwarp();  // an element's N lanes never straddle a warp

  // ------------------------------------------------------------------ phase 2
  if (!active) return;
  constexpr int NK = C::KMAX + 1;
  double acc[NK][DD];
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int q = 0; q < DD; ++q) acc[k][q] = 0.0;
  double Ra[D];
#pragma unroll
  for (int i = 0; i < D; ++i) Ra[i] = 0.0;

  const int a = t;
#pragma unroll 2
  for (int g = 0; g < N; ++g) {
    const double* gp = rec + g * C::GPS;
    const double* vM = gp;
    const double* vG = gp + (C::NV - 1) * D * N;
    const double* sc = gp + C::VEC;
    const double c1 = sc[C::O_C1], c2 = sc[C::O_C2];
    double ma[D], ga[D], p1[D], p2[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      ga[i] = vG[i * N + a];
      ma[i] = (FORM == FORM_LE) ? ga[i] : vM[i * N + a];
      p1[i] = c1 * ma[i];
      p2[i] = c2 * ma[i];
    }
    // residual rows of node a: R_a += wP g_a
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) Ra[i] = fma(sc[C::O_WP + i * D + j], ga[j], Ra[i]);

    double hs[D], sg[D];  // third/fourth-term helpers of node a
    if constexpr (USELAP) {
      // mu * sum_g w_g (g_a.g_b) does not depend on the displacements: it comes from the precomputed table
    } else if constexpr (FORM == FORM_LE) {
#pragma unroll
      for (int i = 0; i < D; ++i) hs[i] = c2 * ga[i];
    } else if constexpr (FORM == FORM_NH) {
      const double c3 = sc[C::O_C3];
#pragma unroll
      for (int i = 0; i < D; ++i) hs[i] = c3 * ga[i];
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(sc[C::O_WS + symIdx<D>(i, j)], ga[j], s);
        sg[i] = s;
      }
    }
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int b = (a + k) & (N - 1);
      double mb[D], gb[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if constexpr (!USELAP || FORM == FORM_LE) gb[i] = vG[i * N + b];
        mb[i] = (FORM == FORM_LE) ? gb[i] : vM[i * N + b];
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) acc[k][i * D + j] = fma(p1[i], mb[j], fma(mb[i], p2[j], acc[k][i * D + j]));
      if constexpr (USELAP) {
      } else if constexpr (FORM == FORM_SVK) {
        double cab = 0.0, sab = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          cab = fma(ga[i], gb[i], cab);
          sab = fma(sg[i], gb[i], sab);
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j)
            acc[k][i * D + j] = fma(cab, sc[C::O_A2 + symIdx<D>(i, j)], acc[k][i * D + j]);
#pragma unroll
        for (int i = 0; i < D; ++i) acc[k][i * D + i] += sab;
      } else {
        double dab = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) dab = fma(hs[i], gb[i], dab);
#pragma unroll
        for (int i = 0; i < D; ++i) acc[k][i * D + i] += dab;
      }
    }
  }

  if constexpr (USELAP) {
    const double mu = A.mu;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      if (k == C::KMAX && a >= N / 2) break;
      const double lap = mu * __ldg(A.Lap + (size_t)(k * N + a) * A.nElem + e);
#pragma unroll
      for (int i = 0; i < D; ++i) acc[k][i * D + i] += lap;
    }
  }

  // ------------------------------------------------------------------ write-out
  // E_e first (it reads the records), then the record area is recycled to transpose K_e so that the element's
  // N lanes store 16-byte vectors to consecutive addresses (full 32-byte sectors instead of 8-byte pieces).
  if ((A.what & IKB_SCALAR) && a == 0) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < N; ++g) s += rec[g * C::GPS + C::VEC + C::O_PSI];
    A.Est[e] = s;
  }
  if (A.what & IKB_MATRIX) {
    __syncwarp(grpMask);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      if (k == C::KMAX && a >= N / 2) break;
      double* dst = rec + (k * N + a) * C::BS;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          // diagonal block: mirror the upper triangle so K_e is exactly symmetric
          dst[i * D + j] = (k == 0 && i > j) ? acc[0][j * D + i] : acc[k][i * D + j];
        }
    }
    __syncwarp(grpMask);
    constexpr int NV2 = C::NPAIR * C::BS / 2;
    const double2* src2 = reinterpret_cast<const double2*>(rec);
    double2* dst2 = reinterpret_cast<double2*>(A.Kst + (size_t)e * C::NPAIR * C::BS);
#pragma unroll 4
    for (int idx = a; idx < NV2; idx += N) dst2[idx] = src2[idx];
  }
  if (A.what & IKB_VECTOR) {
#pragma unroll
    for (int i = 0; i < D; ++i) A.Rst[(size_t)e * (N * D) + a * D + i] = Ra[i];
  }
}

template <int D, int FORM, bool USELAP>
cudaError_t launchElemQ1Impl(const ElemArgs& A, cudaStream_t st) {
  using C = Q1Cfg<D, FORM>;
  // (the opt-in is per device and context: made at every launch, a handle may live on any GPU of the process)
  cudaError_t e = cudaFuncSetAttribute(elem_q1_kernel<D, FORM, USELAP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::SMEM);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((A.elemCount + C::EPC - 1) / C::EPC);
  if (grid == 0) return cudaSuccess;
  elem_q1_kernel<D, FORM, USELAP><<<grid, C::TPB, C::SMEM, st>>>(A);
  return cudaGetLastError();
}

template <int D, int FORM>
cudaError_t launchElemQ1(const ElemArgs& A, cudaStream_t st) {
  if constexpr (FORM != FORM_SVK) {
    if (A.Lap) return launchElemQ1Impl<D, FORM, true>(A, st);
  }
  return launchElemQ1Impl<D, FORM, false>(A, st);
}

// One-time table for the displacement-independent part of the tangent of Q1 elements:
//   Lap[p][e] = sum_g w_g detJ_g  grad N_a . grad N_b,   p = k*N + a, b = (a+k) mod N
// (the mu (grad N_a . grad N_b) I term of NeoHooke / LinearElasticity, see the header comment).
template <int D>
__global__ void __launch_bounds__(128) lap_q1_kernel(const double* __restrict__ X, int64_t nElem, double* Lap) {
  constexpr int N = 1 << D;
  constexpr int KMAX = N / 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nElem) return;
  double x[N][D];
#pragma unroll
  for (int c = 0; c < N; ++c)
#pragma unroll
    for (int k = 0; k < D; ++k) x[c][k] = X[(size_t)(c * D + k) * nElem + e];
  double lap[KMAX + 1][N];
#pragma unroll
  for (int k = 0; k <= KMAX; ++k)
#pragma unroll
    for (int a = 0; a < N; ++a) lap[k][a] = 0.0;
  const double lo = 0.5 - 0.28867513459481287, hi = 0.5 + 0.28867513459481287;
#pragma unroll 1
  for (int g = 0; g < N; ++g) {
    double xi[D], om[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      xi[k] = ((g >> k) & 1) ? hi : lo;
      om[k] = 1.0 - xi[k];
    }
    double dN[N][D], Jt[D][D], Ji[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) Jt[i][k] = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double v = ((c >> i) & 1) ? 1.0 : -1.0;
#pragma unroll
        for (int k = 0; k < D; ++k)
          if (k != i) v *= ((c >> k) & 1) ? xi[k] : om[k];
        dN[c][i] = v;
#pragma unroll
        for (int k = 0; k < D; ++k) Jt[i][k] = fma(v, x[c][k], Jt[i][k]);
      }
    double w = fabs(invSmall<D>(Jt, Ji));
#pragma unroll
    for (int k = 0; k < D; ++k) w *= 0.5;
    double gr[N][D];
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s = fma(Ji[j][i], dN[a][i], s);
        gr[a][j] = s;
      }
#pragma unroll
    for (int k = 0; k <= KMAX; ++k)
#pragma unroll
      for (int a = 0; a < N; ++a) {
        const int b = (a + k) & (N - 1);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s = fma(gr[a][i], gr[b][i], s);
        lap[k][a] = fma(w, s, lap[k][a]);
      }
  }
#pragma unroll
  for (int k = 0; k <= KMAX; ++k)
#pragma unroll
    for (int a = 0; a < N; ++a)
      if (!(k == KMAX && a >= N / 2)) Lap[(size_t)(k * N + a) * nElem + e] = lap[k][a];
}

}  // namespace ikb
