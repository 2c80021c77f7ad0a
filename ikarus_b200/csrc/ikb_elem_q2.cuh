// Fused K_e / R_e / E_e kernel for displacement-based Q2 elements (Quad9, Hex27: 18 / 81 dofs).
//
// Same two-phase scheme as ikb_elem_q1.cuh (see there for the pair formula and the reference lines it
// replaces): N = 3^D threads per element, thread g evaluates Gauss point g of the 3^D-point rule (order
// 2*2, nonlinearelastic.hh:121-124), thread a then owns the node pairs (a, (a+k) mod N), k = 0..(N-1)/2,
// processed in register-sized passes.  Geometry is the grid element's multilinear map from its 2^D corners
// (nonlinearelastic.hh:117), the displacement field uses the quadratic Lagrange basis with DUNE's
// lexicographic node order.
#pragma once
#include "ikb_elem_q1.cuh"
#include "ikb_internal.cuh"

namespace ikb {

template <int D, int FORM>
struct Q2Cfg {
  static constexpr int N = (D == 3) ? 27 : 9;
  static constexpr int NC = 1 << D;
  static constexpr int DD = D * D;
  static constexpr int SYM = D * (D + 1) / 2;
  // The records keep g_a only.  m_a = A g_a (A = F for SVK, F C^-1 for NeoHooke) is formed in phase 2 from the 3x3 matrix A
  // of the Gauss point: 9 more FMAs per node pair and Gauss point, but the record shrinks from 2*D*N to D*N doubles -- Hex27
  // SVK 40.4 -> 24.8 KB per element -- so TWO CTAs fit an SM (8 warps instead of 4; ncu had the FP64 pipe at 45 % with one
  // warp per scheduler).
  static constexpr int NV = 1;
  static constexpr int NPAIR = N * (N + 1) / 2;
  static constexpr int KMAX = (N - 1) / 2;         // pair offsets k = 0..KMAX, all threads
  static constexpr int KPASS = (D == 3) ? 7 : 5;   // pair offsets per register pass
  static constexpr int NPASS = (KMAX + 1 + KPASS - 1) / KPASS;
  static constexpr int O_C1 = 0, O_C2 = 1, O_C3 = 2;
  static constexpr int O_A2 = 2, O_WS = 2 + SYM;
  static constexpr int O_WP = (FORM == FORM_LE) ? 2 : (FORM == FORM_NH ? 3 : 2 + 2 * SYM);
  static constexpr int O_PSI = O_WP + DD;
  static constexpr int O_AM = O_PSI + 1;                                  // A (row-major), nonlinear forms only
  static constexpr int NS = O_AM + ((FORM == FORM_LE) ? 0 : DD);
  static constexpr int VEC = NV * D * N;
  static constexpr int GPS0 = VEC + NS;
  static constexpr int GPS = GPS0 + (1 - GPS0 % 2);  // odd stride: conflict-free phase-1 stores
  static constexpr int S = N * GPS;
  static constexpr int EPW = 32 / N;                 // 1 (Hex27) or 3 (Quad9) elements per warp
  static constexpr int WARPS = (D == 3) ? 4 : 4;
  static constexpr int EPC = EPW * WARPS;
  static constexpr int TPB = 32 * WARPS;
  static constexpr size_t SMEM = (size_t)EPC * S * 8;
};

template <int D, int FORM>
__global__ void __launch_bounds__(Q2Cfg<D, FORM>::TPB) elem_q2_kernel(ElemArgs A) {
  using C = Q2Cfg<D, FORM>;
  constexpr int N = C::N, DD = C::DD, NC = C::NC;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int elw = lane / N;  // element within warp
  const int t = lane - elw * N;
  const int el = warp * C::EPW + elw;
  const int64_t e = (int64_t)blockIdx.x * C::EPC + el;
  const bool active = (elw < C::EPW) && (e < A.nElem);
  double* rec = smem + (size_t)el * C::S;

  if (active) {
    // 3-point Gauss rule on [0,1]
    const double gx[3] = {0.5 - 0.5 * 0.7745966692414834, 0.5, 0.5 + 0.5 * 0.7745966692414834};
    const double gw[3] = {5.0 / 18.0, 4.0 / 9.0, 5.0 / 18.0};
    double xi[D], l[D][3], dl[D][3];
    double w = 1.0;
    {
      int gg = t;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const int gk = gg % 3;
        gg /= 3;
        xi[k] = gx[gk];
        w *= gw[gk];
        const double x = xi[k];
        l[k][0] = 2.0 * (x - 0.5) * (x - 1.0);
        l[k][1] = 4.0 * x * (1.0 - x);
        l[k][2] = 2.0 * x * (x - 0.5);
        dl[k][0] = 4.0 * x - 3.0;
        dl[k][1] = 4.0 - 8.0 * x;
        dl[k][2] = 4.0 * x - 1.0;
      }
    }
    // geometry from the corners
    double Jt[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) Jt[i][k] = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double dn[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double v = ((c >> i) & 1) ? 1.0 : -1.0;
#pragma unroll
        for (int k = 0; k < D; ++k)
          if (k != i) v *= ((c >> k) & 1) ? xi[k] : 1.0 - xi[k];
        dn[i] = v;
      }
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double x = __ldg(A.X + (size_t)(c * D + k) * A.nElem + e);
#pragma unroll
        for (int i = 0; i < D; ++i) Jt[i][k] = fma(dn[i], x, Jt[i][k]);
      }
    }
    double Ji[D][D];
    const double detJ = fabs(invSmall<D>(Jt, Ji));
    w *= detJ;

    double* gp = rec + t * C::GPS;
    double* vG = gp;
    double* sc = gp + C::VEC;
    double H[D][D];
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int j = 0; j < D; ++j) H[c][j] = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a) {
      const int a0 = a % 3, a1 = (a / 3) % 3, a2 = a / 9;
      double dn[D];
      if constexpr (D == 3) {
        dn[0] = dl[0][a0] * l[1][a1] * l[2][a2];
        dn[1] = l[0][a0] * dl[1][a1] * l[2][a2];
        dn[2] = l[0][a0] * l[1][a1] * dl[2][a2];
      } else {
        dn[0] = dl[0][a0] * l[1][a1];
        dn[1] = l[0][a0] * dl[1][a1];
      }
      double g[D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s = fma(Ji[j][i], dn[i], s);
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
      double tr = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) tr += H[i][i];
      double lamR = lam;
      if constexpr (D == 2) {
        if (A.planeStress) {  // see elem_q1_kernel
          double e33;
          if (!reduceE33(lam, mu, A.psTol, tr, e33)) atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
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
      double F[D][D], Cm[D][D], Sm[D][D], Am[D][D];
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
      double psi;
      if constexpr (FORM == FORM_SVK) {
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) tr += 0.5 * (Cm[i][i] - 1.0);
        double ee = 0.0, lamR = lam;
        if constexpr (D == 2) {
          if (A.planeStress) {
            double e33;
            if (!reduceE33(lam, mu, A.psTol, tr, e33)) atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
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
      } else {
        double Ci[D][D];
        const double detC = invSmall<D>(Cm, Ci);
        if (!(detC > 0.0)) atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
        double lnJ = 0.5 * log(detC);
        double c33 = 1.0;
        bool planeStress = false;
        if constexpr (D == 2) {
          if (A.planeStress) {
            planeStress = true;
            if (!reduceC33(lam, mu, A.psTol, detC, c33, lnJ)) atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
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
            Am[i][j] = s;
          }
        sc[C::O_C1] = (planeStress ? condensedLambda(lam, mup) : lam) * w;
        sc[C::O_C2] = mup * w;
        sc[C::O_C3] = mu * w;
      }
      sc[C::O_PSI] = psi * w;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) s = fma(F[i][k], Sm[k][j], s);
          sc[C::O_WP + i * D + j] = w * s;
        }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) sc[C::O_AM + i * D + j] = Am[i][j];
    }
  }
  __syncwarp();
  if (!active) return;

  // ------------------------------------------------------------------ phase 2
  const int a = t;
  double Ra[D];
#pragma unroll
  for (int i = 0; i < D; ++i) Ra[i] = 0.0;
  double* Ke = A.Kst + (size_t)e * C::NPAIR * blockStride(D);
#pragma unroll 1
  for (int pass = 0; pass < C::NPASS; ++pass) {
    const int k0 = pass * C::KPASS;
    double acc[C::KPASS][DD];
#pragma unroll
    for (int k = 0; k < C::KPASS; ++k)
#pragma unroll
      for (int q = 0; q < DD; ++q) acc[k][q] = 0.0;
#pragma unroll 1
    for (int g = 0; g < N; ++g) {
      const double* gp = rec + g * C::GPS;
      const double* vG = gp;
      const double* sc = gp + C::VEC;
      const double c1 = sc[C::O_C1], c2 = sc[C::O_C2];
      double Am[D][D];
      if constexpr (FORM != FORM_LE) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) Am[i][j] = sc[C::O_AM + i * D + j];
      }
      double ma[D], ga[D], p1[D], p2[D];
#pragma unroll
      for (int i = 0; i < D; ++i) ga[i] = vG[i * N + a];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if constexpr (FORM == FORM_LE) {
          ma[i] = ga[i];
        } else {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) s = fma(Am[i][j], ga[j], s);
          ma[i] = s;
        }
        p1[i] = c1 * ma[i];
        p2[i] = c2 * ma[i];
      }
      if (pass == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) Ra[i] = fma(sc[C::O_WP + i * D + j], ga[j], Ra[i]);
      }
      double hs[D], sg[D];
      if constexpr (FORM == FORM_LE) {
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
      for (int k = 0; k < C::KPASS; ++k) {
        int b = a + k0 + k;
        if (b >= N) b -= N;
        double mb[D], gb[D];
#pragma unroll
        for (int i = 0; i < D; ++i) gb[i] = vG[i * N + b];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          if constexpr (FORM == FORM_LE) {
            mb[i] = gb[i];
          } else {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s = fma(Am[i][j], gb[j], s);
            mb[i] = s;
          }
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) acc[k][i * D + j] = fma(p1[i], mb[j], fma(mb[i], p2[j], acc[k][i * D + j]));
        if constexpr (FORM == FORM_SVK) {
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
    if (A.what & IKB_MATRIX) {
#pragma unroll
      for (int k = 0; k < C::KPASS; ++k) {
        const int kk = k0 + k;
        if (kk > C::KMAX) break;
        double* dst = Ke + (size_t)(kk * N + a) * blockStride(D);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) dst[i * D + j] = (kk == 0 && i > j) ? acc[k][j * D + i] : acc[k][i * D + j];
      }
    }
  }
  if (A.what & IKB_VECTOR) {
#pragma unroll
    for (int i = 0; i < D; ++i) A.Rst[(size_t)e * (N * D) + a * D + i] = Ra[i];
  }
  if ((A.what & IKB_SCALAR) && a == 0) {
    double s = 0.0;
    for (int g = 0; g < N; ++g) s += rec[g * C::GPS + C::VEC + C::O_PSI];
    A.Est[e] = s;
  }
}

template <int D, int FORM>
cudaError_t launchElemQ2(const ElemArgs& A, cudaStream_t st) {
  using C = Q2Cfg<D, FORM>;
  // (the opt-in is per device and context: made at every launch, a handle may live on any GPU of the process)
  cudaError_t e =
      cudaFuncSetAttribute(elem_q2_kernel<D, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((A.nElem + C::EPC - 1) / C::EPC);
  if (grid == 0) return cudaSuccess;
  elem_q2_kernel<D, FORM><<<grid, C::TPB, C::SMEM, st>>>(A);
  return cudaGetLastError();
}

}  // namespace ikb
