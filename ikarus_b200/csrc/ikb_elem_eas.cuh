// Q1 elements with enhanced assumed strains (Quad4 + E4/E5/E7, Hex8 + E9/E21): fused K_e / R_e with
// static condensation, and the internal-variable update.
//
// Replaces EnhancedAssumedStrains::calculateMatrixImpl / calculateVectorImpl / calculateDAndLMatrix /
// calculateRtilde / updateStateImpl (ikarus/finiteelements/mechanics/enhancedassumedstrains.hh:225-248,
// 258-348, 378-434), the ansatz matrices E4..E21 (strainenhancements/easvariants/linearandglstrains.hh:
// 71-328) and EAS::GreenLagrangeStrain / EAS::LinearStrain (easfunctions/greenlagrangestrain.hh:40-141,
// linearstrain.hh:40-138).
//
// Every ansatz column has ONE non-zero:  M(xi)[:, j] = T0inv[:, r_j] * p_j(xi) / detJ(xi)  with p_j a
// monomial in t = 2 xi - 1.  With Q = w C T0inv, G = T0inv^T Q, TS = w T0inv^T S per Gauss point:
//   D[j][k]  += s_j s_k G[r_j][r_k]      L[j][a,c] += s_j (Q^T B_a)[r_j][c]      Rt[j] += s_j TS[r_j]
// (s_j = p_j/detJ), so the m x m and m x ndof products of the reference collapse to table lookups.
//
// N = 2^D threads per element, one warp per CTA.  Phase 1: thread = Gauss point (records in shared memory).  Pass C:
// thread a = rows {a, a+N, ..} of D and Rt; pass B: the columns of L belonging to node a.  D is then factorised in
// REGISTERS (LDL^T, right-looking: the rows stay with the thread that accumulated them, the pivot row is broadcast with
// shuffles, Rt rides along as one more column), each thread solves its own right-hand sides, L / Y = Lf^-1 L live in the
// record tails that passes B and C have released, and only then pass A accumulates the K_uu node pairs (a, a+k) and
// R_a, which are condensed in registers:  K_ab -= Y_a^T d^-1 Y_b,  R_a -= Y_a^T d^-1 y_R.
#pragma once
#include "ikb_elem_q1.cuh"
#include "ikb_internal.cuh"

namespace ikb {

struct EasArgs {
  ElemArgs E;
  const double* T0inv;  // [S*S][nElem]  (T(center) detJ0)^-1, row-major S x S, element fastest
  double* alpha;        // [nElem][M]
  const double* dU;     // correction (update mode), indexed like U
  int updateMode;       // 0: K/R   1: alpha -= D^-1 (Rt + L du)
  PsLaw ps;             // FORM_PS (generalised-tangent kernel): the principal-stretch law
};

// ansatz tables: row r_j and monomial id per column.  3D monomials: 0 tx, 1 ty, 2 tz, 3 txty, 4 txtz, 5 tytz;
// 2D: 0 tx, 1 ty, 2 txty.
#define IKB_EAS_TABLE(DIM, MM, ...)                                                                  \
  template <>                                                                                        \
  struct EasTable<DIM, MM> {                                                                         \
    __host__ __device__ static constexpr int row(int j) {                                            \
      constexpr int r[MM][2] = {__VA_ARGS__};                                                        \
      return r[j][0];                                                                                \
    }                                                                                                \
    __host__ __device__ static constexpr int mono(int j) {                                           \
      constexpr int r[MM][2] = {__VA_ARGS__};                                                        \
      return r[j][1];                                                                                \
    }                                                                                                \
    /* the table packed 3 bits per column: a run-time column index costs a shift and a mask instead */ \
    /* of a copy of the table on the thread's stack */                                                 \
    __host__ __device__ static constexpr unsigned long long bits(int which) {                        \
      constexpr int r[MM][2] = {__VA_ARGS__};                                                        \
      unsigned long long b = 0;                                                                      \
      for (int j = 0; j < MM; ++j) b |= (unsigned long long)r[j][which] << (3 * j);                  \
      return b;                                                                                      \
    }                                                                                                \
    __host__ __device__ static int rowRt(int j) {                                                    \
      constexpr unsigned long long b = bits(0);                                                      \
      return (int)((b >> (3 * j)) & 7ull);                                                           \
    }                                                                                                \
    __host__ __device__ static int monoRt(int j) {                                                   \
      constexpr unsigned long long b = bits(1);                                                      \
      return (int)((b >> (3 * j)) & 7ull);                                                           \
    }                                                                                                \
  };
template <int D, int M>
struct EasTable;
// {row, monomial} per ansatz column (easvariants/linearandglstrains.hh:84-92, 118-127, 153-164, 262-277, 299-326)
IKB_EAS_TABLE(2, 4, {0, 0}, {1, 1}, {2, 0}, {2, 1})
IKB_EAS_TABLE(2, 5, {0, 0}, {1, 1}, {2, 0}, {2, 1}, {2, 2})
IKB_EAS_TABLE(2, 7, {0, 0}, {1, 1}, {2, 0}, {2, 1}, {0, 2}, {1, 2}, {2, 2})
IKB_EAS_TABLE(3, 9, {0, 0}, {1, 1}, {2, 2}, {3, 1}, {3, 2}, {4, 0}, {4, 2}, {5, 0}, {5, 1})
IKB_EAS_TABLE(3, 21, {0, 0}, {1, 1}, {2, 2}, {3, 1}, {3, 2}, {4, 0}, {4, 2}, {5, 0}, {5, 1}, {3, 3}, {3, 4}, {4, 3},
              {4, 5}, {5, 4}, {5, 5}, {0, 3}, {0, 4}, {1, 3}, {1, 5}, {2, 4}, {2, 5})
#undef IKB_EAS_TABLE

template <int D, int FORM, int M>
struct EasCfg {
  static constexpr int N = 1 << D;
  static constexpr int DD = D * D;
  static constexpr int S = D * (D + 1) / 2;
  static constexpr int ND = N * D;
  static constexpr int NMONO = D == 3 ? 6 : 3;
  static constexpr int NPAIR = N * (N + 1) / 2;
  static constexpr int KMAX = N / 2;
  static constexpr int ROWS = (M + N - 1) / N;  // rows of D / Rt owned per thread
  // per-Gauss-point record
  static constexpr int O_M = 0, O_G = D * N, VEC = 2 * D * N;
  static constexpr int O_C1 = VEC, O_C2 = VEC + 1, O_IDET = VEC + 2;
  static constexpr int O_A2 = O_IDET + 1;   // c2 F X F^T (sym)
  static constexpr int O_WS = O_A2 + S;     // w S (sym)
  static constexpr int O_X = O_WS + S;      // X (sym)
  static constexpr int O_WP = O_X + S;      // w F S
  static constexpr int O_F = O_WP + DD;     // F
  static constexpr int O_Q = O_F + DD;      // w C T0inv   [S][S]
  static constexpr int O_GG = O_Q + S * S;  // T0inv^T Q   [S][S]
  static constexpr int O_TS = O_GG + S * S; // w T0inv^T S [S]
  static constexpr int GPS0 = O_TS + S;
  static constexpr int GPS = GPS0 + ((5 - GPS0 % 4) % 4);
  static constexpr int REC0 = N * GPS;
  // D [M][M] lives behind the records; the condensation scratch reuses the record area
  static constexpr int SCR = M * ND + M + N * M + M;  // Y [M][ND], Rt [M], partial [N][M], reciprocal pivots [M]
  static constexpr int ES0 = (REC0 > SCR ? REC0 : SCR) + M * M;
  static constexpr int ES = ES0 + ((N % 16) - (ES0 % 16) + 16) % 16;
  static constexpr int EPW = 32 / N;
  // Gauss-point loops of passes A, B, C: two points per trip for the small ansaetze (instruction-level parallelism for a
  // kernel that runs one warp per scheduler); E21 has no registers to spare
  static constexpr int GUNROLL = (M <= 9) ? 2 : 1;
  static constexpr int EPC0 = (100 * 1024) / (ES * 8);
  static constexpr int EPC1 = (EPC0 / EPW) * EPW;
  static constexpr int EPC2 = EPC1 > 128 / N ? 128 / N : EPC1;
  // one warp per CTA: the shared memory of an SM then divides into as many resident warps as fit (E9: 5 x 45 KB instead
  // of 2 CTAs of 2 warps; E21: 4 x 56 KB either way) -- the kernel is latency bound, every resident warp counts
  static constexpr int EPC = EPW + 0 * EPC2;
  static constexpr int TPB = EPC * N;
  static constexpr size_t SMEM = (size_t)EPC * ES * 8;
};

template <int D>
__device__ __forceinline__ void voigtPair(int q, int& i, int& j) {
  if constexpr (D == 3) {
    const int I[6] = {0, 1, 2, 1, 0, 0}, J[6] = {0, 1, 2, 2, 2, 1};
    i = I[q];
    j = J[q];
  } else {
    const int I[3] = {0, 1, 0}, J[3] = {0, 1, 1};
    i = I[q];
    j = J[q];
  }
}

template <int D>
__device__ __forceinline__ void monomials(int g, double invDet, double (&s)[D == 3 ? 6 : 3]) {
  const double c = 0.57735026918962576;  // 2*xi-1 at the 2-point Gauss abscissae
  double t[D];
#pragma unroll
  for (int k = 0; k < D; ++k) t[k] = ((g >> k) & 1) ? c : -c;
  if constexpr (D == 3) {
    s[0] = t[0] * invDet;
    s[1] = t[1] * invDet;
    s[2] = t[2] * invDet;
    s[3] = t[0] * t[1] * invDet;
    s[4] = t[0] * t[2] * invDet;
    s[5] = t[1] * t[2] * invDet;
  } else {
    s[0] = t[0] * invDet;
    s[1] = t[1] * invDet;
    s[2] = t[0] * t[1] * invDet;
  }
}

// s[m] for a run-time m without indexing the register array
template <int NM>
__device__ __forceinline__ double pickMono(const double (&s)[NM], int m) {
  double v = s[0];
#pragma unroll
  for (int q = 1; q < NM; ++q) v = (m == q) ? s[q] : v;
  return v;
}

template <int D, int FORM, int M>
__global__ void __launch_bounds__(EasCfg<D, FORM, M>::TPB) elem_eas_kernel(EasArgs EA) {
  using C = EasCfg<D, FORM, M>;
  using T = EasTable<D, M>;
  constexpr int N = C::N, DD = C::DD, S = C::S, ND = C::ND;
  const ElemArgs& A = EA.E;
  extern __shared__ double smem[];

  const int tid = threadIdx.x;
  const int el = tid / N;
  const int t = tid % N;
  const int64_t eRaw = (int64_t)blockIdx.x * C::EPC + el;
  // lane groups past the last element recompute the last element (their global writes are suppressed) so that
  // every warp-level barrier below can be a plain full-warp __syncwarp()
  const bool active = eRaw < A.nElem;
  const int64_t e = active ? eRaw : A.nElem - 1;
  double* rec = smem + (size_t)el * C::ES;

  // ------------------------------------------------------------------ phase 1: Gauss point t
  {
    // The element's nodes and, behind them, its displacements are fetched first (two dependent global loads), then
    // (T(centre) detJ0)^-1: their latency hides behind the geometry
    double ue[N][D];
    {
      int64_t nodes[N];
#pragma unroll
      for (int a = 0; a < N; ++a) nodes[a] = __ldg(A.elemNode + (size_t)a * A.nElem + e);
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c) ue[a][c] = __ldg(A.U + dofOf(A.layout, D, A.nNodes, nodes[a], c));
    }
    double T0[S][S];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int q = 0; q < S; ++q) T0[p][q] = __ldg(EA.T0inv + (size_t)(p * S + q) * A.nElem + e);
    const double lo = 0.5 - 0.28867513459481287, hi = 0.5 + 0.28867513459481287;
    double xi[D], om[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      xi[k] = ((t >> k) & 1) ? hi : lo;
      om[k] = 1.0 - xi[k];
    }
    double dN[N][D];
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double v = ((c >> i) & 1) ? 1.0 : -1.0;
#pragma unroll
        for (int k = 0; k < D; ++k)
          if (k != i) v *= ((c >> k) & 1) ? xi[k] : om[k];
        dN[c][i] = v;
      }
    double Jt[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) Jt[i][k] = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double x = __ldg(A.X + (size_t)(c * D + k) * A.nElem + e);
#pragma unroll
        for (int i = 0; i < D; ++i) Jt[i][k] = fma(dN[c][i], x, Jt[i][k]);
      }
    double Ji[D][D];
    const double detJ = fabs(invSmall<D>(Jt, Ji));
    double w = detJ;
#pragma unroll
    for (int k = 0; k < D; ++k) w *= 0.5;
    const double invDet = 1.0 / detJ;

    double* gp = rec + t * C::GPS;
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
        gp[C::O_G + j * N + a] = s;
      }
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double u = ue[a][c];
#pragma unroll
        for (int j = 0; j < D; ++j) H[c][j] = fma(u, g[j], H[c][j]);
      }
    }
    // compatible strain (Voigt, shear doubled) and F
    double F[D][D];
    double Ev[S];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) F[i][j] = (FORM == FORM_LE ? 0.0 : H[i][j]) + (i == j ? 1.0 : 0.0);
#pragma unroll
    for (int q = 0; q < S; ++q) {
      int i, j;
      voigtPair<D>(q, i, j);
      double v = 0.5 * (H[i][j] + H[j][i]);
      if constexpr (FORM != FORM_LE) {
#pragma unroll
        for (int k = 0; k < D; ++k) v = fma(0.5 * H[k][i], H[k][j], v);
      }
      Ev[q] = (i == j) ? v : 2.0 * v;
    }
    // enhanced strain: E += T0inv * (sum_j s_j alpha_j e_{r_j})
    double sm[C::NMONO];
    monomials<D>(t, invDet, sm);
    {
      double v[S];
#pragma unroll
      for (int q = 0; q < S; ++q) v[q] = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) v[T::row(j)] = fma(sm[T::mono(j)], EA.alpha[(size_t)e * M + j], v[T::row(j)]);
#pragma unroll
      for (int p = 0; p < S; ++p)
#pragma unroll
        for (int q = 0; q < S; ++q) Ev[p] = fma(T0[p][q], v[q], Ev[p]);
    }
    // material at the enhanced strain: S (tensor), X, l', m'
    const double lam = A.lambda, mu = A.mu;
    double Em[D][D], Sm[D][D], X[D][D];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      int i, j;
      voigtPair<D>(q, i, j);
      Em[i][j] = Em[j][i] = (i == j) ? Ev[q] : 0.5 * Ev[q];
    }
    double mup = mu;
    double lamR = lam;  // lambda of the (condensed, for planeStress) tangent
    if constexpr (FORM == FORM_NH) {
      double Cm[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) Cm[i][j] = 2.0 * Em[i][j] + (i == j ? 1.0 : 0.0);
      const double detC = invSmall<D>(Cm, X);
      if (!(detC > 0.0)) atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
      double lnJ = 0.5 * log(detC);
      bool planeStress = false;
      if constexpr (D == 2) {
        if (A.planeStress) {  // see elem_q1_kernel
          planeStress = true;
          double c33;
          if (!reduceC33(lam, mu, A.psTol, detC, c33, lnJ)) atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
        }
      }
      mup = mu - lam * lnJ;
      if (planeStress) lamR = condensedLambda(lam, mup);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) Sm[i][j] = (i == j ? mu : 0.0) - mup * X[i][j];
    } else {
      double tr = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) tr += Em[i][i];
      if constexpr (D == 2) {
        if (A.planeStress) {
          double e33;
          if (!reduceE33(lam, mu, A.psTol, tr, e33)) atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
          tr += e33;
          lamR = condensedLambda(lam, mu);
        }
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          Sm[i][j] = 2.0 * mu * Em[i][j] + (i == j ? lam * tr : 0.0);
          X[i][j] = (i == j) ? 1.0 : 0.0;
        }
    }
    gp[C::O_C1] = lamR * w;
    gp[C::O_C2] = mup * w;
    gp[C::O_IDET] = invDet;
    // Am = F X ; A2 = c2 Am F^T ; m_a = Am g_a
    double Am[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(F[i][k], X[k][j], s);
        Am[i][j] = s;
      }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = i; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(Am[i][k], F[j][k], s);
        gp[C::O_A2 + symIdx<D>(i, j)] = mup * w * s;
        gp[C::O_WS + symIdx<D>(i, j)] = (FORM == FORM_LE) ? 0.0 : w * Sm[i][j];
        gp[C::O_X + symIdx<D>(i, j)] = X[i][j];
      }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(F[i][k], Sm[k][j], s);
        gp[C::O_WP + i * D + j] = w * s;
        gp[C::O_F + i * D + j] = F[i][j];
      }
#pragma unroll
    for (int a = 0; a < N; ++a) {
      double g[D];
#pragma unroll
      for (int j = 0; j < D; ++j) g[j] = gp[C::O_G + j * N + a];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fma(Am[i][j], g[j], s);
        gp[C::O_M + i * N + a] = s;
      }
    }
    // Voigt tangent CC[p][q] = l' X_ij X_kl + m' (X_ik X_jl + X_il X_jk); Q = w CC T0inv; G = T0inv^T Q
    double CC[S][S];
#pragma unroll
    for (int p = 0; p < S; ++p) {
      int i, j;
      voigtPair<D>(p, i, j);
#pragma unroll
      for (int q = p; q < S; ++q) {
        int k, l;
        voigtPair<D>(q, k, l);
        const double v = w * (lamR * X[i][j] * X[k][l] + mup * (X[i][k] * X[j][l] + X[i][l] * X[j][k]));
        CC[p][q] = v;
        CC[q][p] = v;
      }
    }
    double Q[S][S];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int r = 0; r < S; ++r) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < S; ++q) s = fma(CC[p][q], T0[q][r], s);
        Q[p][r] = s;
        gp[C::O_Q + p * S + r] = s;
      }
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int r2 = 0; r2 < S; ++r2) {
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < S; ++p) s = fma(T0[p][r], Q[p][r2], s);
        gp[C::O_GG + r * S + r2] = s;
      }
#pragma unroll
    for (int r = 0; r < S; ++r) {
      double s = 0.0;
#pragma unroll
      for (int p = 0; p < S; ++p) {
        int i, j;
        voigtPair<D>(p, i, j);
        s = fma(T0[p][r], Sm[i][j], s);
      }
      gp[C::O_TS + r] = w * s;
    }
  }
  __syncwarp();

  const int a = t;
  constexpr int NK = C::KMAX + 1;
  double* Dm = rec + (C::REC0 > C::SCR ? C::REC0 : C::SCR);  // [M][M], outside the record area
  // ------------------------------------------------------------------ pass C: rows {a, a+N, ..} of D and Rt
  double Rt[C::ROWS];
#pragma unroll
  for (int jj = 0; jj < C::ROWS; ++jj) {
    const int j = a + jj * N;
    Rt[jj] = 0.0;
    if (j < M) {
      double drow[M];
#pragma unroll
      for (int k = 0; k < M; ++k) drow[k] = 0.0;
      const int rj = T::rowRt(j), mj = T::monoRt(j);
#pragma unroll(C::GUNROLL)
      for (int g = 0; g < N; ++g) {
        const double* gp = rec + g * C::GPS;
        double sm[C::NMONO];
        monomials<D>(g, gp[C::O_IDET], sm);
        const double sj = pickMono(sm, mj);
        const double* Grow = gp + C::O_GG + rj * S;
#pragma unroll
        for (int k = 0; k < M; ++k) drow[k] = fma(sj * sm[T::mono(k)], Grow[T::row(k)], drow[k]);
        Rt[jj] = fma(sj, gp[C::O_TS + rj], Rt[jj]);
      }
#pragma unroll
      for (int k = 0; k < M; ++k) Dm[j * M + k] = drow[k];
    }
  }

  // ------------------------------------------------------------------ pass B: L columns of node a
  double Lr[M][D];
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int c = 0; c < D; ++c) Lr[j][c] = 0.0;
#pragma unroll(C::GUNROLL)
  for (int g = 0; g < N; ++g) {
    const double* gp = rec + g * C::GPS;
    double sm[C::NMONO];
    monomials<D>(g, gp[C::O_IDET], sm);
    double ga[D];
#pragma unroll
    for (int i = 0; i < D; ++i) ga[i] = gp[C::O_G + i * N + a];
    double B[S][D];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      int i, j;
      voigtPair<D>(q, i, j);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double Fci = gp[C::O_F + c * D + i], Fcj = gp[C::O_F + c * D + j];
        B[q][c] = (i == j) ? ga[i] * Fci : fma(ga[j], Fci, ga[i] * Fcj);
      }
    }
    double QB[S][D];
#pragma unroll
    for (int r = 0; r < S; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < S; ++p) s = fma(gp[C::O_Q + p * S + r], B[p][c], s);
        QB[r][c] = s;
      }
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int c = 0; c < D; ++c) Lr[j][c] = fma(sm[T::mono(j)], QB[T::row(j)][c], Lr[j][c]);
  }
  __syncwarp();  // passes B and C done: the slots [O_F, GPS) of every Gauss-point record become condensation scratch

  // Row j of L (overwritten by Y = Lf^-1 L, D = Lf diag(d) Lf^T) lives in the free tail of Gauss-point record j / RPG;
  // Rt and the reciprocal pivots in the tail of the last record; the partial sums of L du (update mode, no pass A)
  // in the head of record b.  Pass A still finds its slots [0, O_F) intact.
  constexpr int RPG = (C::GPS - C::O_F) / ND;  // rows of L per record tail
  static_assert(RPG >= 1 && (M + RPG - 1) / RPG <= N - 1 && 2 * M <= C::GPS - C::O_F && M <= C::O_F, "scratch layout");
#define IKB_LM(j) (rec + ((j) / RPG) * C::GPS + C::O_F + ((j) % RPG) * ND)
#define IKB_PM(b) (rec + (b) * C::GPS)
  double* Rm = rec + (N - 1) * C::GPS + C::O_F;  // [M]  Rt -> Lf^-1 Rt (assembly) / D^-1 (Rt + L du) (update)
#pragma unroll
  for (int jj = 0; jj < C::ROWS; ++jj) {
    const int j = a + jj * N;
    if (j < M) Rm[j] = Rt[jj];
  }
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int c = 0; c < D; ++c) IKB_LM(j)[a * D + c] = Lr[j][c];
  if (EA.updateMode) {
    // partial_j = sum_c L[j][a,c] du[a,c]
    const int64_t node = __ldg(A.elemNode + (size_t)a * A.nElem + e);
    double du[D];
#pragma unroll
    for (int c = 0; c < D; ++c) du[c] = __ldg(EA.dU + dofOf(A.layout, D, A.nNodes, node, c));
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) s = fma(Lr[j][c], du[c], s);
      IKB_PM(a)[j] = s;
    }
  }
  __syncwarp();
  if (EA.updateMode) {
    // Rt_j + (L du)_j, node contributions added in node order
#pragma unroll
    for (int jj = 0; jj < C::ROWS; ++jj) {
      const int j = a + jj * N;
      if (j < M) {
        double s = Rm[j];
        for (int b = 0; b < N; ++b) s += IKB_PM(b)[j];
        Rm[j] = s;
      }
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------ LDL^T (D = Lf diag(d) Lf^T)
  double* invd = Rm + M;  // [M] reciprocal pivots
#ifndef IKB_EAS_SMEM_LDL
  // In registers: thread a keeps rows {a, a+N, ..} of D (the rows it accumulated in pass C), full rows so that by
  // symmetry row k carries what the elimination of column k needs; per pivot the owner's row is broadcast inside the
  // element's N-lane group with shuffles and every thread updates its own rows.  No shared-memory traffic, no barrier
  // and no cross-thread dependency other than the shuffles: the 21 dependent steps of E21 cost ~150 cycles each instead
  // of a shared-memory round trip per entry (the shared-memory factorisation below held 45 % of the kernel's stall
  // samples at 4 warps per SM).  Same operations per entry as the right-looking elimination below.
  {
    double Dr[C::ROWS][M];
#pragma unroll
    for (int jj = 0; jj < C::ROWS; ++jj)
#pragma unroll
      for (int k = 0; k < M; ++k) Dr[jj][k] = (a + jj * N < M) ? Dm[(a + jj * N) * M + k] : 0.0;
    // the right-hand side rides along as one more column: y_R = Lf^-1 Rt comes out of the same elimination
    double rr[C::ROWS];
#pragma unroll
    for (int jj = 0; jj < C::ROWS; ++jj) rr[jj] = (a + jj * N < M) ? Rm[a + jj * N] : 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const int o = k % N, jo = k / N;  // owner lane (inside the element's group) and its local row
      double v[M];
#pragma unroll
      for (int j = k; j < M; ++j) v[j] = __shfl_sync(0xffffffffu, Dr[jo][j], o, N);
      const double yk = __shfl_sync(0xffffffffu, rr[jo], o, N);
      const double idk = __drcp_rn(v[k]);  // correctly rounded like 1.0 / x, without the division's slow path
      if (a == 0) invd[k] = idk;
#pragma unroll
      for (int jj = 0; jj < C::ROWS; ++jj) {
        if (jj * N + N - 1 <= k) continue;  // every row of this slot is at or above the pivot
        const int i = a + jj * N;
        // rows at or above the pivot (and the padding rows past M) take lik = 0: no divergent branch in the chain
        const bool below = i > k && i < M;
        const double lik = below ? Dr[jj][k] * idk : 0.0;
#pragma unroll
        for (int j = k + 1; j < M; ++j) Dr[jj][j] = fma(-lik, v[j], Dr[jj][j]);
        Dr[jj][k] = below ? lik : Dr[jj][k];
        rr[jj] = fma(-lik, yk, rr[jj]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < C::ROWS; ++jj) {
      const int i = a + jj * N;
      if (i < M) {
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k < i) Dm[i * M + k] = Dr[jj][k];
        Rm[i] = rr[jj];
      }
    }
    __syncwarp();
  }
#else
  // ------------------------------------------------------------------ cooperative LDL^T (lower triangle, in place)
  // The reciprocal of pivot k+1 is computed by the owner of row k+1 while the others scale column k, so no
  // thread waits on a division inside the elimination, and the solves below need no divisions at all.
  if (a == 0) invd[0] = 1.0 / Dm[0];
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k < M; ++k) {
    const double idk = invd[k];
    for (int i = k + 1 + a; i < M; i += N) {
      const double lik = Dm[i * M + k] * idk;
      // Row i is updated by this thread only and column k is not written in this loop, so the loads of a batch of
      // four entries can all be issued before the first store (the compiler cannot prove that through the shared
      // pointer and would serialise load -> fma -> store per entry).  Same operations, same order per entry.
      const double* colk = Dm + k;
      double* rowi = Dm + i * M;
      for (int j = k + 1; j <= i; j += 4) {
        const bool p1 = j + 1 <= i, p2 = j + 2 <= i, p3 = j + 3 <= i;  // the last batch of a row may be short
        const double c0 = colk[j * M], c1 = p1 ? colk[(j + 1) * M] : 0.0, c2 = p2 ? colk[(j + 2) * M] : 0.0,
                     c3 = p3 ? colk[(j + 3) * M] : 0.0;
        const double d0 = rowi[j], d1 = p1 ? rowi[j + 1] : 0.0, d2 = p2 ? rowi[j + 2] : 0.0,
                     d3 = p3 ? rowi[j + 3] : 0.0;
        rowi[j] = fma(-lik, c0, d0);
        if (p1) rowi[j + 1] = fma(-lik, c1, d1);
        if (p2) rowi[j + 2] = fma(-lik, c2, d2);
        if (p3) rowi[j + 3] = fma(-lik, c3, d3);
      }
    }
    __syncwarp();
    for (int i = k + 1 + a; i < M; i += N) Dm[i * M + k] *= idk;
    if (k + 1 < M && ((k + 1) % N) == a) invd[k + 1] = 1.0 / Dm[(k + 1) * M + k + 1];
    __syncwarp();
  }
#endif
  // ------------------------------------------------------------------ solves
  // With D = Lf diag(d) Lf^T:  L^T D^-1 L = Y^T diag(1/d) Y  and  L^T D^-1 Rt = Y^T diag(1/d) y_R  with
  // Y = Lf^-1 L, y_R = Lf^-1 Rt, so the assembly only needs FORWARD substitutions.  Thread a does the D columns
  // of node a together, fully unrolled in registers (each factor entry is loaded once for D columns).
  if (!EA.updateMode) {
    double z[M][D];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) z[i][c] = IKB_LM(i)[a * D + c];  // = Lr, not kept live through the factorisation
#pragma unroll
    for (int i = 1; i < M; ++i)
#pragma unroll
      for (int j = 0; j < i; ++j) {
        const double l = Dm[i * M + j];
#pragma unroll
        for (int c = 0; c < D; ++c) z[i][c] = fma(-l, z[j][c], z[i][c]);
      }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) IKB_LM(i)[a * D + c] = z[i][c];
  }
#ifndef IKB_EAS_SMEM_LDL
  if (a == N - 1 && EA.updateMode) {  // the forward substitution of Rt happened inside the factorisation
#else
  if (a == N - 1) {
#endif
    double zr[M];
#pragma unroll
    for (int i = 0; i < M; ++i) zr[i] = Rm[i];
#ifdef IKB_EAS_SMEM_LDL
#pragma unroll
    for (int i = 1; i < M; ++i)
#pragma unroll
      for (int j = 0; j < i; ++j) zr[i] = fma(-Dm[i * M + j], zr[j], zr[i]);
#endif
    if (EA.updateMode) {
#pragma unroll
      for (int i = 0; i < M; ++i) zr[i] *= invd[i];
#pragma unroll
      for (int i = M - 2; i >= 0; --i)
#pragma unroll
        for (int j = i + 1; j < M; ++j) zr[i] = fma(-Dm[j * M + i], zr[j], zr[i]);
    }
#pragma unroll
    for (int i = 0; i < M; ++i) Rm[i] = zr[i];
  }
  __syncwarp();

  if (EA.updateMode) {
    // alpha -= D^-1 (Rt + L du)     (enhancedassumedstrains.hh:243)
#pragma unroll
    for (int jj = 0; jj < C::ROWS; ++jj) {
      const int j = a + jj * N;
      if (j < M && active) EA.alpha[(size_t)e * M + j] -= Rm[j];
    }
    return;
  }

  // ------------------------------------------------------------------ pass A, then the condensation in registers
  // K_ab = sum_g (...) - Y_a^T diag(1/d) Y_b ; R_a -= Y_a^T diag(1/d) y_R   (enhancedassumedstrains.hh:292-296, 341-345)
  // ------------------------------------------------------------------ pass A: K_uu pairs and R_a
  // (after the solves: the records it reads, [0, O_F) of every Gauss point, are not touched by the scratch below, and
  // the blocks stay in registers until the condensed values are written -- no round trip through the staging array)
  double Ra[D];
#pragma unroll
  for (int i = 0; i < D; ++i) Ra[i] = 0.0;
  double acc[NK][DD];
  {
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
      for (int q = 0; q < DD; ++q) acc[k][q] = 0.0;
#pragma unroll(C::GUNROLL)
    for (int g = 0; g < N; ++g) {
      const double* gp = rec + g * C::GPS;
      const double c1 = gp[C::O_C1], c2 = gp[C::O_C2];
      double ma[D], ga[D], p1[D], p2[D], ha[D], sg[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        ga[i] = gp[C::O_G + i * N + a];
        ma[i] = gp[C::O_M + i * N + a];
        p1[i] = c1 * ma[i];
        p2[i] = c2 * ma[i];
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          Ra[i] = fma(gp[C::O_WP + i * D + j], ga[j], Ra[i]);
          s1 = fma(gp[C::O_X + symIdx<D>(i, j)], ga[j], s1);
          s2 = fma(gp[C::O_WS + symIdx<D>(i, j)], ga[j], s2);
        }
        ha[i] = s1;
        sg[i] = s2;
      }
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int b = (a + k) & (N - 1);
        double mb[D], gb[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          gb[i] = gp[C::O_G + i * N + b];
          mb[i] = gp[C::O_M + i * N + b];
        }
        double cab = 0.0, sab = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          cab = fma(ha[i], gb[i], cab);
          sab = fma(sg[i], gb[i], sab);
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j)
            acc[k][i * D + j] = fma(p1[i], mb[j],
                                    fma(mb[i], p2[j], fma(cab, gp[C::O_A2 + symIdx<D>(i, j)], acc[k][i * D + j])));
#pragma unroll
        for (int i = 0; i < D; ++i) acc[k][i * D + i] += sab;
      }
    }
  }
  __syncwarp();
#pragma unroll 3
  for (int j = 0; j < M; ++j) {
    const double id = invd[j];
    const double* Yj = IKB_LM(j);
    double za[D];
#pragma unroll
    for (int c = 0; c < D; ++c) za[c] = Yj[a * D + c] * id;
    if (A.what & IKB_MATRIX) {
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int b = (a + k) & (N - 1);
        double zb[D];
#pragma unroll
        for (int c = 0; c < D; ++c) zb[c] = Yj[b * D + c];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int c = 0; c < D; ++c) acc[k][i * D + c] = fma(-za[i], zb[c], acc[k][i * D + c]);
      }
    }
    const double zR = Rm[j];
#pragma unroll
    for (int i = 0; i < D; ++i) Ra[i] = fma(-za[i], zR, Ra[i]);
  }
  if (A.what & IKB_MATRIX) {
    // write-out through the (now free) record area: the element's N lanes then store 16-byte vectors to consecutive
    // addresses, full sectors instead of 8-byte pieces 72 bytes apart (the staged K_e of an element is contiguous)
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      if (k == C::KMAX && a >= N / 2) break;
      double* dst = rec + (k * N + a) * DD;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) dst[i * D + j] = (k == 0 && i > j) ? acc[k][j * D + i] : acc[k][i * D + j];
    }
    __syncwarp();
    if (active) {
      constexpr int TOT = C::NPAIR * DD;
      double* Ke = A.Kst + (size_t)e * TOT;
      if constexpr (TOT % 2 == 0) {
        // rec is 16-byte aligned (ES is a multiple of 16 doubles) and so is Ke (TOT even)
        const double2* src2 = reinterpret_cast<const double2*>(rec);
        double2* dst2 = reinterpret_cast<double2*>(Ke);
        for (int q = a; q < TOT / 2; q += N) dst2[q] = src2[q];
      } else {
        for (int q = a; q < TOT; q += N) Ke[q] = rec[q];
      }
    }
  }
  if ((A.what & IKB_VECTOR) && active) {
#pragma unroll
    for (int i = 0; i < D; ++i) A.Rst[(size_t)e * ND + a * D + i] = Ra[i];
  }
#undef IKB_LM
#undef IKB_PM
}

template <int D, int FORM, int M>
cudaError_t launchElemEas(const EasArgs& A, cudaStream_t st) {
  using C = EasCfg<D, FORM, M>;
  // (the opt-in is per device and context: made at every launch, a handle may live on any GPU of the process)
  cudaError_t e =
      cudaFuncSetAttribute(elem_eas_kernel<D, FORM, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((A.E.nElem + C::EPC - 1) / C::EPC);
  if (grid == 0) return cudaSuccess;
  elem_eas_kernel<D, FORM, M><<<grid, C::TPB, C::SMEM, st>>>(A);
  return cudaGetLastError();
}

}  // namespace ikb
