// EAS with an enhanced DISPLACEMENT GRADIENT: EAS::DisplacementGradient and EAS::DisplacementGradientTransposed with the
// H4 (Quad4) / H9 (Hex8) ansatz.
//
// Replaces EnhancedAssumedStrains::calculateMatrixImpl / calculateVectorImpl / updateStateImpl
// (ikarus/finiteelements/mechanics/enhancedassumedstrains.hh:225-248, 258-348, 378-434) over
// strainenhancements/easfunctions/displacementgradient.hh:40-282, displacementgradienttransposed.hh:40-360 and the
// ansatz of strainenhancements/easvariants/displacementgradient.hh:76-163 (helperfunctions.hh:27-36).
//
// The enhanced deformation gradient is  F = I + H_c + sum_p alpha_p Ht_p   (transposed form: I + H_c + F_c0 (sum_p
// alpha_p Ht_p)^T with F_c0 the compatible F at the element centre),  Ht_p = (detJ0/detJ) J0^-T Hhat_p J0^-1  and
// Hhat_(D i + j) = (2 xi_j - 1) e_i (x) e_j, i.e. Ht_p is the rank-one matrix s t_j (J0^-T e_i)(J0^-T e_j)^T.
// Every derivative the reference spells out is an instance of one rule.  For a variation dF_I of F -- e_c (x) g_a for a
// nodal dof (g_a = grad N_a, plus Ht g_a^0 in the transposed form), Ht_p (resp. F_c0 Ht_p^T) for an enhanced one:
//     E,I = sym(F^T dF_I),      E,IJ : S = tr(S dF_I^T dF_J)  (+ P : d_I d_J F, non-zero only for the (alpha, d) pairs of
//     the transposed form: d_p d_(a,c) F = e_c (x) Ht_p g_a^0),
// so K_uu, L, D and R, Rtilde are blocks of ONE generalised tangent over the G = N D + D^2 generalised dofs, with the
// isotropic moduli in the factored form of ikb_elem_q1.cuh:  B_I : CC : B_J = l' (X:B_I)(X:B_J) + 2 m' tr(X B_I X B_J).
//
// The same kernel serves a third mode, ENH_STRAIN: the enhanced dofs contribute dE/dalpha_j = column j of M(xi) (the
// strain enhancements E4..E21, easfunctions/greenlagrangestrain.hh:40-141) and no variation of F; with m = 0 it is the plain
// displacement element.  It is the element kernel of the principal-stretch laws (FORM_PS, ikb_material_ps.cuh), whose
// CC : B_J takes the X B X slot of the record.
//
// One warp per element.  Per Gauss point every lane evaluates the kinematics and the material (redundantly: a few
// hundred flops) and writes the record of its generalised dof -- B, X B X, X:B, dF, dF S -- to shared memory, the nodal
// dofs in one round, the enhanced ones in a second.  The pair loop runs on register tiles: the dofs form NG groups of TS,
// a lane owns the TS x TS block of one group pair and contracts the records channel by channel (2 TS shared-memory loads
// per TS^2 FMAs).  The enhanced block is then eliminated in registers, one lane per column of [D | L | Rtilde]
// (Gauss-Jordan with partial pivoting; the reference inverts D), and the condensed K_e leaves in the symmetric-packed
// staging form of the other element kernels.
#pragma once
#include "ikb_elem_eas.cuh"
#include "ikb_material_ps.cuh"

namespace ikb {

// what the internal variables enhance: the displacement gradient, its transposed form, or (ENH_STRAIN) the
// Green-Lagrange strain with the ansatz tables of ikb_elem_eas.cuh (MM = 0: the plain displacement element)
enum { ENH_DG = 0, ENH_DGT = 1, ENH_STRAIN = 2 };

template <int D, int ENH, int MM = D * D>
struct DgCfg {
  static constexpr bool TR = ENH == ENH_DGT, ST = ENH == ENH_STRAIN;
  static constexpr int N = 1 << D, ND = N * D, M = ST ? MM : D * D, G = ND + M;
  static constexpr int MX = M > 0 ? M : 1;       // array extent that stays legal for M = 0
  static constexpr int SYM = D * (D + 1) / 2;
  // register tiles of the pair loop: the dofs form NG groups of TS, lane t owns the TS x TS block of the group pair
  // (gi <= gj) number t; TS is the smallest size whose NG (NG + 1) / 2 blocks fit the 32 lanes
  static constexpr int tileSize() {
    for (int ts = 1; ts <= 16; ++ts) {
      const int ng = (G + ts - 1) / ts;
      if (ng * (ng + 1) / 2 <= 32) return ts;
    }
    return 16;
  }
  static constexpr int TS = tileSize(), NG = (G + TS - 1) / TS, NT = NG * (NG + 1) / 2;
  // record of a generalised dof at a Gauss point
  static constexpr int O_B = 0, O_XBX = SYM, O_T = 2 * SYM, O_DF = 2 * SYM + 1, O_DFS = O_DF + D * D;
  static constexpr int O_MA = O_DFS + D * D, O_MB = O_MA + D * D;  // transposed form: the mixed second variation
  static constexpr int REC0 = TR ? O_MB + D * D : O_MA;
  static constexpr int REC = REC0 | 1;           // odd stride: lanes on different dofs hit different banks
  static constexpr int GS = G | 1;               // row stride of the generalised tangent
  static constexpr int O_REC = 0, O_KG = O_REC + G * REC, O_RG = O_KG + G * GS, O_X = O_RG + G, O_U = O_X + ND,
                       O_AL = O_U + ND, O_DU = O_AL + M, O_RHS = O_DU + ND, O_T0 = O_RHS + M;
  static constexpr int WARP_DOUBLES = (O_T0 + (ST ? SYM * SYM : 0) + 1) / 2 * 2;
  static constexpr int WARPS = 1;  // one warp per CTA: the register file then holds 9 warps of the 216-register variants instead of 2 CTAs of 4
  static constexpr size_t SMEM = (size_t)WARPS * WARP_DOUBLES * 8;
};

// symmetric index (i,j) -> 0..SYM-1, row-major upper packing
template <int D>
__device__ __forceinline__ constexpr int symI(int i, int j) {
  return symIdx<D>(i, j);
}

template <int D, int FORM, int ENH, int MM>
__global__ void __launch_bounds__(32 * DgCfg<D, ENH, MM>::WARPS) elem_easdg_kernel(EasArgs EA) {
  static_assert(FORM == FORM_SVK || FORM == FORM_NH || FORM == FORM_PS,
                "the displacement gradient enhances the nonlinear element");
  using C = DgCfg<D, ENH, MM>;
  constexpr bool TR = C::TR, ST = C::ST;
  constexpr int N = C::N, ND = C::ND, M = C::M, G = C::G, SYM = C::SYM, REC = C::REC, GS = C::GS;
  const ElemArgs& A = EA.E;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * C::WARPS + warp;
  if (e >= A.nElem) return;
  double* ws = smem + (size_t)warp * C::WARP_DOUBLES;
  double* rec = ws + C::O_REC;
  double* Kg = ws + C::O_KG;
  double* Rg = ws + C::O_RG;
  double* Xs = ws + C::O_X;
  double* us = ws + C::O_U;
  double* al = ws + C::O_AL;
  double* dus = ws + C::O_DU;
  double* rhs = ws + C::O_RHS;

  // ---- element data
  for (int t = lane; t < ND; t += 32) {
    const int a = t / D, c = t - a * D;
    Xs[t] = __ldg(A.X + (size_t)t * A.nElem + e);  // corner a, coordinate c (relative to corner 0)
    const int64_t node = __ldg(A.elemNode + (size_t)a * A.nElem + e);
    const int64_t dof = dofOf(A.layout, D, A.nNodes, node, c);
    us[t] = __ldg(A.U + dof);
    dus[t] = EA.updateMode ? __ldg(EA.dU + dof) : 0.0;
  }
  for (int t = lane; t < M; t += 32) al[t] = EA.alpha[(size_t)e * M + t];
  // strain enhancement: (T(centre) detJ0)^-1 of the element (enhancedassumedstrains.hh:378-390), row-major S x S
  [[maybe_unused]] double* T0 = ws + C::O_T0;
  if constexpr (ST && M > 0)
    for (int t = lane; t < SYM * SYM; t += 32) T0[t] = __ldg(EA.T0inv + (size_t)t * A.nElem + e);
  __syncwarp();

  // the tile of this lane: group pair (gi, gj), gi <= gj, number lane (row gi holds NG - gi tiles); lanes past NT idle
  constexpr int TS = C::TS, NG = C::NG;
  int gi = 0, gj = 0;
  {
    int t = lane < C::NT ? lane : 0, row = NG;
    while (t >= row) {
      t -= row;
      --row;
      ++gi;
    }
    gj = gi + t;
  }
  const bool tileActive = lane < C::NT;
  // first dof of the two groups; slots past G read the last dof (their products are never stored)
  const int I0 = gi * TS, J0 = gj * TS;
  double acc[TS][TS];
#pragma unroll
  for (int i = 0; i < TS; ++i)
#pragma unroll
    for (int j = 0; j < TS; ++j) acc[i][j] = 0.0;
  double rI[2] = {0.0, 0.0};  // R_gen of the nodal dof lane and of the enhanced dof ND + lane
  [[maybe_unused]] double energy = 0.0;

  // ---- element centre: J0^-T, detJ0, (transposed form) grad N^0 and F_c0
  double JI0[D][D], detJ0;
  double Fc0[D][D];
  {
    double Jt[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) Jt[i][k] = 0.0;
    const double half = (D == 3) ? 0.25 : 0.5;  // dN_c/dxi_i at the centre: +-0.5^(D-1)
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double dn = ((c >> i) & 1) ? half : -half;
#pragma unroll
        for (int k = 0; k < D; ++k) Jt[i][k] = fma(dn, Xs[c * D + k], Jt[i][k]);
      }
    detJ0 = fabs(invSmall<D>(Jt, JI0));
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int j = 0; j < D; ++j) Fc0[c][j] = (c == j) ? 1.0 : 0.0;
    if constexpr (TR) {
#pragma unroll
      for (int a = 0; a < N; ++a) {
        double g0[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < D; ++i) s = fma(JI0[j][i], ((a >> i) & 1) ? half : -half, s);
          g0[j] = s;
        }
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
          for (int j = 0; j < D; ++j) Fc0[c][j] = fma(us[a * D + c], g0[j], Fc0[c][j]);
      }
    }
  }

  // ---- Gauss points
#pragma unroll 1
  for (int g = 0; g < N; ++g) {
    const double lo = 0.5 - 0.28867513459481287, hi = 0.5 + 0.28867513459481287;
    double xi[D], om[D], tt[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      xi[k] = ((g >> k) & 1) ? hi : lo;
      om[k] = 1.0 - xi[k];
      tt[k] = 2.0 * xi[k] - 1.0;
    }
    double Jt[D][D], Hx[D][D];  // Hx[c][i] = sum_a u_a[c] dN_a/dxi_i
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) Jt[i][k] = Hx[i][k] = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double dn = ((c >> i) & 1) ? 1.0 : -1.0;
#pragma unroll
        for (int k = 0; k < D; ++k)
          if (k != i) dn *= ((c >> k) & 1) ? xi[k] : om[k];
#pragma unroll
        for (int k = 0; k < D; ++k) {
          Jt[i][k] = fma(dn, Xs[c * D + k], Jt[i][k]);
          Hx[k][i] = fma(dn, us[c * D + k], Hx[k][i]);
        }
      }
    double Ji[D][D];
    const double detJ = fabs(invSmall<D>(Jt, Ji));
    double wd = detJ;
#pragma unroll
    for (int k = 0; k < D; ++k) wd *= 0.5;
    const double sc = detJ0 / detJ;
    // compatible gradient H_c[c][j] = sum_i Hx[c][i] Ji[j][i]; enhanced part Ht = sc J0^-T (alpha_(i,j) t_j) J0^-1
    double H[D][D], Hs[D][D];
    if constexpr (ST) {
#pragma unroll
      for (int c = 0; c < D; ++c)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < D; ++i) s = fma(Hx[c][i], Ji[j][i], s);
          H[c][j] = s;
          Hs[c][j] = 0.0;
        }
    } else {
      double T1[D][D];  // (alpha_(i,j) t_j) J0^-1 :  T1[i][l] = sum_j alpha_(Di+j) t_j JI0[l][j]
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int l = 0; l < D; ++l) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) s = fma(al[D * i + j] * tt[j], JI0[l][j], s);
          T1[i][l] = s;
        }
#pragma unroll
      for (int k = 0; k < D; ++k)
#pragma unroll
        for (int l = 0; l < D; ++l) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < D; ++i) s = fma(JI0[k][i], T1[i][l], s);
          Hs[k][l] = sc * s;
        }
#pragma unroll
      for (int c = 0; c < D; ++c)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < D; ++i) s = fma(Hx[c][i], Ji[j][i], s);
          if constexpr (TR) {
            // H = H_c + F_c0 Hs^T
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(Fc0[c][k], Hs[j][k], s);
          } else {
            s += Hs[c][j];
          }
          H[c][j] = s;
        }
    }
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
    // strain enhancement: E = E_c + T0inv (sum_j s_j alpha_j e_(r_j)), s_j = p_j(xi) / detJ  (easfunctions/
    // greenlagrangestrain.hh:40-141 with the ansatz of easvariants/linearandglstrains.hh), i.e. C += 2 (E - E_c)
    [[maybe_unused]] double sm[D == 3 ? 6 : 3];
    if constexpr (ST && M > 0) {
      using T = EasTable<D, M>;
      monomials<D>(g, 1.0 / detJ, sm);
      double v[SYM];
#pragma unroll
      for (int q = 0; q < SYM; ++q) v[q] = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) v[T::row(j)] = fma(sm[T::mono(j)], al[j], v[T::row(j)]);
#pragma unroll
      for (int p = 0; p < SYM; ++p) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < SYM; ++q) s = fma(T0[p * SYM + q], v[q], s);
        int i, j;
        voigtPair<D>(p, i, j);
        if (i == j) {
          Cm[i][i] = fma(2.0, s, Cm[i][i]);  // Voigt strain: shear entries are doubled
        } else {
          Cm[i][j] += s;
          Cm[j][i] += s;
        }
      }
    }
    // material: S, X, l', m'  (svk.hh; neohooke.hh:79-142 with C = 2E + I; plane strain: the 3D law at zero
    // out-of-plane strain, vanishingstrain.hh)
    double Sm[D][D], Xm[D][D], lp, mp;
    [[maybe_unused]] double Np[3][3], L1[3][3], L2[3][3];  // principal frame and moduli of the principal-stretch laws
    if constexpr (FORM == FORM_PS) {
      double Sp[3], psi;
      if (!principalLaw<D>(EA.ps, Cm, Np, Sp, L1, L2, psi) && lane == 0)
        atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
      principalStress<D>(Np, Sp, Sm);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) Xm[i][j] = (i == j) ? 1.0 : 0.0;
      lp = 0.0;   // the pair loop below then adds wd B_I : (CC : B_J), with CC : B_J in the X B X slot of the record
      mp = 0.5;
      energy = fma(wd, psi, energy);
    } else if constexpr (FORM == FORM_SVK) {
      double tr = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) tr += 0.5 * (Cm[i][i] - 1.0);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const double E = 0.5 * (Cm[i][j] - (i == j ? 1.0 : 0.0));
          Sm[i][j] = 2.0 * A.mu * E + (i == j ? A.lambda * tr : 0.0);
          Xm[i][j] = (i == j) ? 1.0 : 0.0;
        }
      lp = A.lambda;
      mp = A.mu;
    } else {
      const double detC = invSmall<D>(Cm, Xm);
      if (!(detC > 0.0) && lane == 0) atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
      const double lnJ = 0.5 * log(detC);
      lp = A.lambda;
      mp = A.mu - A.lambda * lnJ;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) Sm[i][j] = (i == j ? A.mu : 0.0) - mp * Xm[i][j];
    }

    // ---- record(s) of this lane's generalised dof(s)
    __syncwarp();  // the pair loop of the previous Gauss point has read the records
#pragma unroll
    // round 0: the nodal dofs (lane < ND), round 1: the enhanced ones (lane < M) -- every round runs one of the two
    // record formulas on all its lanes instead of both on a mixed set
    for (int rnd = 0; rnd < (M > 0 ? 2 : 1); ++rnd) {
      const int I = rnd == 0 ? lane : ND + lane;
      if (rnd == 0 ? lane < ND : lane < M) {
        double dF[D][D], mA[D][D], mB[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) dF[i][j] = mA[i][j] = mB[i][j] = 0.0;
        if (I < ND) {
          const int a = I / D, c = I - a * D;
          double dn[D], gph[D];
#pragma unroll
          for (int i = 0; i < D; ++i) {
            double v = ((a >> i) & 1) ? 1.0 : -1.0;
#pragma unroll
            for (int k = 0; k < D; ++k)
              if (k != i) v *= ((a >> k) & 1) ? xi[k] : om[k];
            dn[i] = v;
          }
#pragma unroll
          for (int j = 0; j < D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) s = fma(Ji[j][i], dn[i], s);
            gph[j] = s;
          }
          if constexpr (TR) {
            const double half = (D == 3) ? 0.25 : 0.5;
            double g0[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
              double s = 0.0;
#pragma unroll
              for (int i = 0; i < D; ++i) s = fma(JI0[j][i], ((a >> i) & 1) ? half : -half, s);
              g0[j] = s;
            }
            // g_a + Hs g_a^0 (dNtilde, displacementgradienttransposed.hh:133-141); mA = e_c (x) g_a^0
#pragma unroll
            for (int j = 0; j < D; ++j) {
#pragma unroll
              for (int k = 0; k < D; ++k) gph[j] = fma(Hs[j][k], g0[k], gph[j]);
            }
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
              for (int j = 0; j < D; ++j) mA[r][j] = (r == c) ? g0[j] : 0.0;
          }
#pragma unroll
          for (int r = 0; r < D; ++r)
#pragma unroll
            for (int j = 0; j < D; ++j) dF[r][j] = (r == c) ? gph[j] : 0.0;
        } else if constexpr (!ST) {
          const int p = I - ND, i0 = p / D, j0 = p - i0 * D;
          double Ht[D][D];
          const double f = sc * (j0 == 0 ? tt[0] : (j0 == 1 ? tt[1] : tt[D - 1]));
#pragma unroll
          for (int k = 0; k < D; ++k)
#pragma unroll
            for (int l = 0; l < D; ++l) {
              double ki = 0.0, lj = 0.0;
#pragma unroll
              for (int q = 0; q < D; ++q) {
                ki = (q == i0) ? JI0[k][q] : ki;
                lj = (q == j0) ? JI0[l][q] : lj;
              }
              Ht[k][l] = f * ki * lj;
            }
          if constexpr (TR) {
            // dF = F_c0 Ht^T;  mB = P Ht with P = F S (the work of d_p d_(a,c) F = e_c (x) Ht g_a^0)
            double P[D][D];
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
              for (int j = 0; j < D; ++j) {
                double s = 0.0, pp = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                  s = fma(Fc0[r][k], Ht[j][k], s);
                  pp = fma(F[r][k], Sm[k][j], pp);
                }
                dF[r][j] = s;
                P[r][j] = pp;
              }
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
              for (int k = 0; k < D; ++k) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) s = fma(P[r][j], Ht[j][k], s);
                mB[r][k] = s;
              }
          } else {
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
              for (int j = 0; j < D; ++j) dF[r][j] = Ht[r][j];
          }
        }
        // B = sym(F^T dF), X B X, X:B, dF S
        double FtdF[D][D], Bm[D][D];
        [[maybe_unused]] double XB[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(F[k][i], dF[k][j], s);
            FtdF[i][j] = s;
          }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) Bm[i][j] = 0.5 * (FtdF[i][j] + FtdF[j][i]);
        if constexpr (ST && M > 0) {
          if (I >= ND) {
            // enhanced strain dof: dE/dalpha_j = column j of M(xi) = T0inv[:, r_j] s_j, no variation of F
            using T = EasTable<D, M>;
            const int jj = I - ND;
            const int rj = T::rowRt(jj);
            const double sj = pickMono(sm, T::monoRt(jj));
#pragma unroll
            for (int p = 0; p < SYM; ++p) {
              int i, j;
              voigtPair<D>(p, i, j);
              const double v = T0[p * SYM + rj] * sj;
              Bm[i][j] = Bm[j][i] = (i == j) ? v : 0.5 * v;
            }
          }
        }
        double* r = rec + I * REC;
        double tI = 0.0, rS = 0.0;
        if constexpr (FORM == FORM_PS) {
          double CB[D][D];
          principalTangentTimes<D>(Np, L1, L2, Bm, CB);
#pragma unroll
          for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j) {
              r[C::O_B + symI<D>(i, j)] = Bm[i][j];
              r[C::O_XBX + symI<D>(i, j)] = CB[i][j];
            }
        } else {
#pragma unroll
          for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
              double s = 0.0;
#pragma unroll
              for (int k = 0; k < D; ++k) s = fma(Xm[i][k], Bm[k][j], s);
              XB[i][j] = s;
            }
#pragma unroll
          for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j) {
              double s = 0.0;
#pragma unroll
              for (int k = 0; k < D; ++k) s = fma(XB[i][k], Xm[k][j], s);
              r[C::O_B + symI<D>(i, j)] = Bm[i][j];
              r[C::O_XBX + symI<D>(i, j)] = s;
            }
#pragma unroll
          for (int i = 0; i < D; ++i) tI += XB[i][i];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
          for (int j = 0; j < D; ++j) rS = fma(Bm[i][j], Sm[i][j], rS);
        }
        r[C::O_T] = tI;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(dF[i][k], Sm[k][j], s);
            r[C::O_DF + i * D + j] = dF[i][j];
            r[C::O_DFS + i * D + j] = s;
            if constexpr (TR) {
              r[C::O_MA + i * D + j] = mA[i][j];
              r[C::O_MB + i * D + j] = mB[i][j];
            }
          }
        rI[rnd] = fma(wd, rS, rI[rnd]);
      }
    }
    __syncwarp();

    // ---- pairs, channel by channel: for every scalar channel q of the records (X:B, the 6 (3) strain components, the D^2
    // entries of dF against dF S, and the mixed terms of the transposed form) the tile takes the outer product of the TS
    // row values with the TS column values -- 2 TS shared-memory loads for TS^2 FMAs instead of two whole records per pair
    const double c1 = wd * lp, c2 = 2.0 * wd * mp;
    if (tileActive) {
      const double* ra[TS];
      const double* rb[TS];
#pragma unroll
      for (int i = 0; i < TS; ++i) {
        const int I = I0 + i < G ? I0 + i : G - 1, J = J0 + i < G ? J0 + i : G - 1;
        ra[i] = rec + I * REC;
        rb[i] = rec + J * REC;
      }
      auto channel = [&](int oa, int ob, double coef) {
        double x[TS], y[TS];
#pragma unroll
        for (int i = 0; i < TS; ++i) {
          x[i] = coef * ra[i][oa];
          y[i] = rb[i][ob];
        }
#pragma unroll
        for (int i = 0; i < TS; ++i)
#pragma unroll
          for (int j = 0; j < TS; ++j) acc[i][j] = fma(x[i], y[j], acc[i][j]);
      };
      if constexpr (FORM != FORM_PS) channel(C::O_T, C::O_T, c1);  // lp = 0 for the principal-stretch laws
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = i; j < D; ++j) channel(C::O_B + symI<D>(i, j), C::O_XBX + symI<D>(i, j), (i == j) ? c2 : 2.0 * c2);
#pragma unroll
      for (int q = 0; q < D * D; ++q) channel(C::O_DF + q, C::O_DFS + q, wd);
      if constexpr (TR) {
#pragma unroll
        // mixed second variation: mA is carried by the nodal dofs only, mB by the enhanced ones only, and the tiles hold
        // I <= J (nodal before enhanced), so the term mB_I mA_J of the symmetrised product never contributes
        for (int q = 0; q < D * D; ++q) channel(C::O_MA + q, C::O_MB + q, wd);
      }
    }
  }

  // ---- generalised tangent and residual to shared memory
  __syncwarp();
  if constexpr (FORM == FORM_PS && M == 0) {
    if ((A.what & IKB_SCALAR) && lane == 0) A.Est[e] = energy;  // int psi dV (nonlinearelastic.hh:276-290)
  }
  if (tileActive) {
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
      for (int j = 0; j < TS; ++j) {
        const int I = I0 + i, J = J0 + j;
        if (I < G && J < G && I <= J) {  // diagonal tiles keep their upper triangle
          Kg[I * GS + J] = acc[i][j];
          Kg[J * GS + I] = acc[i][j];
        }
      }
  }
  if (lane < ND) Rg[lane] = rI[0];
  if (lane < M) Rg[ND + lane] = rI[1];
  __syncwarp();
  // right-hand side of the enhanced block: Rtilde, in update mode Rtilde + L du (enhancedassumedstrains.hh:243)
  if (lane < M) {
    double s = Rg[ND + lane];
    if (EA.updateMode)
      for (int j = 0; j < ND; ++j) s = fma(Kg[(ND + lane) * GS + j], dus[j], s);
    rhs[lane] = s;
  }
  __syncwarp();

  // ---- [D | L | rhs] -> [I | D^-1 L | D^-1 rhs]: Gauss-Jordan with partial pivoting (the reference inverts D).  The
  // columns ND.. of the rows below ND keep L^T for the condensation.
#ifndef IKB_DG_SMEM_GJ
  // In registers, one lane per COLUMN of the augmented matrix (G + 1 columns on 32 lanes: NCOL per lane), so a row
  // operation is lane-local.  Per pivot the owner of column ND + k picks the pivot row and hands it, with the column's
  // entries (the multipliers), to all lanes by shuffles; the row exchange is a select chain (no dynamic register index).
  // No shared-memory access and no barrier inside the M dependent steps.
  if constexpr (M > 0) {
    constexpr int NCOL = (G + 1 + 31) / 32;
    double col[NCOL][C::MX];
#pragma unroll
    for (int sl = 0; sl < NCOL; ++sl) {
      const int c = lane + 32 * sl;
#pragma unroll
      for (int r = 0; r < M; ++r) col[sl][r] = c < G ? Kg[(ND + r) * GS + c] : (c == G ? rhs[r] : 0.0);
    }
#pragma unroll
    for (int k = 0; k < M; ++k) {
      constexpr unsigned FULLM = 0xffffffffu;
      const int ol = (ND + k) % 32, os = (ND + k) / 32;  // owner lane and slot of column ND + k
      // pivot row: first maximum of |column| over r >= k, decided by the owner
      int piv = k;
      double best = fabs(col[os][k]);
#pragma unroll
      for (int r = k + 1; r < M; ++r) {
        const double v = fabs(col[os][r]);
        const bool gt = v > best;
        best = gt ? v : best;
        piv = gt ? r : piv;
      }
      piv = __shfl_sync(FULLM, piv, ol);
      double f[C::MX];
#pragma unroll
      for (int r = 0; r < M; ++r) f[r] = __shfl_sync(FULLM, col[os][r], ol);
      // exchange rows k and piv in the multipliers and in this lane's columns
      {
        double fp = f[k];
#pragma unroll
        for (int r = k + 1; r < M; ++r) fp = (r == piv) ? f[r] : fp;
        const double fk = f[k];
        f[k] = fp;
#pragma unroll
        for (int r = k + 1; r < M; ++r) f[r] = (r == piv) ? fk : f[r];
      }
      const double ip = 1.0 / f[k];
#pragma unroll
      for (int sl = 0; sl < NCOL; ++sl) {
        double cp = col[sl][k];
#pragma unroll
        for (int r = k + 1; r < M; ++r) cp = (r == piv) ? col[sl][r] : cp;
        const double ck = col[sl][k];
#pragma unroll
        for (int r = k + 1; r < M; ++r) col[sl][r] = (r == piv) ? ck : col[sl][r];
        const double pk = cp * ip;
        col[sl][k] = pk;
#pragma unroll
        for (int r = 0; r < M; ++r)
          if (r != k) col[sl][r] = fma(-f[r], pk, col[sl][r]);
      }
    }
    // D^-1 L and D^-1 rhs back to shared memory (the columns ND.. have become the identity and are not needed)
#pragma unroll
    for (int sl = 0; sl < NCOL; ++sl) {
      const int c = lane + 32 * sl;
#pragma unroll
      for (int r = 0; r < M; ++r) {
        if (c < ND) Kg[(ND + r) * GS + c] = col[sl][r];
        if (c == G) rhs[r] = col[sl][r];
      }
    }
    __syncwarp();
  }
#else
  for (int k = 0; k < M; ++k) {
    int piv = k;
    double best = fabs(Kg[(ND + k) * GS + ND + k]);
    for (int r = k + 1; r < M; ++r) {
      const double v = fabs(Kg[(ND + r) * GS + ND + k]);
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    __syncwarp();
    if (piv != k) {
      for (int c = lane; c <= G; c += 32) {
        double* x = (c < G) ? &Kg[(ND + k) * GS + c] : &rhs[k];
        double* y = (c < G) ? &Kg[(ND + piv) * GS + c] : &rhs[piv];
        const double t = *x;
        *x = *y;
        *y = t;
      }
      __syncwarp();
    }
    const double ip = 1.0 / Kg[(ND + k) * GS + ND + k];
    double f[C::MX];
#pragma unroll
    for (int r = 0; r < M; ++r) f[r] = Kg[(ND + r) * GS + ND + k];
    __syncwarp();
    for (int c = lane; c <= G; c += 32) {
      double* rowk = (c < G) ? &Kg[(ND + k) * GS + c] : &rhs[k];
      const double pk = *rowk * ip;
      *rowk = pk;
#pragma unroll
      for (int r = 0; r < M; ++r)
        if (r != k) {
          double* x = (c < G) ? &Kg[(ND + r) * GS + c] : &rhs[r];
          *x = fma(-f[r], pk, *x);
        }
    }
    __syncwarp();
  }

#endif

  if (EA.updateMode) {
    if (lane < M) EA.alpha[(size_t)e * M + lane] -= rhs[lane];
    return;
  }

  // ---- condensation: K_uu -= L^T (D^-1 L) on the upper triangle, R_u -= L^T (D^-1 Rtilde)
  // (enhancedassumedstrains.hh:292-296, 341-345)
  for (int t = lane; t < ND * (ND + 1) / 2; t += 32) {
    int I = 0, off = 0;
    while (t >= off + (ND - I)) {
      off += ND - I;
      ++I;
    }
    const int J = I + (t - off);
    double s = Kg[I * GS + J];
#pragma unroll
    for (int p = 0; p < M; ++p) s = fma(-Kg[I * GS + ND + p], Kg[(ND + p) * GS + J], s);
    Kg[I * GS + J] = s;
  }
  if (lane < ND && (A.what & IKB_VECTOR)) {
    double s = Rg[lane];
#pragma unroll
    for (int p = 0; p < M; ++p) s = fma(-Kg[lane * GS + ND + p], rhs[p], s);
    A.Rst[(size_t)e * ND + lane] = s;
  }
  __syncwarp();
  if (A.what & IKB_MATRIX) {
    // symmetric-packed staging: pair p = k*N + a holds block (a, (a+k) mod N), k <= N/2 (k == N/2: a < N/2)
    constexpr int NPAIR = N * (N + 1) / 2, DD = D * D;
    double* Ke = A.Kst + (size_t)e * NPAIR * DD;
    for (int t = lane; t < NPAIR * DD; t += 32) {
      const int p = t / DD, q = t - p * DD, i = q / D, j = q - i * D;
      const int k = p / N, a = p - k * N, b = (a + k) & (N - 1);
      const int I = a * D + i, J = b * D + j;
      Ke[t] = (I <= J) ? Kg[I * GS + J] : Kg[J * GS + I];
    }
  }
}

template <int D, int FORM, int ENH, int MM = D * D>
cudaError_t launchElemEasDg(const EasArgs& A, cudaStream_t st) {
  using C = DgCfg<D, ENH, MM>;
  cudaError_t e = cudaFuncSetAttribute(elem_easdg_kernel<D, FORM, ENH, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::SMEM);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((A.E.nElem + C::WARPS - 1) / C::WARPS);
  if (grid == 0) return cudaSuccess;
  elem_easdg_kernel<D, FORM, ENH, MM><<<grid, 32 * C::WARPS, C::SMEM, st>>>(A);
  return cudaGetLastError();
}

}  // namespace ikb
