// Result evaluation at local positions: fe.calculateAt<ResultTypes::...>(req, local) for every element.
//
// Replaces NonLinearElastic::calculateAtImpl (ikarus/finiteelements/mechanics/nonlinearelastic.hh:237-271),
// LinearElastic::calculateAtImpl (mechanics/linearelastic.hh:~200-240) and
// EnhancedAssumedStrains::calculateAtImpl (mechanics/enhancedassumedstrains.hh:127-187): displacement gradient at xi,
// strain (Green-Lagrange or linear, plus M(xi) alpha with EAS), stress of the material in Voigt notation, optionally
// pushed forward to Kirchhoff / Cauchy (transformStress).  One thread per (element, position); this is post-processing
// (VTK output, stress checks), bandwidth-light and far off the assembly hot path.
#pragma once
#include "ikb_elem_eas.cuh"
#include "ikb_material_ps.cuh"
#include "ikb_internal.cuh"

namespace ikb {

struct ResultArgs {
  const double* X;          // [nc*D][nElem] corner coordinates
  const int32_t* elemNode;  // [N][nElem]
  const double* U;          // [nDof]
  const double* T0inv;      // EAS: [S*S][nElem]
  const double* alpha;      // EAS: [nElem][M]
  const double* local;      // [npts][D] positions in the reference element
  double* out;              // [nElem][npts][ncomp]
  int32_t* errFlag;
  int64_t nElem, nNodes;
  int layout, npts, form, planeStrain, easM, resultType, ncomp;
  int easFunction = 0;  // IKB_EAS_*
  double lambda, mu, psTol;
  PsLaw ps;  // FORM_PS
};

template <int D>
__device__ __forceinline__ void easColumn(int m, int j, int& row, int& mono) {
  if constexpr (D == 2) {
    if (m == 4) {
      row = EasTable<2, 4>::row(j), mono = EasTable<2, 4>::mono(j);
    } else if (m == 5) {
      row = EasTable<2, 5>::row(j), mono = EasTable<2, 5>::mono(j);
    } else {
      row = EasTable<2, 7>::row(j), mono = EasTable<2, 7>::mono(j);
    }
  } else {
    if (m == 9) {
      row = EasTable<3, 9>::row(j), mono = EasTable<3, 9>::mono(j);
    } else {
      row = EasTable<3, 21>::row(j), mono = EasTable<3, 21>::mono(j);
    }
  }
}

// 3D law on the full Voigt strain E6 (shear entries doubled): S6.  Returns false for det C <= 0 (NeoHooke).
__device__ __forceinline__ bool stress3d(int form, double lambda, double mu, const PsLaw& ps, const double (&E6)[6],
                                         double (&S6)[6]) {
  if (form == FORM_PS) {  // principal-stretch laws (hyperelastic/interface.hh:124-141), ikb_material_ps.cuh
    double Cm[3][3], Np[3][3], Sp[3], L1[3][3], L2[3][3], psi, Sm[3][3];
    Cm[0][0] = 2.0 * E6[0] + 1.0, Cm[1][1] = 2.0 * E6[1] + 1.0, Cm[2][2] = 2.0 * E6[2] + 1.0;
    Cm[1][2] = Cm[2][1] = E6[3], Cm[0][2] = Cm[2][0] = E6[4], Cm[0][1] = Cm[1][0] = E6[5];
    if (!principalLaw<3>(ps, Cm, Np, Sp, L1, L2, psi)) return false;
    principalStress<3>(Np, Sp, Sm);
    S6[0] = Sm[0][0], S6[1] = Sm[1][1], S6[2] = Sm[2][2], S6[3] = Sm[1][2], S6[4] = Sm[0][2], S6[5] = Sm[0][1];
    return true;
  }
  if (form != FORM_NH) {  // svk.hh:77-164, linearelasticity.hh:33-136
    const double tr = E6[0] + E6[1] + E6[2];
    for (int i = 0; i < 3; ++i) S6[i] = lambda * tr + 2.0 * mu * E6[i];
    for (int i = 3; i < 6; ++i) S6[i] = mu * E6[i];
    return true;
  }
  // neohooke.hh:79-142 with C = 2E + I (strainconversions.hh:88-106); Voigt [00,11,22,12,02,01]
  const double c00 = 2.0 * E6[0] + 1.0, c11 = 2.0 * E6[1] + 1.0, c22 = 2.0 * E6[2] + 1.0;
  const double c12 = E6[3], c02 = E6[4], c01 = E6[5];
  const double a00 = c11 * c22 - c12 * c12, a01 = c02 * c12 - c01 * c22, a02 = c01 * c12 - c02 * c11;
  const double detC = c00 * a00 + c01 * a01 + c02 * a02;
  if (!(detC > 0.0)) return false;  // checkPositiveOrAbort (relativeWeak compare against 0 == detC <= 0)
  const double id = 1.0 / detC;
  const double i00 = a00 * id, i01 = a01 * id, i02 = a02 * id;
  const double i11 = (c00 * c22 - c02 * c02) * id, i12 = (c01 * c02 - c00 * c12) * id, i22 = (c00 * c11 - c01 * c01) * id;
  const double ll = lambda * 0.5 * log(detC);  // lambda ln J
  S6[0] = mu * (1.0 - i00) + ll * i00;
  S6[1] = mu * (1.0 - i11) + ll * i11;
  S6[2] = mu * (1.0 - i22) + ll * i22;
  S6[3] = (ll - mu) * i12;
  S6[4] = (ll - mu) * i02;
  S6[5] = (ll - mu) * i01;
  return true;
}

template <int D, int ORDER>
__global__ void __launch_bounds__(128) result_at_kernel(ResultArgs A) {
  constexpr int P = ORDER + 1;
  constexpr int N = D == 3 ? P * P * P : P * P;
  constexpr int NC = 1 << D;
  constexpr int S = D * (D + 1) / 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nElem * A.npts) return;
  const int64_t e = t / A.npts;
  const int q = (int)(t - e * A.npts);
  double xi[D];
  for (int k = 0; k < D; ++k) xi[k] = A.local[q * D + k];

  // geometry: multilinear in the 2^D corners (also for Q2 elements)
  double Jt[D][D];
  for (int i = 0; i < D; ++i)
    for (int k = 0; k < D; ++k) Jt[i][k] = 0.0;
  for (int c = 0; c < NC; ++c)
    for (int i = 0; i < D; ++i) {
      double dn = ((c >> i) & 1) ? 1.0 : -1.0;
      for (int k = 0; k < D; ++k)
        if (k != i) dn *= ((c >> k) & 1) ? xi[k] : 1.0 - xi[k];
      for (int k = 0; k < D; ++k) Jt[i][k] += dn * A.X[(size_t)(c * D + k) * A.nElem + e];
    }
  double Ji[D][D];
  const double detJ = invSmall<D>(Jt, Ji);  // Ji = Jt^-1

  // 1D Lagrange polynomials of the basis on [0,1] and their derivatives
  double sh[D][P], ds[D][P];
  for (int k = 0; k < D; ++k) {
    const double x = xi[k];
    if constexpr (ORDER == 1) {
      sh[k][0] = 1.0 - x, sh[k][1] = x;
      ds[k][0] = -1.0, ds[k][1] = 1.0;
    } else {
      sh[k][0] = (2.0 * x - 1.0) * (x - 1.0), sh[k][1] = 4.0 * x * (1.0 - x), sh[k][2] = x * (2.0 * x - 1.0);
      ds[k][0] = 4.0 * x - 3.0, ds[k][1] = 4.0 - 8.0 * x, ds[k][2] = 4.0 * x - 1.0;
    }
  }
  // H = sum_a u_a (x) grad N_a,  grad N_a = Jt^-1 dN_a/dxi
  double H[D][D];
  for (int c = 0; c < D; ++c)
    for (int j = 0; j < D; ++j) H[c][j] = 0.0;
  for (int a = 0; a < N; ++a) {
    int ia[D];
    ia[0] = a % P;
    ia[1] = (a / P) % P;
    if constexpr (D == 3) ia[2] = a / (P * P);
    double dref[D];
    for (int i = 0; i < D; ++i) {
      double v = ds[i][ia[i]];
      for (int k = 0; k < D; ++k)
        if (k != i) v *= sh[k][ia[k]];
      dref[i] = v;
    }
    double g[D];
    for (int j = 0; j < D; ++j) {
      double v = 0.0;
      for (int i = 0; i < D; ++i) v += Ji[j][i] * dref[i];
      g[j] = v;
    }
    const int64_t node = A.elemNode[(size_t)a * A.nElem + e];
    for (int c = 0; c < D; ++c) {
      const double u = A.U[dofOf(A.layout, D, A.nNodes, node, c)];
      for (int j = 0; j < D; ++j) H[c][j] += u * g[j];
    }
  }
  // strain in Voigt notation (shear doubled)
  if (A.easM && A.easFunction != IKB_EAS_STRAIN) {
    // EnhancedStrainFunction::computeDisplacementGradient (easfunctions/displacementgradient.hh:40-64,
    // displacementgradienttransposed.hh:40-58, 343-360): H = H_c + Ht, resp. H_c + F_c0 Ht^T with
    // Ht = (detJ0/detJ) J0^-T (alpha_(i,j) (2 xi_j - 1)) J0^-1 (easvariants/displacementgradient.hh, helperfunctions.hh:27-36)
    if constexpr (ORDER == 1) {
      double Jt0[D][D], JI0[D][D], Fc0[D][D];
      const double half = (D == 3) ? 0.25 : 0.5;
      for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) Jt0[i][k] = 0.0, Fc0[i][k] = (i == k) ? 1.0 : 0.0;
      for (int c = 0; c < NC; ++c)
        for (int i = 0; i < D; ++i)
          for (int k = 0; k < D; ++k) Jt0[i][k] += (((c >> i) & 1) ? half : -half) * A.X[(size_t)(c * D + k) * A.nElem + e];
      const double detJ0 = fabs(invSmall<D>(Jt0, JI0));
      if (A.easFunction == IKB_EAS_DISPLACEMENT_GRADIENT_TRANSPOSED) {
        for (int a = 0; a < NC; ++a) {
          const int64_t node = A.elemNode[(size_t)a * A.nElem + e];
          double g0[D];
          for (int j = 0; j < D; ++j) {
            double v = 0.0;
            for (int i = 0; i < D; ++i) v += JI0[j][i] * (((a >> i) & 1) ? half : -half);
            g0[j] = v;
          }
          for (int c = 0; c < D; ++c) {
            const double u = A.U[dofOf(A.layout, D, A.nNodes, node, c)];
            for (int j = 0; j < D; ++j) Fc0[c][j] += u * g0[j];
          }
        }
      }
      const double sc = detJ0 / fabs(detJ);
      double T1[D][D], Hs[D][D];
      for (int i = 0; i < D; ++i)
        for (int l = 0; l < D; ++l) {
          double v = 0.0;
          for (int j = 0; j < D; ++j) v += A.alpha[(size_t)e * A.easM + D * i + j] * (2.0 * xi[j] - 1.0) * JI0[l][j];
          T1[i][l] = v;
        }
      for (int k = 0; k < D; ++k)
        for (int l = 0; l < D; ++l) {
          double v = 0.0;
          for (int i = 0; i < D; ++i) v += JI0[k][i] * T1[i][l];
          Hs[k][l] = sc * v;
        }
      for (int c = 0; c < D; ++c)
        for (int j = 0; j < D; ++j) {
          if (A.easFunction == IKB_EAS_DISPLACEMENT_GRADIENT_TRANSPOSED) {
            for (int k = 0; k < D; ++k) H[c][j] += Fc0[c][k] * Hs[j][k];
          } else {
            H[c][j] += Hs[c][j];
          }
        }
    }
  }
  const bool gl = A.form != FORM_LE;
  double Ev[S];
  for (int p = 0; p < S; ++p) {
    int i, j;
    voigtPair<D>(p, i, j);
    double v = H[i][j] + H[j][i];
    if (gl)
      for (int k = 0; k < D; ++k) v += H[k][i] * H[k][j];
    Ev[p] = i == j ? 0.5 * v : v;
  }
  if (A.easM && A.easFunction == IKB_EAS_STRAIN) {  // E += M(xi) alpha,  M[:, j] = T0inv[:, r_j] p_j(2 xi - 1) / detJ(xi)
    double tt[D];
    for (int k = 0; k < D; ++k) tt[k] = 2.0 * xi[k] - 1.0;
    const double idet = 1.0 / fabs(detJ);
    for (int j = 0; j < A.easM; ++j) {
      int row, mono;
      easColumn<D>(A.easM, j, row, mono);
      double pj;
      if constexpr (D == 3)
        pj = mono < 3 ? tt[mono] : (mono == 3 ? tt[0] * tt[1] : (mono == 4 ? tt[0] * tt[2] : tt[1] * tt[2]));
      else
        pj = mono < 2 ? tt[mono] : tt[0] * tt[1];
      const double f = pj * idet * A.alpha[(size_t)e * A.easM + j];
      for (int p = 0; p < S; ++p) Ev[p] += A.T0inv[(size_t)(p * S + row) * A.nElem + e] * f;
    }
  }
  // embed into the 3D law (plane strain: free Voigt indices {0,1,5}, vanishingstrain.hh:79-120)
  double E6[6] = {0, 0, 0, 0, 0, 0}, S6[6];
  if constexpr (D == 3) {
    for (int p = 0; p < 6; ++p) E6[p] = Ev[p];
  } else {
    E6[0] = Ev[0], E6[1] = Ev[1], E6[5] = Ev[2];
  }
  bool ok = true;
  if constexpr (D == 2) {
    const bool fullType = A.resultType == IKB_RESULT_LINEAR_STRESS_FULL || A.resultType == IKB_RESULT_PK2_STRESS_FULL;
    if (A.planeStrain == IKB_REDUCE_PLANE_STRESS && !fullType) {
      // planeStress: the out-of-plane normal strain of the reduced solution (vanishingstress.hh:150-199); the *Full
      // types evaluate the 3D law at the zero-extended strain, exactly like enlargeIfReduced in calculateStress
      if (A.form == FORM_NH) {
        const double det2 = (2.0 * E6[0] + 1.0) * (2.0 * E6[1] + 1.0) - E6[5] * E6[5];
        double c33, lnJ;
        ok = reduceC33(A.lambda, A.mu, A.psTol, det2, c33, lnJ);
        E6[2] = 0.5 * (c33 - 1.0);
      } else {
        ok = reduceE33(A.lambda, A.mu, A.psTol, E6[0] + E6[1], E6[2]);
      }
    }
  }
  if (!(ok && stress3d(A.form, A.lambda, A.mu, A.ps, E6, S6))) {
    atomicMin(A.errFlag, (int32_t)(e < 0x7fffffff ? e : 0x7ffffffe));
    for (int p = 0; p < 6; ++p) S6[p] = 0.0;
  }
  double* out = A.out + (size_t)t * A.ncomp;
  if (A.resultType == IKB_RESULT_LINEAR_STRESS_FULL || A.resultType == IKB_RESULT_PK2_STRESS_FULL) {
    for (int p = 0; p < 6; ++p) out[p] = S6[p];  // the underlying 3D law (6 components also in 2D)
    return;
  }
  double Sv[S];
  if constexpr (D == 3) {
    for (int p = 0; p < 6; ++p) Sv[p] = S6[p];
  } else {
    Sv[0] = S6[0], Sv[1] = S6[1], Sv[2] = S6[5];
  }
  if (A.resultType == IKB_RESULT_KIRCHHOFF_STRESS || A.resultType == IKB_RESULT_CAUCHY_STRESS) {
    // tau = F S F^T, sigma = tau / det F  (transformStress<PK2, Kirchhoff|Cauchy>)
    double F[D][D], Sm[D][D];
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) F[i][j] = H[i][j] + (i == j ? 1.0 : 0.0);
    for (int p = 0; p < S; ++p) {
      int i, j;
      voigtPair<D>(p, i, j);
      Sm[i][j] = Sm[j][i] = Sv[p];
    }
    double scale = 1.0;
    if (A.resultType == IKB_RESULT_CAUCHY_STRESS) {
      double Fi[D][D];
      scale = 1.0 / invSmall<D>(F, Fi);
    }
    for (int p = 0; p < S; ++p) {
      int i, j;
      voigtPair<D>(p, i, j);
      double v = 0.0;
      for (int k = 0; k < D; ++k)
        for (int l = 0; l < D; ++l) v += F[i][k] * Sm[k][l] * F[j][l];
      out[p] = v * scale;
    }
    return;
  }
  for (int p = 0; p < S; ++p) out[p] = Sv[p];
}

}  // namespace ikb
