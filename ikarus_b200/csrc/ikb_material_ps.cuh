// Principal-stretch hyperelastic laws on the device: Materials::Hyperelastic<Deviatoric<DF>, Volumetric<VF>>
// (ikarus/finiteelements/mechanics/materials/hyperelastic/interface.hh:99-232, deviatoric/interface.hh:77-115,
// materialhelpers.hh:138-164), input C = 2E + I.
//
//   lambda_i = sqrt(eig_i(C)), N = eigenvectors
//   S   = sum_i S_i N_i (x) N_i,                S_i = W,i / lambda_i          (+ J U'(J) C^-1, none for VF0)
//   CC  = sum_ik L1_ik N_iN_i (x) N_kN_k + sum_{i != k} L2_ik N_iN_k (x) (N_iN_k + N_kN_i)
//   L1_ik = dS(i,k) / (lambda_i lambda_k),      L2_ik = (S_i - S_k) / (lambda_i^2 - lambda_k^2),
//           and 0.5 (L1_ii - L1_ik) when Dune::FloatCmp::eq(lambda_i, lambda_k, 1e-8) (relativeWeak)
//
// The kernels never form the 3^4 tensor: for a symmetric B they need CC : B = N Chat N^T with, in the principal frame,
// Chat_ii = sum_k L1_ik Bhat_kk, Chat_ik = 2 L2_ik Bhat_ik (Bhat = N^T B N).  Plane strain (materials/vanishingstrain.hh)
// is the same law with lambda_3 = 1, N_3 = e_3 and in-plane B: only the in-plane 2 x 2 parts of N, L1, L2 are used.
#pragma once
#include "ikb_internal.cuh"

namespace ikb {

// Deviatoric function: Blatz-Ko, W = mu/2 (sum lambda_i^-2 + 2 J - 5)   (deviatoric/blatzko.hh:60-92)
// in : mu, lam[3];  out: psi, principal PK2 stresses Sp[i] = W,i / lambda_i, L1[i][k] = dS(i,k) / (lambda_i lambda_k)
__device__ __forceinline__ void blatzKoPrincipal(double mu, const double (&lam)[3], double& psi, double (&Sp)[3],
                                                 double (&L1)[3][3]) {
  const double J = lam[0] * lam[1] * lam[2];
  double il2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) il2[i] = 1.0 / (lam[i] * lam[i]);
  psi = 0.5 * mu * (il2[0] + il2[1] + il2[2] + 2.0 * J - 5.0);
#pragma unroll
  for (int i = 0; i < 3; ++i) Sp[i] = mu * (-il2[i] * il2[i] + J * il2[i]);  // (-lam^-3 + J/lam) / lam
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // dS(i,k) = mu J / (lam_i lam_k) (i != k),  mu (lam_i^-2 (lam_i^-2 - J) + 3 lam_i^-4) (i == k)
      L1[i][k] = (i == k) ? mu * (il2[i] * (il2[i] - J) + 3.0 * il2[i] * il2[i]) * il2[i] : mu * J * il2[i] * il2[k];
    }
}

// L2[i][k], i != k, from the principal stresses (deviatoric/interface.hh:98-106)
__device__ __forceinline__ void principalShearModuli(const double (&lam)[3], const double (&Sp)[3], const double (&L1)[3][3],
                                                     double (&L2)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (i == k) {
        L2[i][k] = 0.0;
      } else {
        const bool close = fabs(lam[i] - lam[k]) <= 1e-8 * fmax(fabs(lam[i]), fabs(lam[k]));
        L2[i][k] = close ? 0.5 * (L1[i][i] - L1[i][k]) : (Sp[i] - Sp[k]) / (lam[i] * lam[i] - lam[k] * lam[k]);
      }
    }
}

// Eigen-decomposition of a symmetric 3 x 3 matrix by cyclic Jacobi rotations (converges quadratically; eight sweeps reach
// machine precision for any input).  V's columns are the eigenvectors; no ordering (the laws above are invariant).
__device__ __forceinline__ void eigSym3(const double (&A)[3][3], double (&w)[3], double (&V)[3][3]) {
  double a[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      a[i][j] = A[i][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll 1
  for (int sweep = 0; sweep < 8; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off == 0.0) break;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int p = (r == 2) ? 1 : 0, q = (r == 0) ? 1 : 2;  // (0,1), (0,2), (1,2)
      const double apq = a[p][q];
      if (apq != 0.0) {
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        a[p][p] -= t * apq;
        a[q][q] += t * apq;
        a[p][q] = a[q][p] = 0.0;
        const int o = 3 - p - q;  // the third index
        const double aop = a[o][p], aoq = a[o][q];
        a[o][p] = a[p][o] = c * aop - s * aoq;
        a[o][q] = a[q][o] = s * aop + c * aoq;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) w[i] = a[i][i];
}

// The law at C (D x D in-plane part; plane strain in 2D): principal frame N (3 x 3, in 2D the in-plane 2 x 2 block and
// N_3 = e_3), principal stresses, L1, L2 and psi.  Returns false when C is not positive definite.
template <int D>
__device__ __forceinline__ bool principalLaw(double mu, const double (&Cm)[D][D], double (&N)[3][3], double (&Sp)[3],
                                             double (&L1)[3][3], double (&L2)[3][3], double& psi) {
  double C3[3][3], ev[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C3[i][j] = (i < D && j < D) ? Cm[i < D ? i : 0][j < D ? j : 0] : (i == j ? 1.0 : 0.0);
  eigSym3(C3, ev, N);
  if (!(ev[0] > 0.0 && ev[1] > 0.0 && ev[2] > 0.0)) return false;
  double lam[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) lam[i] = sqrt(ev[i]);
  blatzKoPrincipal(mu, lam, psi, Sp, L1);
  principalShearModuli(lam, Sp, L1, L2);
  return true;
}

// S = N diag(Sp) N^T, in-plane D x D part
template <int D>
__device__ __forceinline__ void principalStress(const double (&N)[3][3], const double (&Sp)[3], double (&Sm)[D][D]) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s = fma(Sp[k] * N[i][k], N[j][k], s);
      Sm[i][j] = s;
    }
}

// CB = CC : B for a symmetric in-plane B (D x D), in-plane part of the result
template <int D>
__device__ __forceinline__ void principalTangentTimes(const double (&N)[3][3], const double (&L1)[3][3],
                                                      const double (&L2)[3][3], const double (&B)[D][D],
                                                      double (&CB)[D][D]) {
  double T[D][3], Bh[3][3];  // T = B N (in-plane rows), Bhat = N^T B N
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fma(B[i][j], N[j][k], s);
      T[i][k] = s;
    }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) s = fma(N[i][a], T[i][k], s);
      Bh[a][k] = s;
    }
  double Ch[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (i == k) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < 3; ++m) s = fma(L1[i][m], Bh[m][m], s);
        Ch[i][i] = s;
      } else {
        Ch[i][k] = 2.0 * L2[i][k] * Bh[i][k];
      }
    }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) s = fma(N[i][a] * Ch[a][k], N[j][k], s);
      CB[i][j] = s;
    }
}

}  // namespace ikb
