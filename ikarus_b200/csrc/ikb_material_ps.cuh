// Principal-stretch hyperelastic laws on the device: Materials::Hyperelastic<Deviatoric<DF>, Volumetric<VF>> with every
// deviatoric function (BlatzKo, Ogden, InvariantBased = Mooney-Rivlin / Yeoh, ArrudaBoyce, Gent) and VF0 .. VF12
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

__device__ __forceinline__ double ipow(double x, int p) {  // x^p for the small integer exponents of InvariantBased
  double r = 1.0;
  const int n = p < 0 ? -p : p;
  for (int i = 0; i < n; ++i) r *= x;
  return p < 0 ? 1.0 / r : r;
}

// Deviatoric function: W, dW/dlambda_i and the reference's "second derivative" array dS = Hess W - diag(W,i / lambda_i)
// (the array Deviatoric::tangentModuli divides by lambda_i lambda_k, deviatoric/interface.hh:95-99).
//   BlatzKo (blatzko.hh:60-92), Ogden on total / deviatoric stretches (ogden.hh:98-190); the invariant-based family
//   (invariantbased.hh:84-188, arrudaboyce.hh:88-160, gent.hh:88-150) is W = f(W1, W2) of the deviatoric invariants
//   W1 = I1 I3^(-1/3), W2 = I2 I3^(-2/3) (deviatoricinvariants.hh:52-107), so with f1, f2, f11, f22, f12
//   dW = f1 dW1 + f2 dW2,   Hess W = f1 ddW1 + f2 ddW2 + f11 dW1 dW1^T + f22 dW2 dW2^T + f12 (dW1 dW2^T + dW2 dW1^T).
// Returns false where the reference throws (Gent: Jm <= W1 - 3).
__device__ __noinline__ bool psDeviatoric(const PsLaw& law, const double* lam, double* Wout, double* dW, double* dS) {
  const double J = lam[0] * lam[1] * lam[2];
  double W = 0.0;
  for (int i = 0; i < 3; ++i) {
    dW[i] = 0.0;
    for (int k = 0; k < 3; ++k) dS[3 * i + k] = 0.0;
  }
  bool ok = true;
  if (law.dev == 1) {
    const double mu = law.par[0];
    double il2[3];
    for (int i = 0; i < 3; ++i) il2[i] = 1.0 / (lam[i] * lam[i]);
    W = 0.5 * mu * (il2[0] + il2[1] + il2[2] + 2.0 * J - 5.0);
    for (int i = 0; i < 3; ++i) dW[i] = mu * (-il2[i] / lam[i] + J / lam[i]);
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k)
        dS[3 * i + k] = (i == k) ? mu * (il2[i] * (il2[i] - J) + 3.0 * il2[i] * il2[i]) : mu * J / (lam[i] * lam[k]);
  } else if (law.dev == 2) {
    const double logJ = log(J);
    for (int p = 0; p < law.n; ++p) {
      const double mu = law.par[p], al = law.ex[p];
      double s = 0.0;
      for (int i = 0; i < 3; ++i) {
        const double la = pow(lam[i], al);
        s += la;
        dW[i] += mu * (la - 1.0) / lam[i];
        dS[4 * i] += (-2.0 * mu * (la - 1.0) + mu * la * al) / (lam[i] * lam[i]);
      }
      W += mu / al * (s - 3.0) - mu * logJ;
    }
  } else if (law.dev == 3) {
    const double jm = pow(J, -1.0 / 3.0);
    double lb[3], dWb[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < 3; ++i) lb[i] = jm * lam[i];
    for (int p = 0; p < law.n; ++p) {
      const double mu = law.par[p], al = law.ex[p];
      double la[3], s = 0.0;
      for (int i = 0; i < 3; ++i) {
        la[i] = pow(lb[i], al);
        s += la[i];
        dWb[i] += mu * la[i] / lb[i];
      }
      W += mu / al * (s - 3.0);
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
          dS[3 * a + b] += (a == b) ? mu * al * (la[a] / 3.0 + s / 9.0) : mu * al * (-(la[a] + la[b]) / 3.0 + s / 9.0);
    }
    const double sum = lb[0] * dWb[0] + lb[1] * dWb[1] + lb[2] * dWb[2];
    for (int i = 0; i < 3; ++i) dW[i] = (lb[i] * dWb[i] - sum / 3.0) / lam[i];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        dS[3 * a + b] /= lam[a] * lam[b];
        if (a == b) dS[3 * a + b] -= 2.0 / lam[a] * dW[a];
      }
  } else if (law.dev >= 4) {
    double l2[3];
    for (int i = 0; i < 3; ++i) l2[i] = lam[i] * lam[i];
    const double I1 = l2[0] + l2[1] + l2[2], I2 = l2[0] * l2[1] + l2[1] * l2[2] + l2[0] * l2[2], I3 = l2[0] * l2[1] * l2[2];
    const double p13 = cbrt(I3), p23 = p13 * p13;
    const double W1 = I1 / p13, W2 = I2 / p23;
    double d1[3], d2[3], dd1[9], dd2[9];
    for (int i = 0; i < 3; ++i) {
      d1[i] = 2.0 * (3.0 * l2[i] - I1) / (3.0 * lam[i] * p13);
      d2[i] = -2.0 * (3.0 * I3 / l2[i] - I2) / (3.0 * lam[i] * p23);
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        if (i == j) {
          dd1[4 * i] = (2.0 / 9.0) * (5.0 * I1 - 3.0 * l2[i]) / (l2[i] * p13);
          dd2[4 * i] = (2.0 / 9.0) * (15.0 * I3 / l2[i] - I2) / (l2[i] * p23);
        } else {
          dd1[3 * i + j] = (4.0 / 9.0) * (I1 - 3.0 * (l2[i] + l2[j])) / (lam[i] * lam[j] * p13);
          dd2[3 * i + j] = (-4.0 / 9.0) * (2.0 * I2 - 3.0 * l2[i] * l2[j]) / (lam[i] * lam[j] * p23);
        }
      }
    double f1 = 0.0, f2 = 0.0, f11 = 0.0, f22 = 0.0, f12 = 0.0;
    if (law.dev == 4) {
      const double a = W1 - 3.0, b = W2 - 3.0;
      for (int t = 0; t < law.n; ++t) {
        const int p = law.pex[t], q = law.qex[t];
        const double c = law.par[t];
        const double ap = ipow(a, p), bq = ipow(b, q);
        const double ap1 = p ? p * ipow(a, p - 1) : 0.0, bq1 = q ? q * ipow(b, q - 1) : 0.0;
        const double ap2 = p > 1 ? p * (p - 1) * ipow(a, p - 2) : 0.0, bq2 = q > 1 ? q * (q - 1) * ipow(b, q - 2) : 0.0;
        W += c * ap * bq;
        f1 += c * ap1 * bq;
        f2 += c * ap * bq1;
        f11 += c * ap2 * bq;
        f22 += c * ap * bq2;
        f12 += c * ap1 * bq1;
      }
    } else if (law.dev == 5) {
      const double mu = law.par[0], beta = 1.0 / (law.par[1] * law.par[1]);
      const double al[5] = {0.5, 1.0 / 20.0, 11.0 / 1050.0, 19.0 / 7000.0, 519.0 / 673750.0};
      double bp = 1.0;
      for (int p = 0; p < 5; ++p) {
        W += mu * al[p] * bp * (ipow(W1, p + 1) - ipow(3.0, p + 1));
        f1 += mu * al[p] * bp * (p + 1) * ipow(W1, p);
        if (p) f11 += mu * al[p] * bp * p * (p + 1) * ipow(W1, p - 1);
        bp *= beta;
      }
    } else {
      const double mu = law.par[0], Jm = law.par[1];
      const double fac = 1.0 - (W1 - 3.0) / Jm;
      ok = Jm > W1 - 3.0;
      W = -(mu / 2.0) * Jm * log(fac);
      f1 = mu / (2.0 * fac);
      f11 = mu / (2.0 * fac * fac * Jm);
    }
    for (int i = 0; i < 3; ++i) dW[i] = f1 * d1[i] + f2 * d2[i];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double v = f1 * dd1[3 * i + j] + f2 * dd2[3 * i + j] + f11 * d1[i] * d1[j] + f22 * d2[i] * d2[j] +
                   f12 * (d1[i] * d2[j] + d1[j] * d2[i]);
        if (i == j) v -= dW[i] / lam[i];
        dS[3 * i + j] = v;
      }
  }
  *Wout = W;
  return ok;
}

// Volumetric function VF0 .. VF12 (volumetric/volumetricfunctions.hh:25-380): U, U', U'' at J
__device__ __noinline__ void psVolumetric(int vf, double b, double J, double* U, double* Up, double* Upp) {
  const double lnJ = log(J);
  double u = 0.0, u1 = 0.0, u2 = 0.0;
  switch (vf) {
    case 1: u = 0.5 * (J - 1.0) * (J - 1.0), u1 = J - 1.0, u2 = 1.0; break;
    case 2: u = 0.25 * ((J - 1.0) * (J - 1.0) + lnJ * lnJ), u1 = 0.5 * (J - 1.0 + lnJ / J), u2 = (1.0 + J * J - lnJ) / (2.0 * J * J); break;
    case 3: u = 0.5 * lnJ * lnJ, u1 = lnJ / J, u2 = (1.0 - lnJ) / (J * J); break;
    case 4: u = (pow(J, -b) - 1.0 + b * lnJ) / (b * b), u1 = (1.0 / J - pow(J, -1.0 - b)) / b, u2 = pow(J, -2.0 - b) * (1.0 + b - pow(J, b)) / b; break;
    case 5: u = 0.25 * (J * J - 1.0 - 2.0 * lnJ), u1 = 0.5 * (J - 1.0 / J), u2 = 0.5 * (1.0 + 1.0 / (J * J)); break;
    case 6: u = J - lnJ - 1.0, u1 = 1.0 - 1.0 / J, u2 = 1.0 / (J * J); break;
    case 7: u = pow(J, b) * (b * lnJ - 1.0) + 1.0, u1 = b * b * pow(J, b - 1.0) * lnJ, u2 = b * b * pow(J, b - 2.0) * (1.0 + (b - 1.0) * lnJ); break;
    case 8: u = J * lnJ - J + 1.0, u1 = lnJ, u2 = 1.0 / J; break;
    case 9: {
      const double J2 = J * J, d = J2 - 1.0 / J2;
      u = d * d / 32.0, u1 = (J2 * J - 1.0 / (J2 * J2 * J)) / 8.0, u2 = (5.0 / (J2 * J2 * J2) + 3.0 * J2) / 8.0;
    } break;
    case 10: u = (J / b) * (1.0 - pow(J, -b) / (1.0 - b)) + 1.0 / (b - 1.0), u1 = (1.0 - pow(J, -b)) / b, u2 = pow(J, -1.0 - b); break;
    case 11: u = (ipow(J, 5) + ipow(J, -5) - 2.0) / 50.0, u1 = (ipow(J, 4) - ipow(J, -6)) / 10.0, u2 = (4.0 * ipow(J, 3) + 6.0 * ipow(J, -7)) / 10.0; break;
    case 12: u = J - 1.0, u1 = 1.0, u2 = 0.0; break;
    default: break;
  }
  *U = u, *Up = u1, *Upp = u2;
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
// N_3 = e_3), principal stresses, L1, L2 and psi of Hyperelastic<Deviatoric<DF>, Volumetric<VF>> (interface.hh:99-216).  With
// C^-1 = N diag(lambda^-2) N^T the volumetric part  S += J U' C^-1,  CC += J ((U' + J U'') C^-1 (x) C^-1 - 2 U' sym(C^-1 (.)
// C^-1))  is diagonal in the same frame:  L1_ik += J (U' + J U'') / (l_i^2 l_k^2) - delta_ik 2 J U' / l_i^4,
// L2_ik -= J U' / (l_i^2 l_k^2).  Returns false where the reference aborts or throws (det C <= 0, Gent's Jm).
template <int D>
__device__ __forceinline__ bool principalLaw(const PsLaw& law, const double (&Cm)[D][D], double (&N)[3][3], double (&Sp)[3],
                                             double (&L1)[3][3], double (&L2)[3][3], double& psi) {
  double C3[3][3], ev[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C3[i][j] = (i < D && j < D) ? Cm[i < D ? i : 0][j < D ? j : 0] : (i == j ? 1.0 : 0.0);
  eigSym3(C3, ev, N);
  if (!(ev[0] > 0.0 && ev[1] > 0.0 && ev[2] > 0.0)) return false;
  double lam[3], dW[3], dS[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) lam[i] = sqrt(ev[i]);
  bool ok = psDeviatoric(law, lam, &psi, dW, dS);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Sp[i] = dW[i] / lam[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) L1[i][k] = dS[3 * i + k] / (lam[i] * lam[k]);
  }
  principalShearModuli(lam, Sp, L1, L2);
  if (law.vf) {
    const double J = lam[0] * lam[1] * lam[2];
    double U, Up, Upp;
    psVolumetric(law.vf, law.beta, J, &U, &Up, &Upp);
    U *= law.K, Up *= law.K, Upp *= law.K;
    psi += U;
    double il2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) il2[i] = 1.0 / ev[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Sp[i] = fma(J * Up, il2[i], Sp[i]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        L1[i][k] += J * (Up + J * Upp) * il2[i] * il2[k];
        if (i == k)
          L1[i][k] -= 2.0 * J * Up * il2[i] * il2[i];
        else
          L2[i][k] -= J * Up * il2[i] * il2[k];
      }
    }
  }
  return ok;
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
