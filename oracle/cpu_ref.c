/*
 * cpu_ref.c -- C restatement of the reference's CPU assembly path (TEST/BASELINE INFRASTRUCTURE ONLY).
 *
 * This is the "port" CPU baseline of bench.py and a second checker for the CUDA path; the product
 * never links or calls it.  It keeps the reference's loop structure on purpose:
 *   - one sweep for R and a separate sweep for K (newtonraphson.hh:242-243 calls residual and
 *     jacobian separately; simpleassemblers.inl:59-76 and :120-137),
 *   - per Gauss point: strain, stress, then for every node pair (i,j) the material tangent is
 *     re-evaluated and B_i^T C B_j + geometric term added (nonlinearelastic.hh:387-398, :316-321),
 *   - NeoHooke through C = 2E+I, inverse, log(sqrt(det C)), 4th-order tensor -> Voigt
 *     (materials/hyperelastic/neohooke.hh:79-142, utils/tensorutils.hh:219-227),
 *   - scatter through precomputed linear indices (simpleassemblers.inl:133-134).
 * Parity: pinned against oracle/ikarus_oracle.py (which reproduces the reference's own known-answer
 * tests) in tests/test_cpu_ref_port.py.  All file:line citations are relative to /root/reference.
 *
 * Q1 elements only (Quad4 plane strain, Hex8); materials 0 = LinearElasticity (linear strain),
 * 1 = StVenantKirchhoff, 2 = NeoHooke (Green-Lagrange strain).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 3
#define MAXN 8
#define MAXS 6

typedef struct {
  int dim, material;
  double lam, mu;
} ikref_cfg;

static const int VP3[6][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 2}, {0, 2}, {0, 1}};
static const int VP2[3][2] = {{0, 0}, {1, 1}, {0, 1}};

static double inv_small(int d, const double A[MAXD][MAXD], double Ai[MAXD][MAXD]) {
  if (d == 2) {
    double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    Ai[0][0] = A[1][1] / det;
    Ai[0][1] = -A[0][1] / det;
    Ai[1][0] = -A[1][0] / det;
    Ai[1][1] = A[0][0] / det;
    return det;
  }
  double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
  Ai[0][0] = c00 / det;
  Ai[1][0] = c01 / det;
  Ai[2][0] = c02 / det;
  Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det;
  Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det;
  Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det;
  Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
  Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
  Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
  return det;
}

/* 3D law on the Voigt GL strain: stress S6 and tangent C66 (and energy) */
static int law3d(const ikref_cfg* c, const double E6[6], double* psi, double S6[6], double C66[6][6]) {
  const double lam = c->lam, mu = c->mu;
  memset(C66, 0, 36 * sizeof(double));
  if (c->material != 2) { /* svk.hh:77-164 */
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) C66[i][j] = lam;
      C66[i][i] += 2 * mu;
      C66[3 + i][3 + i] = mu;
    }
    for (int p = 0; p < 6; ++p) {
      S6[p] = 0;
      for (int q = 0; q < 6; ++q) S6[p] += C66[p][q] * E6[q];
    }
    double tr = E6[0] + E6[1] + E6[2];
    *psi = 0.5 * lam * tr * tr +
           mu * (E6[0] * E6[0] + E6[1] * E6[1] + E6[2] * E6[2] + 0.5 * (E6[3] * E6[3] + E6[4] * E6[4] + E6[5] * E6[5]));
    return 0;
  }
  /* neohooke.hh:79-142 */
  double Cm[MAXD][MAXD], Ci[MAXD][MAXD];
  for (int q = 0; q < 6; ++q) {
    int i = VP3[q][0], j = VP3[q][1];
    double v = (i == j) ? 2 * E6[q] + 1.0 : E6[q]; /* 2 * (E_voigt/2) */
    Cm[i][j] = Cm[j][i] = v;
  }
  double detC = inv_small(3, Cm, Ci);
  if (!(detC > 0.0)) return 1; /* materialhelpers.hh:120-126: FloatCmp::le(det, 0, 1e-10) with the default relativeWeak style == det <= 0 */
  double logJ = log(sqrt(detC));
  *psi = 0.5 * mu * (Cm[0][0] + Cm[1][1] + Cm[2][2] - 3 - 2 * logJ) + 0.5 * lam * logJ * logJ;
  double T[3][3][3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l)
          T[i][j][k][l] = lam * Ci[i][j] * Ci[k][l] +
                          2 * (mu - lam * logJ) * 0.5 * (Ci[i][k] * Ci[j][l] + Ci[i][l] * Ci[j][k]);
  for (int p = 0; p < 6; ++p) {
    S6[p] = mu * ((VP3[p][0] == VP3[p][1]) - Ci[VP3[p][0]][VP3[p][1]]) + lam * logJ * Ci[VP3[p][0]][VP3[p][1]];
    for (int q = 0; q < 6; ++q) C66[p][q] = T[VP3[p][0]][VP3[p][1]][VP3[q][0]][VP3[q][1]];
  }
  return 0;
}

/* material call on a dim-sized Voigt strain, with the planeStrain reduction in 2D (vanishingstrain.hh:79-120) */
static int material(const ikref_cfg* c, const double* Ev, double* psi, double* S, double* C /* s x s */) {
  double E6[6] = {0, 0, 0, 0, 0, 0}, S6[6], C66[6][6];
  static const int free2[3] = {0, 1, 5};
  const int s = c->dim == 3 ? 6 : 3;
  if (c->dim == 3)
    memcpy(E6, Ev, 6 * sizeof(double));
  else
    for (int p = 0; p < 3; ++p) E6[free2[p]] = Ev[p];
  if (law3d(c, E6, psi, S6, C66)) return 1;
  for (int p = 0; p < s; ++p) {
    int pp = c->dim == 3 ? p : free2[p];
    S[p] = S6[pp];
    for (int q = 0; q < s; ++q) C[p * s + q] = C66[pp][c->dim == 3 ? q : free2[q]];
  }
  return 0;
}

typedef struct {
  double gradN[MAXN][MAXD];
  double F[MAXD][MAXD];
  double Ev[MAXS];
  double wdet;
} gp_kin;

static void kinematics(const ikref_cfg* c, const double* X, const double* u, int g, gp_kin* k) {
  const int d = c->dim, n = 1 << d;
  const double lo = 0.5 - 0.5 / sqrt(3.0), hi = 0.5 + 0.5 / sqrt(3.0);
  double xi[MAXD], dN[MAXN][MAXD], Jt[MAXD][MAXD], Ji[MAXD][MAXD], H[MAXD][MAXD];
  for (int a = 0; a < d; ++a) xi[a] = ((g >> a) & 1) ? hi : lo;
  for (int a = 0; a < n; ++a)
    for (int i = 0; i < d; ++i) {
      double v = ((a >> i) & 1) ? 1.0 : -1.0;
      for (int q = 0; q < d; ++q)
        if (q != i) v *= ((a >> q) & 1) ? xi[q] : 1.0 - xi[q];
      dN[a][i] = v;
    }
  memset(Jt, 0, sizeof(Jt));
  for (int a = 0; a < n; ++a)
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) Jt[i][j] += dN[a][i] * X[a * d + j];
  double det = fabs(inv_small(d, Jt, Ji));
  k->wdet = det;
  for (int a = 0; a < d; ++a) k->wdet *= 0.5;
  memset(H, 0, sizeof(H));
  for (int a = 0; a < n; ++a)
    for (int j = 0; j < d; ++j) {
      double s = 0;
      for (int i = 0; i < d; ++i) s += Ji[j][i] * dN[a][i];
      k->gradN[a][j] = s;
    }
  for (int a = 0; a < n; ++a)
    for (int cc = 0; cc < d; ++cc)
      for (int j = 0; j < d; ++j) H[cc][j] += u[a * d + cc] * k->gradN[a][j];
  const int(*vp)[2] = d == 3 ? VP3 : VP2;
  const int s = d * (d + 1) / 2;
  const int gl = c->material != 0;
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) k->F[i][j] = (gl ? H[i][j] : 0.0) + (i == j);
  for (int q = 0; q < s; ++q) {
    int i = vp[q][0], j = vp[q][1];
    double e = 0.5 * (H[i][j] + H[j][i]);
    if (gl)
      for (int m = 0; m < d; ++m) e += 0.5 * H[m][i] * H[m][j];
    k->Ev[q] = (i == j) ? e : 2 * e;
  }
}

/* B_a (s x d): displacementgradient.hh:123-147 restatement of the GL B-operator */
static void bop(const ikref_cfg* c, const gp_kin* k, int a, double* B) {
  const int d = c->dim, s = d * (d + 1) / 2;
  const int(*vp)[2] = d == 3 ? VP3 : VP2;
  for (int q = 0; q < s; ++q) {
    int i = vp[q][0], j = vp[q][1];
    for (int m = 0; m < d; ++m)
      B[q * d + m] = (i == j) ? k->gradN[a][i] * k->F[m][i] : k->gradN[a][j] * k->F[m][i] + k->gradN[a][i] * k->F[m][j];
  }
}

/* calculateVector: R_e (nonlinearelastic.hh:403-430) */
static int elem_vector(const ikref_cfg* c, const double* X, const double* u, double* Re) {
  const int d = c->dim, n = 1 << d, s = d * (d + 1) / 2;
  memset(Re, 0, n * d * sizeof(double));
  for (int g = 0; g < n; ++g) {
    gp_kin k;
    kinematics(c, X, u, g, &k);
    double psi, S[MAXS], C[MAXS * MAXS], B[MAXS * MAXD];
    if (material(c, k.Ev, &psi, S, C)) return 1;
    if (c->material == 0) /* linearelastic.hh: stress = C * eps */
      for (int p = 0; p < s; ++p) {
        S[p] = 0;
        for (int q = 0; q < s; ++q) S[p] += C[p * s + q] * k.Ev[q];
      }
    for (int a = 0; a < n; ++a) {
      bop(c, &k, a, B);
      for (int m = 0; m < d; ++m) {
        double v = 0;
        for (int q = 0; q < s; ++q) v += B[q * d + m] * S[q];
        Re[a * d + m] += v * k.wdet;
      }
    }
  }
  return 0;
}

/* calculateMatrix: K_e with the tangent re-evaluated for every (i,j) (nonlinearelastic.hh:376-400, :316-321) */
static int elem_matrix(const ikref_cfg* c, const double* X, const double* u, double* Ke) {
  const int d = c->dim, n = 1 << d, s = d * (d + 1) / 2, nd = n * d;
  const int(*vp)[2] = d == 3 ? VP3 : VP2;
  memset(Ke, 0, nd * nd * sizeof(double));
  for (int g = 0; g < n; ++g) {
    gp_kin k;
    kinematics(c, X, u, g, &k);
    double psi, S[MAXS], C[MAXS * MAXS], Bi[MAXS * MAXD], Bj[MAXS * MAXD];
    if (material(c, k.Ev, &psi, S, C)) return 1;
    double Sm[MAXD][MAXD];
    for (int q = 0; q < s; ++q) Sm[vp[q][0]][vp[q][1]] = Sm[vp[q][1]][vp[q][0]] = S[q];
    for (int i = 0; i < n; ++i) {
      bop(c, &k, i, Bi);
      for (int j = 0; j < n; ++j) {
        bop(c, &k, j, Bj);
        if (material(c, k.Ev, &psi, S, C)) return 1; /* materialTangent(strain) per node pair */
        for (int a = 0; a < d; ++a)
          for (int b = 0; b < d; ++b) {
            double v = 0;
            for (int p = 0; p < s; ++p) {
              double t = 0;
              for (int q = 0; q < s; ++q) t += C[p * s + q] * Bj[q * d + b];
              v += Bi[p * d + a] * t;
            }
            Ke[(i * d + a) * nd + j * d + b] += v * k.wdet;
          }
        if (c->material != 0) { /* geometric stiffness */
          double kg = 0;
          for (int a = 0; a < d; ++a)
            for (int b = 0; b < d; ++b) kg += k.gradN[i][a] * Sm[a][b] * k.gradN[j][b];
          for (int a = 0; a < d; ++a) Ke[(i * d + a) * nd + j * d + a] += kg * k.wdet;
        }
      }
    }
  }
  return 0;
}

/* Per-element K_e / R_e for checking (K row-major nd x nd) */
int ikref_element(int dim, int material_id, double lam, double mu, const double* X, const double* u, double* Ke,
                  double* Re) {
  ikref_cfg c = {dim, material_id, lam, mu};
  if (Re && elem_vector(&c, X, u, Re)) return 1;
  if (Ke && elem_matrix(&c, X, u, Ke)) return 1;
  return 0;
}

/*
 * Raw global assembly of K (CSR/CSC value array through precomputed linear indices, column-major over
 * K_e: for c, for r -- simpleassemblers.inl:253-266) and R, as two sweeps.  corner[e][n][d],
 * edofs[e][n*d], linidx[e][nd*nd].  nthreads <= 1: serial like the reference; > 1: OpenMP over elements
 * with atomic scatter (baseline only).  Returns the number of failed elements.
 */
int ikref_assemble(int dim, int material_id, double lam, double mu, int64_t nElem, const double* corner,
                   const int64_t* edofs, const int64_t* linidx, const double* d, double* vals, int64_t nnz, double* R,
                   int64_t nDof, int nthreads) {
  ikref_cfg c = {dim, material_id, lam, mu};
  const int n = 1 << dim, nd = n * dim;
  int failed = 0;
  if (R) memset(R, 0, nDof * sizeof(double));
  if (vals) memset(vals, 0, nnz * sizeof(double));
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  if (R) {
#pragma omp parallel for schedule(static) reduction(+ : failed) if (nthreads > 1)
    for (int64_t e = 0; e < nElem; ++e) {
      double u[MAXN * MAXD], Re[MAXN * MAXD];
      for (int i = 0; i < nd; ++i) u[i] = d[edofs[e * nd + i]];
      failed += elem_vector(&c, corner + e * n * dim, u, Re);
      for (int i = 0; i < nd; ++i) {
#pragma omp atomic
        R[edofs[e * nd + i]] += Re[i];
      }
    }
  }
  if (vals) {
#pragma omp parallel for schedule(static) reduction(+ : failed) if (nthreads > 1)
    for (int64_t e = 0; e < nElem; ++e) {
      double u[MAXN * MAXD], Ke[MAXN * MAXD * MAXN * MAXD];
      for (int i = 0; i < nd; ++i) u[i] = d[edofs[e * nd + i]];
      failed += elem_matrix(&c, corner + e * n * dim, u, Ke);
      int64_t q = 0;
      for (int cc = 0; cc < nd; ++cc)
        for (int r = 0; r < nd; ++r, ++q) {
#pragma omp atomic
          vals[linidx[e * nd * nd + q]] += Ke[r * nd + cc];
        }
    }
  }
  return failed;
}

/*
 * "cpu_opt" baseline of SURVEY.md 8d: the SAME math as the device kernels (factored NeoHooke tangent of
 * ikarus_b200/csrc/ikb_elem_q1.cuh -- K_ab = sum_g w [ lam m_a m_b^T + mu' m_b m_a^T + mu (g_a.g_b) I ],
 * m_a = F^-T g_a, mu' = mu - lam ln J; R_a = sum_g w (mu F - mu' F^-T) g_a), K and R in ONE sweep, the block symmetry
 * used, OpenMP over all host cores with atomic scatter.  What a tuned CPU implementation of this path would look like;
 * the faithful port above keeps the reference's loop structure instead.  Hex8 / Quad4-plane-strain NeoHooke only.
 * Same argument list and result layout as ikref_assemble (linidx column-major over K_e).
 */
int ikref_assemble_opt(int dim, int material_id, double lam, double mu, int64_t nElem, const double* corner,
                       const int64_t* edofs, const int64_t* linidx, const double* d, double* vals, int64_t nnz,
                       double* R, int64_t nDof, int nthreads) {
  if (material_id != 2 || !vals || !R) return -1;
  ikref_cfg c = {dim, material_id, lam, mu};
  const int n = 1 << dim, nd = n * dim;
  int failed = 0;
  memset(R, 0, nDof * sizeof(double));
  memset(vals, 0, nnz * sizeof(double));
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : failed) if (nthreads > 1)
  for (int64_t e = 0; e < nElem; ++e) {
    double u[MAXN * MAXD], Re[MAXN * MAXD], Ke[MAXN * MAXD * MAXN * MAXD];
    for (int i = 0; i < nd; ++i) u[i] = d[edofs[e * nd + i]];
    memset(Re, 0, sizeof(Re));
    memset(Ke, 0, sizeof(double) * nd * nd);
    for (int g = 0; g < n; ++g) {
      gp_kin k;
      kinematics(&c, corner + e * n * dim, u, g, &k);
      double Fi[MAXD][MAXD], m[MAXN][MAXD], Pm[MAXD][MAXD];
      const double J = inv_small(dim, k.F, Fi);  /* plane strain: F33 = 1 */
      if (!(J > 0.0)) {
        ++failed;
        continue;
      }
      const double lnJ = log(J), mup = mu - lam * lnJ, w = k.wdet;
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) Pm[i][j] = w * (mu * k.F[i][j] - mup * Fi[j][i]);
      for (int a = 0; a < n; ++a)
        for (int i = 0; i < dim; ++i) {
          double s = 0, r = 0;
          for (int j = 0; j < dim; ++j) {
            s += Fi[j][i] * k.gradN[a][j]; /* (F^-T g_a)_i */
            r += Pm[i][j] * k.gradN[a][j];
          }
          m[a][i] = s;
          Re[a * dim + i] += r;
        }
      const double c1 = w * lam, c2 = w * mup, c3 = w * mu;
      for (int a = 0; a < n; ++a)
        for (int b = a; b < n; ++b) {
          double gg = 0;
          for (int i = 0; i < dim; ++i) gg += k.gradN[a][i] * k.gradN[b][i];
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              Ke[(a * dim + i) * nd + b * dim + j] += c1 * m[a][i] * m[b][j] + c2 * m[b][i] * m[a][j] + (i == j ? c3 * gg : 0.0);
        }
    }
    for (int a = 0; a < n; ++a)  /* lower block triangle from the upper one */
      for (int b = 0; b < a; ++b)
        for (int i = 0; i < dim; ++i)
          for (int j = 0; j < dim; ++j) Ke[(a * dim + i) * nd + b * dim + j] = Ke[(b * dim + j) * nd + a * dim + i];
    for (int i = 0; i < nd; ++i) {
#pragma omp atomic
      R[edofs[e * nd + i]] += Re[i];
    }
    int64_t q = 0;
    for (int cc = 0; cc < nd; ++cc)
      for (int r = 0; r < nd; ++r, ++q) {
#pragma omp atomic
        vals[linidx[e * nd * nd + q]] += Ke[r * nd + cc];
      }
  }
  return failed;
}

int ikref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
