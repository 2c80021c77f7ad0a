// Standalone C++ test of the header-only drop-in wrapper over the C-ABI.
//   build: g++ -std=c++20 -I include tests/cpp/test_deviceflatassembler.cpp -L ikarus_b200 -likb200 -Wl,-rpath,...
// Mode "compile" (no GPU): checks the error path of ikb_create through the wrapper.
// Mode "run" (GPU): 2x2x2 Hex8 NeoHooke cantilever: bind, matrix/vector/scalar in the three DBC modes, invariants
// of tests/src/testassembler.cpp:122-201 of the reference, Newton with the device PCG callable.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include <ikarus_b200/deviceflatassembler.hh>

#include "flatassembler_concept.hh"

using namespace Ikarus::B200;

// the class models the reference's assembler concept (utils/concepts.hh:517-585), checked at compile time
using TestAssembler = DeviceSparseFlatAssembler<std::vector<HostFE>&, HostDirichletValues>;
static_assert(TestConcepts::MatrixFlatAssembler<TestAssembler, std::vector<double>, DBCOption>);

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("CHECK failed at line %d: %s\n", __LINE__, #cond);       \
      return 1;                                                            \
    }                                                                      \
  } while (0)

static std::vector<HostFE> makeMesh(int nx, int ny, int nz, double h, int material, int strain, int easM) {
  std::vector<HostFE> fes;
  const double E = 1000, nu = 0.3;
  const double lam = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
  auto node = [&](int i, int j, int k) { return (std::int64_t)i + (nx + 1) * ((std::int64_t)j + (ny + 1) * (std::int64_t)k); };
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        HostFE fe;
        fe.material = material;
        fe.strain   = strain;
        fe.easM     = easM;
        fe.lambda   = lam;
        fe.mu       = mu;
        for (int a = 0; a < 8; ++a) {
          const int ii = i + (a & 1), jj = j + ((a >> 1) & 1), kk = k + ((a >> 2) & 1);
          for (int c = 0; c < 3; ++c)
            fe.dofs.push_back(3 * node(ii, jj, kk) + c);
          fe.corners.push_back(ii * h);
          fe.corners.push_back(jj * h);
          fe.corners.push_back(kk * h);
        }
        fes.push_back(fe);
      }
  return fes;
}

int main(int argc, char** argv) {
  const bool run = argc > 1 && std::strcmp(argv[1], "run") == 0;
  {
    // unsupported EAS variant must surface as NotImplemented like in the reference
    std::vector<HostFE> fes = makeMesh(1, 1, 1, 1.0, IKB_MAT_NEOHOOKE, IKB_STRAIN_GREEN_LAGRANGE, 11);
    HostDirichletValues dv(24);
    bool thrown = false;
    try {
      auto a = makeDeviceSparseFlatAssembler(fes, dv);
    } catch (const NotImplemented&) {
      thrown = true;
    } catch (const InvalidState&) {
      thrown = !run;  // without a GPU the CUDA runtime may fail first
    }
    CHECK(thrown);
  }
  if (!run) {
    std::printf("compile-mode ok\n");
    return 0;
  }
  const int nx = 4, ny = 2, nz = 2;
  auto fes = makeMesh(nx, ny, nz, 0.5, IKB_MAT_NEOHOOKE, IKB_STRAIN_GREEN_LAGRANGE, 0);
  const std::size_t n = 3 * (nx + 1) * (ny + 1) * (nz + 1);
  HostDirichletValues dv(n);
  for (int k = 0; k <= nz; ++k)
    for (int j = 0; j <= ny; ++j)
      for (int c = 0; c < 3; ++c)
        dv.setSingleDOF(3 * ((nx + 1) * (j + (ny + 1) * k)) + c);
  auto asmb = makeDeviceSparseFlatAssembler(fes, dv);
  using A   = std::remove_cvref_t<decltype(*asmb)>;
  CHECK(asmb->size() == n && asmb->reducedSize() == n - dv.fixedDOFsize());
  bool unbound = false;
  try {
    asmb->requirement();
  } catch (const InvalidState&) {
    unbound = true;
  }
  CHECK(unbound);

  HostRequirement req;
  req.d.assign(n, 0.0);
  for (std::size_t i = 0; i < n; ++i)
    req.d[i] = dv.isConstrained(i) ? 0.0 : 1e-3 * std::sin(0.37 * i);
  std::vector<double> fext(n, 0.0);
  fext[n - 1] = -1.0;
  asmb->setExternalLoad(fext);
  req.lambda = 0.5;
  asmb->bind(req, elastoStatics, DBCOption::Full);
  CHECK(asmb->bound() && &asmb->requirement() == &req);

  const auto& Kraw  = asmb->matrix(DBCOption::Raw);
  const auto& Kfull = asmb->matrix(DBCOption::Full);
  const auto& Kred  = asmb->matrix(DBCOption::Reduced);
  const auto& Rfull = asmb->vector(DBCOption::Full);
  const auto& Rred  = asmb->vector(DBCOption::Reduced);
  CHECK(Kraw.rows() == (std::int64_t)n && Kred.rows() == (std::int64_t)asmb->reducedSize());
  CHECK(Kraw.nonZeros() == Kfull.nonZeros());
  for (std::size_t i = 0; i < n; ++i) {
    if (!dv.isConstrained(i))
      continue;
    CHECK(Kfull.coeff(i, i) == 1.0 && Rfull[i] == 0.0);
    for (std::int64_t p = Kfull.outer[i]; p < Kfull.outer[i + 1]; ++p)
      CHECK(Kfull.inner[p] == (std::int32_t)i || Kfull.values[p] == 0.0);
  }
  // Reduced == Raw with fixed rows/cols removed (testassembler.cpp:185-201)
  for (std::size_t c = 0; c < n; ++c) {
    if (dv.isConstrained(c))
      continue;
    for (std::int64_t p = Kraw.outer[c]; p < Kraw.outer[c + 1]; ++p) {
      const std::size_t r = Kraw.inner[p];
      if (dv.isConstrained(r))
        continue;
      CHECK(Kred.coeff(r - asmb->constraintsBelow(r), c - asmb->constraintsBelow(c)) == Kraw.values[p]);
    }
  }
  // Sparse == Dense (tests/src/testassembler.cpp:31-38, 122-183)
  {
    const auto& dense = asmb->denseMatrix(req, MatrixAffordance::stiffness, DBCOption::Full);
    CHECK(dense.size() == n * n);
    for (std::size_t c = 0; c < n; ++c)
      for (std::int64_t p = Kfull.outer[c]; p < Kfull.outer[c + 1]; ++p)
        CHECK(dense[c * n + Kfull.inner[p]] == Kfull.values[p]);
    double sumDense = 0, sumSparse = 0;
    for (double v : dense) sumDense += std::fabs(v);
    for (double v : Kfull.values) sumSparse += std::fabs(v);
    CHECK(sumDense == sumSparse);
  }
  auto full = asmb->createFullVector(Rred);
  auto back = asmb->createReducedVector(full);
  CHECK(back == Rred);
  const double E0 = asmb->scalar();
  CHECK(std::isfinite(E0));

  // forces due to inhomogeneous Dirichlet values == K_raw * dInc with constrained entries zeroed (functionhelper.hh:170-185)
  {
    std::vector<double> dInc(n, 0.0);
    for (std::size_t i = 0; i < n; ++i)
      if (dv.isConstrained(i))
        dInc[i] = 0.25 + 1e-3 * double(i % 7);
    const auto F = asmb->forcesDueToIDBC(req, dInc);
    CHECK(F.size() == n);
    double maxErr = 0, maxRef = 0;
    std::vector<double> ref(n, 0.0);
    for (std::size_t c = 0; c < n; ++c)  // Kraw is symmetric: CSC column c == CSR row c
      for (std::int64_t p = Kraw.outer[c]; p < Kraw.outer[c + 1]; ++p)
        ref[c] += Kraw.values[p] * dInc[Kraw.inner[p]];
    for (std::size_t i = 0; i < n; ++i) {
      const double r = dv.isConstrained(i) ? 0.0 : ref[i];
      maxErr         = std::max(maxErr, std::fabs(F[i] - r));
      maxRef         = std::max(maxRef, std::fabs(r));
    }
    CHECK(maxRef > 0 && maxErr <= 1e-12 * maxRef);
  }
  // TrustRegion inner solve on the device: a tiny radius ends on the boundary, a huge one on the kappa rule
  {
    asmb->vector();
    asmb->matrix();
    std::vector<double> minusG(asmb->vector());
    for (double& v : minusG) v = -v;
    ikb_tcg_info ti{};
    ti.kappa = 0.1, ti.theta = 1.0, ti.mininner = 1, ti.precond = IKB_PRECOND_IDENTITY;
    ti.delta = 1e-6;
    auto eta = asmb->truncatedCG(DBCOption::Full, minusG, ti);
    CHECK(ti.stop_reason == IKB_TCG_EXCEEDED_TRUST_REGION && std::fabs(ti.eta_norm - 1e-6) < 1e-15);
    ti.delta = 1e6;
    eta      = asmb->truncatedCG(DBCOption::Full, minusG, ti);
    CHECK(ti.stop_reason == IKB_TCG_REACHED_KAPPA_LINEAR || ti.stop_reason == IKB_TCG_REACHED_THETA_SUPERLINEAR);
    CHECK(ti.iterations > 1 && ti.rel_error <= 0.1 && ti.g_dot_eta < 0 && ti.eta_h_eta > 0);
  }

  // estimateOfConnectivity = 8 x number of grid vertices (assembler/interface.hh:141)
  CHECK(asmb->estimateOfConnectivity() == 8u * (nx + 1) * (ny + 1) * (nz + 1));

  // One Newton iteration = ONE element sweep + ONE gather: residual(x) then jacobian(x) at the same state
  // (newtonraphson.hh:242-243); the second call must be served from the device cache, also when the caller hands in a
  // different requirement object holding the same d.
  {
    auto launches = [&]() {
      std::int64_t n0 = 0;
      ikb_launch_count(asmb->handle(), &n0);
      return n0;
    };
    req.d[n - 1] += 1e-9;  // new state
    const std::int64_t l0 = launches();
    asmb->vector();
    const std::int64_t l1 = launches();
    asmb->matrix();
    const std::int64_t l2 = launches();
    HostRequirement copy = req;
    asmb->matrix(copy, MatrixAffordance::stiffness, DBCOption::Full);
    asmb->vector(copy, VectorAffordance::forces, DBCOption::Full);
    const std::int64_t l3 = launches();
    CHECK(l1 - l0 == 3);  // element kernel, residual gather, matrix gather
    CHECK(l2 == l1 && l3 == l2);
    // with the fused sweep off, vector() alone leaves the matrix gather out
    asmb->setFusedSweep(false);
    req.d[n - 1] += 1e-9;
    asmb->vector();
    const std::int64_t l4 = launches();
    CHECK(l4 - l3 == 2);  // element kernel (R only) + residual gather
    asmb->setFusedSweep(true);
  }
  // vector() -> matrix() -> forcesDueToIDBC() -> solve(): the pushes in between must not invalidate the matrix the
  // PCG is about to use (NewtonRaphson with inhomogeneous Dirichlet values, newtonraphson.hh:214-226)
  {
    const auto& rx = asmb->vector();
    asmb->matrix();
    std::vector<double> dInc(n, 0.0);
    asmb->forcesDueToIDBC(req, dInc);
    int its = 0;
    auto x = asmb->solve(DBCOption::Full, rx, 1e-10, -1, &its);
    CHECK(its > 0 && x.size() == n);
  }
  // The reference's AssemblerManipulator pattern: private base, hooks reached from the derived class, callback
  // mutates the returned vector (the cantilever anchors apply their point load this way, testcantileverbeam.hh:56-80)
  {
    using Manip = TestConcepts::Manipulator<TestAssembler, DBCOption>;
    Manip m(fes, dv);
    m.bind(req, elastoStatics, DBCOption::Full);
    const std::vector<double> plain = m.vector(req, VectorAffordance::forces, DBCOption::Full);
    m.vf = [](const TestAssembler& a, const HostRequirement& r, DBCOption, std::vector<double>& v) { v[a.size() - 1] -= -r.parameter(); };
    const std::vector<double>& loaded = m.vector(req, VectorAffordance::forces, DBCOption::Full);
    CHECK(loaded[n - 1] == plain[n - 1] + req.lambda);
    m.mf = [](const TestAssembler&, const HostRequirement&, DBCOption, HostSparseMatrix& K) { K.values[0] += 1.0; };
    const double k00 = m.matrix(req, MatrixAffordance::stiffness, DBCOption::Raw).values[0];
    m.mf          = nullptr;
    CHECK(k00 == m.matrix(req, MatrixAffordance::stiffness, DBCOption::Raw).values[0] + 1.0);
    CHECK(std::isfinite(m.scalar(req, ScalarAffordance::mechanicalPotentialEnergy)));
  }

  // Principal-stretch framework through the wrapper (HostFE::hyperelastic -> ikb_set_hyperelastic): the NeoHooke recovery
  // of tests/src/testhyperelasticity.hh:139-196 -- Ogden<1, total>({mu}, {2}) with VF3 and Lame's first parameter IS the
  // NeoHooke material, so the generalised-tangent kernel must reproduce the factored NeoHooke kernel: K, R and the energy.
  {
    const double E = 1000, nu = 0.3, lam = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
    auto fo = makeMesh(nx, ny, nz, 0.5, IKB_MAT_HYPERELASTIC, IKB_STRAIN_GREEN_LAGRANGE, 0);
    for (auto& fe : fo) {
      fe.hyperelastic            = ikb_hyperelastic{};
      fe.hyperelastic.deviatoric = IKB_DEV_OGDEN_TOTAL;
      fe.hyperelastic.n          = 1;
      fe.hyperelastic.par[0]     = mu;
      fe.hyperelastic.ex[0]      = 2.0;
      fe.hyperelastic.volumetric = 3;
      fe.hyperelastic.K          = lam;
    }
    auto ao = makeDeviceSparseFlatAssembler(fo, dv);
    ao->setExternalLoad(fext);  // the same load as on the NeoHooke assembler
    const HostSparseMatrix Ko = ao->matrix(req, MatrixAffordance::stiffness, DBCOption::Full);
    const HostSparseMatrix Kn = asmb->matrix(req, MatrixAffordance::stiffness, DBCOption::Full);
    CHECK(Ko.values.size() == Kn.values.size());
    double big = 0, err = 0;
    for (std::size_t p = 0; p < Kn.values.size(); ++p) {
      big = std::max(big, std::abs(Kn.values[p]));
      err = std::max(err, std::abs(Ko.values[p] - Kn.values[p]));
    }
    CHECK(big > 1.0 && err <= 1e-11 * big);
    const std::vector<double> Ro = ao->vector(req, VectorAffordance::forces, DBCOption::Full);
    const std::vector<double> Rn = asmb->vector(req, VectorAffordance::forces, DBCOption::Full);
    double rb = 0, re = 0;
    for (std::size_t i = 0; i < n; ++i) {
      rb = std::max(rb, std::abs(Rn[i]));
      re = std::max(re, std::abs(Ro[i] - Rn[i]));
    }
    CHECK(rb > 0 && re <= 1e-11 * rb);
    const double Eo = ao->scalar(req, ScalarAffordance::mechanicalPotentialEnergy);
    const double En = asmb->scalar(req, ScalarAffordance::mechanicalPotentialEnergy);
    CHECK(std::abs(Eo - En) <= 1e-10 * std::max(1.0, std::abs(En)));
    // a law that was never handed over is a state error, not a silent default
    auto fm = makeMesh(1, 1, 1, 1.0, IKB_MAT_HYPERELASTIC, IKB_STRAIN_GREEN_LAGRANGE, 0);
    HostDirichletValues dv1(24);
    bool thrown = false;
    try {
      auto am = makeDeviceSparseFlatAssembler(fm, dv1);  // deviatoric = none, volumetric = VF0: refused by ikb_set_hyperelastic
    } catch (const InvalidState&) {
      thrown = true;
    } catch (const std::exception&) {
      thrown = true;
    }
    CHECK(thrown);
  }

  // Load skills through the wrapper (loads/volume.hh:67-106, loads/traction.hh:70-138): a volume load and a traction on
  // the face x = nx*h with a non-proportional dependence on the load factor.  At d = 0 the internal forces vanish, so
  // R = -fext(lambda): totals against the exact integrals, nodal values against the wrapper's own sampling.
  {
    auto fes2 = makeMesh(nx, ny, nz, 0.5, IKB_MAT_NEOHOOKE, IKB_STRAIN_GREEN_LAGRANGE, 0);
    auto a2   = makeDeviceSparseFlatAssembler(fes2, dv);
    a2->addVolumeLoad([](const std::array<double, 3>& x, double lam) { return std::array<double, 3>{0.0, 0.0, -(lam * lam + 1.0) * x[0]}; });
    std::vector<std::pair<std::int64_t, int>> faces;
    for (std::int64_t e = 0; e < (std::int64_t)fes2.size(); ++e)
      if (e % nx == nx - 1)
        faces.push_back({e, 1});  // face xi_0 = 1 of the last element column
    a2->addNeumannBoundaryLoad(faces, [](const std::array<double, 3>&, double lam) { return std::array<double, 3>{std::cos(lam), 0.0, 0.0}; });
    HostRequirement r2;
    r2.d.assign(n, 0.0);
    for (double lam : {0.0, 1.5}) {
      r2.lambda = lam;
      const std::vector<double> R = a2->vector(r2, VectorAffordance::forces, DBCOption::Raw);
      const std::vector<double> f = a2->sampleLoads(lam);
      double sx = 0, sz = 0, dmax = 0;
      for (std::size_t i = 0; i < n; ++i) {
        dmax = std::max(dmax, std::abs(R[i] + f[i]));
        if (i % 3 == 0)
          sx += f[i];
        if (i % 3 == 2)
          sz += f[i];
      }
      // box 2 x 1 x 1: int x dV = 2, face area 1
      CHECK(dmax <= 1e-13);
      CHECK(std::abs(sz + (lam * lam + 1.0) * 2.0) <= 1e-12 && std::abs(sx - std::cos(lam)) <= 1e-12);
      const double E = a2->scalar(r2, ScalarAffordance::mechanicalPotentialEnergy);
      CHECK(std::abs(E) <= 1e-13);  // d = 0
    }
  }

  // ResultFunction mirror (io/resultfunction.hh:57-157): PK2 stress of single elements at local positions, served from
  // one device evaluation of all elements per position; a user function (von Mises like) with its own ncomps/name.
  {
    auto rf = makeResultFunction(asmb, IKB_RESULT_PK2_STRESS);
    CHECK(rf->ncomps() == 6 && rf->name() == "PK2Stress");
    const double centre[3] = {0.5, 0.5, 0.5}, corner[3] = {0.0, 1.0, 0.0};
    const std::vector<double> all = asmb->calculateAt(IKB_RESULT_PK2_STRESS, req, centre, 1);
    CHECK(all.size() == fes.size() * 6);
    for (std::int64_t e : {std::int64_t(0), std::int64_t(5), std::int64_t(fes.size() - 1)})
      for (int c = 0; c < 6; ++c)
        CHECK(rf->evaluate(c, e, centre) == all[e * 6 + c]);
    CHECK(rf->evaluate(0, 3, corner) != rf->evaluate(0, 3, centre));
    struct Trace
    {
      double operator()(const double* r, const double*, const HostFE&, int) const { return r[0] + r[1] + r[2]; }
      int ncomps() const { return 1; }
      std::string name() const { return "trace"; }
    };
    auto tr = makeResultFunction(asmb, IKB_RESULT_PK2_STRESS, Trace{});
    CHECK(tr->ncomps() == 1 && tr->name() == "trace");
    CHECK(std::abs(tr->evaluate(0, 2, centre) - (all[12] + all[13] + all[14])) <= 1e-12 * std::abs(all[12]));
    // a new state of the bound requirement drops the cached tables
    const double before = rf->evaluate(0, 1, centre);
    req.d[n - 1] += 1e-3;
    const double after = rf->evaluate(0, (std::int64_t)fes.size() - 1, centre);
    CHECK(after != all[(fes.size() - 1) * 6]);
    req.d[n - 1] -= 1e-3;
    CHECK(rf->evaluate(0, 1, centre) == before);
    // the linear result type is rejected for the nonlinear element like supportsResultType does
    bool rejected = false;
    try {
      makeResultFunction(asmb, IKB_RESULT_LINEAR_STRESS)->evaluate(0, 0, centre);
    } catch (const NotImplemented&) {
      rejected = true;
    }
    CHECK(rejected);
  }

  // The reference's cantilever known answer for EAS::DisplacementGradient with H9 and NeoHooke
  // (tests/src/testcantileverbeamEAS.cpp:81, problem tests/src/testcantileverbeam.hh:83-198: 10 x 1 x 1 Hex8 on
  // 10 x 2 x 2, E = 100, nu = 0.3, x = 0 clamped, unit point loads -lambda e_y at (10,2,2) and (10,2,0), LoadControl with 20
  // steps to lambda = 1, Newton tolerance 1e-10, alpha updated on CORRECTION_UPDATED before the solution): 80 Newton
  // iterations, max |d| = 4.763101490723167 -- through the C++ wrapper, the device PCG as linear solver.
  {
    const double E = 100, nu = 0.3;
    const double lam = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
    std::vector<HostFE> cfes;
    auto cnode = [](int i, int j, int k) { return (std::int64_t)i + 11 * ((std::int64_t)j + 2 * (std::int64_t)k); };
    for (int i = 0; i < 10; ++i) {
      HostFE fe;
      fe.material = IKB_MAT_NEOHOOKE, fe.strain = IKB_STRAIN_GREEN_LAGRANGE, fe.easM = 9;
      fe.easFunction = IKB_EAS_DISPLACEMENT_GRADIENT, fe.lambda = lam, fe.mu = mu;
      for (int a = 0; a < 8; ++a) {
        const int ii = i + (a & 1), jj = (a >> 1) & 1, kk = (a >> 2) & 1;
        for (int c = 0; c < 3; ++c)
          fe.dofs.push_back(3 * cnode(ii, jj, kk) + c);
        fe.corners.push_back(ii * 1.0), fe.corners.push_back(jj * 2.0), fe.corners.push_back(kk * 2.0);
      }
      cfes.push_back(fe);
    }
    const std::size_t cn = 3 * 11 * 2 * 2;
    HostDirichletValues cdv(cn);
    for (int k = 0; k < 2; ++k)
      for (int j = 0; j < 2; ++j)
        for (int c = 0; c < 3; ++c)
          cdv.setSingleDOF(3 * cnode(0, j, k) + c);
    auto ca  = makeDeviceSparseFlatAssembler(cfes, cdv);
    using CA = std::remove_cvref_t<decltype(*ca)>;
    std::vector<double> load(cn, 0.0);
    load[3 * cnode(10, 1, 1) + 1] = -1.0;
    load[3 * cnode(10, 1, 0) + 1] = -1.0;
    ca->setExternalLoad(load);
    HostRequirement creq;
    creq.d.assign(cn, 0.0);
    ca->bind(creq, elastoStatics, DBCOption::Full);
    DevicePCG<CA> cls{ca, 1e-14};
    int total = 0;
    for (int step = 1; step <= 20; ++step) {
      creq.lambda = step / 20.0;
      for (int it = 0;; ++it) {
        CHECK(it < 20);
        const auto& rx = ca->vector();
        const auto& Ax = ca->matrix();
        double nrm   = 0;
        for (double v : rx)
          nrm += v * v;
        if (std::sqrt(nrm) <= 1e-10)
          break;
        auto corr = cls(rx, Ax);
        for (double& v : corr)
          v = -v;
        ca->updateInternalVariables(creq, corr);  // CORRECTION_UPDATED comes before the solution update
        for (std::size_t i = 0; i < cn; ++i)
          creq.d[i] += corr[i];
        ++total;
      }
    }
    double maxd = 0;
    for (double v : creq.d)
      maxd = std::max(maxd, std::abs(v));
    std::printf("cantilever H9 DisplacementGradient through the wrapper: %d Newton iterations, max|d| = %.15f\n", total, maxd);
    CHECK(total == 80);
    CHECK(std::abs(maxd - 4.763101490723167) < 1e-10);
  }

  // Newton iteration with the device PCG callable, as NewtonRaphson::solve does (newtonraphson.hh:196-257)
  DevicePCG<A> ls{asmb, 1e-13};
  double rnorm = 0;
  int iter = 0;
  for (; iter < 20; ++iter) {
    const auto& rx = asmb->vector();
    const auto& Ax = asmb->matrix();
    rnorm          = 0;
    for (double v : rx)
      rnorm += v * v;
    rnorm = std::sqrt(rnorm);
    if (rnorm <= 1e-10)
      break;
    auto corr = ls(rx, Ax);
    CHECK(ls.lastIterations > 0);
    for (std::size_t i = 0; i < n; ++i)
      req.d[i] -= corr[i];
  }
  CHECK(rnorm <= 1e-10 && iter > 1 && iter < 10);
  std::printf("run-mode ok: Newton converged in %d iterations, |R| = %.3e\n", iter, rnorm);
  return 0;
}
