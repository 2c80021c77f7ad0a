// The IKB_HAVE_IKARUS branch of include/ikarus_b200/deviceflatassembler.hh, compiled against the stand-in headers of
// tests/cpp/stubs (Eigen, dune-common and Ikarus are not installed in this image) and run on the GPU.
//   build: g++ -std=c++20 -I tests/cpp/stubs -I include tests/cpp/test_ikarus_branch.cpp -L ikarus_b200 -likb200 ...
// Mode "compile" (no GPU): the static_asserts below are the test.  Mode "run": Ikarus-shaped finite elements
// (localView(), material(), gridElement().geometry().corner()) through the adapter, a Newton solve, the manipulator
// pattern and the single CORRECTION_UPDATED listener.
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <vector>

#include <ikarus_b200/deviceflatassembler.hh>

#include "flatassembler_concept.hh"

#if !IKB_HAVE_IKARUS
  #error "this test must see tests/cpp/stubs on the include path"
#endif

using namespace Ikarus::B200;

#define CHECK(cond)                                                  \
  do {                                                               \
    if (!(cond)) {                                                   \
      std::printf("CHECK failed at line %d: %s\n", __LINE__, #cond); \
      return 1;                                                      \
    }                                                                \
  } while (0)

// ---- Ikarus-shaped inputs --------------------------------------------------------------------------------
struct GridViewS
{
  static constexpr int dimension = 3;
  std::size_t nVertices{0};
  std::size_t size(int codim) const { return codim == dimension ? nVertices : 0; }
};
struct BasisS
{
  using GridView = GridViewS;
  GridViewS gv;
  const GridViewS& gridView() const { return gv; }
};
struct DirichletValuesS
{
  using Basis = BasisS;
  BasisS b;
  std::vector<bool> flags;
  std::size_t size() const { return flags.size(); }
  bool isConstrained(std::size_t i) const { return flags[i]; }
  std::size_t fixedDOFsize() const {
    std::size_t c = 0;
    for (bool f : flags) c += f;
    return c;
  }
  const BasisS& basis() const { return b; }
};
struct GeometryS
{
  std::array<std::array<double, 3>, 8> c;
  int corners() const { return 8; }
  const std::array<double, 3>& corner(int i) const { return c[static_cast<std::size_t>(i)]; }
};
struct ElementS
{
  GeometryS g;
  const GeometryS& geometry() const { return g; }
};
template <typename MAT>
struct FES
{
  struct Traits
  {
    static constexpr int mydim = 3, worlddim = 3;
  };
  using Requirement = Ikarus::FERequirements;
  using Material    = MAT;
  using GlobalIndex = std::array<std::size_t, 1>;
  MAT mat;
  ElementS el;
  std::vector<std::size_t> dofs_;
  int order() const { return 1; }
  int localView() const { return 0; }
  const MAT& material() const { return mat; }
  const ElementS& gridElement() const { return el; }
};

using NH   = Ikarus::Materials::NeoHookeT<double>;
using FE   = FES<NH>;
using Asm  = DeviceSparseFlatAssembler<std::vector<FE>&, DirichletValuesS>;
using Vec  = Eigen::VectorXd;
using DBC  = Ikarus::DBCOption;
static_assert(TestConcepts::MatrixFlatAssembler<Asm, Vec, DBC>, "Concepts::MatrixFlatAssembler (utils/concepts.hh:517-585)");
static_assert(std::is_same_v<Asm::VectorType, Eigen::VectorXd> && std::is_same_v<Asm::MatrixType, Eigen::SparseMatrix<double>>);
static_assert(std::is_same_v<Asm::GridView, GridViewS> && std::is_same_v<Asm::Basis, BasisS> &&
              std::is_same_v<Asm::GlobalIndex, std::array<std::size_t, 1>>);  // assembler/interface.hh:32-42
// material law and reduction are read off the TYPE
constexpr std::array<Ikarus::Materials::MatrixIndexPair, 1> onePair{{{2, 2}}};
using PStrain = Ikarus::Materials::VanishingStrain<onePair, Ikarus::Materials::StVenantKirchhoffT<double>>;
using PStress = Ikarus::Materials::VanishingStress<onePair, Ikarus::Materials::LinearElasticityT<double>>;
static_assert(MaterialCode<NH>::material == IKB_MAT_NEOHOOKE && MaterialCode<NH>::reduction == IKB_REDUCE_NONE);
static_assert(MaterialCode<PStrain>::material == IKB_MAT_SVK && MaterialCode<PStrain>::reduction == IKB_REDUCE_PLANE_STRAIN);
static_assert(MaterialCode<PStress>::material == IKB_MAT_LINEAR_ELASTICITY &&
              MaterialCode<PStress>::reduction == IKB_REDUCE_PLANE_STRESS);
using BK        = Ikarus::Materials::Hyperelastic<Ikarus::Materials::Deviatoric<Ikarus::Materials::BlatzKoT<double>>>;
using BKPStrain = Ikarus::Materials::VanishingStrain<onePair, BK>;
static_assert(MaterialCode<BK>::material == IKB_MAT_BLATZKO && MaterialCode<BKPStrain>::material == IKB_MAT_BLATZKO &&
              MaterialCode<BKPStrain>::reduction == IKB_REDUCE_PLANE_STRAIN);  // makeBlatzKo(mu), factory.hh:34-39
using GentVF4 = Ikarus::Materials::Hyperelastic<Ikarus::Materials::Deviatoric<Ikarus::Materials::GentT<double>>,
                                             Ikarus::Materials::Volumetric<Ikarus::Materials::VF4>>;
static_assert(MaterialCode<GentVF4>::material == IKB_MAT_HYPERELASTIC && VolumetricIndex<Ikarus::Materials::VF10>::value == 10);
static_assert(MaterialCode<int>::material < 0);

// a nonlinear solver's broadcaster, as far as subscribeTo() needs it (utils/broadcaster)
struct SolverStateS
{
  const Ikarus::FERequirements& domain;
  const Eigen::VectorXd& correction;
};
struct BroadcasterS
{
  using State = SolverStateS;
  std::vector<std::function<void(Ikarus::NonLinearSolverMessages, const SolverStateS&)>> listeners;
  template <typename M>
  BroadcasterS& station() { return *this; }
  template <typename F>
  int registerListener(F&& f) {
    listeners.emplace_back(std::forward<F>(f));
    return static_cast<int>(listeners.size());
  }
  void notify(Ikarus::NonLinearSolverMessages m, const SolverStateS& s) {
    for (auto& l : listeners) l(m, s);
  }
};

int main(int argc, char** argv) {
  if (!(argc > 1 && std::strcmp(argv[1], "run") == 0)) {
    std::printf("compile-mode ok\n");
    return 0;
  }
  const int nx = 4, ny = 2, nz = 2;
  const double h = 0.5, E = 1000, nu = 0.3;
  const double lam = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
  auto node = [&](int i, int j, int k) { return (std::size_t)i + (nx + 1) * ((std::size_t)j + (ny + 1) * (std::size_t)k); };
  std::vector<FE> fes;
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        FE fe;
        fe.mat.p = {lam, mu};
        for (int a = 0; a < 8; ++a) {
          const int ii = i + (a & 1), jj = j + ((a >> 1) & 1), kk = k + ((a >> 2) & 1);
          for (std::size_t c = 0; c < 3; ++c) fe.dofs_.push_back(3 * node(ii, jj, kk) + c);
          fe.el.g.c[static_cast<std::size_t>(a)] = {ii * h, jj * h, kk * h};
        }
        fes.push_back(fe);
      }
  const std::size_t n = 3 * (nx + 1) * (ny + 1) * (nz + 1);
  DirichletValuesS dv;
  dv.b.gv.nVertices = n / 3;
  dv.flags.assign(n, false);
  for (int k = 0; k <= nz; ++k)
    for (int j = 0; j <= ny; ++j)
      for (std::size_t c = 0; c < 3; ++c) dv.flags[3 * node(0, j, k) + c] = true;
  auto asmb = makeDeviceSparseFlatAssembler(fes, dv);
  CHECK(asmb->size() == n && asmb->gridView().size(3) == n / 3 && asmb->estimateOfConnectivity() == 8 * (n / 3));
  bool unbound = false;
  try {
    asmb->requirement();
  } catch (const Dune::InvalidStateException&) {
    unbound = true;
  }
  CHECK(unbound);

  Ikarus::FERequirements req;
  req.d.resize(static_cast<Eigen::Index>(n));
  for (std::size_t i = 0; i < n; ++i) req.d[(Eigen::Index)i] = dv.flags[i] ? 0.0 : 1e-3 * std::sin(0.37 * double(i));
  Eigen::VectorXd fext(static_cast<Eigen::Index>(n));
  fext.setZero();
  fext[(Eigen::Index)n - 1] = -1.0;
  asmb->setExternalLoad(fext);
  req.lambda = 0.5;
  asmb->bind(req, Ikarus::AffordanceCollections::elastoStatics, DBC::Full);
  CHECK(asmb->bound() && &asmb->requirement() == &req);
  const Eigen::SparseMatrix<double>& K = asmb->matrix();
  const Eigen::VectorXd& R             = asmb->vector();
  CHECK(K.rows() == (Eigen::Index)n && K.nonZeros() > 0 && R.size() == (Eigen::Index)n);
  for (std::size_t i = 0; i < n; ++i)
    if (dv.flags[i]) CHECK(K.coeff((Eigen::Index)i, (Eigen::Index)i) == 1.0 && R[(Eigen::Index)i] == 0.0);
  // symmetric
  for (Eigen::Index c = 0; c < K.cols(); ++c)
    for (int p = K.outerIndexPtr()[c]; p < K.outerIndexPtr()[c + 1]; ++p)
      CHECK(K.valuePtr()[p] == K.coeff(c, K.innerIndexPtr()[p]));
  // the manipulator pattern of the reference
  {
    using Manip = TestConcepts::Manipulator<Asm, DBC>;
    Manip m(fes, dv);
    m.bind(req, Ikarus::AffordanceCollections::elastoStatics, DBC::Full);
    const Eigen::VectorXd plain = m.vector(req, Ikarus::VectorAffordance::forces, DBC::Full);
    m.vf = [](const Asm& a, const Ikarus::FERequirements& r, DBC, Eigen::VectorXd& v) {
      v[(Eigen::Index)a.size() - 2] -= -r.parameter();
    };
    const Eigen::VectorXd& loaded = m.vector(req, Ikarus::VectorAffordance::forces, DBC::Full);
    CHECK(loaded[(Eigen::Index)n - 2] == plain[(Eigen::Index)n - 2] + req.lambda);
  }
  // ONE listener for CORRECTION_UPDATED instead of one per element (controlroutinefactory.hh:43-46): plain elements
  // have no internal variables, the call must go through and change nothing
  {
    BroadcasterS bc;
    asmb->subscribeTo(bc);
    CHECK(bc.listeners.size() == 1);
    Eigen::VectorXd corr(static_cast<Eigen::Index>(n));
    corr.setZero();
    bc.notify(Ikarus::NonLinearSolverMessages::CORRECTION_UPDATED, SolverStateS{req, corr});
  }
  // Newton with the device PCG callable (newtonraphson.hh:196-257)
  using A = std::remove_cvref_t<decltype(*asmb)>;
  DevicePCG<A> ls{asmb, 1e-13};
  double rnorm = 0;
  int iter     = 0;
  for (; iter < 20; ++iter) {
    const auto& rx = asmb->vector();
    const auto& Ax = asmb->matrix();
    rnorm          = 0;
    for (double v : rx) rnorm += v * v;
    rnorm = std::sqrt(rnorm);
    if (rnorm <= 1e-10) break;
    auto corr = ls(rx, Ax);
    for (std::size_t i = 0; i < n; ++i) req.d[(Eigen::Index)i] -= corr[(Eigen::Index)i];
  }
  CHECK(rnorm <= 1e-10 && iter > 1 && iter < 10);

  // A material of the principal-stretch framework read off the TYPE and the public accessors (makeGent({mu, Jm}, K,
  // VF4{beta}), hyperelastic/factory.hh:133-153).  At d = 0 every such law is linear elasticity with shear modulus mu and
  // bulk modulus K U''(1) = K, and so is NeoHooke with lambda = K - 2 mu / 3: the two device kernels must agree there.
  {
    using FEG = FES<GentVF4>;
    const double Kb = 3.0 * lam;
    std::vector<FEG> fg;
    std::vector<FE> fn;
    for (const auto& f : fes) {
      FEG g;
      g.mat.dev_.deviatoricFunction_.p = {mu, 2.5};
      g.mat.vol_.matPar_               = Kb;
      g.mat.vol_.volumetricFunction_   = {0.5};
      g.dofs_                          = f.dofs_;
      g.el                             = f.el;
      fg.push_back(g);
      FE nh = f;
      nh.mat.p = {Kb - 2.0 * mu / 3.0, mu};
      fn.push_back(nh);
    }
    auto ag = makeDeviceSparseFlatAssembler(fg, dv);
    auto an = makeDeviceSparseFlatAssembler(fn, dv);
    Ikarus::FERequirements rq;
    rq.d.resize(static_cast<Eigen::Index>(n));
    rq.d.setZero();
    rq.lambda = 0.0;
    const Eigen::SparseMatrix<double> Kg = ag->matrix(rq, Ikarus::MatrixAffordance::stiffness, DBC::Full);
    const Eigen::SparseMatrix<double> Kn = an->matrix(rq, Ikarus::MatrixAffordance::stiffness, DBC::Full);
    CHECK(Kg.nonZeros() == Kn.nonZeros() && Kg.nonZeros() > 0);
    double big = 0, err = 0;
    for (Eigen::Index p = 0; p < Kg.nonZeros(); ++p) {
      big = std::max(big, std::abs(Kn.valuePtr()[p]));
      err = std::max(err, std::abs(Kg.valuePtr()[p] - Kn.valuePtr()[p]));
    }
    CHECK(big > 1.0 && err <= 1e-11 * big);
    // and away from d = 0 the Gent law differs from NeoHooke
    const Eigen::SparseMatrix<double> Kg2 = ag->matrix(req, Ikarus::MatrixAffordance::stiffness, DBC::Full);
    const Eigen::SparseMatrix<double> Kn2 = an->matrix(req, Ikarus::MatrixAffordance::stiffness, DBC::Full);
    double diff = 0;
    for (Eigen::Index p = 0; p < Kg2.nonZeros(); ++p) diff = std::max(diff, std::abs(Kg2.valuePtr()[p] - Kn2.valuePtr()[p]));
    CHECK(diff > 1e-6 * big);
  }
  std::printf("run-mode ok: Ikarus branch, Newton converged in %d iterations, |R| = %.3e\n", iter, rnorm);
  return 0;
}
