// SPDX-License-Identifier: LGPL-3.0-or-later
/**
 * \file hosttypes.hh
 * \brief Value types behind DeviceSparseFlatAssembler.  With Ikarus + Eigen on the include path the reference's
 * own types are used; standalone (this repository's image has no DUNE/Eigen) small equivalents with the same
 * member names are provided so the wrapper, its tests and examples compile with g++ alone.
 */
#pragma once

#include <cstdint>
#include <vector>

#if !defined(IKB_HAVE_IKARUS)
  #if __has_include(<ikarus/assembler/interface.hh>) && __has_include(<Eigen/Sparse>)
    #define IKB_HAVE_IKARUS 1
  #else
    #define IKB_HAVE_IKARUS 0
  #endif
#endif

#if IKB_HAVE_IKARUS
  #include <Eigen/Core>
  #include <Eigen/Sparse>

  #include <dune/common/exceptions.hh>
  #include <dune/common/referencehelper.hh>

  #include <ikarus/assembler/dirichletbcenforcement.hh>
  #include <ikarus/finiteelements/fehelper.hh>
  #include <ikarus/finiteelements/ferequirements.hh>
  #include <ikarus/utils/broadcaster/broadcastermessages.hh>
#endif

namespace Ikarus::B200 {

#if IKB_HAVE_IKARUS
struct HostTraits
{
  template <typename FE>
  using Requirement          = typename FE::Requirement;
  using AffordanceCollection = Ikarus::AffordanceCollection<Ikarus::ScalarAffordance, Ikarus::VectorAffordance,
                                                            Ikarus::MatrixAffordance>;
  using ScalarAffordance     = Ikarus::ScalarAffordance;
  using VectorAffordance     = Ikarus::VectorAffordance;
  using MatrixAffordance     = Ikarus::MatrixAffordance;
  using DBCOption            = Ikarus::DBCOption;
  using Vector               = Eigen::VectorXd;
  using SparseMatrix         = Eigen::SparseMatrix<double>;
  /** Eigen keeps column-major compressed storage; pattern and K are symmetric so the CSR arrays of the device
   *  are exactly Eigen's outerIndexPtr/innerIndexPtr (ikarus/assembler/simpleassemblers.inl:206-251). */
  static void setPattern(SparseMatrix& A, std::int64_t rows, const std::vector<std::int64_t>& outer,
                         const std::vector<std::int32_t>& inner) {
    A.resize(rows, rows);
    A.resizeNonZeros(static_cast<Eigen::Index>(inner.size()));
    for (std::int64_t i = 0; i <= rows; ++i)
      A.outerIndexPtr()[i] = static_cast<int>(outer[i]);
    for (std::size_t p = 0; p < inner.size(); ++p)
      A.innerIndexPtr()[p] = inner[p];
  }
  static double* valuePtr(SparseMatrix& A) { return A.valuePtr(); }
};
#else
enum class DBCOption
{
  Raw,
  Reduced,
  Full
};
enum class ScalarAffordance
{
  noAffordance,
  mechanicalPotentialEnergy
};
enum class VectorAffordance
{
  noAffordance,
  forces
};
enum class MatrixAffordance
{
  noAffordance,
  stiffness
};
/** finiteelements/ferequirements.hh:104-169 */
struct AffordanceCollection
{
  ScalarAffordance s{ScalarAffordance::noAffordance};
  VectorAffordance v{VectorAffordance::noAffordance};
  MatrixAffordance m{MatrixAffordance::noAffordance};
  ScalarAffordance scalarAffordance() const { return s; }
  VectorAffordance vectorAffordance() const { return v; }
  MatrixAffordance matrixAffordance() const { return m; }
};
inline constexpr AffordanceCollection elastoStatics{ScalarAffordance::mechanicalPotentialEnergy, VectorAffordance::forces,
                                                    MatrixAffordance::stiffness};

/** finiteelements/ferequirements.hh:222-407 */
struct HostRequirement
{
  std::vector<double> d;
  double lambda{0.0};
  std::vector<double>& globalSolution() { return d; }
  const std::vector<double>& globalSolution() const { return d; }
  double& parameter() { return lambda; }
  const double& parameter() const { return lambda; }
};

/** Compressed column storage with Eigen::SparseMatrix<double>'s accessor names. */
struct HostSparseMatrix
{
  std::int64_t n{0};
  std::vector<std::int64_t> outer;
  std::vector<std::int32_t> inner;
  std::vector<double> values;
  std::int64_t rows() const { return n; }
  std::int64_t cols() const { return n; }
  std::int64_t nonZeros() const { return static_cast<std::int64_t>(values.size()); }
  const double* valuePtr() const { return values.data(); }
  const std::int64_t* outerIndexPtr() const { return outer.data(); }
  const std::int32_t* innerIndexPtr() const { return inner.data(); }
  double coeff(std::int64_t r, std::int64_t c) const {
    for (std::int64_t p = outer[c]; p < outer[c + 1]; ++p)
      if (inner[p] == r)
        return values[p];
    return 0.0;
  }
};

/** One bound finite element as plain data (what the FE walk of the Ikarus adapter extracts). */
struct HostFE
{
  int dim{3}, order{1}, strain{1}, material{2}, easM{0};
  int easFunction{0};  // IKB_EAS_*
  bool planeStrain{false}, planeStress{false};
  double reduceTol{1e-12};
  double lambda{0.0}, mu{0.0};
  ikb_hyperelastic hyperelastic{};  // material == IKB_MAT_HYPERELASTIC: the law (hyperelastic/factory.hh)
  std::vector<std::int64_t> dofs;
  std::vector<double> corners;
};

/** utils/dirichletvalues.hh:73-311 (flags only) */
struct HostDirichletValues
{
  std::vector<bool> flags;
  explicit HostDirichletValues(std::size_t n = 0)
      : flags(n, false) {}
  std::size_t size() const { return flags.size(); }
  bool isConstrained(std::size_t i) const { return flags[i]; }
  void setSingleDOF(std::size_t i, bool f = true) { flags[i] = f; }
  std::size_t fixedDOFsize() const {
    std::size_t c = 0;
    for (bool f : flags)
      c += f;
    return c;
  }
};

struct HostTraits
{
  template <typename FE>
  using Requirement          = HostRequirement;
  using AffordanceCollection = Ikarus::B200::AffordanceCollection;
  using ScalarAffordance     = Ikarus::B200::ScalarAffordance;
  using VectorAffordance     = Ikarus::B200::VectorAffordance;
  using MatrixAffordance     = Ikarus::B200::MatrixAffordance;
  using DBCOption            = Ikarus::B200::DBCOption;
  using Vector               = std::vector<double>;
  using SparseMatrix         = HostSparseMatrix;
  static void setPattern(SparseMatrix& A, std::int64_t rows, const std::vector<std::int64_t>& outer,
                         const std::vector<std::int32_t>& inner) {
    A.n     = rows;
    A.outer = outer;
    A.inner = inner;
    A.values.assign(inner.size(), 0.0);
  }
  static double* valuePtr(SparseMatrix& A) { return A.values.data(); }
};
#endif

} // namespace Ikarus::B200
