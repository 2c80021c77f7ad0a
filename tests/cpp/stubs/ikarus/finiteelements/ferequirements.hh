// stand-in for ikarus/finiteelements/ferequirements.hh: affordances, their collection, FERequirements
#pragma once
#include <Eigen/Core>
namespace Ikarus {
enum class ScalarAffordance { noAffordance, mechanicalPotentialEnergy };
enum class VectorAffordance { noAffordance, forces };
enum class MatrixAffordance { noAffordance, stiffness };
template <typename S, typename V, typename M>
struct AffordanceCollection
{
  S s{};
  V v{};
  M m{};
  S scalarAffordance() const { return s; }
  V vectorAffordance() const { return v; }
  M matrixAffordance() const { return m; }
};
namespace AffordanceCollections {
  inline constexpr AffordanceCollection<ScalarAffordance, VectorAffordance, MatrixAffordance> elastoStatics{
      ScalarAffordance::mechanicalPotentialEnergy, VectorAffordance::forces, MatrixAffordance::stiffness};
}
struct FERequirements
{
  Eigen::VectorXd d;
  double lambda{0.0};
  Eigen::VectorXd& globalSolution() { return d; }
  const Eigen::VectorXd& globalSolution() const { return d; }
  double& parameter() { return lambda; }
  const double& parameter() const { return lambda; }
};
}  // namespace Ikarus
