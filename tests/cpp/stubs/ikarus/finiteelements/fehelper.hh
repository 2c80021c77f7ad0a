// stand-in for ikarus/finiteelements/fehelper.hh plus the material / finite-element shapes the adapter reads
#pragma once
#include <array>
#include <cstddef>
#include <string>
#include <vector>
namespace Ikarus {
enum class StrainTags { linear, deformationGradient, displacementGradient, greenLagrangian, rightCauchyGreenTensor };
struct LamesFirstParameterAndShearModulus
{
  double lambda, mu;
};
namespace Materials {
  template <typename ST>
  struct LinearElasticityT
  {
    static constexpr auto strainTag = StrainTags::linear;
    static constexpr bool isReduced = false;
    LamesFirstParameterAndShearModulus p;
    static std::string name() { return "LinearElasticity"; }
    const auto& materialParameters() const { return p; }
  };
  template <typename ST>
  struct StVenantKirchhoffT
  {
    static constexpr auto strainTag = StrainTags::greenLagrangian;
    static constexpr bool isReduced = false;
    LamesFirstParameterAndShearModulus p;
    static std::string name() { return "StVenantKirchhoff"; }
    const auto& materialParameters() const { return p; }
  };
  template <typename ST>
  struct NeoHookeT
  {
    static constexpr auto strainTag = StrainTags::rightCauchyGreenTensor;
    static constexpr bool isReduced = false;
    LamesFirstParameterAndShearModulus p;
    static std::string name() { return "NeoHooke"; }
    const auto& materialParameters() const { return p; }
  };
  struct MatrixIndexPair
  {
    std::size_t row, col;
  };
  template <auto pairs, typename MI>
  struct VanishingStrain
  {
    static constexpr auto strainTag = MI::strainTag;
    static constexpr bool isReduced = true;
    MI mat;
    static std::string name() { return "VanishingStrain_" + MI::name(); }
    const auto& materialParameters() const { return mat.materialParameters(); }
  };
  template <auto pairs, typename MI>
  struct VanishingStress
  {
    static constexpr auto strainTag = MI::strainTag;
    static constexpr bool isReduced = true;
    MI mat;
    double tol;
    static std::string name() { return "VanishingStress_" + MI::name(); }
    const auto& materialParameters() const { return mat.materialParameters(); }
  };
}  // namespace Materials
namespace FEHelper {
  /** globalIndices(fe, ids): flat multi-indices of the element's dofs, node-major / component-minor */
  template <typename FE, typename Ids>
  void globalIndices(const FE& fe, Ids& ids) {
    for (auto d : fe.dofs_) ids.push_back({d});
  }
}  // namespace FEHelper
}  // namespace Ikarus
