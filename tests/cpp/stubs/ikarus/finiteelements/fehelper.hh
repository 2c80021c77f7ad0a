// stand-in for ikarus/finiteelements/fehelper.hh plus the material / finite-element shapes the adapter reads
#pragma once
#include <array>
#include <cstddef>
#include <string>
#include <type_traits>
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
  // the principal-stretch framework as far as the adapter reads it (materials/hyperelastic/interface.hh:33-97,
  // deviatoric/interface.hh:35-66, deviatoric/blatzko.hh:34-63, volumetric/interface.hh:31, volumetricfunctions.hh:30-45)
  template <typename ST>
  struct BlatzKoT
  {
    using ScalarType         = ST;
    using MaterialParameters = double;
    double mu_;
    MaterialParameters materialParametersImpl() const { return mu_; }
    static std::string name() { return "BlatzKo"; }
  };
  template <typename DF>
  struct Deviatoric
  {
    using ScalarType         = typename DF::ScalarType;
    using DeviatoricFunction = DF;
    using MaterialParameters = typename DF::MaterialParameters;
    DF deviatoricFunction_;
    const MaterialParameters materialParameters() const { return deviatoricFunction_.materialParametersImpl(); }
    static std::string name() { return "Deviatoric function: " + DF::name(); }
  };
  template <typename ST>
  struct VF0T
  {
    static std::string name() { return "None"; }
  };
  template <typename VF>
  struct Volumetric
  {
    using VolumetricFunction = VF;
    using MaterialParameter  = double;
    static std::string name() { return "Volumetric function: " + VF::name(); }
  };
  using NoVolumetricPart = Volumetric<VF0T<double>>;
  template <typename DEV, typename VOL = NoVolumetricPart>
  struct Hyperelastic
  {
    static constexpr bool hasVolumetricPart = !std::is_same_v<VOL, NoVolumetricPart>;
    static constexpr auto strainTag         = StrainTags::rightCauchyGreenTensor;
    static constexpr bool isReduced         = false;
    using MaterialParameters = typename DEV::MaterialParameters;
    DEV dev_;
    static std::string name() { return "Hyperelastic (" + DEV::name() + ")"; }
    const MaterialParameters materialParameters() const { return dev_.materialParameters(); }
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
