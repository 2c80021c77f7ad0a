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
  struct ArrudaBoyceMatParameters
  {
    double mu, lambdaM;
  };
  struct GentMatParameters
  {
    double mu, Jm;
  };
  template <typename ST>
  struct ArrudaBoyceT
  {
    using ScalarType         = ST;
    using MaterialParameters = ArrudaBoyceMatParameters;
    MaterialParameters p;
    MaterialParameters materialParametersImpl() const { return p; }
    static std::string name() { return "ArrudaBoyce"; }
  };
  template <typename ST>
  struct GentT
  {
    using ScalarType         = ST;
    using MaterialParameters = GentMatParameters;
    MaterialParameters p;
    MaterialParameters materialParametersImpl() const { return p; }
    static std::string name() { return "Gent"; }
  };
  // volumetricfunctions.hh:25-380: plain structs VF0 .. VF12, beta() on VF4 / VF7 / VF10
  struct VF0 { static std::string name() { return "None"; } };
  struct VF1 { static std::string name() { return "Function 1"; } };
  struct VF2 { static std::string name() { return "Function 2"; } };
  struct VF3 { static std::string name() { return "Function 3"; } };
  struct VF4 { double beta_; double beta() const { return beta_; } static std::string name() { return "Function 4"; } };
  struct VF5 { static std::string name() { return "Function 5"; } };
  struct VF6 { static std::string name() { return "Function 6"; } };
  struct VF7 { double beta_; double beta() const { return beta_; } static std::string name() { return "Function 7"; } };
  struct VF8 { static std::string name() { return "Function 8"; } };
  struct VF9 { static std::string name() { return "Function 9"; } };
  struct VF10 { double beta_; double beta() const { return beta_; } static std::string name() { return "Function 10"; } };
  struct VF11 { static std::string name() { return "Function 11"; } };
  struct VF12 { static std::string name() { return "Function 12"; } };
  template <typename VF>
  struct Volumetric
  {
    using VolumetricFunction = VF;
    using MaterialParameter  = double;
    MaterialParameter matPar_{};
    VF volumetricFunction_{};
    const VolumetricFunction& volumetricFunction() const { return volumetricFunction_; }
    const MaterialParameter materialParameter() const { return matPar_; }
    static std::string name() { return "Volumetric function: " + VF::name(); }
  };
  using NoVolumetricPart = Volumetric<VF0>;
  template <typename DEV, typename VOL = NoVolumetricPart>
  struct Hyperelastic
  {
    static constexpr bool hasVolumetricPart = !std::is_same_v<VOL, NoVolumetricPart>;
    static constexpr auto strainTag         = StrainTags::rightCauchyGreenTensor;
    static constexpr bool isReduced         = false;
    using MaterialParameters = typename DEV::MaterialParameters;
    DEV dev_;
    VOL vol_{};
    static std::string name() { return "Hyperelastic (" + DEV::name() + ")"; }
    const MaterialParameters materialParameters() const { return dev_.materialParameters(); }
    const DEV& deviatoricFunction() const { return dev_; }
    const VOL& volumetricFunction() const { return vol_; }
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
    decltype(auto) materialParameters() const { return mat.materialParameters(); }
    auto& underlying() const { return mat; }
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
