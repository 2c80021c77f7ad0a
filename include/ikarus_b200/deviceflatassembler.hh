// SPDX-License-Identifier: LGPL-3.0-or-later
/**
 * \file deviceflatassembler.hh
 * \brief Header-only C++20 host side of the B200 device flat assembler.
 *
 * DeviceSparseFlatAssembler<FEC, DV> models the reference's Concepts::MatrixFlatAssembler
 * (ikarus/utils/concepts.hh:517-585) with the semantics of FlatAssemblerBase / ScalarAssembler /
 * VectorAssembler / MatrixAssembler (ikarus/assembler/interface.hh:28-468) and SparseFlatAssembler
 * (ikarus/assembler/simpleassemblers.hh:106-177), but every get*Impl is one call into libikb200.so
 * (include/ikb200.h) instead of a serial loop over finite elements.
 *
 * Two build modes:
 *   - with Ikarus/DUNE/Eigen on the include path (IKB_HAVE_IKARUS, auto-detected): VectorType is
 *     Eigen::VectorXd, MatrixType is Eigen::SparseMatrix<double> ("mirror mode": the device writes the
 *     values in Eigen's compressed order and they are copied into valuePtr()), exceptions are Dune's,
 *     and the element walk uses FEHelper::globalIndices / geometry().corner().  NewtonRaphson,
 *     TrustRegion and LoadControl then drive this class unchanged.
 *   - standalone (this repository, no DUNE in the image): the same class over the small value types of
 *     ikarus_b200/hosttypes.hh, used by tests/cpp.
 * The element container is adapted through ElementAccess<FE> (customisation point below).
 */
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <functional>
#include <iterator>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "../ikb200.h"
#include "hosttypes.hh"

// DeviceResultFunction derives from Dune::VTKFunction (as io/resultfunction.hh:57 does) where dune-grid is present
#if !defined(IKB_HAVE_DUNE_VTK)
  #if IKB_HAVE_IKARUS && defined(__has_include)
    #if __has_include(<dune/grid/io/file/vtk/function.hh>)
      #define IKB_HAVE_DUNE_VTK 1
    #endif
  #endif
  #if !defined(IKB_HAVE_DUNE_VTK)
    #define IKB_HAVE_DUNE_VTK 0
  #endif
#endif
#if IKB_HAVE_DUNE_VTK
  #include <dune/common/fvector.hh>
  #include <dune/grid/io/file/vtk/function.hh>
#endif

namespace Ikarus::B200 {

/** Thrown where the reference throws Dune::InvalidStateException / Dune::NotImplemented. */
#if IKB_HAVE_IKARUS
using InvalidState   = Dune::InvalidStateException;
using NotImplemented = Dune::NotImplemented;
  #define IKB_THROW(Type, msg) DUNE_THROW(Type, msg)
#else
struct InvalidState : std::logic_error
{
  using std::logic_error::logic_error;
};
struct NotImplemented : std::logic_error
{
  using std::logic_error::logic_error;
};
  #define IKB_THROW(Type, msg) throw Type(std::string(msg))
#endif
/** The reference aborts the process on det C <= 0 (materials/materialhelpers.hh:120-126); here it is an exception. */
struct MaterialFailure : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

/**
 * \brief Customisation point: how to read one finite element.  Specialise for your FE type, or provide the
 * members used by the primary template (the standalone HostFE of hosttypes.hh does).
 */
template <typename FE>
struct ElementAccess
{
  static int dim(const FE& fe) { return fe.dim; }
  static int order(const FE& fe) { return fe.order; }
  static int strain(const FE& fe) { return fe.strain; }      // IKB_STRAIN_*
  static int material(const FE& fe) { return fe.material; }  // IKB_MAT_*
  static bool planeStrain(const FE& fe) { return fe.planeStrain; }
  /** IKB_REDUCE_*: planeStrain / planeStress wrapper of the material; tolerance of the planeStress reduction */
  static int reduction(const FE& fe) { return fe.planeStress ? IKB_REDUCE_PLANE_STRESS : (fe.planeStrain ? IKB_REDUCE_PLANE_STRAIN : IKB_REDUCE_NONE); }
  static double reductionTolerance(const FE& fe) { return fe.reduceTol; }
  static double lambda(const FE& fe) { return fe.lambda; }
  static double mu(const FE& fe) { return fe.mu; }
  /** IKB_MAT_HYPERELASTIC: the law of the principal-stretch framework; false for the other materials */
  static bool hyperelastic(const FE& fe, ikb_hyperelastic& law) {
    if (fe.material != IKB_MAT_HYPERELASTIC)
      return false;
    law = fe.hyperelastic;
    return true;
  }
  static int numberOfInternalVariables(const FE& fe) { return fe.easM; }
  /** IKB_EAS_*: the ES of eas<ES>(m) (mechanics/enhancedassumedstrains.hh:69) */
  static int easFunction(const FE& fe) { return fe.easFunction; }
  /** FEHelper::globalIndices(fe, dofs) (finiteelements/fehelper.hh:194-197), flat index [0] */
  static void globalIndices(const FE& fe, std::vector<std::int64_t>& dofs) {
    dofs.insert(dofs.end(), fe.dofs.begin(), fe.dofs.end());
  }
  /** fe.gridElement().geometry().corner(c)[k], c-major */
  static void corners(const FE& fe, std::vector<double>& x) { x.insert(x.end(), fe.corners.begin(), fe.corners.end()); }
};

#if IKB_HAVE_IKARUS
/** Which device material law and reduction an Ikarus material type maps to, decided on the TYPE (not on name()). */
template <typename M>
struct MaterialCode
{
  static constexpr int material  = -1;  // outside the device hot path
  static constexpr int reduction = IKB_REDUCE_NONE;
};
template <typename ST>
struct MaterialCode<Ikarus::Materials::LinearElasticityT<ST>>
{
  static constexpr int material = IKB_MAT_LINEAR_ELASTICITY, reduction = IKB_REDUCE_NONE;
};
template <typename ST>
struct MaterialCode<Ikarus::Materials::StVenantKirchhoffT<ST>>
{
  static constexpr int material = IKB_MAT_SVK, reduction = IKB_REDUCE_NONE;
};
template <typename ST>
struct MaterialCode<Ikarus::Materials::NeoHookeT<ST>>
{
  static constexpr int material = IKB_MAT_NEOHOOKE, reduction = IKB_REDUCE_NONE;
};
/** makeBlatzKo(mu) = Hyperelastic<Deviatoric<BlatzKoT<ST>>> without a volumetric part (materials/hyperelastic/
 *  factory.hh:34-39, interface.hh:33-41); its material parameter is the scalar mu (deviatoric/blatzko.hh:49-63) */
template <typename ST, typename VOL>
requires(!Ikarus::Materials::Hyperelastic<Ikarus::Materials::Deviatoric<Ikarus::Materials::BlatzKoT<ST>>, VOL>::hasVolumetricPart)
struct MaterialCode<Ikarus::Materials::Hyperelastic<Ikarus::Materials::Deviatoric<Ikarus::Materials::BlatzKoT<ST>>, VOL>>
{
  static constexpr int material = IKB_MAT_BLATZKO, reduction = IKB_REDUCE_NONE;
};
/** Any other Hyperelastic<Deviatoric<DF>, Volumetric<VF>> (hyperelastic/interface.hh:33-97) goes through
 *  ikb_set_hyperelastic; HyperelasticLaw below reads what the reference's public accessors expose. */
template <typename DEV, typename VOL>
struct MaterialCode<Ikarus::Materials::Hyperelastic<DEV, VOL>>
{
  static constexpr int material = IKB_MAT_HYPERELASTIC, reduction = IKB_REDUCE_NONE;
};
/** VF0 .. VF12 -> index (volumetric/volumetricfunctions.hh:25-380); beta() of VF4 / VF7 / VF10 */
template <typename VF>
struct VolumetricIndex
{
  static constexpr int value = -1;
};
  #define IKB_VF(N)                                   \
    template <>                                       \
    struct VolumetricIndex<Ikarus::Materials::VF##N>  \
    {                                                 \
      static constexpr int value = N;                 \
    };
IKB_VF(0) IKB_VF(1) IKB_VF(2) IKB_VF(3) IKB_VF(4) IKB_VF(5) IKB_VF(6) IKB_VF(7) IKB_VF(8) IKB_VF(9) IKB_VF(10) IKB_VF(11) IKB_VF(12)
  #undef IKB_VF
/** Fills ikb_hyperelastic from a material object.  Deviatoric<DF> exposes only DF::materialParametersImpl()
 *  (deviatoric/interface.hh:66): enough for BlatzKo {mu}, ArrudaBoyce {mu, lambdaM} and Gent {mu, Jm}.  The exponents of
 *  Ogden and InvariantBased are not reachable through the reference's public interface -- specialise
 *  ElementAccess<FE>::hyperelastic for those (the factories' arguments are at hand where the material is built). */
template <typename M>
struct HyperelasticLaw
{
  static bool fill(const M&, ikb_hyperelastic&) { return false; }
};
template <typename DEV, typename VOL>
struct HyperelasticLaw<Ikarus::Materials::Hyperelastic<DEV, VOL>>
{
  using Mat = Ikarus::Materials::Hyperelastic<DEV, VOL>;
  using DF  = typename DEV::DeviatoricFunction;
  static bool fill(const Mat& mat, ikb_hyperelastic& law) {
    law = ikb_hyperelastic{};
    const auto p = mat.deviatoricFunction().materialParameters();
    if constexpr (requires { p.lambdaM; }) {
      law.deviatoric = IKB_DEV_ARRUDA_BOYCE, law.par[0] = p.mu, law.par[1] = p.lambdaM;
    } else if constexpr (requires { p.Jm; }) {
      law.deviatoric = IKB_DEV_GENT, law.par[0] = p.mu, law.par[1] = p.Jm;
    } else if constexpr (std::is_arithmetic_v<std::remove_cvref_t<decltype(p)>>) {
      law.deviatoric = IKB_DEV_BLATZKO, law.par[0] = static_cast<double>(p);
    } else {
      // Ogden / InvariantBased: the exponents are private in the reference; the caller hands the law over with
      // DeviceSparseFlatAssembler::setHyperelasticLaw() before the first assembly (IKB_ESTATE otherwise)
      return false;
    }
    if constexpr (Mat::hasVolumetricPart) {
      using VF = typename VOL::VolumetricFunction;
      static_assert(VolumetricIndex<VF>::value >= 0, "unknown volumetric function");
      law.volumetric = VolumetricIndex<VF>::value;
      law.K          = mat.volumetricFunction().materialParameter();
      if constexpr (requires { mat.volumetricFunction().volumetricFunction().beta(); })
        law.beta = mat.volumetricFunction().volumetricFunction().beta();
    }
    return true;
  }
};
template <auto pairs, typename MI>
struct HyperelasticLaw<Ikarus::Materials::VanishingStrain<pairs, MI>>
{
  static bool fill(const Ikarus::Materials::VanishingStrain<pairs, MI>& mat, ikb_hyperelastic& law) {
    return HyperelasticLaw<MI>::fill(mat.underlying(), law);
  }
};
/** planeStrain(mat) (materials/vanishingstrain.hh:186-198) */
template <auto pairs, typename MI>
struct MaterialCode<Ikarus::Materials::VanishingStrain<pairs, MI>>
{
  static constexpr int material = MaterialCode<MI>::material, reduction = IKB_REDUCE_PLANE_STRAIN;
};
/** planeStress(mat, tol) (materials/vanishingstress.hh:254-265) */
template <auto pairs, typename MI>
struct MaterialCode<Ikarus::Materials::VanishingStress<pairs, MI>>
{
  static constexpr int material = MaterialCode<MI>::material, reduction = IKB_REDUCE_PLANE_STRESS;
};

/** Adapter for real Ikarus finite elements FE<PreFE, Skills...> (finiteelements/mixin.hh, febase.hh). */
template <typename FE>
requires requires(const FE& fe) { fe.localView(); fe.material(); }
struct ElementAccess<FE>
{
  using Mat = std::remove_cvref_t<decltype(std::declval<const FE&>().material())>;
  static int dim(const FE&) { return FE::Traits::mydim; }
  static int order(const FE& fe) { return fe.order(); }
  static int strain(const FE&) {
    return Mat::strainTag == Ikarus::StrainTags::linear ? IKB_STRAIN_LINEAR : IKB_STRAIN_GREEN_LAGRANGE;
  }
  static int material(const FE& fe) {
    if constexpr (MaterialCode<Mat>::material < 0)
      IKB_THROW(NotImplemented, "material " + fe.material().name() + " is outside the device hot path");
    return MaterialCode<Mat>::material;
  }
  static bool planeStrain(const FE&) { return MaterialCode<Mat>::reduction == IKB_REDUCE_PLANE_STRAIN; }
  static int reduction(const FE&) { return MaterialCode<Mat>::reduction; }
  /** VanishingStress keeps its tolerance private (vanishingstress.hh:231); planeStress() defaults to 1e-8
   *  (vanishingstress.hh:254-257).  Specialise ElementAccess to pass another one. */
  static double reductionTolerance(const FE&) { return 1e-8; }
  /** Lame parameters of the two-parameter laws; a principal-stretch law with one scalar parameter passes it as mu */
  static double lambda(const FE& fe) {
    if constexpr (requires { fe.material().materialParameters().lambda; })
      return fe.material().materialParameters().lambda;
    else
      return 0.0;
  }
  static double mu(const FE& fe) {
    if constexpr (requires { fe.material().materialParameters().mu; })
      return fe.material().materialParameters().mu;
    else
      return static_cast<double>(fe.material().materialParameters());
  }
  static bool hyperelastic(const FE& fe, ikb_hyperelastic& law) {
    if constexpr (MaterialCode<Mat>::material == IKB_MAT_HYPERELASTIC)
      return HyperelasticLaw<Mat>::fill(fe.material(), law);
    else
      return false;
  }
  static int numberOfInternalVariables(const FE& fe) {
    if constexpr (requires { fe.numberOfInternalVariables(); })
      return fe.numberOfInternalVariables();
    else
      return 0;
  }
  static int easFunction(const FE&) {
    if constexpr (requires { typename FE::EnhancedStrainFunction; }) {
      using ES = typename FE::EnhancedStrainFunction;
      if constexpr (requires { ES::name(); }) {
        // EAS::DisplacementGradient / DisplacementGradientTransposed (strainenhancements/easfunctions/
        // displacementgradient.hh:256, displacementgradienttransposed.hh:322)
        constexpr std::string_view n = "Displacement Gradient (Transposed)";
        const std::string name = ES::name();
        if (name == n)
          return IKB_EAS_DISPLACEMENT_GRADIENT_TRANSPOSED;
        if (name == n.substr(0, 21))
          return IKB_EAS_DISPLACEMENT_GRADIENT;
      }
    }
    return IKB_EAS_STRAIN;
  }
  static void globalIndices(const FE& fe, std::vector<std::int64_t>& dofs) {
    std::vector<typename FE::GlobalIndex> ids;
    Ikarus::FEHelper::globalIndices(fe, ids);
    for (auto& id : ids)
      dofs.push_back(static_cast<std::int64_t>(id[0]));
  }
  static void corners(const FE& fe, std::vector<double>& x) {
    const auto geo = fe.gridElement().geometry();
    for (int c = 0; c < geo.corners(); ++c)
      for (int k = 0; k < FE::Traits::worlddim; ++k)
        x.push_back(geo.corner(c)[k]);
  }
};
#endif

/**
 * \brief Device matrix handle: what DeviceSparseFlatAssembler::deviceMatrix() returns; the JacobianType when
 * NewtonRaphson runs in resident mode (matrix() may return any type, utils/concepts.hh:575-585).
 */
struct DeviceMatrixHandle
{
  ikb_handle handle{nullptr};
  int dbc{IKB_DBC_FULL};
  std::int64_t rows{0};
};

template <typename FEC, typename DV>
class DeviceSparseFlatAssembler
{
public:
  using FEContainer              = FEC;
  using FE                       = std::remove_cvref_t<decltype(*std::begin(std::declval<FEC&>()))>;
  using FERequirement            = HostTraits::template Requirement<FE>;
  using DirichletValuesType      = DV;
  using SizeType                 = std::size_t;
  using AffordanceCollectionType = HostTraits::AffordanceCollection;
  using ScalarType               = double;
  using VectorType               = HostTraits::Vector;
  using MatrixType               = HostTraits::SparseMatrix;
  using DBCOption                = HostTraits::DBCOption;
#if IKB_HAVE_IKARUS
  // assembler/interface.hh:32-42
  using GlobalIndex = typename FE::GlobalIndex;
  using Basis       = typename DV::Basis;
  using GridView    = typename Basis::GridView;
#endif

  /** FlatAssemblerBase ctor (assembler/interface.hh:51-62) + one-time device upload. */
  DeviceSparseFlatAssembler(FEC&& fes, const DV& dirichletValues, int device = -1)
      : fes_{std::forward<FEC>(fes)},
        dirichletValues_{dirichletValues} {
    const std::size_t n = dirichletValues_.size();
    constraintsBelow_.reserve(n);
    std::size_t counter = 0;
    std::vector<std::uint8_t> flags(n);
    for (std::size_t i = 0; i < n; ++i) {
      constraintsBelow_.push_back(counter);
      if (dirichletValues_.isConstrained(i)) {
        ++counter;
        flags[i] = 1;
      }
    }
    fixedDofs_ = dirichletValues_.fixedDOFsize();

    // single walk over the container: connectivity + corner coordinates
    std::vector<std::int64_t> dofs;
    std::vector<double> corners;
    std::int64_t nElem = 0;
    ikb_desc desc{};
    desc.abi_version = IKB_ABI_VERSION;
    ikb_hyperelastic law{};
    bool hasLaw = false;
    for (const auto& fe : fes_) {
      using A = ElementAccess<FE>;
      if (nElem == 0) {
        desc.dim          = A::dim(fe);
        desc.order        = A::order(fe);
        desc.strain       = A::strain(fe);
        desc.material     = A::material(fe);
        desc.plane_strain = A::reduction(fe);
        desc.reduce_tol   = A::reductionTolerance(fe);
        desc.eas_m        = A::numberOfInternalVariables(fe);
        desc.eas_function = A::easFunction(fe);
        desc.lambda       = A::lambda(fe);
        desc.mu           = A::mu(fe);
        hasLaw            = A::hyperelastic(fe, law);
      }
      A::globalIndices(fe, dofs);
      A::corners(fe, corners);
      ++nElem;
    }
    // number of grid vertices for estimateOfConnectivity(): the distinct nodes at element corners (for Q2 the corner
    // nodes are the lattice positions with every coordinate in {0, 2}; local order is lexicographic, x fastest)
    if (nElem > 0) {
      const int dim = desc.dim, n1 = desc.order + 1;
      int nn = 1;
      for (int k = 0; k < dim; ++k)
        nn *= n1;
      std::vector<std::uint8_t> seen(n / static_cast<std::size_t>(dim) + 1, 0);
      std::vector<std::int64_t> minDof(static_cast<std::size_t>(nn));
      for (std::int64_t e = 0; e < nElem; ++e) {
        for (int a = 0; a < nn; ++a) {
          bool corner = true;
          for (int k = 0, q = a; k < dim; ++k, q /= n1)
            corner = corner && (q % n1 == 0 || q % n1 == desc.order);
          if (!corner)
            continue;
          const std::int64_t d0 = dofs[static_cast<std::size_t>(e) * nn * dim + static_cast<std::size_t>(a) * dim];
          // FlatInterleaved: dim*node + 0, FlatLexicographic: node
          const std::size_t node = static_cast<std::size_t>(dofs[1] == dofs[0] + 1 ? d0 / dim : d0);
          if (node < seen.size() && !seen[node]) {
            seen[node] = 1;
            ++vertexCount_;
          }
        }
      }
    }
    desc.device = device;
    desc.n_elem = nElem;
    desc.n_dof  = static_cast<std::int64_t>(n);
    const int rc = ikb_create(&h_, &desc);
    if (rc == IKB_ENOTIMPL)
      IKB_THROW(NotImplemented, "EAS is only supported for Q1, Q2 and H1 elements");
    if (rc != IKB_OK)
      IKB_THROW(InvalidState, "ikb_create failed (" + std::to_string(rc) + "): unsupported element description");
    if (hasLaw)
      check(ikb_set_hyperelastic(h_, &law));
    check(ikb_upload_mesh(h_, corners.data(), dofs.data()));
    check(ikb_upload_dirichlet(h_, flags.data()));
    check(ikb_build_pattern(h_));
    dim_     = desc.dim;
    order_   = desc.order;
    nElem_   = nElem;
    corners_ = std::move(corners);  // kept for the host-side load sampling
    dofs_    = std::move(dofs);
  }
  /** The law of a principal-stretch material (IKB_MAT_HYPERELASTIC) where it cannot be read off the material object:
   *  Deviatoric<DF> exposes only materialParameters() (deviatoric/interface.hh:61-66), so the exponents of makeOgden /
   *  makeMooneyRivlin / makeYeoh / makeInvariantBased are handed over here, before the first assembly. */
  void setHyperelasticLaw(const ikb_hyperelastic& law) { check(ikb_set_hyperelastic(h_, &law)); }
  ~DeviceSparseFlatAssembler() {
    if (h_)
      ikb_destroy(h_);
  }
  DeviceSparseFlatAssembler(const DeviceSparseFlatAssembler&)            = delete;
  DeviceSparseFlatAssembler& operator=(const DeviceSparseFlatAssembler&) = delete;

  // ---- FlatAssemblerBase (assembler/interface.hh:68-132) ------------------------------------------
  std::size_t size() const { return dirichletValues_.size(); }
  std::size_t reducedSize() const { return size() - fixedDofs_; }
  auto& finiteElements() const { return fes_; }
  const auto& dirichletValues() const { return dirichletValues_; }
  std::size_t constraintsBelow(std::size_t i) const { return constraintsBelow_[i]; }
  bool isConstrained(std::size_t i) const { return dirichletValues_.isConstrained(i); }
  /** assembler/interface.hh:141: gridView.size(dim) * 8, i.e. eight times the number of grid vertices */
  std::size_t estimateOfConnectivity() const { return vertexCount_ * 8; }
#if IKB_HAVE_IKARUS
  /** assembler/interface.hh:110-116 */
  const auto& basis() const { return dirichletValues_.basis(); }
  const auto& gridView() const { return Dune::resolveRef(dirichletValues_.basis().gridView()); }
#endif

  VectorType createFullVector(const VectorType& reducedVector) const {
    assert(static_cast<std::size_t>(reducedVector.size()) == reducedSize() &&
           "The reduced vector you passed has the wrong dimensions.");
    VectorType full(size());
    std::size_t reducedCounter = 0;
    for (std::size_t i = 0; i < size(); ++i) {
      if (isConstrained(i)) {
        ++reducedCounter;
        full[i] = 0.0;
      } else
        full[i] = reducedVector[i - reducedCounter];
    }
    return full;
  }
  VectorType createReducedVector(const VectorType& fullVector) const {
    assert(static_cast<std::size_t>(fullVector.size()) == size() &&
           "The full vector you passed has the wrong dimensions.");
    VectorType red(reducedSize());
    std::size_t reducedCounter = 0;
    for (std::size_t i = 0; i < size(); ++i) {
      if (isConstrained(i))
        ++reducedCounter;
      else
        red[i - reducedCounter] = fullVector[i];
    }
    return red;
  }

  // ---- binding (assembler/interface.hh:150-256) ----------------------------------------------------
  void bind(const FERequirement& req, AffordanceCollectionType aff, DBCOption dbc = DBCOption::Full) {
    req_ = std::cref(req);
    aff_ = aff;
    dbc_ = dbc;
  }
  void bind(const FERequirement& req) { req_ = std::cref(req); }
  void bind(AffordanceCollectionType aff) { aff_ = aff; }
  void bind(DBCOption dbc) { dbc_ = dbc; }
  bool bound() const { return boundToRequirement() and boundToAffordanceCollection() and boundToDBCOption(); }
  bool boundToRequirement() const { return req_.has_value(); }
  bool boundToAffordanceCollection() const { return aff_.has_value(); }
  bool boundToDBCOption() const { return dbc_.has_value(); }
  const FERequirement& requirement() const {
    if (req_.has_value())
      return req_.value().get();
    IKB_THROW(InvalidState, "The requirement can only be obtained after binding");
  }
  AffordanceCollectionType affordanceCollection() const {
    if (aff_.has_value())
      return aff_.value();
    IKB_THROW(InvalidState, "The affordance can only be obtained after binding");
  }
  DBCOption dBCOption() const {
    if (dbc_.has_value())
      return dbc_.value();
    IKB_THROW(InvalidState, "The dBCOption can only be obtained after binding");
  }

  // ---- ScalarAssembler / VectorAssembler / MatrixAssembler (interface.hh:300-462) -----------------
  // Public calls dispatch on the DBCOption to the get*Impl hooks exactly like the reference's CRTP interfaces
  // (interface.hh:355-389, 430-462); the hooks are protected so that the reference's AssemblerManipulator
  // (assemblermanipulatorfuser.hh:242-385, which derives privately from the wrapped assembler and calls
  // A::get*Impl) can wrap this class and run its callbacks on the returned quantity.
  const ScalarType& scalar(const FERequirement& req, HostTraits::ScalarAffordance aff) { return getScalarImpl(req, aff); }
  const ScalarType& scalar() { return scalar(requirement(), affordanceCollection().scalarAffordance()); }

  const VectorType& vector(const FERequirement& req, HostTraits::VectorAffordance aff,
                           DBCOption dbc = DBCOption::Full) {
    if (dbc == DBCOption::Raw)
      return getRawVectorImpl(req, aff);
    if (dbc == DBCOption::Reduced)
      return getReducedVectorImpl(req, aff);
    return getVectorImpl(req, aff);
  }
  const VectorType& vector(DBCOption dbc) { return vector(requirement(), affordanceCollection().vectorAffordance(), dbc); }
  const VectorType& vector() { return vector(dBCOption()); }

  const MatrixType& matrix(const FERequirement& req, HostTraits::MatrixAffordance aff,
                           DBCOption dbc = DBCOption::Full) {
    if (dbc == DBCOption::Raw)
      return getRawMatrixImpl(req, aff);
    if (dbc == DBCOption::Reduced)
      return getReducedMatrixImpl(req, aff);
    return getMatrixImpl(req, aff);
  }
  const MatrixType& matrix(DBCOption dbc) { return matrix(requirement(), affordanceCollection().matrixAffordance(), dbc); }
  const MatrixType& matrix() { return matrix(dBCOption()); }

  /**
   * One element sweep yields K and R together.  With the fused sweep on (default) vector() also asks for the matrix
   * whenever a stiffness affordance is bound (or nothing is bound), so that NewtonRaphson's `residual(x); jacobian(x)`
   * pair (solver/nonlinearsolver/newtonraphson.hh:204-205, 242-243) costs ONE element kernel + ONE gather; the matrix
   * call that follows is served from the device cache because push() leaves an unchanged state alone.  Switch it off
   * where gradients are evaluated without the Hessian (TrustRegion's rejected steps, trustregion.hh:309-340).
   */
  void setFusedSweep(bool on) { fusedSweep_ = on; }
  bool fusedSweep() const { return fusedSweep_; }

protected:
  ScalarType& getScalarImpl(const FERequirement& req, HostTraits::ScalarAffordance aff) {
    if (aff != HostTraits::ScalarAffordance::mechanicalPotentialEnergy)
      IKB_THROW(NotImplemented, "ScalarAffordance not implemented");
    push(req);
    check(ikb_assemble(h_, IKB_SCALAR, IKB_DBC_RAW));
    check(ikb_get_scalar(h_, &scal_));
    return scal_;
  }
  VectorType& getRawVectorImpl(const FERequirement& req, HostTraits::VectorAffordance aff) {
    return vectorImpl(req, aff, DBCOption::Raw);
  }
  VectorType& getVectorImpl(const FERequirement& req, HostTraits::VectorAffordance aff) {
    return vectorImpl(req, aff, DBCOption::Full);
  }
  VectorType& getReducedVectorImpl(const FERequirement& req, HostTraits::VectorAffordance aff) {
    return vectorImpl(req, aff, DBCOption::Reduced);
  }
  MatrixType& getRawMatrixImpl(const FERequirement& req, HostTraits::MatrixAffordance aff) {
    return matrixImpl(req, aff, DBCOption::Raw);
  }
  MatrixType& getMatrixImpl(const FERequirement& req, HostTraits::MatrixAffordance aff) {
    return matrixImpl(req, aff, DBCOption::Full);
  }
  MatrixType& getReducedMatrixImpl(const FERequirement& req, HostTraits::MatrixAffordance aff) {
    return matrixImpl(req, aff, DBCOption::Reduced);
  }

private:
  bool matrixLikelyNext() const {
    return fusedSweep_ && (!aff_.has_value() || aff_->matrixAffordance() == HostTraits::MatrixAffordance::stiffness);
  }
  VectorType& vectorImpl(const FERequirement& req, HostTraits::VectorAffordance aff, DBCOption dbc) {
    if (aff != HostTraits::VectorAffordance::forces)
      IKB_THROW(NotImplemented, "VectorAffordance not implemented");
    push(req);
    const int d = toCode(dbc);
    check(ikb_assemble(h_, IKB_VECTOR | (matrixLikelyNext() ? IKB_MATRIX : 0u), d));
    VectorType& out = vec_[d];
    out.resize(dbc == DBCOption::Reduced ? reducedSize() : size());
    check(ikb_get_vector(h_, d, out.data()));
    return out;
  }
  MatrixType& matrixImpl(const FERequirement& req, HostTraits::MatrixAffordance aff, DBCOption dbc) {
    if (aff != HostTraits::MatrixAffordance::stiffness)
      IKB_THROW(NotImplemented, "MatrixAffordance not implemented");
    push(req);
    const int d = toCode(dbc);
    check(ikb_assemble(h_, IKB_MATRIX | (fusedSweep_ ? IKB_VECTOR : 0u), d));
    MatrixType& A = mat_[d];
    if (not patternReady_[d]) {  // preProcessSparseMatrix(Reduced) (simpleassemblers.inl:289-299)
      std::int64_t rows = 0, nnz = 0;
      check(ikb_pattern_nnz(h_, d, &rows, &nnz));
      std::vector<std::int64_t> outer(rows + 1);
      std::vector<std::int32_t> inner(nnz);
      check(ikb_get_pattern(h_, d, outer.data(), inner.data()));
      HostTraits::setPattern(A, rows, outer, inner);
      patternReady_[d] = true;
    }
    check(ikb_get_matrix_values(h_, d, HostTraits::valuePtr(A)));
    return A;
  }

public:
  // ---- DenseFlatAssembler view (assembler/simpleassemblers.hh:188-231, simpleassemblers.inl:301-375) --------
  /** Column-major rows x rows copy of the assembled matrix (Eigen::MatrixXd layout); small problems only. */
  const std::vector<double>& denseMatrix(const FERequirement& req, HostTraits::MatrixAffordance aff,
                                         DBCOption dbc = DBCOption::Full) {
    if (aff != HostTraits::MatrixAffordance::stiffness)
      IKB_THROW(NotImplemented, "MatrixAffordance not implemented");
    push(req);
    const int d = toCode(dbc);
    check(ikb_assemble(h_, IKB_MATRIX, d));
    const std::size_t n = dbc == DBCOption::Reduced ? reducedSize() : size();
    dense_.assign(n * n, 0.0);
    check(ikb_get_dense_matrix(h_, d, dense_.data()));
    return dense_;
  }

  // ---- resident mode ---------------------------------------------------------------------------------
  /** Assemble on the device and return a handle instead of mirroring K to the host. */
  DeviceMatrixHandle deviceMatrix(const FERequirement& req, DBCOption dbc = DBCOption::Full) {
    push(req);
    check(ikb_assemble(h_, IKB_MATRIX | IKB_VECTOR, toCode(dbc)));
    return {h_, toCode(dbc), static_cast<std::int64_t>(dbc == DBCOption::Reduced ? reducedSize() : size())};
  }
  /** Jacobi-PCG on the device for the currently assembled matrix: returns K^-1 rhs (host vector). */
  VectorType solve(DBCOption dbc, const VectorType& rhs, double relTol = 1e-13, int maxIt = -1, int* iterations = nullptr) {
    VectorType x(rhs.size());
    int it = 0;
    double rel = 0;
    check(ikb_pcg_solve(h_, toCode(dbc), rhs.data(), x.data(), relTol, maxIt < 0 ? 2 * static_cast<int>(rhs.size()) : maxIt,
                        &it, &rel));
    if (iterations)
      *iterations = it;
    return x;
  }

  /** Inner problem of TrustRegion on the device: Steihaug-Toint truncated CG for H eta = rhs under the radius
   *  info.delta (linearalgebra/truncatedconjugategradient.hh:68-168); the stop reason, iteration count and the
   *  model terms |eta|, g.eta, eta.H eta come back in `info`. */
  VectorType truncatedCG(DBCOption dbc, const VectorType& rhs, ikb_tcg_info& info) {
    VectorType x(rhs.size());
    check(ikb_tcg_solve(h_, toCode(dbc), rhs.data(), x.data(), &info));
    return x;
  }
  /** utils::obtainForcesDueToIDBC (utils/functionhelper.hh:170-185) with the SpMV on the device:
   *  K_raw * dInc, zeroed at constrained dofs (Full) or reduced.  dInc = d(d_D)/d(lambda), N entries. */
  VectorType forcesDueToIDBC(const FERequirement& req, const VectorType& dInc) {
    if (static_cast<std::size_t>(dInc.size()) != size())
      IKB_THROW(InvalidState, "The increment of the Dirichlet values must have full size.");
    push(req);
    const DBCOption dbc = dBCOption();
    VectorType f(static_cast<typename VectorType::size_type>(dbc == DBCOption::Full ? size() : reducedSize()));
    check(ikb_idbc_forces(h_, toCode(dbc), dInc.data(), f.data()));
    return f;
  }

  // ---- EAS state (mechanics/enhancedassumedstrains.hh:225-248, 350-359) -----------------------------
  /** One call replaces the per-element CORRECTION_UPDATED subscriptions (controlroutinefactory.hh:43-46). */
  void updateInternalVariables(const FERequirement& req, const VectorType& correction) {
    if (static_cast<std::size_t>(correction.size()) != size())
      IKB_THROW(NotImplemented,
                "Solution vector and correction vector should be of the same size. Check if DBCOption::Full is used.");
    push(req);
    check(ikb_eas_update(h_, correction.data()));
  }
#if IKB_HAVE_IKARUS
  /** Register ONE listener on a nonlinear solver / broadcaster for CORRECTION_UPDATED. */
  template <typename BC>
  auto subscribeTo(BC& bc) {
    using NLSState = typename BC::State;
    return bc.template station<Ikarus::NonLinearSolverMessages>().registerListener(
        [this](Ikarus::NonLinearSolverMessages message, const NLSState& state) {
          if (message == Ikarus::NonLinearSolverMessages::CORRECTION_UPDATED)
            this->updateInternalVariables(state.domain, state.correction);
        });
  }
#endif
  // ---- results at local positions ---------------------------------------------------------------------
  /** Number of components of result type rt (IKB_RESULT_*): dim(dim+1)/2 in Voigt form, 6 for the *Full types. */
  int resultComponents(int rt) const {
    return (rt == IKB_RESULT_LINEAR_STRESS_FULL || rt == IKB_RESULT_PK2_STRESS_FULL) ? 6 : dim_ * (dim_ + 1) / 2;
  }
  /** `fe.calculateAt<RT>(req, local)` of EVERY element at nPoints local positions at once (mechanics/
   *  nonlinearelastic.hh:237-271, linearelastic.hh, enhancedassumedstrains.hh:127-187), evaluated on the device:
   *  out[(e*nPoints + q)*ncomp + c], Voigt order.  local: nPoints x dim reference coordinates in [0,1]^dim. */
  std::vector<double> calculateAt(int rt, const FERequirement& req, const double* local, int nPoints) {
    push(req);
    std::vector<double> out(static_cast<std::size_t>(nElem_) * nPoints * resultComponents(rt));
    check(ikb_calculate_at(h_, rt, local, nPoints, out.data()));
    check(ikb_sync(h_));
    return out;
  }
  int worldDimension() const { return dim_; }
  std::int64_t numberOfElements() const { return nElem_; }

  // ---- external loads (mechanics/loads/volume.hh, loads/traction.hh) -----------------------------------
  /** f(x, lambda): the reference's `std::function<Eigen::Vector<double, worldDim>(const FieldVector&, const double&)>`
   *  (loads/volume.hh:26, loads/traction.hh:29) on plain arrays; the first dim entries are used. */
  using LoadFunction = std::function<std::array<double, 3>(const std::array<double, 3>& x, double lambda)>;

  /** volumeLoad<dim>(f) (loads/volume.hh:67-106): R_a -= N_a f(x, lambda) detJ w, E -= u . f(x, lambda) detJ w with the
   *  element's own Gauss rule (order + 1 points per direction). */
  void addVolumeLoad(LoadFunction f) {
    volumeLoads_.push_back(std::move(f));
    loadsSampled_ = false;
  }
  /** neumannBoundaryLoad(&patch, t) (loads/traction.hh:70-138): the faces of the patch as (element index in the
   *  container, DUNE face id: 2k -> xi_k = 0, 2k+1 -> xi_k = 1); face rule of order = basis order (traction.hh:122). */
  void addNeumannBoundaryLoad(std::vector<std::pair<std::int64_t, int>> faces, LoadFunction t) {
    neumannLoads_.push_back({std::move(faces), std::move(t)});
    loadsSampled_ = false;
  }
  /** Extra nodal load vector, proportional to the load factor by default (what the reference's tests add through an
   *  AssemblerManipulator vector callback, tests/src/testcantileverbeam.hh:56-80). */
  void setExternalLoad(const VectorType& fext, bool scalesWithLambda = true) {
    userLoad_.assign(fext.data(), fext.data() + size());
    userScales_   = scalesWithLambda;
    loadsSampled_ = false;
    if (volumeLoads_.empty() && neumannLoads_.empty())
      check(ikb_set_external_load(h_, userLoad_.data(), scalesWithLambda ? 1 : 0));
  }
  /** The nodal load vector of the load skills at load factor lambda, integrated on the host (any dependence on lambda). */
  std::vector<double> sampleLoads(double lambda) const {
    std::vector<double> fext(size(), 0.0);
    const int dim = dim_, n1 = order_ + 1, nc = 1 << dim;
    int nn = 1;
    for (int k = 0; k < dim; ++k)
      nn *= n1;
    auto lagrange = [](int order, double xi, double* v) {
      if (order == 1) {
        v[0] = 1.0 - xi;
        v[1] = xi;
      } else {
        v[0] = 2.0 * (xi - 0.5) * (xi - 1.0);
        v[1] = 4.0 * xi * (1.0 - xi);
        v[2] = 2.0 * xi * (xi - 0.5);
      }
    };
    auto gauss01 = [](int n, double* x, double* w) {
      if (n == 1) {
        x[0] = 0.5, w[0] = 1.0;
      } else if (n == 2) {
        const double a = 0.5 / std::sqrt(3.0);
        x[0] = 0.5 - a, x[1] = 0.5 + a, w[0] = w[1] = 0.5;
      } else {
        const double a = 0.5 * std::sqrt(0.6);
        x[0] = 0.5 - a, x[1] = 0.5, x[2] = 0.5 + a, w[0] = w[2] = 5.0 / 18.0, w[1] = 8.0 / 18.0;
      }
    };
    // shape functions of the basis, geometry (multilinear over the corners) and its Jacobian at xi of element e
    auto at = [&](std::int64_t e, const double* xi, double* N, std::array<double, 3>& x, double Jt[3][3]) {
      double v[3][3];
      for (int k = 0; k < dim; ++k)
        lagrange(order_, xi[k], v[k]);
      for (int a = 0; a < nn; ++a) {
        double p = 1.0;
        for (int k = 0, q = a; k < dim; ++k, q /= n1)
          p *= v[k][q % n1];
        N[a] = p;
      }
      x = {0.0, 0.0, 0.0};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          Jt[i][j] = 0.0;
      const double* X = corners_.data() + static_cast<std::size_t>(e) * nc * dim;
      for (int c = 0; c < nc; ++c) {
        double Ng = 1.0;
        for (int k = 0; k < dim; ++k)
          Ng *= ((c >> k) & 1) ? xi[k] : 1.0 - xi[k];
        for (int j = 0; j < dim; ++j)
          x[j] += Ng * X[c * dim + j];
        for (int i = 0; i < dim; ++i) {
          double dN = ((c >> i) & 1) ? 1.0 : -1.0;
          for (int k = 0; k < dim; ++k)
            if (k != i)
              dN *= ((c >> k) & 1) ? xi[k] : 1.0 - xi[k];
          for (int j = 0; j < dim; ++j)
            Jt[i][j] += dN * X[c * dim + j];  // Jt[i][j] = dx_j / dxi_i
        }
      }
    };
    auto scatter = [&](std::int64_t e, const double* N, const std::array<double, 3>& f, double weight) {
      const std::int64_t* dofs = dofs_.data() + static_cast<std::size_t>(e) * nn * dim;
      for (int a = 0; a < nn; ++a)
        for (int k = 0; k < dim; ++k)
          fext[static_cast<std::size_t>(dofs[a * dim + k])] += N[a] * f[k] * weight;
    };
    double N[27], Jt[3][3];
    std::array<double, 3> x{};
    if (!volumeLoads_.empty()) {
      double gx[3], gw[3];
      gauss01(n1, gx, gw);
      int npts = 1;
      for (int k = 0; k < dim; ++k)
        npts *= n1;
      for (std::int64_t e = 0; e < nElem_; ++e)
        for (int g = 0; g < npts; ++g) {
          double xi[3] = {0, 0, 0}, w = 1.0;
          for (int k = 0, q = g; k < dim; ++k, q /= n1) {
            xi[k] = gx[q % n1];
            w *= gw[q % n1];
          }
          at(e, xi, N, x, Jt);
          const double detJ =
              dim == 2 ? std::abs(Jt[0][0] * Jt[1][1] - Jt[0][1] * Jt[1][0])
                       : std::abs(Jt[0][0] * (Jt[1][1] * Jt[2][2] - Jt[1][2] * Jt[2][1]) - Jt[0][1] * (Jt[1][0] * Jt[2][2] - Jt[1][2] * Jt[2][0]) +
                                  Jt[0][2] * (Jt[1][0] * Jt[2][1] - Jt[1][1] * Jt[2][0]));
          std::array<double, 3> f{0.0, 0.0, 0.0};
          for (const auto& fn : volumeLoads_) {
            const auto v = fn(x, lambda);
            for (int k = 0; k < dim; ++k)
              f[k] += v[k];
          }
          scatter(e, N, f, detJ * w);
        }
    }
    for (const auto& nl : neumannLoads_) {
      const int np = order_ / 2 + 1;
      double gx[3], gw[3];
      gauss01(np, gx, gw);
      int npts = 1;
      for (int k = 0; k + 1 < dim; ++k)
        npts *= np;
      for (const auto& [e, face] : nl.faces) {
        const int kf = face / 2, side = face % 2;
        int dirs[2] = {0, 0}, nd = 0;
        for (int j = 0; j < dim; ++j)
          if (j != kf)
            dirs[nd++] = j;
        for (int g = 0; g < npts; ++g) {
          double xi[3] = {0, 0, 0}, w = 1.0;
          xi[kf] = side;
          for (int t = 0, q = g; t < nd; ++t, q /= np) {
            xi[dirs[t]] = gx[q % np];
            w *= gw[q % np];
          }
          at(e, xi, N, x, Jt);
          double area;
          if (dim == 2) {
            area = std::hypot(Jt[dirs[0]][0], Jt[dirs[0]][1]);
          } else {
            const double* a = Jt[dirs[0]];
            const double* b = Jt[dirs[1]];
            const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
            area = std::sqrt(cx * cx + cy * cy + cz * cz);
          }
          scatter(e, N, nl.t(x, lambda), w * area);
        }
      }
    }
    return fext;
  }
  ikb_handle handle() const { return h_; }

private:
  static int toCode(DBCOption dbc) {
    return dbc == DBCOption::Raw ? IKB_DBC_RAW : (dbc == DBCOption::Reduced ? IKB_DBC_REDUCED : IKB_DBC_FULL);
  }
  /** Uploads d only when it differs from what the device holds: ikb_set_solution invalidates every cached result, and
   *  a Newton iteration asks for residual and tangent of the SAME state one after the other. */
  void push(const FERequirement& req) {
    const auto& d = req.globalSolution();
    assert(static_cast<std::size_t>(d.size()) == size());
    const double* p = d.data();
    if (lastD_.size() != size() || !std::equal(lastD_.begin(), lastD_.end(), p)) {
      check(ikb_set_solution(h_, p));
      check(ikb_sync(h_));  // the copy has read the caller's buffer
      lastD_.assign(p, p + size());
    }
    check(ikb_set_parameter(h_, req.parameter()));
    // load skills: f(x, lambda) is evaluated at the requirement's load factor, like the reference does at every call
    if ((!volumeLoads_.empty() || !neumannLoads_.empty()) && (!loadsSampled_ || sampledLambda_ != req.parameter())) {
      const double lambda = req.parameter();
      loadNow_            = sampleLoads(lambda);
      for (std::size_t i = 0; i < userLoad_.size(); ++i)
        loadNow_[i] += (userScales_ ? lambda : 1.0) * userLoad_[i];
      check(ikb_set_external_load(h_, loadNow_.data(), 0));
      check(ikb_sync(h_));
      loadsSampled_  = true;
      sampledLambda_ = lambda;
    }
  }
public:
  /** The device state was changed behind the wrapper's back (ikb_update_solution, ikb_set_solution on handle()). */
  void invalidateSolutionCache() { lastD_.clear(); }
private:
  void check(int rc) const {
    if (rc == IKB_OK)
      return;
    char buf[512];
    ikb_last_error(h_, buf, sizeof(buf));
    if (rc == IKB_EMATERIAL)
      throw MaterialFailure(buf);
    if (rc == IKB_ENOTIMPL)
      IKB_THROW(NotImplemented, buf);
    IKB_THROW(InvalidState, std::string("libikb200: ") + buf);
  }

  FEC fes_;
  DV dirichletValues_;  // copied like in the reference (interface.hh:260)
  std::optional<std::reference_wrapper<const FERequirement>> req_;
  std::optional<AffordanceCollectionType> aff_;
  std::optional<DBCOption> dbc_;
  std::vector<std::size_t> constraintsBelow_;
  std::size_t fixedDofs_{};
  std::size_t vertexCount_{0};
  bool fusedSweep_{true};
  std::vector<double> lastD_;
  // host copies of the mesh and the load skills
  int dim_{0}, order_{0};
  std::int64_t nElem_{0};
  std::vector<double> corners_;
  std::vector<std::int64_t> dofs_;
  struct NeumannLoad
  {
    std::vector<std::pair<std::int64_t, int>> faces;
    LoadFunction t;
  };
  std::vector<LoadFunction> volumeLoads_;
  std::vector<NeumannLoad> neumannLoads_;
  std::vector<double> userLoad_, loadNow_;
  bool userScales_{true}, loadsSampled_{false};
  double sampledLambda_{0.0};
  ikb_handle h_{nullptr};
  ScalarType scal_{0.0};
  VectorType vec_[3];
  MatrixType mat_[3];
  bool patternReady_[3]{false, false, false};
  std::vector<double> dense_;
};

/** makeSparseFlatAssembler analogue (assembler/simpleassemblers.hh:174-177). */
template <typename FEC, typename DV>
auto makeDeviceSparseFlatAssembler(FEC&& fes, const DV& dirichletValues, int device = -1) {
  return std::make_shared<DeviceSparseFlatAssembler<FEC, DV>>(std::forward<FEC>(fes), dirichletValues, device);
}

/** io/resultfunction.hh:19-37: the default user function returns the requested component unchanged. */
struct DefaultResultUserFunction
{
  template <typename R, typename FiniteElement>
  double operator()(const R& resultArray, const double* /*pos*/, const FiniteElement& /*fe*/, int comp) const {
    return resultArray[comp];
  }
};

/**
 * \brief Mirror of Ikarus::ResultFunction (io/resultfunction.hh:57-157): evaluates a result type of the bound
 * requirement at local positions of single elements, for a VTK writer.
 * \details The reference calls `fe.calculateAt<RT>(requirement(), local)` for one element at a time; a writer asks for
 * the same few local positions (the element vertices) on every element, so here the first request for a position
 * evaluates ALL elements on the device (ikb_calculate_at) and the following ones are served from that table.  The
 * table is dropped when the solution or the load factor of the bound requirement changes.
 * With Ikarus and dune-grid present (IKB_HAVE_DUNE_VTK) the class derives from Dune::VTKFunction<GridView> and `evaluate(comp, entity, local)` maps
 * the entity through `gridView().indexSet().index(e)` exactly like resultfunction.hh:87-90.
 * \tparam AS DeviceSparseFlatAssembler; RT: IKB_RESULT_* code of the result type; UF: user function
 * `(resultArray, pos, fe, comp) -> double` with optional `ncomps()` / `name()` (resultfunction.hh:100-121).
 */
template <typename AS, typename UserFunction = DefaultResultUserFunction>
class DeviceResultFunction
#if IKB_HAVE_DUNE_VTK
    : public Dune::VTKFunction<typename AS::GridView>
#endif
{
public:
  using Assembler = AS;

  DeviceResultFunction(std::shared_ptr<AS> assembler, int resultType, UserFunction userFunction = {})
      : assembler_{std::move(assembler)},
        rt_{resultType},
        userFunction_{std::move(userFunction)} {}

  /** resultfunction.hh:143-148 with the element given by its index in the container */
  double evaluate(int comp, std::int64_t elementIndex, const double* local) const {
    const int nc        = assembler_->resultComponents(rt_);
    const double* table = tableFor(local);
    const double* r     = table + static_cast<std::size_t>(elementIndex) * nc;
    auto it             = std::begin(assembler_->finiteElements());
    std::advance(it, elementIndex);
    return userFunction_(r, local, *it, comp);
  }
#if IKB_HAVE_DUNE_VTK
  using GridView = typename AS::GridView;
  using Entity   = typename GridView::template Codim<0>::Entity;
  using ctype    = typename GridView::ctype;
  double evaluate(int comp, const Entity& e, const Dune::FieldVector<ctype, GridView::dimension>& local) const override {
    double xi[3] = {0, 0, 0};
    for (int k = 0; k < GridView::dimension; ++k)
      xi[k] = local[k];
    return evaluate(comp, static_cast<std::int64_t>(assembler_->gridView().indexSet().index(e)), xi);
  }
  [[nodiscard]] int ncomps() const override { return ncompsImpl(); }
  [[nodiscard]] std::string name() const override { return nameImpl(); }
#else
  [[nodiscard]] int ncomps() const { return ncompsImpl(); }
  [[nodiscard]] std::string name() const { return nameImpl(); }
#endif

private:
  int ncompsImpl() const {
    if constexpr (requires { userFunction_.ncomps(); })
      return userFunction_.ncomps();
    else
      return assembler_->resultComponents(rt_);
  }
  std::string nameImpl() const {
    if constexpr (requires { userFunction_.name(); })
      return userFunction_.name();
    else {
      static const char* names[] = {"linearStress", "PK2Stress", "linearStressFull", "PK2StressFull", "kirchhoffStress", "cauchyStress"};
      return names[rt_];
    }
  }
  const double* tableFor(const double* local) const {
    const auto& req = assembler_->requirement();
    const auto& d   = req.globalSolution();
    // a new state of the bound requirement invalidates every table
    if (lambda_ != req.parameter() || lastD_.size() != static_cast<std::size_t>(d.size()) ||
        !std::equal(lastD_.begin(), lastD_.end(), d.data())) {
      tables_.clear();
      lastD_.assign(d.data(), d.data() + d.size());
      lambda_ = req.parameter();
    }
    const int dim = assembler_->worldDimension();
    std::array<double, 3> key{0, 0, 0};
    for (int k = 0; k < dim; ++k)
      key[k] = local[k];
    for (const auto& t : tables_)
      if (t.first == key)
        return t.second.data();
    tables_.emplace_back(key, assembler_->calculateAt(rt_, req, key.data(), 1));
    return tables_.back().second.data();
  }

  std::shared_ptr<AS> assembler_;
  int rt_;
  UserFunction userFunction_;
  mutable std::vector<std::pair<std::array<double, 3>, std::vector<double>>> tables_;
  mutable std::vector<double> lastD_;
  mutable double lambda_{std::numeric_limits<double>::quiet_NaN()};
};

/** makeResultFunction<RT>(assembler, ...) (io/resultfunction.hh:172-178) with the result type as IKB_RESULT_* code */
template <typename AS, typename UserFunction = DefaultResultUserFunction>
auto makeResultFunction(std::shared_ptr<AS> assembler, int resultType, UserFunction&& userFunction = {}) {
  return std::make_shared<DeviceResultFunction<AS, std::remove_cvref_t<UserFunction>>>(std::move(assembler), resultType,
                                                                                       std::forward<UserFunction>(userFunction));
}

/**
 * \brief Linear-solver callable for NewtonRaphson: `correction_ = -linearSolver_(rx, Ax)`
 * (solver/nonlinearsolver/newtonraphson.hh:152, 221-226).  Runs Jacobi-PCG on the device on the matrix the
 * assembler holds; Ax is only used for its type.  The analogue of SolverTypeTag::si_ConjugateGradient.
 */
template <typename Assembler>
struct DevicePCG
{
  std::shared_ptr<Assembler> assembler;
  double relTol{1e-13};
  int maxIt{-1};
  mutable int lastIterations{0};
  template <typename R, typename K>
  auto operator()(const R& rx, const K&) const {
    return assembler->solve(assembler->dBCOption(), rx, relTol, maxIt, &lastIterations);
  }
};

} // namespace Ikarus::B200
