// Restatement of Concepts::FlatAssembler / ScalarFlatAssembler / VectorFlatAssembler / MatrixFlatAssembler
// (ikarus/utils/concepts.hh:517-585 of the reference) for the tests: the same member list, parametrised on the vector
// and DBCOption types so that it can be checked in both build modes of deviceflatassembler.hh.
#pragma once
#include <concepts>
#include <cstddef>

namespace TestConcepts {

template <typename T, typename Vec, typename Dbc>
concept FlatAssembler = requires(T t, const typename T::FERequirement& req, typename T::AffordanceCollectionType aff, Dbc dbc) {
  { t.requirement() } -> std::convertible_to<const typename T::FERequirement&>;
  { t.affordanceCollection() } -> std::convertible_to<typename T::AffordanceCollectionType>;
  { t.dBCOption() } -> std::convertible_to<Dbc>;
  { t.bind(req, aff, dbc) } -> std::same_as<void>;
  { t.bind(req) } -> std::same_as<void>;
  { t.bind(aff) } -> std::same_as<void>;
  { t.bind(dbc) } -> std::same_as<void>;
  { t.bound() } -> std::convertible_to<bool>;
  { t.boundToRequirement() } -> std::convertible_to<bool>;
  { t.boundToAffordanceCollection() } -> std::convertible_to<bool>;
  { t.boundToDBCOption() } -> std::convertible_to<bool>;
  { t.estimateOfConnectivity() } -> std::convertible_to<std::size_t>;
  { t.createFullVector(std::declval<const Vec&>()) } -> std::convertible_to<Vec>;
  { t.constraintsBelow(std::declval<std::size_t>()) } -> std::convertible_to<std::size_t>;
  { t.isConstrained(std::declval<std::size_t>()) } -> std::convertible_to<bool>;
  { t.size() } -> std::convertible_to<std::size_t>;
  { t.reducedSize() } -> std::convertible_to<std::size_t>;
};

template <typename T, typename Vec, typename Dbc>
concept ScalarFlatAssembler = FlatAssembler<T, Vec, Dbc> and
    requires(T t, const typename T::FERequirement& req, typename T::AffordanceCollectionType aff) {
      { t.scalar(req, aff.scalarAffordance()) } -> std::convertible_to<const double&>;
      { t.scalar() } -> std::convertible_to<const double&>;
    };

template <typename T, typename Vec, typename Dbc>
concept VectorFlatAssembler = ScalarFlatAssembler<T, Vec, Dbc> and
    requires(T t, const typename T::FERequirement& req, typename T::AffordanceCollectionType aff, Dbc dbc) {
      { t.vector(req, aff.vectorAffordance(), dbc) } -> std::convertible_to<const Vec&>;
      { t.vector(dbc) } -> std::convertible_to<const Vec&>;
      { t.vector() } -> std::convertible_to<const Vec&>;
    };

template <typename T, typename Vec, typename Dbc>
concept MatrixFlatAssembler = VectorFlatAssembler<T, Vec, Dbc> and
    requires(T t, const typename T::FERequirement& req, typename T::AffordanceCollectionType aff, Dbc dbc) {
      { t.matrix(req, aff.matrixAffordance(), dbc) };
      { t.matrix(dbc) };
      { t.matrix() };
    };

/**
 * What the reference's AssemblerManipulator does to the assembler it wraps (assemblermanipulatorfuser.hh:242-385,
 * assemblermanipulatorbuildingblocks.hh:29-189): it derives PRIVATELY from it, reaches the get*Impl hooks of the base
 * and runs callbacks (assembler, requirement, affordance, dbcOption, quantity&) on what they return.  Restated here
 * for the vector and matrix hooks; compiles only if the wrapped class exposes the hooks to derived classes.
 */
template <typename A, typename Dbc>
class Manipulator : private A
{
public:
  using A::A;
  using A::bind;
  using A::dBCOption;
  using A::handle;
  using A::requirement;
  using A::size;
  using VectorType = typename A::VectorType;
  using MatrixType = typename A::MatrixType;
  using VecFn = void (*)(const A&, const typename A::FERequirement&, Dbc, VectorType&);
  using MatFn = void (*)(const A&, const typename A::FERequirement&, Dbc, MatrixType&);
  VecFn vf{nullptr};
  MatFn mf{nullptr};
  template <typename VA>
  const VectorType& vector(const typename A::FERequirement& req, VA aff, Dbc dbc) {
    VectorType& v = dbc == Dbc::Raw ? A::getRawVectorImpl(req, aff)
                                    : (dbc == Dbc::Reduced ? A::getReducedVectorImpl(req, aff) : A::getVectorImpl(req, aff));
    if (vf) vf(base(), req, dbc, v);
    return v;
  }
  template <typename MA>
  const MatrixType& matrix(const typename A::FERequirement& req, MA aff, Dbc dbc) {
    MatrixType& m = dbc == Dbc::Raw ? A::getRawMatrixImpl(req, aff)
                                    : (dbc == Dbc::Reduced ? A::getReducedMatrixImpl(req, aff) : A::getMatrixImpl(req, aff));
    if (mf) mf(base(), req, dbc, m);
    return m;
  }
  template <typename SA>
  const double& scalar(const typename A::FERequirement& req, SA aff) { return A::getScalarImpl(req, aff); }
  const A& base() const { return *this; }
};

}  // namespace TestConcepts
