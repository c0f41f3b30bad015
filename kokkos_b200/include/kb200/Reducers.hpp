// kb200/Reducers.hpp -- the built-in reducers of the hot path, same names and semantics as
// core/src/Kokkos_Parallel_Reduce.hpp:33-1352 and identities as core/src/Kokkos_ReductionIdentity.hpp.
//
//   Sum, Prod, Min, Max, LAnd, LOr, BAnd, BOr, MinLoc, MaxLoc, MinMax, MinMaxLoc,
//   MinFirstLoc, MaxFirstLoc, MinMaxFirstLastLoc, FirstLoc, LastLoc
//   ValLocScalar, MinMaxScalar, MinMaxLocScalar, reduction_identity<T>
//
// A reducer is constructed from a scalar reference (blocking, result on return) or from anything
// with .data() pointing at device-accessible memory, e.g. a rank-0 View (asynchronous) --
// Kokkos_Parallel_Reduce.hpp:1417-1468.
//
// One deliberate strengthening: the loc reducers break value ties towards the LOWER location.
// The reference keeps `dest` on ties (Kokkos_Parallel_Reduce.hpp:441-449,628-644), which makes the
// result depend on the combine order; on its OpenMP backend (static schedule, strict-compare functor,
// thread-ordered joins: OpenMP/Kokkos_OpenMP_Parallel_Reduce.hpp:147-151) that order yields exactly
// "lowest index wins".  Using that rule in join() makes it commutative, so the B200 result is
// bit-identical to the OpenMP oracle for any grid shape, ties included.
#ifndef KB200_REDUCERS_HPP
#define KB200_REDUCERS_HPP

#include "Macros.hpp"
#include "Complex.hpp"
#include "impl/Collectives.hpp"
#include <cfloat>
#include <climits>
#include <type_traits>

namespace kb200 {

struct B200Space;  // device memory space tag (View.hpp)
struct HostSpace;
template <class DataType, class... Props>
class View;

// ---------------------------------------------------------------- identities
template <class T, class Enable = void>
struct reduction_identity;  // user types specialise this, as with Kokkos::reduction_identity

#define KB200_IDENTITY_INT(T, TMIN, TMAX)                                    \
  template <>                                                                \
  struct reduction_identity<T> {                                             \
    KB200_FORCEINLINE_FUNCTION constexpr static T sum() { return (T)0; }      \
    KB200_FORCEINLINE_FUNCTION constexpr static T prod() { return (T)1; }     \
    KB200_FORCEINLINE_FUNCTION constexpr static T max() { return TMIN; }     \
    KB200_FORCEINLINE_FUNCTION constexpr static T min() { return TMAX; }     \
    KB200_FORCEINLINE_FUNCTION constexpr static T bor() { return (T)0; }      \
    KB200_FORCEINLINE_FUNCTION constexpr static T band() { return (T) ~(T)0; } \
    KB200_FORCEINLINE_FUNCTION constexpr static T lor() { return (T)0; }      \
    KB200_FORCEINLINE_FUNCTION constexpr static T land() { return (T)1; }     \
  };
KB200_IDENTITY_INT(char, CHAR_MIN, CHAR_MAX)
KB200_IDENTITY_INT(signed char, SCHAR_MIN, SCHAR_MAX)
KB200_IDENTITY_INT(unsigned char, 0, UCHAR_MAX)
KB200_IDENTITY_INT(short, SHRT_MIN, SHRT_MAX)
KB200_IDENTITY_INT(unsigned short, 0, USHRT_MAX)
KB200_IDENTITY_INT(int, INT_MIN, INT_MAX)
KB200_IDENTITY_INT(unsigned int, 0u, UINT_MAX)
KB200_IDENTITY_INT(long, LONG_MIN, LONG_MAX)
KB200_IDENTITY_INT(unsigned long, 0ul, ULONG_MAX)
KB200_IDENTITY_INT(long long, LLONG_MIN, LLONG_MAX)
KB200_IDENTITY_INT(unsigned long long, 0ull, ULLONG_MAX)
#undef KB200_IDENTITY_INT
template <>
struct reduction_identity<bool> {
  KB200_FORCEINLINE_FUNCTION constexpr static bool lor() { return false; }
  KB200_FORCEINLINE_FUNCTION constexpr static bool land() { return true; }
};
#define KB200_IDENTITY_FP(T, TMAX)                                         \
  template <>                                                              \
  struct reduction_identity<T> {                                           \
    KB200_FORCEINLINE_FUNCTION constexpr static T sum() { return T(0); }   \
    KB200_FORCEINLINE_FUNCTION constexpr static T prod() { return T(1); }  \
    KB200_FORCEINLINE_FUNCTION constexpr static T max() { return -TMAX; }  \
    KB200_FORCEINLINE_FUNCTION constexpr static T min() { return TMAX; }   \
  };
KB200_IDENTITY_FP(float, FLT_MAX)
KB200_IDENTITY_FP(double, DBL_MAX)
#undef KB200_IDENTITY_FP
template <class T>
struct reduction_identity<complex<T>> {  // core/src/Kokkos_Complex.hpp:905-925
  KB200_FORCEINLINE_FUNCTION constexpr static complex<T> sum() noexcept { return complex<T>(reduction_identity<T>::sum(), reduction_identity<T>::sum()); }
  KB200_FORCEINLINE_FUNCTION constexpr static complex<T> prod() noexcept { return complex<T>(reduction_identity<T>::prod(), reduction_identity<T>::sum()); }
};

// ---------------------------------------------------------------- value structs
template <class Scalar, class Index>
struct ValLocScalar { Scalar val; Index loc; };
template <class Scalar>
struct MinMaxScalar { Scalar min_val, max_val; };
template <class Scalar, class Index>
struct MinMaxLocScalar { Scalar min_val, max_val; Index min_loc, max_loc; };

namespace Impl {
// common storage: where the result goes and whether that is a host scalar
template <class V>
struct ReducerBase {
  using value_type = V;
  value_type* m_ptr;
  bool m_scalar;
  KB200_INLINE_FUNCTION ReducerBase(value_type& v) : m_ptr(&v), m_scalar(true) {}
  // result View: device memory => written by the kernel (asynchronous); host-space View => treated like a scalar (the value
  // is stored after the wait, core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:344-360 "result ptr host accessible")
  template <class ViewLike, class = void>
  struct view_on_host : std::false_type {};
  template <class ViewLike>
  struct view_on_host<ViewLike, std::void_t<typename ViewLike::memory_space>>
      : std::integral_constant<bool, !ViewLike::is_device> {};
  template <class ViewLike, class = decltype(std::declval<const ViewLike&>().data())>
  KB200_INLINE_FUNCTION ReducerBase(const ViewLike& v, int = 0) : m_ptr(v.data()), m_scalar(view_on_host<ViewLike>::value) {}
  KB200_INLINE_FUNCTION value_type& reference() const { return *m_ptr; }
  KB200_INLINE_FUNCTION value_type* data() const { return m_ptr; }
  KB200_INLINE_FUNCTION bool references_scalar() const { return m_scalar; }
  KB200_INLINE_FUNCTION void final(value_type&) const {}
};
}  // namespace Impl

#define KB200_REDUCER_HEAD(Name, ...)                         \
  using reducer = Name;                                       \
  using value_type = __VA_ARGS__;                             \
  using result_view_type = View<__VA_ARGS__, Space>;          \
  using Base = Impl::ReducerBase<__VA_ARGS__>;                \
  using Base::Base;

template <class Scalar, class Space = HostSpace>
struct Sum : Impl::ReducerBase<std::remove_cv_t<Scalar>> {
  KB200_REDUCER_HEAD(Sum, std::remove_cv_t<Scalar>)
  static constexpr int redux_op = Impl::ReduxAdd;
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const { dest += src; }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v = reduction_identity<value_type>::sum(); }
};
template <class Scalar, class Space = HostSpace>
struct Prod : Impl::ReducerBase<std::remove_cv_t<Scalar>> {
  KB200_REDUCER_HEAD(Prod, std::remove_cv_t<Scalar>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const { dest *= src; }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v = reduction_identity<value_type>::prod(); }
};
template <class Scalar, class Space = HostSpace>
struct Min : Impl::ReducerBase<std::remove_cv_t<Scalar>> {
  KB200_REDUCER_HEAD(Min, std::remove_cv_t<Scalar>)
  static constexpr int redux_op = Impl::ReduxMin;
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const { if (src < dest) dest = src; }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v = reduction_identity<value_type>::min(); }
};
template <class Scalar, class Space = HostSpace>
struct Max : Impl::ReducerBase<std::remove_cv_t<Scalar>> {
  KB200_REDUCER_HEAD(Max, std::remove_cv_t<Scalar>)
  static constexpr int redux_op = Impl::ReduxMax;
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const { if (src > dest) dest = src; }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v = reduction_identity<value_type>::max(); }
};
template <class Scalar, class Space = HostSpace>
struct LAnd : Impl::ReducerBase<std::remove_cv_t<Scalar>> {
  KB200_REDUCER_HEAD(LAnd, std::remove_cv_t<Scalar>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const { dest = dest && src; }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v = reduction_identity<value_type>::land(); }
};
template <class Scalar, class Space = HostSpace>
struct LOr : Impl::ReducerBase<std::remove_cv_t<Scalar>> {
  KB200_REDUCER_HEAD(LOr, std::remove_cv_t<Scalar>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const { dest = dest || src; }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v = reduction_identity<value_type>::lor(); }
};
template <class Scalar, class Space = HostSpace>
struct BAnd : Impl::ReducerBase<std::remove_cv_t<Scalar>> {
  KB200_REDUCER_HEAD(BAnd, std::remove_cv_t<Scalar>)
  static constexpr int redux_op = Impl::ReduxAnd;
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const { dest = dest & src; }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v = reduction_identity<value_type>::band(); }
};
template <class Scalar, class Space = HostSpace>
struct BOr : Impl::ReducerBase<std::remove_cv_t<Scalar>> {
  KB200_REDUCER_HEAD(BOr, std::remove_cv_t<Scalar>)
  static constexpr int redux_op = Impl::ReduxOr;
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const { dest = dest | src; }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v = reduction_identity<value_type>::bor(); }
};

template <class Scalar, class Index, class Space = HostSpace>
struct MinLoc : Impl::ReducerBase<ValLocScalar<std::remove_cv_t<Scalar>, std::remove_cv_t<Index>>> {
  using scalar_type = std::remove_cv_t<Scalar>;
  using index_type = std::remove_cv_t<Index>;
  KB200_REDUCER_HEAD(MinLoc, ValLocScalar<scalar_type, index_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (src.val < dest.val || (src.val == dest.val && src.loc < dest.loc)) dest = src;
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const {
    v.val = reduction_identity<scalar_type>::min();
    v.loc = reduction_identity<index_type>::min();
  }
};
template <class Scalar, class Index, class Space = HostSpace>
struct MaxLoc : Impl::ReducerBase<ValLocScalar<std::remove_cv_t<Scalar>, std::remove_cv_t<Index>>> {
  using scalar_type = std::remove_cv_t<Scalar>;
  using index_type = std::remove_cv_t<Index>;
  KB200_REDUCER_HEAD(MaxLoc, ValLocScalar<scalar_type, index_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (src.val > dest.val || (src.val == dest.val && src.loc < dest.loc)) dest = src;
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const {
    v.val = reduction_identity<scalar_type>::max();
    v.loc = reduction_identity<index_type>::min();
  }
};
template <class Scalar, class Space = HostSpace>
struct MinMax : Impl::ReducerBase<MinMaxScalar<std::remove_cv_t<Scalar>>> {
  using scalar_type = std::remove_cv_t<Scalar>;
  KB200_REDUCER_HEAD(MinMax, MinMaxScalar<scalar_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (src.min_val < dest.min_val) dest.min_val = src.min_val;
    if (src.max_val > dest.max_val) dest.max_val = src.max_val;
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const {
    v.max_val = reduction_identity<scalar_type>::max();
    v.min_val = reduction_identity<scalar_type>::min();
  }
};
template <class Scalar, class Index, class Space = HostSpace>
struct MinMaxLoc : Impl::ReducerBase<MinMaxLocScalar<std::remove_cv_t<Scalar>, std::remove_cv_t<Index>>> {
  using scalar_type = std::remove_cv_t<Scalar>;
  using index_type = std::remove_cv_t<Index>;
  KB200_REDUCER_HEAD(MinMaxLoc, MinMaxLocScalar<scalar_type, index_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (src.min_val < dest.min_val || (src.min_val == dest.min_val && src.min_loc < dest.min_loc)) {
      dest.min_val = src.min_val;
      dest.min_loc = src.min_loc;
    }
    if (src.max_val > dest.max_val || (src.max_val == dest.max_val && src.max_loc < dest.max_loc)) {
      dest.max_val = src.max_val;
      dest.max_loc = src.max_loc;
    }
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const {
    v.max_val = reduction_identity<scalar_type>::max();
    v.min_val = reduction_identity<scalar_type>::min();
    v.max_loc = reduction_identity<index_type>::min();
    v.min_loc = reduction_identity<index_type>::min();
  }
};

// ---- first/last-location reducers used by the std-algorithm layer (core/src/Kokkos_Parallel_Reduce.hpp:679-1228):
//      MaxFirstLoc / MinFirstLoc keep the LOWEST location among equal extrema, MinMaxFirstLastLoc the lowest location of the
//      minimum and the HIGHEST location of the maximum (std::minmax_element), FirstLoc / LastLoc the lowest / highest
//      location at which a predicate held.  All joins are commutative.
template <class Scalar, class Index, class Space = HostSpace>
struct MaxFirstLoc : Impl::ReducerBase<ValLocScalar<std::remove_cv_t<Scalar>, std::remove_cv_t<Index>>> {
  using scalar_type = std::remove_cv_t<Scalar>;
  using index_type = std::remove_cv_t<Index>;
  KB200_REDUCER_HEAD(MaxFirstLoc, ValLocScalar<scalar_type, index_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (dest.val < src.val) dest = src;
    else if (!(src.val < dest.val) && src.loc < dest.loc) dest.loc = src.loc;
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const {
    v.val = reduction_identity<scalar_type>::max();
    v.loc = reduction_identity<index_type>::min();
  }
};
template <class Scalar, class Index, class Space = HostSpace>
struct MinFirstLoc : Impl::ReducerBase<ValLocScalar<std::remove_cv_t<Scalar>, std::remove_cv_t<Index>>> {
  using scalar_type = std::remove_cv_t<Scalar>;
  using index_type = std::remove_cv_t<Index>;
  KB200_REDUCER_HEAD(MinFirstLoc, ValLocScalar<scalar_type, index_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (src.val < dest.val) dest = src;
    else if (!(dest.val < src.val) && src.loc < dest.loc) dest.loc = src.loc;
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const {
    v.val = reduction_identity<scalar_type>::min();
    v.loc = reduction_identity<index_type>::min();
  }
};
template <class Scalar, class Index, class Space = HostSpace>
struct MinMaxFirstLastLoc : Impl::ReducerBase<MinMaxLocScalar<std::remove_cv_t<Scalar>, std::remove_cv_t<Index>>> {
  using scalar_type = std::remove_cv_t<Scalar>;
  using index_type = std::remove_cv_t<Index>;
  KB200_REDUCER_HEAD(MinMaxFirstLastLoc, MinMaxLocScalar<scalar_type, index_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (src.min_val < dest.min_val) { dest.min_val = src.min_val; dest.min_loc = src.min_loc; }
    else if (!(dest.min_val < src.min_val) && src.min_loc < dest.min_loc) dest.min_loc = src.min_loc;
    if (dest.max_val < src.max_val) { dest.max_val = src.max_val; dest.max_loc = src.max_loc; }
    else if (!(src.max_val < dest.max_val) && src.max_loc > dest.max_loc) dest.max_loc = src.max_loc;
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const {
    v.max_val = reduction_identity<scalar_type>::max();
    v.min_val = reduction_identity<scalar_type>::min();
    v.max_loc = reduction_identity<index_type>::max();
    v.min_loc = reduction_identity<index_type>::min();
  }
};
template <class Index>
struct FirstLocScalar { Index min_loc_true; };
template <class Index>
struct LastLocScalar { Index max_loc_true; };
template <class Index, class Space = HostSpace>
struct FirstLoc : Impl::ReducerBase<FirstLocScalar<std::remove_cv_t<Index>>> {
  using index_type = std::remove_cv_t<Index>;
  KB200_REDUCER_HEAD(FirstLoc, FirstLocScalar<index_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (src.min_loc_true < dest.min_loc_true) dest.min_loc_true = src.min_loc_true;
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v.min_loc_true = reduction_identity<index_type>::min(); }
};
template <class Index, class Space = HostSpace>
struct LastLoc : Impl::ReducerBase<LastLocScalar<std::remove_cv_t<Index>>> {
  using index_type = std::remove_cv_t<Index>;
  KB200_REDUCER_HEAD(LastLoc, LastLocScalar<index_type>)
  KB200_FORCEINLINE_FUNCTION void join(value_type& dest, const value_type& src) const {
    if (src.max_loc_true > dest.max_loc_true) dest.max_loc_true = src.max_loc_true;
  }
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { v.max_loc_true = reduction_identity<index_type>::max(); }
};
#undef KB200_REDUCER_HEAD

template <class T>
struct is_reducer : std::false_type {};
template <class S, class Sp> struct is_reducer<Sum<S, Sp>> : std::true_type {};
template <class S, class Sp> struct is_reducer<Prod<S, Sp>> : std::true_type {};
template <class S, class Sp> struct is_reducer<Min<S, Sp>> : std::true_type {};
template <class S, class Sp> struct is_reducer<Max<S, Sp>> : std::true_type {};
template <class S, class Sp> struct is_reducer<LAnd<S, Sp>> : std::true_type {};
template <class S, class Sp> struct is_reducer<LOr<S, Sp>> : std::true_type {};
template <class S, class Sp> struct is_reducer<BAnd<S, Sp>> : std::true_type {};
template <class S, class Sp> struct is_reducer<BOr<S, Sp>> : std::true_type {};
template <class S, class I, class Sp> struct is_reducer<MinLoc<S, I, Sp>> : std::true_type {};
template <class S, class I, class Sp> struct is_reducer<MaxLoc<S, I, Sp>> : std::true_type {};
template <class S, class Sp> struct is_reducer<MinMax<S, Sp>> : std::true_type {};
template <class S, class I, class Sp> struct is_reducer<MinMaxLoc<S, I, Sp>> : std::true_type {};
template <class S, class I, class Sp> struct is_reducer<MaxFirstLoc<S, I, Sp>> : std::true_type {};
template <class S, class I, class Sp> struct is_reducer<MinFirstLoc<S, I, Sp>> : std::true_type {};
template <class S, class I, class Sp> struct is_reducer<MinMaxFirstLastLoc<S, I, Sp>> : std::true_type {};
template <class I, class Sp> struct is_reducer<FirstLoc<I, Sp>> : std::true_type {};
template <class I, class Sp> struct is_reducer<LastLoc<I, Sp>> : std::true_type {};
// user-defined reducers: anything that names itself in a nested `reducer` typedef
// (the reference's detection: Kokkos_Parallel_Reduce.hpp is_reducer via T::reducer)
template <class T, class = void>
struct has_reducer_typedef : std::false_type {};
template <class T>
struct has_reducer_typedef<T, std::void_t<typename T::reducer>> : std::true_type {};
template <class T>
constexpr bool is_reducer_v = is_reducer<T>::value || has_reducer_typedef<T>::value;

}  // namespace kb200
#endif
