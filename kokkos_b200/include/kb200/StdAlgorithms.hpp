// kb200/StdAlgorithms.hpp -- the scan / reduction members of Kokkos' std-algorithm layer on the B200 execution space,
// plus ViewFill-style deep_copy(view, value).  SURVEY.md 8(f) ranks 2-3: direct consumers of parallel_scan / parallel_reduce.
//
// Mirrors the View overloads of algorithms/src/std_algorithms (Kokkos::Experimental::):
//   exclusive_scan  (Kokkos_ExclusiveScan.hpp)   inclusive_scan (Kokkos_InclusiveScan.hpp)   reduce (Kokkos_Reduce.hpp)
//   transform_reduce (Kokkos_TransformReduce.hpp) min_element / max_element / minmax_element (Kokkos_MinMaxMinmaxElement.hpp)
//   fill / copy / transform / count_if / for_each / iota-free subset (Kokkos_Fill.hpp, Kokkos_Copy.hpp, ...)
// with one deliberate difference: the reference returns iterators (begin(view) + n); rank-1 Views here have no iterator
// type, so the element-returning functions return the INDEX (= std::distance(begin, it)) and the range-writing functions
// return the number of elements written.
//
// On the reference's Cuda backend inclusive/exclusive_scan call Thrust/CUB (impl/Kokkos_InclusiveScan.hpp:156-176).  Here
// the default-operator scans and sums over contiguous 4/8-byte arithmetic Views go straight to the single-pass TMA +
// decoupled-look-back kernel / the 256-bit-load reduction of libkokkos_b200.so (the typed C-ABI fast paths); everything else
// (custom operators, other value types) runs the generic parallel_scan / parallel_reduce with a lambda.
#ifndef KB200_STDALGORITHMS_HPP
#define KB200_STDALGORITHMS_HPP

#include "Parallel.hpp"
#include <limits>
#include <string>
#include <utility>

namespace kb200 {

// ---------------------------------------------------------------- ViewCopy for strided / layout-changing pairs
// (core/src/Kokkos_CopyViews.hpp:300-560 ViewCopy).  The index box is flattened with the dimension of smallest destination
// stride running fastest, so the writes of a warp are contiguous whenever the destination has a unit-stride dimension.
namespace Impl {
template <int R>
struct StridedBox {
  size_t ext[R > 0 ? R : 1], dst_stride[R > 0 ? R : 1], src_stride[R > 0 ? R : 1];
  KB200_INLINE_FUNCTION void offsets(size_t flat, size_t& od, size_t& os) const {
    od = 0; os = 0;
    for (int r = 0; r < R; ++r) {
      const size_t c = flat % ext[r];
      flat /= ext[r];
      od += c * dst_stride[r];
      os += c * src_stride[r];
    }
  }
};
template <class V1, class V2>
StridedBox<V1::rank> make_strided_box(const V1& dst, const V2& src) {
  constexpr int R = V1::rank;
  StridedBox<R> b;
  int order[R > 0 ? R : 1];
  for (int r = 0; r < R; ++r) order[r] = r;
  for (int i = 1; i < R; ++i)  // insertion sort by destination stride
    for (int j = i; j > 0 && dst.stride(order[j]) < dst.stride(order[j - 1]); --j) std::swap(order[j], order[j - 1]);
  for (int r = 0; r < R; ++r) { b.ext[r] = dst.extent(order[r]); b.dst_stride[r] = dst.stride(order[r]); b.src_stride[r] = src.stride(order[r]); }
  return b;
}
template <class V1, class V2>
void strided_copy(const B200& space, const V1& dst, const V2& src) {
  using T = typename V1::non_const_value_type;
  const size_t n = dst.size();
  if (n == 0) return;
  const StridedBox<V1::rank> box = make_strided_box(dst, src);
  T* dp = const_cast<T*>(dst.data());
  const T* sp = src.data();
  if constexpr (V1::is_device && V2::is_device) {
    parallel_for("kb200::ViewCopy", RangePolicy<>(space, 0, (long long)n), KB200_LAMBDA(const long long i) {
      size_t od, os;
      box.offsets((size_t)i, od, os);
      dp[od] = sp[os];
    });
  } else if constexpr (!V1::is_device && !V2::is_device) {
    space.fence("kb200::deep_copy(host strided view, host view)");
    for (size_t i = 0; i < n; ++i) {
      size_t od, os;
      box.offsets(i, od, os);
      dp[od] = sp[os];
    }
  } else {
    throw std::runtime_error("kb200::deep_copy: a strided or layout-changing copy between host and device needs a contiguous mirror on one side");
  }
}
}  // namespace Impl

// ---------------------------------------------------------------- ViewFill: deep_copy(view, value)
// (core/src/Kokkos_CopyViews.hpp:58-300 ViewFill; zero bit patterns take the memset path like ZeroMemset<Cuda>)
template <class D, class... P>
void deep_copy(const B200& space, const View<D, P...>& dst, const typename View<D, P...>::non_const_value_type& value) {
  using V = View<D, P...>;
  using T = typename V::non_const_value_type;
  const size_t n = dst.size();
  if (n == 0) return;
  T* p = const_cast<T*>(dst.data());
  if constexpr (V::is_strided) {  // fill element by element through the strides
    if (!dst.span_is_contiguous()) {
      const Impl::StridedBox<V::rank> box = Impl::make_strided_box(dst, dst);
      const T v = value;
      if constexpr (V::is_device) {
        parallel_for("kb200::ViewFill", RangePolicy<>(space, 0, (long long)n), KB200_LAMBDA(const long long i) {
          size_t od, os;
          box.offsets((size_t)i, od, os);
          p[od] = v;
        });
      } else {
        space.fence("kb200::deep_copy(host view, value)");
        for (size_t i = 0; i < n; ++i) { size_t od, os; box.offsets(i, od, os); p[od] = v; }
      }
      return;
    }
  }
  if constexpr (V::is_device) {
    bool all_zero = true;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(&value);
    for (size_t k = 0; k < sizeof(T); ++k) all_zero = all_zero && b[k] == 0;
    if (all_zero) {
      Impl::throw_on_error(b200_memset_async(space.impl_instance(), p, 0, n * sizeof(T)));
    } else if constexpr (std::is_same<T, double>::value) {
      Impl::throw_on_error(b200_stream_set_f64(space.impl_instance(), p, value, (int64_t)n));
    } else {
      const T v = value;
      parallel_for("kb200::ViewFill", RangePolicy<>(space, 0, (long long)n), KB200_LAMBDA(const long long i) { p[i] = v; });
    }
  } else {
    space.fence("kb200::deep_copy(host view, value)");
    for (size_t i = 0; i < n; ++i) p[i] = value;
  }
}
template <class D, class... P>
void deep_copy(const View<D, P...>& dst, const typename View<D, P...>::non_const_value_type& value) {
  B200 space;
  deep_copy(space, dst, value);
  space.fence("kb200::deep_copy(view, value): fence after fill");
}

namespace Experimental {
namespace Impl2 {
// (flag, value) pairs: scans / reductions with a user operator need no identity element of that operator
// (the reference's ValueWrapperForNoNeutralElement, algorithms/src/std_algorithms/impl/Kokkos_ValueWrapperForNoNeutralElement.hpp)
template <class T>
struct Wrapped { T val; int is_initial; };
template <class T, class Op>
struct WrappedJoin {
  using value_type = Wrapped<T>;
  Op op;
  KB200_INLINE_FUNCTION void init(value_type& w) const { w.val = T(); w.is_initial = 1; }
  KB200_INLINE_FUNCTION void join(value_type& d, const value_type& s) const {
    if (s.is_initial) return;
    if (d.is_initial) { d = s; return; }
    d.val = op(d.val, s.val);
  }
};
template <class VIn, class VOut, class T, class Op>
struct ExclScanOp : WrappedJoin<T, Op> {
  VIn in; VOut out; T seed;
  KB200_INLINE_FUNCTION void operator()(const long long i, Wrapped<T>& u, const bool fin) const {
    if (fin) out(i) = u.is_initial ? seed : this->op(seed, u.val);
    Wrapped<T> me{(T)in(i), 0};
    this->join(u, me);
  }
};
template <class VIn, class VOut, class T, class Op>
struct InclScanOp : WrappedJoin<T, Op> {
  VIn in; VOut out;
  KB200_INLINE_FUNCTION void operator()(const long long i, Wrapped<T>& u, const bool fin) const {
    Wrapped<T> me{(T)in(i), 0};
    this->join(u, me);
    if (fin) out(i) = u.val;
  }
};
template <class V, class T, class Op>
struct ReduceOp : WrappedJoin<T, Op> {
  V v;
  KB200_INLINE_FUNCTION void operator()(const long long i, Wrapped<T>& u) const { Wrapped<T> me{(T)v(i), 0}; this->join(u, me); }
};
template <class V, class T, class Op, class UnaryOp>
struct TransformReduceOp : WrappedJoin<T, Op> {
  V v; UnaryOp uop;
  KB200_INLINE_FUNCTION void operator()(const long long i, Wrapped<T>& u) const { Wrapped<T> me{(T)uop(v(i)), 0}; this->join(u, me); }
};
template <class T>
constexpr bool typed_scan_v = std::is_same<T, long long>::value || std::is_same<T, long>::value || std::is_same<T, double>::value ||
                              std::is_same<T, int>::value;
template <class VIn, class VOut>
void check_scan_views(const VIn& in, const VOut& out, const char* what) {
  static_assert(VIn::rank == 1 && VOut::rank == 1, "std algorithms: rank-1 Views only");
  if (out.extent(0) < in.extent(0)) throw std::runtime_error(std::string("kb200::Experimental::") + what + ": destination is shorter than the source");
}
// typed C-ABI scan for the default operator (plus) over contiguous arithmetic Views
template <class T>
int typed_scan(b200_instance* inst, bool inclusive, const T* x, T* y, int64_t n, T init) {
  if constexpr (std::is_same<T, double>::value) {
    return inclusive ? b200_scan_incl_f64(inst, x, y, n, init, nullptr, nullptr) : b200_scan_excl_f64(inst, x, y, n, init, nullptr, nullptr);
  } else if constexpr (sizeof(T) == 8) {
    return inclusive ? b200_scan_incl_i64(inst, (const int64_t*)x, (int64_t*)y, n, (int64_t)init, nullptr, nullptr)
                     : b200_scan_excl_i64(inst, (const int64_t*)x, (int64_t*)y, n, (int64_t)init, nullptr, nullptr);
  } else {
    static_assert(std::is_same<T, int>::value, "");
    if (!inclusive) return b200_scan_excl_i32(inst, x, y, n, init, nullptr, nullptr);
    return B200_EUNSUPPORTED;
  }
}
}  // namespace Impl2

// exclusive_scan(exec, in, out, init [, op]): out[i] = init (+) in[0] (+) ... (+) in[i-1]; returns the number of elements written
template <class VIn, class VOut, class T>
size_t exclusive_scan(const std::string& label, const B200& space, const VIn& in, const VOut& out, T init) {
  using TV = typename VOut::non_const_value_type;
  Impl2::check_scan_views(in, out, "exclusive_scan");
  const long long n = (long long)in.extent(0);
  if (n == 0) return 0;
  if constexpr (Impl2::typed_scan_v<TV> && std::is_same<typename VIn::non_const_value_type, TV>::value && VIn::is_device && VOut::is_device) {
    const int rc = Impl2::typed_scan<TV>(space.impl_instance(), false, in.data(), const_cast<TV*>(out.data()), n, (TV)init);
    if (rc != B200_EUNSUPPORTED) { Impl::throw_on_error(rc); return (size_t)n; }
  }
  const TV seed = (TV)init;
  TV total{};
  parallel_scan(label, RangePolicy<>(space, 0, n), KB200_LAMBDA(const long long i, TV& u, const bool fin) {
    if (fin) out(i) = seed + u;
    u += in(i);
  }, total);
  return (size_t)n;
}
template <class VIn, class VOut, class T>
size_t exclusive_scan(const B200& space, const VIn& in, const VOut& out, T init) {
  return exclusive_scan("Kokkos::exclusive_scan_default_functors_view_api", space, in, out, init);
}
// custom associative operator (value type must be trivially copyable; identity = `init` folded in front)
template <class VIn, class VOut, class T, class BinaryOp>
size_t exclusive_scan(const B200& space, const VIn& in, const VOut& out, T init, BinaryOp op) {
  using TV = typename VOut::non_const_value_type;
  Impl2::check_scan_views(in, out, "exclusive_scan");
  const long long n = (long long)in.extent(0);
  if (n == 0) return 0;
  Impl2::ExclScanOp<VIn, VOut, TV, BinaryOp> f;
  f.op = op; f.in = in; f.out = out; f.seed = (TV)init;
  parallel_scan("Kokkos::exclusive_scan_custom_functors_view_api", RangePolicy<>(space, 0, n), f);
  return (size_t)n;
}

// inclusive_scan(exec, in, out [, op [, init]])
template <class VIn, class VOut>
size_t inclusive_scan(const std::string& label, const B200& space, const VIn& in, const VOut& out) {
  using TV = typename VOut::non_const_value_type;
  Impl2::check_scan_views(in, out, "inclusive_scan");
  const long long n = (long long)in.extent(0);
  if (n == 0) return 0;
  if constexpr (Impl2::typed_scan_v<TV> && std::is_same<typename VIn::non_const_value_type, TV>::value && VIn::is_device && VOut::is_device) {
    const int rc = Impl2::typed_scan<TV>(space.impl_instance(), true, in.data(), const_cast<TV*>(out.data()), n, TV(0));
    if (rc != B200_EUNSUPPORTED) { Impl::throw_on_error(rc); return (size_t)n; }
  }
  TV total{};
  parallel_scan(label, RangePolicy<>(space, 0, n), KB200_LAMBDA(const long long i, TV& u, const bool fin) {
    u += in(i);
    if (fin) out(i) = u;
  }, total);
  return (size_t)n;
}
template <class VIn, class VOut>
size_t inclusive_scan(const B200& space, const VIn& in, const VOut& out) {
  return inclusive_scan("Kokkos::inclusive_scan_default_functors_view_api", space, in, out);
}
template <class VIn, class VOut, class BinaryOp>
size_t inclusive_scan(const B200& space, const VIn& in, const VOut& out, BinaryOp op) {
  using TV = typename VOut::non_const_value_type;
  Impl2::check_scan_views(in, out, "inclusive_scan");
  const long long n = (long long)in.extent(0);
  if (n == 0) return 0;
  Impl2::InclScanOp<VIn, VOut, TV, BinaryOp> f;
  f.op = op; f.in = in; f.out = out;
  parallel_scan("Kokkos::inclusive_scan_custom_functors_view_api", RangePolicy<>(space, 0, n), f);
  return (size_t)n;
}

// reduce(exec, view [, init [, op]])
template <class V, class T>
T reduce(const B200& space, const V& v, T init) {
  static_assert(V::rank == 1, "std algorithms: rank-1 Views only");
  using TV = typename V::non_const_value_type;
  const long long n = (long long)v.extent(0);
  if (n == 0) return init;
  if constexpr (V::is_device && std::is_same<TV, T>::value &&
                (std::is_same<T, double>::value || std::is_same<T, float>::value || std::is_same<T, int>::value || std::is_same<T, long long>::value ||
                 std::is_same<T, long>::value)) {
    T r{};
    int rc;
    if constexpr (std::is_same<T, double>::value) rc = b200_reduce_sum_f64(space.impl_instance(), v.data(), n, &r, nullptr);
    else if constexpr (std::is_same<T, float>::value) rc = b200_reduce_sum_f32(space.impl_instance(), v.data(), n, &r, nullptr);
    else if constexpr (std::is_same<T, int>::value) rc = b200_reduce_sum_i32(space.impl_instance(), v.data(), n, &r, nullptr);
    else rc = b200_reduce_sum_i64(space.impl_instance(), (const int64_t*)v.data(), n, (int64_t*)&r, nullptr);
    Impl::throw_on_error(rc);
    return init + r;
  } else {
    T r{};
    parallel_reduce("Kokkos::reduce_default_functors_view_api", RangePolicy<>(space, 0, n), KB200_LAMBDA(const long long i, T& u) { u += (T)v(i); }, r);
    return init + r;
  }
}
template <class V>
typename V::non_const_value_type reduce(const B200& space, const V& v) {
  return reduce(space, v, typename V::non_const_value_type{});
}
template <class V, class T, class BinaryOp>
T reduce(const B200& space, const V& v, T init, BinaryOp op) {
  static_assert(V::rank == 1, "std algorithms: rank-1 Views only");
  const long long n = (long long)v.extent(0);
  if (n == 0) return init;
  using W = Impl2::Wrapped<T>;
  Impl2::ReduceOp<V, T, BinaryOp> f;
  f.op = op; f.v = v;
  W r;
  parallel_reduce("Kokkos::reduce_custom_functors_view_api", RangePolicy<>(space, 0, n), f, r);
  return r.is_initial ? init : op(init, r.val);
}

// transform_reduce: inner product form and (join, unary transform) form
template <class V1, class V2, class T>
T transform_reduce(const B200& space, const V1& a, const V2& b, T init) {
  const long long n = (long long)a.extent(0);
  if ((long long)b.extent(0) < n) throw std::runtime_error("kb200::Experimental::transform_reduce: second view is shorter than the first");
  if (n == 0) return init;
  T r{};
  parallel_reduce("Kokkos::transform_reduce_default_functors_view_api", RangePolicy<>(space, 0, n),
                  KB200_LAMBDA(const long long i, T& u) { u += (T)a(i) * (T)b(i); }, r);
  return init + r;
}
template <class V, class T, class JoinOp, class UnaryOp>
T transform_reduce(const B200& space, const V& v, T init, JoinOp jop, UnaryOp uop) {
  const long long n = (long long)v.extent(0);
  if (n == 0) return init;
  using W = Impl2::Wrapped<T>;
  Impl2::TransformReduceOp<V, T, JoinOp, UnaryOp> f;
  f.op = jop; f.v = v; f.uop = uop;
  W r;
  parallel_reduce("Kokkos::transform_reduce_custom_functors_view_api", RangePolicy<>(space, 0, n), f, r);
  return r.is_initial ? init : jop(init, r.val);
}

// min_element / max_element: index of the FIRST smallest / largest element; extent on an empty view (= end())
template <class V>
size_t min_element(const B200& space, const V& v) {
  static_assert(V::rank == 1, "std algorithms: rank-1 Views only");
  using T = typename V::non_const_value_type;
  const long long n = (long long)v.extent(0);
  if (n == 0) return 0;
  using R = MinFirstLoc<T, long long>;
  typename R::value_type r;
  parallel_reduce("Kokkos::min_element_view_api_default", RangePolicy<>(space, 0, n),
                  KB200_LAMBDA(const long long i, typename R::value_type& u) {
                    const T x = v(i);
                    if (x < u.val) { u.val = x; u.loc = i; }
                  }, R(r));
  return (size_t)r.loc;
}
template <class V>
size_t max_element(const B200& space, const V& v) {
  static_assert(V::rank == 1, "std algorithms: rank-1 Views only");
  using T = typename V::non_const_value_type;
  const long long n = (long long)v.extent(0);
  if (n == 0) return 0;
  using R = MaxFirstLoc<T, long long>;
  typename R::value_type r;
  parallel_reduce("Kokkos::max_element_view_api_default", RangePolicy<>(space, 0, n),
                  KB200_LAMBDA(const long long i, typename R::value_type& u) {
                    const T x = v(i);
                    if (u.val < x) { u.val = x; u.loc = i; }
                  }, R(r));
  return (size_t)r.loc;
}
// minmax_element: (index of the first smallest, index of the LAST largest), as std::minmax_element
template <class V>
std::pair<size_t, size_t> minmax_element(const B200& space, const V& v) {
  static_assert(V::rank == 1, "std algorithms: rank-1 Views only");
  using T = typename V::non_const_value_type;
  const long long n = (long long)v.extent(0);
  if (n == 0) return {0, 0};
  using R = MinMaxFirstLastLoc<T, long long>;
  typename R::value_type r;
  parallel_reduce("Kokkos::minmax_element_view_api_default", RangePolicy<>(space, 0, n),
                  KB200_LAMBDA(const long long i, typename R::value_type& u) {
                    const T x = v(i);
                    if (x < u.min_val) { u.min_val = x; u.min_loc = i; }
                    if (x >= u.max_val) { u.max_val = x; u.max_loc = i; }   // >= : a thread visits its indices in increasing order
                  }, R(r));
  return {(size_t)r.min_loc, (size_t)r.max_loc};
}

// elementwise members
template <class V, class T>
void fill(const B200& space, const V& v, const T& value) { deep_copy(space, v, (typename V::non_const_value_type)value); }
template <class VIn, class VOut>
size_t copy(const B200& space, const VIn& in, const VOut& out) {
  Impl2::check_scan_views(in, out, "copy");
  const long long n = (long long)in.extent(0);
  parallel_for("Kokkos::copy_view_api_default", RangePolicy<>(space, 0, n), KB200_LAMBDA(const long long i) { out(i) = in(i); });
  return (size_t)n;
}
template <class VIn, class VOut, class UnaryOp>
size_t transform(const B200& space, const VIn& in, const VOut& out, UnaryOp op) {
  Impl2::check_scan_views(in, out, "transform");
  const long long n = (long long)in.extent(0);
  parallel_for("Kokkos::transform_view_api_default", RangePolicy<>(space, 0, n), KB200_LAMBDA(const long long i) { out(i) = op(in(i)); });
  return (size_t)n;
}
template <class V, class UnaryOp>
void for_each(const B200& space, const V& v, UnaryOp op) {
  const long long n = (long long)v.extent(0);
  parallel_for("Kokkos::for_each_view_api_default", RangePolicy<>(space, 0, n), KB200_LAMBDA(const long long i) { op(v(i)); });
}
template <class V, class Pred>
size_t count_if(const B200& space, const V& v, Pred pred) {
  const long long n = (long long)v.extent(0);
  long long c = 0;
  if (n) parallel_reduce("Kokkos::count_if_view_api_default", RangePolicy<>(space, 0, n), KB200_LAMBDA(const long long i, long long& u) { u += pred(v(i)) ? 1 : 0; }, c);
  return (size_t)c;
}
// find_if: index of the first element satisfying pred, extent(0) if none (FirstLoc reducer, as the reference's find_if)
template <class V, class Pred>
size_t find_if(const B200& space, const V& v, Pred pred) {
  const long long n = (long long)v.extent(0);
  if (n == 0) return 0;
  using R = FirstLoc<long long>;
  typename R::value_type r;
  parallel_reduce("Kokkos::find_if_view_api_default", RangePolicy<>(space, 0, n),
                  KB200_LAMBDA(const long long i, typename R::value_type& u) { if (pred(v(i)) && i < u.min_loc_true) u.min_loc_true = i; }, R(r));
  return r.min_loc_true == reduction_identity<long long>::min() ? (size_t)n : (size_t)r.min_loc_true;
}

}  // namespace Experimental

// ---------------------------------------------------------------- Crs row-map construction (core/src/Kokkos_Crs.hpp:165-208,293-370)
// get_crs_row_map_from_counts: row_map[i] = sum_{j<i} counts[j], row_map[n] = total; returns the total (number of entries).
// The reference runs a parallel_scan with a functor that also writes the last entry; here it is ONE exclusive scan
// (typed single-pass kernel for int32/int64 counts) whose total lands in row_map[n].
template <class RowMap, class Counts>
typename RowMap::non_const_value_type get_crs_row_map_from_counts(const B200& space, const RowMap& row_map, const Counts& counts) {
  using T = typename RowMap::non_const_value_type;
  const long long n = (long long)counts.extent(0);
  if ((long long)row_map.extent(0) != n + 1) throw std::runtime_error("kb200::get_crs_row_map_from_counts: row_map must have counts.extent(0)+1 entries");
  T total{};
  T* rm = const_cast<T*>(row_map.data());
  if constexpr (std::is_same<T, typename Counts::non_const_value_type>::value && (std::is_same<T, long long>::value || std::is_same<T, long>::value)) {
    Impl::throw_on_error(b200_scan_excl_i64(space.impl_instance(), (const int64_t*)counts.data(), (int64_t*)rm, n, 0, (int64_t*)&total, (int64_t*)(rm + n)));
  } else if constexpr (std::is_same<T, typename Counts::non_const_value_type>::value && std::is_same<T, int>::value) {
    Impl::throw_on_error(b200_scan_excl_i32(space.impl_instance(), counts.data(), rm, n, 0, &total, rm + n));
  } else {
    parallel_scan("Kokkos::get_crs_row_map_from_counts", RangePolicy<>(space, 0, n), KB200_LAMBDA(const long long i, T& u, const bool fin) {
      if (fin) { rm[i] = u; }
      u += (T)counts(i);
      if (fin && i == n - 1) rm[n] = u;
    }, total);
  }
  return total;
}

}  // namespace kb200
#endif
