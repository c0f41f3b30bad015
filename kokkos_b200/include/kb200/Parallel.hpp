// kb200/Parallel.hpp -- front-end of the hot path: kb200::parallel_for / parallel_reduce / parallel_scan
// over RangePolicy and MDRangePolicy, generic in the user functor (compiled by nvcc in the user's TU).
//
// Same call forms as core/src/Kokkos_Parallel.hpp:130-173,348-454 and the parallel_reduce overload set of
// core/src/Kokkos_Parallel_Reduce.hpp:1679-1837: optional label, policy or plain count, functor, and for
// reductions a scalar reference / rank-0 View / reducer object.  Functor analysis follows
// core/src/impl/Kokkos_FunctorAnalysis.hpp:865-958: value_type from the result argument, optional
// functor members init / join / final, work tag passed as first argument.
//
// What executes is not the reference's design:
//   ParallelFor   -> impl/ForKernel.hpp     (UNROLL independent iterations per thread, persistent-free grid)
//   ParallelReduce-> impl/ReduceKernel.hpp  (register partials, shuffle/redux.sync, ticketed ordered combine,
//                                            result written to a mapped pinned slot)
//   ParallelScan  -> impl/ScanGeneric.hpp   (single pass, decoupled look-back; functor called exactly twice
//                                            per index: final=false for its contribution, final=true once)
//   MDRange       -> impl/MDRangeKernel.hpp (tile = thread block, hardware threadIdx gives the in-tile
//                                            coordinates: no div/mod per element)
// Scalar results block (fence) and View/device results do not, as in the reference
// (Kokkos_Parallel_Reduce.hpp:1592-1638).
#ifndef KB200_PARALLEL_HPP
#define KB200_PARALLEL_HPP

#include "Policy.hpp"
#include "Reducers.hpp"
#include "View.hpp"
#include "impl/ForKernel.hpp"
#include "impl/ReduceKernel.hpp"
#include "impl/ScanGeneric.hpp"
#include "impl/MDRangeKernel.hpp"
#include "impl/ArrayReduceKernel.hpp"

namespace kb200 {
namespace Impl {

// ---- invoke a functor with or without a work tag --------------------------------------------------
template <class Tag, class F, class... Args>
KB200_FORCEINLINE_FUNCTION void invoke(const F& f, Args&&... args) {
  if constexpr (std::is_void<Tag>::value) f(static_cast<Args&&>(args)...);
  else f(Tag{}, static_cast<Args&&>(args)...);
}

// ---- functor analysis: does the functor bring its own init / join / final ? -----------------------
template <class F, class V, class = void> struct has_join : std::false_type {};
template <class F, class V> struct has_join<F, V, std::void_t<decltype(std::declval<const F&>().join(std::declval<V&>(), std::declval<const V&>()))>> : std::true_type {};
template <class F, class V, class = void> struct has_init : std::false_type {};
template <class F, class V> struct has_init<F, V, std::void_t<decltype(std::declval<const F&>().init(std::declval<V&>()))>> : std::true_type {};
template <class F, class V, class = void> struct has_final : std::false_type {};
template <class F, class V> struct has_final<F, V, std::void_t<decltype(std::declval<const F&>().final(std::declval<V&>()))>> : std::true_type {};
template <class Tag, class F, class V, class = void> struct has_tagged_join : std::false_type {};
template <class Tag, class F, class V>
struct has_tagged_join<Tag, F, V, std::void_t<decltype(std::declval<const F&>().join(std::declval<Tag>(), std::declval<V&>(), std::declval<const V&>()))>> : std::true_type {};

template <class Tag, class F, class V, class = void> struct has_tagged_init : std::false_type {};
template <class Tag, class F, class V>
struct has_tagged_init<Tag, F, V, std::void_t<decltype(std::declval<const F&>().init(std::declval<Tag>(), std::declval<V&>()))>> : std::true_type {};
template <class Tag, class F, class V, class = void> struct has_tagged_final : std::false_type {};
template <class Tag, class F, class V>
struct has_tagged_final<Tag, F, V, std::void_t<decltype(std::declval<const F&>().final(std::declval<Tag>(), std::declval<V&>()))>> : std::true_type {};
// does the functor bring any of init / join / final (plain or tagged with the policy's work tag)?
template <class F, class V, class Tag, bool = std::is_void<Tag>::value>
struct brings_reduction_members : std::integral_constant<bool, has_join<F, V>::value || has_init<F, V>::value || has_final<F, V>::value> {};
template <class F, class V, class Tag>
struct brings_reduction_members<F, V, Tag, false>
    : std::integral_constant<bool, has_join<F, V>::value || has_init<F, V>::value || has_final<F, V>::value || has_tagged_join<Tag, F, V>::value ||
                                       has_tagged_init<Tag, F, V>::value || has_tagged_final<Tag, F, V>::value> {};

// Uniform init/join/final over "functor with optional members" (default: value-init and operator+=,
// FunctorAnalysis.hpp:604-613,724-732) -- the reducer used when the result argument is a plain scalar or View.
template <class F, class V, class Tag>
struct FunctorReducer {
  using value_type = V;
  F f;
  KB200_FORCEINLINE_FUNCTION void init(V& v) const {
    if constexpr (!std::is_void<Tag>::value && has_tagged_init<Tag, F, V>::value) f.init(Tag{}, v);
    else if constexpr (has_init<F, V>::value) f.init(v);
    else v = V();
  }
  KB200_FORCEINLINE_FUNCTION void join(V& d, const V& s) const {
    if constexpr (!std::is_void<Tag>::value && has_tagged_join<Tag, F, V>::value) f.join(Tag{}, d, s);
    else if constexpr (has_join<F, V>::value) f.join(d, s);
    else d += s;
  }
  KB200_FORCEINLINE_FUNCTION void final(V& v) const {
    if constexpr (!std::is_void<Tag>::value && has_tagged_final<Tag, F, V>::value) f.final(Tag{}, v);
    else if constexpr (has_final<F, V>::value) f.final(v);
  }
};
// plain-old-data fast case: no functor copy inside the reducer, and redux.sync for 32-bit integers
template <class V>
struct DefaultSumReducer {
  using value_type = V;
  static constexpr int redux_op = ReduxAdd;
  KB200_FORCEINLINE_FUNCTION void init(V& v) const { v = V(); }
  KB200_FORCEINLINE_FUNCTION void join(V& d, const V& s) const { d += s; }
  KB200_FORCEINLINE_FUNCTION void final(V&) const {}
};
// wrap a reducer object (built-in or user-defined: has ::reducer, value_type, init, join, maybe final)
template <class R>
struct ReducerAdapter {
  using value_type = typename R::value_type;
  R r;
  template <class Q, class = void> struct has_redux : std::false_type {};
  template <class Q> struct has_redux<Q, std::void_t<decltype(Q::redux_op)>> : std::true_type {};
  KB200_FORCEINLINE_FUNCTION void init(value_type& v) const { r.init(v); }
  KB200_FORCEINLINE_FUNCTION void join(value_type& d, const value_type& s) const { r.join(d, s); }
  KB200_FORCEINLINE_FUNCTION void final(value_type& v) const {
    if constexpr (has_final<R, value_type>::value) r.final(v);
  }
};
template <class R>
struct redux_op_of<ReducerAdapter<R>, void> { static constexpr int value = redux_op_of<R>::value; };

// ---- bodies adapting a Kokkos functor to the kernel skeletons -----------------------------------------
struct EmptyPacket {};
template <class F, class Tag, class Index>
struct FunctorForBody {
  using packet = EmptyPacket;
  F f;
  Index begin;
  KB200_DEVICE_FUNCTION packet load(int64) const { return packet{}; }
  KB200_DEVICE_FUNCTION void store(const packet&, int64 u) const { invoke<Tag>(f, (Index)(begin + (Index)u)); }
  KB200_FUNCTION int64 edge_count() const { return 0; }
  KB200_DEVICE_FUNCTION void edge(int64) const {}
};
template <class F, class Tag, class Index, class V>
struct FunctorReduceBody {
  using packet = EmptyPacket;
  F f;
  Index begin;
  KB200_DEVICE_FUNCTION packet load(int64) const { return packet{}; }
  KB200_DEVICE_FUNCTION void consume(const packet&, int64 u, V& acc) const { invoke<Tag>(f, (Index)(begin + (Index)u), acc); }
  KB200_FUNCTION int64 edge_count() const { return 0; }
  KB200_DEVICE_FUNCTION void edge(int64, V&) const {}
};

// ---- where does a reduction result go ----------------------------------------------------------------
template <class V>
struct ResultTarget { V* host; V* dev; };
template <class V>
ResultTarget<V> target_of_scalar(V& v) { return ResultTarget<V>{&v, nullptr}; }
template <class ViewT>
ResultTarget<typename ViewT::non_const_value_type> target_of_view(const ViewT& v) {
  using V = typename ViewT::non_const_value_type;
  if (ViewT::is_device) return ResultTarget<V>{nullptr, (V*)v.data()};
  return ResultTarget<V>{(V*)v.data(), nullptr};  // host View: written after a fence, like a scalar
}

}  // namespace Impl

// =====================================================================================================
// parallel_for
// =====================================================================================================
template <class... P, class F>
void parallel_for(const std::string& /*label*/, const RangePolicy<P...>& policy, const F& f) {
  using Policy = RangePolicy<P...>;
  using Body = Impl::FunctorForBody<F, typename Policy::work_tag, typename Policy::index_type>;
  const int64 n = (int64)(policy.end() - policy.begin());
  if (n <= 0) return;
  Body body{f, policy.begin()};
  // Static schedule: persistent grid (SMs x up to 8 resident CTAs) walking block-interleaved tiles; Dynamic: one CTA per tile so
  // the hardware scheduler balances uneven iterations.  B200 probe, copy lambda 2^28 (profiles/r01_for_probe.log): one element
  // per thread (the reference's mapping) 3.7 TB/s, 256x4 plain 5.59, 256x4 persistent 5.82.
  constexpr bool dynamic = std::is_same<typename Policy::schedule_type, Schedule<Dynamic>>::value;
  using Launch = Impl::RangeForLaunch<Body, 256, 4>;
  static const int static_cap = [] { int v = 8; b200_tune_get("for.bps", &v); return v; }();  // probe knobs, read once
  static const int waves = [] { int v = 16; b200_tune_get("for.waves", &v); return v; }();
  int cap = dynamic ? 0 : static_cap;
  if constexpr (Policy::experimental_contains_desired_occupancy) {  // Experimental::prefer(policy, DesiredOccupancy{p})
    cap = policy.impl_occupancy_cap(Launch::resident_blocks_per_sm());
    if (!dynamic && cap > 8) cap = 8;
  }
  Impl::throw_on_error(Launch::run(policy.space().impl_instance(), body, n, cap, waves));
}
template <class... P, class F>
void parallel_for(const RangePolicy<P...>& policy, const F& f) { parallel_for(std::string(), policy, f); }
template <class F, class = std::enable_if_t<!std::is_class<std::decay_t<F>>::value || true>>
void parallel_for(const std::string& label, size_t n, const F& f) { parallel_for(label, RangePolicy<>(0, (long long)n), f); }
template <class F>
void parallel_for(size_t n, const F& f) { parallel_for(std::string(), RangePolicy<>(0, (long long)n), f); }

template <class... P, class F>
void parallel_for(const std::string& /*label*/, const MDRangePolicy<P...>& policy, const F& f) {
  Impl::throw_on_error(Impl::MDRangeFor<MDRangePolicy<P...>, F>::run(policy, f));
}
template <class... P, class F>
void parallel_for(const MDRangePolicy<P...>& policy, const F& f) { parallel_for(std::string(), policy, f); }

// =====================================================================================================
// parallel_reduce
// =====================================================================================================
namespace Impl {
template <class Policy, class F, class Red>
void reduce_dispatch(const Policy& policy, const F& f, const Red& red, ResultTarget<typename Red::value_type> t);

template <class... P, class F, class Red>
void reduce_dispatch(const RangePolicy<P...>& policy, const F& f, const Red& red, ResultTarget<typename Red::value_type> t) {
  using Policy = RangePolicy<P...>;
  using V = typename Red::value_type;
  using Body = FunctorReduceBody<F, typename Policy::work_tag, typename Policy::index_type, V>;
  int64 n = (int64)(policy.end() - policy.begin());
  if (n < 0) n = 0;
  Body body{f, policy.begin()};
  // wider unroll for small values: more independent loads in flight per thread
  constexpr int UNROLL = sizeof(V) <= 8 ? 8 : (sizeof(V) <= 32 ? 4 : 2);
  // __launch_bounds__(256, 4): caps the kernel at 64 registers so >= 1024 threads stay resident per SM.  Without it the
  // 32-byte MinMaxLoc kernels took 142 registers (shuffle trees of the epilogue), one CTA per SM, 12 % occupancy and
  // 3.2 TB/s (profiles/r01_reduce_minmaxloc_ncu.txt); the hot loop itself needs < 48.
  using Launch = RangeReduceLaunch<Body, Red, 256, UNROLL, 4>;
  int cap = 0;
  if constexpr (Policy::experimental_contains_desired_occupancy) cap = policy.impl_occupancy_cap(Launch::resident_blocks_per_sm());
  throw_on_error(Launch::run(policy.space().impl_instance(), body, red, n, t.host, t.dev, cap));
}
template <class... P, class F, class Red>
void reduce_dispatch(const MDRangePolicy<P...>& policy, const F& f, const Red& red, ResultTarget<typename Red::value_type> t) {
  throw_on_error(MDRangeReduce<MDRangePolicy<P...>, F, Red>::run(policy, f, red, t.host, t.dev));
}

template <class T> struct is_policy : std::false_type {};
template <class... P> struct is_policy<RangePolicy<P...>> : std::true_type {};
template <class... P> struct is_policy<MDRangePolicy<P...>> : std::true_type {};
template <class... P> struct is_policy<TeamPolicy<P...>> : std::true_type {};

// ---- runtime-length array reductions: functor::value_type = T[] + value_count (row a4) -----------------------------------
template <class F, class = void> struct is_array_reduce_functor : std::false_type {};
template <class F>
struct is_array_reduce_functor<F, std::void_t<typename F::value_type>>
    : std::integral_constant<bool, std::is_array<typename F::value_type>::value && std::extent<typename F::value_type>::value == 0> {};

template <class T, class Tag, class F, class... P, int CAP>
int array_reduce_launch(const RangePolicy<P...>& policy, const F& f, int count, T* rh, T* rd, std::integral_constant<int, CAP>) {
  using Index = typename RangePolicy<P...>::index_type;
  constexpr int BLOCK = 256;
  b200_instance* inst = policy.space().impl_instance();
  HostRuntime rt(inst);
  const int64 n = (int64)(policy.end() - policy.begin()) > 0 ? (int64)(policy.end() - policy.begin()) : 0;
  auto k = array_range_reduce_kernel<F, Tag, Index, T, CAP, BLOCK>;
  int bps = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, BLOCK, 0);
  if (bps < 1) bps = 1;
  const int64 blocks = (n + BLOCK - 1) / BLOCK, cap = (int64)rt.sm_count() * bps;
  const int grid = (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
  return array_reduce_run<T>(inst, count, grid, rh, rd, [&](T* partials, unsigned* ticket, T* dst) {
    k<<<grid, BLOCK, 0, rt.stream()>>>(f, policy.begin(), n, count, partials, ticket, dst);
  });
}
template <class T, class Tag, class F, class... P, int CAP>
int array_reduce_launch(const MDRangePolicy<P...>& policy_in, const F& f, int count, T* rh, T* rd, std::integral_constant<int, CAP>) {
  using Policy = MDRangePolicy<P...>;
  using Index = typename Policy::index_type;
  Policy policy(policy_in);
  b200_instance* inst = policy.space().impl_instance();
  HostRuntime rt(inst);
  auto k = array_mdrange_reduce_kernel<F, Tag, T, CAP, Policy::rank, Index>;
  for (;;) {  // per-thread accumulator arrays are register-hungry: shrink a default tile until the kernel fits
    int fit = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, k, MDLaunchShape<Policy>(policy).threads, 0);
    if (fit >= 1 || !policy.impl_shrink_default_tile()) break;
  }
  MDLaunchShape<Policy> sh(policy);
  int bps = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, sh.threads, 0);
  if (bps < 1) bps = 1;
  const long long cap = (long long)rt.sm_count() * bps, tiles = sh.p.num_tiles;
  const int grid = (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
  sh.set_grid(grid);
  if (tiles < 1) { sh.p.num_tiles = 0; sh.block = dim3(32, 1, 1); }
  return array_reduce_run<T>(inst, count, grid, rh, rd, [&](T* partials, unsigned* ticket, T* dst) {
    k<<<grid, sh.block, 0, rt.stream()>>>(f, sh.p, count, partials, ticket, dst);
  });
}
template <class T, class Tag, class F, class... P, int CAP>
int array_reduce_launch(const TeamPolicy<P...>& pol, const F& f, int count, T* rh, T* rd, std::integral_constant<int, CAP>);  // Team.hpp

// value_count > 64: accumulator arrays in global memory (ArrayReduceKernel.hpp, "big" kernels)
template <class T, class Tag, class F, class... P>
int array_reduce_launch_big(const RangePolicy<P...>& policy, const F& f, int count, T* rh, T* rd) {
  using Index = typename RangePolicy<P...>::index_type;
  constexpr int BLOCK = 128;
  b200_instance* inst = policy.space().impl_instance();
  HostRuntime rt(inst);
  const int64 n = (int64)(policy.end() - policy.begin()) > 0 ? (int64)(policy.end() - policy.begin()) : 0;
  const int64 blocks = (n + BLOCK - 1) / BLOCK, cap = (int64)rt.sm_count() * 2;
  const int grid = array_big_grid(blocks < cap ? blocks : cap, BLOCK, count, sizeof(T));
  return array_reduce_big_run<T>(inst, count, grid, BLOCK, rh, rd, [&](T* slabs, unsigned* ticket, T* dst) {
    array_range_reduce_big_kernel<F, Tag, Index, T><<<grid, BLOCK, 0, rt.stream()>>>(f, policy.begin(), n, count, slabs, ticket, dst);
  });
}
template <class T, class Tag, class F, class... P>
int array_reduce_launch_big(const MDRangePolicy<P...>& policy, const F& f, int count, T* rh, T* rd) {
  using Policy = MDRangePolicy<P...>;
  using Index = typename Policy::index_type;
  b200_instance* inst = policy.space().impl_instance();
  HostRuntime rt(inst);
  MDLaunchShape<Policy> sh(policy);
  const long long cap = (long long)rt.sm_count() * 2, tiles = sh.p.num_tiles;
  const int grid = array_big_grid(tiles < cap ? tiles : cap, sh.threads, count, sizeof(T));
  sh.set_grid(grid);
  if (tiles < 1) { sh.p.num_tiles = 0; sh.block = dim3(32, 1, 1); sh.threads = 32; }
  return array_reduce_big_run<T>(inst, count, grid, sh.threads, rh, rd, [&](T* slabs, unsigned* ticket, T* dst) {
    array_mdrange_reduce_big_kernel<F, Tag, T, Policy::rank, Index><<<grid, sh.block, 0, rt.stream()>>>(f, sh.p, count, slabs, ticket, dst);
  });
}
template <class T, class Tag, class F, class... P>
int array_reduce_launch_big(const TeamPolicy<P...>&, const F&, int, T*, T*) {
  return b200_report_error(B200_EUNSUPPORTED, "kb200::parallel_reduce(TeamPolicy, value_type[]): value_count above 64 is not supported");
}

template <class Policy, class F, class R>
void array_reduce_entry(const Policy& policy, const F& f, R&& result) {
  using T = std::remove_extent_t<typename F::value_type>;
  using Tag = typename Policy::work_tag;
  using RD = std::decay_t<R>;
  const int count = (int)f.value_count;
  T *rh = nullptr, *rd = nullptr;
  if constexpr (is_view_v<RD>) {
    if ((int)result.size() < count) throw std::runtime_error("kb200::parallel_reduce(value_type[]): result View is shorter than value_count");
    if (RD::is_device) rd = (T*)result.data(); else rh = (T*)result.data();
  } else {
    rh = &result[0];  // C array or pointer on the host
  }
  int rc;
  if (count <= 8) rc = array_reduce_launch<T, Tag>(policy, f, count, rh, rd, std::integral_constant<int, 8>{});
  else if (count <= 32) rc = array_reduce_launch<T, Tag>(policy, f, count, rh, rd, std::integral_constant<int, 32>{});
  else if (count <= 64) rc = array_reduce_launch<T, Tag>(policy, f, count, rh, rd, std::integral_constant<int, 64>{});
  else rc = array_reduce_launch_big<T, Tag>(policy, f, count, rh, rd);
  throw_on_error(rc);
}

// where a reducer's result lives: the built-in reducers know (a scalar or a host View vs a device View); a user-written
// reducer declares it through the memory space of its result_view_type (core/unit_test/TestReduceCombinatorical.hpp:27-56)
template <class R, class = void> struct has_references_scalar : std::false_type {};
template <class R> struct has_references_scalar<R, std::void_t<decltype(std::declval<const R&>().references_scalar())>> : std::true_type {};
template <class R>
bool reducer_result_on_host(const R& r) {
  if constexpr (has_references_scalar<R>::value) return r.references_scalar();
  else return !R::result_view_type::is_device;
}

template <class Policy, class F, class R>
void reduce_entry(const Policy& policy, const F& f, R&& result) {
  using RD = std::decay_t<R>;
  using Tag = typename Policy::work_tag;
  if constexpr (is_array_reduce_functor<F>::value) {
    array_reduce_entry(policy, f, static_cast<R&&>(result));
    return;
  } else
  if constexpr (is_reducer_v<RD>) {
    using V = typename RD::value_type;
    ReducerAdapter<RD> red{result};
    ResultTarget<V> t = reducer_result_on_host(result) ? ResultTarget<V>{&result.reference(), nullptr}
                                                       : ResultTarget<V>{nullptr, &result.reference()};
    reduce_dispatch(policy, f, red, t);
  } else if constexpr (is_view_v<RD>) {
    using V = typename RD::non_const_value_type;
    static_assert(RD::rank == 0, "parallel_reduce: View results must be rank 0 (array reductions are not on this path)");
    if constexpr (brings_reduction_members<F, V, Tag>::value) {
      reduce_dispatch(policy, f, FunctorReducer<F, V, Tag>{f}, target_of_view(result));
    } else {
      reduce_dispatch(policy, f, DefaultSumReducer<V>{}, target_of_view(result));
    }
  } else {
    using V = RD;
    static_assert(!std::is_const<std::remove_reference_t<R>>::value, "parallel_reduce: result must be a non-const reference");
    if constexpr (brings_reduction_members<F, V, Tag>::value) {
      reduce_dispatch(policy, f, FunctorReducer<F, V, Tag>{f}, target_of_scalar(result));
    } else {
      reduce_dispatch(policy, f, DefaultSumReducer<V>{}, target_of_scalar(result));
    }
  }
}
}  // namespace Impl

template <class Policy, class F, class R, class = std::enable_if_t<Impl::is_policy<Policy>::value>>
void parallel_reduce(const std::string& /*label*/, const Policy& policy, const F& f, R&& result) {
  Impl::reduce_entry(policy, f, static_cast<R&&>(result));
}
template <class Policy, class F, class R, class = std::enable_if_t<Impl::is_policy<Policy>::value>>
void parallel_reduce(const Policy& policy, const F& f, R&& result) {
  Impl::reduce_entry(policy, f, static_cast<R&&>(result));
}
template <class F, class R>
void parallel_reduce(const std::string& /*label*/, size_t n, const F& f, R&& result) {
  Impl::reduce_entry(RangePolicy<>(0, (long long)n), f, static_cast<R&&>(result));
}
template <class F, class R>
void parallel_reduce(size_t n, const F& f, R&& result) {
  Impl::reduce_entry(RangePolicy<>(0, (long long)n), f, static_cast<R&&>(result));
}

// ---- no result argument: the functor's final() consumes the value on the device (Kokkos_Parallel_Reduce.hpp:1555-1640;
// core/unit_test/TestReduceCombinatorical.hpp:460-490).  The value type is F::value_type, else the last parameter of the
// (non-template) call operator, as the reference's FunctorAnalysis deduces it.
namespace Impl {
template <class F>
struct reduce_op_value {
  template <class C, class R, class... A>
  static std::remove_reference_t<std::tuple_element_t<sizeof...(A) - 1, std::tuple<A...>>> pick(R (C::*)(A...) const);
  using type = decltype(pick(&F::operator()));
};
template <class F, class = void> struct reduce_value_type_of { using type = typename reduce_op_value<F>::type; };
template <class F> struct reduce_value_type_of<F, std::void_t<typename F::value_type>> { using type = typename F::value_type; };
template <class Policy, class F>
void reduce_entry_no_result(const Policy& policy, const F& f) {
  using V = typename reduce_value_type_of<F>::type;
  static_assert(!std::is_array<V>::value, "parallel_reduce without a result argument: array value types need a result");
  V unused{};  // handing it back costs one stream sync; final() has already run on the device by then
  reduce_entry(policy, f, unused);
}
}  // namespace Impl
template <class Policy, class F, class = std::enable_if_t<Impl::is_policy<Policy>::value>>
void parallel_reduce(const std::string& /*label*/, const Policy& policy, const F& f) { Impl::reduce_entry_no_result(policy, f); }
template <class Policy, class F, class = std::enable_if_t<Impl::is_policy<Policy>::value>>
void parallel_reduce(const Policy& policy, const F& f) { Impl::reduce_entry_no_result(policy, f); }
template <class F, class = std::enable_if_t<!Impl::is_policy<F>::value && std::is_class<F>::value>>
void parallel_reduce(const std::string& /*label*/, size_t n, const F& f) { Impl::reduce_entry_no_result(RangePolicy<>(0, (long long)n), f); }
template <class F, class = std::enable_if_t<std::is_class<F>::value>>
void parallel_reduce(size_t n, const F& f) { Impl::reduce_entry_no_result(RangePolicy<>(0, (long long)n), f); }

// =====================================================================================================
// parallel_reduce with several results: parallel_reduce(label, policy, f, r0, r1, ...)   (row a22)
// =====================================================================================================
// The reference packs N reducers into one struct value_type and funnels it through the ordinary reduction
// (core/src/impl/Kokkos_Combined_Reducer.hpp:273-288,389-404,533-581).  Same here: the values travel as one trivially
// copyable CombinedValue through the register/shuffle/ticket machinery; the functor is called as f(i..., v0, v1, ...).
// Each result may be a scalar reference or a rank-0 View (summed), or any reducer object.  The call blocks, then each
// component is written to its destination (the reference fences as well, :567-580).
namespace Impl {
template <class... Vs> struct CombinedValue;
template <> struct CombinedValue<> {};
template <class V0, class... Vs>
struct CombinedValue<V0, Vs...> { V0 head; CombinedValue<Vs...> tail; };
template <int I, class V0, class... Vs>
KB200_FORCEINLINE_FUNCTION auto& combined_get(CombinedValue<V0, Vs...>& v) {
  if constexpr (I == 0) return v.head; else return combined_get<I - 1>(v.tail);
}
template <int I, class V0, class... Vs>
KB200_FORCEINLINE_FUNCTION const auto& combined_get(const CombinedValue<V0, Vs...>& v) {
  if constexpr (I == 0) return v.head; else return combined_get<I - 1>(v.tail);
}

// how one result argument takes part: value type, reducer, and where the value goes afterwards
template <class R, class = void>
struct combined_slot {  // plain scalar reference: summed
  using value_type = std::remove_reference_t<R>;
  using reducer_type = DefaultSumReducer<value_type>;
  KB200_INLINE_FUNCTION static reducer_type reducer(R&) { return {}; }
  static void store(const B200&, R& dst, const value_type& v) { dst = v; }
};
template <class R>
struct combined_slot<R, std::enable_if_t<is_view_v<R>>> {  // rank-0 View: summed
  using VT = std::decay_t<R>;
  using value_type = typename VT::non_const_value_type;
  using reducer_type = DefaultSumReducer<value_type>;
  KB200_INLINE_FUNCTION static reducer_type reducer(const VT&) { return {}; }
  static void store(const B200& space, const VT& dst, const value_type& v) {
    copy_bytes<typename VT::memory_space, HostSpace>(space, (void*)dst.data(), &v, sizeof(value_type));
  }
};
template <class R>
struct combined_slot<R, std::enable_if_t<is_reducer_v<std::decay_t<R>>>> {  // reducer object
  using RD = std::decay_t<R>;
  using value_type = typename RD::value_type;
  using reducer_type = ReducerAdapter<RD>;
  KB200_INLINE_FUNCTION static reducer_type reducer(const RD& r) { return reducer_type{r}; }
  static void store(const B200& space, const RD& r, const value_type& v) {
    if (reducer_result_on_host(r)) r.reference() = v;
    else throw_on_error(b200_memcpy_h2d_async(space.impl_instance(), (void*)r.data(), &v, sizeof(value_type)));
  }
};

template <class... Rs> struct CombinedReducers;
template <> struct CombinedReducers<> {
  template <class CV> KB200_FORCEINLINE_FUNCTION void init(CV&) const {}
  template <class CV> KB200_FORCEINLINE_FUNCTION void join(CV&, const CV&) const {}
  template <class CV> KB200_FORCEINLINE_FUNCTION void final(CV&) const {}
};
template <class R0, class... Rs>
struct CombinedReducers<R0, Rs...> {
  R0 head;
  CombinedReducers<Rs...> tail;
  template <class CV> KB200_FORCEINLINE_FUNCTION void init(CV& v) const { head.init(v.head); tail.init(v.tail); }
  template <class CV> KB200_FORCEINLINE_FUNCTION void join(CV& d, const CV& s) const { head.join(d.head, s.head); tail.join(d.tail, s.tail); }
  template <class CV> KB200_FORCEINLINE_FUNCTION void final(CV& v) const { head.final(v.head); tail.final(v.tail); }
};
template <class CV, class... Rs>
struct CombinedReducer {
  using value_type = CV;
  CombinedReducers<Rs...> rs;
  KB200_FORCEINLINE_FUNCTION void init(CV& v) const { rs.init(v); }
  KB200_FORCEINLINE_FUNCTION void join(CV& d, const CV& s) const { rs.join(d, s); }
  KB200_FORCEINLINE_FUNCTION void final(CV& v) const { rs.final(v); }
};
KB200_INLINE_FUNCTION CombinedReducers<> make_combined_reducers() { return {}; }
template <class R0, class... Rs>
KB200_INLINE_FUNCTION CombinedReducers<R0, Rs...> make_combined_reducers(const R0& r0, const Rs&... rs) { return CombinedReducers<R0, Rs...>{r0, make_combined_reducers(rs...)}; }

// calls f(leading args..., v0, v1, ...): leading args are the index / indices / team handle (and the work tag)
template <class F, class CV, int N>
struct CombinedFunctor {
  F f;
  template <class... Lead, size_t... Is>
  KB200_FORCEINLINE_FUNCTION void call(CV& v, std::index_sequence<Is...>, Lead&&... lead) const {
    f(static_cast<Lead&&>(lead)..., combined_get<(int)Is>(v)...);
  }
  template <class A0> KB200_FORCEINLINE_FUNCTION void operator()(A0&& a0, CV& v) const { call(v, std::make_index_sequence<N>{}, static_cast<A0&&>(a0)); }
  template <class A0, class A1> KB200_FORCEINLINE_FUNCTION void operator()(A0&& a0, A1&& a1, CV& v) const {
    call(v, std::make_index_sequence<N>{}, static_cast<A0&&>(a0), static_cast<A1&&>(a1)); }
  template <class A0, class A1, class A2> KB200_FORCEINLINE_FUNCTION void operator()(A0&& a0, A1&& a1, A2&& a2, CV& v) const {
    call(v, std::make_index_sequence<N>{}, static_cast<A0&&>(a0), static_cast<A1&&>(a1), static_cast<A2&&>(a2)); }
  template <class A0, class A1, class A2, class A3> KB200_FORCEINLINE_FUNCTION void operator()(A0&& a0, A1&& a1, A2&& a2, A3&& a3, CV& v) const {
    call(v, std::make_index_sequence<N>{}, static_cast<A0&&>(a0), static_cast<A1&&>(a1), static_cast<A2&&>(a2), static_cast<A3&&>(a3)); }
  template <class A0, class A1, class A2, class A3, class A4> KB200_FORCEINLINE_FUNCTION void operator()(A0&& a0, A1&& a1, A2&& a2, A3&& a3, A4&& a4, CV& v) const {
    call(v, std::make_index_sequence<N>{}, static_cast<A0&&>(a0), static_cast<A1&&>(a1), static_cast<A2&&>(a2), static_cast<A3&&>(a3), static_cast<A4&&>(a4)); }
  template <class A0, class A1, class A2, class A3, class A4, class A5> KB200_FORCEINLINE_FUNCTION void operator()(A0&& a0, A1&& a1, A2&& a2, A3&& a3, A4&& a4, A5&& a5, CV& v) const {
    call(v, std::make_index_sequence<N>{}, static_cast<A0&&>(a0), static_cast<A1&&>(a1), static_cast<A2&&>(a2), static_cast<A3&&>(a3), static_cast<A4&&>(a4), static_cast<A5&&>(a5)); }
  template <class A0, class A1, class A2, class A3, class A4, class A5, class A6> KB200_FORCEINLINE_FUNCTION void operator()(A0&& a0, A1&& a1, A2&& a2, A3&& a3, A4&& a4, A5&& a5, A6&& a6, CV& v) const {
    call(v, std::make_index_sequence<N>{}, static_cast<A0&&>(a0), static_cast<A1&&>(a1), static_cast<A2&&>(a2), static_cast<A3&&>(a3), static_cast<A4&&>(a4), static_cast<A5&&>(a5), static_cast<A6&&>(a6)); }
};

template <class CV, class... Rs, size_t... Is>
void combined_store(const B200& space, const CV& v, std::index_sequence<Is...>, Rs&&... rs) {
  (combined_slot<Rs>::store(space, rs, combined_get<(int)Is>(v)), ...);
}
template <class Policy, class F, class... Rs>
void combined_reduce_entry(const Policy& policy, const F& f, Rs&&... rs) {
  using CV = CombinedValue<typename combined_slot<Rs>::value_type...>;
  using Red = CombinedReducer<CV, typename combined_slot<Rs>::reducer_type...>;
  Red red{make_combined_reducers(combined_slot<Rs>::reducer(rs)...)};
  CombinedFunctor<F, CV, (int)sizeof...(Rs)> cf{f};
  CV result;
  red.init(result);
  reduce_dispatch(policy, cf, red, ResultTarget<CV>{&result, nullptr});
  combined_store(policy.space(), result, std::index_sequence_for<Rs...>{}, static_cast<Rs&&>(rs)...);
  policy.space().fence("kb200::parallel_reduce(combined): results delivered");
}
}  // namespace Impl

template <class Policy, class F, class R0, class R1, class... Rs, class = std::enable_if_t<Impl::is_policy<Policy>::value>>
void parallel_reduce(const std::string& /*label*/, const Policy& policy, const F& f, R0&& r0, R1&& r1, Rs&&... rs) {
  Impl::combined_reduce_entry(policy, f, static_cast<R0&&>(r0), static_cast<R1&&>(r1), static_cast<Rs&&>(rs)...);
}
template <class Policy, class F, class R0, class R1, class... Rs, class = std::enable_if_t<Impl::is_policy<Policy>::value>>
void parallel_reduce(const Policy& policy, const F& f, R0&& r0, R1&& r1, Rs&&... rs) {
  Impl::combined_reduce_entry(policy, f, static_cast<R0&&>(r0), static_cast<R1&&>(r1), static_cast<Rs&&>(rs)...);
}
template <class F, class R0, class R1, class... Rs>
void parallel_reduce(const std::string& /*label*/, size_t n, const F& f, R0&& r0, R1&& r1, Rs&&... rs) {
  Impl::combined_reduce_entry(RangePolicy<>(0, (long long)n), f, static_cast<R0&&>(r0), static_cast<R1&&>(r1), static_cast<Rs&&>(rs)...);
}
template <class F, class R0, class R1, class... Rs>
void parallel_reduce(size_t n, const F& f, R0&& r0, R1&& r1, Rs&&... rs) {
  Impl::combined_reduce_entry(RangePolicy<>(0, (long long)n), f, static_cast<R0&&>(r0), static_cast<R1&&>(r1), static_cast<Rs&&>(rs)...);
}

// =====================================================================================================
// parallel_scan (RangePolicy)
// =====================================================================================================
namespace Impl {
// value type of a nested-scan lambda f(i, T& partial, bool final): deduced from its call operator
template <class L, class I>
struct scan_arg_of {
  template <class C, class R, class A0, class A1, class A2> static A1 pick(R (C::*)(A0, A1, A2) const);
  template <class C, class R, class A0, class A1, class A2> static A1 pick(R (C::*)(A0, A1, A2));
  using type = decltype(pick(&L::operator()));
};
// value type of a scan functor used without a total: F::value_type if published, else deduced from the (non-template)
// call operator's second parameter, as the reference's FunctorAnalysis does (impl/Kokkos_FunctorAnalysis.hpp:77-160)
template <class F>
struct scan_op_value {
  template <class C, class R, class A0, class A1, class A2> static std::remove_reference_t<A1> pick(R (C::*)(A0, A1, A2) const);
  template <class C, class R, class T, class A0, class A1, class A2> static std::remove_reference_t<A1> pick(R (C::*)(T, A0, A1, A2) const);
  using type = decltype(pick(&F::operator()));
};
template <class F, class = void> struct scan_value_type_of { using type = typename scan_op_value<F>::type; };
template <class F> struct scan_value_type_of<F, std::void_t<typename F::value_type>> { using type = typename F::value_type; };
}  // namespace Impl

// value type deduced from the total argument
template <class... P, class F, class V, class = std::enable_if_t<!is_view_v<V>>>
void parallel_scan(const std::string& /*label*/, const RangePolicy<P...>& policy, const F& f, V& total) {
  using Policy = RangePolicy<P...>;
  using Red = Impl::FunctorReducer<F, V, typename Policy::work_tag>;
  Impl::throw_on_error(Impl::GenericScan<Policy, F, Red>::run(policy, f, Red{f}, &total, nullptr));
}
template <class... P, class F, class VT, class = std::enable_if_t<is_view_v<VT>>, class = void>
void parallel_scan(const std::string& /*label*/, const RangePolicy<P...>& policy, const F& f, const VT& total_view) {
  using Policy = RangePolicy<P...>;
  using V = typename VT::non_const_value_type;
  using Red = Impl::FunctorReducer<F, V, typename Policy::work_tag>;
  auto t = Impl::target_of_view(total_view);
  Impl::throw_on_error(Impl::GenericScan<Policy, F, Red>::run(policy, f, Red{f}, t.host, t.dev));
}
// no total: the functor must publish its value_type (as in the reference, Kokkos_Parallel.hpp:348-367)
template <class... P, class F>
void parallel_scan(const std::string& /*label*/, const RangePolicy<P...>& policy, const F& f) {
  using Policy = RangePolicy<P...>;
  using V = typename Impl::scan_value_type_of<F>::type;
  using Red = Impl::FunctorReducer<F, V, typename Policy::work_tag>;
  Impl::throw_on_error(Impl::GenericScan<Policy, F, Red>::run(policy, f, Red{f}, (V*)nullptr, (V*)nullptr));
}
template <class... P, class F, class... R>
void parallel_scan(const RangePolicy<P...>& policy, const F& f, R&&... r) { parallel_scan(std::string(), policy, f, static_cast<R&&>(r)...); }
template <class F, class... R>
void parallel_scan(const std::string& label, size_t n, const F& f, R&&... r) { parallel_scan(label, RangePolicy<>(0, (long long)n), f, static_cast<R&&>(r)...); }
template <class F, class... R>
void parallel_scan(size_t n, const F& f, R&&... r) { parallel_scan(std::string(), RangePolicy<>(0, (long long)n), f, static_cast<R&&>(r)...); }

}  // namespace kb200
#endif
