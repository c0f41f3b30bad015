// Kokkos_B200_StdAlgorithms.hpp -- Kokkos::Experimental::{exclusive,inclusive}_scan on `Kokkos::B200` (optional header: include it
// instead of <Kokkos_StdAlgorithms.hpp>; needs <reference>/algorithms/src on the include path).
//
// The reference's std_algorithms express a prefix sum over Views as a parallel_scan with ONE fixed functor type per algorithm
// (algorithms/src/std_algorithms/impl/Kokkos_ExclusiveScan.hpp:37-92, Kokkos_InclusiveScan.hpp:85-106,190-230), and for its own
// Cuda backend short-cuts inclusive_scan to thrust/CUB (Kokkos_InclusiveScan.hpp:156-172).  Here the same hook is the functor
// type: Impl::ParallelScan specialised on those functors knows that the loop body is "out[i] = prefix (+ init); prefix += in[i]"
// over two contiguous rank-1 Views and runs the typed single-pass look-back kernel (kokkos_b200/include/kb200/impl/ScanContig.hpp:
// TMA bulk loads/stores through a shared-memory ring, 16 B/element) instead of the generic functor kernel.  Strided Views and
// value types without a typed kernel take the generic path, exactly as a user lambda would.
#ifndef KOKKOS_B200_STDALGORITHMS_HPP
#define KOKKOS_B200_STDALGORITHMS_HPP

#include <Kokkos_B200_Space.hpp>
#include <Kokkos_StdAlgorithms.hpp>
#include <kokkos_b200.h>

namespace Kokkos {
namespace Impl {
namespace B200Adapter {

// typed entry points of libkokkos_b200.so for a value type (nullptr: none)
template <class V> struct typed_scan { static constexpr bool available = false; };
#define KOKKOS_B200_TYPED_SCAN(TYPE, CTYPE, EXCL, INCL)                                                                          \
  template <> struct typed_scan<TYPE> {                                                                                         \
    static constexpr bool available = true;                                                                                     \
    static int excl(b200_instance* i, const TYPE* x, TYPE* y, int64_t n, TYPE seed) { return EXCL(i, (const CTYPE*)x, (CTYPE*)y, n, (CTYPE)seed, nullptr, nullptr); } \
    static int incl(b200_instance* i, const TYPE* x, TYPE* y, int64_t n, TYPE seed) { return INCL(i, (const CTYPE*)x, (CTYPE*)y, n, (CTYPE)seed, nullptr, nullptr); } \
  };
KOKKOS_B200_TYPED_SCAN(long, int64_t, b200_scan_excl_i64, b200_scan_incl_i64)
KOKKOS_B200_TYPED_SCAN(long long, int64_t, b200_scan_excl_i64, b200_scan_incl_i64)
KOKKOS_B200_TYPED_SCAN(unsigned long, int64_t, b200_scan_excl_i64, b200_scan_incl_i64)  // two's complement: same bits
KOKKOS_B200_TYPED_SCAN(unsigned long long, int64_t, b200_scan_excl_i64, b200_scan_incl_i64)
KOKKOS_B200_TYPED_SCAN(double, double, b200_scan_excl_f64, b200_scan_incl_f64)
#undef KOKKOS_B200_TYPED_SCAN

template <class It>
inline constexpr bool is_view_iterator = false;
template <class D, class... A>
inline constexpr bool is_view_iterator<Kokkos::Experimental::Impl::RandomAccessIterator<Kokkos::View<D, A...>>> = true;

// the body shared by the two specialisations below
template <bool Inclusive, class Policy, class Functor, class V, class In, class Out>
void std_scan_execute(const Policy& policy, const Functor& f, const In& from, const Out& dest, V init) {
  before_launch();
  using KP  = kb_range_policy<Policy>;
  using Red = kb200::Impl::FunctorReducer<Functor, V, typename Policy::work_tag>;
  if constexpr (typed_scan<V>::available && is_view_iterator<In> && is_view_iterator<Out> &&
                std::is_same_v<std::remove_const_t<typename In::value_type>, V> && std::is_same_v<typename Out::value_type, V>) {
    if (from.stride() == 1 && dest.stride() == 1 && policy.end() > policy.begin()) {
      const int64_t b = (int64_t)policy.begin(), n = (int64_t)policy.end() - b;
      b200_instance* inst = policy.space().impl_kb200().impl_instance();
      kb200::Impl::throw_on_error(Inclusive ? typed_scan<V>::incl(inst, (const V*)from.data() + b, (V*)dest.data() + b, n, init)
                                            : typed_scan<V>::excl(inst, (const V*)from.data() + b, (V*)dest.data() + b, n, init));
      return;
    }
  }
  kb200::Impl::throw_on_error(kb200::Impl::GenericScan<KP, Functor, Red>::run(to_kb(policy), f, Red{f}, (V*)nullptr, (V*)nullptr));
}
}  // namespace B200Adapter

template <class Index, class V, class In, class Out, class... Traits>
class ParallelScan<Kokkos::Experimental::Impl::ExclusiveScanDefaultFunctorForKnownNeutralElement<Kokkos::B200, Index, V, In, Out>,
                   Kokkos::RangePolicy<Traits...>, Kokkos::B200> {
 public:
  using FunctorType  = Kokkos::Experimental::Impl::ExclusiveScanDefaultFunctorForKnownNeutralElement<Kokkos::B200, Index, V, In, Out>;
  using Policy       = Kokkos::RangePolicy<Traits...>;
  using functor_type = FunctorType;
  ParallelScan(const FunctorType& arg_functor, const Policy& arg_policy) : m_functor(arg_functor), m_policy(arg_policy) {}
  Policy const& get_policy() const { return m_policy; }
  void execute() const {
    B200Adapter::std_scan_execute<false>(m_policy, m_functor, m_functor.m_first_from, m_functor.m_first_dest, m_functor.m_init_value);
  }

 private:
  const FunctorType m_functor;
  const Policy m_policy;
};

template <class Index, class V, class In, class Out, class... Traits>
class ParallelScan<Kokkos::Experimental::Impl::InclusiveScanDefaultFunctorForKnownIdentityElement<Kokkos::B200, Index, V, In, Out>,
                   Kokkos::RangePolicy<Traits...>, Kokkos::B200> {
 public:
  using FunctorType  = Kokkos::Experimental::Impl::InclusiveScanDefaultFunctorForKnownIdentityElement<Kokkos::B200, Index, V, In, Out>;
  using Policy       = Kokkos::RangePolicy<Traits...>;
  using functor_type = FunctorType;
  ParallelScan(const FunctorType& arg_functor, const Policy& arg_policy) : m_functor(arg_functor), m_policy(arg_policy) {}
  Policy const& get_policy() const { return m_policy; }
  void execute() const { B200Adapter::std_scan_execute<true>(m_policy, m_functor, m_functor.m_first_from, m_functor.m_first_dest, V()); }

 private:
  const FunctorType m_functor;
  const Policy m_policy;
};

}  // namespace Impl
}  // namespace Kokkos
#endif
