// kb200/Compat.hpp -- small pieces of the Kokkos Core surface that user code and the reference's own unit tests
// (core/unit_test/incremental/*.hpp, built by tests/ref_unit) touch next to the hot path.
//   kokkos_malloc / kokkos_free / kokkos_realloc   core/src/Kokkos_Core.hpp:155-200 (raw allocations in a memory space)
//   ExecutionSpace::memory_space / device_type helpers, Kokkos::Timer, Kokkos::abort, is_execution_space / is_memory_space
#ifndef KB200_COMPAT_HPP
#define KB200_COMPAT_HPP

#include "View.hpp"
#include "Atomic.hpp"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <typeinfo>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>

namespace kb200 {

namespace Impl {
template <class Space>
inline void* raw_allocate(size_t bytes) {
  if (bytes == 0) return nullptr;
  void* p = nullptr;
  if constexpr (std::is_same<Space, HostSpace>::value) {
    p = std::malloc(bytes);
    if (!p) throw RawMemoryAllocationFailure("kb200::kokkos_malloc<HostSpace>: out of memory");
  } else if constexpr (std::is_same<Space, B200HostPinnedSpace>::value) {
    throw_on_error(b200_malloc_host_pinned(bytes, &p));
  } else {
    throw_on_error(b200_malloc(B200().impl_instance(), bytes, &p));
  }
  return p;
}
template <class Space>
inline void raw_deallocate(void* p) {
  if (!p) return;
  if constexpr (std::is_same<Space, HostSpace>::value) std::free(p);
  else if constexpr (std::is_same<Space, B200HostPinnedSpace>::value) throw_on_error(b200_free_host_pinned(p));
  else throw_on_error(b200_free(B200().impl_instance(), p));
}
}  // namespace Impl

template <class Space = B200Space>
inline void* kokkos_malloc(const std::string& /*label*/, size_t bytes) { return Impl::raw_allocate<typename Space::memory_space>(bytes); }
template <class Space = B200Space>
inline void* kokkos_malloc(size_t bytes) { return Impl::raw_allocate<typename Space::memory_space>(bytes); }
template <class Space = B200Space>
inline void kokkos_free(void* p) { Impl::raw_deallocate<typename Space::memory_space>(p); }

KB200_INLINE_FUNCTION void abort(const char* msg) {
#ifdef __CUDA_ARCH__
  // device-side assertion: the message reaches the host's stderr and the context fails with cudaErrorAssert
  // (core/src/Cuda/Kokkos_Abort.hpp does the same)
  __assert_fail(msg, "kb200::abort", 0, "");
#else
  std::fprintf(stderr, "kb200::abort: %s\n", msg);
  std::abort();
#endif
}

// Host-side tag type: enough for create_mirror_view_and_copy(DefaultHostExecutionSpace(), view) and memory_space queries.
// It is NOT an execution space that can run patterns (this package has no host backend by design).
struct DefaultHostExecutionSpace {
  using execution_space = DefaultHostExecutionSpace;
  using memory_space = HostSpace;
  using device_type = Device<DefaultHostExecutionSpace, HostSpace>;
  static const char* name() { return "HostTag"; }
  void fence(const std::string& = "") const {}
};

// device-wide memory fence (core/src/Kokkos_Atomics_Desul_Wrapper.hpp: Kokkos::memory_fence)
KB200_FORCEINLINE_FUNCTION void memory_fence() {
#ifdef __CUDA_ARCH__
  __threadfence();
#else
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
#endif
}

namespace numbers {  // core/src/Kokkos_MathematicalConstants.hpp:30-75 (the C++20 <numbers> set; floating-point types only)
#define KB200_MATH_CONSTANT(NAME, VALUE)                                                                               \
  template <class T> inline constexpr auto NAME##_v = std::enable_if_t<std::is_floating_point<T>::value, T>(VALUE##L); \
  inline constexpr double NAME = NAME##_v<double>;
KB200_MATH_CONSTANT(e, 2.718281828459045235360287471352662498)
KB200_MATH_CONSTANT(log2e, 1.442695040888963407359924681001892137)
KB200_MATH_CONSTANT(log10e, 0.434294481903251827651128918916605082)
KB200_MATH_CONSTANT(pi, 3.141592653589793238462643383279502884)
KB200_MATH_CONSTANT(inv_pi, 0.318309886183790671537767526745028724)
KB200_MATH_CONSTANT(inv_sqrtpi, 0.564189583547756286948079451560772586)
KB200_MATH_CONSTANT(ln2, 0.693147180559945309417232121458176568)
KB200_MATH_CONSTANT(ln10, 2.302585092994045684017991454684364208)
KB200_MATH_CONSTANT(sqrt2, 1.414213562373095048801688724209698079)
KB200_MATH_CONSTANT(sqrt3, 1.732050807568877293527446341505872367)
KB200_MATH_CONSTANT(inv_sqrt3, 0.577350269189625764509148780501957456)
KB200_MATH_CONSTANT(egamma, 0.577215664901532860606512090082402431)
KB200_MATH_CONSTANT(phi, 1.618033988749894848204586834365638118)
#undef KB200_MATH_CONSTANT
}  // namespace numbers

// Kokkos::printf (core/src/Kokkos_Printf.hpp): callable from host and device code
template <class... Args>
KB200_FORCEINLINE_FUNCTION void printf(const char* fmt, Args... args) {
  if constexpr (sizeof...(Args) == 0) ::printf("%s", fmt);
  else ::printf(fmt, args...);
}

struct InvalidType {};  // core/src/Kokkos_Core_fwd.hpp:56: "no argument" marker in variadic test helpers

// Kokkos::pair (core/src/Kokkos_Pair.hpp): std::pair usable in device code
template <class T1, class T2>
struct pair {
  using first_type = T1;
  using second_type = T2;
  T1 first;
  T2 second;
  constexpr pair() = default;
  KB200_FORCEINLINE_FUNCTION constexpr pair(const T1& f, const T2& s) : first(f), second(s) {}
  template <class U, class V>
  KB200_FORCEINLINE_FUNCTION constexpr pair(const pair<U, V>& p) : first(p.first), second(p.second) {}
  template <class U, class V>
  pair(const std::pair<U, V>& p) : first(p.first), second(p.second) {}
  std::pair<T1, T2> to_std_pair() const { return std::make_pair(first, second); }
};
template <class T1, class T2>
KB200_FORCEINLINE_FUNCTION constexpr bool operator==(const pair<T1, T2>& a, const pair<T1, T2>& b) { return a.first == b.first && a.second == b.second; }
template <class T1, class T2>
KB200_FORCEINLINE_FUNCTION constexpr bool operator!=(const pair<T1, T2>& a, const pair<T1, T2>& b) { return !(a == b); }
template <class T1, class T2>
KB200_FORCEINLINE_FUNCTION constexpr bool operator<(const pair<T1, T2>& a, const pair<T1, T2>& b) {
  return a.first < b.first || (!(b.first < a.first) && a.second < b.second);
}
template <class T1, class T2>
KB200_FORCEINLINE_FUNCTION constexpr pair<T1, T2> make_pair(T1 a, T2 b) { return pair<T1, T2>(a, b); }

// Kokkos math functions used in functors (core/src/Kokkos_MathematicalFunctions.hpp, Kokkos_MinMax.hpp).  Only declared in
// Kokkos-namespace mode: a translation unit that says `using namespace kb200;` before CUDA's math headers would otherwise see
// two candidates for every unqualified sqrt()/fabs() call inside those headers.
#ifdef KB200_AS_KOKKOS
// min / max / minmax / clamp with the std:: tie rules (core/src/Kokkos_MinMax.hpp:30-200): min and max return the FIRST of
// equivalent arguments, minmax returns {leftmost smallest, rightmost largest}; each with an optional comparator and over a braced list
template <class T> KB200_FORCEINLINE_FUNCTION constexpr const T& min(const T& a, const T& b) { return b < a ? b : a; }
template <class T> KB200_FORCEINLINE_FUNCTION constexpr const T& max(const T& a, const T& b) { return a < b ? b : a; }
template <class T, class C> KB200_FORCEINLINE_FUNCTION constexpr const T& min(const T& a, const T& b, C comp) { return comp(b, a) ? b : a; }
template <class T, class C> KB200_FORCEINLINE_FUNCTION constexpr const T& max(const T& a, const T& b, C comp) { return comp(a, b) ? b : a; }
template <class T, class C>
KB200_FORCEINLINE_FUNCTION constexpr T min(std::initializer_list<T> l, C comp) {
  auto it = l.begin();
  auto best = it;
  for (++it; it != l.end(); ++it)
    if (comp(*it, *best)) best = it;
  return *best;
}
template <class T, class C>
KB200_FORCEINLINE_FUNCTION constexpr T max(std::initializer_list<T> l, C comp) {
  auto it = l.begin();
  auto best = it;
  for (++it; it != l.end(); ++it)
    if (comp(*best, *it)) best = it;
  return *best;
}
namespace Impl { struct less_than { template <class A, class B> KB200_FORCEINLINE_FUNCTION constexpr bool operator()(const A& a, const B& b) const { return a < b; } }; }
template <class T> KB200_FORCEINLINE_FUNCTION constexpr T min(std::initializer_list<T> l) { return kb200::min(l, Impl::less_than{}); }
template <class T> KB200_FORCEINLINE_FUNCTION constexpr T max(std::initializer_list<T> l) { return kb200::max(l, Impl::less_than{}); }
template <class T, class C>
KB200_FORCEINLINE_FUNCTION constexpr pair<const T&, const T&> minmax(const T& a, const T& b, C comp) {
  using R = pair<const T&, const T&>;
  return comp(b, a) ? R{b, a} : R{a, b};
}
template <class T> KB200_FORCEINLINE_FUNCTION constexpr pair<const T&, const T&> minmax(const T& a, const T& b) { return kb200::minmax(a, b, Impl::less_than{}); }
template <class T, class C>
KB200_FORCEINLINE_FUNCTION constexpr pair<T, T> minmax(std::initializer_list<T> l, C comp) {
  auto it = l.begin();
  auto lo = it, hi = it;
  for (++it; it != l.end(); ++it) {
    if (comp(*it, *lo)) lo = it;
    else if (!comp(*it, *hi)) hi = it;
  }
  return pair<T, T>{*lo, *hi};
}
template <class T> KB200_FORCEINLINE_FUNCTION constexpr pair<T, T> minmax(std::initializer_list<T> l) { return kb200::minmax(l, Impl::less_than{}); }
template <class T> KB200_FORCEINLINE_FUNCTION constexpr const T& clamp(const T& v, const T& lo, const T& hi) { return v < lo ? lo : (hi < v ? hi : v); }
template <class T, class C> KB200_FORCEINLINE_FUNCTION constexpr const T& clamp(const T& v, const T& lo, const T& hi, C comp) { return comp(v, lo) ? lo : (comp(hi, v) ? hi : v); }
#define KB200_MATH_1(NAME) \
  KB200_FORCEINLINE_FUNCTION double NAME(double x) { return ::NAME(x); } \
  KB200_FORCEINLINE_FUNCTION float NAME(float x) { return ::NAME##f(x); }
KB200_MATH_1(sqrt) KB200_MATH_1(fabs) KB200_MATH_1(exp) KB200_MATH_1(log) KB200_MATH_1(sin) KB200_MATH_1(cos) KB200_MATH_1(floor) KB200_MATH_1(ceil)
#undef KB200_MATH_1
KB200_FORCEINLINE_FUNCTION double fmin(double a, double b) { return ::fmin(a, b); }
KB200_FORCEINLINE_FUNCTION double fmax(double a, double b) { return ::fmax(a, b); }
KB200_FORCEINLINE_FUNCTION float fmin(float a, float b) { return ::fminf(a, b); }
KB200_FORCEINLINE_FUNCTION float fmax(float a, float b) { return ::fmaxf(a, b); }
KB200_FORCEINLINE_FUNCTION double pow(double a, double b) { return ::pow(a, b); }
template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
KB200_FORCEINLINE_FUNCTION T abs(T x) { return x < T(0) ? T(-x) : x; }
#endif  // KB200_AS_KOKKOS

namespace Experimental {
using half_t = __half;          // core/src/Kokkos_Half_FloatingPointWrapper.hpp (sm_100 has native fp16 / bf16)
using bhalf_t = __nv_bfloat16;
// Kokkos::Experimental::require(policy, WorkItemProperty): launch hints; accepted and ignored on this backend
struct WorkItemProperty {  // (a class with constant members, as in the reference, so that `using Experimental::WorkItemProperty;` works)
  struct HintLightWeight_t {};
  struct HintHeavyWeight_t {};
  struct None_t {};
  static constexpr HintLightWeight_t HintLightWeight{};
  static constexpr HintHeavyWeight_t HintHeavyWeight{};
  static constexpr None_t None{};
};
template <class Policy, class Property>
inline Policy require(const Policy& p, Property) { return p; }
}  // namespace Experimental
// reduction identities of the 16-bit floating types, from their bit patterns (core/src/Kokkos_Half_NumericTraits.hpp:40-230)
template <>
struct reduction_identity<__half> {
  KB200_FORCEINLINE_FUNCTION static __half bits(unsigned short b) { __half_raw r; r.x = b; return __half(r); }
  KB200_FORCEINLINE_FUNCTION static __half sum() { return bits(0x0000); }
  KB200_FORCEINLINE_FUNCTION static __half prod() { return bits(0x3C00); }
  KB200_FORCEINLINE_FUNCTION static __half max() { return bits(0xFBFF); }  // -65504
  KB200_FORCEINLINE_FUNCTION static __half min() { return bits(0x7BFF); }  // +65504
};
template <>
struct reduction_identity<__nv_bfloat16> {
  KB200_FORCEINLINE_FUNCTION static __nv_bfloat16 bits(unsigned short b) { __nv_bfloat16_raw r; r.x = b; return __nv_bfloat16(r); }
  KB200_FORCEINLINE_FUNCTION static __nv_bfloat16 sum() { return bits(0x0000); }
  KB200_FORCEINLINE_FUNCTION static __nv_bfloat16 prod() { return bits(0x3F80); }
  KB200_FORCEINLINE_FUNCTION static __nv_bfloat16 max() { return bits(0xFF7F); }
  KB200_FORCEINLINE_FUNCTION static __nv_bfloat16 min() { return bits(0x7F7F); }
};

namespace Impl {
template <class T>
struct TypeInfo { static std::string name() { return typeid(T).name(); } };  // core/src/impl/Kokkos_TypeInfo.hpp
}  // namespace Impl

class Timer {  // core/src/Kokkos_Timer.hpp
  std::chrono::steady_clock::time_point m_t0 = std::chrono::steady_clock::now();
 public:
  void reset() { m_t0 = std::chrono::steady_clock::now(); }
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - m_t0).count(); }
};

template <class T> struct is_execution_space : std::is_same<std::remove_cv_t<T>, B200> {};
template <class T> inline constexpr bool is_execution_space_v = is_execution_space<T>::value;
template <class T> struct is_memory_space
    : std::integral_constant<bool, std::is_same<T, HostSpace>::value || std::is_same<T, B200Space>::value || std::is_same<T, B200HostPinnedSpace>::value> {};
template <class T> inline constexpr bool is_memory_space_v = is_memory_space<T>::value;

}  // namespace kb200

// The two wrap-around counters of desul that Kokkos does not re-export (tpls/desul/include/desul/atomics/Fetch_Op_Generic.hpp;
// used directly by core/unit_test/TestAtomicOperations.hpp:336-380): old value returned, new value = (old >= wrap) ? 0 : old + 1
// resp. (old == 0 || old > wrap) ? wrap : old - 1.  Memory order / scope tags are accepted (relaxed, device scope is what the
// Kokkos wrappers always request: core/src/Kokkos_Atomics_Desul_Wrapper.hpp:72-148).
#ifdef KB200_AS_KOKKOS  // (with the real desul in the translation unit -- the Kokkos::B200 adapter -- its own definitions are used)
namespace desul {
struct MemoryOrderRelaxed {};
struct MemoryOrderSeqCst {};
struct MemoryOrderAcqRel {};
struct MemoryScopeDevice {};
struct MemoryScopeCore {};
struct MemoryScopeSystem {};
template <class T, class Order, class Scope>
KB200_FORCEINLINE_FUNCTION T atomic_fetch_inc_mod(T* p, T wrap, Order, Scope) {
  T old = kb200::atomic_load(p);
  while (true) {
    const T nw = old >= wrap ? T(0) : (T)(old + T(1));
    const T seen = kb200::atomic_compare_exchange(p, old, nw);
    if (seen == old) return old;
    old = seen;
  }
}
template <class T, class Order, class Scope>
KB200_FORCEINLINE_FUNCTION T atomic_fetch_dec_mod(T* p, T wrap, Order, Scope) {
  T old = kb200::atomic_load(p);
  while (true) {
    const T nw = (old == T(0) || old > wrap) ? wrap : (T)(old - T(1));
    const T seen = kb200::atomic_compare_exchange(p, old, nw);
    if (seen == old) return old;
    old = seen;
  }
}
}  // namespace desul

// warning-control macros the reference's sources use around tests (core/src/Kokkos_Macros.hpp)
#define KOKKOS_IMPL_DISABLE_UNREACHABLE_WARNINGS_PUSH()
#define KOKKOS_IMPL_DISABLE_UNREACHABLE_WARNINGS_POP()
#define KOKKOS_IMPL_DISABLE_DEPRECATED_WARNINGS_PUSH()
#define KOKKOS_IMPL_DISABLE_DEPRECATED_WARNINGS_POP()
#endif  // KB200_AS_KOKKOS
#endif
