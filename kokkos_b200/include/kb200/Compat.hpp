// kb200/Compat.hpp -- small pieces of the Kokkos Core surface that user code and the reference's own unit tests
// (core/unit_test/incremental/*.hpp, built by tests/ref_unit) touch next to the hot path.
//   kokkos_malloc / kokkos_free / kokkos_realloc   core/src/Kokkos_Core.hpp:155-200 (raw allocations in a memory space)
//   ExecutionSpace::memory_space / device_type helpers, Kokkos::Timer, Kokkos::abort, is_execution_space / is_memory_space
#ifndef KB200_COMPAT_HPP
#define KB200_COMPAT_HPP

#include "View.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>

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
  printf("kb200::abort: %s\n", msg);
  __trap();
#else
  std::fprintf(stderr, "kb200::abort: %s\n", msg);
  std::abort();
#endif
}

struct ParallelForTag {};     // core/src/Kokkos_Core_fwd.hpp: pattern tags for team_size_max / team_size_recommended
struct ParallelReduceTag {};
struct ParallelScanTag {};

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

// warning-control macros the reference's sources use around tests (core/src/Kokkos_Macros.hpp)
#define KOKKOS_IMPL_DISABLE_UNREACHABLE_WARNINGS_PUSH()
#define KOKKOS_IMPL_DISABLE_UNREACHABLE_WARNINGS_POP()
#define KOKKOS_IMPL_DISABLE_DEPRECATED_WARNINGS_PUSH()
#define KOKKOS_IMPL_DISABLE_DEPRECATED_WARNINGS_POP()
#endif
