// kb200/B200.hpp -- the execution-space class `kb200::B200` and process-level initialize/finalize/fence.
//
// Mirrors the members a Kokkos backend must provide (impl/Kokkos_ExecSpaceManager.hpp:30-108; model:
// core/src/Cuda/Kokkos_Cuda.hpp:95-247): nested type names, impl_initialize/impl_finalize/impl_static_fence,
// fence(label), concurrency(), print_configuration(), name(), impl_instance_id(), equality, and cheap copies
// (sizeof <= 2 pointers: the class is a ref-counted handle on a C-ABI instance).
// Error convention (core/src/Cuda/Kokkos_Cuda_Error.hpp:45-68): context-poisoning CUDA errors abort, every
// other failure throws std::runtime_error; allocation failure throws RawMemoryAllocationFailure.
#ifndef KB200_B200_HPP
#define KB200_B200_HPP

#include "Macros.hpp"
#include <kokkos_b200.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <array>
#include <memory>
#include <utility>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace kb200 {

struct RawMemoryAllocationFailure : std::runtime_error {
  using std::runtime_error::runtime_error;
};

namespace Impl {
inline bool is_sticky(int rc) {
  switch (rc) {
    case cudaErrorIllegalAddress: case cudaErrorLaunchFailure: case cudaErrorHardwareStackError:
    case cudaErrorIllegalInstruction: case cudaErrorMisalignedAddress: case cudaErrorInvalidAddressSpace:
    case cudaErrorInvalidPc: case cudaErrorAssert: case cudaErrorECCUncorrectable: case cudaErrorLaunchTimeout:
      return true;
    default: return false;
  }
}
// rc -> the reference's throw / abort split
inline void throw_on_error(int rc) {
  if (rc == 0) return;
  const std::string msg = b200_last_error_string();
  if (rc > 0 && is_sticky(rc)) {
    std::fprintf(stderr, "kb200: unrecoverable CUDA error: %s\n", msg.c_str());
    std::abort();
  }
  if (rc == B200_ENOMEM) throw RawMemoryAllocationFailure(msg);
  throw std::runtime_error(msg);
}
struct InstanceDeleter {
  void operator()(b200_instance* p) const { if (p) b200_finalize(p); }
};
inline std::shared_ptr<b200_instance>& default_instance() {
  static std::shared_ptr<b200_instance> inst;
  return inst;
}
}  // namespace Impl

struct B200Space;     // device memory (View.hpp)
template <class ExecSpace, class MemSpace>
struct Device {  // core/src/Kokkos_Core_fwd.hpp: execution space + memory space pair
  using execution_space = ExecSpace;
  using memory_space = MemSpace;
  using device_type = Device<ExecSpace, MemSpace>;
};
struct LayoutLeft;
template <class ExecSpace> class ScratchMemorySpace;

struct InitializationSettings {
  int device_id = 0;
  InitializationSettings& set_device_id(int d) { device_id = d; return *this; }
  int get_device_id() const { return device_id; }
};

class B200 {
 public:
  using execution_space = B200;
  using memory_space = B200Space;
  using device_type = Device<B200, B200Space>;
  using array_layout = LayoutLeft;
  using size_type = long long;  // the default index type of policies on this space (64-bit: no 2^31 special case, Policy.hpp)
  using scratch_memory_space = ScratchMemorySpace<B200>;

  // default instance (Kokkos::Cuda()): requires kb200::initialize()
  B200() : m_inst(Impl::default_instance()) {
    if (!m_inst) throw std::runtime_error("kb200::B200(): kb200::initialize() has not been called");
  }
  // instance on a caller-provided stream (Kokkos::Cuda(cudaStream_t)); the stream is not owned
  explicit B200(cudaStream_t stream, int device = -1) {
    if (device < 0) cudaGetDevice(&device);
    b200_instance* p = nullptr;
    Impl::throw_on_error(b200_instance_create(device, (void*)stream, &p));
    m_inst.reset(p, Impl::InstanceDeleter());
  }
  // fresh instance with its own stream on `device` (one per GPU: TestMultiGPU.hpp:21-96)
  static B200 on_device(int device) {
    b200_instance* p = nullptr;
    Impl::throw_on_error(b200_init(device, &p));
    B200 s{std::shared_ptr<b200_instance>(p, Impl::InstanceDeleter())};
    return s;
  }

  static void impl_initialize(const InitializationSettings& s) {
    if (Impl::default_instance()) return;
    b200_instance* p = nullptr;
    Impl::throw_on_error(b200_init(s.device_id, &p));
    Impl::default_instance().reset(p, Impl::InstanceDeleter());
  }
  static void impl_finalize() { Impl::default_instance().reset(); }
  static bool impl_is_initialized() { return (bool)Impl::default_instance(); }
  static void impl_static_fence(const std::string& label) {
    if (Impl::default_instance()) Impl::throw_on_error(b200_fence(Impl::default_instance().get(), label.c_str()));
    cudaError_t e = cudaDeviceSynchronize();  // all instances on the device (Cuda_Instance.cpp:139-157)
    if (e != cudaSuccess) Impl::throw_on_error(b200_report_error((int)e, label.c_str()));
  }

  void fence(const std::string& label = "kb200::B200::fence(): unnamed instance fence") const {
    Impl::throw_on_error(b200_fence(m_inst.get(), label.c_str()));
  }
  int concurrency() const { return props().concurrency; }
  static const char* name() { return "B200"; }
  uint32_t impl_instance_id() const noexcept { return b200_instance_id(m_inst.get()); }
  cudaStream_t cuda_stream() const { return (cudaStream_t)b200_instance_stream(m_inst.get()); }
  int cuda_device() const { return props().device; }
  b200_instance* impl_instance() const { return m_inst.get(); }
  b200_props props() const {
    b200_props p;
    Impl::throw_on_error(b200_device_props(m_inst.get(), &p));
    return p;
  }
  void print_configuration(std::ostream& os, bool /*verbose*/ = false) const {
    const b200_props p = props();
    os << "Device Execution Space:\n  KB200_ENABLE_B200: yes (sm_100a only, no host fallback)\n"
       << "B200 device " << p.device << ": " << p.name << ", cc " << p.cc_major << "." << p.cc_minor << ", " << p.sm_count
       << " SMs, " << (p.total_mem >> 20) << " MiB, L2 " << (p.l2_bytes >> 20) << " MiB, concurrency " << p.concurrency << "\n";
  }
  friend bool operator==(const B200& a, const B200& b) { return a.m_inst == b.m_inst; }
  friend bool operator!=(const B200& a, const B200& b) { return !(a == b); }

 private:
  explicit B200(std::shared_ptr<b200_instance> p) : m_inst(std::move(p)) {}
  std::shared_ptr<b200_instance> m_inst;
};
static_assert(sizeof(B200) <= 2 * sizeof(void*), "execution space handles must stay cheap to copy");

using DefaultExecutionSpace = B200;

inline void initialize(const InitializationSettings& s = InitializationSettings()) { B200::impl_initialize(s); }
inline void initialize(int& /*argc*/, char** /*argv*/) { B200::impl_initialize(InitializationSettings()); }
inline void finalize() { B200::impl_finalize(); }
inline bool is_initialized() { return B200::impl_is_initialized(); }
inline void fence(const std::string& label = "kb200::fence") { B200::impl_static_fence(label); }

struct ScopeGuard {
  explicit ScopeGuard(const InitializationSettings& s = InitializationSettings()) { initialize(s); }
  ~ScopeGuard() { finalize(); }
  ScopeGuard(const ScopeGuard&) = delete;
  ScopeGuard& operator=(const ScopeGuard&) = delete;
};

namespace Experimental {
// Kokkos::Experimental::partition_space: n independent instances (own streams) on the same device
// (core/src/Cuda/Kokkos_Cuda_Instance.hpp:368-385)
namespace detail {
template <size_t... Is>
std::array<B200, sizeof...(Is)> make_instances(int device, std::index_sequence<Is...>) {
  return std::array<B200, sizeof...(Is)>{((void)Is, B200::on_device(device))...};
}
}  // namespace detail
// weights given as arguments: a std::array (usable with structured bindings); as a vector: a std::vector
template <class... W, class = std::enable_if_t<(std::is_arithmetic<W>::value && ...)>>
std::array<B200, sizeof...(W)> partition_space(const B200& base, W... /*weights*/) {
  return detail::make_instances(base.cuda_device(), std::make_index_sequence<sizeof...(W)>{});
}
inline std::vector<B200> partition_space(const B200& base, const std::vector<int>& weights) {
  std::vector<B200> out;
  for (size_t k = 0; k < weights.size(); ++k) out.push_back(B200::on_device(base.cuda_device()));
  return out;
}
}  // namespace Experimental

}  // namespace kb200
#endif
