// kb200/impl/HostRuntime.hpp -- C++ convenience over the C ABI (include/kokkos_b200.h) used by the
// header-only launchers.  No state of its own: the instance owns stream and scratch
// (the role CudaInternal plays in core/src/Cuda/Kokkos_Cuda_Instance.hpp:80-365).
#ifndef KB200_IMPL_HOSTRUNTIME_HPP
#define KB200_IMPL_HOSTRUNTIME_HPP

#include "../Macros.hpp"
#include <kokkos_b200.h>
#include <cuda_runtime.h>

namespace kb200 {
namespace Impl {

struct HostRuntime {
  b200_instance* inst;
  explicit HostRuntime(b200_instance* i) : inst(i) {}
  cudaStream_t stream() const { return static_cast<cudaStream_t>(b200_instance_stream(inst)); }
  int sm_count() const { return b200_instance_sm_count(inst); }
  int reduce_scratch(size_t partial_bytes, size_t value_bytes, bool want_slot, void** partials, unsigned** ticket,
                     void** slot_dev, void** slot_host) const {
    return b200_reduce_scratch(inst, partial_bytes, value_bytes, want_slot ? 1 : 0, partials, ticket, slot_dev, slot_host);
  }
  int result_slot(size_t value_bytes, void** slot_dev, void** slot_host, unsigned long long** seq_dev, unsigned long long* seq_value) const {
    return b200_result_slot(inst, value_bytes, slot_dev, slot_host, seq_dev, seq_value);
  }
  int result_wait(const void* slot_host, unsigned long long seq_value, const char* label) const {
    return b200_result_wait(inst, slot_host, seq_value, label);
  }
  int check_launch(const char* where) const { return b200_report_error((int)cudaGetLastError(), where); }
  int fence(const char* label) const { return b200_fence(inst, label); }
};

}  // namespace Impl
}  // namespace kb200
#endif
