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

// One-time per-kernel set-up (cudaFuncSetAttribute opt-ins, occupancy queries) is per DEVICE: a process may drive several
// GPUs through B200::on_device / partition_space (core/unit_test/TestMultiGPU.hpp).  `here()` is the slot of the current device.
struct PerDeviceInt {
  int v[64] = {};
  int& here() {
    int d = 0;
    cudaGetDevice(&d);
    return v[d & 63];
  }
};

// Every launcher builds one of these first: it makes the instance's device current for the duration of the call (launches,
// scratch growth and occupancy queries all act on the current device) and restores the caller's device afterwards.
struct HostRuntime {
  b200_instance* inst;
  int prev_device = -1;
  explicit HostRuntime(b200_instance* i) : inst(i) {
    const int want = b200_instance_device(i);
    int cur = -1;
    if (want >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != want) {
      cudaSetDevice(want);
      prev_device = cur;
    }
  }
  ~HostRuntime() {
    if (prev_device >= 0) cudaSetDevice(prev_device);
  }
  HostRuntime(const HostRuntime&) = delete;
  HostRuntime& operator=(const HostRuntime&) = delete;
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
