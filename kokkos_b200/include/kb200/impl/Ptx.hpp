// kb200/impl/Ptx.hpp -- sm_100a PTX wrappers: mbarrier, 1-D bulk async copies (TMA engine,
// SASS UBLKCP), proxy fences, relaxed 128-bit descriptor accesses.
#ifndef KB200_IMPL_PTX_HPP
#define KB200_IMPL_PTX_HPP

#include "../Macros.hpp"

namespace kb200 {
namespace Impl {
namespace ptx {

KB200_DEVICE_FUNCTION unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

KB200_DEVICE_FUNCTION void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
KB200_DEVICE_FUNCTION void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
KB200_DEVICE_FUNCTION void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
KB200_DEVICE_FUNCTION void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
KB200_DEVICE_FUNCTION bool mbar_try_wait(void* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
KB200_DEVICE_FUNCTION void mbar_wait(void* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both 16-byte aligned)
KB200_DEVICE_FUNCTION void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// global -> L2 only (no destination in shared memory, nothing to wait for): SASS UBLKPF
KB200_DEVICE_FUNCTION void bulk_prefetch_l2(const void* src_gmem, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// shared -> global, tracked by bulk async-groups
KB200_DEVICE_FUNCTION void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
KB200_DEVICE_FUNCTION void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
KB200_DEVICE_FUNCTION void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
KB200_DEVICE_FUNCTION void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (before a bulk store)
KB200_DEVICE_FUNCTION void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16-byte tile descriptors: one relaxed, device-scope 128-bit access each way
KB200_DEVICE_FUNCTION void st_relaxed_v2(void* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
KB200_DEVICE_FUNCTION void ld_relaxed_v2(const void* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
KB200_DEVICE_FUNCTION void st_release_u64(void* p, unsigned long long a) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}
KB200_DEVICE_FUNCTION unsigned long long ld_acquire_u64(const void* p) {
  unsigned long long a;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
  return a;
}

}  // namespace ptx
}  // namespace Impl
}  // namespace kb200
#endif
