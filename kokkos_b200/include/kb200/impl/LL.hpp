// kb200/impl/LL.hpp -- "LL" (low-latency) synchronisation words for values exchanged between CTAs of one GPU or between
// GPUs over NVLink: every 8-byte word carries 32 bits of payload and a 32-bit tag, so a reader needs no ordering fence
// (a word is valid iff its tag matches) and nothing is ever cleared (tags only grow).  An 8-byte value = two words.
// Same idea as NCCL's LL protocol; the reference has no counterpart (single-GPU kernels only).
#ifndef KB200_IMPL_LL_HPP
#define KB200_IMPL_LL_HPP

#include "../Macros.hpp"
#include <cstring>

namespace kb200 {
namespace Impl {

template <class T>
KB200_DEVICE_FUNCTION unsigned long long to_bits(T v) {
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
KB200_DEVICE_FUNCTION T from_bits(unsigned long long b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}

namespace ll {
template <class T>
KB200_DEVICE_FUNCTION void pack(T v, unsigned tag, unsigned long long& w0, unsigned long long& w1) {
  const unsigned long long b = to_bits(v);
  w0 = ((unsigned long long)tag << 32) | (b & 0xffffffffull);
  w1 = ((unsigned long long)tag << 32) | (b >> 32);
}
KB200_DEVICE_FUNCTION bool ok(unsigned long long w0, unsigned long long w1, unsigned tag) {
  return (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag;
}
template <class T>
KB200_DEVICE_FUNCTION T unpack(unsigned long long w0, unsigned long long w1) {
  return from_bits<T>((w0 & 0xffffffffull) | (w1 << 32));
}
KB200_DEVICE_FUNCTION void st_gpu(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
KB200_DEVICE_FUNCTION void ld_gpu(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
KB200_DEVICE_FUNCTION void st_sys(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
KB200_DEVICE_FUNCTION void ld_sys(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
KB200_DEVICE_FUNCTION unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// a peer (or a CTA of this GPU) never showed up: record why and stop the kernel instead of hanging the GPU
KB200_DEVICE_FUNCTION void give_up(unsigned* err, unsigned code) {
  if (err) {
    *reinterpret_cast<volatile unsigned*>(err) = code;
    __threadfence_system();
  }
  __trap();
}
// poll an LL pair until it carries `tag`
template <class T, bool SYS>
KB200_DEVICE_FUNCTION T wait_value(const unsigned long long* p, unsigned tag, unsigned long long timeout_ns, unsigned* err, unsigned code) {
  unsigned long long w0, w1;
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    if (SYS) ld_sys(p, w0, w1); else ld_gpu(p, w0, w1);
    if (ok(w0, w1, tag)) return unpack<T>(w0, w1);
    if ((spin & 1023u) == 1023u) {
      const unsigned long long t = now_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > timeout_ns) give_up(err, code);
    }
  }
}
}  // namespace ll

}  // namespace Impl
}  // namespace kb200
#endif
