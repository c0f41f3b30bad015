// kb200/impl/ContigBody.hpp -- reduction "bodies" over contiguous typed arrays with wide streaming loads.
//
// A generic Kokkos functor `f(i, update)` hides its loads from the backend; for the recognised
// contiguous-View cases the execution space issues them itself: 8/16/32-byte
// ld.global.nc.L1::no_allocate (SASS LDG.E.{64,128,ENL2.256}.CONSTANT), sm_100's widest access.
// The reference's ParallelReduce loop issues one scalar LDG.E.64 per trip
// (core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:210-215).
#ifndef KB200_IMPL_CONTIGBODY_HPP
#define KB200_IMPL_CONTIGBODY_HPP

#include "../Macros.hpp"
#include <cstring>

namespace kb200 {
namespace Impl {

template <int BYTES>
struct RawVec;
template <>
struct RawVec<4> { unsigned w[1]; };
template <>
struct RawVec<8> { unsigned long long w[1]; };
template <>
struct RawVec<16> { unsigned long long w[2]; };
template <>
struct RawVec<32> { unsigned long long w[4]; };

// read-only streaming load: non-coherent path, no L1 allocation (each byte is used once)
template <int BYTES>
KB200_DEVICE_FUNCTION RawVec<BYTES> ld_stream(const void* p) {
  RawVec<BYTES> r;
  if constexpr (BYTES == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r.w[0]) : "l"(p));
  } else if constexpr (BYTES == 8) {
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(r.w[0]) : "l"(p));
  } else if constexpr (BYTES == 16) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(r.w[0]), "=l"(r.w[1]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3])
                 : "l"(p));
  }
  return r;
}
// coherent variant for data the same kernel may also write (in-place operations)
template <int BYTES>
KB200_DEVICE_FUNCTION RawVec<BYTES> ld_plain(const void* p) {
  RawVec<BYTES> r;
  if constexpr (BYTES == 4) {
    asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(r.w[0]) : "l"(p) : "memory");
  } else if constexpr (BYTES == 8) {
    asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(r.w[0]) : "l"(p) : "memory");
  } else if constexpr (BYTES == 16) {
    asm volatile("ld.global.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(r.w[0]), "=l"(r.w[1]) : "l"(p) : "memory");
  } else {
    asm volatile("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3])
                 : "l"(p)
                 : "memory");
  }
  return r;
}
template <int BYTES>
KB200_DEVICE_FUNCTION void st_stream(void* p, const RawVec<BYTES>& r) {
  if constexpr (BYTES == 4) {
    asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(p), "r"(r.w[0]) : "memory");
  } else if constexpr (BYTES == 8) {
    asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" ::"l"(p), "l"(r.w[0]) : "memory");
  } else if constexpr (BYTES == 16) {
    asm volatile("st.global.L1::no_allocate.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(r.w[0]), "l"(r.w[1]) : "memory");
  } else {
    asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(r.w[0]), "l"(r.w[1]), "l"(r.w[2]),
                 "l"(r.w[3])
                 : "memory");
  }
}

// split [0,n) of T at address x into  head | nvec vectors of VBYTES | tail
template <class T, int VBYTES>
struct VecSplit {
  static constexpr int EPV = VBYTES / (int)sizeof(T);
  int64 head, nvec, tail;
  __host__ VecSplit(const T* x, int64 n) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(x);
    int64 h = 0;
    if (a % VBYTES) h = (int64)((VBYTES - a % VBYTES) / sizeof(T));
    if (a % sizeof(T)) { h = n; }  // not even element aligned: everything goes through the scalar edge path
    if (h > n) h = n;
    head = h;
    nvec = (n - h) / EPV;
    tail = n - h - nvec * EPV;
  }
};

// ElemOp: static void apply(V& acc, T x, int64 i)
template <class T, int VBYTES, class ElemOp, class V>
struct ContigBody {
  static constexpr int EPV = VBYTES / (int)sizeof(T);
  using packet = RawVec<VBYTES>;
  const T* x;
  int64 head, nvec, tail;
  int64 index_base;

  __host__ ContigBody(const T* x_, int64 n, int64 base) : x(x_), index_base(base) {
    VecSplit<T, VBYTES> s(x_, n);
    head = s.head; nvec = s.nvec; tail = s.tail;
  }
  KB200_DEVICE_FUNCTION packet load(int64 u) const { return ld_stream<VBYTES>(x + head + u * EPV); }
  KB200_DEVICE_FUNCTION void consume(const packet& p, int64 u, V& acc) const {
    T e[EPV];
    memcpy(e, p.w, VBYTES);
    const int64 i0 = index_base + head + u * EPV;
#pragma unroll
    for (int k = 0; k < EPV; ++k) ElemOp::apply(acc, e[k], i0 + k);
  }
  KB200_FUNCTION int64 edge_count() const { return head + tail; }
  KB200_DEVICE_FUNCTION void edge(int64 k, V& acc) const {
    const int64 i = k < head ? k : head + nvec * EPV + (k - head);
    ElemOp::apply(acc, x[i], index_base + i);
  }
};

}  // namespace Impl
}  // namespace kb200
#endif
