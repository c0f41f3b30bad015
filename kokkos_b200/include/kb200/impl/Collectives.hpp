// kb200/impl/Collectives.hpp -- intra-warp / intra-block / inter-block combine machinery.
//
// Replaces core/src/Cuda/Kokkos_Cuda_ReduceScan.hpp (cuda_intra_warp_reduction :43-100,
// cuda_inter_block_reduction :102-186, cuda_intra_block_reduce_scan :409-554,
// cuda_single_inter_block_reduce_scan :565-697) and the word-wise shuffles of
// core/src/Cuda/Kokkos_Cuda_Vectorization.hpp:107-123.
//
// Design (B200-first, nothing taken from the reference's smem trees):
//   * partials live in REGISTERS; a warp combines them with 5 shuffle steps (redux.sync for
//     32-bit integer add/min/max/and/or), one shared-memory hop joins the <=32 warps;
//   * every combine is ORDER PRESERVING: join(dest = lower ranks, src = higher ranks), so a
//     non-commutative (but associative) join sees operands in thread/block order;
//   * blocks publish one partial each; a self-resetting ticket elects the last block, which
//     folds the partials in block order -> bitwise reproducible results for a given grid.
#ifndef KB200_IMPL_COLLECTIVES_HPP
#define KB200_IMPL_COLLECTIVES_HPP

#include "../Macros.hpp"
#include <type_traits>
#include <cstring>

namespace kb200 {
namespace Impl {

constexpr unsigned kFullMask = 0xffffffffu;

// ---- shuffles of arbitrary trivially-copyable values, as N 32-bit words ------------------
template <class T, class ShflOp>
KB200_DEVICE_FUNCTION T shfl_words(const T& v, ShflOp op) {
  // Kokkos moves reduction values between threads bit-wise; user types with hand-written (but member-wise) copy operations
  // are accepted like the reference does (e.g. MyComplex in incremental/Test14_MDRangeReduce.hpp:30-52)
  static_assert(std::is_trivially_destructible<T>::value, "reduction value_type must be bit-wise movable (trivially destructible)");
  constexpr int W = (sizeof(T) + 3) / 4;
  unsigned w[W];
#pragma unroll
  for (int k = 0; k < W; ++k) w[k] = 0;
  memcpy(w, &v, sizeof(T));
#pragma unroll
  for (int k = 0; k < W; ++k) w[k] = op(w[k]);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}
template <class T>
KB200_DEVICE_FUNCTION T shfl_down(const T& v, unsigned delta, unsigned mask = kFullMask) {
  return shfl_words(v, [=](unsigned x) { return __shfl_down_sync(mask, x, delta); });
}
template <class T>
KB200_DEVICE_FUNCTION T shfl_up(const T& v, unsigned delta, unsigned mask = kFullMask) {
  return shfl_words(v, [=](unsigned x) { return __shfl_up_sync(mask, x, delta); });
}
template <class T>
KB200_DEVICE_FUNCTION T shfl_idx(const T& v, int lane, unsigned mask = kFullMask) {
  return shfl_words(v, [=](unsigned x) { return __shfl_sync(mask, x, lane); });
}
template <class T>
KB200_DEVICE_FUNCTION T shfl_xor(const T& v, int m) {
  return shfl_words(v, [=](unsigned x) { return __shfl_xor_sync(kFullMask, x, m); });
}

// ---- warp reduce: lane 0 ends with x0 (+) x1 (+) ... (+) x31, folded in lane order --------
// `Red` provides value_type, join(value_type& dest, const value_type& src).
// A reducer may opt into redux.sync by defining  static constexpr int redux_op  (see Reducers.hpp).
enum ReduxOp { ReduxNone = 0, ReduxAdd, ReduxMin, ReduxMax, ReduxAnd, ReduxOr };

template <class Red, class = void>
struct redux_op_of { static constexpr int value = ReduxNone; };
template <class Red>
struct redux_op_of<Red, std::void_t<decltype(Red::redux_op)>> { static constexpr int value = Red::redux_op; };

// `nl` = number of live lanes of this warp (32 except in the last warp of a block whose size is not a multiple of 32)
template <class Red>
KB200_DEVICE_FUNCTION void warp_reduce(const Red& red, typename Red::value_type& v, int nl = kWarp) {
  using V = typename Red::value_type;
  constexpr int op = redux_op_of<Red>::value;
  const unsigned mask = nl >= kWarp ? kFullMask : ((1u << nl) - 1u);
  if constexpr (op != ReduxNone && std::is_integral<V>::value && sizeof(V) == 4) {
    // redux.sync: one instruction instead of 5 shuffle+op rounds (SASS: REDUX)
    if constexpr (op == ReduxAdd) v = (V)__reduce_add_sync(mask, v);
    if constexpr (op == ReduxMin) v = (V)__reduce_min_sync(mask, v);
    if constexpr (op == ReduxMax) v = (V)__reduce_max_sync(mask, v);
    if constexpr (op == ReduxAnd) v = (V)__reduce_and_sync(mask, (unsigned)v);
    if constexpr (op == ReduxOr) v = (V)__reduce_or_sync(mask, (unsigned)v);
  } else {
    const int lane = (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31;
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
      V hi = ::kb200::Impl::shfl_down(v, d, mask);  // value of lane+d: the HIGHER-ranked operand
      // a partner beyond the live lanes contributes nothing; within a full warp an out-of-range partner
      // returns the lane's own value, whose result lane 0's fold never consumes
      if (nl >= kWarp || lane + d < nl) red.join(v, hi);
    }
  }
}

// ---- block reduce: thread 0 ends with the fold over all threads in thread order ----------
// smem: at least 32 * sizeof(value_type) bytes, 16-byte aligned.  Contains __syncthreads().
template <class Red>
KB200_DEVICE_FUNCTION void block_reduce(const Red& red, typename Red::value_type& v, void* smem) {
  using V = typename Red::value_type;
  V* s = reinterpret_cast<V*>(smem);
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
  const int live = nthreads - (warp << 5);
  warp_reduce(red, v, live < kWarp ? live : kWarp);
  if (nwarps == 1) return;
  if (lane == 0) s[warp] = v;
  __syncthreads();
  if (warp == 0) {
    V w;
    red.init(w);
    if (lane < nwarps) w = s[lane];
    warp_reduce(red, w);
    if (lane == 0) v = w;
  }
}

// __ldcg on arbitrary PODs: fall back to a volatile word copy for sizes the intrinsic lacks
template <class V>
KB200_DEVICE_FUNCTION V load_cg(const V* p) {
  if constexpr (sizeof(V) == 4) {
    unsigned w = __ldcg(reinterpret_cast<const unsigned*>(p));
    V r; memcpy(&r, &w, 4); return r;
  } else if constexpr (sizeof(V) == 8) {
    unsigned long long w = __ldcg(reinterpret_cast<const unsigned long long*>(p));
    V r; memcpy(&r, &w, 8); return r;
  } else if constexpr (sizeof(V) % 16 == 0 && alignof(V) >= 16) {
    V r; uint4* d = reinterpret_cast<uint4*>(&r); const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (unsigned k = 0; k < sizeof(V) / 16; ++k) d[k] = __ldcg(q + k);
    return r;
  } else {
    V r;
    const volatile unsigned char* q = reinterpret_cast<const volatile unsigned char*>(p);
    unsigned char* d = reinterpret_cast<unsigned char*>(&r);
    for (unsigned k = 0; k < sizeof(V); ++k) d[k] = q[k];
    return r;
  }
}

// ---- inter-block combine -----------------------------------------------------------------
struct ReduceScratch {
  void* partials = nullptr;      // >= gridDim.x * sizeof(value_type), device
  unsigned* ticket = nullptr;    // zero before the launch; reset by the last block
  void* result0 = nullptr;       // device-accessible destination (pinned mapped slot or device memory), may be null
  void* result1 = nullptr;       // second destination, may be null
  unsigned long long* seq_ptr = nullptr;  // if set: completion word in mapped host memory, written after the result
  unsigned long long seq_val = 0;         // (system-scope fence in between) so the host can poll instead of synchronising
};
KB200_DEVICE_FUNCTION void reduce_signal_host(const ReduceScratch& s) {
  if (s.seq_ptr) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(s.seq_ptr) = s.seq_val;
  }
}

// An EMPTY iteration range: the result is init() -> final() with no join at all.  A user reducer's join need not be neutral
// on identities (core/unit_test/TestReduceCombinatorical.hpp:27-45 adds 1 per join and expects exactly 0 for N = 0), so the
// kernels return through here instead of combining per-thread identities.  Launched with one block.
template <class Red>
KB200_DEVICE_FUNCTION void reduce_store_identity(const Red& red, const ReduceScratch& s) {
  using V = typename Red::value_type;
  if (threadIdx.x + threadIdx.y + threadIdx.z == 0 && blockIdx.x == 0) {
    V v;
    red.init(v);
    red.final(v);
    if (s.result0) *reinterpret_cast<V*>(s.result0) = v;
    if (s.result1) *reinterpret_cast<V*>(s.result1) = v;
    reduce_signal_host(s);
  }
}

// All threads of every block must call this (it contains barriers).  `v` is meaningful in
// thread 0 (the block's partial).  Red additionally provides init(value_type&) and
// final(value_type&).
template <class Red>
KB200_DEVICE_FUNCTION void grid_reduce_and_store(const Red& red, typename Red::value_type v, const ReduceScratch& s,
                                                 void* smem) {
  using V = typename Red::value_type;
  __shared__ bool is_last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  V* partials = reinterpret_cast<V*>(s.partials);

  if (nblocks == 1) {
    if (tid == 0) {
      red.final(v);
      if (s.result0) *reinterpret_cast<V*>(s.result0) = v;
      if (s.result1) *reinterpret_cast<V*>(s.result1) = v;
      reduce_signal_host(s);
    }
    return;
  }
  if (tid == 0) {
    partials[bid] = v;
    __threadfence();  // partial visible device-wide before the ticket is taken
    const unsigned t = atomicAdd(s.ticket, 1u);
    is_last = (t == nblocks - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();  // acquire side: order the partial reads after the ticket
  // each thread folds a contiguous run of partials (block order), then an ordered block reduce
  const unsigned per = (nblocks + nthreads - 1) / nthreads;
  const unsigned b0 = tid * per;
  V acc;
  red.init(acc);
  for (unsigned b = b0; b < b0 + per && b < nblocks; ++b) {
    V p = load_cg(&partials[b]);  // L2 read: skip any stale L1 line
    red.join(acc, p);
  }
  __syncthreads();  // smem reuse by block_reduce
  block_reduce(red, acc, smem);
  if (tid == 0) {
    red.final(acc);
    if (s.result0) *reinterpret_cast<V*>(s.result0) = acc;
    if (s.result1) *reinterpret_cast<V*>(s.result1) = acc;
    *s.ticket = 0u;  // self-reset: stream order makes this safe for the next launch
    reduce_signal_host(s);
  }
}

}  // namespace Impl
}  // namespace kb200
#endif
