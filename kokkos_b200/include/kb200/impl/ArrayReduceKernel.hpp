// kb200/impl/ArrayReduceKernel.hpp -- parallel_reduce whose value_type is a RUNTIME-LENGTH array: the functor declares
// `using value_type = T[]; unsigned value_count;` and is called as f(i..., T dst[]) with optional init(T[]), join(T[], const T[]),
// final(T[])  (core/src/impl/Kokkos_FunctorAnalysis.hpp:604-613,724-732,865-958; tests: core/unit_test/TestReduce.hpp:124-232,
// TestMDRange.hpp:31-160).  The reference's Cuda backend keeps such accumulators in shared memory for the whole kernel
// (Cuda/Kokkos_Cuda_Parallel_Range.hpp:191-259 with value_size from the functor).
//
// Here the accumulator is a per-thread array of compile-time CAPACITY (8, 32 or 64 elements, chosen from value_count at
// launch) so the inner loop never touches shared memory; whole arrays are exchanged by shuffles for the warp combine
// (join() is called on complete arrays: it need not be element-wise), one shared-memory hop joins the warps, CTAs publish
// their partial arrays and a ticket elects the last CTA to fold them.
// value_count > 64 (any length, as in the reference): the accumulators live in GLOBAL memory, one array per thread of a grid sized so
// that they fit a scratch arena (array_*_big_kernel below); the block combine is a tree over those arrays.
#ifndef KB200_IMPL_ARRAYREDUCEKERNEL_HPP
#define KB200_IMPL_ARRAYREDUCEKERNEL_HPP

#include "Collectives.hpp"
#include "HostRuntime.hpp"
#include "MDRangeKernel.hpp"

namespace kb200 {
namespace Impl {

template <class F, class T, class = void> struct has_array_init : std::false_type {};
template <class F, class T> struct has_array_init<F, T, std::void_t<decltype(std::declval<const F&>().init(std::declval<T*>()))>> : std::true_type {};
template <class F, class T, class = void> struct has_array_join : std::false_type {};
template <class F, class T>
struct has_array_join<F, T, std::void_t<decltype(std::declval<const F&>().join(std::declval<T*>(), std::declval<const T*>()))>> : std::true_type {};
template <class F, class T, class = void> struct has_array_final : std::false_type {};
template <class F, class T> struct has_array_final<F, T, std::void_t<decltype(std::declval<const F&>().final(std::declval<T*>()))>> : std::true_type {};

template <class F, class T>
struct ArrayOps {
  KB200_DEVICE_FUNCTION static void init(const F& f, T* a, int count) {
    if constexpr (has_array_init<F, T>::value) f.init(a);
    else for (int c = 0; c < count; ++c) a[c] = T();
  }
  KB200_DEVICE_FUNCTION static void join(const F& f, T* d, const T* s, int count) {
    if constexpr (has_array_join<F, T>::value) f.join(d, s);
    else for (int c = 0; c < count; ++c) d[c] += s[c];
  }
  KB200_DEVICE_FUNCTION static void final(const F& f, T* a) {
    if constexpr (has_array_final<F, T>::value) f.final(a);
  }
};

// block-wide combine of per-thread arrays; thread 0 ends with the block's array.  smem: 32 * CAP elements.
template <class F, class T, int CAP>
KB200_DEVICE_FUNCTION void array_block_reduce(const F& f, T (&acc)[CAP], int count, T* smem) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
  const int live = (nthreads - (warp << 5)) < 32 ? (nthreads - (warp << 5)) : 32;  // lanes of this warp that exist
  const unsigned mask = live >= 32 ? kFullMask : ((1u << live) - 1u);
  T tmp[CAP];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
    for (int c = 0; c < CAP; ++c)
      if (c < count) tmp[c] = ::kb200::Impl::shfl_down(acc[c], d, mask);
    if (lane + d < live) ArrayOps<F, T>::join(f, acc, tmp, count);  // never join a lane that does not exist
  }
  if (nwarps == 1) return;
  if (lane == 0)
    for (int c = 0; c < count; ++c) smem[warp * CAP + c] = acc[c];
  __syncthreads();
  if (warp == 0) {
    ArrayOps<F, T>::init(f, acc, count);
    if (lane < nwarps)
      for (int c = 0; c < count; ++c) acc[c] = smem[lane * CAP + c];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
      for (int c = 0; c < CAP; ++c)
        if (c < count) tmp[c] = ::kb200::Impl::shfl_down(acc[c], d);
      if (lane + d < 32) ArrayOps<F, T>::join(f, acc, tmp, count);
    }
  }
}

// grid-wide: publish the block's array, the last block folds all of them, applies final() and stores `count` values
template <class F, class T, int CAP>
KB200_DEVICE_FUNCTION void array_grid_reduce_and_store(const F& f, T (&acc)[CAP], int count, T* partials, unsigned* ticket, T* result, T* smem) {
  __shared__ bool is_last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const unsigned nblocks = gridDim.x;
  if (nblocks > 1) {
    if (tid == 0) {
      for (int c = 0; c < count; ++c) partials[(size_t)blockIdx.x * count + c] = acc[c];
      __threadfence();
      is_last = (atomicAdd(ticket, 1u) == nblocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    ArrayOps<F, T>::init(f, acc, count);
    T tmp[CAP];
    for (unsigned b = tid; b < nblocks; b += nthreads) {
      for (int c = 0; c < count; ++c) tmp[c] = load_cg(&partials[(size_t)b * count + c]);
      ArrayOps<F, T>::join(f, acc, tmp, count);
    }
    __syncthreads();
    array_block_reduce<F, T, CAP>(f, acc, count, smem);
  }
  if (tid == 0) {
    ArrayOps<F, T>::final(f, acc);
    for (int c = 0; c < count; ++c) result[c] = acc[c];
    if (nblocks > 1) *ticket = 0u;
  }
}

template <class F, class Tag, class Index, class T, int CAP, int BLOCK>
__global__ void __launch_bounds__(BLOCK) array_range_reduce_kernel(const __grid_constant__ F f, const Index begin, const int64 n, const int count,
                                                                   T* partials, unsigned* ticket, T* result) {
  __shared__ T smem[32 * CAP];
  T acc[CAP];
  ArrayOps<F, T>::init(f, acc, count);
  for (int64 i = (int64)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (int64)gridDim.x * BLOCK) {
    if constexpr (std::is_void<Tag>::value) f((Index)(begin + (Index)i), acc);
    else f(Tag{}, (Index)(begin + (Index)i), acc);
  }
  array_block_reduce<F, T, CAP>(f, acc, count, smem);
  __syncthreads();
  array_grid_reduce_and_store<F, T, CAP>(f, acc, count, partials, ticket, result, smem);
}

template <class F, class Tag, class T, int CAP, int RANK, class Index>
__global__ void array_mdrange_reduce_kernel(const __grid_constant__ F f, const __grid_constant__ MDParams<RANK, Index> p, const int count,
                                            T* partials, unsigned* ticket, T* result) {
  __shared__ T smem[32 * CAP];
  T acc[CAP];
  ArrayOps<F, T>::init(f, acc, count);
  T* accp = acc;
  md_walk<RANK, 1, Index>(p, [&](const Index* idx) { md_invoke<Tag>(f, idx, std::make_index_sequence<RANK>{}, accp); });
  array_block_reduce<F, T, CAP>(f, acc, count, smem);
  __syncthreads();
  array_grid_reduce_and_store<F, T, CAP>(f, acc, count, partials, ticket, result, smem);
}

// ---- value_count above the register-array capacities: accumulator arrays in global memory ----------------------------------
// slabs: [gridDim.x][threads per block][count].  After the work loop: tree-join inside the block (whole arrays, through
// pointers), thread 0's array is the block partial; the last block (ticket) joins the block partials the same way.
template <class F, class T>
KB200_DEVICE_FUNCTION void array_big_finish(const F& f, T* slabs, int count, unsigned* ticket, T* result) {
  __shared__ bool is_last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  T* const block_slabs = slabs + (size_t)blockIdx.x * nthreads * count;
  auto tree = [&]() {
    int top = 1;
    while (top < nthreads) top <<= 1;
    for (int sft = top >> 1; sft >= 1; sft >>= 1) {
      __syncthreads();
      if (tid < sft && tid + sft < nthreads) ArrayOps<F, T>::join(f, block_slabs + (size_t)tid * count, block_slabs + (size_t)(tid + sft) * count, count);
    }
    __syncthreads();
  };
  tree();
  const unsigned nblocks = gridDim.x;
  if (nblocks > 1) {
    if (tid == 0) {
      __threadfence();
      is_last = (atomicAdd(ticket, 1u) == nblocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // every thread folds a strided subset of the OTHER blocks' partial arrays (array 0 of each block) into its own array
    T* mine = block_slabs + (size_t)tid * count;
    if (tid != 0) ArrayOps<F, T>::init(f, mine, count);
    for (unsigned b = tid; b < nblocks; b += nthreads)
      if (b != blockIdx.x) ArrayOps<F, T>::join(f, mine, slabs + (size_t)b * nthreads * count, count);
    tree();
  }
  if (tid == 0) {
    ArrayOps<F, T>::final(f, block_slabs);
    for (int c = 0; c < count; ++c) result[c] = block_slabs[c];
    if (nblocks > 1) *ticket = 0u;
  }
}

template <class F, class Tag, class Index, class T>
__global__ void array_range_reduce_big_kernel(const __grid_constant__ F f, const Index begin, const int64 n, const int count, T* slabs,
                                              unsigned* ticket, T* result) {
  T* acc = slabs + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * count;
  ArrayOps<F, T>::init(f, acc, count);
  for (int64 i = (int64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64)gridDim.x * blockDim.x) {
    if constexpr (std::is_void<Tag>::value) f((Index)(begin + (Index)i), acc);
    else f(Tag{}, (Index)(begin + (Index)i), acc);
  }
  array_big_finish<F, T>(f, slabs, count, ticket, result);
}

template <class F, class Tag, class T, int RANK, class Index>
__global__ void array_mdrange_reduce_big_kernel(const __grid_constant__ F f, const __grid_constant__ MDParams<RANK, Index> p, const int count,
                                                T* slabs, unsigned* ticket, T* result) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  T* acc = slabs + ((size_t)blockIdx.x * nthreads + tid) * count;
  ArrayOps<F, T>::init(f, acc, count);
  md_walk<RANK, 1, Index>(p, [&](const Index* idx) { md_invoke<Tag>(f, idx, std::make_index_sequence<RANK>{}, acc); });
  array_big_finish<F, T>(f, slabs, count, ticket, result);
}

// how many blocks of `threads` threads may keep an accumulator array each without exceeding the slab budget
inline int array_big_grid(long long wanted, int threads, int count, size_t elem_bytes) {
  const size_t budget = (size_t)512 << 20;
  long long fit = (long long)(budget / ((size_t)threads * (size_t)count * elem_bytes));
  if (fit < 1) fit = 1;
  return (int)(wanted < 1 ? 1 : (wanted < fit ? wanted : fit));
}

template <class T, class Launch>
int array_reduce_big_run(b200_instance* inst, int count, int grid, int threads, T* result_host, T* result_dev, Launch&& launch) {
  HostRuntime rt(inst);
  int rc;
  void* slabs = nullptr;
  unsigned* ticket = nullptr;
  const size_t need = ((size_t)grid * threads + 1) * (size_t)count * sizeof(T);  // accumulators + a staging array for the result
  if ((rc = rt.reduce_scratch(need, 0, false, &slabs, &ticket, nullptr, nullptr))) return rc;
  T* stage = reinterpret_cast<T*>(slabs) + (size_t)grid * threads * count;
  T* dst = result_dev ? result_dev : stage;
  launch(reinterpret_cast<T*>(slabs), ticket, dst);
  if ((rc = rt.check_launch("kb200::array_reduce_big_kernel"))) return rc;
  if (result_host) {
    if ((rc = b200_memcpy_d2h_async(inst, result_host, dst, (size_t)count * sizeof(T)))) return rc;
    return rt.fence("kb200::parallel_reduce(value_type[]): fence to hand the array result to the host");
  }
  return 0;
}

// host side, shared by both policies: scratch, launch through `launch(cap_tag, grid-independent args...)`, result hand-back
template <class T, class Launch>
int array_reduce_run(b200_instance* inst, int count, int grid, T* result_host, T* result_dev, Launch&& launch) {
  HostRuntime rt(inst);
  if (count < 0 || count > 64) return b200_report_error(B200_EUNSUPPORTED, "kb200::parallel_reduce(value_type[]): value_count above 64 is not supported");
  int rc;
  void* partials = nullptr;
  unsigned* ticket = nullptr;
  // partials for every CTA + a device staging area for the result (the tail of the same arena)
  const size_t need = ((size_t)grid + 1) * (size_t)(count > 0 ? count : 1) * sizeof(T);
  if ((rc = rt.reduce_scratch(need, 0, false, &partials, &ticket, nullptr, nullptr))) return rc;
  T* stage = reinterpret_cast<T*>(partials) + (size_t)grid * count;
  T* dst = result_dev ? result_dev : stage;
  launch(reinterpret_cast<T*>(partials), ticket, dst);
  if ((rc = rt.check_launch("kb200::array_reduce_kernel"))) return rc;
  if (result_host) {
    if ((rc = b200_memcpy_d2h_async(inst, result_host, dst, (size_t)count * sizeof(T)))) return rc;
    return rt.fence("kb200::parallel_reduce(value_type[]): fence to hand the array result to the host");
  }
  return 0;
}

}  // namespace Impl
}  // namespace kb200
#endif
