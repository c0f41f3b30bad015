// kb200/Macros.hpp -- function annotations of the B200 execution space.
// Mirrors KOKKOS_FUNCTION / KOKKOS_INLINE_FUNCTION / KOKKOS_LAMBDA as the reference defines them
// when its CUDA backend is on (core/src/setup/Kokkos_Setup_Cuda.hpp:56-66).
#ifndef KB200_MACROS_HPP
#define KB200_MACROS_HPP

#if !defined(__CUDACC__)
#error "kb200 is device code for sm_100a: compile with nvcc -gencode arch=compute_100a,code=sm_100a --extended-lambda (there is no host fallback)"
#endif

#define KB200_FUNCTION __host__ __device__
#define KB200_INLINE_FUNCTION __host__ __device__ inline
#define KB200_FORCEINLINE_FUNCTION __host__ __device__ __forceinline__
#define KB200_DEVICE_FUNCTION __device__ __forceinline__
#define KB200_DEFAULTED_FUNCTION __host__ __device__ inline
#define KB200_LAMBDA [=] __host__ __device__
#define KB200_CLASS_LAMBDA [ =, *this ] __host__ __device__

// namespace name as it appears in diagnostics (the layer is declared as `Kokkos` in KB200_AS_KOKKOS mode)
#ifdef KB200_AS_KOKKOS
#define KB200_NS_STR "Kokkos"
#else
#define KB200_NS_STR "kb200"
#endif

#include <cstdint>
#include <cstddef>

namespace kb200 {
using int64 = long long;  // same width as int64_t; matches CUDA's atomic/shuffle overloads
static_assert(sizeof(int64) == 8, "");
constexpr int kWarp = 32;
}  // namespace kb200

#endif
