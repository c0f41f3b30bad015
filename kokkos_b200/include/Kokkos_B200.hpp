// Kokkos_B200.hpp -- umbrella header of the B200-native execution space (C++ layer over include/kokkos_b200.h).
//
//   #include <Kokkos_B200.hpp>          // namespace kb200: B200, View, RangePolicy, MDRangePolicy, TeamPolicy,
//                                        // parallel_for/reduce/scan, reducers, atomic_*, deep_copy, ...
//   #define KB200_AS_KOKKOS before the include to have the layer declared as `namespace Kokkos` plus the KOKKOS_* macros,
//   so that code written against the reference's hot-path API compiles unchanged.
// Compile with: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --extended-lambda, link libkokkos_b200.so.
#ifndef KOKKOS_B200_HPP
#define KOKKOS_B200_HPP

#ifdef KB200_AS_KOKKOS
// The whole layer is DECLARED in namespace Kokkos in this translation unit (not aliased): the reference's sources reopen
// `namespace Kokkos { template <> struct reduction_identity<MyType> ... }`, which a namespace alias would not allow.
// libkokkos_b200.so exports only C symbols, so translation units with and without this macro link together.
#define kb200 Kokkos
#endif
#include <nv/target>
#include "kb200/Macros.hpp"
#include "kb200/B200.hpp"
#include "kb200/View.hpp"
#include "kb200/Policy.hpp"
#include "kb200/Reducers.hpp"
#include "kb200/Atomic.hpp"
#include "kb200/Parallel.hpp"
#include "kb200/Team.hpp"
#include "kb200/TeamMDRange.hpp"
#include "kb200/UniqueToken.hpp"
#include "kb200/StdAlgorithms.hpp"
#include "kb200/Compat.hpp"

#ifdef KB200_AS_KOKKOS
#define KOKKOS_FUNCTION KB200_FUNCTION
#define KOKKOS_INLINE_FUNCTION KB200_INLINE_FUNCTION
#define KOKKOS_FORCEINLINE_FUNCTION KB200_FORCEINLINE_FUNCTION
#define KOKKOS_LAMBDA KB200_LAMBDA
#define KOKKOS_CLASS_LAMBDA KB200_CLASS_LAMBDA
#define KOKKOS_DEFAULTED_FUNCTION KB200_DEFAULTED_FUNCTION
#define KOKKOS_IMPL_HOST_FUNCTION __host__
#define KOKKOS_IMPL_DEVICE_FUNCTION __device__
#define KOKKOS_ENABLE_CUDA_LAMBDA
#define KOKKOS_COMPILER_NVCC (__CUDACC_VER_MAJOR__ * 100 + __CUDACC_VER_MINOR__ * 10)  // as core/src/Kokkos_Macros.hpp
#define KOKKOS_IF_ON_DEVICE(CODE) NV_IF_TARGET(NV_IS_DEVICE, CODE)
#define KOKKOS_IF_ON_HOST(CODE) NV_IF_TARGET(NV_IS_HOST, CODE)
#endif

#endif
