// Kokkos_B200.hpp -- umbrella header of the B200-native execution space (C++ layer over include/kokkos_b200.h).
//
//   #include <Kokkos_B200.hpp>          // namespace kb200: B200, View, RangePolicy, MDRangePolicy, TeamPolicy,
//                                        // parallel_for/reduce/scan, reducers, atomic_*, deep_copy, ...
//   #define KB200_AS_KOKKOS before the include to also get `namespace Kokkos = kb200;` and the KOKKOS_* macros,
//   so that code written against the reference's hot-path API compiles unchanged.
// Compile with: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --extended-lambda, link libkokkos_b200.so.
#ifndef KOKKOS_B200_HPP
#define KOKKOS_B200_HPP

#include "kb200/Macros.hpp"
#include "kb200/B200.hpp"
#include "kb200/View.hpp"
#include "kb200/Policy.hpp"
#include "kb200/Reducers.hpp"
#include "kb200/Atomic.hpp"
#include "kb200/Parallel.hpp"
#include "kb200/Team.hpp"
#include "kb200/StdAlgorithms.hpp"
#include "kb200/Compat.hpp"

#ifdef KB200_AS_KOKKOS
namespace Kokkos = kb200;
#define KOKKOS_FUNCTION KB200_FUNCTION
#define KOKKOS_INLINE_FUNCTION KB200_INLINE_FUNCTION
#define KOKKOS_FORCEINLINE_FUNCTION KB200_FORCEINLINE_FUNCTION
#define KOKKOS_LAMBDA KB200_LAMBDA
#define KOKKOS_CLASS_LAMBDA KB200_CLASS_LAMBDA
#endif

#endif
