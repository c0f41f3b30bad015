/* Hand-written stand-in for the header the reference's cmake generates (template: cmake/KokkosCore_config.h.in):
 * the configuration -DKokkos_ENABLE_CUDA=ON -DKokkos_ENABLE_OPENMP=ON -DKokkos_ENABLE_SERIAL=ON -DKokkos_ARCH_BLACKWELL100=ON
 * of the UNMODIFIED reference, so it can be compiled in place with plain nvcc (no cmake).
 * Kokkos::B200 (../Kokkos_B200_Space.hpp) is added on top of this configuration; Kokkos::Cuda stays in the binary as the
 * GPU comparator and Kokkos::OpenMP as the CPU checker. */
#if !defined(KOKKOS_MACROS_HPP) || defined(KOKKOS_CORE_CONFIG_H)
#error "include Kokkos_Macros.hpp, not KokkosCore_config.h"
#else
#define KOKKOS_CORE_CONFIG_H
#endif
#define KOKKOS_VERSION 40699
#define KOKKOS_VERSION_MAJOR 4
#define KOKKOS_VERSION_MINOR 6
#define KOKKOS_VERSION_PATCH 99
#define KOKKOS_ENABLE_SERIAL
#define KOKKOS_ENABLE_OPENMP
#define KOKKOS_ENABLE_CUDA
#define KOKKOS_ENABLE_CUDA_LAMBDA
#define KOKKOS_ENABLE_CUDA_CONSTEXPR
#define KOKKOS_ENABLE_CXX17
#define KOKKOS_ENABLE_LIBDL
#define KOKKOS_ENABLE_IMPL_MDSPAN
#define KOKKOS_ENABLE_IMPL_REF_COUNT_BRANCH_UNLIKELY
#define KOKKOS_ARCH_BLACKWELL
#define KOKKOS_ARCH_BLACKWELL100
