/* Hand-written stand-in for the header the reference's cmake would generate
 * (template: cmake/KokkosCore_config.h.in).  Test infrastructure only: it lets
 * oracle/Makefile compile the UNMODIFIED reference sources in place under
 * /root/reference with plain g++ (OpenMP + Serial host backends, C++17). */
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
#define KOKKOS_ENABLE_CXX17
#define KOKKOS_ENABLE_LIBDL
#define KOKKOS_ENABLE_IMPL_MDSPAN
#define KOKKOS_ENABLE_IMPL_REF_COUNT_BRANCH_UNLIKELY
