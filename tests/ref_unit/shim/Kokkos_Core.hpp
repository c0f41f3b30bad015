// Shim used ONLY by tests/ref_unit: lets the reference's own unit-test sources (compiled unmodified from
// /root/reference/core/unit_test) see the B200 execution space under the name `Kokkos`.
#ifndef KB200_REF_UNIT_SHIM_KOKKOS_CORE_HPP
#define KB200_REF_UNIT_SHIM_KOKKOS_CORE_HPP
#define KB200_AS_KOKKOS
#include <Kokkos_B200.hpp>
#endif
