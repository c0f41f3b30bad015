// shim: <Kokkos_Timer.hpp> for the reference's test sources -> kb200::Timer (kb200/Compat.hpp)
#include <Kokkos_Core.hpp>
