// shim: <Kokkos_Pair.hpp> for the reference's test sources -> kb200::pair (kb200/Compat.hpp)
#include <Kokkos_Core.hpp>
