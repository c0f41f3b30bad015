// hand-written (see KokkosCore_config.h): enabled host backends, forward decls
#include <fwd/Kokkos_Fwd_SERIAL.hpp>
#include <fwd/Kokkos_Fwd_OPENMP.hpp>
