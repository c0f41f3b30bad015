// hand-written (see KokkosCore_config.h): enabled backends, forward declarations
#include <fwd/Kokkos_Fwd_SERIAL.hpp>
#include <fwd/Kokkos_Fwd_OPENMP.hpp>
#include <fwd/Kokkos_Fwd_CUDA.hpp>
