// hand-written (see KokkosCore_config.h): enabled host backends, declarations
#include <decl/Kokkos_Declare_SERIAL.hpp>
#include <decl/Kokkos_Declare_OPENMP.hpp>
