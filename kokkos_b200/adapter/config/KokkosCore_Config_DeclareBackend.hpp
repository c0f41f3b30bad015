// hand-written (see KokkosCore_config.h): enabled backends, declarations
#include <decl/Kokkos_Declare_SERIAL.hpp>
#include <decl/Kokkos_Declare_OPENMP.hpp>
#include <decl/Kokkos_Declare_CUDA.hpp>
