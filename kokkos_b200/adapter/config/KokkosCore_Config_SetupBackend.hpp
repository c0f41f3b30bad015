// hand-written (see KokkosCore_config.h): device backend set-up (function annotations, KOKKOS_LAMBDA)
#include <setup/Kokkos_Setup_Cuda.hpp>
