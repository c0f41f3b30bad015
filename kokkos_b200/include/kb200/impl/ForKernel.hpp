// kb200/impl/ForKernel.hpp -- the elementwise (parallel_for) skeleton of the B200 execution space.
//
// Replaces ParallelFor<F,RangePolicy,Cuda>::operator()/execute
// (core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:72-112: one element per thread, 8-byte accesses,
// ~10^6 tiny blocks for N=2^28).  Here each thread handles UNROLL units per tile with all loads
// issued before the first store; a typed body moves 32 bytes per unit (LDG/STG.E.ENL2.256).
//
// Body concept (device side):
//   using packet = ...;
//   packet load(int64 u) const;                // issue the loads of unit u
//   void   store(const packet&, int64 u) const;// compute + store unit u
//   int64  edge_count() const;   void edge(int64 k) const;   // scalar head/tail elements
#ifndef KB200_IMPL_FORKERNEL_HPP
#define KB200_IMPL_FORKERNEL_HPP

#include "HostRuntime.hpp"

namespace kb200 {
namespace Impl {

template <class Body, int BLOCK, int UNROLL>
KB200_DEVICE_FUNCTION void range_for_tiles(const Body& body, const int64 n_units) {
  constexpr int64 TILE = (int64)BLOCK * UNROLL;
  const int64 full_tiles = n_units / TILE;
  for (int64 tile = blockIdx.x; tile < full_tiles; tile += gridDim.x) {
    const int64 base = tile * TILE + threadIdx.x;
    typename Body::packet p[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) p[j] = body.load(base + (int64)j * BLOCK);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) body.store(p[j], base + (int64)j * BLOCK);
  }
  if ((int64)blockIdx.x == full_tiles % gridDim.x) {
    for (int64 u = full_tiles * TILE + threadIdx.x; u < n_units; u += BLOCK) {
      typename Body::packet p = body.load(u);
      body.store(p, u);
    }
  }
  const int64 edges = body.edge_count();
  for (int64 k = (int64)blockIdx.x * BLOCK + threadIdx.x; k < edges; k += (int64)gridDim.x * BLOCK) body.edge(k);
}

template <class Body, int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK) range_for_kernel(const __grid_constant__ Body body, const int64 n_units) {
  range_for_tiles<Body, BLOCK, UNROLL>(body, n_units);
}
// closures beyond the kernel parameter space (the reference's "global memory launch", Cuda/Kokkos_Cuda_KernelLaunch.hpp:317-420):
// the closure is copied into the instance's functor scratch in stream order and the kernel reads it from there
template <class Body, int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK) range_for_kernel_global(const Body* __restrict__ body, const int64 n_units) {
  range_for_tiles<Body, BLOCK, UNROLL>(*body, n_units);
}

template <class Body, int BLOCK = 256, int UNROLL = 4>
struct RangeForLaunch {
  static int resident_blocks_per_sm() {
    static PerDeviceInt cache;
    int& cached = cache.here();
    if (cached == 0) {
      int nb = 0;
      if constexpr (sizeof(Body) > 32000) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, range_for_kernel_global<Body, BLOCK, UNROLL>, BLOCK, 0);
      else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, range_for_kernel<Body, BLOCK, UNROLL>, BLOCK, 0);
      cached = nb > 0 ? nb : 1;
    }
    return cached;
  }
  // blocks_per_sm_cap <= 0: one block per tile (plain grid); > 0: persistent grid of SMs x min(resident, cap) x waves blocks.
  // waves > 1 oversubscribes the machine a little: each block still walks many tiles, but the hardware block scheduler evens out
  // what a fixed tile->block assignment cannot when iterations differ in latency (random atomics: the last blocks of a one-wave grid
  // finish 9% late -- profiles/r02_for_waves.log).
  static int run(b200_instance* inst, const Body& body, int64 n_units, int blocks_per_sm_cap = 0, int waves = 1) {
    HostRuntime rt(inst);
    constexpr int64 TILE = (int64)BLOCK * UNROLL;
    constexpr bool kGlobal = sizeof(Body) > 32000;  // beyond the kernel parameter space (32 764 bytes on sm_100)
    int64 tiles = (n_units + TILE - 1) / TILE;
    if (tiles < 1) {
      if (body.edge_count() == 0) return 0;  // empty range: nothing to launch
      tiles = 1;
    }
    int64 grid = tiles;
    if (blocks_per_sm_cap > 0) {
      int bps = resident_blocks_per_sm();
      if (blocks_per_sm_cap < bps) bps = blocks_per_sm_cap;
      const int64 cap = (int64)rt.sm_count() * bps * (waves > 1 ? waves : 1);
      if (grid > cap) grid = cap;
    }
    if (grid > 0x7fffffffll) grid = 0x7fffffffll;
    if constexpr (kGlobal) {
      void* dev = nullptr;
      int rc;
      if ((rc = b200_scratch_get(inst, B200_SCRATCH_FUNCTOR, sizeof(Body), &dev, nullptr))) return rc;
      if ((rc = b200_memcpy_h2d_async(inst, dev, &body, sizeof(Body)))) return rc;  // pageable source: staged before the call returns
      range_for_kernel_global<Body, BLOCK, UNROLL><<<(unsigned)grid, BLOCK, 0, rt.stream()>>>(reinterpret_cast<const Body*>(dev), n_units);
    } else {
      range_for_kernel<Body, BLOCK, UNROLL><<<(unsigned)grid, BLOCK, 0, rt.stream()>>>(body, n_units);
    }
    return rt.check_launch("kb200::range_for_kernel");
  }
};

}  // namespace Impl
}  // namespace kb200
#endif
