// kb200/impl/ReduceKernel.hpp -- the one grid-wide reduction skeleton of the B200 execution space.
//
// Replaces ParallelReduce<CombinedFunctorReducer,RangePolicy,Cuda>::operator()/execute
// (core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:191-259,290-368) and is reused by the MDRange and
// the typed fast paths.  Differences by design:
//   * a persistent grid (SMs x resident blocks) walks block-interleaved tiles of BLOCK*UNROLL
//     "units"; all UNROLL loads of a thread are issued before the first is consumed, so each
//     thread keeps UNROLL independent (up to 32-byte) requests in flight;
//   * partials never touch shared memory until the single cross-warp hop (Collectives.hpp);
//   * the result goes straight to a pinned, device-mapped host slot followed by a completion word; the host polls
//     that word: no stream synchronisation and no memcpy (the reference: unified scratch + fence + copy, :344-360).
//
// Body concept (device side):
//   using packet = ...;                                  // what load() returns (may be empty)
//   packet load(int64 u) const;                          // issue the memory request(s) of unit u
//   void   consume(const packet&, int64 u, V& acc) const;// fold unit u into the accumulator
//   int64  edge_count() const;                           // un-vectorisable head/tail elements
//   void   edge(int64 k, V& acc) const;
#ifndef KB200_IMPL_REDUCEKERNEL_HPP
#define KB200_IMPL_REDUCEKERNEL_HPP

#include "Collectives.hpp"
#include "HostRuntime.hpp"

namespace kb200 {
namespace Impl {

template <class Body, class Red, int BLOCK, int UNROLL, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
    range_reduce_kernel(const __grid_constant__ Body body, const __grid_constant__ Red red, const int64 n_units,
                        const ReduceScratch scratch) {
  using V = typename Red::value_type;
  __shared__ __align__(16) unsigned char smem[32 * sizeof(V)];
  if (n_units <= 0 && body.edge_count() <= 0) return reduce_store_identity(red, scratch);  // (grid is one block then)
  V acc;
  red.init(acc);

  constexpr int64 TILE = (int64)BLOCK * UNROLL;
  const int64 full_tiles = n_units / TILE;
  for (int64 tile = blockIdx.x; tile < full_tiles; tile += gridDim.x) {
    const int64 base = tile * TILE + threadIdx.x;
    typename Body::packet p[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) p[j] = body.load(base + (int64)j * BLOCK);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) body.consume(p[j], base + (int64)j * BLOCK, acc);
  }
  // ragged last tile: handled by the block that would own tile `full_tiles`
  if ((int64)blockIdx.x == full_tiles % gridDim.x) {
    for (int64 u = full_tiles * TILE + threadIdx.x; u < n_units; u += BLOCK) {
      typename Body::packet p = body.load(u);
      body.consume(p, u, acc);
    }
  }
  const int64 edges = body.edge_count();
  for (int64 k = (int64)blockIdx.x * BLOCK + threadIdx.x; k < edges; k += (int64)gridDim.x * BLOCK) body.edge(k, acc);

  block_reduce(red, acc, smem);
  __syncthreads();
  grid_reduce_and_store(red, acc, scratch, smem);
}

// Host side: occupancy-sized persistent launch + result hand-back.
template <class Body, class Red, int BLOCK = 256, int UNROLL = 4, int MIN_BLOCKS = 1>
struct RangeReduceLaunch {
  using V = typename Red::value_type;

  static int resident_blocks_per_sm() {
    static PerDeviceInt cache;  // per instantiation and device, like the reference's func-attr cache (KernelLaunch.hpp:131-145)
    int& cached = cache.here();
    if (cached == 0) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, range_reduce_kernel<Body, Red, BLOCK, UNROLL, MIN_BLOCKS>, BLOCK, 0);
      cached = nb > 0 ? nb : 1;
    }
    return cached;
  }

  // blocks_per_sm_cap <= 0: use every resident slot
  static int run(b200_instance* inst, const Body& body, const Red& red, int64 n_units, V* result_host, V* result_dev,
                 int blocks_per_sm_cap = 0) {
    static_assert(sizeof(Body) + sizeof(Red) <= 32000, "closure exceeds the kernel parameter space");
    HostRuntime rt(inst);
    int bps = resident_blocks_per_sm();
    if (blocks_per_sm_cap > 0 && blocks_per_sm_cap < bps) bps = blocks_per_sm_cap;
    constexpr int64 TILE = (int64)BLOCK * UNROLL;
    int64 tiles = (n_units + TILE - 1) / TILE;
    int64 max_grid = (int64)rt.sm_count() * bps;
    int grid = (int)(tiles < 1 ? 1 : (tiles < max_grid ? tiles : max_grid));

    ReduceScratch s;
    void* slot_dev = nullptr;
    void* slot_host = nullptr;
    int rc;
    if ((rc = rt.reduce_scratch((size_t)grid * sizeof(V), sizeof(V), false, &s.partials, &s.ticket, nullptr, nullptr))) return rc;
    // scalar result: pinned slot + completion word polled by the host (no stream synchronisation on this path)
    if (result_host && (rc = rt.result_slot(sizeof(V), &slot_dev, &slot_host, &s.seq_ptr, &s.seq_val))) return rc;
    s.result0 = result_host ? slot_dev : (void*)result_dev;
    s.result1 = result_host ? (void*)result_dev : nullptr;
    range_reduce_kernel<Body, Red, BLOCK, UNROLL, MIN_BLOCKS><<<grid, BLOCK, 0, rt.stream()>>>(body, red, n_units, s);
    if ((rc = rt.check_launch("kb200::range_reduce_kernel"))) return rc;
    if (result_host) {
      if ((rc = rt.result_wait(slot_host, s.seq_val, "kb200::parallel_reduce: wait for the scalar result"))) return rc;
      memcpy(result_host, slot_host, sizeof(V));
    }
    return 0;
  }
};

}  // namespace Impl
}  // namespace kb200
#endif
