// kb200/impl/MDRangeKernel.hpp -- MDRangePolicy<Rank<2..6>> parallel_for / parallel_reduce.
//
// Replaces ParallelFor/ParallelReduce<...,MDRangePolicy,Cuda> (core/src/Cuda/Kokkos_Cuda_Parallel_MDRange.hpp:
// 69-246,248-497) and the DeviceIterateTile maps (core/src/impl/KokkosExp_IterateTileGPU.hpp:71-1156,1206-1305).
//
// Mapping: a tile is a thread block whose SHAPE is the tile (blockDim.x = tile[0] along the contiguous
// dimension, blockDim.y = tile[1], blockDim.z = product of the remaining tile extents -- or, when that product
// is beyond the hardware's 64, one linear block of the same size), so the hardware's
// threadIdx supplies the in-tile coordinates and no division is executed per element.  A persistent grid
// (SMs x resident blocks) strides over the tiles; the tile coordinates advance by a pre-decomposed stride
// with carries (adds and compares only).  The reference's reduce variant instead launches <= 512 blocks
// with tile_prod of 256 threads active and pays a div/mod per rank per element.
// Thread coarsening: when the slowest dimension is long enough each thread handles C = 4 CONSECUTIVE indices of it per tile
// visit (a tile visit then covers tile[R-1]*C indices there), so the tile bookkeeping (carry chain, bounds tests, index
// construction) is paid once per 4 points and the compiler can reuse loads between neighbouring points of a stencil.
// Reductions reuse the block/grid combine of Collectives.hpp (ordered, ticketed, result to a pinned slot).
#ifndef KB200_IMPL_MDRANGEKERNEL_HPP
#define KB200_IMPL_MDRANGEKERNEL_HPP

#include "Collectives.hpp"
#include "HostRuntime.hpp"
#include <utility>

namespace kb200 {
namespace Impl {

template <int RANK, class Index>
struct MDParams {
  Index lower[RANK], upper[RANK], tile[RANK], tile_end[RANK];
  Index stride[RANK];  // gridDim.x decomposed in the mixed radix tile_end[] (dimension 0 fastest)
  long long num_tiles;
  int linear;  // 1: the block is one-dimensional and threadIdx.x enumerates the tile (dimension 0 fastest)
};

template <class Tag, class F, class Index, size_t... Is, class... Extra>
KB200_DEVICE_FUNCTION void md_invoke(const F& f, const Index* idx, std::index_sequence<Is...>, Extra&... extra) {
  if constexpr (std::is_void<Tag>::value) f(idx[Is]..., extra...);
  else f(Tag{}, idx[Is]..., extra...);
}

// walks the tiles owned by this block; calls op(idx) for every in-range point handled by this thread
template <int RANK, int C, class Index, class Op>
KB200_DEVICE_FUNCTION void md_walk(const MDParams<RANK, Index>& p, Op op) {
  // this thread's fixed offset inside any tile
  Index off[RANK];
  {
    Index zz;
    if (p.linear) {
      zz = (Index)threadIdx.x;
      off[0] = zz % p.tile[0]; zz /= p.tile[0];
      off[1] = zz % p.tile[1]; zz /= p.tile[1];
    } else {
      off[0] = (Index)threadIdx.x;
      off[1] = (Index)threadIdx.y;
      zz = (Index)threadIdx.z;
    }
#pragma unroll
    for (int d = 2; d < RANK; ++d) { off[d] = zz % p.tile[d]; zz /= p.tile[d]; }
  }
  // tile coordinates of the first tile of this block (one decomposition per block, not per element)
  Index t[RANK];
  {
    long long rem = blockIdx.x;
#pragma unroll
    for (int d = 0; d < RANK; ++d) { t[d] = (Index)(rem % p.tile_end[d]); rem /= p.tile_end[d]; }
  }
  for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    Index idx[RANK];
    bool in = true;
#pragma unroll
    for (int d = 0; d < RANK - 1; ++d) {
      idx[d] = p.lower[d] + t[d] * p.tile[d] + off[d];
      in = in && (idx[d] < p.upper[d]);
    }
    const Index last0 = p.lower[RANK - 1] + (t[RANK - 1] * p.tile[RANK - 1] + off[RANK - 1]) * (Index)C;
    if (in) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        idx[RANK - 1] = last0 + (Index)c;
        if (idx[RANK - 1] < p.upper[RANK - 1]) op(idx);
      }
    }
    // advance by the grid stride: mixed-radix add with carry
    Index carry = 0;
#pragma unroll
    for (int d = 0; d < RANK; ++d) {
      t[d] += p.stride[d] + carry;
      carry = 0;
      if (t[d] >= p.tile_end[d]) { t[d] -= p.tile_end[d]; carry = 1; }
    }
  }
}

template <class F, class Tag, int RANK, class Index, int C>
__global__ void mdrange_for_kernel(const __grid_constant__ F f, const __grid_constant__ MDParams<RANK, Index> p) {
  md_walk<RANK, C, Index>(p, [&](const Index* idx) { md_invoke<Tag>(f, idx, std::make_index_sequence<RANK>{}); });
}

template <class F, class Tag, class Red, int RANK, class Index, int C>
__global__ void mdrange_reduce_kernel(const __grid_constant__ F f, const __grid_constant__ Red red,
                                      const __grid_constant__ MDParams<RANK, Index> p, const ReduceScratch scratch) {
  using V = typename Red::value_type;
  __shared__ __align__(16) unsigned char smem[32 * sizeof(V)];
  if (p.num_tiles <= 0) return reduce_store_identity(red, scratch);
  V acc;
  red.init(acc);
  md_walk<RANK, C, Index>(p, [&](const Index* idx) { md_invoke<Tag>(f, idx, std::make_index_sequence<RANK>{}, acc); });
  block_reduce(red, acc, smem);
  __syncthreads();
  grid_reduce_and_store(red, acc, scratch, smem);
}

template <class Policy>
struct MDLaunchShape {
  static constexpr int RANK = Policy::rank;
  using Index = typename Policy::index_type;
  MDParams<RANK, Index> p;
  dim3 block;
  int threads;
  int coarsen = 1;
  static constexpr int kCoarsen = 4;
  // coarsen the slowest dimension by kCoarsen if that still leaves >= 8 tile visits per SM-resident block slot
  void maybe_coarsen(int sm_count) {
    const long long e = (long long)p.tile_end[RANK - 1];
    if (e < 2 * kCoarsen) return;
    const long long e2 = (e + kCoarsen - 1) / kCoarsen;
    const long long nt2 = p.num_tiles / e * e2;
    if (nt2 < (long long)sm_count * 64) return;
    p.tile_end[RANK - 1] = (Index)e2;
    p.num_tiles = nt2;
    coarsen = kCoarsen;
  }
  explicit MDLaunchShape(const Policy& pol) {
    long long z = 1;
    for (int d = 0; d < RANK; ++d) {
      p.lower[d] = pol.m_lower[d]; p.upper[d] = pol.m_upper[d]; p.tile[d] = pol.m_tile[d]; p.tile_end[d] = pol.m_tile_end[d];
      if (d >= 2) z *= pol.m_tile[d];
    }
    p.num_tiles = (long long)pol.m_num_tiles;
    block = dim3((unsigned)pol.m_tile[0], (unsigned)pol.m_tile[1], (unsigned)z);
    threads = (int)(block.x * block.y * block.z);
    p.linear = 0;
    // blockDim.z stops at 64: a tile whose slow extents multiply past that (the reference's default tiling of a rank-6
    // Iterate::Right policy is {2,2,2,2,2,16}, KokkosExp_MDRangePolicy.hpp:330-372) runs as a one-dimensional block that
    // decomposes threadIdx.x once per thread instead
    if (z > 64) { block = dim3((unsigned)threads, 1, 1); p.linear = 1; }
  }
  void set_grid(int grid) {
    long long rem = grid;
    for (int d = 0; d < RANK; ++d) {
      const long long e = p.tile_end[d] > 0 ? (long long)p.tile_end[d] : 1;
      p.stride[d] = (Index)(rem % e);
      rem /= e;
    }
  }
};

template <class Policy, class F>
struct MDRangeFor {
  static int run(const Policy& pol_in, const F& f) {
    if (pol_in.m_num_tiles <= 0) return 0;
    constexpr int RANK = Policy::rank;
    using Index = typename Policy::index_type;
    using Tag = typename Policy::work_tag;
    Policy pol(pol_in);
    HostRuntime rt(pol.space().impl_instance());
    for (;;) {  // a default tile the kernel's register count does not allow is shrunk until it launches
      MDLaunchShape<Policy> probe(pol);
      probe.maybe_coarsen(rt.sm_count());
      auto kp = probe.coarsen > 1 ? mdrange_for_kernel<F, Tag, RANK, Index, MDLaunchShape<Policy>::kCoarsen> : mdrange_for_kernel<F, Tag, RANK, Index, 1>;
      int fit = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, kp, probe.threads, 0);
      if (fit >= 1 || !pol.impl_shrink_default_tile()) break;
    }
    MDLaunchShape<Policy> sh(pol);
    sh.maybe_coarsen(rt.sm_count());
    auto k = sh.coarsen > 1 ? mdrange_for_kernel<F, Tag, RANK, Index, MDLaunchShape<Policy>::kCoarsen> : mdrange_for_kernel<F, Tag, RANK, Index, 1>;
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, sh.threads, 0);
    if (bps < 1) bps = 1;
    bps = pol.impl_occupancy_cap(bps);
    const long long cap = (long long)rt.sm_count() * bps * 4;  // a few waves: tiles are short
    const int grid = (int)(sh.p.num_tiles < cap ? sh.p.num_tiles : cap);
    sh.set_grid(grid);
    k<<<grid, sh.block, 0, rt.stream()>>>(f, sh.p);
    return rt.check_launch("kb200::mdrange_for_kernel");
  }
};

template <class Policy, class F, class Red>
struct MDRangeReduce {
  using V = typename Red::value_type;
  static int run(const Policy& pol_in, const F& f, const Red& red, V* result_host, V* result_dev) {
    constexpr int RANK = Policy::rank;
    using Index = typename Policy::index_type;
    using Tag = typename Policy::work_tag;
    Policy pol(pol_in);
    HostRuntime rt(pol.space().impl_instance());
    for (;;) {  // (as in MDRangeFor: shrink a default tile the kernel cannot launch with)
      MDLaunchShape<Policy> probe(pol);
      probe.maybe_coarsen(rt.sm_count());
      auto kp = probe.coarsen > 1 ? mdrange_reduce_kernel<F, Tag, Red, RANK, Index, MDLaunchShape<Policy>::kCoarsen> : mdrange_reduce_kernel<F, Tag, Red, RANK, Index, 1>;
      int fit = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, kp, probe.threads, 0);
      if (fit >= 1 || !pol.impl_shrink_default_tile()) break;
    }
    MDLaunchShape<Policy> sh(pol);
    sh.maybe_coarsen(rt.sm_count());
    auto k = sh.coarsen > 1 ? mdrange_reduce_kernel<F, Tag, Red, RANK, Index, MDLaunchShape<Policy>::kCoarsen> : mdrange_reduce_kernel<F, Tag, Red, RANK, Index, 1>;
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, sh.threads, 0);
    if (bps < 1) bps = 1;
    bps = pol.impl_occupancy_cap(bps);
    const long long cap = (long long)rt.sm_count() * bps;
    long long tiles = sh.p.num_tiles;
    const int grid = (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
    sh.set_grid(grid);
    if (tiles < 1) {  // zero-length range: the result is the identity (TestMDRangeReduce.hpp:48-63)
      sh.p.num_tiles = 0;
      sh.block = dim3(32, 1, 1);
    }
    ReduceScratch s;
    void *slot_dev = nullptr, *slot_host = nullptr;
    int rc;
    if ((rc = rt.reduce_scratch((size_t)grid * sizeof(V), sizeof(V), result_host != nullptr, &s.partials, &s.ticket, &slot_dev, &slot_host))) return rc;
    s.result0 = result_host ? slot_dev : (void*)result_dev;
    s.result1 = result_host ? (void*)result_dev : nullptr;
    k<<<grid, sh.block, 0, rt.stream()>>>(f, red, sh.p, s);
    if ((rc = rt.check_launch("kb200::mdrange_reduce_kernel"))) return rc;
    if (result_host) {
      if ((rc = rt.fence("kb200::parallel_reduce(MDRange): fence to hand the scalar result to the host"))) return rc;
      memcpy(result_host, slot_host, sizeof(V));
    }
    return 0;
  }
};

}  // namespace Impl
}  // namespace kb200
#endif
