// stencil.cu -- MDRangePolicy<Rank<3>> parallel_reduce fast path (C ABI): 7-point stencil with a
// MinMaxLoc reducer over the interior of a LayoutLeft View<double***> (config C4).
//
// Stands for   parallel_reduce(MDRangePolicy<B200,Rank<3>>({1,1,1},{n0-1,n1-1,n2-1}),
//                  KB200_LAMBDA(i,j,k, MinMaxLoc::value_type& r){ v = ...; update r with loc (i*n1+j)*n2+k },
//                  MinMaxLoc<double,int64>(result));
// replacing ParallelReduce<...,MDRangePolicy,Cuda> (core/src/Cuda/Kokkos_Cuda_Parallel_MDRange.hpp:248-497),
// whose reduce grid is capped at <=512 blocks with 64 of 256 threads active and a div/mod per element
// (impl/KokkosExp_IterateTileGPU.hpp:1206-1305).
//
// Mapping: one warp per (j,k) row, lanes along the contiguous i dimension (coalesced 8-byte loads, the
// i+-1 neighbours come from the same lines, j+-1 rows from L1 via the neighbouring warps of the block,
// k+-1 planes from L2); one integer division per ROW, none per element; persistent grid.
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/Reducers.hpp>
#include <kb200/impl/Collectives.hpp>
#include <kb200/impl/HostRuntime.hpp>

using namespace kb200;
using namespace kb200::Impl;

namespace {
using Red = MinMaxLoc<double, int64>;
using V = Red::value_type;

template <int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK) stencil7_minmaxloc_kernel(const double* __restrict__ u, double* __restrict__ vout,
                                                                    int64 n0, int64 n1, int64 n2, double c0, double c1,
                                                                    ReduceScratch scratch) {
  __shared__ __align__(16) unsigned char smem[32 * sizeof(V)];
  V dummy;
  const Red red(dummy);
  V acc;
  red.init(acc);
  const int lane = threadIdx.x & 31;
  const int64 warps_per_grid = (int64)gridDim.x * (BLOCK / 32);
  const int64 m1 = n1 - 2, m2 = n2 - 2;
  const int64 rows = m1 * m2;
  const int64 sj = n0, sk = n0 * n1;
  // consecutive warps of a block take consecutive rows (j fastest) so j+-1 neighbours share L1
  for (int64 row = (int64)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); row < rows; row += warps_per_grid) {
    const int64 k = row / m1 + 1;
    const int64 j = row - (k - 1) * m1 + 1;
    const double* c = u + j * sj + k * sk;
    const int64 locbase = j * n2 + k;  // loc = i*n1*n2 + j*n2 + k
    for (int64 i0 = 1; i0 < n0 - 1; i0 += 32 * UNROLL) {
      double ctr[UNROLL], xm[UNROLL], xp[UNROLL], ym[UNROLL], yp[UNROLL], zm[UNROLL], zp[UNROLL];
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) {
        const int64 i = i0 + q * 32 + lane;
        if (i < n0 - 1) {
          ctr[q] = c[i];
          xm[q] = c[i - 1];
          xp[q] = c[i + 1];
          ym[q] = c[i - sj];
          yp[q] = c[i + sj];
          zm[q] = c[i - sk];
          zp[q] = c[i + sk];
        }
      }
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) {
        const int64 i = i0 + q * 32 + lane;
        if (i < n0 - 1) {
          // evaluation order as written in the header, no FMA contraction
          double s = __dadd_rn(xm[q], xp[q]);
          s = __dadd_rn(s, ym[q]);
          s = __dadd_rn(s, yp[q]);
          s = __dadd_rn(s, zm[q]);
          s = __dadd_rn(s, zp[q]);
          const double v = __dadd_rn(__dmul_rn(c0, ctr[q]), __dmul_rn(c1, s));
          if (vout) vout[i + j * sj + k * sk] = v;
          const int64 loc = i * (n1 * n2) + locbase;
          if (v < acc.min_val) { acc.min_val = v; acc.min_loc = loc; }
          if (v > acc.max_val) { acc.max_val = v; acc.max_loc = loc; }
        }
      }
    }
  }
  block_reduce(red, acc, smem);
  __syncthreads();
  grid_reduce_and_store(red, acc, scratch, smem);
}
}  // namespace

extern "C" int b200_stencil7_minmaxloc_f64(b200_instance* I, const double* u, double* v_out, int64_t n0, int64_t n1, int64_t n2,
                                           double c0, double c1, b200_minmaxloc_f64* rh, b200_minmaxloc_f64* rd) {
  const char* where = "b200_stencil7_minmaxloc_f64";
  B200_CHECK_INST(I, where);
  if (n0 < 0 || n1 < 0 || n2 < 0) return b200_set_error(B200_EINVAL, where, "negative extent");
  if (!rh && !rd) return b200_set_error(B200_EINVAL, where, "no result destination");
  const bool empty = (n0 < 3 || n1 < 3 || n2 < 3);
  if (!empty && !u) return b200_set_error(B200_EINVAL, where, "u is NULL");
  constexpr int BLOCK = 256, UNROLL = 4;
  HostRuntime rt(I);
  static int bps = 0;
  if (!bps) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, stencil7_minmaxloc_kernel<BLOCK, UNROLL>, BLOCK, 0);
    if (bps < 1) bps = 1;
  }
  const int64 rows = empty ? 0 : (int64)(n1 - 2) * (n2 - 2);
  int64 blocks = (rows + BLOCK / 32 - 1) / (BLOCK / 32);
  const int cap = b200_tune("stencil.bps", 0);
  const int64 max_grid = (int64)rt.sm_count() * ((cap > 0 && cap < bps) ? cap : bps);
  const int grid = (int)(blocks < 1 ? 1 : (blocks < max_grid ? blocks : max_grid));
  ReduceScratch s;
  void *slot_dev = nullptr, *slot_host = nullptr;
  int rc;
  if ((rc = rt.reduce_scratch((size_t)grid * sizeof(V), sizeof(V), rh != nullptr, &s.partials, &s.ticket, &slot_dev, &slot_host))) return rc;
  s.result0 = rh ? slot_dev : (void*)rd;
  s.result1 = rh ? (void*)rd : nullptr;
  stencil7_minmaxloc_kernel<BLOCK, UNROLL><<<grid, BLOCK, 0, rt.stream()>>>(u, v_out, empty ? 2 : n0, empty ? 2 : n1, empty ? 2 : n2, c0, c1, s);
  if ((rc = rt.check_launch(where))) return rc;
  if (rh) {
    if ((rc = rt.fence(where))) return rc;
    memcpy(rh, slot_host, sizeof(V));
  }
  return 0;
}
