// spmv.cu -- TeamPolicy nested-reduce CRS SpMV fast path (C ABI), y = A x  (config C5b).
//
// Stands for the hierarchical pattern of
// example/tutorial/Hierarchical_Parallelism/03_vectorization/vectorization.cpp:51-76:
//   parallel_for(TeamPolicy<B200>(rows/rows_per_team, AUTO, VL), KB200_LAMBDA(member){
//     parallel_for(TeamThreadRange(member, rows_per_team), [&](int r){
//       double s; parallel_reduce(ThreadVectorRange(member, row_map(row), row_map(row+1)),
//                                 [&](int64 k, double& u){ u += values(k)*x(col_idx(k)); }, s);
//       single(PerThread(member), [&]{ y(row) = s; }); }); });
// replacing ParallelFor<F,TeamPolicy,Cuda> + CudaTeamMember::vector_reduce
// (core/src/Cuda/Kokkos_Cuda_Parallel_Team.hpp:431-587, Kokkos_Cuda_Team.hpp:299-334,721-751).
//
// Mapping: a "thread" of the team is a group of VL lanes (VL = 4..32 picked from the mean row length),
// lanes stride the row's nonzeros (coalesced values/col_idx), x is gathered through the read-only path,
// and the vector reduce is a log2(VL) xor-shuffle tree.  Products and sums are not contracted to FMA.
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/impl/HostRuntime.hpp>
#include <kb200/impl/Collectives.hpp>
#include <mutex>

using namespace kb200;
using namespace kb200::Impl;

namespace {
// UR consecutive rows per lane group and iteration (independent row_map / col_idx / x requests in flight).  Measured on B200
// (profiles/r01_spmv_probe.log): UR > 1 does NOT help -- the kernel is bound by L1 wavefronts of the x gather (one 32-byte
// sector per lane for random columns, ~1 wavefront/clk/SM => ~0.48 ms for 2^27 gathers), not by requests in flight; the
// shipped configuration is UR = 1 and 16 lanes per row at 32 nnz/row (3.1 TB/s of algorithmic bytes).
// Per-row arithmetic is unchanged: lane `sub` folds nonzeros sub, sub+VL, ... left to right, then a log2(VL) xor tree.
template <int BLOCK, int VL, int UR>
__global__ void __launch_bounds__(BLOCK) spmv_crs_kernel(int64 nrows, const int64* __restrict__ row_map,
                                                         const int* __restrict__ col_idx, const double* __restrict__ values,
                                                         const double* __restrict__ x, double* __restrict__ y) {
  const int sub = threadIdx.x % VL;
  const int64 groups_per_grid = (int64)gridDim.x * (BLOCK / VL);
  for (int64 row0 = ((int64)blockIdx.x * (BLOCK / VL) + threadIdx.x / VL) * UR; row0 < nrows; row0 += groups_per_grid * UR) {
    int64 kb[UR], ke[UR];
#pragma unroll
    for (int q = 0; q < UR; ++q) {
      const int64 r = row0 + q < nrows ? row0 + q : nrows - 1;
      kb[q] = __ldg(row_map + r);
      ke[q] = row0 + q < nrows ? __ldg(row_map + r + 1) : kb[q];  // rows past the end are empty
    }
    double acc[UR];
    int64 longest = 0;
#pragma unroll
    for (int q = 0; q < UR; ++q) { acc[q] = 0.0; longest = ke[q] - kb[q] > longest ? ke[q] - kb[q] : longest; }
    for (int64 t = sub; t < longest; t += VL) {
      int c[UR];
      double v[UR];
#pragma unroll
      for (int q = 0; q < UR; ++q) {
        const bool in = kb[q] + t < ke[q];
        c[q] = in ? __ldg(col_idx + kb[q] + t) : -1;
        v[q] = in ? __ldg(values + kb[q] + t) : 0.0;
      }
      double xv[UR];
#pragma unroll
      for (int q = 0; q < UR; ++q) xv[q] = c[q] >= 0 ? __ldg(x + c[q]) : 0.0;
#pragma unroll
      for (int q = 0; q < UR; ++q)
        if (c[q] >= 0) acc[q] = __dadd_rn(acc[q], __dmul_rn(v[q], xv[q]));
    }
#pragma unroll
    for (int q = 0; q < UR; ++q) {
#pragma unroll
      for (int m = VL / 2; m > 0; m >>= 1) acc[q] = __dadd_rn(acc[q], shfl_xor(acc[q], m));
    }
    if (sub < UR && row0 + sub < nrows) {
      double out = acc[0];
#pragma unroll
      for (int q = 1; q < UR; ++q) out = sub == q ? acc[q] : out;
      y[row0 + sub] = out;  // UR consecutive rows written by UR lanes: one coalesced store
    }
  }
}

template <int VL, int UR>
int launch(b200_instance* I, int64 nrows, const int64* row_map, const int* col_idx, const double* values, const double* x, double* y) {
  constexpr int BLOCK = 256;
  HostRuntime rt(I);
  static int bps = 0;
  if (!bps) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, spmv_crs_kernel<BLOCK, VL, UR>, BLOCK, 0);
    if (bps < 1) bps = 1;
  }
  const int64 rows_per_block = (int64)(BLOCK / VL) * UR;
  int64 blocks = (nrows + rows_per_block - 1) / rows_per_block;
  const int64 max_grid = (int64)rt.sm_count() * bps * 4;  // a few waves: rows differ in length
  const int grid = (int)(blocks < max_grid ? blocks : max_grid);
  spmv_crs_kernel<BLOCK, VL, UR><<<grid, BLOCK, 0, rt.stream()>>>(nrows, row_map, col_idx, values, x, y);
  return rt.check_launch("b200_spmv_crs_f64");
}
}  // namespace

namespace {
// (instance, row_map, nrows) -> nnz, remembered so that repeated products with one matrix do not read the row map back
struct SpmvMeta { b200_instance* inst; const int64_t* rm; int64_t nrows, nnz; };
SpmvMeta g_spmv_cache[8];
int g_spmv_cache_next = 0;
std::mutex g_spmv_cache_mu;
}  // namespace

// b200_finalize: forget what was remembered for a dying instance (a later instance may reuse the address)
void b200_spmv_release(b200_instance* I) {
  std::lock_guard<std::mutex> g(g_spmv_cache_mu);
  for (SpmvMeta& m : g_spmv_cache) if (m.inst == I) m = SpmvMeta{nullptr, nullptr, 0, 0};
}

extern "C" int b200_spmv_crs_f64(b200_instance* I, int64_t nrows, const int64_t* row_map, const int32_t* col_idx,
                                 const double* values, const double* x, double* y) {
  const char* where = "b200_spmv_crs_f64";
  B200_CHECK_INST(I, where);
  if (nrows < 0) return b200_set_error(B200_EINVAL, where, "negative row count");
  if (nrows == 0) return 0;
  if (!row_map || !y) return b200_set_error(B200_EINVAL, where, "NULL array");
  // the vector length is a launch-time choice (TeamPolicy's third argument) made from the mean row length: a View-metadata
  // query in the reference's terms (Crs::numRows / nnz).  Reading it back costs a stream synchronisation, so the answer is
  // remembered per (instance, row_map, nrows): repeated products with the same matrix launch without touching the host.
  // The choice only affects speed -- every vector length is correct for every row length.
  int64 nnz = -1;
  {
    std::lock_guard<std::mutex> g(g_spmv_cache_mu);
    for (const SpmvMeta& m : g_spmv_cache) if (m.inst == I && m.rm == row_map && m.nrows == nrows) nnz = m.nnz;
  }
  if (nnz < 0) {
    int64_t ends[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(&ends[0], row_map, 8, cudaMemcpyDeviceToHost, (cudaStream_t)b200_instance_stream(I));
    if (e == cudaSuccess) e = cudaMemcpyAsync(&ends[1], row_map + nrows, 8, cudaMemcpyDeviceToHost, (cudaStream_t)b200_instance_stream(I));
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)b200_instance_stream(I));
    if (e != cudaSuccess) return b200_set_error((int)e, where, "row_map read-back");
    nnz = ends[1] - ends[0];
    std::lock_guard<std::mutex> g(g_spmv_cache_mu);
    g_spmv_cache[g_spmv_cache_next] = SpmvMeta{I, row_map, nrows, nnz};
    g_spmv_cache_next = (g_spmv_cache_next + 1) % 8;
  }
  if (nnz > 0 && (!col_idx || !values || !x)) return b200_set_error(B200_EINVAL, where, "NULL array");
  int vl = b200_tune("spmv.vl", 0);
  if (vl == 0) {
    const double mean = (double)nnz / (double)nrows;
    // B200 probe at 32 nnz/row (profiles/r02_spmv_probe.log): 8 lanes 3.80 TB/s, 16 lanes 3.32, 4 lanes 3.24, 32 lanes 2.24 -- about four
    // nonzeros per lane amortise the shuffle tree best
    vl = mean <= 6 ? 4 : mean <= 48 ? 8 : mean <= 96 ? 16 : 32;
  }
  const int64* rm = (const int64*)row_map;
  const int ur = b200_tune("spmv.ur", 1);
#define SPMV_CFG(VL, UR) if (vl == VL && ur == UR) return launch<VL, UR>(I, nrows, rm, col_idx, values, x, y);
  SPMV_CFG(4, 1) SPMV_CFG(8, 1) SPMV_CFG(16, 1) SPMV_CFG(32, 1) SPMV_CFG(8, 2) SPMV_CFG(4, 2)
#ifdef B200_SWEEP
  SPMV_CFG(4, 4) SPMV_CFG(8, 4) SPMV_CFG(16, 4) SPMV_CFG(32, 4) SPMV_CFG(32, 2) SPMV_CFG(32, 8) SPMV_CFG(16, 2) SPMV_CFG(16, 8) SPMV_CFG(8, 8)
#endif
#undef SPMV_CFG
  return b200_set_error(B200_EUNSUPPORTED, where, "vector length must be 4, 8, 16 or 32 (and spmv.ur 4)");
}
