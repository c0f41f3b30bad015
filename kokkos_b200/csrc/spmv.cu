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

using namespace kb200;
using namespace kb200::Impl;

namespace {
template <int BLOCK, int VL>
__global__ void __launch_bounds__(BLOCK) spmv_crs_kernel(int64 nrows, const int64* __restrict__ row_map,
                                                         const int* __restrict__ col_idx, const double* __restrict__ values,
                                                         const double* __restrict__ x, double* __restrict__ y) {
  const int sub = threadIdx.x % VL;
  const int64 groups_per_grid = (int64)gridDim.x * (BLOCK / VL);
  for (int64 row = (int64)blockIdx.x * (BLOCK / VL) + threadIdx.x / VL; row < nrows; row += groups_per_grid) {
    const int64 kb = __ldg(row_map + row), ke = __ldg(row_map + row + 1);
    double acc = 0.0;
    for (int64 k = kb + sub; k < ke; k += VL) acc = __dadd_rn(acc, __dmul_rn(__ldg(values + k), __ldg(x + __ldg(col_idx + k))));
#pragma unroll
    for (int m = VL / 2; m > 0; m >>= 1) acc = __dadd_rn(acc, shfl_xor(acc, m));
    if (sub == 0) y[row] = acc;
  }
}

template <int VL>
int launch(b200_instance* I, int64 nrows, const int64* row_map, const int* col_idx, const double* values, const double* x, double* y) {
  constexpr int BLOCK = 256;
  HostRuntime rt(I);
  static int bps = 0;
  if (!bps) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, spmv_crs_kernel<BLOCK, VL>, BLOCK, 0);
    if (bps < 1) bps = 1;
  }
  int64 blocks = (nrows + BLOCK / VL - 1) / (BLOCK / VL);
  const int64 max_grid = (int64)rt.sm_count() * bps * 4;  // a few waves: rows differ in length
  const int grid = (int)(blocks < max_grid ? blocks : max_grid);
  spmv_crs_kernel<BLOCK, VL><<<grid, BLOCK, 0, rt.stream()>>>(nrows, row_map, col_idx, values, x, y);
  return rt.check_launch("b200_spmv_crs_f64");
}
}  // namespace

extern "C" int b200_spmv_crs_f64(b200_instance* I, int64_t nrows, const int64_t* row_map, const int32_t* col_idx,
                                 const double* values, const double* x, double* y) {
  const char* where = "b200_spmv_crs_f64";
  B200_CHECK_INST(I, where);
  if (nrows < 0) return b200_set_error(B200_EINVAL, where, "negative row count");
  if (nrows == 0) return 0;
  if (!row_map || !y) return b200_set_error(B200_EINVAL, where, "NULL array");
  // the vector length is a launch-time choice (TeamPolicy's third argument); the mean row length is
  // read back once per call -- a View-metadata query in the reference's terms (Crs::numRows / nnz).
  int64_t ends[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(&ends[0], row_map, 8, cudaMemcpyDeviceToHost, (cudaStream_t)b200_instance_stream(I));
  if (e == cudaSuccess) e = cudaMemcpyAsync(&ends[1], row_map + nrows, 8, cudaMemcpyDeviceToHost, (cudaStream_t)b200_instance_stream(I));
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)b200_instance_stream(I));
  if (e != cudaSuccess) return b200_set_error((int)e, where, "row_map read-back");
  const int64 nnz = ends[1] - ends[0];
  if (nnz > 0 && (!col_idx || !values || !x)) return b200_set_error(B200_EINVAL, where, "NULL array");
  int vl = b200_tune("spmv.vl", 0);
  if (vl == 0) {
    const double mean = (double)nnz / (double)nrows;
    vl = mean <= 6 ? 4 : mean <= 12 ? 8 : mean <= 24 ? 16 : 32;
  }
  const int64* rm = (const int64*)row_map;
  switch (vl) {
    case 4: return launch<4>(I, nrows, rm, col_idx, values, x, y);
    case 8: return launch<8>(I, nrows, rm, col_idx, values, x, y);
    case 16: return launch<16>(I, nrows, rm, col_idx, values, x, y);
    case 32: return launch<32>(I, nrows, rm, col_idx, values, x, y);
  }
  return b200_set_error(B200_EUNSUPPORTED, where, "vector length must be 4, 8, 16 or 32");
}
