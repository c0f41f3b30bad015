// scan.cu -- typed parallel_scan fast paths (C ABI) over contiguous View<T*> data:
// instantiations of the single-pass TMA + decoupled-look-back kernel (kb200/impl/ScanContig.hpp).
// The functor they stand for is the canonical prefix-sum functor of
// core/unit_test/TestParallelScanRangePolicy.hpp:41-84 / incremental/Test16_ParallelScan.hpp:94-143.
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/impl/ScanContig.hpp>

using namespace kb200;
using kb200::Impl::ContigScanLaunch;

namespace {
template <class T, bool INCL>
int scan_entry(b200_instance* I, const char* where, const T* x, T* y, int64_t n, T seed, const T* seed_dev, T* th, T* td) {
  B200_CHECK_INST(I, where);
  if (n < 0) return b200_set_error(B200_EINVAL, where, "negative length");
  if (n > 0 && (!x || !y)) return b200_set_error(B200_EINVAL, where, "x or y is NULL");
  const int nv = b200_tune("scan.nv", 9), nbuf = b200_tune("scan.nbuf", 2), lbw = b200_tune("scan.lbw", 2);
  const int block = b200_tune("scan.block", 256), bps = b200_tune("scan.bps", 0);
#define CFG(BL, NV, NB, LB) \
  if (block == BL && nv == NV && nbuf == NB && lbw == LB) return ContigScanLaunch<T, BL, NV, NB, LB, INCL>::run(I, x, y, n, seed, seed_dev, th, td, bps);
  CFG(256, 9, 2, 2)
#ifdef B200_SWEEP
  if constexpr (sizeof(T) == 8 && !INCL) {
    CFG(256, 9, 2, 1) CFG(256, 9, 2, 4) CFG(256, 9, 3, 2) CFG(256, 9, 3, 4)
    CFG(256, 7, 2, 2) CFG(256, 7, 3, 2) CFG(256, 7, 3, 4) CFG(256, 11, 2, 2) CFG(256, 11, 2, 4)
    CFG(512, 9, 2, 2) CFG(512, 9, 2, 4) CFG(512, 7, 2, 2) CFG(512, 7, 3, 4) CFG(512, 5, 3, 4) CFG(512, 5, 2, 2)
    CFG(128, 9, 2, 2) CFG(128, 9, 3, 2) CFG(128, 11, 3, 2) CFG(128, 13, 2, 2) CFG(128, 13, 3, 4)
    CFG(256, 13, 2, 4) CFG(256, 13, 2, 2) CFG(1024, 5, 2, 4) CFG(1024, 7, 2, 4) CFG(1024, 3, 3, 4)
  }
#endif
#undef CFG
  return b200_set_error(B200_EUNSUPPORTED, where, "tuning combination not compiled in");
}
}  // namespace

extern "C" {
int b200_scan_excl_i64(b200_instance* I, const int64_t* x, int64_t* y, int64_t n, int64_t seed, int64_t* th, int64_t* td) {
  return scan_entry<int64, false>(I, "b200_scan_excl_i64", (const int64*)x, (int64*)y, n, (int64)seed, nullptr, (int64*)th, (int64*)td);
}
int b200_scan_incl_i64(b200_instance* I, const int64_t* x, int64_t* y, int64_t n, int64_t seed, int64_t* th, int64_t* td) {
  return scan_entry<int64, true>(I, "b200_scan_incl_i64", (const int64*)x, (int64*)y, n, (int64)seed, nullptr, (int64*)th, (int64*)td);
}
int b200_scan_excl_f64(b200_instance* I, const double* x, double* y, int64_t n, double seed, double* th, double* td) {
  return scan_entry<double, false>(I, "b200_scan_excl_f64", x, y, n, seed, nullptr, th, td);
}
int b200_scan_incl_f64(b200_instance* I, const double* x, double* y, int64_t n, double seed, double* th, double* td) {
  return scan_entry<double, true>(I, "b200_scan_incl_f64", x, y, n, seed, nullptr, th, td);
}
int b200_scan_excl_i32(b200_instance* I, const int32_t* x, int32_t* y, int64_t n, int32_t seed, int32_t* th, int32_t* td) {
  return scan_entry<int, false>(I, "b200_scan_excl_i32", x, y, n, seed, nullptr, th, td);
}
int b200_scan_excl_i64_seed_dev(b200_instance* I, const int64_t* x, int64_t* y, int64_t n, const int64_t* seed_dev, int64_t* td) {
  if (!seed_dev) return b200_set_error(B200_EINVAL, "b200_scan_excl_i64_seed_dev", "seed_dev is NULL");
  return scan_entry<int64, false>(I, "b200_scan_excl_i64_seed_dev", (const int64*)x, (int64*)y, n, 0, (const int64*)seed_dev, nullptr, (int64*)td);
}
}  // extern "C"
