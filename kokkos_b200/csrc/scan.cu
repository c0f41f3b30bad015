// scan.cu -- typed parallel_scan fast paths (C ABI) over contiguous View<T*> data:
// instantiations of the single-pass TMA + decoupled-look-back kernel (kb200/impl/ScanContig.hpp).
// The functor they stand for is the canonical prefix-sum functor of
// core/unit_test/TestParallelScanRangePolicy.hpp:41-84 / incremental/Test16_ParallelScan.hpp:94-143.
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/impl/ScanContig.hpp>
#include <kb200/impl/ScanChunked.hpp>

using namespace kb200;
using kb200::Impl::ContigScanLaunch;

namespace {
template <class T, bool INCL>
int scan_entry(b200_instance* I, const char* where, const T* x, T* y, int64_t n, T seed, const T* seed_dev, T* th, T* td, int nseeds = 1) {
  B200_CHECK_INST(I, where);
  if (n < 0) return b200_set_error(B200_EINVAL, where, "negative length");
  if (n > 0 && (!x || !y)) return b200_set_error(B200_EINVAL, where, "x or y is NULL");
  const int nv = b200_tune("scan.nv", 9), nbuf = b200_tune("scan.nbuf", 4), lbw = b200_tune("scan.lbw", 1);
  const int block = b200_tune("scan.block", 128), bps = b200_tune("scan.bps", 0);
  const int pfd = b200_tune("scan.pfd", -1);  // L2 prefetch distance in tiles (0 = off, -1 = one wave of CTAs)
  // the warp-specialised kernel needs both Views 16-byte aligned (bulk copies); otherwise the uniform kernel
  const bool aligned = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0);
  const int ws = aligned ? b200_tune("scan.ws", 2) : 0;
  #ifdef B200_SWEEP
  const int sleep_ns = b200_tune("scan.sleep", 0), dbg = b200_tune("scan.dbg", 0);
#else
  const int sleep_ns = 0, dbg = 0;  // the experiment switches (skip look-back / skip scan) exist in sweep builds only
#endif
  if (aligned && b200_tune("scan.chunked", 0)) {  // chunk-synchronous kernel (the multi-GPU scan at world 1)
    using L = kb200::Impl::ChunkScanLaunch<T, 384, 9, 4, 4, INCL>;
    const int grid = L::max_grid(I->device, I->props.sm_count);
    if (grid > 0) return L::run(I, nullptr, grid, L::steps_for(n, grid), x, y, n, seed, seed_dev, nseeds, th, td);
  }
#define CFG(BL, NV, NB, LB) \
  if (!ws && block == BL && nv == NV && nbuf == NB && lbw == LB) return ContigScanLaunch<T, BL, NV, NB, LB, INCL, 0>::run(I, x, y, n, seed, seed_dev, th, td, bps, sleep_ns, dbg, nseeds);
#define ZCFG(BL, NV, NB, LB) \
  if (ws == 4 && block == BL && nv == NV && nbuf == NB && lbw == LB) return ContigScanLaunch<T, BL, NV, NB, LB, INCL, 4>::run(I, x, y, n, seed, seed_dev, th, td, bps, sleep_ns, dbg, nseeds);
#define YCFG(BL, NV, NB, LB) \
  if (ws == 3 && block == BL && nv == NV && nbuf == NB && lbw == LB) return ContigScanLaunch<T, BL, NV, NB, LB, INCL, 3>::run(I, x, y, n, seed, seed_dev, th, td, bps, sleep_ns, dbg, nseeds);
#define XCFG(BL, NV, NB, LB) \
  if (ws == 2 && block == BL && nv == NV && nbuf == NB && lbw == LB) return ContigScanLaunch<T, BL, NV, NB, LB, INCL, 2>::run(I, x, y, n, seed, seed_dev, th, td, bps, sleep_ns, dbg, nseeds, pfd);
#define WCFG(BL, NV, NB, LB) \
  if (ws == 1 && block == BL && nv == NV && nbuf == NB && lbw == LB) return ContigScanLaunch<T, BL, NV, NB, LB, INCL, 1>::run(I, x, y, n, seed, seed_dev, th, td, bps, sleep_ns, dbg, nseeds);
  XCFG(128, 9, 4, 1)   // shipped: 5 variants x ~90 configurations measured, profiles/r01_scan_probe_v*.log
  if (!ws) return ContigScanLaunch<T, 256, 9, 2, 2, INCL, 0>::run(I, x, y, n, seed, seed_dev, th, td, bps, sleep_ns, dbg, nseeds);
#ifdef B200_SWEEP
  WCFG(256, 9, 2, 2)
  if constexpr (sizeof(T) == 8 && !INCL) {
    CFG(256, 9, 2, 1) CFG(256, 9, 2, 4) CFG(256, 9, 3, 2) CFG(256, 9, 3, 4)
    CFG(256, 7, 2, 2) CFG(256, 7, 3, 2) CFG(256, 7, 3, 4) CFG(256, 11, 2, 2) CFG(256, 11, 2, 4)
    CFG(512, 9, 2, 2) CFG(512, 9, 2, 4) CFG(512, 7, 2, 2) CFG(512, 7, 3, 4) CFG(512, 5, 3, 4) CFG(512, 5, 2, 2)
    CFG(128, 9, 2, 2) CFG(128, 9, 3, 2) CFG(128, 11, 3, 2) CFG(128, 13, 2, 2) CFG(128, 13, 3, 4)
    CFG(256, 13, 2, 4) CFG(256, 13, 2, 2) CFG(1024, 5, 2, 4) CFG(1024, 7, 2, 4) CFG(1024, 3, 3, 4)
    WCFG(256, 9, 2, 1) WCFG(256, 9, 2, 4) WCFG(256, 9, 2, 8) WCFG(256, 9, 3, 2) WCFG(256, 9, 3, 4) WCFG(256, 9, 3, 8)
    WCFG(256, 7, 2, 4) WCFG(256, 7, 3, 4) WCFG(256, 7, 3, 8) WCFG(256, 7, 4, 4) WCFG(256, 5, 3, 4) WCFG(256, 5, 4, 8)
    WCFG(256, 11, 2, 4) WCFG(256, 11, 2, 8) WCFG(256, 13, 2, 4) WCFG(256, 13, 2, 8)
    WCFG(512, 9, 2, 4) WCFG(512, 9, 2, 8) WCFG(512, 9, 3, 8) WCFG(512, 7, 2, 4) WCFG(512, 7, 3, 8) WCFG(512, 5, 3, 4) WCFG(512, 5, 4, 8)
    WCFG(128, 9, 2, 4) WCFG(128, 9, 3, 4) WCFG(128, 13, 2, 4) WCFG(128, 13, 3, 8) WCFG(128, 7, 3, 4)
    WCFG(1024, 5, 2, 8) WCFG(1024, 3, 3, 8) WCFG(1024, 7, 2, 8)
    ZCFG(256, 9, 3, 1) ZCFG(256, 9, 4, 1) ZCFG(256, 7, 4, 1) ZCFG(256, 5, 5, 1) ZCFG(256, 5, 4, 1) ZCFG(256, 9, 2, 1) ZCFG(256, 13, 3, 1) ZCFG(256, 11, 3, 1) ZCFG(256, 13, 2, 1)
    ZCFG(128, 9, 4, 1) ZCFG(128, 9, 6, 1) ZCFG(128, 13, 4, 1) ZCFG(128, 17, 3, 1) ZCFG(128, 17, 4, 1) ZCFG(128, 25, 2, 1) ZCFG(128, 25, 3, 1)
    ZCFG(512, 9, 3, 1) ZCFG(512, 9, 2, 1) ZCFG(512, 7, 3, 1) ZCFG(512, 5, 4, 1) ZCFG(512, 7, 4, 1)
    ZCFG(256, 9, 3, 2) ZCFG(256, 7, 4, 2) ZCFG(128, 9, 4, 2) ZCFG(256, 9, 3, 4) ZCFG(512, 9, 3, 2) ZCFG(512, 9, 3, 4)
    YCFG(256, 9, 3, 1) YCFG(256, 9, 2, 1) YCFG(256, 9, 4, 1) YCFG(256, 7, 4, 1) YCFG(256, 7, 3, 1) YCFG(256, 5, 5, 1) YCFG(256, 5, 4, 1) YCFG(256, 5, 6, 1) YCFG(256, 11, 3, 1) YCFG(256, 13, 2, 1) YCFG(256, 13, 3, 1)
    YCFG(128, 9, 4, 1) YCFG(128, 9, 3, 1) YCFG(128, 9, 6, 1) YCFG(128, 13, 4, 1) YCFG(128, 13, 3, 1) YCFG(128, 7, 4, 1) YCFG(128, 7, 6, 1) YCFG(128, 17, 3, 1) YCFG(128, 17, 2, 1)
    YCFG(512, 9, 3, 1) YCFG(512, 9, 2, 1) YCFG(512, 7, 3, 1) YCFG(512, 5, 4, 1) YCFG(512, 5, 3, 1) YCFG(64, 17, 4, 1) YCFG(64, 17, 6, 1) YCFG(64, 25, 4, 1)
    YCFG(128, 9, 4, 2) YCFG(256, 7, 4, 2) YCFG(256, 9, 3, 2)
    XCFG(256, 9, 2, 1) XCFG(256, 9, 2, 2) XCFG(256, 9, 2, 4) XCFG(256, 9, 2, 8) XCFG(256, 9, 3, 1) XCFG(256, 9, 3, 2) XCFG(256, 9, 3, 8)
    XCFG(256, 7, 3, 4) XCFG(256, 7, 4, 4) XCFG(256, 5, 4, 4) XCFG(256, 11, 2, 4) XCFG(256, 13, 2, 4) XCFG(256, 11, 3, 4)
    XCFG(512, 9, 2, 4) XCFG(512, 9, 3, 4) XCFG(512, 7, 3, 4) XCFG(512, 5, 3, 4) XCFG(512, 5, 4, 4) XCFG(512, 3, 4, 4)
    XCFG(128, 9, 3, 4) XCFG(128, 13, 3, 4) XCFG(128, 9, 4, 4) XCFG(128, 13, 2, 4) XCFG(256, 9, 3, 4)
    XCFG(256, 5, 5, 1) XCFG(256, 5, 6, 1) XCFG(256, 3, 8, 1) XCFG(256, 7, 4, 1) XCFG(256, 7, 5, 1) XCFG(256, 5, 4, 1)
    XCFG(128, 9, 5, 1) XCFG(128, 9, 6, 1) XCFG(128, 7, 6, 1) XCFG(128, 7, 8, 1) XCFG(128, 5, 8, 1) XCFG(128, 13, 4, 1) XCFG(128, 11, 4, 1)
    XCFG(256, 7, 4, 2) XCFG(256, 5, 5, 2) XCFG(128, 9, 5, 2) XCFG(256, 9, 4, 1) XCFG(512, 5, 4, 1) XCFG(512, 3, 6, 1) XCFG(512, 7, 3, 1) XCFG(512, 9, 3, 1)
  }
#endif
#undef CFG
#undef WCFG
#undef XCFG
#undef YCFG
#undef ZCFG
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
int b200_scan_excl_i64_seeds_dev(b200_instance* I, const int64_t* x, int64_t* y, int64_t n, const int64_t* seeds_dev, int nseeds, int64_t* td) {
  if (nseeds < 0 || (nseeds > 0 && !seeds_dev)) return b200_set_error(B200_EINVAL, "b200_scan_excl_i64_seeds_dev", "bad seed array");
  if (nseeds == 0) return scan_entry<int64, false>(I, "b200_scan_excl_i64_seeds_dev", (const int64*)x, (int64*)y, n, 0, nullptr, nullptr, (int64*)td);
  return scan_entry<int64, false>(I, "b200_scan_excl_i64_seeds_dev", (const int64*)x, (int64*)y, n, 0, (const int64*)seeds_dev, nullptr, (int64*)td, nseeds);
}
}  // extern "C"

#ifdef B200_SWEEP
extern "C" int b200_debug_chunk_stats(unsigned long long* out32) {
  return (int)cudaMemcpy(out32, kb200::Impl::chunk_stats_buffer(), 32 * 8, cudaMemcpyDeviceToHost);
}
// diagnostics for tools/scan_probe.py (sweep build only; not declared in the public header)
extern "C" int b200_debug_scan_stats(unsigned long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, kb200::Impl::g_scan_stats, 16 * sizeof(unsigned long long));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(kb200::Impl::g_scan_stats, z, sizeof z); }
  return 0;
}
#endif
