// reduce.cu -- typed parallel_reduce fast paths (C ABI) = the generic skeleton
// (kb200/impl/ReduceKernel.hpp) instantiated with the built-in reducers (kb200/Reducers.hpp) and
// wide-load bodies over contiguous View<T*> data (kb200/impl/ContigBody.hpp).
//
// Each entry point is what `Kokkos::parallel_reduce(RangePolicy<B200>(0,n), f, Reducer(result))`
// dispatches to when f is the canonical functor of that reducer over one contiguous View:
//   Sum:  u += x(i)            Min: if (x(i) < u) u = x(i)        Max: if (x(i) > u) u = x(i)
//   MinLoc: if (x(i) < u.val) { u.val = x(i); u.loc = i; }  (MaxLoc, MinMax, MinMaxLoc alike)
// -- the functors of core/unit_test/TestReducers.hpp:66-135.
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/Reducers.hpp>
#include <kb200/impl/ReduceKernel.hpp>
#include <kb200/impl/ContigBody.hpp>

using namespace kb200;
using kb200::Impl::ContigBody;
using kb200::Impl::RangeReduceLaunch;

namespace {
struct SumOp { template <class V, class T> KB200_DEVICE_FUNCTION static void apply(V& a, T x, int64) { a += x; } };
struct MinOp { template <class V, class T> KB200_DEVICE_FUNCTION static void apply(V& a, T x, int64) { if (x < a) a = x; } };
struct MaxOp { template <class V, class T> KB200_DEVICE_FUNCTION static void apply(V& a, T x, int64) { if (x > a) a = x; } };
struct MinMaxOp {
  template <class V, class T> KB200_DEVICE_FUNCTION static void apply(V& a, T x, int64) {
    if (x < a.min_val) a.min_val = x;
    if (x > a.max_val) a.max_val = x;
  }
};
// loc reducers: equal extrema keep the LOWEST location whatever order a thread meets them in (the scalar head elements of
// an unaligned View are folded after the vector body, and they carry the lowest indices)
struct MinLocOp { template <class V, class T> KB200_DEVICE_FUNCTION static void apply(V& a, T x, int64 i) { if (x < a.val || (x == a.val && i < a.loc)) { a.val = x; a.loc = i; } } };
struct MaxLocOp { template <class V, class T> KB200_DEVICE_FUNCTION static void apply(V& a, T x, int64 i) { if (x > a.val || (x == a.val && i < a.loc)) { a.val = x; a.loc = i; } } };
struct MinMaxLocOp {
  template <class V, class T> KB200_DEVICE_FUNCTION static void apply(V& a, T x, int64 i) {
    if (x < a.min_val || (x == a.min_val && i < a.min_loc)) { a.min_val = x; a.min_loc = i; }
    if (x > a.max_val || (x == a.max_val && i < a.max_loc)) { a.max_val = x; a.max_loc = i; }
  }
};

template <class T, class Op, class Red, int VB, int BLOCK, int UNROLL>
int run_cfg(b200_instance* I, const T* x, int64_t n, int64_t base, typename Red::value_type* rh, typename Red::value_type* rd, int bps) {
  using V = typename Red::value_type;
  using Body = ContigBody<T, VB, Op, V>;
  Body body(x, (int64)n, (int64)base);
  V dummy;
  Red red(dummy);
  return RangeReduceLaunch<Body, Red, BLOCK, UNROLL, 1024 / BLOCK>::run(I, body, red, body.nvec, rh, rd, bps);  // <= 64 registers: see Parallel.hpp
}

// shipped configuration (see DESIGN.md "reduce kernel: tuning"); the f64 Sum path additionally
// exposes the sweep space through b200_tune_set for tools/sweep.py.
template <class T, class Op, class Red>
int run_default(b200_instance* I, const char* where, const T* x, int64_t n, int64_t base, typename Red::value_type* rh,
                typename Red::value_type* rd) {
  B200_CHECK_INST(I, where);
  if (n < 0) return b200_set_error(B200_EINVAL, where, "negative length");
  if (n > 0 && !x) return b200_set_error(B200_EINVAL, where, "x is NULL");
  if (!rh && !rd) return b200_set_error(B200_EINVAL, where, "no result destination");
  return run_cfg<T, Op, Red, 32, 256, 4>(I, x, n, base, rh, rd, 0);
}
}  // namespace

extern "C" {

int b200_reduce_sum_f64(b200_instance* I, const double* x, int64_t n, double* rh, double* rd) {
  const char* where = "b200_reduce_sum_f64";
  B200_CHECK_INST(I, where);
  if (n < 0) return b200_set_error(B200_EINVAL, where, "negative length");
  if (n > 0 && !x) return b200_set_error(B200_EINVAL, where, "x is NULL");
  if (!rh && !rd) return b200_set_error(B200_EINVAL, where, "no result destination");
  const int vb = b200_tune("reduce.vbytes", 32), un = b200_tune("reduce.unroll", 4), bl = b200_tune("reduce.block", 256);
  const int bps = b200_tune("reduce.bps", 0);
  using R = Sum<double>;
#define CFG(VB, BL, UN) if (vb == VB && bl == BL && un == UN) return run_cfg<double, SumOp, R, VB, BL, UN>(I, x, n, 0, rh, rd, bps);
  CFG(32, 256, 4)
#ifdef B200_SWEEP
  CFG(32, 256, 1) CFG(32, 256, 2) CFG(32, 256, 8)
  CFG(32, 512, 1) CFG(32, 512, 2) CFG(32, 512, 4) CFG(32, 512, 8)
  CFG(32, 128, 2) CFG(32, 128, 4) CFG(32, 128, 8)
  CFG(16, 256, 2) CFG(16, 256, 4) CFG(16, 256, 8)
  CFG(16, 512, 2) CFG(16, 512, 4) CFG(16, 512, 8)
  CFG(8, 256, 2) CFG(8, 256, 4) CFG(8, 256, 8) CFG(8, 256, 16)
  CFG(8, 512, 4) CFG(8, 512, 8)
#endif
#undef CFG
  return b200_set_error(B200_EUNSUPPORTED, where, "tuning combination not compiled in");
}
int b200_reduce_sum_f32(b200_instance* I, const float* x, int64_t n, float* rh, float* rd) { return run_default<float, SumOp, Sum<float>>(I, "b200_reduce_sum_f32", x, n, 0, rh, rd); }
int b200_reduce_sum_i64(b200_instance* I, const int64_t* x, int64_t n, int64_t* rh, int64_t* rd) {
  return run_default<int64, SumOp, Sum<int64>>(I, "b200_reduce_sum_i64", (const int64*)x, n, 0, (int64*)rh, (int64*)rd);
}
int b200_reduce_sum_i32(b200_instance* I, const int32_t* x, int64_t n, int32_t* rh, int32_t* rd) { return run_default<int, SumOp, Sum<int>>(I, "b200_reduce_sum_i32", x, n, 0, rh, rd); }
int b200_reduce_min_f64(b200_instance* I, const double* x, int64_t n, double* rh, double* rd) { return run_default<double, MinOp, Min<double>>(I, "b200_reduce_min_f64", x, n, 0, rh, rd); }
int b200_reduce_max_f64(b200_instance* I, const double* x, int64_t n, double* rh, double* rd) { return run_default<double, MaxOp, Max<double>>(I, "b200_reduce_max_f64", x, n, 0, rh, rd); }
int b200_reduce_min_i64(b200_instance* I, const int64_t* x, int64_t n, int64_t* rh, int64_t* rd) {
  return run_default<int64, MinOp, Min<int64>>(I, "b200_reduce_min_i64", (const int64*)x, n, 0, (int64*)rh, (int64*)rd);
}
int b200_reduce_max_i64(b200_instance* I, const int64_t* x, int64_t n, int64_t* rh, int64_t* rd) {
  return run_default<int64, MaxOp, Max<int64>>(I, "b200_reduce_max_i64", (const int64*)x, n, 0, (int64*)rh, (int64*)rd);
}
int b200_reduce_min_i32(b200_instance* I, const int32_t* x, int64_t n, int32_t* rh, int32_t* rd) { return run_default<int, MinOp, Min<int>>(I, "b200_reduce_min_i32", x, n, 0, rh, rd); }
int b200_reduce_max_i32(b200_instance* I, const int32_t* x, int64_t n, int32_t* rh, int32_t* rd) { return run_default<int, MaxOp, Max<int>>(I, "b200_reduce_max_i32", x, n, 0, rh, rd); }

// the C structs are layout-identical to the kb200 value structs
static_assert(sizeof(b200_valloc_f64) == sizeof(ValLocScalar<double, int64>), "");
static_assert(sizeof(b200_minmaxloc_f64) == sizeof(MinMaxLocScalar<double, int64>), "");
static_assert(sizeof(b200_minmax_f64) == sizeof(MinMaxScalar<double>), "");

int b200_reduce_minmax_f64(b200_instance* I, const double* x, int64_t n, b200_minmax_f64* rh, b200_minmax_f64* rd) {
  using R = MinMax<double>;
  return run_default<double, MinMaxOp, R>(I, "b200_reduce_minmax_f64", x, n, 0, (R::value_type*)rh, (R::value_type*)rd);
}
int b200_reduce_minloc_f64(b200_instance* I, const double* x, int64_t n, int64_t base, b200_valloc_f64* rh, b200_valloc_f64* rd) {
  using R = MinLoc<double, int64>;
  return run_default<double, MinLocOp, R>(I, "b200_reduce_minloc_f64", x, n, base, (R::value_type*)rh, (R::value_type*)rd);
}
int b200_reduce_maxloc_f64(b200_instance* I, const double* x, int64_t n, int64_t base, b200_valloc_f64* rh, b200_valloc_f64* rd) {
  using R = MaxLoc<double, int64>;
  return run_default<double, MaxLocOp, R>(I, "b200_reduce_maxloc_f64", x, n, base, (R::value_type*)rh, (R::value_type*)rd);
}
int b200_reduce_minmaxloc_f64(b200_instance* I, const double* x, int64_t n, int64_t base, b200_minmaxloc_f64* rh, b200_minmaxloc_f64* rd) {
  using R = MinMaxLoc<double, int64>;
  return run_default<double, MinMaxLocOp, R>(I, "b200_reduce_minmaxloc_f64", x, n, base, (R::value_type*)rh, (R::value_type*)rd);
}

}  // extern "C"
