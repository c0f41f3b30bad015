// cases_algorithms.cu -- std-algorithm layer + ViewFill + Crs row-map cases for the B200 execution space, written as the
// reference's users / tests write them (algorithms/unit_tests/TestStdAlgorithms{Exclusive,Inclusive}Scan.cpp,
// TestStdAlgorithmsReduce.cpp, TestStdAlgorithmsTransformReduce.cpp, TestStdAlgorithmsMinMaxElementOps.cpp,
// core/unit_test/TestCrs.hpp:176-198), exported with a C ABI for tests/test_gpu_algorithms.py which checks them against numpy.
#include <Kokkos_B200.hpp>
#include <cstdint>
#include <string>

using namespace kb200;
namespace KE = kb200::Experimental;
using i64 = long long;

namespace {
std::string g_alg_err;
template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_alg_err = e.what();
    return -100;
  }
}
template <class T>
using HostU = View<T*, HostSpace, MemoryTraits<Unmanaged>>;
template <class T>
View<T*> to_device(const T* h, i64 n) {
  View<T*> d(view_alloc(WithoutInitializing, "in"), (size_t)n);
  deep_copy(d, HostU<const T>(h, (size_t)n));
  return d;
}
template <class T>
void to_host(T* h, const View<T*>& d) { deep_copy(HostU<T>(h, d.extent(0)), d); }

struct MaxOp { template <class T> KB200_INLINE_FUNCTION T operator()(const T& a, const T& b) const { return a < b ? b : a; } };
struct MulMod { KB200_INLINE_FUNCTION unsigned operator()(unsigned a, unsigned b) const { return (unsigned)(((unsigned long long)a * b) % 1000003u); } };
struct Square { template <class T> KB200_INLINE_FUNCTION T operator()(const T& a) const { return a * a; } };
struct IsNegative { template <class T> KB200_INLINE_FUNCTION bool operator()(const T& a) const { return a < T(0); } };

template <class T>
int scan_case(int kind, const T* hx, T* hy, i64 n, T init) {
  B200 space;
  View<T*> x = to_device(hx, n), y("y", (size_t)n);
  size_t wrote = 0;
  switch (kind) {
    case 0: wrote = KE::exclusive_scan(space, x, y, init); break;
    case 1: wrote = KE::inclusive_scan(space, x, y); break;
    case 2: wrote = KE::exclusive_scan(space, x, y, init, MaxOp{}); break;
    case 3: wrote = KE::inclusive_scan(space, x, y, MaxOp{}); break;
    case 4: wrote = KE::exclusive_scan(space, x, x, init); space.fence(); to_host(hy, x); return (int)(wrote != (size_t)n);  // in place
    default: return -1;
  }
  space.fence();
  to_host(hy, y);
  return wrote == (size_t)n ? 0 : 1;
}
}  // namespace

extern "C" {
const char* kb200_alg_last_error() { return g_alg_err.c_str(); }

int kb200_alg_scan_i64(int kind, const i64* hx, i64* hy, i64 n, i64 init) { return guarded([&] { return scan_case<i64>(kind, hx, hy, n, init); }); }
int kb200_alg_scan_f64(int kind, const double* hx, double* hy, i64 n, double init) { return guarded([&] { return scan_case<double>(kind, hx, hy, n, init); }); }
int kb200_alg_scan_i32(int kind, const int* hx, int* hy, i64 n, int init) { return guarded([&] { return scan_case<int>(kind, hx, hy, n, init); }); }
int kb200_alg_scan_f32(int kind, const float* hx, float* hy, i64 n, float init) { return guarded([&] { return scan_case<float>(kind, hx, hy, n, init); }); }
int kb200_alg_scan_u32_mulmod(const unsigned* hx, unsigned* hy, i64 n, unsigned init, int inclusive) {
  return guarded([&] {
    B200 space;
    View<unsigned*> x = to_device(hx, n), y("y", (size_t)n);
    if (inclusive) KE::inclusive_scan(space, x, y, MulMod{});
    else KE::exclusive_scan(space, x, y, init, MulMod{});
    space.fence();
    to_host(hy, y);
    return 0;
  });
}

// out: [0] reduce(double) [1] reduce(double, init 2.5) [2] reduce with MaxOp [3] dot(x,x) [4] transform_reduce(max of squares)
//      [5] reduce(float view) [6] count_if(negative) ; iout: [0] reduce(i64 view) [1] min_element [2] max_element [3],[4] minmax_element
//      [5] find_if(negative) [6] reduce(int view) [7] find_if on a view without negatives (must be n)
int kb200_alg_reductions(const double* hx, i64 n, double* out, i64* iout) {
  return guarded([&] {
    B200 space;
    View<double*> x = to_device(hx, n);
    View<float*> xf("xf", (size_t)n);
    View<i64*> xi("xi", (size_t)n);
    View<int*> x32("x32", (size_t)n);
    View<double*> xabs("xabs", (size_t)n);
    parallel_for(n, KB200_LAMBDA(const i64 i) {
      xf(i) = (float)(int)(x(i) * 8.0); xi(i) = (i64)(x(i) * 1000.0); x32(i) = (int)(x(i) * 100.0); xabs(i) = x(i) < 0 ? -x(i) : x(i);
    });
    out[0] = KE::reduce(space, x);
    out[1] = KE::reduce(space, x, 2.5);
    out[2] = KE::reduce(space, x, -1e300, MaxOp{});
    out[3] = KE::transform_reduce(space, x, x, 0.0);
    out[4] = KE::transform_reduce(space, x, 0.0, MaxOp{}, Square{});
    out[5] = (double)KE::reduce(space, xf);
    out[6] = (double)KE::count_if(space, x, IsNegative{});
    iout[0] = KE::reduce(space, xi);
    iout[1] = (i64)KE::min_element(space, x);
    iout[2] = (i64)KE::max_element(space, x);
    auto mm = KE::minmax_element(space, x);
    iout[3] = (i64)mm.first; iout[4] = (i64)mm.second;
    iout[5] = (i64)KE::find_if(space, x, IsNegative{});
    iout[6] = (i64)KE::reduce(space, x32);
    iout[7] = (i64)KE::find_if(space, xabs, IsNegative{});
    return 0;
  });
}

// fill / copy / transform / deep_copy(view, value) on several types; results copied back for checking
int kb200_alg_elementwise(i64 n, double* hd, int* hi, double* hcopy, double* hsq, unsigned char* hb) {
  return guarded([&] {
    B200 space;
    View<double*> d("d", (size_t)n), c("c", (size_t)n), sq("sq", (size_t)n);
    View<int*> iv("iv", (size_t)n);
    View<unsigned char*> bv("bv", (size_t)n);
    KE::fill(space, d, 3.25);
    deep_copy(iv, 7);           // non-zero fill, generic path
    deep_copy(bv, (unsigned char)0);  // zero pattern: memset path
    deep_copy(bv, (unsigned char)201);
    KE::for_each(space, d, KB200_LAMBDA(double& v) { v += 1.0; });
    KE::copy(space, d, c);
    KE::transform(space, c, sq, Square{});
    space.fence();
    to_host(hd, d); to_host(hi, iv); to_host(hcopy, c); to_host(hsq, sq); to_host(hb, bv);
    View<int*> z("z", (size_t)n);
    deep_copy(z, 5); deep_copy(z, 0);
    i64 s = -1;
    parallel_reduce(n, KB200_LAMBDA(const i64 i, i64& u) { u += z(i); }, s);
    return s == 0 ? 0 : 2;
  });
}

// Crs: row_map from counts (TestCrs.hpp), three index types; returns totals
int kb200_alg_crs_row_map(const i64* hcounts, i64 n, i64* hrm64, int* hrm32, unsigned* hrmu, i64* totals) {
  return guarded([&] {
    B200 space;
    View<i64*> c64 = to_device(hcounts, n), rm64("rm64", (size_t)n + 1);
    View<int*> c32("c32", (size_t)n), rm32("rm32", (size_t)n + 1);
    View<unsigned*> cu("cu", (size_t)n), rmu("rmu", (size_t)n + 1);
    parallel_for(n, KB200_LAMBDA(const i64 i) { c32(i) = (int)c64(i); cu(i) = (unsigned)c64(i); });
    totals[0] = get_crs_row_map_from_counts(space, rm64, c64);
    totals[1] = get_crs_row_map_from_counts(space, rm32, c32);
    totals[2] = (i64)get_crs_row_map_from_counts(space, rmu, cu);
    space.fence();
    to_host(hrm64, rm64); to_host(hrm32, rm32); to_host(hrmu, rmu);
    return 0;
  });
}
}  // extern "C"
