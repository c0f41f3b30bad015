// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.  A C ABI over the UNMODIFIED reference
// (kokkos 4.6.99, Kokkos::OpenMP backend) compiled in place from /root/reference by oracle/Makefile.
// Every function below is a plain call of the reference's public API -- Kokkos::parallel_for /
// parallel_reduce / parallel_scan over RangePolicy / MDRangePolicy / TeamPolicy on Kokkos::OpenMP with
// unmanaged HostSpace Views over the caller's buffers -- using the functors of the reference's own tests
// and benchmarks (TestReducers.hpp:66-135, TestParallelScanRangePolicy.hpp:41-84,
// benchmarks/stream/stream-kokkos.cpp:55-77, benchmarks/gups/gups.cpp:83-97,
// example/tutorial/Hierarchical_Parallelism/03_vectorization/vectorization.cpp:51-76).
// It is the parity oracle and the "reference" CPU baseline; nothing in the product links to it.
#include <Kokkos_Core.hpp>
#include <cstdint>
#include <cstdio>

using Exec = Kokkos::OpenMP;
using Host = Kokkos::HostSpace;
template <class T>
using UView = Kokkos::View<T*, Host, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
template <class T>
using CUView = Kokkos::View<const T*, Host, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
using Range = Kokkos::RangePolicy<Exec, Kokkos::IndexType<int64_t>>;

extern "C" {

struct ref_valloc_f64 { double val; int64_t loc; };
struct ref_minmaxloc_f64 { double min_val, max_val; int64_t min_loc, max_loc; };
struct ref_minmax_f64 { double min_val, max_val; };

int ref_init(int threads) {
  if (Kokkos::is_initialized()) return Exec().concurrency();
  Kokkos::InitializationSettings s;
  if (threads > 0) s.set_num_threads(threads);
  s.set_disable_warnings(true);
  Kokkos::initialize(s);
  return Exec().concurrency();
}
void ref_finalize() {
  if (Kokkos::is_initialized() && !Kokkos::is_finalized()) Kokkos::finalize();
}
int ref_concurrency() { return Exec().concurrency(); }
const char* ref_version() { return "kokkos 4.6.99 (reference, unmodified) Kokkos::OpenMP"; }

// ---------------------------------------------------------------- reductions
#define REF_SUM(NAME, T)                                                                         \
  T NAME(const T* x, int64_t n) {                                                                \
    CUView<T> a(x, n);                                                                           \
    T r = 0;                                                                                     \
    Kokkos::parallel_reduce("ref_sum", Range(0, n), KOKKOS_LAMBDA(const int64_t i, T& u) { u += a(i); }, r); \
    return r;                                                                                    \
  }
REF_SUM(ref_reduce_sum_f64, double)
REF_SUM(ref_reduce_sum_f32, float)
REF_SUM(ref_reduce_sum_i64, int64_t)
REF_SUM(ref_reduce_sum_i32, int32_t)

#define REF_MINMAX1(NAME, T, RED, CMP)                                                           \
  T NAME(const T* x, int64_t n) {                                                                \
    CUView<T> a(x, n);                                                                           \
    T r;                                                                                         \
    Kokkos::parallel_reduce("ref_minmax", Range(0, n),                                           \
        KOKKOS_LAMBDA(const int64_t i, T& u) { if (a(i) CMP u) u = a(i); }, Kokkos::RED<T>(r));  \
    return r;                                                                                    \
  }
REF_MINMAX1(ref_reduce_min_f64, double, Min, <)
REF_MINMAX1(ref_reduce_max_f64, double, Max, >)
REF_MINMAX1(ref_reduce_min_i64, int64_t, Min, <)
REF_MINMAX1(ref_reduce_max_i64, int64_t, Max, >)
REF_MINMAX1(ref_reduce_min_i32, int32_t, Min, <)
REF_MINMAX1(ref_reduce_max_i32, int32_t, Max, >)

ref_minmax_f64 ref_reduce_minmax_f64(const double* x, int64_t n) {
  CUView<double> a(x, n);
  using R = Kokkos::MinMax<double>;
  R::value_type r;
  Kokkos::parallel_reduce("ref_minmax", Range(0, n), KOKKOS_LAMBDA(const int64_t i, R::value_type& u) {
    if (a(i) < u.min_val) u.min_val = a(i);
    if (a(i) > u.max_val) u.max_val = a(i);
  }, R(r));
  return ref_minmax_f64{r.min_val, r.max_val};
}
ref_valloc_f64 ref_reduce_minloc_f64(const double* x, int64_t n, int64_t base) {
  CUView<double> a(x, n);
  using R = Kokkos::MinLoc<double, int64_t>;
  R::value_type r;
  Kokkos::parallel_reduce("ref_minloc", Range(0, n), KOKKOS_LAMBDA(const int64_t i, R::value_type& u) {
    if (a(i) < u.val) { u.val = a(i); u.loc = base + i; }
  }, R(r));
  return ref_valloc_f64{r.val, r.loc};
}
ref_valloc_f64 ref_reduce_maxloc_f64(const double* x, int64_t n, int64_t base) {
  CUView<double> a(x, n);
  using R = Kokkos::MaxLoc<double, int64_t>;
  R::value_type r;
  Kokkos::parallel_reduce("ref_maxloc", Range(0, n), KOKKOS_LAMBDA(const int64_t i, R::value_type& u) {
    if (a(i) > u.val) { u.val = a(i); u.loc = base + i; }
  }, R(r));
  return ref_valloc_f64{r.val, r.loc};
}
ref_minmaxloc_f64 ref_reduce_minmaxloc_f64(const double* x, int64_t n, int64_t base) {
  CUView<double> a(x, n);
  using R = Kokkos::MinMaxLoc<double, int64_t>;
  R::value_type r;
  Kokkos::parallel_reduce("ref_minmaxloc", Range(0, n), KOKKOS_LAMBDA(const int64_t i, R::value_type& u) {
    if (a(i) < u.min_val) { u.min_val = a(i); u.min_loc = base + i; }
    if (a(i) > u.max_val) { u.max_val = a(i); u.max_loc = base + i; }
  }, R(r));
  return ref_minmaxloc_f64{r.min_val, r.max_val, r.min_loc, r.max_loc};
}

// ---------------------------------------------------------------- scans
#define REF_SCAN(NAME, T)                                                                        \
  T NAME(const T* x, T* y, int64_t n, T seed, int inclusive) {                                   \
    CUView<T> a(x, n);                                                                           \
    UView<T> b(y, n);                                                                            \
    T total = 0;                                                                                 \
    if (inclusive)                                                                               \
      Kokkos::parallel_scan("ref_scan", Range(0, n), KOKKOS_LAMBDA(const int64_t i, T& u, const bool fin) { \
        u += a(i); if (fin) b(i) = seed + u; }, total);                                          \
    else                                                                                         \
      Kokkos::parallel_scan("ref_scan", Range(0, n), KOKKOS_LAMBDA(const int64_t i, T& u, const bool fin) { \
        const T xi = a(i); if (fin) b(i) = seed + u; u += xi; }, total);                         \
    return total;                                                                                \
  }
REF_SCAN(ref_scan_i64, int64_t)
REF_SCAN(ref_scan_i32, int32_t)
REF_SCAN(ref_scan_f64, double)

// ---------------------------------------------------------------- stream
void ref_stream_set_f64(double* a_, double v, int64_t n) {
  UView<double> a(a_, n);
  Kokkos::parallel_for("set", Range(0, n), KOKKOS_LAMBDA(const int64_t i) { a(i) = v; });
}
void ref_stream_copy_f64(const double* a_, double* b_, int64_t n) {
  CUView<double> a(a_, n); UView<double> b(b_, n);
  Kokkos::parallel_for("copy", Range(0, n), KOKKOS_LAMBDA(const int64_t i) { b(i) = a(i); });
}
void ref_stream_scale_f64(double* b_, const double* c_, double s, int64_t n) {
  UView<double> b(b_, n); CUView<double> c(c_, n);
  Kokkos::parallel_for("scale", Range(0, n), KOKKOS_LAMBDA(const int64_t i) { b(i) = s * c(i); });
}
void ref_stream_add_f64(const double* a_, const double* b_, double* c_, int64_t n) {
  CUView<double> a(a_, n), b(b_, n); UView<double> c(c_, n);
  Kokkos::parallel_for("add", Range(0, n), KOKKOS_LAMBDA(const int64_t i) { c(i) = a(i) + b(i); });
}
void ref_stream_triad_f64(double* a_, const double* b_, const double* c_, double s, int64_t n) {
  UView<double> a(a_, n); CUView<double> b(b_, n), c(c_, n);
  Kokkos::parallel_for("triad", Range(0, n), KOKKOS_LAMBDA(const int64_t i) { a(i) = b(i) + s * c(i); });
}

// ---------------------------------------------------------------- MDRange stencil + MinMaxLoc
ref_minmaxloc_f64 ref_stencil7_minmaxloc_f64(const double* u_, double* v_, int64_t n0, int64_t n1, int64_t n2, double c0, double c1) {
  using V3 = Kokkos::View<const double***, Kokkos::LayoutLeft, Host, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  using W3 = Kokkos::View<double***, Kokkos::LayoutLeft, Host, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  V3 u(u_, n0, n1, n2);
  W3 v(v_, v_ ? n0 : 0, v_ ? n1 : 0, v_ ? n2 : 0);
  const bool store = v_ != nullptr;
  using R = Kokkos::MinMaxLoc<double, int64_t>;
  R::value_type r;
  using MD = Kokkos::MDRangePolicy<Exec, Kokkos::Rank<3>, Kokkos::IndexType<int64_t>>;
  if (n0 < 3 || n1 < 3 || n2 < 3) { R(r).init(r); return ref_minmaxloc_f64{r.min_val, r.max_val, r.min_loc, r.max_loc}; }
  Kokkos::parallel_reduce("stencil7", MD({1, 1, 1}, {n0 - 1, n1 - 1, n2 - 1}),
      KOKKOS_LAMBDA(const int64_t i, const int64_t j, const int64_t k, R::value_type& m) {
        double s = u(i - 1, j, k) + u(i + 1, j, k);
        s = s + u(i, j - 1, k);
        s = s + u(i, j + 1, k);
        s = s + u(i, j, k - 1);
        s = s + u(i, j, k + 1);
        const double val = c0 * u(i, j, k) + c1 * s;
        if (store) v(i, j, k) = val;
        const int64_t loc = (i * n1 + j) * n2 + k;
        if (val < m.min_val) { m.min_val = val; m.min_loc = loc; }
        if (val > m.max_val) { m.max_val = val; m.max_loc = loc; }
      }, R(r));
  return ref_minmaxloc_f64{r.min_val, r.max_val, r.min_loc, r.max_loc};
}

// ---------------------------------------------------------------- atomics
void ref_gups_add_i64(int64_t* t_, int64_t len, const int64_t* idx_, int64_t m, int64_t d) {
  UView<int64_t> t(t_, len); CUView<int64_t> idx(idx_, m);
  Kokkos::parallel_for("gups_add", Range(0, m), KOKKOS_LAMBDA(const int64_t i) { Kokkos::atomic_add(&t(idx(i)), d); });
}
void ref_gups_xor_i64(int64_t* t_, int64_t len, const int64_t* idx_, int64_t m, int64_t d) {
  UView<int64_t> t(t_, len); CUView<int64_t> idx(idx_, m);
  Kokkos::parallel_for("gups_xor", Range(0, m), KOKKOS_LAMBDA(const int64_t i) { Kokkos::atomic_fetch_xor(&t(idx(i)), d); });
}

// ---------------------------------------------------------------- TeamPolicy SpMV
void ref_spmv_crs_f64(int64_t nrows, const int64_t* rm_, const int32_t* ci_, const double* va_, int64_t nnz, const double* x_,
                      int64_t ncols, double* y_) {
  CUView<int64_t> row_map(rm_, nrows + 1); CUView<int32_t> col(ci_, nnz); CUView<double> val(va_, nnz), x(x_, ncols);
  UView<double> y(y_, nrows);
  using TP = Kokkos::TeamPolicy<Exec>;
  const int rows_per_team = 64;
  const int league = (int)((nrows + rows_per_team - 1) / rows_per_team);
  Kokkos::parallel_for("spmv", TP(league, Kokkos::AUTO), KOKKOS_LAMBDA(const TP::member_type& team) {
    const int64_t first = (int64_t)team.league_rank() * rows_per_team;
    const int64_t last = first + rows_per_team < nrows ? first + rows_per_team : nrows;
    Kokkos::parallel_for(Kokkos::TeamThreadRange(team, first, last), [&](const int64_t row) {
      double s = 0;
      Kokkos::parallel_reduce(Kokkos::ThreadVectorRange(team, row_map(row), row_map(row + 1)),
                              [&](const int64_t k, double& u) { u += val(k) * x(col(k)); }, s);
      Kokkos::single(Kokkos::PerThread(team), [&]() { y(row) = s; });
    });
  });
}

// ---------------------------------------------------------------- timed legs for bench.py --impl reference / cpu_baseline
// reps timed calls after one warm-up over reference-allocated Views; returns best seconds
double ref_time_reduce_sum_f64(int64_t n, int reps, double* result) {
  Kokkos::View<double*, Host> a("a", n);
  Kokkos::parallel_for("fill", Range(0, n), KOKKOS_LAMBDA(const int64_t i) { a(i) = (double)((i * 2654435761ull >> 7) % 100); });
  double best = 1e30, r = 0;
  for (int k = 0; k <= reps; ++k) {
    Kokkos::Timer t;
    Kokkos::parallel_reduce("sum", Range(0, n), KOKKOS_LAMBDA(const int64_t i, double& u) { u += a(i); }, r);
    const double s = t.seconds();
    if (k > 0 && s < best) best = s;
  }
  if (result) *result = r;
  return best;
}
double ref_time_scan_excl_i64(int64_t n, int reps, int64_t* total_out) {
  Kokkos::View<int64_t*, Host> x("x", n), y("y", n);
  Kokkos::parallel_for("fill", Range(0, n), KOKKOS_LAMBDA(const int64_t i) { x(i) = (int64_t)((i * 2654435761ull >> 7) % 7) - 3; });
  double best = 1e30;
  int64_t total = 0;
  for (int k = 0; k <= reps; ++k) {
    Kokkos::Timer t;
    Kokkos::parallel_scan("scan", Range(0, n), KOKKOS_LAMBDA(const int64_t i, int64_t& u, const bool fin) {
      if (fin) y(i) = u; u += x(i); }, total);
    const double s = t.seconds();
    if (k > 0 && s < best) best = s;
  }
  if (total_out) *total_out = total;
  return best;
}
double ref_time_stream_triad_f64(int64_t n, int reps) {
  Kokkos::View<double*, Host> a("a", n), b("b", n), c("c", n);
  Kokkos::deep_copy(a, 1.0); Kokkos::deep_copy(b, 2.0); Kokkos::deep_copy(c, 0.0);
  const double s3 = 3.0;
  double best = 1e30;
  for (int k = 0; k <= reps; ++k) {
    Kokkos::Timer t;
    Kokkos::parallel_for("triad", Range(0, n), KOKKOS_LAMBDA(const int64_t i) { a(i) = b(i) + s3 * c(i); });
    Kokkos::fence();
    const double s = t.seconds();
    if (k > 0 && s < best) best = s;
  }
  return best;
}

}  // extern "C"
