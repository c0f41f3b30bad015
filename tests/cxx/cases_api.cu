// cases_api.cu -- Kokkos-API-level test cases for the B200 execution space, exported with a C ABI so the
// pytest suite (tests/test_gpu_cxx_api.py) can drive them with host buffers and compare against the oracle.
//
// Every case is written the way a Kokkos user writes it -- kb200::View, RangePolicy / MDRangePolicy /
// TeamPolicy, KB200_LAMBDA functors, reducer objects, nested parallelism, atomics -- using the functors of
// the reference's own unit tests and benchmarks:
//   core/unit_test/TestReducers.hpp:66-135,450-1105   core/unit_test/TestParallelScanRangePolicy.hpp:41-84
//   core/unit_test/TestMDRange.hpp, TestMDRangeReduce.hpp:24-63   core/unit_test/TestTeam.hpp, TestTeamVector.hpp
//   core/unit_test/TestAtomics.hpp:456-598              benchmarks/stream/stream-kokkos.cpp:55-77
//   example/tutorial/Hierarchical_Parallelism/03_vectorization/vectorization.cpp:51-76
// They exercise the GENERIC template path (impl/ReduceKernel, ScanGeneric, ForKernel, MDRangeKernel, Team),
// i.e. what any user functor gets, not the typed fast paths of libkokkos_b200.so.
#include <Kokkos_B200.hpp>
#include <cmath>
#include <cstdint>
#include <string>

using namespace kb200;
using i64 = long long;

namespace {
// products and sums without FMA contraction (the oracle is built with -ffp-contract=off)
KB200_INLINE_FUNCTION double nf_add(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
KB200_INLINE_FUNCTION double nf_mul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
std::string g_err;
template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -100;
  }
}
template <class T>
using HostU = View<T*, HostSpace, MemoryTraits<Unmanaged>>;

template <class T>
View<T*> to_device(const T* h, i64 n, const char* label = "in") {
  View<T*> d(view_alloc(WithoutInitializing, label), (size_t)n);
  deep_copy(d, HostU<const T>(h, (size_t)n));
  return d;
}
template <class T>
void to_host(T* h, const View<T*>& d) { deep_copy(HostU<T>(h, d.extent(0)), d); }

struct TagA {};
struct SumSqFunctor {  // functor-provided init / join / final (TestReduce.hpp style)
  using value_type = double;
  View<const double*> x;
  KB200_INLINE_FUNCTION void operator()(const i64 i, double& u) const { u += x(i) * x(i); }
  KB200_INLINE_FUNCTION void init(double& u) const { u = 0.0; }
  KB200_INLINE_FUNCTION void join(double& d, const double& s) const { d += s; }
  KB200_INLINE_FUNCTION void final(double& u) const { u = u + 1.0; }
};
struct TaggedFunctor {
  View<const double*> x;
  KB200_INLINE_FUNCTION void operator()(const TagA&, const int i, double& u) const { u += x(i); }
};
struct Stats {  // multi-word value type with operator+= (what CombinedReducer produces)
  double s, s2;
  i64 cnt;
  int pad;
  KB200_INLINE_FUNCTION Stats() : s(0), s2(0), cnt(0), pad(0) {}
  KB200_INLINE_FUNCTION Stats& operator+=(const Stats& o) { s += o.s; s2 += o.s2; cnt += o.cnt; return *this; }
};
// affine maps x -> a*x+b over Z/2^64: composition is associative and NOT commutative
struct Affine {
  unsigned long long a, b;
};
struct AffineScan {
  using value_type = Affine;
  View<const unsigned long long*> a, b;
  View<unsigned long long*> ya, yb;
  int inclusive;
  KB200_INLINE_FUNCTION void init(Affine& v) const { v.a = 1; v.b = 0; }
  // dest = (apply dest first, then src)
  KB200_INLINE_FUNCTION void join(Affine& d, const Affine& s) const { d.b = s.a * d.b + s.b; d.a = s.a * d.a; }
  KB200_INLINE_FUNCTION void operator()(const i64 i, Affine& u, const bool fin) const {
    const Affine me{a(i), b(i)};
    if (inclusive) { join(u, me); if (fin) { ya(i) = u.a; yb(i) = u.b; } }
    else { if (fin) { ya(i) = u.a; yb(i) = u.b; } join(u, me); }
  }
};
}  // namespace

extern "C" {

const char* kb200_case_last_error() { return g_err.c_str(); }
int kb200_case_init(int device) {
  return guarded([&] { initialize(InitializationSettings().set_device_id(device)); return 0; });
}
int kb200_case_finalize() { return guarded([] { finalize(); return 0; }); }

// ------------------------------------------------------------------ parallel_reduce over RangePolicy
int kb200_case_reduce_f64(int op, const double* hx, i64 n, i64 base, double* out, i64* out_loc) {
  return guarded([&] {
    View<const double*> x = to_device(hx, n);
    switch (op) {
      case 0: {  // plain scalar result => Sum
        double r = -1;
        parallel_reduce("sum", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, double& u) { u += x(i); }, r);
        out[0] = r; break; }
      case 1: { double r = -1; parallel_reduce(n, KB200_LAMBDA(const i64 i, double& u) { u += x(i); }, Sum<double>(r)); out[0] = r; break; }
      case 2: { double r = 0; parallel_reduce(RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, double& u) { if (x(i) < u) u = x(i); }, Min<double>(r)); out[0] = r; break; }
      case 3: { double r = 0; parallel_reduce(RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, double& u) { if (x(i) > u) u = x(i); }, Max<double>(r)); out[0] = r; break; }
      case 4: { using R = MinLoc<double, i64>; R::value_type r;
        parallel_reduce(RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, R::value_type& u) { if (x(i) < u.val) { u.val = x(i); u.loc = base + i; } }, R(r));
        out[0] = r.val; out_loc[0] = r.loc; break; }
      case 5: { using R = MaxLoc<double, i64>; R::value_type r;
        parallel_reduce(RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, R::value_type& u) { if (x(i) > u.val) { u.val = x(i); u.loc = base + i; } }, R(r));
        out[0] = r.val; out_loc[0] = r.loc; break; }
      case 6: { using R = MinMax<double>; R::value_type r;
        parallel_reduce(RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, R::value_type& u) { if (x(i) < u.min_val) u.min_val = x(i); if (x(i) > u.max_val) u.max_val = x(i); }, R(r));
        out[0] = r.min_val; out[1] = r.max_val; break; }
      case 7: { using R = MinMaxLoc<double, i64>; R::value_type r;
        parallel_reduce(RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, R::value_type& u) {
          if (x(i) < u.min_val) { u.min_val = x(i); u.min_loc = base + i; }
          if (x(i) > u.max_val) { u.max_val = x(i); u.max_loc = base + i; } }, R(r));
        out[0] = r.min_val; out[1] = r.max_val; out_loc[0] = r.min_loc; out_loc[1] = r.max_loc; break; }
      case 8: {  // result in a device View: asynchronous (TestReductions_DeviceView.hpp)
        View<double> r("r");
        parallel_reduce(RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, double& u) { u += x(i); }, r);
        View<double, HostSpace> h("h");
        deep_copy(h, r);
        out[0] = h(); break; }
      case 9: {  // work tag + IndexType<int> + Schedule<Dynamic> + LaunchBounds (TestRange.hpp:331-375)
        double r = -1;
        parallel_reduce(RangePolicy<B200, Schedule<Dynamic>, IndexType<int>, TagA, LaunchBounds<256, 2>>(0, (int)n), TaggedFunctor{x}, r);
        out[0] = r; break; }
      case 10: {  // functor with init/join/final
        double r = -1;
        parallel_reduce(RangePolicy<>(0, n), SumSqFunctor{x}, r);
        out[0] = r; break; }
      case 11: {  // struct value type (3 components, 32 bytes)
        Stats r;
        parallel_reduce(RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, Stats& u) { u.s += x(i); u.s2 += x(i) * x(i); u.cnt += 1; }, r);
        out[0] = r.s; out[1] = r.s2; out_loc[0] = r.cnt; break; }
      case 12: {  // non-zero begin
        double r = -1;
        const i64 b = n / 3;
        parallel_reduce(RangePolicy<>(b, n), KB200_LAMBDA(const i64 i, double& u) { u += x(i); }, r);
        out[0] = r; break; }
      case 13: {  // Prod over a short range
        double r = -1;
        const i64 m = n < 20 ? n : 20;
        parallel_reduce(RangePolicy<>(0, m), KB200_LAMBDA(const i64 i, double& u) { u *= (1.0 + x(i) * 0.0 + (double)(i % 3 + 1)); }, Prod<double>(r));
        out[0] = r; break; }
      default: return -1;
    }
    return 0;
  });
}

int kb200_case_reduce_i32(int op, const int32_t* hx, i64 n, int32_t* out) {
  return guarded([&] {
    View<const int*> x = to_device((const int*)hx, n);
    int r = 0;
    switch (op) {
      case 0: parallel_reduce(n, KB200_LAMBDA(const i64 i, int& u) { u += x(i); }, r); break;  // redux.sync add
      case 1: parallel_reduce(n, KB200_LAMBDA(const i64 i, int& u) { if (x(i) < u) u = x(i); }, Min<int>(r)); break;
      case 2: parallel_reduce(n, KB200_LAMBDA(const i64 i, int& u) { if (x(i) > u) u = x(i); }, Max<int>(r)); break;
      case 3: parallel_reduce(n, KB200_LAMBDA(const i64 i, int& u) { u &= (x(i) | 0x0f0f0000); }, BAnd<int>(r)); break;
      case 4: parallel_reduce(n, KB200_LAMBDA(const i64 i, int& u) { u |= (x(i) & 0x00ff00ff); }, BOr<int>(r)); break;
      case 5: parallel_reduce(n, KB200_LAMBDA(const i64 i, int& u) { u = u && (x(i) != 12345678); }, LAnd<int>(r)); break;
      case 6: parallel_reduce(n, KB200_LAMBDA(const i64 i, int& u) { u = u || (x(i) == 3); }, LOr<int>(r)); break;
      default: return -1;
    }
    out[0] = r;
    return 0;
  });
}

// multi-result parallel_reduce (CombinedReducer, core/unit_test/TestReduce.hpp:545-640 "int_combined_reduce"):
// out[0..2] = sum, min, max over a RangePolicy; out[3..4] = sum + count over a 3-D MDRangePolicy; out[5] = device-View sum,
// out[6] = host scalar summed in the same call
int kb200_case_combined_reduce(const i64* hx, i64 n, i64 n0, i64 n1, i64 n2, i64* out) {
  return guarded([&] {
    View<const i64*> x = to_device(hx, n);
    i64 s = -1, mn = 0, mx = 0;
    parallel_reduce("combined", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, i64& a, i64& lo, i64& hi) {
      a += x(i);
      if (x(i) < lo) lo = x(i);
      if (x(i) > hi) hi = x(i);
    }, s, Min<i64>(mn), Max<i64>(mx));
    out[0] = s; out[1] = mn; out[2] = mx;
    i64 s3 = -1, cnt = -1;
    parallel_reduce("combined_md", MDRangePolicy<Rank<3>>({0, 0, 0}, {n0, n1, n2}), KB200_LAMBDA(const i64 i, const i64 j, const i64 k, i64& a, i64& c) {
      a += i + 10 * j + 100 * k;
      c += 1;
    }, s3, cnt);
    out[3] = s3; out[4] = cnt;
    View<i64> dv("dv");
    i64 hs = -1;
    parallel_reduce(n, KB200_LAMBDA(const i64 i, i64& a, i64& b) { a += 2 * x(i); b += 1; }, dv, hs);
    View<i64, HostSpace> hv("hv");
    deep_copy(hv, dv);
    out[5] = hv(); out[6] = hs;
    return 0;
  });
}

// ------------------------------------------------------------------ parallel_scan
int kb200_case_scan_i64(const i64* hx, i64* hy, i64 n, int inclusive, i64* total) {
  return guarded([&] {
    View<const i64*> x = to_device(hx, n);
    View<i64*> y("y", (size_t)n);
    i64 t = -7;
    if (inclusive)
      parallel_scan("scan", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, i64& u, const bool fin) { u += x(i); if (fin) y(i) = u; }, t);
    else
      parallel_scan("scan", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, i64& u, const bool fin) { if (fin) y(i) = u; u += x(i); }, t);
    to_host(hy, y);
    *total = t;
    return 0;
  });
}
int kb200_case_scan_f64_to_view_total(const double* hx, double* hy, i64 n, double* total) {
  return guarded([&] {
    View<const double*> x = to_device(hx, n);
    View<double*> y("y", (size_t)n);
    View<double> t("t");
    parallel_scan(RangePolicy<B200, IndexType<int>>(0, (int)n), KB200_LAMBDA(const int i, double& u, const bool fin) { if (fin) y(i) = u; u += x(i); }, t);
    fence();
    to_host(hy, y);
    View<double, HostSpace> h("h");
    deep_copy(h, t);
    *total = h();
    return 0;
  });
}
int kb200_case_scan_affine(const unsigned long long* ha, const unsigned long long* hb, i64 n, int inclusive, unsigned long long* hya,
                           unsigned long long* hyb, unsigned long long* total2) {
  return guarded([&] {
    AffineScan f;
    f.a = to_device(ha, n); f.b = to_device(hb, n);
    View<unsigned long long*> ya("ya", (size_t)n), yb("yb", (size_t)n);
    f.ya = ya; f.yb = yb; f.inclusive = inclusive;
    Affine t{0, 0};
    parallel_scan(RangePolicy<>(0, n), f, t);
    to_host(hya, ya); to_host(hyb, yb);
    total2[0] = t.a; total2[1] = t.b;
    return 0;
  });
}

// ------------------------------------------------------------------ parallel_for: stream kernels as lambdas
int kb200_case_stream(double* ha, double* hb, double* hc, i64 n, int iters, double scalar) {
  return guarded([&] {
    View<double*> a = to_device((const double*)ha, n), b = to_device((const double*)hb, n), c = to_device((const double*)hc, n);
    for (int it = 0; it < iters; ++it) {
      parallel_for("copy", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i) { c(i) = a(i); });
      parallel_for("scale", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i) { b(i) = nf_mul(scalar, c(i)); });
      parallel_for("add", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i) { c(i) = nf_add(a(i), b(i)); });
      parallel_for("triad", n, KB200_LAMBDA(const i64 i) { a(i) = nf_add(b(i), nf_mul(scalar, c(i))); });
    }
    fence();
    to_host(ha, a); to_host(hb, b); to_host(hc, c);
    return 0;
  });
}

// ------------------------------------------------------------------ MDRangePolicy
int kb200_case_mdrange_stencil(const double* hu, double* hv, i64 n0, i64 n1, i64 n2, double c0, double c1, double* out2, i64* loc2) {
  return guarded([&] {
    View<double***> u(view_alloc(WithoutInitializing, "u"), (size_t)n0, (size_t)n1, (size_t)n2);  // LayoutLeft on the device
    View<double***> v("v", (size_t)n0, (size_t)n1, (size_t)n2);
    deep_copy(u, View<const double***, LayoutLeft, HostSpace, MemoryTraits<Unmanaged>>(hu, (size_t)n0, (size_t)n1, (size_t)n2));
    using R = MinMaxLoc<double, i64>;
    R::value_type r;
    parallel_reduce("stencil7", MDRangePolicy<Rank<3>>({1, 1, 1}, {n0 - 1, n1 - 1, n2 - 1}),
        KB200_LAMBDA(const i64 i, const i64 j, const i64 k, R::value_type& m) {
          double s = nf_add(u(i - 1, j, k), u(i + 1, j, k));
          s = nf_add(s, u(i, j - 1, k));
          s = nf_add(s, u(i, j + 1, k));
          s = nf_add(s, u(i, j, k - 1));
          s = nf_add(s, u(i, j, k + 1));
          const double val = nf_add(nf_mul(c0, u(i, j, k)), nf_mul(c1, s));
          v(i, j, k) = val;
          const i64 loc = (i * n1 + j) * n2 + k;
          if (val < m.min_val) { m.min_val = val; m.min_loc = loc; }
          if (val > m.max_val) { m.max_val = val; m.max_loc = loc; }
        }, R(r));
    deep_copy(View<double***, LayoutLeft, HostSpace, MemoryTraits<Unmanaged>>(hv, (size_t)n0, (size_t)n1, (size_t)n2), v);
    out2[0] = r.min_val; out2[1] = r.max_val; loc2[0] = r.min_loc; loc2[1] = r.max_loc;
    return 0;
  });
}

}  // extern "C"

// rank 2..6 parallel_for + parallel_reduce with optional tiles and negative lower bounds (TestMDRange.hpp):
// every point (i0..) adds 1 to a hit counter at its flattened offset; the reduce sums a polynomial of the indices
template <int RANK, class... Idx>
struct MDHit {
  View<int*> hits;
  i64 lo[6], ext[6];
  KB200_INLINE_FUNCTION i64 flat(const i64* ix) const {
    i64 f = 0;
    for (int d = RANK - 1; d >= 0; --d) f = f * ext[d] + (ix[d] - lo[d]);
    return f;
  }
};
template <int RANK>
int mdrange_case(const i64* lower, const i64* upper, const i64* tile, int use_tile, int* hhits, i64 total_points, i64* poly_sum) {
  View<int*> hits("hits", (size_t)total_points);
  MDHit<RANK> h;
  h.hits = hits;
  for (int d = 0; d < 6; ++d) { h.lo[d] = d < RANK ? lower[d] : 0; h.ext[d] = d < RANK ? upper[d] - lower[d] : 1; }
  using Pol = MDRangePolicy<Rank<RANK>>;
  typename Pol::point_type lo, up;
  typename Pol::tile_type tl;
  for (int d = 0; d < RANK; ++d) { lo[d] = lower[d]; up[d] = upper[d]; tl[d] = use_tile ? tile[d] : 0; }
  Pol pol(lo, up, tl);
  i64 sum = 0;
  if constexpr (RANK == 2) {
    parallel_for(pol, KB200_LAMBDA(i64 a, i64 b) { i64 ix[2] = {a, b}; atomic_add(&h.hits(h.flat(ix)), 1); });
    parallel_reduce(pol, KB200_LAMBDA(i64 a, i64 b, i64& u) { u += a + 3 * b; }, sum);
  } else if constexpr (RANK == 3) {
    parallel_for(pol, KB200_LAMBDA(i64 a, i64 b, i64 c) { i64 ix[3] = {a, b, c}; atomic_add(&h.hits(h.flat(ix)), 1); });
    parallel_reduce(pol, KB200_LAMBDA(i64 a, i64 b, i64 c, i64& u) { u += a + 3 * b + 5 * c; }, sum);
  } else if constexpr (RANK == 4) {
    parallel_for(pol, KB200_LAMBDA(i64 a, i64 b, i64 c, i64 d) { i64 ix[4] = {a, b, c, d}; atomic_add(&h.hits(h.flat(ix)), 1); });
    parallel_reduce(pol, KB200_LAMBDA(i64 a, i64 b, i64 c, i64 d, i64& u) { u += a + 3 * b + 5 * c + 7 * d; }, sum);
  } else if constexpr (RANK == 5) {
    parallel_for(pol, KB200_LAMBDA(i64 a, i64 b, i64 c, i64 d, i64 e) { i64 ix[5] = {a, b, c, d, e}; atomic_add(&h.hits(h.flat(ix)), 1); });
    parallel_reduce(pol, KB200_LAMBDA(i64 a, i64 b, i64 c, i64 d, i64 e, i64& u) { u += a + 3 * b + 5 * c + 7 * d + 11 * e; }, sum);
  } else {
    parallel_for(pol, KB200_LAMBDA(i64 a, i64 b, i64 c, i64 d, i64 e, i64 f) { i64 ix[6] = {a, b, c, d, e, f}; atomic_add(&h.hits(h.flat(ix)), 1); });
    parallel_reduce(pol, KB200_LAMBDA(i64 a, i64 b, i64 c, i64 d, i64 e, i64 f, i64& u) { u += a + 3 * b + 5 * c + 7 * d + 11 * e + 13 * f; }, sum);
  }
  fence();
  to_host(hhits, hits);
  *poly_sum = sum;
  return 0;
}
extern "C" {
int kb200_case_mdrange(int rank, const i64* lower, const i64* upper, const i64* tile, int use_tile, int* hits, i64 total_points, i64* poly_sum) {
  return guarded([&] {
    switch (rank) {
      case 2: return mdrange_case<2>(lower, upper, tile, use_tile, hits, total_points, poly_sum);
      case 3: return mdrange_case<3>(lower, upper, tile, use_tile, hits, total_points, poly_sum);
      case 4: return mdrange_case<4>(lower, upper, tile, use_tile, hits, total_points, poly_sum);
      case 5: return mdrange_case<5>(lower, upper, tile, use_tile, hits, total_points, poly_sum);
      case 6: return mdrange_case<6>(lower, upper, tile, use_tile, hits, total_points, poly_sum);
    }
    return -1;
  });
}

// ------------------------------------------------------------------ TeamPolicy
int kb200_case_team_spmv(i64 nrows, const i64* hrm, const int* hci, const double* hva, i64 nnz, const double* hx, i64 ncols, double* hy,
                         int rows_per_team, int team_size, int vec) {
  return guarded([&] {
    View<const i64*> row_map = to_device(hrm, nrows + 1);
    View<const int*> col = to_device(hci, nnz);
    View<const double*> val = to_device(hva, nnz), x = to_device(hx, ncols);
    View<double*> y("y", (size_t)nrows);
    const int league = (int)((nrows + rows_per_team - 1) / rows_per_team);
    using TP = TeamPolicy<>;
    TP pol = team_size > 0 ? TP(league, team_size, vec) : TP(league, AUTO, vec);
    parallel_for("spmv", pol, KB200_LAMBDA(const TP::member_type& team) {
      const i64 first = (i64)team.league_rank() * rows_per_team;
      const i64 last = first + rows_per_team < nrows ? first + rows_per_team : nrows;
      parallel_for(TeamThreadRange(team, first, last), [&](const i64 row) {
        double s = 0;
        parallel_reduce(ThreadVectorRange(team, row_map(row), row_map(row + 1)),
                        [&](const i64 k, double& u) { u = nf_add(u, nf_mul(val(k), x(col(k)))); }, s);
        single(PerThread(team), [&]() { y(row) = s; });
      });
    });
    fence();
    to_host(hy, y);
    return 0;
  });
}

// league-level reduce + team_reduce / team_scan / team_broadcast / nested scans / scratch (TestTeam.hpp, TestTeamVector.hpp)
int kb200_case_team_collectives(int league, int team_size, int vec, int n_inner, i64* out) {
  return guarded([&] {
    using TP = TeamPolicy<>;
    View<i64*> errors("errors", 8);
    View<i64*> scan_out("scan_out", (size_t)league * n_inner);
    View<i64> gaccum("gaccum");
    TP pol(league, team_size, vec);
    pol.set_scratch_size(0, PerTeam(View<i64*>::shmem_size(n_inner)), PerThread(64));
    pol.set_scratch_size(1, PerTeam(1024), PerThread(128));
    i64 league_sum = 0;
    parallel_reduce("team_collectives", pol, KB200_LAMBDA(const TP::member_type& t, i64& update) {
      const int lr = t.league_rank(), tr = t.team_rank(), ts = t.team_size();
      // (1) nested TeamThreadRange reduce: sum_{i<n_inner} (i + lr)
      i64 s1 = 0;
      parallel_reduce(TeamThreadRange(t, n_inner), [&](const int i, i64& a) { a += i + lr; }, s1);
      if (s1 != (i64)n_inner * (n_inner - 1) / 2 + (i64)n_inner * lr) atomic_add(&errors(0), (i64)1);
      // (2) ThreadVectorRange reduce with a Max reducer
      i64 vmax = -1;
      parallel_reduce(ThreadVectorRange(t, 37), [&](const int i, i64& m) { if ((i * 7) % 37 > m) m = (i * 7) % 37; }, Max<i64>(vmax));
      if (vmax != 36) atomic_add(&errors(1), (i64)1);
      // (3) team_scan (exclusive over team_rank) with a global accumulator
      const i64 sc = t.team_scan((i64)(tr + 1));
      if (sc != (i64)tr * (tr + 1) / 2) atomic_add(&errors(2), (i64)1);
      (void)t.team_scan((i64)1, gaccum.data());
      // (4) team_broadcast + team_reduce with a Min reducer
      i64 bval = tr == ts - 1 ? 1000 + lr : -1;
      t.team_broadcast(bval, ts - 1);
      if (bval != 1000 + lr) atomic_add(&errors(3), (i64)1);
      i64 mn = 100 + tr;
      t.team_reduce(Min<i64>(mn));
      if (mn != 100) atomic_add(&errors(4), (i64)1);
      // (5) level-0 team scratch shared by the team + nested TeamThreadRange scan written through it
      View<i64*, ScratchMemorySpace<B200>, MemoryTraits<Unmanaged>> buf(t.team_scratch(0), (size_t)n_inner);
      if (buf.data() == nullptr) atomic_add(&errors(5), (i64)1);
      parallel_for(TeamVectorRange(t, n_inner), [&](const int i) { buf(i) = i % 5 + lr; });
      t.team_barrier();
      parallel_scan(TeamThreadRange(t, n_inner), [&](const int i, i64& p, const bool fin) {
        if (fin) scan_out((i64)lr * n_inner + i) = p;
        p += buf(i);
      });
      // (6) per-thread level-1 scratch is private to the thread
      i64* mine = (i64*)t.thread_scratch(1).get_shmem(64);
      if (mine == nullptr) atomic_add(&errors(6), (i64)1);
      else {
        single(PerThread(t), [&]() { mine[0] = 7000 + tr; });
        t.team_barrier();
        if (mine[0] != 7000 + tr) atomic_add(&errors(6), (i64)1);
      }
      // (7) ThreadVectorRange scan
      i64 vtotal = 0;
      parallel_scan(ThreadVectorRange(t, 19), [&](const int i, i64& p, const bool fin) { if (fin && i == 18) vtotal = p; p += i; });
      i64 vt = 0;
      parallel_reduce(ThreadVectorRange(t, 1), [&](const int, i64& m) { m += 0; }, vt);
      // league-level contribution: once per team
      single(PerTeam(t), [&]() { update += s1; });
    }, league_sum);
    fence();
    View<i64*, HostSpace> herr("herr", 8);
    deep_copy(herr, errors);
    for (int k = 0; k < 8; ++k) out[k] = herr(k);
    out[8] = league_sum;
    View<i64, HostSpace> hg("hg");
    deep_copy(hg, gaccum);
    out[9] = hg();
    // check the nested scan on the host: exclusive prefix of (i%5 + lr)
    View<i64*, HostSpace> hs("hs", (size_t)league * n_inner);
    deep_copy(hs, scan_out);
    i64 bad = 0;
    for (int lr = 0; lr < league; ++lr) {
      i64 run = 0;
      for (int i = 0; i < n_inner; ++i) { if (hs((size_t)lr * n_inner + i) != run) ++bad; run += i % 5 + lr; }
    }
    out[10] = bad;
    return 0;
  });
}

// multi-level scratch carving (TestTeam.hpp:920-1036 pattern): three team-level and three thread-level Views per level, carved by
// successive team_scratch(l) / thread_scratch(l) calls; out[0..3] = mismatches in team L0, thread L0, team L1, thread L1
int kb200_case_multilevel_scratch(int league, int team_size, int vec, i64* out) {
  return guarded([&] {
    using TP = TeamPolicy<>;
    using UV = View<double*, B200, MemoryTraits<Unmanaged>>;
    View<i64*> err("err", 4);
    TP pol(league, team_size, vec);
    pol.set_scratch_size(0, PerTeam(3 * UV::shmem_size(128)), PerThread(3 * UV::shmem_size(16)))
       .set_scratch_size(1, PerTeam(3 * UV::shmem_size(12800)), PerThread(3 * UV::shmem_size(1600)));
    parallel_for("mls", pol, KB200_LAMBDA(const TP::member_type& team) {
      UV at1(team.team_scratch(0), 128), ah1(team.thread_scratch(0), 16), at2(team.team_scratch(0), 128), ah2(team.thread_scratch(0), 16);
      UV bt1(team.team_scratch(1), 12800), bh1(team.thread_scratch(1), 1600), bt2(team.team_scratch(1), 12800), bh2(team.thread_scratch(1), 1600);
      UV at3(team.team_scratch(0), 128), ah3(team.thread_scratch(0), 16), bt3(team.team_scratch(1), 12800), bh3(team.thread_scratch(1), 1600);
      const int lr = team.league_rank(), tr = team.team_rank();
      parallel_for(TeamThreadRange(team, int(0), unsigned(128)), [&](const int& i) { at1(i) = 1e6 + i + lr * 1e5; at2(i) = 2e6 + i + lr * 1e5; at3(i) = 3e6 + i + lr * 1e5; });
      team.team_barrier();
      parallel_for(ThreadVectorRange(team, int(0), unsigned(16)), [&](const int& i) { ah1(i) = 1e6 + 1e5 * tr + 16 - i + lr * 1e5; ah2(i) = 2e6 + 1e5 * tr + 16 - i + lr * 1e5; ah3(i) = 3e6 + 1e5 * tr + 16 - i + lr * 1e5; });
      parallel_for(TeamThreadRange(team, int(0), unsigned(12800)), [&](const int& i) { bt1(i) = 1e6 + i + lr * 1e5; bt2(i) = 2e6 + i + lr * 1e5; bt3(i) = 3e6 + i + lr * 1e5; });
      team.team_barrier();
      parallel_for(ThreadVectorRange(team, 1600), [&](const int& i) { bh1(i) = 1e6 + 1e5 * tr + 16 - i + lr * 1e5; bh2(i) = 2e6 + 1e5 * tr + 16 - i + lr * 1e5; bh3(i) = 3e6 + 1e5 * tr + 16 - i + lr * 1e5; });
      team.team_barrier();
      parallel_for(TeamThreadRange(team, 0, 128), [&](const int& i) {
        if (at1(i) != 1e6 + i + lr * 1e5 || at2(i) != 2e6 + i + lr * 1e5 || at3(i) != 3e6 + i + lr * 1e5) atomic_add(&err(0), (i64)1); });
      team.team_barrier();
      parallel_for(ThreadVectorRange(team, 16), [&](const int& i) {
        if (ah1(i) != 1e6 + 1e5 * tr + 16 - i + lr * 1e5 || ah2(i) != 2e6 + 1e5 * tr + 16 - i + lr * 1e5 || ah3(i) != 3e6 + 1e5 * tr + 16 - i + lr * 1e5) atomic_add(&err(1), (i64)1); });
      parallel_for(TeamThreadRange(team, 0, 12800), [&](const int& i) {
        if (bt1(i) != 1e6 + i + lr * 1e5 || bt2(i) != 2e6 + i + lr * 1e5 || bt3(i) != 3e6 + i + lr * 1e5) atomic_add(&err(2), (i64)1); });
      team.team_barrier();
      parallel_for(ThreadVectorRange(team, 1600), [&](const int& i) {
        if (bh1(i) != 1e6 + 1e5 * tr + 16 - i + lr * 1e5 || bh2(i) != 2e6 + 1e5 * tr + 16 - i + lr * 1e5 || bh3(i) != 3e6 + 1e5 * tr + 16 - i + lr * 1e5) atomic_add(&err(3), (i64)1); });
      if (at1.data() == nullptr || at3.data() == nullptr || ah3.data() == nullptr || bt3.data() == nullptr || bh3.data() == nullptr) atomic_add(&err(0), (i64)1000000);
    });
    fence();
    View<i64*, HostSpace> h("h", 4);
    deep_copy(h, err);
    for (int k = 0; k < 4; ++k) out[k] = h(k);
    return 0;
  });
}

// ------------------------------------------------------------------ atomics (TestAtomics.hpp:456-598 style loops)
int kb200_case_atomics(i64 n, double* out) {
  return guarded([&] {
    View<i64*> ci("ci", 8);
    View<int*> c32("c32", 8);
    View<double*> cd("cd", 4);
    View<float*> cf("cf", 4);
    View<unsigned char*> c8("c8", 4);
    parallel_for(1, KB200_LAMBDA(const i64) { ci(2) = 1 << 30; ci(3) = -1; c32(2) = 1 << 30; c32(3) = -5; ci(4) = -1; c32(4) = 0; cd(1) = 1e300; cd(2) = -1e300; });
    parallel_for("atomics", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i) {
      atomic_add(&ci(0), i);                      // sum i
      (void)atomic_fetch_add(&ci(1), (i64)1);     // count
      atomic_min(&ci(2), (i64)((i * 7919) % 10007 + 5));
      atomic_max(&ci(3), (i64)((i * 7919) % 10007));
      atomic_and(&ci(4), (i64)~(1ll << (i % 40)));
      atomic_or(&ci(5), (i64)(1ll << (i % 50)));
      atomic_xor(&ci(6), (i64)(i * 0x9E3779B97F4A7C15ull));
      atomic_inc(&c32(0));
      atomic_sub(&c32(1), 2);
      atomic_min(&c32(2), (int)((i * 31) % 977 + 3));
      atomic_max(&c32(3), (int)((i * 31) % 977));
      atomic_add(&cd(0), 0.5);
      atomic_min(&cd(1), (double)((i * 13) % 1000) - 3.25);   // CAS-loop path
      atomic_max(&cd(2), (double)((i * 13) % 1000) + 0.75);
      atomic_add(&cf(0), 1.0f);
      atomic_add(&c8(i % 4), (unsigned char)1);               // 1-byte CAS path
      if (i == n / 2) { (void)atomic_exchange(&c32(5), 77); (void)atomic_compare_exchange(&c32(6), 0, 99); (void)atomic_compare_exchange(&c32(6), 5, 11); }
      atomic_store(&c32(7), 5);
    });
    fence();
    View<i64*, HostSpace> hi("hi", 8); View<int*, HostSpace> h32("h32", 8); View<double*, HostSpace> hd("hd", 4);
    View<float*, HostSpace> hf("hf", 4); View<unsigned char*, HostSpace> h8("h8", 4);
    deep_copy(hi, ci); deep_copy(h32, c32); deep_copy(hd, cd); deep_copy(hf, cf); deep_copy(h8, c8);
    for (int k = 0; k < 7; ++k) out[k] = (double)hi(k);
    for (int k = 0; k < 8; ++k) out[8 + k] = (double)h32(k);
    for (int k = 0; k < 3; ++k) out[16 + k] = hd(k);
    out[19] = hf(0);
    for (int k = 0; k < 4; ++k) out[20 + k] = (double)h8(k);
    // exact 64-bit values for the integer counters that exceed 2^53
    std::memcpy(out + 24, &hi(0), 8); std::memcpy(out + 25, &hi(4), 8); std::memcpy(out + 26, &hi(5), 8); std::memcpy(out + 27, &hi(6), 8);
    return 0;
  });
}

// ------------------------------------------------------------------ views / deep_copy / subview / mirrors / errors
int kb200_case_views(i64 n, i64* out) {
  return guarded([&] {
    View<i64*> a("a", (size_t)n);
    out[0] = (a.extent(0) == (size_t)n) && a.label() == "a" && a.use_count() == 1;
    auto h = create_mirror_view(a);
    for (i64 i = 0; i < n; ++i) h(i) = i * i;
    deep_copy(a, h);
    auto sub = subview(a, std::make_pair((i64)2, n - 1));
    out[1] = a.use_count();  // 2: a + sub share the record
    i64 s = 0;
    parallel_reduce(sub.extent(0), KB200_LAMBDA(const i64 i, i64& u) { u += sub(i); }, s);
    out[2] = s;
    auto back = create_mirror_view_and_copy(HostSpace(), a);
    out[3] = back(n - 1);
    View<i64*> zero("zero", (size_t)n);  // value-initialised
    i64 z = -1;
    parallel_reduce(n, KB200_LAMBDA(const i64 i, i64& u) { u += zero(i); }, z);
    out[4] = z;
    View<double**> m2("m2", 5, 7);       // LayoutLeft
    out[5] = (i64)m2.stride(1);
    View<double**, LayoutRight> m2r("m2r", 5, 7);
    out[6] = (i64)m2r.stride(0);
    int threw = 0;
    try { (void)subview(a, std::make_pair((i64)0, n + 1)); } catch (const std::runtime_error&) { threw = 1; }
    out[7] = threw;
    B200 space;
    out[8] = space.concurrency();
    out[9] = std::string(B200::name()) == "B200";
    return 0;
  });
}

}  // extern "C"

// ------------------------------------------------------------------ Array, MDRangePolicy from Arrays, result-less reduce, user reducer
namespace {
struct FinalWritesView {  // TestReduceCombinatorical.hpp:262-300: no result argument, final() stores on the device
  View<i64> result;
  View<const i64*> x;
  KB200_INLINE_FUNCTION void operator()(const i64 i, i64& u) const { u += x(i); }
  KB200_INLINE_FUNCTION void final(i64& u) const { result() = u + 7; }
};
template <class Space>
struct PlainSum {  // a user-written reducer: the memory space of result_view_type says where the result lives
  using reducer = PlainSum;
  using value_type = i64;
  using result_view_type = View<i64, Space, MemoryTraits<Unmanaged>>;
  result_view_type result;
  explicit PlainSum(i64* r) : result(r) {}
  KB200_INLINE_FUNCTION void join(i64& d, const i64& s) const { d += s; }
  KB200_INLINE_FUNCTION void init(i64& v) const { v = 0; }
  KB200_INLINE_FUNCTION i64& reference() const { return *result.data(); }
  KB200_INLINE_FUNCTION result_view_type view() const { return result; }
};
struct JoinAddsOne {  // TestReduceCombinatorical.hpp:27-56 (AddPlus)
  using reducer = JoinAddsOne;
  using value_type = i64;
  using result_view_type = View<i64, HostSpace, MemoryTraits<Unmanaged>>;
  result_view_type result;
  explicit JoinAddsOne(i64* r) : result(r) {}
  KB200_INLINE_FUNCTION void join(i64& d, const i64& s) const { d += s + 1; }
  KB200_INLINE_FUNCTION void init(i64& v) const { v = 0; }
  KB200_INLINE_FUNCTION i64& reference() const { return *result.data(); }
  KB200_INLINE_FUNCTION result_view_type view() const { return result; }
};
}  // namespace

extern "C" {
int kb200_case_utilities(const i64* hx, i64 n, i64* out) {
  return guarded([&] {
    auto x = to_device(hx, n);
    View<const i64*> xc = x;
    // out[0]: result-less parallel_reduce(n, functor) and (policy, functor)
    View<i64> r("r");
    parallel_reduce(n, FinalWritesView{r, xc});
    i64 h = 0;
    deep_copy(h, r);
    out[0] = h;
    parallel_reduce("labelled", RangePolicy<>(1, n), FinalWritesView{r, xc});
    deep_copy(h, r);
    out[1] = h;
    // out[2..3]: user reducer with a host result, then a device result
    i64 host_sum = -1;
    parallel_reduce(n, KB200_LAMBDA(const i64 i, i64& u) { u += xc(i); }, PlainSum<HostSpace>(&host_sum));
    out[2] = host_sum;
    View<i64> dsum("dsum");
    parallel_reduce(n, KB200_LAMBDA(const i64 i, i64& u) { u += 2 * xc(i); }, PlainSum<B200Space>(dsum.data()));
    deep_copy(h, dsum);
    out[3] = h;
    // out[4..7]: MDRangePolicy built from Arrays of another integer type, tile naming only the first dimension
    using Pol = MDRangePolicy<Rank<3>>;
    Pol p(Array<int, 3>{{1, 2, 3}}, Array<int, 3>{{41, 22, 13}}, Array<int, 1>{{8}});
    out[4] = p.m_tile[0] * 10000 + p.m_tile[1] * 100 + p.m_tile[2];
    Pol q({0, 0, 0}, {100, 100, 100});
    const auto rec = q.tile_size_recommended();
    out[5] = (rec[0] == q.m_tile[0] && rec[1] == q.m_tile[1] && rec[2] == q.m_tile[2]) ? (i64)(rec[0] * rec[1] * rec[2]) : -1;
    out[6] = q.max_total_tile_size();
    i64 cnt = 0;
    parallel_reduce(p, KB200_LAMBDA(const i64 i, const i64 j, const i64 k, i64& u) { u += i + 100 * j + 10000 * k; }, cnt);
    out[7] = cnt;
    // out[8]: Array / kokkos_swap / numbers on the device
    i64 bad = 0;
    parallel_reduce(64, KB200_LAMBDA(const i64 i, i64& u) {
      Array<i64, 3> a{{i, i + 1, i + 2}}, b{{-i, -i - 1, -i - 2}};
      kokkos_swap(a, b);
      i64 c[2] = {1, 2}, d[2] = {3, 4};
      kokkos_swap(c, d);
      auto [a0, a1, a2] = a;
      if (a0 != -i || a1 != -i - 1 || a2 != -i - 2 || b[2] != i + 2 || c[0] != 3 || d[1] != 2 || a.size() != 3 || !(a != b)) ++u;
      if (numbers::pi_v<float> != 3.14159265358979323846f || numbers::sqrt2 * numbers::sqrt2 < 1.999999 || numbers::inv_pi * numbers::pi > 1.000001) ++u;
    }, bad);
    out[8] = bad;
    // out[9..11]: an empty range is init() -> final() with no join, for every policy kind (the join here is not neutral)
    i64 e0 = -1, e1 = -1, e2 = -1;
    parallel_reduce(RangePolicy<>(5, 5), KB200_LAMBDA(const i64, i64& u) { u += 1; }, JoinAddsOne(&e0));
    parallel_reduce(MDRangePolicy<Rank<2>>({0, 0}, {0, 9}), KB200_LAMBDA(const i64, const i64, i64& u) { u += 1; }, JoinAddsOne(&e1));
    parallel_reduce(TeamPolicy<>(0, AUTO), KB200_LAMBDA(const TeamPolicy<>::member_type&, i64& u) { u += 1; }, JoinAddsOne(&e2));
    out[9] = e0; out[10] = e1; out[11] = e2;
    return 0;
  });
}

// ------------------------------------------------------------------ multi-dimensional subviews (LayoutStride), strided ViewCopy / ViewFill
int kb200_case_strided_subview(i64 n0, i64 n1, i64 n2, i64* out) {
  return guarded([&] {
    View<i64***> a("a", (size_t)n0, (size_t)n1, (size_t)n2);
    parallel_for(MDRangePolicy<Rank<3>>({0, 0, 0}, {n0, n1, n2}), KB200_LAMBDA(const i64 i, const i64 j, const i64 k) { a(i, j, k) = i + 100 * j + 10000 * k; });
    const i64 jfix = n1 / 2;
    auto sub = subview(a, std::make_pair((i64)1, n0 - 1), jfix, ALL);  // rank 2: (n0-2) x n2, strides (1, n0*n1)
    out[0] = (i64)sub.extent(0) * 1000 + (i64)sub.extent(1);
    out[1] = (i64)sub.stride(0) * 1000000 + (i64)sub.stride(1);
    out[2] = sub.span_is_contiguous() ? 1 : 0;
    out[3] = a.use_count();
    i64 bad = 0;
    parallel_reduce(MDRangePolicy<Rank<2>>({0, 0}, {(i64)sub.extent(0), (i64)sub.extent(1)}),
                    KB200_LAMBDA(const i64 i, const i64 k, i64& u) { if (sub(i, k) != (i + 1) + 100 * jfix + 10000 * k) ++u; }, bad);
    out[4] = bad;
    // strided -> contiguous copy on the device, then to the host
    View<i64**> dense("dense", sub.extent(0), sub.extent(1));
    deep_copy(dense, sub);
    auto hd = create_mirror_view_and_copy(HostSpace(), dense);
    i64 bad2 = 0;
    for (size_t i = 0; i < hd.extent(0); ++i)
      for (size_t k = 0; k < hd.extent(1); ++k)
        if (hd(i, k) != (i64)(i + 1) + 100 * jfix + 10000 * (i64)k) ++bad2;
    out[5] = bad2;
    // strided fill touches the window only; contiguous -> strided copy writes it back
    deep_copy(sub, (i64)-7);
    i64 cnt = 0;
    parallel_reduce(MDRangePolicy<Rank<3>>({0, 0, 0}, {n0, n1, n2}), KB200_LAMBDA(const i64 i, const i64 j, const i64 k, i64& u) { if (a(i, j, k) == -7) ++u; }, cnt);
    out[6] = cnt;
    deep_copy(sub, dense);
    i64 bad3 = 0;
    parallel_reduce(MDRangePolicy<Rank<3>>({0, 0, 0}, {n0, n1, n2}), KB200_LAMBDA(const i64 i, const i64 j, const i64 k, i64& u) { if (a(i, j, k) != i + 100 * j + 10000 * k) ++u; }, bad3);
    out[7] = bad3;
    // a rank-1 column out of a rank-2 LayoutRight view: stride = row length
    View<i64**, LayoutRight> m("m", 6, 9);
    parallel_for(MDRangePolicy<Rank<2>>({0, 0}, {6, 9}), KB200_LAMBDA(const i64 i, const i64 j) { m(i, j) = 10 * i + j; });
    auto col = subview(m, ALL, 4);
    i64 colsum = 0;
    parallel_reduce(col.extent(0), KB200_LAMBDA(const i64 i, i64& u) { u += col(i); }, colsum);
    out[8] = colsum;
    out[9] = (i64)col.stride(0);
    return 0;
  });
}

// ------------------------------------------------------------------ UniqueToken + resize / realloc
int kb200_case_tokens_resize(i64 n, i64* out) {
  return guarded([&] {
    using Experimental::UniqueToken;
    using Experimental::UniqueTokenScope;
    UniqueToken<B200, UniqueTokenScope::Instance> tokens{B200()};
    UniqueToken<B200, UniqueTokenScope::Instance> few(64, B200());
    out[0] = tokens.size();
    View<int*> inuse("inuse", (size_t)tokens.size()), inuse_few("inuse_few", 64);
    View<i64*> hits("hits", (size_t)tokens.size());
    View<i64> errors("errors");
    parallel_for(RangePolicy<>(0, n), KB200_LAMBDA(const i64) {
      Experimental::AcquireUniqueToken<B200, UniqueTokenScope::Instance> t(tokens);
      const int id = t.value();
      if (id < 0 || id >= tokens.size() || atomic_fetch_add(&inuse(id), 1) != 0) atomic_add(&errors(), (i64)1);
      hits(id) += 1;  // exclusive while the token is held
      if (atomic_fetch_add(&inuse(id), -1) != 1) atomic_add(&errors(), (i64)1);
    });
    // a caller-sized token set, used from one warp's worth of threads (never more holders than tokens)
    parallel_for(RangePolicy<>(0, 32), KB200_LAMBDA(const i64) {
      const int id = few.acquire();
      if (id < 0 || id >= few.size() || atomic_fetch_add(&inuse_few(id), 1) != 0) atomic_add(&errors(), (i64)1);
      if (atomic_fetch_add(&inuse_few(id), -1) != 1) atomic_add(&errors(), (i64)1);
      few.release(id);
    });
    i64 e = -1, total = 0;
    deep_copy(e, errors);
    out[1] = e;
    parallel_reduce(tokens.size(), KB200_LAMBDA(const i64 i, i64& u) { u += hits(i); }, total);
    out[2] = total;
    // resize keeps the common box, realloc does not
    View<i64*> a("a", (size_t)n);
    parallel_for(n, KB200_LAMBDA(const i64 i) { a(i) = i + 1; });
    resize(a, (size_t)(2 * n));
    i64 s1 = 0, s2 = 0;
    parallel_reduce(2 * n, KB200_LAMBDA(const i64 i, i64& u) { u += a(i); }, s1);
    out[3] = (i64)a.extent(0); out[4] = s1;
    resize(a, (size_t)(n / 2));
    parallel_reduce(n / 2, KB200_LAMBDA(const i64 i, i64& u) { u += a(i); }, s2);
    out[5] = (i64)a.extent(0); out[6] = s2;
    View<i64**> m("m", 5, 7);
    parallel_for(MDRangePolicy<Rank<2>>({0, 0}, {5, 7}), KB200_LAMBDA(const i64 i, const i64 j) { m(i, j) = 10 * i + j; });
    resize(m, 8, 4);
    i64 s3 = 0;
    parallel_reduce(MDRangePolicy<Rank<2>>({0, 0}, {8, 4}), KB200_LAMBDA(const i64 i, const i64 j, i64& u) { u += m(i, j); }, s3);
    out[7] = (i64)m.extent(0) * 10 + (i64)m.extent(1); out[8] = s3;
    realloc(m, 3, 3);
    i64 s4 = -1;
    parallel_reduce(MDRangePolicy<Rank<2>>({0, 0}, {3, 3}), KB200_LAMBDA(const i64 i, const i64 j, i64& u) { u += m(i, j); }, s4);
    out[9] = s4;
    return 0;
  });
}

// ------------------------------------------------------------------ MDRange default tile vs a register-heavy functor
// 72 live 64-bit values per thread: more than the 128 registers a 512-thread default tile leaves, so the launcher has to shrink
// the tile (Policy::impl_shrink_default_tile) instead of failing with "too many resources requested".
int kb200_case_mdrange_heavy(i64 n0, i64 n1, i64 n2, unsigned long long* hy) {
  return guarded([&] {
    using u64 = unsigned long long;
    View<u64***> y("y", (size_t)n0, (size_t)n1, (size_t)n2);
    parallel_for(MDRangePolicy<Rank<3>>({0, 0, 0}, {n0, n1, n2}), KB200_LAMBDA(const i64 i, const i64 j, const i64 k) {
      u64 a[72];
#pragma unroll
      for (int t = 0; t < 72; ++t) a[t] = (u64)(i + 3 * j + 7 * k) * (u64)(t + 1) + 0x9E3779B97F4A7C15ull;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int t = 0; t < 72; ++t) a[t] = a[t] * a[(t + 7) % 72] + (u64)(r + 1);
      }
      u64 s = 0;
#pragma unroll
      for (int t = 0; t < 72; ++t) s ^= a[t] + (u64)t;
      y(i, j, k) = s;
    });
    auto h = create_mirror_view_and_copy(HostSpace(), y);
    for (i64 k = 0; k < n2; ++k)
      for (i64 j = 0; j < n1; ++j)
        for (i64 i = 0; i < n0; ++i) hy[(k * n1 + j) * n0 + i] = h(i, j, k);
    return 0;
  });
}

}  // extern "C"

// ---- runtime-length array reductions of ANY length (value_type = T[], value_count; FunctorAnalysis.hpp:865-958): above 64 elements
// the accumulators live in global memory (impl/ArrayReduceKernel.hpp, "big" kernels)
namespace {
struct BigHistogram {
  using value_type = i64[];
  unsigned value_count;
  KB200_INLINE_FUNCTION void operator()(const i64 i, i64 dst[]) const { dst[i % (i64)value_count] += i; }
};
struct BigHistogramMD {
  using value_type = i64[];
  unsigned value_count;
  KB200_INLINE_FUNCTION void operator()(const int i, const int j, i64 dst[]) const { dst[(i * 7 + j) % (int)value_count] += 1 + j; }
  KB200_INLINE_FUNCTION void init(i64 dst[]) const { for (unsigned c = 0; c < value_count; ++c) dst[c] = 0; }
  KB200_INLINE_FUNCTION void join(i64 dst[], const i64 src[]) const { for (unsigned c = 0; c < value_count; ++c) dst[c] += src[c]; }
  KB200_INLINE_FUNCTION void final(i64 dst[]) const { dst[0] += 1000000; }
};
}  // namespace
extern "C" int kb200_case_array_reduce_big(i64 n, int count, i64 n0, i64 n1, i64* out_range, i64* out_md) {
  return guarded([&] {
    parallel_reduce("big array", RangePolicy<>(0, n), BigHistogram{(unsigned)count}, out_range);                    // result: host array
    View<i64*> r("r", (size_t)count);
    parallel_reduce("big array md", MDRangePolicy<Rank<2>>({0, 0}, {n0, n1}), BigHistogramMD{(unsigned)count}, r);  // result: device View
    fence();
    auto h = create_mirror_view_and_copy(HostSpace(), r);
    for (int c = 0; c < count; ++c) out_md[c] = h(c);
    return 0;
  });
}

// ---- a closure larger than the kernel parameter space (32 764 bytes): the reference's "global memory launch"
// (Cuda/Kokkos_Cuda_KernelLaunch.hpp:317-420; core/unit_test/TestGraph.hpp:423-445 builds such a functor)
namespace {
struct HugeClosure {
  View<i64*> out;
  unsigned char ballast[40000];
  KB200_INLINE_FUNCTION void operator()(const i64 i) const { out(i) = i * 3 + (i64)ballast[i % 40000]; }
};
}  // namespace
extern "C" int kb200_case_huge_closure(i64 n, i64* checksum) {
  return guarded([&] {
    static_assert(sizeof(HugeClosure) > 32764, "must not fit the kernel parameter space");
    View<i64*> out("out", (size_t)n);
    auto f = std::make_unique<HugeClosure>();
    f->out = out;
    for (int k = 0; k < 40000; ++k) f->ballast[k] = (unsigned char)(k * 7 + 1);
    parallel_for("huge closure", RangePolicy<>(0, n), *f);
    i64 s = 0;
    parallel_reduce("check", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, i64& u) { u += out(i); }, s);
    *checksum = s;
    return 0;
  });
}

// ---- a rank-6 tile whose four slow extents multiply to 128, beyond blockDim.z's 64: the reference's default tiling of a rank-6
// Iterate::Right policy ({2,2,2,2,2,16}, KokkosExp_MDRangePolicy.hpp:330-372; its ViewFill / ViewCopy of rank >= 6 LayoutRight Views
// launch exactly that, Kokkos_CopyViews.hpp:214-260).  The launch falls back to a linear block; every point is visited once.
extern "C" int kb200_case_mdrange_wide_tile(i64* out_for_sum, i64* out_reduce, i64* out_points) {
  return guarded([&] {
    const i64 e[6] = {5, 3, 4, 3, 5, 37};
    const size_t total = (size_t)(e[0] * e[1] * e[2] * e[3] * e[4] * e[5]);
    View<i64*> hit("hit", total);
    using P6 = MDRangePolicy<Rank<6>>;
    const P6 pol({0, 0, 0, 0, 0, 0}, {e[0], e[1], e[2], e[3], e[4], e[5]}, {2, 2, 2, 2, 2, 16});
    const i64 e1 = e[1], e2 = e[2], e3 = e[3], e4 = e[4], e5 = e[5];
    parallel_for("wide tile for", pol, KB200_LAMBDA(const i64 a, const i64 b, const i64 c, const i64 d, const i64 f, const i64 g) {
      const i64 flat = ((((a * e1 + b) * e2 + c) * e3 + d) * e4 + f) * e5 + g;
      atomic_add(&hit(flat), flat + 1);
    });
    i64 s = 0, r = 0, pts = 0;
    parallel_reduce("wide tile check", RangePolicy<>(0, (i64)total), KB200_LAMBDA(const i64 i, i64& u) { u += hit(i); }, s);
    parallel_reduce("wide tile reduce", pol, KB200_LAMBDA(const i64 a, const i64 b, const i64 c, const i64 d, const i64 f, const i64 g, i64& u) {
      u += ((((a * e1 + b) * e2 + c) * e3 + d) * e4 + f) * e5 + g + 1;
    }, r);
    parallel_reduce("wide tile once", RangePolicy<>(0, (i64)total), KB200_LAMBDA(const i64 i, i64& u) { u += (hit(i) == i + 1) ? 1 : 0; }, pts);
    *out_for_sum = s; *out_reduce = r; *out_points = pts;
    return 0;
  });
}
