// cases_perf.cu -- timing harness for the GENERIC (user-lambda) path of the B200 execution space.
//
// Each entry builds the Views on the device, runs `warm` untimed and `reps` timed invocations of ONE Kokkos-style
// call written exactly as a user writes it (KB200_LAMBDA functors, the same functors as the reference's
// benchmarks/stream, benchmarks/gups, core/unit_test/TestReducers.hpp, example/tutorial/Hierarchical_Parallelism),
// and reports {best, median} milliseconds measured with CUDA events on the instance's stream.  tools/configs_bench.py
// prints them next to the typed C-ABI fast paths; nothing here is a reported bench.py value.
#include <Kokkos_B200.hpp>
#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

using namespace kb200;
using i64 = long long;

namespace {
std::string g_perf_err;
template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_perf_err = e.what();
    return -100;
  }
}
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
KB200_INLINE_FUNCTION unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <class Call>
void time_call(Call&& call, int warm, int reps, double* out_ms) {
  B200 space;
  cudaStream_t s = space.cuda_stream();
  for (int k = 0; k < warm; ++k) call();
  space.fence();
  std::vector<float> ms((size_t)reps);
  std::vector<cudaEvent_t> e0((size_t)reps), e1((size_t)reps);
  for (int k = 0; k < reps; ++k) { cudaEventCreate(&e0[k]); cudaEventCreate(&e1[k]); }
  for (int k = 0; k < reps; ++k) {
    cudaEventRecord(e0[k], s);
    call();
    cudaEventRecord(e1[k], s);
  }
  space.fence();
  for (int k = 0; k < reps; ++k) { cudaEventElapsedTime(&ms[k], e0[k], e1[k]); cudaEventDestroy(e0[k]); cudaEventDestroy(e1[k]); }
  std::sort(ms.begin(), ms.end());
  out_ms[0] = ms[0];
  out_ms[1] = ms[ms.size() / 2];
}
}  // namespace

extern "C" {
const char* kb200_perf_last_error() { return g_perf_err.c_str(); }

// C1: parallel_reduce(RangePolicy(0,n), lambda(i, double& u){ u += a(i); }, sum)   [op 0]
//     MinMaxLoc reducer over the same View                                           [op 1]
//     result in a device View (no fence per call)                                    [op 2]
int kb200_perf_reduce(int op, i64 n, int warm, int reps, double* out_ms, double* check) {
  return guarded([&] {
    View<double*> a(view_alloc(WithoutInitializing, "a"), (size_t)n);
    parallel_for("fill", n, KB200_LAMBDA(const i64 i) { a(i) = (double)((((unsigned long long)i * 2654435761ull) >> 7) % 100); });
    fence();
    if (op == 0) {
      double r = 0;
      time_call([&] { parallel_reduce("sum", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, double& u) { u += a(i); }, r); }, warm, reps, out_ms);
      check[0] = r;
    } else if (op == 1) {
      using R = MinMaxLoc<double, i64>;
      R::value_type r;
      time_call([&] {
        parallel_reduce("mml", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, R::value_type& u) {
          const double v = a(i);
          if (v < u.min_val) { u.min_val = v; u.min_loc = i; }
          if (v > u.max_val) { u.max_val = v; u.max_loc = i; } }, R(r)); }, warm, reps, out_ms);
      check[0] = r.min_val; check[1] = r.max_val;
    } else {
      View<double> r("r");
      time_call([&] { parallel_reduce("sumv", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, double& u) { u += a(i); }, r); }, warm, reps, out_ms);
      View<double, HostSpace> h("h");
      deep_copy(h, r);
      check[0] = h();
    }
    return 0;
  });
}

// C2: benchmarks/stream copy [op 0] / triad [op 1] as lambdas
int kb200_perf_stream(int op, i64 n, int warm, int reps, double* out_ms) {
  return guarded([&] {
    View<double*> a(view_alloc(WithoutInitializing, "a"), (size_t)n), b(view_alloc(WithoutInitializing, "b"), (size_t)n),
        c(view_alloc(WithoutInitializing, "c"), (size_t)n);
    parallel_for("init", n, KB200_LAMBDA(const i64 i) { a(i) = 1.0; b(i) = 2.0; c(i) = 0.5; });
    fence();
    const double s = 3.0;
    if (op == 0) time_call([&] { parallel_for("copy", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i) { c(i) = a(i); }); }, warm, reps, out_ms);
    else time_call([&] { parallel_for("triad", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i) { a(i) = nf_add(b(i), nf_mul(s, c(i))); }); }, warm, reps, out_ms);
    return 0;
  });
}

// C2 tuning variants of the generic parallel_for launch shape (BLOCK x UNROLL, plain or persistent grid), copy lambda
int kb200_perf_for_variant(int variant, i64 n, int warm, int reps, double* out_ms) {
  return guarded([&] {
    View<double*> a(view_alloc(WithoutInitializing, "a"), (size_t)n), c(view_alloc(WithoutInitializing, "c"), (size_t)n);
    parallel_for("init", n, KB200_LAMBDA(const i64 i) { a(i) = 1.0; c(i) = 0.5; });
    fence();
    auto f = KB200_LAMBDA(const i64 i) { c(i) = a(i); };
    using F = decltype(f);
    using Body = Impl::FunctorForBody<F, void, i64>;
    Body body{f, 0};
    b200_instance* inst = B200().impl_instance();
    auto run = [&](auto launcher, int bps) {
      using L = decltype(launcher);
      time_call([&] { Impl::throw_on_error(L::run(inst, body, n, bps)); }, warm, reps, out_ms);
    };
    switch (variant) {
      case 0: run(Impl::RangeForLaunch<Body, 256, 4>{}, 0); break;
      case 1: run(Impl::RangeForLaunch<Body, 256, 1>{}, 0); break;
      case 2: run(Impl::RangeForLaunch<Body, 256, 2>{}, 0); break;
      case 3: run(Impl::RangeForLaunch<Body, 256, 8>{}, 0); break;
      case 4: run(Impl::RangeForLaunch<Body, 128, 1>{}, 0); break;
      case 5: run(Impl::RangeForLaunch<Body, 512, 1>{}, 0); break;
      case 6: run(Impl::RangeForLaunch<Body, 1024, 1>{}, 0); break;
      case 7: run(Impl::RangeForLaunch<Body, 256, 4>{}, 8); break;   // persistent
      case 8: run(Impl::RangeForLaunch<Body, 256, 1>{}, 8); break;
      case 9: run(Impl::RangeForLaunch<Body, 512, 2>{}, 0); break;
      default: return -1;
    }
    return 0;
  });
}

// C3 tuning variants of the generic scan's tile shape (BLOCK x ITEMS), same lambda; variant 0 = the public parallel_scan
int kb200_perf_scan_variant(int variant, i64 n, int warm, int reps, double* out_ms, i64* total) {
  return guarded([&] {
    View<i64*> x(view_alloc(WithoutInitializing, "x"), (size_t)n), y(view_alloc(WithoutInitializing, "y"), (size_t)n);
    parallel_for("fill", n, KB200_LAMBDA(const i64 i) { x(i) = (i64)((((unsigned long long)i * 2654435761ull) >> 7) % 7) - 3; });
    fence();
    i64 t = 0;
    auto f = KB200_LAMBDA(const i64 i, i64& u, const bool fin) { if (fin) y(i) = u; u += x(i); };
    using F = decltype(f);
    using Pol = RangePolicy<>;
    using Red = Impl::FunctorReducer<F, i64, void>;
    Pol pol(0, n);
    auto run = [&](auto tag) {
      using G = decltype(tag);
      time_call([&] { Impl::throw_on_error(G::run(pol, f, Red{f}, &t, nullptr)); }, warm, reps, out_ms);
    };
    switch (variant) {
      case 0: run(Impl::GenericScan<Pol, F, Red>{}); break;
      case 1: run(Impl::GenericScan<Pol, F, Red, 1024, 17, 1>{}); break;
      case 2: run(Impl::GenericScan<Pol, F, Red, 1024, 13, 4>{}); break;
      case 3: run(Impl::GenericScan<Pol, F, Red, 512, 13, 4>{}); break;
      case 4: run(Impl::GenericScan<Pol, F, Red, 512, 17, 4>{}); break;
      case 5: run(Impl::GenericScan<Pol, F, Red, 512, 13, 8>{}); break;
      case 6: run(Impl::GenericScan<Pol, F, Red, 1024, 19, 2>{}); break;
      case 7: run(Impl::GenericScan<Pol, F, Red, 1024, 21, 2>{}); break;
      case 8: run(Impl::GenericScan<Pol, F, Red, 256, 13, 8>{}); break;
      case 9: run(Impl::GenericScan<Pol, F, Red, 512, 21, 4>{}); break;
      case 10: run(Impl::GenericScan<Pol, F, Red, 768, 13, 4>{}); break;
      case 11: run(Impl::GenericScan<Pol, F, Red, 1024, 13, 8>{}); break;
      case 12: run(Impl::GenericScan<Pol, F, Red, 512, 25, 4>{}); break;
      case 13: run(Impl::GenericScan<Pol, F, Red, 1024, 21, 4>{}); break;
      default: return -1;
    }
    *total = t;
    return 0;
  });
}

// C3: parallel_scan exclusive prefix sum, lambda form, with total
int kb200_perf_scan(i64 n, int warm, int reps, double* out_ms, i64* total) {
  return guarded([&] {
    View<i64*> x(view_alloc(WithoutInitializing, "x"), (size_t)n), y(view_alloc(WithoutInitializing, "y"), (size_t)n);
    parallel_for("fill", n, KB200_LAMBDA(const i64 i) { x(i) = (i64)((((unsigned long long)i * 2654435761ull) >> 7) % 7) - 3; });
    fence();
    i64 t = 0;
    time_call([&] { parallel_scan("scan", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, i64& u, const bool fin) { if (fin) y(i) = u; u += x(i); }, t); }, warm, reps, out_ms);
    *total = t;
    return 0;
  });
}

// C4: MDRangePolicy<Rank<3>> 7-point stencil + MinMaxLoc (optionally storing v)
int kb200_perf_mdrange_stencil(i64 n0, i64 n1, i64 n2, int store, int warm, int reps, double* out_ms, double* check) {
  return guarded([&] {
    View<double***> u(view_alloc(WithoutInitializing, "u"), (size_t)n0, (size_t)n1, (size_t)n2);
    View<double***> v(view_alloc(WithoutInitializing, "v"), store ? (size_t)n0 : 1, store ? (size_t)n1 : 1, store ? (size_t)n2 : 1);
    parallel_for("fill", MDRangePolicy<Rank<3>>({0, 0, 0}, {n0, n1, n2}), KB200_LAMBDA(const i64 i, const i64 j, const i64 k) {
      u(i, j, k) = (double)(mix64((unsigned long long)((i * n1 + j) * n2 + k)) >> 11) * (1.0 / 9007199254740992.0);
    });
    fence();
    using R = MinMaxLoc<double, i64>;
    R::value_type r;
    const double c0 = 0.5, c1 = 0.125;
    time_call([&] {
      parallel_reduce("stencil7", MDRangePolicy<Rank<3>>({1, 1, 1}, {n0 - 1, n1 - 1, n2 - 1}),
          KB200_LAMBDA(const i64 i, const i64 j, const i64 k, R::value_type& m) {
            double s = nf_add(u(i - 1, j, k), u(i + 1, j, k));
            s = nf_add(s, u(i, j - 1, k));
            s = nf_add(s, u(i, j + 1, k));
            s = nf_add(s, u(i, j, k - 1));
            s = nf_add(s, u(i, j, k + 1));
            const double val = nf_add(nf_mul(c0, u(i, j, k)), nf_mul(c1, s));
            if (store) v(i, j, k) = val;
            const i64 loc = (i * n1 + j) * n2 + k;
            if (val < m.min_val) { m.min_val = val; m.min_loc = loc; }
            if (val > m.max_val) { m.max_val = val; m.max_loc = loc; }
          }, R(r)); }, warm, reps, out_ms);
    check[0] = r.min_val; check[1] = r.max_val; check[2] = (double)r.min_loc; check[3] = (double)r.max_loc;
    return 0;
  });
}

// C5a: GUPS update loop as a lambda: atomic_add [op 0] / atomic_fetch_xor [op 1] (benchmarks/gups/gups.cpp:83-97)
int kb200_perf_gups(int op, i64 table_len, i64 m, int warm, int reps, double* out_ms) {
  return guarded([&] {
    View<i64*> table(view_alloc(WithoutInitializing, "table"), (size_t)table_len), idx(view_alloc(WithoutInitializing, "idx"), (size_t)m);
    parallel_for("fill", table_len, KB200_LAMBDA(const i64 i) { table(i) = 10101010101ll; });
    parallel_for("idx", m, KB200_LAMBDA(const i64 i) { idx(i) = (i64)(mix64((unsigned long long)i + 20230913ull) % (unsigned long long)table_len); });
    fence();
    const i64 datum = -1;
    if (op == 0) time_call([&] { parallel_for("gups", RangePolicy<>(0, m), KB200_LAMBDA(const i64 i) { atomic_add(&table(idx(i)), datum); }); }, warm, reps, out_ms);
    else time_call([&] { parallel_for("gups", RangePolicy<>(0, m), KB200_LAMBDA(const i64 i) { (void)atomic_fetch_xor(&table(idx(i)), datum); }); }, warm, reps, out_ms);
    return 0;
  });
}

// C5b: TeamPolicy + TeamThreadRange + ThreadVectorRange nested-reduce CRS SpMV on caller-provided DEVICE arrays
int kb200_perf_team_spmv(i64 nrows, const i64* d_row_map, const int* d_col, const double* d_val, const double* d_x, double* d_y,
                         int rows_per_team, int team_size, int vec, int warm, int reps, double* out_ms) {
  return guarded([&] {
    using U = MemoryTraits<Unmanaged>;
    View<const i64*, U> row_map(d_row_map, (size_t)nrows + 1);
    View<const int*, U> col(d_col, 1);      // extents of the unmanaged wrappers are not checked on the device
    View<const double*, U> val(d_val, 1), x(d_x, 1);
    View<double*, U> y(d_y, (size_t)nrows);
    const int league = (int)((nrows + rows_per_team - 1) / rows_per_team);
    using TP = TeamPolicy<>;
    TP pol = team_size > 0 ? TP(league, team_size, vec) : TP(league, AUTO, vec);
    time_call([&] {
      parallel_for("spmv", pol, KB200_LAMBDA(const TP::member_type& team) {
        const i64 first = (i64)team.league_rank() * rows_per_team;
        const i64 last = first + rows_per_team < nrows ? first + rows_per_team : nrows;
        parallel_for(TeamThreadRange(team, first, last), [&](const i64 row) {
          double s = 0;
          parallel_reduce(ThreadVectorRange(team, row_map(row), row_map(row + 1)),
                          [&](const i64 k, double& u) { u = nf_add(u, nf_mul(val(k), x(col(k)))); }, s);
          single(PerThread(team), [&]() { y(row) = s; });
        });
      }); }, warm, reps, out_ms);
    return 0;
  });
}

// launch latency (benchmarks/launch_latency/launch_latency.cpp): batches of `batch` tiny kernels, mean microseconds per call
//   op 0: parallel_for over n;  op 1: parallel_reduce to a scalar (kernel + fence per call);  op 2: parallel_reduce to a device View
int kb200_perf_launch_latency(int op, i64 n, int batch, int reps, double* out_us) {
  return guarded([&] {
    View<double*> a("a", (size_t)(n > 0 ? n : 1));
    View<double> rv("rv");
    B200 space;
    double best = 1e30;
    for (int r = 0; r < reps + 1; ++r) {
      space.fence();
      const auto t0 = std::chrono::steady_clock::now();
      double s = 0;
      for (int k = 0; k < batch; ++k) {
        if (op == 0) parallel_for("lat", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i) { a(i) = 1.0; });
        else if (op == 1) parallel_reduce("lat", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, double& u) { u += a(i); }, s);
        else parallel_reduce("lat", RangePolicy<>(0, n), KB200_LAMBDA(const i64 i, double& u) { u += a(i); }, rv);
      }
      space.fence();
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / batch;
      if (r > 0 && us < best) best = us;
    }
    out_us[0] = best;
    return 0;
  });
}
}  // extern "C"
