// kokkos_arms.cu -- the headline workloads written the way a Kokkos user writes them (KOKKOS_LAMBDA functors over
// Kokkos::View, RangePolicy / MDRangePolicy / TeamPolicy), compiled ONCE against the UNMODIFIED reference headers and
// dispatched to three arms that share the same device buffers and the same CUDA stream:
//   arm 0  Kokkos::B200   this repository's execution space, attached by kokkos_b200/adapter/Kokkos_B200_Space.hpp
//   arm 1  Kokkos::Cuda   the reference's own generic CUDA backend built for sm_100 (comparator)
//   arm 2  CUB            cub::DeviceReduce / cub::DeviceScan, the path the reference's std_algorithms take on CUDA
//                         (algorithms/src/std_algorithms/impl/Kokkos_InclusiveScan.hpp:156-176)
// bench.py times the calls with CUDA events on that stream; nothing here times itself.  Every call is asynchronous
// (results go to device memory), so a timed region contains kernels only.
// Functors follow the reference's own benchmark sources: benchmarks/stream/stream-kokkos.cpp:217-231 (copy, triad),
// benchmarks/gups/gups-kokkos.cpp (atomic update loop), core/unit_test/TestReducers.hpp (MinMaxLoc),
// example/tutorial/Hierarchical_Parallelism (nested TeamThreadRange reduce as CRS SpMV).
#include <Kokkos_B200_StdAlgorithms.hpp>

#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

#include <chrono>
#include <cstdint>
#include <memory>
#include <string>

namespace {
using i64 = long long;
template <class T>
using DView = Kokkos::View<T*, Kokkos::CudaSpace, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
template <class T>
using DScalar = Kokkos::View<T, Kokkos::CudaSpace, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
template <class T>
using DField = Kokkos::View<T***, Kokkos::LayoutLeft, Kokkos::CudaSpace, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;

// separately rounded add / multiply (no FMA contraction), so results are bit-identical to the OpenMP oracle's
KOKKOS_INLINE_FUNCTION double nf_add(double a, double b) {
  KOKKOS_IF_ON_DEVICE((return __dadd_rn(a, b);))
  KOKKOS_IF_ON_HOST((return a + b;))
}
KOKKOS_INLINE_FUNCTION double nf_mul(double a, double b) {
  KOKKOS_IF_ON_DEVICE((return __dmul_rn(a, b);))
  KOKKOS_IF_ON_HOST((return a * b;))
}

struct Arms {
  cudaStream_t stream = nullptr;
  std::unique_ptr<Kokkos::B200> b200;
  std::unique_ptr<Kokkos::Cuda> cuda;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  bool we_initialized = false;
};
Arms g;
std::string g_err;

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

int cub_scratch(size_t bytes) {
  if (bytes <= g.cub_tmp_bytes) return 0;
  if (g.cub_tmp) cudaFree(g.cub_tmp);
  g.cub_tmp_bytes = 0;
  if (cudaMalloc(&g.cub_tmp, bytes) != cudaSuccess) { g_err = "cub scratch allocation failed"; return -1; }
  g.cub_tmp_bytes = bytes;
  return 0;
}

// ---- the user-level calls, templated on the execution space -------------------------------------------------------
template <class Space>
void reduce_sum(const Space& s, const double* x, i64 n, double* result_dev) {
  DView<const double> a(x, (size_t)n);
  DScalar<double> r(result_dev);
  Kokkos::parallel_reduce("arms::reduce_sum", Kokkos::RangePolicy<Space, Kokkos::IndexType<i64>>(s, 0, n),
                          KOKKOS_LAMBDA(const i64 i, double& u) { u += a(i); }, r);
}

template <class Space>
void scan_excl(const Space& s, const i64* x, i64* y, i64 n, i64* total_dev) {
  DView<const i64> a(x, (size_t)n);
  DView<i64> b(y, (size_t)n);
  DScalar<i64> t(total_dev);
  Kokkos::parallel_scan("arms::scan_excl", Kokkos::RangePolicy<Space, Kokkos::IndexType<i64>>(s, 0, n),
                        KOKKOS_LAMBDA(const i64 i, i64& u, const bool fin) {
                          const i64 v = a(i);
                          if (fin) b(i) = u;
                          u += v;
                        }, t);
}

// the std_algorithms form of the same prefix sum (blocking: the reference fences at its end)
template <class Space>
void std_exclusive_scan(const Space& s, const i64* x, i64* y, i64 n) {
  DView<const i64> a(x, (size_t)n);
  DView<i64> b(y, (size_t)n);
  Kokkos::Experimental::exclusive_scan("arms::std_exclusive_scan", s, a, b, (i64)0);
}

template <class Space>
void stream_copy(const Space& s, const double* a_, double* c_, i64 n) {
  DView<const double> a(a_, (size_t)n);
  DView<double> c(c_, (size_t)n);
  Kokkos::parallel_for("arms::copy", Kokkos::RangePolicy<Space, Kokkos::IndexType<i64>>(s, 0, n), KOKKOS_LAMBDA(const i64 i) { c(i) = a(i); });
}

template <class Space>
void stream_triad(const Space& s, double* a_, const double* b_, const double* c_, double scalar, i64 n) {
  DView<double> a(a_, (size_t)n);
  DView<const double> b(b_, (size_t)n), c(c_, (size_t)n);
  // a = b + scalar*c with separately rounded multiply and add (what the OpenMP oracle computes without contraction)
  Kokkos::parallel_for("arms::triad", Kokkos::RangePolicy<Space, Kokkos::IndexType<i64>>(s, 0, n),
                       KOKKOS_LAMBDA(const i64 i) { a(i) = nf_add(b(i), nf_mul(scalar, c(i))); });
}

using MML = Kokkos::MinMaxLoc<double, i64, Kokkos::CudaSpace>;
template <class Space>
void stencil7_minmaxloc(const Space& s, const double* u_, i64 n0, i64 n1, i64 n2, double c0, double c1, MML::value_type* result_dev,
                        i64 t0 = 0, i64 t1 = 0, i64 t2 = 0) {
  DField<const double> u(u_, (size_t)n0, (size_t)n1, (size_t)n2);
  Kokkos::View<MML::value_type, Kokkos::CudaSpace, Kokkos::MemoryTraits<Kokkos::Unmanaged>> r(result_dev);
  using Policy = Kokkos::MDRangePolicy<Space, Kokkos::Rank<3>>;  // default index type, as in the reference's own MDRange tests
  // t0 == 0: the backend's default tile (what a user who does not tune gets)
  const Policy policy = t0 > 0 ? Policy(s, {1, 1, 1}, {n0 - 1, n1 - 1, n2 - 1}, {t0, t1, t2}) : Policy(s, {1, 1, 1}, {n0 - 1, n1 - 1, n2 - 1});
  Kokkos::parallel_reduce("arms::stencil7", policy,
                          KOKKOS_LAMBDA(const int i, const int j, const int k, MML::value_type& m) {
                            const double nb = nf_add(nf_add(nf_add(nf_add(nf_add(u(i - 1, j, k), u(i + 1, j, k)), u(i, j - 1, k)), u(i, j + 1, k)), u(i, j, k - 1)), u(i, j, k + 1));
                            const double v = nf_add(nf_mul(c0, u(i, j, k)), nf_mul(c1, nb));
                            const i64 loc = ((i64)i * n1 + j) * n2 + k;
                            if (v < m.min_val || (v == m.min_val && loc < m.min_loc)) { m.min_val = v; m.min_loc = loc; }
                            if (v > m.max_val || (v == m.max_val && loc < m.max_loc)) { m.max_val = v; m.max_loc = loc; }
                          }, MML(r));
}

// int64_t (= long on LP64) as in benchmarks/gups: desul has a native red/atom path for it, `long long` would take its CAS loop
template <class Space>
void gups_add(const Space& s, i64* table_, i64 table_len, const i64* idx_, i64 m, i64 datum_) {
  DView<int64_t> table((int64_t*)table_, (size_t)table_len);
  DView<const int64_t> idx((const int64_t*)idx_, (size_t)m);
  const int64_t datum = (int64_t)datum_;
  Kokkos::parallel_for("arms::gups", Kokkos::RangePolicy<Space, Kokkos::IndexType<i64>>(s, 0, m),
                       KOKKOS_LAMBDA(const i64 i) { Kokkos::atomic_add(&table(idx(i)), datum); });
}

template <class Space>
void spmv(const Space& s, i64 nrows, const i64* row_map_, const int* col_, const double* val_, const double* x_, double* y_, i64 nnz, i64 ncols) {
  DView<const i64> row_map(row_map_, (size_t)nrows + 1);
  DView<const int> col(col_, (size_t)nnz);
  DView<const double> val(val_, (size_t)nnz), x(x_, (size_t)ncols);
  DView<double> y(y_, (size_t)nrows);
  using Policy = Kokkos::TeamPolicy<Space>;
  using Member = typename Policy::member_type;
  const int rows_per_team = 16, vec = 8;
  const int league = (int)((nrows + rows_per_team - 1) / rows_per_team);
  Kokkos::parallel_for("arms::spmv", Policy(s, league, rows_per_team, vec), KOKKOS_LAMBDA(const Member& t) {
    const i64 row0 = (i64)t.league_rank() * rows_per_team;
    Kokkos::parallel_for(Kokkos::TeamThreadRange(t, rows_per_team), [&](const int r) {
      const i64 row = row0 + r;
      if (row >= nrows) return;
      double acc = 0;
      Kokkos::parallel_reduce(Kokkos::ThreadVectorRange(t, (int)(row_map(row + 1) - row_map(row))),
                              [&](const int k, double& u) { const i64 e = row_map(row) + k; u = nf_add(u, nf_mul(val(e), x(col(e)))); }, acc);
      Kokkos::single(Kokkos::PerThread(t), [&]() { y(row) = acc; });
    });
  });
}

// benchmarks/launch_latency/launch_latency.cpp in miniature: `batch` back-to-back launches over n elements, host clock around the
// batch including the closing fence.  op 0: parallel_for; 1: parallel_reduce into a device View (asynchronous); 2: parallel_reduce
// into a host scalar (each call fences).
template <class Space>
double launch_latency_us(const Space& s, int op, i64 n, int batch, double* scratch_dev) {
  DView<double> a(scratch_dev, (size_t)(n > 0 ? n : 1));
  DScalar<double> r(scratch_dev + (n > 0 ? n : 1));
  using Policy = Kokkos::RangePolicy<Space, Kokkos::IndexType<i64>>;
  auto one = [&]() {
    if (op == 0) Kokkos::parallel_for("arms::lat_for", Policy(s, 0, n), KOKKOS_LAMBDA(const i64 i) { a(i) = 1.0; });
    else if (op == 1) Kokkos::parallel_reduce("arms::lat_red_view", Policy(s, 0, n), KOKKOS_LAMBDA(const i64 i, double& u) { u += a(i); }, r);
    else { double h = 0; Kokkos::parallel_reduce("arms::lat_red_scalar", Policy(s, 0, n), KOKKOS_LAMBDA(const i64 i, double& u) { u += a(i); }, h); }
  };
  for (int k = 0; k < 20; ++k) one();
  s.fence();
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < batch; ++k) one();
  s.fence();
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::micro>(t1 - t0).count() / batch;
}

template <class F>
int dispatch(int arm, F&& f) {
  if (!g.b200) { g_err = "kka_init has not been called"; return -2; }
  if (arm == 0) return guarded([&] { f(*g.b200); });
  if (arm == 1) return guarded([&] { f(*g.cuda); });
  g_err = "no such arm for this workload";
  return -3;
}
}  // namespace

extern "C" {
const char* kka_last_error() { return g_err.c_str(); }

// Kokkos::initialize on `device`, one Kokkos::B200 and one Kokkos::Cuda instance on the caller's stream
int kka_init(int device, void* stream) {
  return guarded([&] {
    if (!Kokkos::is_initialized()) {
      Kokkos::initialize(Kokkos::InitializationSettings().set_device_id(device).set_num_threads(1).set_disable_warnings(true));
      g.we_initialized = true;
    }
    g.stream = (cudaStream_t)stream;
    g.b200.reset(new Kokkos::B200(g.stream));
    g.cuda.reset(new Kokkos::Cuda(g.stream));
  });
}
int kka_finalize() {
  return guarded([&] {
    g.b200.reset();
    g.cuda.reset();
    if (g.cub_tmp) cudaFree(g.cub_tmp);
    g.cub_tmp = nullptr;
    g.cub_tmp_bytes = 0;
    if (g.we_initialized && Kokkos::is_initialized() && !Kokkos::is_finalized()) Kokkos::finalize();
    g.we_initialized = false;
  });
}

int kka_reduce_sum_f64(int arm, const double* x, i64 n, double* result_dev) {
  if (arm == 2) {
    size_t need = 0;
    cub::DeviceReduce::Sum(nullptr, need, x, result_dev, n, g.stream);
    if (cub_scratch(need)) return -1;
    return cub::DeviceReduce::Sum(g.cub_tmp, need, x, result_dev, n, g.stream) == cudaSuccess ? 0 : -1;
  }
  return dispatch(arm, [&](auto& s) { reduce_sum(s, x, n, result_dev); });
}
// total_dev receives the scan total on arms 0/1; CUB's ExclusiveSum has no total (left untouched)
int kka_scan_excl_i64(int arm, const i64* x, i64* y, i64 n, i64* total_dev) {
  if (arm == 2) {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, x, y, n, g.stream);
    if (cub_scratch(need)) return -1;
    return cub::DeviceScan::ExclusiveSum(g.cub_tmp, need, x, y, n, g.stream) == cudaSuccess ? 0 : -1;
  }
  return dispatch(arm, [&](auto& s) { scan_excl(s, x, y, n, total_dev); });
}
// Kokkos::Experimental::exclusive_scan(exec, in, out, 0) (std_algorithms); arm 2 is the same CUB call as above
int kka_std_exclusive_scan_i64(int arm, const i64* x, i64* y, i64 n) {
  if (arm == 2) return kka_scan_excl_i64(2, x, y, n, nullptr);
  return dispatch(arm, [&](auto& s) { std_exclusive_scan(s, x, y, n); });
}
int kka_stream_copy_f64(int arm, const double* a, double* c, i64 n) {
  return dispatch(arm, [&](auto& s) { stream_copy(s, a, c, n); });
}
int kka_stream_triad_f64(int arm, double* a, const double* b, const double* c, double scalar, i64 n) {
  return dispatch(arm, [&](auto& s) { stream_triad(s, a, b, c, scalar, n); });
}
// result_dev: {min_val, max_val, min_loc, max_loc} = Kokkos::MinMaxLoc<double, int64>::value_type (32 bytes)
int kka_stencil7_minmaxloc_f64(int arm, const double* u, i64 n0, i64 n1, i64 n2, double c0, double c1, void* result_dev) {
  static_assert(sizeof(MML::value_type) == 32, "MinMaxLoc value layout");
  return dispatch(arm, [&](auto& s) { stencil7_minmaxloc(s, u, n0, n1, n2, c0, c1, (MML::value_type*)result_dev); });
}
int kka_stencil7_minmaxloc_f64_tiled(int arm, const double* u, i64 n0, i64 n1, i64 n2, double c0, double c1, void* result_dev, i64 t0, i64 t1, i64 t2) {
  return dispatch(arm, [&](auto& s) { stencil7_minmaxloc(s, u, n0, n1, n2, c0, c1, (MML::value_type*)result_dev, t0, t1, t2); });
}
// scratch_dev: at least n + 1 doubles of device memory
int kka_launch_latency_us(int arm, int op, i64 n, int batch, double* scratch_dev, double* out_us) {
  return dispatch(arm, [&](auto& s) { *out_us = launch_latency_us(s, op, n, batch, scratch_dev); });
}
int kka_gups_add_i64(int arm, i64* table, i64 table_len, const i64* idx, i64 m, i64 datum) {
  return dispatch(arm, [&](auto& s) { gups_add(s, table, table_len, idx, m, datum); });
}
int kka_spmv_crs_f64(int arm, i64 nrows, const i64* row_map, const int* col, const double* val, const double* x, double* y, i64 nnz, i64 ncols) {
  return dispatch(arm, [&](auto& s) { spmv(s, nrows, row_map, col, val, x, y, nnz, ncols); });
}
}  // extern "C"
