/* omp_oracle.c -- TEST INFRASTRUCTURE ONLY.  See omp_oracle.h.
 *
 * Restates, in plain C, what Kokkos::OpenMP computes for the hot path.  Reference lines followed
 * (relative to /root/reference/core/src):
 *   auto chunk size ............ Kokkos_ExecPolicy.hpp:202-231          (oracle_auto_chunk)
 *   reduce work partition ...... impl/Kokkos_HostThreadTeam.hpp:357-389 (oracle_reduce_partition)
 *   reduce execute ............. OpenMP/Kokkos_OpenMP_Parallel_Reduce.hpp:69-162
 *                                 per-thread init + left fold :126-139, thread-ordered join :147-151
 *   scan work range ............ Kokkos_ExecPolicy.hpp:303-337          (oracle_scan_partition)
 *   scan execute ............... OpenMP/Kokkos_OpenMP_Parallel_Scan.hpp:69-140
 *                                 pass 1 final=false :100-107, serial scan of thread totals :109-131,
 *                                 pass 2 final=true from the thread's base :133-138
 *   reducer init/join .......... Kokkos_Parallel_Reduce.hpp:33-72 (Sum) 123-212 (Min/Max)
 *                                 411-463,471-523 (MinLoc/MaxLoc) 536-587 (MinMax) 600-662 (MinMaxLoc)
 *   identities ................. Kokkos_ReductionIdentity.hpp:143-166,357-380
 *   default join (dst += src) .. impl/Kokkos_FunctorAnalysis.hpp:604-613
 * The "threads" are simulated one after another: only the association order matters for the result.
 * Floating-point expressions are evaluated exactly as written (build with -ffp-contract=off).
 */
#include "omp_oracle.h"
#include <float.h>
#include <limits.h>
#include <stddef.h>

#define IDX_IDENTITY INT64_MAX /* reduction_identity<int64_t>::min() */

int64_t oracle_auto_chunk(int64_t n, int threads) {
  int64_t concurrency = threads > 0 ? threads : 1;
  int64_t c = 1;
  while (c * 100 * concurrency < n) c *= 2;
  if (c < 128) {
    c = 1;
    while ((c * 40 * concurrency < n) && (c < 128)) c *= 2;
  }
  return c;
}

void oracle_reduce_partition(int64_t n, int threads, int rank, int64_t* begin, int64_t* end) {
  const int64_t chunk_req = oracle_auto_chunk(n, threads);
  const int64_t chunk_min = (n + INT_MAX) / INT_MAX;
  const int64_t chunk = chunk_req > chunk_min ? chunk_req : chunk_min;
  const int64_t num = (n + chunk - 1) / chunk;
  const int64_t part = (num + threads - 1) / threads;
  int64_t first = part * rank * chunk;
  int64_t second = (part * rank + part) * chunk;
  if (second > n) second = n;
  if (first > second) first = second;
  *begin = first;
  *end = second;
}

void oracle_scan_partition(int64_t n, int threads, int rank, int64_t* begin, int64_t* end) {
  const int64_t mask = oracle_auto_chunk(n, threads) - 1;
  const int64_t part = (((n + (threads - 1)) / threads) + mask) & ~mask;
  int64_t b = part * rank, e = b + part;
  if (n < b) b = n;
  if (n < e) e = n;
  *begin = b;
  *end = e;
}

/* ---------------------------------------------------------------- reductions */
#define DEFINE_REDUCE(NAME, T, INIT, STEP, JOIN)                                   \
  T NAME(const T* x, int64_t n, int threads) {                                     \
    T result = INIT;                                                               \
    if (n <= 0) return result;                                                     \
    if (threads < 1) threads = 1;                                                  \
    for (int r = 0; r < threads; ++r) {                                            \
      int64_t b, e;                                                                \
      oracle_reduce_partition(n, threads, r, &b, &e);                              \
      T u = INIT;                                                                  \
      for (int64_t i = b; i < e; ++i) { const T v = x[i]; STEP; }                  \
      if (r == 0) result = u; else { T* dst = &result; const T src = u; JOIN; }    \
    }                                                                              \
    return result;                                                                 \
  }
DEFINE_REDUCE(oracle_reduce_sum_f64, double, 0.0, u += v, *dst += src)
DEFINE_REDUCE(oracle_reduce_sum_f32, float, 0.0f, u += v, *dst += src)
DEFINE_REDUCE(oracle_reduce_sum_i64, int64_t, 0, u = (int64_t)((uint64_t)u + (uint64_t)v), *dst = (int64_t)((uint64_t)*dst + (uint64_t)src))
DEFINE_REDUCE(oracle_reduce_sum_i32, int32_t, 0, u = (int32_t)((uint32_t)u + (uint32_t)v), *dst = (int32_t)((uint32_t)*dst + (uint32_t)src))
DEFINE_REDUCE(oracle_reduce_min_f64, double, DBL_MAX, if (v < u) u = v, if (src < *dst) *dst = src)
DEFINE_REDUCE(oracle_reduce_max_f64, double, -DBL_MAX, if (v > u) u = v, if (src > *dst) *dst = src)
DEFINE_REDUCE(oracle_reduce_min_i64, int64_t, INT64_MAX, if (v < u) u = v, if (src < *dst) *dst = src)
DEFINE_REDUCE(oracle_reduce_max_i64, int64_t, INT64_MIN, if (v > u) u = v, if (src > *dst) *dst = src)
DEFINE_REDUCE(oracle_reduce_min_i32, int32_t, INT32_MAX, if (v < u) u = v, if (src < *dst) *dst = src)
DEFINE_REDUCE(oracle_reduce_max_i32, int32_t, INT32_MIN, if (v > u) u = v, if (src > *dst) *dst = src)

oracle_minmax_f64 oracle_reduce_minmax_f64(const double* x, int64_t n, int threads) {
  oracle_minmax_f64 result = {DBL_MAX, -DBL_MAX};
  if (n <= 0) return result;
  if (threads < 1) threads = 1;
  for (int r = 0; r < threads; ++r) {
    int64_t b, e;
    oracle_reduce_partition(n, threads, r, &b, &e);
    oracle_minmax_f64 u = {DBL_MAX, -DBL_MAX};
    for (int64_t i = b; i < e; ++i) {
      if (x[i] < u.min_val) u.min_val = x[i];
      if (x[i] > u.max_val) u.max_val = x[i];
    }
    if (r == 0) result = u;
    else {
      if (u.min_val < result.min_val) result.min_val = u.min_val;
      if (u.max_val > result.max_val) result.max_val = u.max_val;
    }
  }
  return result;
}

/* MinLoc::join, Kokkos_Parallel_Reduce.hpp:441-449 */
static void minloc_join(oracle_valloc_f64* dest, const oracle_valloc_f64* src) {
  if (src->val < dest->val) *dest = *src;
  else if (src->val == dest->val && dest->loc == IDX_IDENTITY) dest->loc = src->loc;
}
/* MaxLoc::join, :501-509 */
static void maxloc_join(oracle_valloc_f64* dest, const oracle_valloc_f64* src) {
  if (src->val > dest->val) *dest = *src;
  else if (src->val == dest->val && dest->loc == IDX_IDENTITY) dest->loc = src->loc;
}
/* MinMaxLoc::join, :628-644 */
static void minmaxloc_join(oracle_minmaxloc_f64* dest, const oracle_minmaxloc_f64* src) {
  if (src->min_val < dest->min_val) { dest->min_val = src->min_val; dest->min_loc = src->min_loc; }
  else if (dest->min_val == src->min_val && dest->min_loc == IDX_IDENTITY) dest->min_loc = src->min_loc;
  if (src->max_val > dest->max_val) { dest->max_val = src->max_val; dest->max_loc = src->max_loc; }
  else if (dest->max_val == src->max_val && dest->max_loc == IDX_IDENTITY) dest->max_loc = src->max_loc;
}

oracle_valloc_f64 oracle_reduce_minloc_f64(const double* x, int64_t n, int64_t base, int threads) {
  oracle_valloc_f64 result = {DBL_MAX, IDX_IDENTITY};
  if (n <= 0) return result;
  if (threads < 1) threads = 1;
  for (int r = 0; r < threads; ++r) {
    int64_t b, e;
    oracle_reduce_partition(n, threads, r, &b, &e);
    oracle_valloc_f64 u = {DBL_MAX, IDX_IDENTITY};
    for (int64_t i = b; i < e; ++i)
      if (x[i] < u.val) { u.val = x[i]; u.loc = base + i; }
    if (r == 0) result = u; else minloc_join(&result, &u);
  }
  return result;
}
oracle_valloc_f64 oracle_reduce_maxloc_f64(const double* x, int64_t n, int64_t base, int threads) {
  oracle_valloc_f64 result = {-DBL_MAX, IDX_IDENTITY};
  if (n <= 0) return result;
  if (threads < 1) threads = 1;
  for (int r = 0; r < threads; ++r) {
    int64_t b, e;
    oracle_reduce_partition(n, threads, r, &b, &e);
    oracle_valloc_f64 u = {-DBL_MAX, IDX_IDENTITY};
    for (int64_t i = b; i < e; ++i)
      if (x[i] > u.val) { u.val = x[i]; u.loc = base + i; }
    if (r == 0) result = u; else maxloc_join(&result, &u);
  }
  return result;
}
oracle_minmaxloc_f64 oracle_reduce_minmaxloc_f64(const double* x, int64_t n, int64_t base, int threads) {
  oracle_minmaxloc_f64 result = {DBL_MAX, -DBL_MAX, IDX_IDENTITY, IDX_IDENTITY};
  if (n <= 0) return result;
  if (threads < 1) threads = 1;
  for (int r = 0; r < threads; ++r) {
    int64_t b, e;
    oracle_reduce_partition(n, threads, r, &b, &e);
    oracle_minmaxloc_f64 u = {DBL_MAX, -DBL_MAX, IDX_IDENTITY, IDX_IDENTITY};
    for (int64_t i = b; i < e; ++i) {
      if (x[i] < u.min_val) { u.min_val = x[i]; u.min_loc = base + i; }
      if (x[i] > u.max_val) { u.max_val = x[i]; u.max_loc = base + i; }
    }
    if (r == 0) result = u; else minmaxloc_join(&result, &u);
  }
  return result;
}

/* ---------------------------------------------------------------- scans */
#define DEFINE_SCAN(NAME, T, ADD)                                                              \
  T NAME(const T* x, T* y, int64_t n, T seed, int inclusive, int threads) {                    \
    if (n <= 0) return (T)0;                                                                   \
    if (threads < 1) threads = 1;                                                              \
    T base = (T)0; /* exclusive scan of the thread totals, Parallel_Scan.hpp:109-131 */        \
    T total = (T)0; /* ScanWithTotal: the LAST thread's running value after pass 2, :262-264 */  \
    for (int r = 0; r < threads; ++r) {                                                        \
      int64_t b, e;                                                                            \
      oracle_scan_partition(n, threads, r, &b, &e);                                            \
      T sum = (T)0; /* pass 1, final=false */                                                  \
      for (int64_t i = b; i < e; ++i) sum = ADD(sum, x[i]);                                    \
      T upd = base; /* pass 2, final=true, from this thread's base */                          \
      for (int64_t i = b; i < e; ++i) {                                                        \
        const T xi = x[i];                                                                     \
        if (inclusive) { upd = ADD(upd, xi); y[i] = ADD(seed, upd); }                          \
        else { y[i] = ADD(seed, upd); upd = ADD(upd, xi); }                                    \
      }                                                                                        \
      base = ADD(base, sum);                                                                   \
      total = upd;                                                                             \
    }                                                                                          \
    return total;                                                                              \
  }
#define ADD_I64(a, b) ((int64_t)((uint64_t)(a) + (uint64_t)(b)))
#define ADD_I32(a, b) ((int32_t)((uint32_t)(a) + (uint32_t)(b)))
#define ADD_F64(a, b) ((a) + (b))
DEFINE_SCAN(oracle_scan_i64, int64_t, ADD_I64)
DEFINE_SCAN(oracle_scan_i32, int32_t, ADD_I32)
DEFINE_SCAN(oracle_scan_f64, double, ADD_F64)

/* ---------------------------------------------------------------- parallel_for (stream) */
/* functors: benchmarks/stream/stream-kokkos.cpp:55-77; elementwise, so no partition matters */
void oracle_stream_set_f64(double* a, double v, int64_t n) { for (int64_t i = 0; i < n; ++i) a[i] = v; }
void oracle_stream_copy_f64(const double* a, double* b, int64_t n) { for (int64_t i = 0; i < n; ++i) b[i] = a[i]; }
void oracle_stream_scale_f64(double* b, const double* c, double s, int64_t n) { for (int64_t i = 0; i < n; ++i) b[i] = s * c[i]; }
void oracle_stream_add_f64(const double* a, const double* b, double* c, int64_t n) { for (int64_t i = 0; i < n; ++i) c[i] = a[i] + b[i]; }
void oracle_stream_triad_f64(double* a, const double* b, const double* c, double s, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { const double t = s * c[i]; a[i] = b[i] + t; }
}

/* ---------------------------------------------------------------- MDRange stencil + MinMaxLoc */
/* Reducer semantics as above; the combine order of the OpenMP MDRange reduce is tile order
 * (OpenMP/Kokkos_OpenMP_Parallel_Reduce.hpp:217-299), which only matters for ties: the data this is
 * used on plants unique extrema, as the reference's own loc tests do (TestReducers.hpp:1004-1023). */
oracle_minmaxloc_f64 oracle_stencil7_minmaxloc_f64(const double* u, double* v_out, int64_t n0, int64_t n1, int64_t n2,
                                                   double c0, double c1) {
  oracle_minmaxloc_f64 r = {DBL_MAX, -DBL_MAX, IDX_IDENTITY, IDX_IDENTITY};
  const int64_t sj = n0, sk = n0 * n1;
  for (int64_t i = 1; i < n0 - 1; ++i)
    for (int64_t j = 1; j < n1 - 1; ++j)
      for (int64_t k = 1; k < n2 - 1; ++k) {
        const double* c = u + i + j * sj + k * sk;
        double s = c[-1] + c[1];
        s = s + c[-sj];
        s = s + c[sj];
        s = s + c[-sk];
        s = s + c[sk];
        const double a = c0 * c[0];
        const double b = c1 * s;
        const double v = a + b;
        if (v_out) v_out[i + j * sj + k * sk] = v;
        const int64_t loc = (i * n1 + j) * n2 + k;
        if (v < r.min_val) { r.min_val = v; r.min_loc = loc; }
        if (v > r.max_val) { r.max_val = v; r.max_loc = loc; }
      }
  return r;
}

/* ---------------------------------------------------------------- atomics */
/* benchmarks/gups/gups.cpp:83-97: one RMW per index; integer add/xor commute, so a serial loop is the
 * same final table as any interleaving */
void oracle_gups_add_i64(int64_t* t, const int64_t* idx, int64_t m, int64_t d) {
  for (int64_t i = 0; i < m; ++i) t[idx[i]] = (int64_t)((uint64_t)t[idx[i]] + (uint64_t)d);
}
void oracle_gups_xor_i64(int64_t* t, const int64_t* idx, int64_t m, int64_t d) {
  for (int64_t i = 0; i < m; ++i) t[idx[i]] ^= d;
}
void oracle_atomic_add_f64(double* t, const int64_t* idx, const double* v, int64_t m) {
  for (int64_t i = 0; i < m; ++i) t[idx[i]] += v[i];
}

/* ---------------------------------------------------------------- TeamPolicy SpMV */
/* host ThreadVectorRange reduce has vector length 1: left-to-right sum per row
 * (impl/Kokkos_HostThreadTeam.hpp:781-1060) */
void oracle_spmv_crs_f64(int64_t nrows, const int64_t* row_map, const int32_t* col_idx, const double* values,
                         const double* x, double* y) {
  for (int64_t r = 0; r < nrows; ++r) {
    double s = 0.0;
    for (int64_t k = row_map[r]; k < row_map[r + 1]; ++k) { const double p = values[k] * x[col_idx[k]]; s += p; }
    y[r] = s;
  }
}
