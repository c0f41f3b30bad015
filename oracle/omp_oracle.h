/* omp_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into, loaded by, or called from the product).
 *
 * Plain-C restatement ("port") of what the reference's Kokkos::OpenMP backend computes for the hot
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it, and only as
 * the checker.  The association order of a P-thread OpenMP run is reproduced exactly (each simulated
 * thread folds its contiguous block left to right, thread partials are joined in thread order), so
 * floating-point results are bit-identical to oracle/_ref/libkokkos_ref_omp.so run with P threads --
 * tests/test_oracle.py pins that, and pins both against the reference's closed-form known answers.
 */
#ifndef OMP_ORACLE_H
#define OMP_ORACLE_H
#include <stdint.h>

typedef struct { double val; int64_t loc; } oracle_valloc_f64;
typedef struct { double min_val, max_val; int64_t min_loc, max_loc; } oracle_minmaxloc_f64;
typedef struct { double min_val, max_val; } oracle_minmax_f64;

/* partitions (exposed for tests) */
int64_t oracle_auto_chunk(int64_t n, int threads);
void oracle_reduce_partition(int64_t n, int threads, int rank, int64_t* begin, int64_t* end);
void oracle_scan_partition(int64_t n, int threads, int rank, int64_t* begin, int64_t* end);

/* parallel_reduce, RangePolicy(0,n), canonical functors of TestReducers.hpp */
double oracle_reduce_sum_f64(const double* x, int64_t n, int threads);
float oracle_reduce_sum_f32(const float* x, int64_t n, int threads);
int64_t oracle_reduce_sum_i64(const int64_t* x, int64_t n, int threads);
int32_t oracle_reduce_sum_i32(const int32_t* x, int64_t n, int threads);
double oracle_reduce_min_f64(const double* x, int64_t n, int threads);
double oracle_reduce_max_f64(const double* x, int64_t n, int threads);
int64_t oracle_reduce_min_i64(const int64_t* x, int64_t n, int threads);
int64_t oracle_reduce_max_i64(const int64_t* x, int64_t n, int threads);
int32_t oracle_reduce_min_i32(const int32_t* x, int64_t n, int threads);
int32_t oracle_reduce_max_i32(const int32_t* x, int64_t n, int threads);
oracle_minmax_f64 oracle_reduce_minmax_f64(const double* x, int64_t n, int threads);
oracle_valloc_f64 oracle_reduce_minloc_f64(const double* x, int64_t n, int64_t index_base, int threads);
oracle_valloc_f64 oracle_reduce_maxloc_f64(const double* x, int64_t n, int64_t index_base, int threads);
oracle_minmaxloc_f64 oracle_reduce_minmaxloc_f64(const double* x, int64_t n, int64_t index_base, int threads);

/* parallel_scan, RangePolicy(0,n); returns the total; y may alias x */
int64_t oracle_scan_i64(const int64_t* x, int64_t* y, int64_t n, int64_t seed, int inclusive, int threads);
double oracle_scan_f64(const double* x, double* y, int64_t n, double seed, int inclusive, int threads);
int32_t oracle_scan_i32(const int32_t* x, int32_t* y, int64_t n, int32_t seed, int inclusive, int threads);

/* parallel_for: benchmarks/stream kernels */
void oracle_stream_set_f64(double* a, double v, int64_t n);
void oracle_stream_copy_f64(const double* a, double* b, int64_t n);
void oracle_stream_scale_f64(double* b, const double* c, double s, int64_t n);
void oracle_stream_add_f64(const double* a, const double* b, double* c, int64_t n);
void oracle_stream_triad_f64(double* a, const double* b, const double* c, double s, int64_t n);

/* MDRange<Rank<3>> 7-point stencil + MinMaxLoc; u LayoutLeft (i fastest); v_out may be NULL */
oracle_minmaxloc_f64 oracle_stencil7_minmaxloc_f64(const double* u, double* v_out, int64_t n0, int64_t n1, int64_t n2,
                                                   double c0, double c1);

/* atomics */
void oracle_gups_add_i64(int64_t* table, const int64_t* idx, int64_t m, int64_t datum);
void oracle_gups_xor_i64(int64_t* table, const int64_t* idx, int64_t m, int64_t datum);
void oracle_atomic_add_f64(double* table, const int64_t* idx, const double* v, int64_t m);

/* TeamPolicy nested-reduce CRS SpMV (host vector length 1: each row summed left to right) */
void oracle_spmv_crs_f64(int64_t nrows, const int64_t* row_map, const int32_t* col_idx, const double* values,
                         const double* x, double* y);
#endif
