/* kokkos_b200.h -- C ABI of the B200-native execution space for the Kokkos Core hot path.
 *
 * The reference (kokkos/kokkos 4.6.99) has NO C ABI for execution: a backend plugs in by
 * partially specialising Impl::ParallelFor/Reduce/Scan/ScanWithTotal on its execution-space
 * type (core/src/Kokkos_Core_fwd.hpp:292-329) and by providing an instance class, a memory
 * space and atomics.  This header is the thin C layer those specialisations sit on
 * (last row).  Each entry point names the reference interface it
 * replaces.  All pointers are plain host or device addresses; no C++/torch types cross it.
 *
 * Conventions
 *   - every function returns int: 0 = ok, >0 = cudaError_t value, <0 = B200_E* below;
 *     b200_last_error_string() describes the last failure on the calling thread.  The C++
 *     layer (kokkos_b200/include/kb200) turns these into the reference's throw/abort split
 *     (core/src/Cuda/Kokkos_Cuda_Error.hpp:45-68).
 *   - work submitted to one instance is ordered on its CUDA stream
 *     (core/src/Cuda/Kokkos_Cuda_Instance.hpp:368-385); different instances may overlap.
 *   - reductions/scans: `result_host` non-NULL  => blocking, value written on return
 *     (scalar-result semantics, core/src/Kokkos_Parallel_Reduce.hpp:1592-1638);
 *     `result_dev` non-NULL => asynchronous, value written by the device (View result).
 *     Both may be given.  Empty ranges yield the reducer identity (TestReducers.hpp:484-489).
 *   - there is no CPU fallback: with no usable sm_100 device every call fails.
 */
#ifndef KOKKOS_B200_H
#define KOKKOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_EINVAL (-1)      /* bad argument (null pointer, negative length, bad enum)   */
#define B200_ENOTINIT (-2)    /* instance is NULL / finalised                             */
#define B200_EUNSUPPORTED (-3)/* request outside what the kernels implement               */
#define B200_ENOMEM (-4)      /* allocation failure (maps to Kokkos bad_alloc path)        */
#define B200_EARCH (-5)       /* device is not compute capability 10.x                    */

typedef struct b200_instance b200_instance;

/* Value/location pairs: layout-compatible with Kokkos::ValLocScalar<double,int64_t> and
 * Kokkos::MinMaxLocScalar<double,int64_t> (core/src/Kokkos_Parallel_Reduce.hpp:405-409,594-598). */
typedef struct { double val; int64_t loc; } b200_valloc_f64;
typedef struct { double min_val, max_val; int64_t min_loc, max_loc; } b200_minmaxloc_f64;
typedef struct { double min_val, max_val; } b200_minmax_f64;

typedef struct {
  int device, cc_major, cc_minor, sm_count, max_threads_per_sm, warp_size;
  size_t smem_per_block_optin, smem_per_sm, l2_bytes, total_mem;
  int concurrency;              /* max_threads_per_sm * sm_count, as Cuda::concurrency()
                                   (core/src/Cuda/Kokkos_Cuda_Instance.cpp:194-198)        */
  char name[64];
} b200_props;

/* ---- instance / runtime:  replaces Cuda::impl_initialize, CudaInternal::initialize/finalize,
 *      Cuda(stream) ctor, Cuda::fence (core/src/Cuda/Kokkos_Cuda_Instance.cpp:268-329,462-504,
 *      542-665; core/src/Cuda/Kokkos_Cuda.hpp:95-247) ---- */
int b200_init(int device, b200_instance** out);                 /* new instance + own stream */
int b200_instance_create(int device, void* cuda_stream, b200_instance** out); /* on a caller stream */
int b200_finalize(b200_instance* inst);
int b200_fence(b200_instance* inst, const char* label);
int b200_device_props(b200_instance* inst, b200_props* out);
int b200_device_count(int* count);
void* b200_instance_stream(b200_instance* inst);               /* the cudaStream_t           */
uint32_t b200_instance_id(b200_instance* inst);                /* Cuda::impl_instance_id()   */
int b200_instance_device(b200_instance* inst);                 /* Cuda::cuda_device(); -1 for NULL */
const char* b200_last_error_string(void);
const char* b200_version(void);

/* ---- memory: replaces CudaSpace::allocate/deallocate, ZeroMemset<Cuda>, DeepCopy<>
 *      (core/src/Cuda/Kokkos_CudaSpace.cpp:151-260; Kokkos_Cuda_ZeroMemset.hpp:26-33;
 *      Kokkos_CudaSpace.hpp:473-590) ---- */
int b200_malloc(b200_instance* inst, size_t bytes, void** out);
int b200_free(b200_instance* inst, void* ptr);
int b200_malloc_host_pinned(size_t bytes, void** out);         /* CudaHostPinnedSpace        */
int b200_free_host_pinned(void* ptr);
int b200_memset_async(b200_instance* inst, void* dst, int byte, size_t bytes);
int b200_memcpy_h2d_async(b200_instance* inst, void* dst_dev, const void* src_host, size_t bytes);
int b200_memcpy_d2h_async(b200_instance* inst, void* dst_host, const void* src_dev, size_t bytes);
int b200_memcpy_d2d_async(b200_instance* inst, void* dst_dev, const void* src_dev, size_t bytes);

/* ---- scratch + launch: replaces CudaInternal::scratch_space/flags/unified/functor and the
 *      team scratch pool (Kokkos_Cuda_Instance.cpp:333-458), CudaParallelLaunch
 *      (Kokkos_Cuda_KernelLaunch.hpp:317-752) and the occupancy search
 *      (Kokkos_Cuda_BlockSize_Deduction.hpp:28-238).  Used by the header-only template layer. */
typedef enum {
  B200_SCRATCH_PARTIALS = 0, /* device: per-block partial values                            */
  B200_SCRATCH_FLAGS = 1,    /* device: zero-initialised, self-resetting tickets (256 B)    */
  B200_SCRATCH_RESULT = 2,   /* pinned+mapped host slot the last block writes the result to */
  B200_SCRATCH_FUNCTOR = 3,  /* device: closure spill for functors larger than 4 KiB        */
  B200_SCRATCH_TEAM_L1 = 4,  /* device: level-1 team scratch arena                          */
  B200_SCRATCH_SCAN_DESC = 5,  /* device: 16-byte look-back tile descriptors (epoch tagged, never cleared) */
  B200_SCRATCH_SCAN_STATUS = 6,/* device: 8-byte status words of the generic scan (epoch tagged, never cleared) */
  B200_SCRATCH_SCAN_VALUES = 7 /* device: aggregate/inclusive value slots of the generic scan             */
} b200_scratch_kind;
int b200_scratch_get(b200_instance* inst, int kind, size_t bytes, void** dev_ptr, void** host_ptr);
/* start a look-back launch: returns the epoch to tag descriptors with and the device tile-id counter
 * (counter_dev[0] = next tile id, reset to 0 by the last CTA of the launch; counter_dev[1] = finished CTAs);
 * counter_base is always 0 and kept for ABI stability */
int b200_scan_begin(b200_instance* inst, uint64_t ntiles, uint64_t* epoch, uint64_t* counter_base,
                    unsigned long long** counter_dev);
/* start a chunk-synchronous scan launch (kb200/impl/ScanChunked.hpp): first LL step tag of the launch, the instance's
 * descriptor ring (16 rows x 256 CTAs x 16 bytes, zero-initialised once) and a pinned error word (device address) */
int b200_chunk_begin(b200_instance* inst, uint64_t nsteps, unsigned* tag_base, unsigned long long** desc_dev, unsigned** err_dev);
unsigned b200_chunk_error(b200_instance* inst);
/* one call for everything a reduction launch needs: partials (>= partial_bytes), the ticket word,
 * and (if want_result_slot) a fresh pinned+mapped result slot of >= value_bytes from a 64-entry ring */
int b200_reduce_scratch(b200_instance* inst, size_t partial_bytes, size_t value_bytes, int want_result_slot,
                        void** partials_dev, unsigned** ticket_dev, void** slot_dev, void** slot_host);
/* low-latency scalar hand-back: a result slot plus a completion sequence word that the kernel writes (after a system-scope
 * fence) when the value is in place; b200_result_wait polls it instead of synchronising the stream (it still surfaces
 * launch failures through periodic cudaStreamQuery).  Replaces the fence + copy of
 * core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:344-360. */
int b200_result_slot(b200_instance* inst, size_t value_bytes, void** slot_dev, void** slot_host, unsigned long long** seq_dev,
                     unsigned long long* seq_value);
int b200_result_wait(b200_instance* inst, const void* slot_host, unsigned long long seq_value, const char* label);
int b200_instance_sm_count(b200_instance* inst);
/* record a CUDA error code raised in the header-only layer; returns the code (0 stays 0) */
int b200_report_error(int code, const char* where);
int b200_launch(b200_instance* inst, const void* func, unsigned gx, unsigned gy, unsigned gz,
                unsigned bx, unsigned by, unsigned bz, size_t smem, void** args);
int b200_occupancy(const void* func, int block_threads, size_t smem, int* blocks_per_sm);

/* ---- typed fast paths over contiguous View<T*> data.
 *      parallel_reduce over RangePolicy with the built-in reducers: replaces
 *      ParallelReduce<...,RangePolicy,Cuda> + cuda_single_inter_block_reduce_scan
 *      (Kokkos_Cuda_Parallel_Range.hpp:118-388; Kokkos_Cuda_ReduceScan.hpp:188-697) for
 *      Sum/Min/Max/MinLoc/MaxLoc/MinMax/MinMaxLoc (Kokkos_Parallel_Reduce.hpp:33-662).
 *      `loc` values are element indices offset by `index_base` (for range-sharded views). ---- */
int b200_reduce_sum_f64(b200_instance*, const double* x, int64_t n, double* result_host, double* result_dev);
int b200_reduce_sum_f32(b200_instance*, const float* x, int64_t n, float* result_host, float* result_dev);
int b200_reduce_sum_i64(b200_instance*, const int64_t* x, int64_t n, int64_t* result_host, int64_t* result_dev);
int b200_reduce_sum_i32(b200_instance*, const int32_t* x, int64_t n, int32_t* result_host, int32_t* result_dev);
int b200_reduce_min_f64(b200_instance*, const double* x, int64_t n, double* result_host, double* result_dev);
int b200_reduce_max_f64(b200_instance*, const double* x, int64_t n, double* result_host, double* result_dev);
int b200_reduce_min_i64(b200_instance*, const int64_t* x, int64_t n, int64_t* result_host, int64_t* result_dev);
int b200_reduce_max_i64(b200_instance*, const int64_t* x, int64_t n, int64_t* result_host, int64_t* result_dev);
int b200_reduce_min_i32(b200_instance*, const int32_t* x, int64_t n, int32_t* result_host, int32_t* result_dev);
int b200_reduce_max_i32(b200_instance*, const int32_t* x, int64_t n, int32_t* result_host, int32_t* result_dev);
int b200_reduce_minmax_f64(b200_instance*, const double* x, int64_t n, b200_minmax_f64* result_host, b200_minmax_f64* result_dev);
int b200_reduce_minloc_f64(b200_instance*, const double* x, int64_t n, int64_t index_base, b200_valloc_f64* result_host, b200_valloc_f64* result_dev);
int b200_reduce_maxloc_f64(b200_instance*, const double* x, int64_t n, int64_t index_base, b200_valloc_f64* result_host, b200_valloc_f64* result_dev);
int b200_reduce_minmaxloc_f64(b200_instance*, const double* x, int64_t n, int64_t index_base, b200_minmaxloc_f64* result_host, b200_minmaxloc_f64* result_dev);

/* parallel_scan over RangePolicy: replaces ParallelScan / ParallelScanWithTotal<...,RangePolicy,Cuda>
 * (Kokkos_Cuda_Parallel_Range.hpp:390-701,704-1047) for the functor
 *   (i, upd, final) { if(final) y(i)=upd; upd+=x(i); }   (exclusive)   or
 *   (i, upd, final) { upd+=x(i); if(final) y(i)=upd; }   (inclusive)
 * `seed` is added to every output (the exclusive prefix of lower-ranked shards);
 * total = seed-free sum of x.  In-place (y==x) is allowed.  One pass, 16 B/element. */
int b200_scan_excl_i64(b200_instance*, const int64_t* x, int64_t* y, int64_t n, int64_t seed, int64_t* total_host, int64_t* total_dev);
int b200_scan_incl_i64(b200_instance*, const int64_t* x, int64_t* y, int64_t n, int64_t seed, int64_t* total_host, int64_t* total_dev);
int b200_scan_excl_f64(b200_instance*, const double* x, double* y, int64_t n, double seed, double* total_host, double* total_dev);
int b200_scan_incl_f64(b200_instance*, const double* x, double* y, int64_t n, double seed, double* total_host, double* total_dev);
int b200_scan_excl_i32(b200_instance*, const int32_t* x, int32_t* y, int64_t n, int32_t seed, int32_t* total_host, int32_t* total_dev);
/* as b200_scan_excl_i64 but the seed is read from device memory when the kernel runs
 * (lets a distributed scan chain reduce -> all-gather -> scan on one stream with no host sync) */
int b200_scan_excl_i64_seed_dev(b200_instance*, const int64_t* x, int64_t* y, int64_t n, const int64_t* seed_dev, int64_t* total_dev);
/* seed = seeds_dev[0] + ... + seeds_dev[nseeds-1], summed by the kernel: a rank of a range-sharded scan passes the
 * all-gathered shard totals and nseeds = its rank (nseeds == 0 => seed 0), so the NCCL all-gather output feeds the scan
 * directly -- no host round trip and no helper kernel between the collective and the scan */
int b200_scan_excl_i64_seeds_dev(b200_instance*, const int64_t* x, int64_t* y, int64_t n, const int64_t* seeds_dev, int nseeds, int64_t* total_dev);

/* parallel_for over RangePolicy, the benchmarks/stream kernels
 * (benchmarks/stream/stream-kokkos.cpp:55-77): replaces ParallelFor<F,RangePolicy,Cuda>
 * (Kokkos_Cuda_Parallel_Range.hpp:38-116) for these five functors. */
int b200_stream_set_f64(b200_instance*, double* a, double value, int64_t n);
int b200_stream_copy_f64(b200_instance*, const double* a, double* b, int64_t n);          /* b = a       */
int b200_stream_scale_f64(b200_instance*, double* b, const double* c, double s, int64_t n);/* b = s*c     */
int b200_stream_add_f64(b200_instance*, const double* a, const double* b, double* c, int64_t n); /* c = a+b */
int b200_stream_triad_f64(b200_instance*, double* a, const double* b, const double* c, double s, int64_t n); /* a = b+s*c, no FMA contraction */

/* MDRangePolicy<Rank<3>> parallel_reduce with MinMaxLoc (config C4): 7-point stencil
 *   v = c0*u(i,j,k) + c1*(u(i-1,j,k)+u(i+1,j,k)+u(i,j-1,k)+u(i,j+1,k)+u(i,j,k-1)+u(i,j,k+1))
 * over the interior [1,n0-1)x[1,n1-1)x[1,n2-1); u is LayoutLeft (i fastest) as a device View is
 * (core/src/Cuda/Kokkos_Cuda_MDRangePolicy.hpp:25-35); loc = (i*n1+j)*n2+k; sums evaluated in
 * the order written, without FMA contraction.  `v_out` (may be NULL) receives v, same layout.
 * Replaces ParallelReduce<...,MDRangePolicy,Cuda> (Kokkos_Cuda_Parallel_MDRange.hpp:248-497). */
int b200_stencil7_minmaxloc_f64(b200_instance*, const double* u, double* v_out, int64_t n0, int64_t n1, int64_t n2,
                                double c0, double c1, b200_minmaxloc_f64* result_host, b200_minmaxloc_f64* result_dev);

/* Kokkos::atomic_add / atomic_fetch_xor loops of benchmarks/gups (gups.cpp:83-97): one relaxed
 * device-scope RMW per index (desul cuda_cc7_asm_atomic_op.inc_isglobal:5-106 -> red.global). */
int b200_gups_add_i64(b200_instance*, int64_t* table, int64_t table_len, const int64_t* indices, int64_t m, int64_t datum);
int b200_gups_xor_i64(b200_instance*, int64_t* table, int64_t table_len, const int64_t* indices, int64_t m, int64_t datum);
int b200_atomic_add_f64(b200_instance*, double* table, int64_t table_len, const int64_t* indices, const double* values, int64_t m);

/* TeamPolicy + TeamThreadRange + ThreadVectorRange nested-reduce CRS SpMV  y = A x
 * (pattern: example/tutorial/Hierarchical_Parallelism/03_vectorization/vectorization.cpp:51-76;
 * replaces ParallelFor<F,TeamPolicy,Cuda> + CudaTeamMember vector reduce,
 * Kokkos_Cuda_Parallel_Team.hpp:431-587, Kokkos_Cuda_Team.hpp:299-334,721-751). */
int b200_spmv_crs_f64(b200_instance*, int64_t nrows, const int64_t* row_map, const int32_t* col_idx,
                      const double* values, const double* x, double* y);

/* ---- host-buffer forms: deep_copy(device,host) -> pattern -> deep_copy(host,device) as ONE chunked, double-buffered
 *      pipeline (H2D of chunk c+1 and D2H of chunk c-1 overlap the kernel of chunk c; chunks are chained on the
 *      device, no host sync inside).  Replaces the user-level sequence around core/src/Kokkos_CopyViews.hpp:897-1100.
 *      Host buffers should be pinned (b200_malloc_host_pinned).  Blocking. ---- */
int b200_reduce_sum_f64_host(b200_instance*, const double* host_x, int64_t n, double* result);
int b200_scan_excl_i64_host(b200_instance*, const int64_t* host_x, int64_t* host_y, int64_t n, int64_t seed, int64_t* total);

/* ---- one-box communicator (SURVEY.md 8b last row, 8e): one process per GPU, peer-mapped mailboxes over NVLink/NVSwitch
 *      and our own kernels -- no NCCL, no torch.  The reference has no collectives: a Kokkos program holds one Kokkos::Cuda
 *      instance per device (core/src/Cuda/Kokkos_Cuda_Instance.cpp:268-283, core/unit_test/TestMultiGPU.hpp) and combines
 *      the per-device results on the host.  Every call is stream-ordered on the instance's stream and must be made by all
 *      ranks in the same order (like MPI/NCCL collectives).  Payloads are at most 1024 bytes per rank. ---- */
typedef struct b200_comm b200_comm;
int b200_comm_unique_id(char* out, size_t capacity);           /* rank 0 makes it, hands it to the others (>= 40 bytes) */
int b200_comm_init(b200_instance* inst, int rank, int world, const char* unique_id, b200_comm** out); /* collective, blocking */
int b200_comm_finalize(b200_comm* comm);                       /* collective, blocking */
int b200_comm_rank(b200_comm* comm);
int b200_comm_world(b200_comm* comm);
int b200_comm_barrier(b200_comm* comm);                        /* device-side, stream ordered */
int b200_comm_host_barrier(b200_comm* comm);                   /* host-side (processes) */
unsigned b200_comm_error(b200_comm* comm);                     /* non-zero after a device-side time-out: 0xA.. all-gather, 0xD.. scan mailbox */
int b200_allgather_bytes(b200_comm* comm, const void* src_dev, void* dst_dev, size_t bytes_per_rank); /* dst: world * bytes, rank order */
/* in-place folds of `count` values per rank, evaluated in RANK ORDER on every rank (bitwise identical everywhere) */
int b200_allreduce_sum_f64(b200_comm* comm, double* buf_dev, int count);
int b200_allreduce_min_f64(b200_comm* comm, double* buf_dev, int count);
int b200_allreduce_max_f64(b200_comm* comm, double* buf_dev, int count);
int b200_allreduce_sum_i64(b200_comm* comm, int64_t* buf_dev, int count);
int b200_allreduce_min_i64(b200_comm* comm, int64_t* buf_dev, int count);
int b200_allreduce_max_i64(b200_comm* comm, int64_t* buf_dev, int count);
/* MinLoc / MaxLoc / MinMaxLoc joins (core/src/Kokkos_Parallel_Reduce.hpp:441-449,501-509,628-644); equal extrema keep the
 * lowest location, which is what the OpenMP reference yields with its thread-ordered joins */
int b200_allreduce_minloc_f64(b200_comm* comm, b200_valloc_f64* buf_dev);
int b200_allreduce_maxloc_f64(b200_comm* comm, b200_valloc_f64* buf_dev);
int b200_allreduce_minmaxloc_f64(b200_comm* comm, b200_minmaxloc_f64* buf_dev);
/* Block-cyclic distributed parallel_scan, ONE fused kernel per GPU, 16 B/element at any world size
 * (kb200/impl/ScanChunked.hpp): global block c (block_elems elements) lives on rank c % world as local block c / world;
 * x/y point at the rank's local blocks stored back to back.  b200_comm_cyclic_layout reports the block size, this rank's
 * local length and the number of lock-step rounds for a global length.  total = sum over ALL ranks' elements.
 * Replaces ParallelScanWithTotal<...,Cuda> (Kokkos_Cuda_Parallel_Range.hpp:704-1047) + the user's cross-device fix-up. */
int b200_comm_cyclic_layout(b200_comm* comm, int elem_bytes, int64_t n_global, int64_t* block_elems, int64_t* n_local, int64_t* nsteps);
int b200_comm_scan_excl_i64(b200_comm* comm, const int64_t* x_local, int64_t* y_local, int64_t n_global, int64_t* total_host, int64_t* total_dev);
int b200_comm_scan_incl_i64(b200_comm* comm, const int64_t* x_local, int64_t* y_local, int64_t n_global, int64_t* total_host, int64_t* total_dev);
int b200_comm_scan_excl_f64(b200_comm* comm, const double* x_local, double* y_local, int64_t n_global, double* total_host, double* total_dev);

/* ---- tuning knobs (benchmark harness only; defaults are the shipped configuration) ---- */
int b200_tune_set(const char* key, int value);
int b200_tune_get(const char* key, int* value);

#ifdef __cplusplus
}
#endif
#endif /* KOKKOS_B200_H */
