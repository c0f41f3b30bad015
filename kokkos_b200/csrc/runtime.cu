// runtime.cu -- instance, stream, memory and scratch management behind include/kokkos_b200.h.
//
// Plays the role of CudaInternal (core/src/Cuda/Kokkos_Cuda_Instance.{hpp,cpp}) for the B200
// execution space, redesigned so that the hot path never allocates, memsets or takes a global lock:
//   * all scratch is sized once at instance creation (partials 1 MiB, look-back descriptors for
//     2^22 tiles, a 64-slot pinned+mapped result ring) and only grows, outside the steady state,
//     when a larger request first appears (the reference frees+mallocs on growth as well:
//     Kokkos_Cuda_Instance.cpp:360-379);
//   * tickets self-reset (Collectives.hpp) and look-back descriptors are tagged with a
//     monotonically increasing epoch (scan), so nothing is cleared between launches;
//   * results are written by the device into mapped pinned memory: one stream sync, no memcpy.
#include <kokkos_b200.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "runtime_internal.h"

namespace {
thread_local std::string g_last_error;
std::atomic<uint32_t> g_next_instance_id{1};
std::mutex g_tune_mutex;
std::map<std::string, int>& tune_map() {
  static std::map<std::string, int> m;
  return m;
}
}  // namespace

int b200_set_error(int code, const char* where, const char* detail) {
  if (code == 0) return 0;
  char buf[512];
  if (code > 0) {
    snprintf(buf, sizeof buf, "%s: CUDA error %d (%s): %s%s%s", where ? where : "kokkos_b200", code,
             cudaGetErrorName((cudaError_t)code), cudaGetErrorString((cudaError_t)code), detail ? " -- " : "",
             detail ? detail : "");
  } else {
    static const char* names[] = {"", "invalid argument", "instance not initialised", "unsupported request",
                                  "out of memory", "device is not a compute-capability 10.x (B200) GPU"};
    int k = -code;
    snprintf(buf, sizeof buf, "%s: %s%s%s", where ? where : "kokkos_b200", k < 6 ? names[k] : "error",
             detail ? " -- " : "", detail ? detail : "");
  }
  g_last_error = buf;
  return code;
}

// makes `device` current for a scope and restores the caller's device (multi-GPU processes: every entry that allocates,
// copies or launches acts on the instance's device, never on whatever happens to be current)
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != device) {
      cudaSetDevice(device);
      prev = cur;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define CU_TRY(expr, where)                                            \
  do {                                                                 \
    cudaError_t e__ = (expr);                                          \
    if (e__ != cudaSuccess) return b200_set_error((int)e__, where, #expr); \
  } while (0)

extern "C" {

const char* b200_last_error_string(void) { return g_last_error.c_str(); }
const char* b200_version(void) { return "kokkos_b200 0.1 (sm_100a; mirrors kokkos 4.6.99 hot path)"; }

int b200_report_error(int code, const char* where) { return b200_set_error(code, where, nullptr); }

int b200_device_count(int* count) {
  if (!count) return b200_set_error(B200_EINVAL, "b200_device_count", "count is NULL");
  CU_TRY(cudaGetDeviceCount(count), "b200_device_count");
  return 0;
}

static int instance_setup(int device, cudaStream_t stream, bool owns_stream, b200_instance** out) {
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev), "b200_init");
  if (device < 0 || device >= ndev) return b200_set_error(B200_EINVAL, "b200_init", "device id out of range");
  DeviceGuard guard(device);
  cudaDeviceProp p;
  CU_TRY(cudaGetDeviceProperties(&p, device), "b200_init");
  if (p.major != 10)  // kernels are sm_100a SASS only: anything else cannot run them
    return b200_set_error(B200_EARCH, "b200_init", p.name);

  b200_instance* I = new b200_instance();
  I->device = device;
  I->owns_stream = owns_stream;
  I->stream = stream;
  if (owns_stream) {
    cudaError_t e = cudaStreamCreateWithFlags(&I->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete I; return b200_set_error((int)e, "b200_init", "cudaStreamCreate"); }
  }
  I->id = g_next_instance_id++;
  memset(&I->props, 0, sizeof I->props);
  I->props.device = device;
  I->props.cc_major = p.major;
  I->props.cc_minor = p.minor;
  I->props.sm_count = p.multiProcessorCount;
  I->props.max_threads_per_sm = p.maxThreadsPerMultiProcessor;
  I->props.warp_size = p.warpSize;
  I->props.smem_per_block_optin = p.sharedMemPerBlockOptin;
  I->props.smem_per_sm = p.sharedMemPerMultiprocessor;
  I->props.l2_bytes = (size_t)p.l2CacheSize;
  I->props.total_mem = p.totalGlobalMem;
  I->props.concurrency = p.maxThreadsPerMultiProcessor * p.multiProcessorCount;
  strncpy(I->props.name, p.name, sizeof(I->props.name) - 1);

  // ---- scratch, sized once ----
  auto fail = [&](cudaError_t e, const char* what) {
    b200_finalize(I);
    return b200_set_error(e == cudaErrorMemoryAllocation ? B200_ENOMEM : (int)e, "b200_init", what);
  };
  cudaError_t e;
  I->partials_bytes = 1u << 20;
  if ((e = cudaMalloc(&I->partials, I->partials_bytes)) != cudaSuccess) return fail(e, "partials");
  if ((e = cudaMalloc((void**)&I->flags, 256)) != cudaSuccess) return fail(e, "flags");
  if ((e = cudaMemsetAsync(I->flags, 0, 256, I->stream)) != cudaSuccess) return fail(e, "flags memset");
  I->slot_bytes = 256;  // value bytes; each slot is followed by a 64-byte trailer whose first word is the completion sequence
  if ((e = cudaHostAlloc(&I->result_ring, (size_t)kResultSlots * (I->slot_bytes + kSlotTrailer), cudaHostAllocMapped)) != cudaSuccess)
    return fail(e, "pinned result ring");
  if ((e = cudaHostGetDevicePointer(&I->result_ring_dev, I->result_ring, 0)) != cudaSuccess) return fail(e, "map ring");
  I->scan_desc_bytes = (size_t)32 << 20;
  if ((e = cudaMalloc(&I->scan_desc, I->scan_desc_bytes)) != cudaSuccess) return fail(e, "scan descriptors");
  if ((e = cudaMemsetAsync(I->scan_desc, 0, I->scan_desc_bytes, I->stream)) != cudaSuccess) return fail(e, "scan memset");
  if ((e = cudaMalloc((void**)&I->tile_counter, 256)) != cudaSuccess) return fail(e, "tile counter");
  if ((e = cudaMemsetAsync(I->tile_counter, 0, 256, I->stream)) != cudaSuccess) return fail(e, "counter memset");
  I->scan_epoch = 0;
  I->tiles_issued = 0;
  if ((e = cudaStreamSynchronize(I->stream)) != cudaSuccess) return fail(e, "init sync");
  *out = I;
  return 0;
}

int b200_init(int device, b200_instance** out) {
  if (!out) return b200_set_error(B200_EINVAL, "b200_init", "out is NULL");
  *out = nullptr;
  return instance_setup(device, nullptr, true, out);
}

int b200_instance_create(int device, void* cuda_stream, b200_instance** out) {
  if (!out) return b200_set_error(B200_EINVAL, "b200_instance_create", "out is NULL");
  *out = nullptr;
  // a NULL stream asks for a fresh one (partition_space semantics); a legacy-default-stream
  // instance is requested with the explicit handle cudaStreamLegacy / cudaStreamPerThread.
  if (cuda_stream == nullptr) return instance_setup(device, nullptr, true, out);
  return instance_setup(device, (cudaStream_t)cuda_stream, false, out);
}

int b200_finalize(b200_instance* I) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_finalize", nullptr);
  DeviceGuard guard(I->device);
  b200_hostpath_release(I);
  b200_spmv_release(I);
  if (I->stream || !I->owns_stream) cudaStreamSynchronize(I->stream);
  if (I->partials) cudaFree(I->partials);
  if (I->flags) cudaFree(I->flags);
  if (I->result_ring) cudaFreeHost(I->result_ring);
  if (I->scan_desc) cudaFree(I->scan_desc);
  if (I->tile_counter) cudaFree(I->tile_counter);
  if (I->scan_status) cudaFree(I->scan_status);
  if (I->scan_values) cudaFree(I->scan_values);
  if (I->chunk_desc) cudaFree(I->chunk_desc);
  if (I->chunk_err) cudaFreeHost(I->chunk_err);
  if (I->functor_spill) cudaFree(I->functor_spill);
  if (I->team_l1) cudaFree(I->team_l1);
  if (I->owns_stream && I->stream) cudaStreamDestroy(I->stream);
  delete I;
  return 0;
}

int b200_fence(b200_instance* I, const char* label) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_fence", nullptr);
  cudaError_t e = cudaStreamSynchronize(I->stream);
  if (e != cudaSuccess) return b200_set_error((int)e, label ? label : "b200_fence", "cudaStreamSynchronize");
  return 0;
}

int b200_device_props(b200_instance* I, b200_props* out) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_device_props", nullptr);
  if (!out) return b200_set_error(B200_EINVAL, "b200_device_props", "out is NULL");
  *out = I->props;
  return 0;
}
void* b200_instance_stream(b200_instance* I) { return I ? (void*)I->stream : nullptr; }
uint32_t b200_instance_id(b200_instance* I) { return I ? I->id : 0; }
int b200_instance_sm_count(b200_instance* I) { return I ? I->props.sm_count : 0; }
int b200_instance_device(b200_instance* I) { return I ? I->device : -1; }

// ---------------------------------------------------------------- memory
int b200_malloc(b200_instance* I, size_t bytes, void** out) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_malloc", nullptr);
  if (!out) return b200_set_error(B200_EINVAL, "b200_malloc", "out is NULL");
  *out = nullptr;
  if (bytes == 0) return 0;  // zero-length View: null data pointer, as the reference
  DeviceGuard guard(I->device);
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    char d[96];
    snprintf(d, sizeof d, "cudaMalloc of %zu bytes failed", bytes);
    return b200_set_error(B200_ENOMEM, "b200_malloc", d);
  }
  if (e != cudaSuccess) return b200_set_error((int)e, "b200_malloc", "cudaMalloc");
  return 0;
}
int b200_free(b200_instance* I, void* ptr) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_free", nullptr);
  if (!ptr) return 0;
  DeviceGuard guard(I->device);
  CU_TRY(cudaFree(ptr), "b200_free");
  return 0;
}
int b200_malloc_host_pinned(size_t bytes, void** out) {
  if (!out) return b200_set_error(B200_EINVAL, "b200_malloc_host_pinned", "out is NULL");
  *out = nullptr;
  if (bytes == 0) return 0;
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return b200_set_error(B200_ENOMEM, "b200_malloc_host_pinned", nullptr); }
  if (e != cudaSuccess) return b200_set_error((int)e, "b200_malloc_host_pinned", "cudaHostAlloc");
  return 0;
}
int b200_free_host_pinned(void* ptr) {
  if (!ptr) return 0;
  CU_TRY(cudaFreeHost(ptr), "b200_free_host_pinned");
  return 0;
}
int b200_memset_async(b200_instance* I, void* dst, int byte, size_t bytes) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_memset_async", nullptr);
  if (bytes == 0) return 0;
  if (!dst) return b200_set_error(B200_EINVAL, "b200_memset_async", "dst is NULL");
  DeviceGuard guard(I->device);
  CU_TRY(cudaMemsetAsync(dst, byte, bytes, I->stream), "b200_memset_async");
  return 0;
}
static int copy_async(b200_instance* I, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, const char* where) {
  if (!I) return b200_set_error(B200_ENOTINIT, where, nullptr);
  if (bytes == 0) return 0;
  if (!dst || !src) return b200_set_error(B200_EINVAL, where, "NULL pointer");
  DeviceGuard guard(I->device);
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, kind, I->stream), where);
  return 0;
}
int b200_memcpy_h2d_async(b200_instance* I, void* d, const void* s, size_t n) { return copy_async(I, d, s, n, cudaMemcpyHostToDevice, "b200_memcpy_h2d_async"); }
int b200_memcpy_d2h_async(b200_instance* I, void* d, const void* s, size_t n) { return copy_async(I, d, s, n, cudaMemcpyDeviceToHost, "b200_memcpy_d2h_async"); }
int b200_memcpy_d2d_async(b200_instance* I, void* d, const void* s, size_t n) { return copy_async(I, d, s, n, cudaMemcpyDeviceToDevice, "b200_memcpy_d2d_async"); }

// ---------------------------------------------------------------- scratch
static int grow(b200_instance* I, void** ptr, size_t* have, size_t want, bool zero, const char* what) {
  if (want <= *have) return 0;
  DeviceGuard guard(I->device);
  // growth is a cold path: drain the stream so nothing in flight still uses the old block
  CU_TRY(cudaStreamSynchronize(I->stream), what);
  if (*ptr) CU_TRY(cudaFree(*ptr), what);
  *ptr = nullptr;
  *have = 0;
  size_t sz = 1;
  while (sz < want) sz <<= 1;
  cudaError_t e = cudaMalloc(ptr, sz);
  if (e != cudaSuccess) { cudaGetLastError(); return b200_set_error(B200_ENOMEM, what, "scratch growth failed"); }
  if (zero) {
    CU_TRY(cudaMemsetAsync(*ptr, 0, sz, I->stream), what);
    CU_TRY(cudaStreamSynchronize(I->stream), what);
  }
  *have = sz;
  return 0;
}

int b200_scratch_get(b200_instance* I, int kind, size_t bytes, void** dev_ptr, void** host_ptr) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_scratch_get", nullptr);
  if (!dev_ptr) return b200_set_error(B200_EINVAL, "b200_scratch_get", "dev_ptr is NULL");
  std::lock_guard<std::mutex> lock(I->mutex);
  int rc = 0;
  if (host_ptr) *host_ptr = nullptr;
  switch (kind) {
    case B200_SCRATCH_PARTIALS:
      if ((rc = grow(I, &I->partials, &I->partials_bytes, bytes, false, "b200_scratch_get(partials)"))) return rc;
      *dev_ptr = I->partials;
      return 0;
    case B200_SCRATCH_FLAGS:
      if (bytes > 256) return b200_set_error(B200_EINVAL, "b200_scratch_get(flags)", "at most 256 bytes of flags");
      *dev_ptr = I->flags;
      return 0;
    case B200_SCRATCH_RESULT: {
      if (bytes > I->slot_bytes) return b200_set_error(B200_EUNSUPPORTED, "b200_scratch_get(result)", "value larger than a result slot");
      unsigned k = I->next_slot++ % kResultSlots;
      *dev_ptr = (char*)I->result_ring_dev + (size_t)k * (I->slot_bytes + kSlotTrailer);
      if (host_ptr) *host_ptr = (char*)I->result_ring + (size_t)k * (I->slot_bytes + kSlotTrailer);
      return 0;
    }
    case B200_SCRATCH_FUNCTOR:
      if ((rc = grow(I, &I->functor_spill, &I->functor_spill_bytes, bytes, false, "b200_scratch_get(functor)"))) return rc;
      *dev_ptr = I->functor_spill;
      return 0;
    case B200_SCRATCH_TEAM_L1:
      if ((rc = grow(I, &I->team_l1, &I->team_l1_bytes, bytes, false, "b200_scratch_get(team_l1)"))) return rc;
      *dev_ptr = I->team_l1;
      return 0;
    case B200_SCRATCH_SCAN_DESC:
      if ((rc = grow(I, &I->scan_desc, &I->scan_desc_bytes, bytes, true, "b200_scratch_get(scan_desc)"))) return rc;
      *dev_ptr = I->scan_desc;
      return 0;
    case B200_SCRATCH_SCAN_STATUS:
      if ((rc = grow(I, &I->scan_status, &I->scan_status_bytes, bytes < (4u << 20) ? (4u << 20) : bytes, true, "b200_scratch_get(scan_status)"))) return rc;
      *dev_ptr = I->scan_status;
      return 0;
    case B200_SCRATCH_SCAN_VALUES:
      if ((rc = grow(I, &I->scan_values, &I->scan_values_bytes, bytes < (8u << 20) ? (8u << 20) : bytes, false, "b200_scratch_get(scan_values)"))) return rc;
      *dev_ptr = I->scan_values;
      return 0;
  }
  return b200_set_error(B200_EINVAL, "b200_scratch_get", "unknown scratch kind");
}

int b200_reduce_scratch(b200_instance* I, size_t partial_bytes, size_t value_bytes, int want_slot, void** partials,
                        unsigned** ticket, void** slot_dev, void** slot_host) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_reduce_scratch", nullptr);
  if (!partials || !ticket) return b200_set_error(B200_EINVAL, "b200_reduce_scratch", "NULL out pointer");
  if (partial_bytes > I->partials_bytes) {
    std::lock_guard<std::mutex> lock(I->mutex);
    int rc = grow(I, &I->partials, &I->partials_bytes, partial_bytes, false, "b200_reduce_scratch");
    if (rc) return rc;
  }
  *partials = I->partials;
  *ticket = I->flags;
  if (want_slot) {
    if (!slot_dev || !slot_host) return b200_set_error(B200_EINVAL, "b200_reduce_scratch", "NULL slot pointer");
    if (value_bytes > I->slot_bytes)
      return b200_set_error(B200_EUNSUPPORTED, "b200_reduce_scratch", "reduction value larger than 256 bytes needs a device result");
    unsigned k = I->next_slot++ % kResultSlots;
    *slot_dev = (char*)I->result_ring_dev + (size_t)k * (I->slot_bytes + kSlotTrailer);
    *slot_host = (char*)I->result_ring + (size_t)k * (I->slot_bytes + kSlotTrailer);
  }
  return 0;
}

int b200_result_slot(b200_instance* I, size_t value_bytes, void** slot_dev, void** slot_host, unsigned long long** seq_dev,
                     unsigned long long* seq_value) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_result_slot", nullptr);
  if (!slot_dev || !slot_host || !seq_dev || !seq_value) return b200_set_error(B200_EINVAL, "b200_result_slot", "NULL out pointer");
  if (value_bytes > I->slot_bytes)
    return b200_set_error(B200_EUNSUPPORTED, "b200_result_slot", "reduction value larger than 256 bytes needs a device result");
  const unsigned long long c = I->result_seq++;  // 64-bit, never wraps in practice; value c+1 is unique per hand-out
  const unsigned k = (unsigned)(c % kResultSlots);
  const size_t off = (size_t)k * (I->slot_bytes + kSlotTrailer);
  *slot_dev = (char*)I->result_ring_dev + off;
  *slot_host = (char*)I->result_ring + off;
  *seq_dev = (unsigned long long*)((char*)I->result_ring_dev + off + I->slot_bytes);
  *seq_value = c + 1;
  return 0;
}

int b200_result_wait(b200_instance* I, const void* slot_host, unsigned long long seq_value, const char* label) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_result_wait", nullptr);
  const volatile unsigned long long* seq = (const volatile unsigned long long*)((const char*)slot_host + I->slot_bytes);
  // The kernel's last block stores the value, __threadfence_system(), then the sequence word: poll it instead of paying
  // a driver stream synchronisation.  Every 2048 polls the stream is queried so a failed launch cannot hang the host.
  for (unsigned spin = 1;; ++spin) {
    if (*seq == seq_value) break;
    if ((spin & 2047u) == 0) {
      cudaError_t e = cudaStreamQuery(I->stream);
      if (e == cudaSuccess) {
        if (*seq == seq_value) break;
        return b200_set_error(B200_EINVAL, label ? label : "b200_result_wait", "stream drained but the result slot was not written");
      }
      if (e != cudaErrorNotReady) return b200_set_error((int)e, label ? label : "b200_result_wait", "cudaStreamQuery");
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return 0;
}

int b200_scan_begin(b200_instance* I, uint64_t ntiles, uint64_t* epoch, uint64_t* counter_base,
                    unsigned long long** counter_dev) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_scan_begin", nullptr);
  if (!epoch || !counter_base || !counter_dev) return b200_set_error(B200_EINVAL, "b200_scan_begin", "NULL out pointer");
  std::lock_guard<std::mutex> lock(I->mutex);
  // counter_dev[0] hands out tile ids and is reset to 0 by the last CTA of each launch (counter_dev[1] counts
  // finished CTAs), so a launch that fails on the host cannot desynchronise host and device state.
  (void)ntiles;
  *epoch = ++I->scan_epoch;
  *counter_base = 0;
  *counter_dev = I->tile_counter;
  return 0;
}

int b200_chunk_begin(b200_instance* I, uint64_t nsteps, unsigned* tag_base, unsigned long long** desc_dev, unsigned** err_dev) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_chunk_begin", nullptr);
  if (!tag_base || !desc_dev || !err_dev) return b200_set_error(B200_EINVAL, "b200_chunk_begin", "NULL out pointer");
  std::lock_guard<std::mutex> lock(I->mutex);
  if (!I->chunk_desc) {  // first use: 64 KiB of LL descriptors (zero = tag 0 = never valid) + one pinned error word
    constexpr size_t kBytes = 16 * 256 * 16;
    DeviceGuard guard(I->device);
    CU_TRY(cudaMalloc((void**)&I->chunk_desc, kBytes), "b200_chunk_begin");
    CU_TRY(cudaMemsetAsync(I->chunk_desc, 0, kBytes, I->stream), "b200_chunk_begin");
    CU_TRY(cudaHostAlloc((void**)&I->chunk_err, 64, cudaHostAllocMapped), "b200_chunk_begin");
    I->chunk_err[0] = 0;
    CU_TRY(cudaHostGetDevicePointer((void**)&I->chunk_err_dev, I->chunk_err, 0), "b200_chunk_begin");
  }
  *tag_base = I->chunk_tag;
  I->chunk_tag += (uint32_t)nsteps;  // tags only ever grow (mod 2^32): a ring entry is rewritten every 16 steps, so no stale tag can match
  if (I->chunk_tag == 0) I->chunk_tag = 1;
  *desc_dev = I->chunk_desc;
  *err_dev = I->chunk_err_dev;
  return 0;
}

unsigned b200_chunk_error(b200_instance* I) { return (I && I->chunk_err) ? I->chunk_err[0] : 0u; }

int b200_launch(b200_instance* I, const void* func, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                unsigned bz, size_t smem, void** args) {
  if (!I) return b200_set_error(B200_ENOTINIT, "b200_launch", nullptr);
  if (!func) return b200_set_error(B200_EINVAL, "b200_launch", "func is NULL");
  DeviceGuard guard(I->device);
  CU_TRY(cudaLaunchKernel(func, dim3(gx, gy, gz), dim3(bx, by, bz), args, smem, I->stream), "b200_launch");
  return 0;
}

int b200_occupancy(const void* func, int block_threads, size_t smem, int* blocks_per_sm) {
  if (!func || !blocks_per_sm) return b200_set_error(B200_EINVAL, "b200_occupancy", "NULL pointer");
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, func, block_threads, smem), "b200_occupancy");
  return 0;
}

// ---------------------------------------------------------------- tuning knobs
int b200_tune_set(const char* key, int value) {
  if (!key) return b200_set_error(B200_EINVAL, "b200_tune_set", "key is NULL");
  std::lock_guard<std::mutex> lock(g_tune_mutex);
  tune_map()[key] = value;
  return 0;
}
int b200_tune_get(const char* key, int* value) {
  if (!key || !value) return b200_set_error(B200_EINVAL, "b200_tune_get", "NULL pointer");
  std::lock_guard<std::mutex> lock(g_tune_mutex);
  auto it = tune_map().find(key);
  if (it == tune_map().end()) return b200_set_error(B200_EINVAL, "b200_tune_get", "unknown key");
  *value = it->second;
  return 0;
}

}  // extern "C"

int b200_tune(const char* key, int dflt) {
  std::lock_guard<std::mutex> lock(g_tune_mutex);
  auto it = tune_map().find(key);
  return it == tune_map().end() ? dflt : it->second;
}
