// hostpath.cu -- host-buffer entry points of the C ABI: what a caller holding HOST memory gets when it asks the
// B200 execution space for a reduction or a scan.  In the reference this is the user-level sequence
//   create_mirror_view / deep_copy(device, host)  ->  parallel_reduce | parallel_scan  ->  deep_copy(host, device)
// (core/src/Kokkos_CopyViews.hpp:897-1100 + the pattern launch); here the three stages are one chunked,
// double-buffered pipeline so that the PCIe copies in both directions overlap the kernels:
//   copy-in stream : H2D chunk c+1            | main stream : kernel(chunk c)        | copy-out stream : D2H chunk c-1
// Chunks are chained on the device (the scan seed of chunk c+1 is produced by chunk c; reduction partials are
// folded in chunk order), so no host synchronisation happens inside the pipeline.
// Host buffers should be pinned (b200_malloc_host_pinned); pageable memory works but copies are staged by the driver.
#include <kokkos_b200.h>
#include "runtime_internal.h"

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

namespace {

constexpr int64_t kChunkBytes = 64ll << 20;  // per staging buffer: small enough that the un-overlapped first H2D / last D2H are ~1 ms
constexpr int kDepth = 2;

struct HostPipe {
  cudaStream_t in = nullptr, out = nullptr;
  void* din[kDepth] = {nullptr, nullptr};
  void* dout[kDepth] = {nullptr, nullptr};
  cudaEvent_t loaded[kDepth], computed[kDepth], drained[kDepth], consumed[kDepth];
  void* carry = nullptr;  // device scalars: [0..127] running seeds / partials
  bool ready = false;
};

__global__ void carry_add_i64(const long long* seed_in, const long long* total, long long* seed_out) {
  *seed_out = *seed_in + *total;
}

std::mutex g_pipe_mutex;
std::map<b200_instance*, HostPipe*>& pipe_map() {
  static std::map<b200_instance*, HostPipe*> pipes;
  return pipes;
}

int pipe_get(b200_instance* I, HostPipe** out) {
  std::lock_guard<std::mutex> lock(g_pipe_mutex);
  auto& pipes = pipe_map();
  auto it = pipes.find(I);
  if (it != pipes.end()) { *out = it->second; return 0; }
  HostPipe* P = new HostPipe();
  cudaError_t e = cudaSetDevice(I->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&P->in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&P->out, cudaStreamNonBlocking);
  for (int k = 0; k < kDepth && e == cudaSuccess; ++k) {
    e = cudaMalloc(&P->din[k], kChunkBytes);
    if (e == cudaSuccess) e = cudaMalloc(&P->dout[k], kChunkBytes);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->loaded[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->computed[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->drained[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->consumed[k], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaMalloc(&P->carry, 4096 * 8);
  if (e != cudaSuccess) { delete P; return b200_set_error(e == cudaErrorMemoryAllocation ? B200_ENOMEM : (int)e, "b200 host pipeline", "setup"); }
  P->ready = true;
  pipes[I] = P;
  *out = P;
  return 0;
}

}  // namespace

// b200_finalize: the staging pipeline dies with its instance (a later instance at the same address must not inherit it)
void b200_hostpath_release(b200_instance* I) {
  std::lock_guard<std::mutex> lock(g_pipe_mutex);
  auto& pipes = pipe_map();
  auto it = pipes.find(I);
  if (it == pipes.end()) return;
  HostPipe* P = it->second;
  pipes.erase(it);
  if (P->in) cudaStreamSynchronize(P->in);
  if (P->out) cudaStreamSynchronize(P->out);
  for (int k = 0; k < kDepth; ++k) {
    if (P->din[k]) cudaFree(P->din[k]);
    if (P->dout[k]) cudaFree(P->dout[k]);
    if (P->ready) { cudaEventDestroy(P->loaded[k]); cudaEventDestroy(P->computed[k]); cudaEventDestroy(P->drained[k]); cudaEventDestroy(P->consumed[k]); }
  }
  if (P->carry) cudaFree(P->carry);
  if (P->in) cudaStreamDestroy(P->in);
  if (P->out) cudaStreamDestroy(P->out);
  delete P;
}

namespace {
#define CU(expr)                                                                \
  do {                                                                          \
    cudaError_t e__ = (expr);                                                   \
    if (e__ != cudaSuccess) return b200_set_error((int)e__, where, #expr);      \
  } while (0)
}  // namespace

extern "C" {

int b200_reduce_sum_f64_host(b200_instance* I, const double* host_x, int64_t n, double* result) {
  const char* where = "b200_reduce_sum_f64_host";
  B200_CHECK_INST(I, where);
  if (n < 0 || !result || (n > 0 && !host_x)) return b200_set_error(B200_EINVAL, where, "bad argument");
  HostPipe* P = nullptr;
  int rc = pipe_get(I, &P);
  if (rc) return rc;
  const int64_t per = kChunkBytes / 8;
  const int64_t nchunks = (n + per - 1) / per;
  if (nchunks > 4096) return b200_set_error(B200_EUNSUPPORTED, where, "more than 4096 chunks");
  double* partials = reinterpret_cast<double*>(P->carry);
  CU(cudaEventRecord(P->consumed[0], I->stream));  // order after earlier work on the instance
  CU(cudaStreamWaitEvent(P->in, P->consumed[0], 0));
  for (int64_t c = 0; c < nchunks; ++c) {
    const int b = (int)(c % kDepth);
    const int64_t off = c * per, cnt = std::min(per, n - off);
    if (c >= kDepth) CU(cudaStreamWaitEvent(P->in, P->computed[b], 0));  // buffer free once its kernel ran
    CU(cudaMemcpyAsync(P->din[b], host_x + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, P->in));
    CU(cudaEventRecord(P->loaded[b], P->in));
    CU(cudaStreamWaitEvent(I->stream, P->loaded[b], 0));
    if ((rc = b200_reduce_sum_f64(I, (const double*)P->din[b], cnt, nullptr, partials + c))) return rc;
    CU(cudaEventRecord(P->computed[b], I->stream));
  }
  std::vector<double> h((size_t)std::max<int64_t>(nchunks, 1), 0.0);
  if (nchunks) CU(cudaMemcpyAsync(h.data(), partials, (size_t)nchunks * 8, cudaMemcpyDeviceToHost, I->stream));
  CU(cudaStreamSynchronize(I->stream));
  double s = 0.0;
  for (int64_t c = 0; c < nchunks; ++c) s += h[(size_t)c];  // chunk order: deterministic
  *result = s;
  return 0;
}

int b200_scan_excl_i64_host(b200_instance* I, const int64_t* host_x, int64_t* host_y, int64_t n, int64_t seed, int64_t* total) {
  const char* where = "b200_scan_excl_i64_host";
  B200_CHECK_INST(I, where);
  if (n < 0 || (n > 0 && (!host_x || !host_y))) return b200_set_error(B200_EINVAL, where, "bad argument");
  HostPipe* P = nullptr;
  int rc = pipe_get(I, &P);
  if (rc) return rc;
  const int64_t per = kChunkBytes / 8;
  const int64_t nchunks = (n + per - 1) / per;
  if (nchunks > 2000) return b200_set_error(B200_EUNSUPPORTED, where, "more than 2000 chunks");
  long long* seeds = reinterpret_cast<long long*>(P->carry);          // seeds[c] = seed of chunk c
  long long* totals = seeds + 2048;                                    // totals[c] = seed-free sum of chunk c
  const long long seed0 = seed;
  CU(cudaMemcpyAsync(seeds, &seed0, 8, cudaMemcpyHostToDevice, I->stream));
  CU(cudaEventRecord(P->consumed[0], I->stream));
  CU(cudaStreamWaitEvent(P->in, P->consumed[0], 0));
  CU(cudaStreamWaitEvent(P->out, P->consumed[0], 0));
  for (int64_t c = 0; c < nchunks; ++c) {
    const int b = (int)(c % kDepth);
    const int64_t off = c * per, cnt = std::min(per, n - off);
    if (c >= kDepth) CU(cudaStreamWaitEvent(P->in, P->computed[b], 0));     // input buffer consumed by its kernel
    CU(cudaMemcpyAsync(P->din[b], host_x + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, P->in));
    CU(cudaEventRecord(P->loaded[b], P->in));
    CU(cudaStreamWaitEvent(I->stream, P->loaded[b], 0));
    if (c >= kDepth) CU(cudaStreamWaitEvent(I->stream, P->drained[b], 0));  // output buffer copied out
    if ((rc = b200_scan_excl_i64_seed_dev(I, (const int64_t*)P->din[b], (int64_t*)P->dout[b], cnt, (const int64_t*)(seeds + c),
                                          (int64_t*)(totals + c))))
      return rc;
    carry_add_i64<<<1, 1, 0, I->stream>>>(seeds + c, totals + c, seeds + c + 1);
    CU(cudaGetLastError());
    CU(cudaEventRecord(P->computed[b], I->stream));
    CU(cudaStreamWaitEvent(P->out, P->computed[b], 0));
    CU(cudaMemcpyAsync(host_y + off, P->dout[b], (size_t)cnt * 8, cudaMemcpyDeviceToHost, P->out));
    CU(cudaEventRecord(P->drained[b], P->out));
  }
  long long last = seed0;
  CU(cudaMemcpyAsync(&last, seeds + nchunks, 8, cudaMemcpyDeviceToHost, I->stream));
  CU(cudaStreamSynchronize(I->stream));
  CU(cudaStreamSynchronize(P->out));
  if (total) *total = (int64_t)(last - seed0);
  return 0;
}

}  // extern "C"
