// runtime_internal.h -- the instance record shared by the .cu files of libkokkos_b200.so.
// Not part of the ABI: callers only ever see the opaque b200_instance*.
#ifndef KOKKOS_B200_RUNTIME_INTERNAL_H
#define KOKKOS_B200_RUNTIME_INTERNAL_H

#include <kokkos_b200.h>
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>

constexpr unsigned kResultSlots = 64;
constexpr size_t kSlotTrailer = 64;  // bytes after each result slot; word 0 = completion sequence

struct b200_instance {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  uint32_t id = 0;
  b200_props props;
  std::mutex mutex;  // cold paths only (growth, epoch hand-out)

  // reduction scratch
  void* partials = nullptr;
  size_t partials_bytes = 0;
  unsigned* flags = nullptr;  // [0] = reduction ticket
  void* result_ring = nullptr;      // pinned host
  void* result_ring_dev = nullptr;  // same memory, device address
  size_t slot_bytes = 0;
  std::atomic<unsigned> next_slot{0};
  std::atomic<unsigned long long> result_seq{0};

  // scan scratch
  void* scan_desc = nullptr;
  size_t scan_desc_bytes = 0;
  unsigned long long* tile_counter = nullptr;  // monotonic dynamic tile id source
  uint64_t scan_epoch = 0;
  uint64_t tiles_issued = 0;

  void* scan_status = nullptr;
  size_t scan_status_bytes = 0;
  void* scan_values = nullptr;
  size_t scan_values_bytes = 0;

  // chunk-synchronous scan (impl/ScanChunked.hpp): LL descriptor ring, step-tag counter, pinned error word
  unsigned long long* chunk_desc = nullptr;
  uint32_t chunk_tag = 1;
  unsigned* chunk_err = nullptr;      // pinned host
  unsigned* chunk_err_dev = nullptr;  // same word, device address

  void* functor_spill = nullptr;
  size_t functor_spill_bytes = 0;
  void* team_l1 = nullptr;
  size_t team_l1_bytes = 0;
};

int b200_set_error(int code, const char* where, const char* detail);
void b200_hostpath_release(b200_instance* inst);  // hostpath.cu: drop the instance's staging pipeline (called by b200_finalize)
void b200_spmv_release(b200_instance* inst);      // spmv.cu: forget the matrix metadata remembered for the instance (called by b200_finalize)
int b200_tune(const char* key, int dflt);  // value of a tuning knob, or dflt

#define B200_CHECK_INST(I, where) \
  if (!(I)) return b200_set_error(B200_ENOTINIT, where, nullptr)

#endif
