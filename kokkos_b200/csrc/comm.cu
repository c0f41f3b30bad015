// comm.cu -- one-box communicator of the B200 execution space: one process per GPU, peer-mapped mailboxes over
// NVLink/NVSwitch, our own kernels (no NCCL, no torch).  SURVEY.md section 8(b) last row / 8(e):
//   b200_comm_init / b200_comm_finalize      bootstrap through a POSIX shared-memory segment named by a unique id
//                                            (the role ncclGetUniqueId + ncclCommInitRank play), CUDA IPC handles
//   b200_allgather_bytes                     few-byte all-gather, LL words (4 B payload + 4 B sequence tag per 8-byte
//                                            word) stored straight into every peer's mailbox; stream ordered, no host sync
//   b200_allreduce_{sum,min,max}_{f64,i64}   all-gather + fold in RANK ORDER on every rank: bitwise identical on all
//                                            ranks and run to run (the reference combines per-device results on the host
//                                            in a fixed order too: core/unit_test/TestMultiGPU.hpp)
//   b200_allreduce_{minloc,maxloc,minmaxloc}_f64   rank-ordered join, equal extrema keep the LOWEST location
//                                            (core/src/Kokkos_Parallel_Reduce.hpp:441-449,501-509,628-644 under ordered joins)
//   b200_comm_scan_{excl,incl}_{i64,f64}     the fused block-cyclic scan (kb200/impl/ScanChunked.hpp)
// The reference has no collectives of its own (one Kokkos::Cuda instance per device, results combined by the user).
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/impl/ScanChunked.hpp>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>

using namespace kb200;
using namespace kb200::Impl;

namespace {

constexpr int kMaxWorld = kChunkMaxWorld;
constexpr int kAgRing = 4;              // all-gather ring slots (a rank can be at most one call ahead of a peer)
constexpr int kAgMaxBytes = 1024;       // payload bytes per rank and call
constexpr int kAgWords = kAgMaxBytes / 4;
// mailbox layout on every rank (one cudaMalloc, zero = tag 0 = never valid)
constexpr size_t kOffAg = 0;                                                           // [kAgRing][kMaxWorld][kAgWords] u64
constexpr size_t kAgBytes = (size_t)kAgRing * kMaxWorld * kAgWords * 8;
constexpr size_t kOffScan = kOffAg + kAgBytes;                                         // [kChunkRing][kMaxWorld][2] u64 (lock-step kernel)
constexpr size_t kOffRound = kOffScan + kChunkMboxBytes;                               // [kRoundRing][kMaxWorld][2] u64 (rounds kernel)
constexpr size_t kRoundMboxBytes = (size_t)kRoundRing * kRoundMaxWorld * 16;
constexpr size_t kOffRdesc = kOffRound + kRoundMboxBytes;                              // [kRoundRing][4] u64, local only
constexpr size_t kOffRacc = kOffRdesc + (size_t)kRoundRing * 32;                       // [kRoundRing][2] u64, local only
constexpr size_t kMailboxBytes = ((kOffRacc + (size_t)kRoundRing * 16 + 4095) / 4096) * 4096;
static_assert(kRoundMaxWorld == kChunkMaxWorld, "one world limit");

struct ShmSlot {
  cudaIpcMemHandle_t handle;
  int device, sm_count;
  long pid;
};
struct ShmHeader {
  std::atomic<unsigned> magic;
  std::atomic<int> arrived[4];
  std::atomic<int> host_barrier[2];
  int world;
  ShmSlot slot[kMaxWorld];
};
constexpr unsigned kMagic = 0xB2000C01u;

}  // namespace

struct b200_comm {
  b200_instance* inst = nullptr;
  int rank = 0, world = 1;
  int grid = 0;  // CTAs of the fused scan = min SM count over the ranks
  unsigned char* mailbox = nullptr;  // this rank's (device)
  unsigned char* peer[kMaxWorld] = {};  // every rank's mailbox as mapped here (peer[rank] == mailbox)
  unsigned ag_seq = 0;     // all-gather calls so far (identical on all ranks)
  unsigned scan_tag = 1;   // next step tag of the lock-step scan (identical on all ranks)
  unsigned round_tag = 1;  // next round tag of the rounds scan (identical on all ranks)
  unsigned* err = nullptr;      // pinned
  unsigned* err_dev = nullptr;
  ShmHeader* shm = nullptr;
  char shm_name[80] = {};
  int host_barrier_gen = 0;
  void* scratch = nullptr;  // device, kMaxWorld * kAgMaxBytes
};

namespace {

#define CU_TRY(expr, where)                                                \
  do {                                                                     \
    cudaError_t e__ = (expr);                                              \
    if (e__ != cudaSuccess) return b200_set_error((int)e__, where, #expr); \
  } while (0)

double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// wait until an shm counter reaches `target` (every rank increments it once)
bool shm_wait(std::atomic<int>& c, int target, double timeout_s) {
  const double t0 = now_s();
  while (c.load(std::memory_order_acquire) < target) {
    if (now_s() - t0 > timeout_s) return false;
    usleep(200);
  }
  return true;
}

enum FoldOp { kGather = 0, kSumF64, kSumI64, kMinF64, kMaxF64, kMinI64, kMaxI64, kMinLoc, kMaxLoc, kMinMaxLoc };

struct AgParams {
  unsigned long long* peer_ag[kMaxWorld];  // base of the all-gather area of every rank
  int rank, world;
  unsigned seq;
  int nwords;  // 4-byte words per rank
  const unsigned* src;
  unsigned* dst;  // kGather: world * nwords words; folds: nwords words (may alias src)
  int op;
  unsigned* err;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned ag_wait_word(const unsigned long long* p, unsigned seq, unsigned long long timeout_ns, unsigned* err, unsigned code) {
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if ((unsigned)(w >> 32) == seq) return (unsigned)w;
    if ((spin & 1023u) == 1023u) {
      const unsigned long long t = ll::now_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > timeout_ns) ll::give_up(err, code);
    }
  }
}

// one CTA.  Phase 1: my payload -> every rank's slot [ring][my rank].  Phase 2: wait for everyone's payload in MY slot row,
// land it in shared memory.  Phase 3: write the gathered bytes or fold them in rank order.
__global__ void __launch_bounds__(256) ll_allgather_kernel(const AgParams p) {
  __shared__ unsigned s_data[kMaxWorld * kAgWords];
  const int ring = (int)(p.seq % kAgRing);
  const size_t row = (size_t)ring * kMaxWorld * kAgWords;
  for (int i = threadIdx.x; i < p.world * p.nwords; i += blockDim.x) {
    const int q = i / p.nwords, w = i % p.nwords;
    const unsigned long long word = ((unsigned long long)p.seq << 32) | p.src[w];
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.peer_ag[q] + row + (size_t)p.rank * kAgWords + w), "l"(word) : "memory");
  }
  for (int i = threadIdx.x; i < p.world * p.nwords; i += blockDim.x) {
    const int q = i / p.nwords, w = i % p.nwords;
    s_data[q * kAgWords + w] = ag_wait_word(p.peer_ag[p.rank] + row + (size_t)q * kAgWords + w, p.seq, p.timeout_ns, p.err, 0xA0000000u | (unsigned)q);
  }
  __syncthreads();
  if (p.op == kGather) {
    for (int i = threadIdx.x; i < p.world * p.nwords; i += blockDim.x) p.dst[i] = s_data[(i / p.nwords) * kAgWords + (i % p.nwords)];
    return;
  }
  auto f64_at = [&](int q, int e) { double v; memcpy(&v, &s_data[q * kAgWords + 2 * e], 8); return v; };
  auto i64_at = [&](int q, int e) { long long v; memcpy(&v, &s_data[q * kAgWords + 2 * e], 8); return v; };
  if (p.op >= kSumF64 && p.op <= kMaxI64) {
    const int count = p.nwords / 2;
    for (int e = threadIdx.x; e < count; e += blockDim.x) {
      if (p.op == kSumF64 || p.op == kMinF64 || p.op == kMaxF64) {
        double a = f64_at(0, e);
        for (int q = 1; q < p.world; ++q) {
          const double v = f64_at(q, e);
          a = p.op == kSumF64 ? __dadd_rn(a, v) : (p.op == kMinF64 ? (v < a ? v : a) : (v > a ? v : a));
        }
        memcpy(&p.dst[2 * e], &a, 8);
      } else {
        long long a = i64_at(0, e);
        for (int q = 1; q < p.world; ++q) {
          const long long v = i64_at(q, e);
          a = p.op == kSumI64 ? a + v : (p.op == kMinI64 ? (v < a ? v : a) : (v > a ? v : a));
        }
        memcpy(&p.dst[2 * e], &a, 8);
      }
    }
    return;
  }
  if (threadIdx.x == 0) {
    // loc reducers: ranks joined in order; a tie keeps the lower location (ranks hold ascending index ranges for a
    // range-sharded View, but a k-slab sharded MDRange does not order locations by rank: compare explicitly)
    if (p.op == kMinLoc || p.op == kMaxLoc) {
      double bv = f64_at(0, 0);
      long long bl = i64_at(0, 1);
      for (int q = 1; q < p.world; ++q) {
        const double v = f64_at(q, 0);
        const long long l = i64_at(q, 1);
        const bool better = p.op == kMinLoc ? (v < bv) : (v > bv);
        if (better || (v == bv && l < bl)) { bv = v; bl = l; }
      }
      memcpy(&p.dst[0], &bv, 8);
      memcpy(&p.dst[2], &bl, 8);
    } else {  // {min_val, max_val, min_loc, max_loc}
      double mn = f64_at(0, 0), mx = f64_at(0, 1);
      long long mnl = i64_at(0, 2), mxl = i64_at(0, 3);
      for (int q = 1; q < p.world; ++q) {
        const double v0 = f64_at(q, 0), v1 = f64_at(q, 1);
        const long long l0 = i64_at(q, 2), l1 = i64_at(q, 3);
        if (v0 < mn || (v0 == mn && l0 < mnl)) { mn = v0; mnl = l0; }
        if (v1 > mx || (v1 == mx && l1 < mxl)) { mx = v1; mxl = l1; }
      }
      memcpy(&p.dst[0], &mn, 8); memcpy(&p.dst[2], &mx, 8); memcpy(&p.dst[4], &mnl, 8); memcpy(&p.dst[6], &mxl, 8);
    }
  }
}

int comm_collective(b200_comm* C, const char* where, const void* src, void* dst, size_t bytes, int op) {
  if (!C) return b200_set_error(B200_ENOTINIT, where, "communicator is NULL");
  if (bytes == 0) return 0;
  if (!src || !dst) return b200_set_error(B200_EINVAL, where, "NULL buffer");
  if (bytes > (size_t)kAgMaxBytes || bytes % 4) return b200_set_error(B200_EUNSUPPORTED, where, "payload must be a multiple of 4 and at most 1024 bytes per rank");
  CU_TRY(cudaSetDevice(C->inst->device), where);
  if (C->world == 1) {  // nothing to exchange: gather == copy, a fold of one value == that value
    if (dst != src) CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, C->inst->stream), where);
    return 0;
  }
  AgParams p;
  memset(&p, 0, sizeof p);
  for (int q = 0; q < C->world; ++q) p.peer_ag[q] = reinterpret_cast<unsigned long long*>(C->peer[q] + kOffAg);
  p.rank = C->rank; p.world = C->world;
  p.seq = ++C->ag_seq;
  if (p.seq == 0) p.seq = ++C->ag_seq;
  p.nwords = (int)(bytes / 4);
  p.src = static_cast<const unsigned*>(src);
  p.dst = static_cast<unsigned*>(dst);
  p.op = op;
  p.err = C->err_dev;
  p.timeout_ns = 30ull * 1000000000ull;
  ll_allgather_kernel<<<1, 256, 0, C->inst->stream>>>(p);
  return b200_report_error((int)cudaGetLastError(), where);
}

// tiles per round of the rounds kernel: a round (= the block of the block-cyclic distribution) is tpr tiles of 18 KiB
int64_t round_tpr() { return b200_tune("comm.tpr", 128); }  // 128 x 18 KiB = 2.25 MiB per round and rank (profiles/r02_cyclic_scan_probe.log)
template <class T>
int64_t round_block_elems() {
  int64_t tile = ContigScanLaunch<T, 128, 9, 4, 1, false, 2>::TILE;
#ifdef B200_SWEEP
  const int tiles[] = {128 * 9, 128 * 7, 128 * 5, 64 * 9, 256 * 9, 128 * 9};
  const int c = b200_tune("comm.cfg", 0);
  if (c >= 0 && c <= 5) tile = (int64_t)tiles[c] * 16 / (int64_t)sizeof(T);
#endif
  return round_tpr() * tile;
}

template <class T, bool INCL>
int comm_scan(b200_comm* C, const char* where, const T* x, T* y, int64_t n_global, T* total_host, T* total_dev) {
  if (!C) return b200_set_error(B200_ENOTINIT, where, "communicator is NULL");
  if (n_global < 0) return b200_set_error(B200_EINVAL, where, "negative length");
  CU_TRY(cudaSetDevice(C->inst->device), where);
  const bool lockstep = b200_tune("comm.scan_algo", 0) == 1;  // 1: the lock-step kernel (ScanChunked.hpp), kept for comparison
  using LS = ChunkScanLaunch<T, 384, 9, 4, 4, INCL>;
  using LR = ContigScanLaunch<T, 128, 9, 4, 1, INCL, 2>;
  if (lockstep && LS::max_grid(C->inst->device, C->grid) <= 0) return b200_set_error(B200_EUNSUPPORTED, where, "the lock-step scan kernel does not fit this device");
  const int64_t block = lockstep ? LS::block_elems(C->grid) : round_block_elems<T>();
  const int64_t nblocks = (n_global + block - 1) / block;
  const int64_t nsteps = (nblocks + C->world - 1) / C->world;
  // local length: full blocks owned by this rank + the (possibly short) last global block if it is ours
  int64_t n_local = 0;
  for (int64_t c = C->rank; c < nblocks; c += C->world) n_local += (c == nblocks - 1) ? (n_global - c * block) : block;
  if (n_local > 0 && (!x || !y)) return b200_set_error(B200_EINVAL, where, "x or y is NULL");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16) return b200_set_error(B200_EUNSUPPORTED, where, "local blocks must be 16-byte aligned");
  if (lockstep) {
    ChunkPeers peers;
    peers.rank = C->rank; peers.world = C->world;
    peers.mbox = reinterpret_cast<unsigned long long*>(C->mailbox + kOffScan);
    for (int q = 0; q < C->world; ++q) peers.peer_mbox[q] = reinterpret_cast<unsigned long long*>(C->peer[q] + kOffScan);
    peers.mbox_tag_base = C->scan_tag;
    C->scan_tag += (unsigned)nsteps;
    if (C->scan_tag == 0) C->scan_tag = 1;
    return LS::run(C->inst, &peers, C->grid, nsteps, x, y, n_local, T(0), nullptr, 0, total_host, total_dev);
  }
  RoundPeers peers;
  peers.rank = C->rank; peers.world = C->world;
  peers.rdesc = reinterpret_cast<unsigned long long*>(C->mailbox + kOffRdesc);
  peers.racc = reinterpret_cast<unsigned long long*>(C->mailbox + kOffRacc);
  peers.mbox = reinterpret_cast<unsigned long long*>(C->mailbox + kOffRound);
  for (int q = 0; q < C->world; ++q) peers.peer_mbox[q] = reinterpret_cast<unsigned long long*>(C->peer[q] + kOffRound);
  peers.rtag_base = C->round_tag;
  peers.err = C->err_dev;
  C->round_tag += (unsigned)nsteps;
  if (C->round_tag == 0) C->round_tag = 1;
#ifdef B200_SWEEP
  switch (b200_tune("comm.cfg", 0)) {  // tile shape experiments (same tile BYTES per round only if tpr is scaled by the caller)
    case 1: return ContigScanLaunch<T, 128, 7, 5, 1, INCL, 2>::run_rounds(C->inst, peers, round_tpr(), nsteps, x, y, n_local, total_host, total_dev);
    case 2: return ContigScanLaunch<T, 128, 5, 7, 1, INCL, 2>::run_rounds(C->inst, peers, round_tpr(), nsteps, x, y, n_local, total_host, total_dev);
    case 3: return ContigScanLaunch<T, 64, 9, 8, 1, INCL, 2>::run_rounds(C->inst, peers, round_tpr(), nsteps, x, y, n_local, total_host, total_dev);
    case 4: return ContigScanLaunch<T, 256, 9, 2, 1, INCL, 2>::run_rounds(C->inst, peers, round_tpr(), nsteps, x, y, n_local, total_host, total_dev);
    case 5: return ContigScanLaunch<T, 128, 9, 4, 2, INCL, 2>::run_rounds(C->inst, peers, round_tpr(), nsteps, x, y, n_local, total_host, total_dev);
    default: break;
  }
#endif
  return LR::run_rounds(C->inst, peers, round_tpr(), nsteps, x, y, n_local, total_host, total_dev, b200_tune("scan.pfd", -1));
}

}  // namespace

extern "C" {

int b200_comm_unique_id(char* out, size_t capacity) {
  if (!out || capacity < 40) return b200_set_error(B200_EINVAL, "b200_comm_unique_id", "buffer of at least 40 bytes needed");
  unsigned long long r = 0;
  int fd = open("/dev/urandom", O_RDONLY);
  if (fd >= 0) { if (read(fd, &r, sizeof r) != (ssize_t)sizeof r) r = 0; close(fd); }
  timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  snprintf(out, capacity, "kb200-%ld-%llx-%llx", (long)getpid(), (unsigned long long)ts.tv_nsec ^ ((unsigned long long)ts.tv_sec << 20), r);
  return 0;
}

int b200_comm_init(b200_instance* I, int rank, int world, const char* unique_id, b200_comm** out) {
  const char* where = "b200_comm_init";
  if (!I) return b200_set_error(B200_ENOTINIT, where, nullptr);
  if (!out) return b200_set_error(B200_EINVAL, where, "out is NULL");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return b200_set_error(B200_EINVAL, where, "rank/world out of range (one box: at most 8 ranks)");
  if (world > 1 && (!unique_id || !unique_id[0] || strlen(unique_id) > 60)) return b200_set_error(B200_EINVAL, where, "unique_id missing or longer than 60 characters");
  CU_TRY(cudaSetDevice(I->device), where);
  b200_comm* C = new (std::nothrow) b200_comm();
  if (!C) return b200_set_error(B200_ENOMEM, where, nullptr);
  C->inst = I; C->rank = rank; C->world = world;
  C->grid = I->props.sm_count < kChunkMaxGrid ? I->props.sm_count : kChunkMaxGrid;
  auto fail = [&](int code, const char* what) {
    b200_set_error(code, where, what);
    if (C->shm) { munmap(C->shm, sizeof(ShmHeader)); if (rank == 0) shm_unlink(C->shm_name); }
    if (C->mailbox) cudaFree(C->mailbox);
    if (C->scratch) cudaFree(C->scratch);
    if (C->err) cudaFreeHost(C->err);
    delete C;
    return code;
  };
  cudaError_t e;
  if ((e = cudaMalloc((void**)&C->mailbox, kMailboxBytes)) != cudaSuccess) return fail(B200_ENOMEM, "mailbox");
  if ((e = cudaMemset(C->mailbox, 0, kMailboxBytes)) != cudaSuccess) return fail((int)e, "mailbox memset");
  if ((e = cudaMalloc(&C->scratch, (size_t)kMaxWorld * kAgMaxBytes)) != cudaSuccess) return fail(B200_ENOMEM, "scratch");
  if ((e = cudaHostAlloc((void**)&C->err, 64, cudaHostAllocMapped)) != cudaSuccess) return fail((int)e, "error word");
  C->err[0] = 0;
  if ((e = cudaHostGetDevicePointer((void**)&C->err_dev, C->err, 0)) != cudaSuccess) return fail((int)e, "map error word");
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail((int)e, "sync");
  C->peer[rank] = C->mailbox;
  if (world == 1) { *out = C; return 0; }

  // ---- bootstrap: rank 0 creates the segment, everybody publishes an IPC handle, everybody maps everybody ----
  snprintf(C->shm_name, sizeof C->shm_name, "/%s", unique_id);
  for (char* c = C->shm_name + 1; *c; ++c) if (*c == '/') *c = '_';
  int fd = -1;
  const double t0 = now_s();
  if (rank == 0) {
    fd = shm_open(C->shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return fail(B200_EINVAL, "shm_open(create) failed: is the unique id really unique?");
    if (ftruncate(fd, sizeof(ShmHeader)) != 0) { close(fd); return fail(B200_ENOMEM, "ftruncate"); }
  } else {
    while ((fd = shm_open(C->shm_name, O_RDWR, 0600)) < 0) {
      if (now_s() - t0 > 120.0) return fail(B200_EINVAL, "timed out waiting for rank 0 to create the bootstrap segment");
      usleep(1000);
    }
    struct stat sb;  // rank 0 may not have sized it yet
    while (fstat(fd, &sb) == 0 && (size_t)sb.st_size < sizeof(ShmHeader)) {
      if (now_s() - t0 > 120.0) { close(fd); return fail(B200_EINVAL, "bootstrap segment never sized"); }
      usleep(1000);
    }
  }
  void* m = mmap(nullptr, sizeof(ShmHeader), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return fail(B200_ENOMEM, "mmap");
  C->shm = static_cast<ShmHeader*>(m);
  if (rank == 0) {
    C->shm->world = world;
    C->shm->magic.store(kMagic, std::memory_order_release);
  } else {
    while (C->shm->magic.load(std::memory_order_acquire) != kMagic) {
      if (now_s() - t0 > 120.0) return fail(B200_EINVAL, "bootstrap segment never initialised");
      usleep(200);
    }
    if (C->shm->world != world) return fail(B200_EINVAL, "ranks disagree on the world size");
  }
  ShmSlot& mine = C->shm->slot[rank];
  if ((e = cudaIpcGetMemHandle(&mine.handle, C->mailbox)) != cudaSuccess) return fail((int)e, "cudaIpcGetMemHandle");
  mine.device = I->device; mine.sm_count = I->props.sm_count; mine.pid = (long)getpid();
  C->shm->arrived[0].fetch_add(1, std::memory_order_acq_rel);
  if (!shm_wait(C->shm->arrived[0], world, 120.0)) return fail(B200_EINVAL, "timed out waiting for the other ranks (handles)");
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    const ShmSlot& s = C->shm->slot[q];
    if (s.pid == (long)getpid()) return fail(B200_EUNSUPPORTED, "one process per GPU: two ranks share a process");
    void* ptr = nullptr;
    if ((e = cudaIpcOpenMemHandle(&ptr, s.handle, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) return fail((int)e, "cudaIpcOpenMemHandle (is peer access available between the GPUs?)");
    C->peer[q] = static_cast<unsigned char*>(ptr);
    if (s.sm_count < C->grid) C->grid = s.sm_count;
  }
  C->shm->arrived[1].fetch_add(1, std::memory_order_acq_rel);
  if (!shm_wait(C->shm->arrived[1], world, 120.0)) return fail(B200_EINVAL, "timed out waiting for the other ranks (mapping)");
  if (rank == 0) shm_unlink(C->shm_name);  // the mapping stays; the name is gone, nothing is left behind on a crash
  *out = C;
  return 0;
}

int b200_comm_host_barrier(b200_comm* C) {
  if (!C) return b200_set_error(B200_ENOTINIT, "b200_comm_host_barrier", nullptr);
  if (C->world == 1) return 0;
  // sense-reversing pair of counters: generation g uses counter g & 1, reset by the last rank of generation g + 1
  const int g = C->host_barrier_gen++;
  std::atomic<int>& c = C->shm->host_barrier[g & 1];
  std::atomic<int>& other = C->shm->host_barrier[(g + 1) & 1];
  if (c.fetch_add(1, std::memory_order_acq_rel) + 1 == C->world) other.store(0, std::memory_order_release);
  if (!shm_wait(c, C->world, 300.0)) return b200_set_error(B200_EINVAL, "b200_comm_host_barrier", "timed out");
  return 0;
}

int b200_comm_finalize(b200_comm* C) {
  if (!C) return b200_set_error(B200_ENOTINIT, "b200_comm_finalize", nullptr);
  cudaSetDevice(C->inst->device);
  cudaStreamSynchronize(C->inst->stream);
  if (C->world > 1) {
    b200_comm_host_barrier(C);  // nobody may still be writing into a mailbox that is about to be unmapped
    for (int q = 0; q < C->world; ++q)
      if (q != C->rank && C->peer[q]) cudaIpcCloseMemHandle(C->peer[q]);
    b200_comm_host_barrier(C);
    munmap(C->shm, sizeof(ShmHeader));
  }
  if (C->mailbox) cudaFree(C->mailbox);
  if (C->scratch) cudaFree(C->scratch);
  if (C->err) cudaFreeHost(C->err);
  delete C;
  return 0;
}

int b200_comm_rank(b200_comm* C) { return C ? C->rank : -1; }
int b200_comm_world(b200_comm* C) { return C ? C->world : 0; }
unsigned b200_comm_error(b200_comm* C) { return (C && C->err) ? C->err[0] : 0u; }

int b200_comm_barrier(b200_comm* C) {
  if (!C) return b200_set_error(B200_ENOTINIT, "b200_comm_barrier", nullptr);
  unsigned* s = static_cast<unsigned*>(C->scratch);
  return comm_collective(C, "b200_comm_barrier", s, s + 64, 4, kGather);
}

int b200_allgather_bytes(b200_comm* C, const void* src_dev, void* dst_dev, size_t bytes_per_rank) {
  return comm_collective(C, "b200_allgather_bytes", src_dev, dst_dev, bytes_per_rank, kGather);
}
int b200_allreduce_sum_f64(b200_comm* C, double* buf_dev, int count) { return comm_collective(C, "b200_allreduce_sum_f64", buf_dev, buf_dev, (size_t)count * 8, kSumF64); }
int b200_allreduce_min_f64(b200_comm* C, double* buf_dev, int count) { return comm_collective(C, "b200_allreduce_min_f64", buf_dev, buf_dev, (size_t)count * 8, kMinF64); }
int b200_allreduce_max_f64(b200_comm* C, double* buf_dev, int count) { return comm_collective(C, "b200_allreduce_max_f64", buf_dev, buf_dev, (size_t)count * 8, kMaxF64); }
int b200_allreduce_sum_i64(b200_comm* C, int64_t* buf_dev, int count) { return comm_collective(C, "b200_allreduce_sum_i64", buf_dev, buf_dev, (size_t)count * 8, kSumI64); }
int b200_allreduce_min_i64(b200_comm* C, int64_t* buf_dev, int count) { return comm_collective(C, "b200_allreduce_min_i64", buf_dev, buf_dev, (size_t)count * 8, kMinI64); }
int b200_allreduce_max_i64(b200_comm* C, int64_t* buf_dev, int count) { return comm_collective(C, "b200_allreduce_max_i64", buf_dev, buf_dev, (size_t)count * 8, kMaxI64); }
int b200_allreduce_minloc_f64(b200_comm* C, b200_valloc_f64* buf_dev) { return comm_collective(C, "b200_allreduce_minloc_f64", buf_dev, buf_dev, sizeof(b200_valloc_f64), kMinLoc); }
int b200_allreduce_maxloc_f64(b200_comm* C, b200_valloc_f64* buf_dev) { return comm_collective(C, "b200_allreduce_maxloc_f64", buf_dev, buf_dev, sizeof(b200_valloc_f64), kMaxLoc); }
int b200_allreduce_minmaxloc_f64(b200_comm* C, b200_minmaxloc_f64* buf_dev) { return comm_collective(C, "b200_allreduce_minmaxloc_f64", buf_dev, buf_dev, sizeof(b200_minmaxloc_f64), kMinMaxLoc); }

int b200_comm_cyclic_layout(b200_comm* C, int elem_bytes, int64_t n_global, int64_t* block_elems, int64_t* n_local, int64_t* nsteps) {
  const char* where = "b200_comm_cyclic_layout";
  if (!C) return b200_set_error(B200_ENOTINIT, where, nullptr);
  if ((elem_bytes != 4 && elem_bytes != 8) || n_global < 0) return b200_set_error(B200_EINVAL, where, "element size must be 4 or 8, length >= 0");
  const int64_t block = b200_tune("comm.scan_algo", 0) == 1 ? (int64_t)C->grid * 384 * (9 * 16 / elem_bytes)
                                                            : (elem_bytes == 8 ? round_block_elems<int64>() : round_block_elems<int>());
  const int64_t nblocks = (n_global + block - 1) / block;
  int64_t nl = 0;
  for (int64_t c = C->rank; c < nblocks; c += C->world) nl += (c == nblocks - 1) ? (n_global - c * block) : block;
  if (block_elems) *block_elems = block;
  if (n_local) *n_local = nl;
  if (nsteps) *nsteps = (nblocks + C->world - 1) / C->world;
  return 0;
}

int b200_comm_scan_excl_i64(b200_comm* C, const int64_t* x, int64_t* y, int64_t n_global, int64_t* total_host, int64_t* total_dev) {
  return comm_scan<int64, false>(C, "b200_comm_scan_excl_i64", (const int64*)x, (int64*)y, n_global, (int64*)total_host, (int64*)total_dev);
}
int b200_comm_scan_incl_i64(b200_comm* C, const int64_t* x, int64_t* y, int64_t n_global, int64_t* total_host, int64_t* total_dev) {
  return comm_scan<int64, true>(C, "b200_comm_scan_incl_i64", (const int64*)x, (int64*)y, n_global, (int64*)total_host, (int64*)total_dev);
}
int b200_comm_scan_excl_f64(b200_comm* C, const double* x, double* y, int64_t n_global, double* total_host, double* total_dev) {
  return comm_scan<double, false>(C, "b200_comm_scan_excl_f64", x, y, n_global, total_host, total_dev);
}

}  // extern "C"

#ifdef B200_SWEEP
// the counters of THIS translation unit's kernels (tools/cyclic_probe.py; sweep build only)
extern "C" int b200_debug_comm_scan_stats(unsigned long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, kb200::Impl::g_scan_stats, 16 * sizeof(unsigned long long));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(kb200::Impl::g_scan_stats, z, sizeof z); }
  return 0;
}
extern "C" int b200_debug_round_ts(unsigned long long* out_4x8192) {
  return (int)cudaMemcpyFromSymbol(out_4x8192, kb200::Impl::g_round_ts, 4 * 8192 * sizeof(unsigned long long));
}
#endif
