// kb200/impl/ScanChunked.hpp -- chunk-synchronous single-pass prefix sum: the scan that stays at 16 B/element
// on ANY number of GPUs (8 read + 8 written per int64 element), one kernel per GPU, compute and exchange fused.
//
// Why it exists.  With contiguous per-GPU shards every element on rank r depends on ALL data of the ranks below,
// so a distributed scan has to read its input twice (shard totals, then the seeded scan: 24 B/element, the
// reference's two-launch Cuda scan costs the same on ONE GPU, core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:390-1047).
// Here the View is distributed BLOCK-CYCLICALLY with a block = what one GPU keeps ON CHIP:
//     block  = G x SLICE elements      (G = CTAs = SMs, SLICE = one shared-memory stage of one CTA)
//     global block c lives on rank c % world as local block c / world
// and all GPUs advance in lock step, one block per step:
//     step k, every CTA b:  bulk-load its slice (TMA, cp.async.bulk) into a shared-memory stage
//                           sum it                          -> desc[k][b]        (this GPU, L2)
//                           all CTAs read all G descriptors -> exclusive offset of the slice inside the block
//                                                              and the block AGGREGATE
//                           the aggregate is stored into every peer's mailbox over NVLink (st.relaxed.sys, LL words)
//                           all CTAs read the `world` mailbox entries of the step
//                           prefix = running + sum(aggregates of lower ranks) + offset;  running += sum(all aggregates)
//                           scan the slice in registers out of the SAME shared-memory stage, bulk-store it
// The slice never leaves shared memory between the sum and the scan, so HBM sees each element once in and once out;
// the exchange per step is 16 bytes per peer and its latency is hidden by the NSTAGE-deep stage ring (the loads of
// steps k+1.. are in flight while step k waits for its prefix).  world == 1 is the same kernel without mailboxes.
//
// Synchronisation words are "LL" (low-latency) pairs: every 8-byte word carries 32 bits of payload and the 32-bit
// step tag, so a reader never needs an ordering fence and nothing is ever cleared (tags only grow).
// All CTAs must be co-resident (they wait for each other every step): the launch is cooperative, grid = SMs.
#ifndef KB200_IMPL_SCANCHUNKED_HPP
#define KB200_IMPL_SCANCHUNKED_HPP

#include "ScanContig.hpp"

namespace kb200 {
namespace Impl {

constexpr int kChunkRing = 16;      // ring depth (steps) of descriptor rows and mailbox rows; must exceed NSTAGE
constexpr int kChunkMaxWorld = 8;   // GPUs of one NVSwitch box
constexpr int kChunkMaxGrid = 256;  // CTAs (>= SM count)
constexpr size_t kChunkDescBytes = (size_t)kChunkRing * kChunkMaxGrid * 16;
constexpr size_t kChunkMboxBytes = (size_t)kChunkRing * kChunkMaxWorld * 16;

template <class T>
struct ChunkScanParams {
  const T* x;
  T* y;
  int64 n;        // local elements
  int64 nsteps;   // identical on every rank
  T seed;
  const T* seed_dev;
  int seed_count;
  int rank, world;
  unsigned desc_tag_base;  // tag of step k = base + k (never 0); descriptors are local, mailboxes shared: two tag spaces
  unsigned mbox_tag_base;
  unsigned long long* desc;                       // [kChunkRing][kChunkMaxGrid][2]
  unsigned long long* mbox;                       // [kChunkRing][kChunkMaxWorld][2], written by the peers
  unsigned long long* peer_mbox[kChunkMaxWorld];  // the same array on every rank (peer-mapped)
  T* total0;
  T* total1;
  unsigned* err;  // pinned host word: set before a time-out trap
  int bulk_load, bulk_store;
  unsigned long long timeout_ns;
  unsigned long long* stats;  // sweep builds: 16 counters written by CTA 0 (cycles), else unused
};

#ifdef B200_SWEEP
#define KB200_CS_T0() const long long cs_t0 = clock64()
#define KB200_CS_ADD(var) var += clock64() - cs_t0
#else
#define KB200_CS_T0() do {} while (0)
#define KB200_CS_ADD(var) do {} while (0)
#endif


// warps: [0, SCAN_WARPS) scan | [.., +RED_WARPS) reduce | LOAD | STORE | NSTAGE x SYNC
template <class T, int SCAN_THREADS, int NV, int NSTAGE, int RED_WARPS, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS + RED_WARPS * 32 + 64 + 32 * NSTAGE, 1) chunk_scan_kernel(const ChunkScanParams<T> p) {
  static_assert(NV % 2 == 1, "odd vector count keeps blocked smem accesses conflict free");
  static_assert(NSTAGE < kChunkRing, "descriptor ring must be deeper than the stage ring");
  constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  constexpr int EPV = 16 / (int)sizeof(T);
  constexpr int SLICE = SCAN_THREADS * ITEMS;
  constexpr unsigned SLICE_BYTES = SLICE * sizeof(T);
  constexpr int NPIECE = 4;  // bulk copies per slice (keeps several TMA requests of one CTA in flight)
  static_assert(SLICE_BYTES % (NPIECE * 16) == 0, "pieces must stay 16-byte multiples");
  constexpr unsigned PIECE_BYTES = SLICE_BYTES / NPIECE;
  constexpr int SCAN_WARPS = SCAN_THREADS / 32;
  constexpr int RED_THREADS = RED_WARPS * 32;
  constexpr int MAXSLOT = kChunkMaxGrid / 32;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* const bufs = reinterpret_cast<T*>(smem_raw);
  __shared__ __align__(8) unsigned long long full[NSTAGE], aggready[NSTAGE], prefready[NSTAGE], outready[NSTAGE], empty[NSTAGE];
  __shared__ T s_prefix[NSTAGE];
  __shared__ T s_running[NSTAGE];
  __shared__ T s_red[NSTAGE][RED_WARPS];
  __shared__ T s_warp[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, G = gridDim.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      ptx::mbar_init(&full[s], 1); ptx::mbar_init(&aggready[s], 1); ptx::mbar_init(&prefready[s], 1);
      ptx::mbar_init(&outready[s], 1); ptx::mbar_init(&empty[s], 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
#ifdef B200_SWEEP
  long long c0 = 0, c1 = 0, c2 = 0;
  unsigned long long* const stat = (p.stats && (b == 0 || b == G - 1)) ? p.stats + (b == 0 ? 0 : 16) : nullptr;
#endif

  // local element range of this CTA's slice in step k
  auto slice_base = [&](int64 k) -> int64 { return (k * G + b) * (int64)SLICE; };
  auto slice_valid = [&](int64 k) -> int {
    const int64 rem = p.n - slice_base(k);
    return rem <= 0 ? 0 : (rem >= SLICE ? SLICE : (int)rem);
  };

  if (warp == SCAN_WARPS + RED_WARPS) {
    // ================= LOAD warp =================
    for (int64 k = 0; k < p.nsteps; ++k) {
      const int st = (int)(k % NSTAGE);
      if (k >= NSTAGE) {
        KB200_CS_T0();
        ptx::mbar_wait(&empty[st], (unsigned)(((k / NSTAGE) - 1) & 1));
        KB200_CS_ADD(c0);
      }
      const int valid = slice_valid(k);
      T* const buf = bufs + (size_t)st * SLICE;
      const T* const src = p.x + slice_base(k);
      if (valid == SLICE && p.bulk_load) {
        if (lane == 0) {
          ptx::mbar_expect_tx(&full[st], SLICE_BYTES);
#pragma unroll
          for (int q = 0; q < NPIECE; ++q)
            ptx::bulk_g2s(reinterpret_cast<unsigned char*>(buf) + q * PIECE_BYTES, reinterpret_cast<const unsigned char*>(src) + q * PIECE_BYTES,
                          PIECE_BYTES, &full[st]);
        }
      } else {
        if (valid > 0)
          for (int i = lane; i < SLICE; i += 32) buf[i] = (i < valid) ? src[i] : T(0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
      }
    }
#ifdef B200_SWEEP
    if (stat && lane == 0) stat[0] = c0;  // LOAD: cycles waiting for a free stage
#endif
    return;
  }

  if (warp == SCAN_WARPS + RED_WARPS + 1) {
    // ================= STORE warp =================
    for (int64 k = 0; k < p.nsteps; ++k) {
      const int st = (int)(k % NSTAGE);
      {
        KB200_CS_T0();
        ptx::mbar_wait(&outready[st], (unsigned)((k / NSTAGE) & 1));
        KB200_CS_ADD(c0);
      }
      const int valid = slice_valid(k);
      const T* const buf = bufs + (size_t)st * SLICE;
      T* const dst = p.y + slice_base(k);
      KB200_CS_T0();
      if (valid == SLICE && p.bulk_store) {
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < NPIECE; ++q)
            ptx::bulk_s2g(reinterpret_cast<unsigned char*>(dst) + q * PIECE_BYTES, reinterpret_cast<const unsigned char*>(buf) + q * PIECE_BYTES, PIECE_BYTES);
          ptx::bulk_commit();
          ptx::bulk_wait_read<0>();
        }
      } else {
        for (int i = lane; i < valid; i += 32) dst[i] = buf[i];
      }
      __syncwarp();
      KB200_CS_ADD(c1);
      if (lane == 0) mbar_arrive(&empty[st]);
    }
#ifdef B200_SWEEP
    if (stat && lane == 0) { stat[1] = c0; stat[2] = c1; }  // STORE: waiting for results; store issue + smem drain
#endif
    return;
  }

  if (warp >= SCAN_WARPS + RED_WARPS + 2) {
    // ================= SYNC warps, one per stage: block offsets on this GPU, aggregates across GPUs =================
    // Warp s owns the steps k = s (mod NSTAGE), so the exchanges of NSTAGE consecutive steps overlap (a single warp walking
    // the steps one after the other made (local round trip + NVLink round trip) the step time).  Each lane keeps its share
    // of the G descriptors of a step in flight TOGETHER (one L2 round trip under load costs more than a microsecond).
    // The running total is handed from the warp of step k-1 to the warp of step k through shared memory (prefready[]).
    const int sw = warp - (SCAN_WARPS + RED_WARPS + 2);
    const int nslot = (G + 31) / 32;
    T running = T(0);
    for (int64 k = sw; k < p.nsteps; k += NSTAGE) {
      const int st = sw;
      const unsigned par = (unsigned)((k / NSTAGE) & 1);
      const int row = (int)(k % kChunkRing);
      const unsigned dtag = p.desc_tag_base + (unsigned)k;
      const unsigned long long* const drow = p.desc + (size_t)row * kChunkMaxGrid * 2;
      ptx::mbar_wait(&aggready[st], par);  // this CTA has published: the others are about as far
      T val[MAXSLOT];
      unsigned pend = 0;
#pragma unroll
      for (int j = 0; j < MAXSLOT; ++j) {
        val[j] = T(0);
        if (j < nslot && j * 32 + lane < G) pend |= 1u << j;
      }
      KB200_CS_T0();
      {
        unsigned long long t0 = 0;
        for (unsigned spin = 0; __any_sync(kFullMask, pend != 0); ++spin) {
          unsigned long long w0[MAXSLOT], w1[MAXSLOT];
#pragma unroll
          for (int j = 0; j < MAXSLOT; ++j)
            if (pend & (1u << j)) ll::ld_gpu(drow + 2 * (j * 32 + lane), w0[j], w1[j]);
#pragma unroll
          for (int j = 0; j < MAXSLOT; ++j)
            if ((pend & (1u << j)) && ll::ok(w0[j], w1[j], dtag)) {
              val[j] = ll::unpack<T>(w0[j], w1[j]);
              pend &= ~(1u << j);
            }
          if ((spin & 1023u) == 1023u) {
            const unsigned long long t = ll::now_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > p.timeout_ns) ll::give_up(p.err, 0xC0000000u | (unsigned)b);
          }
        }
      }
      T excl = T(0), tot = T(0);
#pragma unroll
      for (int j = 0; j < MAXSLOT; ++j)
        if (j < nslot) {
          tot += val[j];
          if (j * 32 + lane < b) excl += val[j];
        }
      tot = warp_sum_all<T>(tot);
      excl = warp_sum_all<T>(excl);
      KB200_CS_ADD(c0);
      T before = T(0), all = tot;
      if (p.world > 1) {
        KB200_CS_T0();
        const unsigned mtag = p.mbox_tag_base + (unsigned)k;
        const size_t moff = ((size_t)row * kChunkMaxWorld) * 2;
        // CTA q (and CTA q + world, for redundancy against a late CTA) sends this GPU's aggregate to peer q
        if (lane == 0 && b < 2 * p.world) {
          const int q = b % p.world;
          if (q != p.rank) {
            unsigned long long w0, w1;
            ll::pack(tot, mtag, w0, w1);
            ll::st_sys(p.peer_mbox[q] + moff + 2 * p.rank, w0, w1);
          }
        }
        T v = T(0);
        if (lane < p.world) v = (lane == p.rank) ? tot : ll::wait_value<T, true>(p.mbox + moff + 2 * lane, mtag, p.timeout_ns, p.err, 0xD0000000u | (unsigned)lane);
        __syncwarp();
        all = warp_sum_all<T>(v);
        before = warp_sum_all<T>(lane < p.rank ? v : T(0));
        KB200_CS_ADD(c1);
      }
      if (k > 0) {  // running total through step k-1, from the warp that owns it
        KB200_CS_T0();
        const int pst = (int)((k - 1) % NSTAGE);
        ptx::mbar_wait(&prefready[pst], (unsigned)(((k - 1) / NSTAGE) & 1));
        running = s_running[pst];
        KB200_CS_ADD(c2);
      }
      __syncwarp();
      if (lane == 0) {
        s_prefix[st] = running + before + excl;
        s_running[st] = running + all;
        mbar_arrive(&prefready[st]);
      }
      if (k == p.nsteps - 1 && b == 0 && lane == 0) {
        if (p.total0) *p.total0 = running + all;
        if (p.total1) *p.total1 = running + all;
      }
    }
#ifdef B200_SWEEP
    if (stat && lane == 0 && sw == 0) { stat[3] = c0 * NSTAGE; stat[4] = c1 * NSTAGE; stat[11] = c2 * NSTAGE; }  // SYNC (warp 0, scaled to all steps)
#endif
    return;
  }

  if (warp >= SCAN_WARPS) {
    // ================= REDUCE warps: slice sum as soon as the bytes land, published straight to the descriptor row =================
    const int rt = tid - SCAN_THREADS, rw = warp - SCAN_WARPS;
    for (int64 k = 0; k < p.nsteps; ++k) {
      const int st = (int)(k % NSTAGE);
      {
        KB200_CS_T0();
        ptx::mbar_wait(&full[st], (unsigned)((k / NSTAGE) & 1));
        KB200_CS_ADD(c0);
      }
      KB200_CS_T0();
      T part = T(0);
      if (slice_valid(k) > 0) {
        const uint4* src = reinterpret_cast<const uint4*>(bufs + (size_t)st * SLICE);
        T acc[4] = {T(0), T(0), T(0), T(0)};
        constexpr int NVEC = (int)(SLICE_BYTES / 16);
#pragma unroll 8
        for (int i = rt; i < NVEC; i += RED_THREADS) {
          const uint4 q = src[i];
          T e[EPV];
          memcpy(e, &q, 16);
#pragma unroll
          for (int j = 0; j < EPV; ++j) acc[j & 3] += e[j];
        }
        part = warp_sum_all<T>((acc[0] + acc[1]) + (acc[2] + acc[3]));
      }
      if (lane == 0) s_red[st][rw] = part;
      named_bar_sync(2, RED_THREADS);
      if (rt == 0) {
        T a = T(0);
#pragma unroll
        for (int w = 0; w < RED_WARPS; ++w) a += s_red[st][w];
        unsigned long long w0, w1;
        ll::pack(a, p.desc_tag_base + (unsigned)k, w0, w1);
        ll::st_gpu(p.desc + ((size_t)(k % kChunkRing) * kChunkMaxGrid + b) * 2, w0, w1);
        mbar_arrive(&aggready[st]);
      }
      KB200_CS_ADD(c1);
    }
#ifdef B200_SWEEP
    if (stat && rt == 0) { stat[5] = c0; stat[6] = c1; }  // REDUCE: waiting for the load; summing + publishing
#endif
    return;
  }

  // ================= SCAN warps =================
  T seed = p.seed;
  if (p.seed_dev) {
    seed = T(0);
    for (int j = 0; j < p.seed_count; ++j) seed += p.seed_dev[j];
  }
  for (int64 k = 0; k < p.nsteps; ++k) {
    const int st = (int)(k % NSTAGE);
    const unsigned par = (unsigned)((k / NSTAGE) & 1);
    {
      KB200_CS_T0();
      ptx::mbar_wait(&full[st], par);
      KB200_CS_ADD(c0);
    }
    T* const buf = bufs + (size_t)st * SLICE;
    T v[ITEMS];
    {
      const uint4* src = reinterpret_cast<const uint4*>(buf + (size_t)tid * ITEMS);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const uint4 q = src[j];
        memcpy(&v[j * EPV], &q, 16);
      }
    }
    T tsum = T(0);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) tsum += v[j];
    const T tincl = warp_incl_scan(tsum, lane);
    if (lane == 31) s_warp[warp] = tincl;
    named_bar_sync(1, SCAN_THREADS);
    T woff = T(0);
    {
      const T w = lane < SCAN_WARPS ? s_warp[lane] : T(0);
      const T wi = warp_incl_scan(w, lane);
      woff = ::kb200::Impl::shfl_idx((T)(wi - w), warp);
    }
    {
      KB200_CS_T0();
      ptx::mbar_wait(&prefready[st], par);
      KB200_CS_ADD(c1);
    }
    KB200_CS_T0();
    T run = seed + s_prefix[st] + woff + (tincl - tsum);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const T in = v[j];
      if (INCLUSIVE) { run += in; v[j] = run; } else { v[j] = run; run += in; }
    }
    {
      uint4* dst = reinterpret_cast<uint4*>(buf + (size_t)tid * ITEMS);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        uint4 q;
        memcpy(&q, &v[j * EPV], 16);
        dst[j] = q;
      }
    }
    ptx::fence_proxy_async_smem();
    named_bar_sync(1, SCAN_THREADS);  // also orders the s_warp reads above before the next step's writes
    if (tid == 0) mbar_arrive(&outready[st]);
    KB200_CS_ADD(c2);
  }
#ifdef B200_SWEEP
  if (stat && tid == 0) { stat[7] = c0; stat[8] = c1; stat[9] = c2; stat[10] = (unsigned long long)p.nsteps; }  // SCAN: wait load; wait prefix; finish
#endif
}

#ifdef B200_SWEEP
inline unsigned long long* chunk_stats_buffer() {  // 32 counters, device memory, one per process (probe builds only)
  static unsigned long long* buf = nullptr;
  if (!buf) { cudaMalloc((void**)&buf, 32 * 8); cudaMemset(buf, 0, 32 * 8); }
  return buf;
}
#endif

// peer-side view of a communicator as the launcher needs it (csrc/comm.cu fills it; nullptr = single GPU)
struct ChunkPeers {
  int rank = 0, world = 1;
  unsigned long long* mbox = nullptr;
  unsigned long long* peer_mbox[kChunkMaxWorld] = {};
  unsigned mbox_tag_base = 0;  // first tag of this launch; the communicator advances it by nsteps on every rank alike
};

template <class T, int SCAN_THREADS, int NV, int NSTAGE, int RED_WARPS, bool INCLUSIVE>
struct ChunkScanLaunch {
  static constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  static constexpr int SLICE = SCAN_THREADS * ITEMS;
  static constexpr size_t SMEM = (size_t)NSTAGE * SLICE * sizeof(T);
  static constexpr int THREADS = SCAN_THREADS + RED_WARPS * 32 + 64 + 32 * NSTAGE;

  static auto kernel() { return chunk_scan_kernel<T, SCAN_THREADS, NV, NSTAGE, RED_WARPS, INCLUSIVE>; }

  // CTAs the kernel can keep co-resident on `device` (0 = it does not fit)
  static int max_grid(int device, int sm_count) {
    static int cached[64] = {};
    if (device < 0 || device >= 64) return 0;
    if (cached[device] == 0) {
      auto k = kernel();
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM) != cudaSuccess) { cudaGetLastError(); return 0; }
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, THREADS, SMEM) != cudaSuccess) { cudaGetLastError(); return 0; }
      cached[device] = nb > 0 ? 1 : -1;
    }
    if (cached[device] < 0) return 0;
    return sm_count < kChunkMaxGrid ? sm_count : kChunkMaxGrid;
  }
  static int64 block_elems(int grid) { return (int64)grid * SLICE; }
  static int64 steps_for(int64 n_local, int grid) { return (n_local + block_elems(grid) - 1) / block_elems(grid); }

  // x, y: this rank's local blocks, contiguous; nsteps/grid identical on all ranks of `peers`
  static int run(b200_instance* inst, const ChunkPeers* peers, int grid, int64 nsteps, const T* x, T* y, int64 n, T seed, const T* seed_dev,
                 int seed_count, T* total_host, T* total_dev) {
    HostRuntime rt(inst);
    int rc;
    if (nsteps <= 0) {
      if (total_dev && (rc = b200_memset_async(inst, total_dev, 0, sizeof(T)))) return rc;
      if (total_host) {
        if ((rc = rt.fence("kb200::parallel_scan (empty)"))) return rc;
        *total_host = T(0);
      }
      return 0;
    }
    ChunkScanParams<T> p;
    memset(&p, 0, sizeof p);
    p.x = x; p.y = y; p.n = n; p.nsteps = nsteps; p.seed = seed; p.seed_dev = seed_dev; p.seed_count = seed_count;
    p.rank = peers ? peers->rank : 0;
    p.world = peers ? peers->world : 1;
    unsigned long long* desc = nullptr;
    unsigned* err = nullptr;
    if ((rc = b200_chunk_begin(inst, (uint64_t)nsteps, &p.desc_tag_base, &desc, &err))) return rc;
    p.desc = desc;
    p.err = err;
    if (peers && peers->world > 1) {
      p.mbox = peers->mbox;
      for (int q = 0; q < peers->world; ++q) p.peer_mbox[q] = peers->peer_mbox[q];
      p.mbox_tag_base = peers->mbox_tag_base;
    }
    void *slot_dev = nullptr, *slot_host = nullptr, *unused_p = nullptr;
    unsigned* unused_t = nullptr;
    if (total_host && (rc = rt.reduce_scratch(0, sizeof(T), true, &unused_p, &unused_t, &slot_dev, &slot_host))) return rc;
    p.total0 = total_host ? reinterpret_cast<T*>(slot_dev) : total_dev;
    p.total1 = total_host ? total_dev : nullptr;
    p.bulk_load = (reinterpret_cast<uintptr_t>(x) % 16 == 0);
    p.bulk_store = (reinterpret_cast<uintptr_t>(y) % 16 == 0);
    p.timeout_ns = 20ull * 1000000000ull;
#ifdef B200_SWEEP
    p.stats = chunk_stats_buffer();
#endif
    void* args[] = {&p};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)kernel(), dim3(grid), dim3(THREADS), args, SMEM, rt.stream());
    if (e != cudaSuccess) return b200_report_error((int)e, "kb200::chunk_scan_kernel (cooperative launch)");
    if (total_host) {
      if ((rc = rt.fence("kb200::parallel_scan: fence to hand the total to the host"))) return rc;
      memcpy(total_host, slot_host, sizeof(T));
    }
    return 0;
  }
};

}  // namespace Impl
}  // namespace kb200
#endif
