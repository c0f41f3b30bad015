// kb200/impl/ScanContig.hpp -- single-pass prefix sum over a contiguous View<T*> (T = 4 or 8 byte
// arithmetic), the B200 replacement for ParallelScan/ParallelScanWithTotal<...,RangePolicy,Cuda>
// (core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:390-701,704-1047: two launches, input read twice,
// functor called three times per index, >= 24 B/element for int64).
//
// Here: ONE launch, 16 B/element (8 read + 8 written), chained scan with decoupled look-back.
//   * persistent CTAs (SMs x resident) take tile ids from a monotonic atomic counter, so a tile's
//     predecessors are always owned by CTAs that already run (no co-residency assumption);
//   * tiles move global->shared and shared->global with 1-D bulk async copies (TMA engine,
//     cp.async.bulk, SASS UBLKCP/UBLKPF) through an NBUF-deep ring: the load of tile k+1 and the
//     store of tile k-1 overlap the scan of tile k, and no register is spent on staging;
//   * each thread owns an ODD number of 16-byte vectors, so its blocked LDS.128/STS.128 accesses
//     are bank-conflict free without padding (stride = odd x 16 B);
//   * tile descriptors are 16 bytes {value, epoch<<2|state} written/read with one relaxed
//     device-scope 128-bit access; the epoch tag means the arena is never cleared;
//   * look-back inspects 32*LBW predecessors per step to keep the chain shorter than the tile
//     arrival rate ("scan: look-back depth").
#ifndef KB200_IMPL_SCANCONTIG_HPP
#define KB200_IMPL_SCANCONTIG_HPP

#include "LL.hpp"
#include <type_traits>

#include "Collectives.hpp"
#include "HostRuntime.hpp"
#include "Ptx.hpp"

namespace kb200 {
namespace Impl {

#ifdef B200_SWEEP
// diagnostic counters (sweep build only): [0] look-backs [1] window steps [2] polls that found an unpublished predecessor
// [3] cycles inside look-back [4] compute-warp cycles waiting for the prefix [5] compute-warp cycles waiting for data
// [6] tiles (compute) [7] cycles the look-back warp waits for its own aggregate
__device__ unsigned long long g_scan_stats[16];
__device__ unsigned long long g_round_ts[4][8192];  // per round: [0] first tile landed [1] last tile landed [2] round aggregate sent [3] base published
// accumulated in registers, flushed once per warp at exit: per-event global atomics perturb the kernel badly
struct ScanStats {
  unsigned long long v[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // [8] cycles resolving the round base (non-leaders) [9] same, leaders [10] leaders [11] late bases
  KB200_DEVICE_FUNCTION void flush() {
    for (int i = 0; i < 12; ++i) if (v[i]) atomicAdd(&g_scan_stats[i], v[i]);
  }
};
#define KB200_STATS_DECL ScanStats kb_stats
#define KB200_STAT_ADD(i, x) (kb_stats.v[i] += (unsigned long long)(x))
#define KB200_STATS_FLUSH() kb_stats.flush()
#define KB200_STATS_ARG , ScanStats& kb_stats
#define KB200_STATS_PASS , kb_stats
#else
#define KB200_STATS_DECL
#define KB200_STAT_ADD(i, x) ((void)0)
#define KB200_STATS_FLUSH() ((void)0)
#define KB200_STATS_ARG
#define KB200_STATS_PASS
#endif

struct alignas(16) ScanDesc16 {
  unsigned long long payload;
  unsigned long long status;  // (epoch << 2) | state ; state 1 = tile aggregate, 2 = inclusive prefix
};
constexpr unsigned long long kDescAgg = 1ull, kDescIncl = 2ull;

template <class T>
struct ScanContigParams {
  const T* x;
  T* y;
  int64 n;
  int64 ntiles;
  T seed;
  const T* seed_dev;  // if non-null, the seed is the sum of seed_dev[0..seed_count) read when the kernel runs
  int seed_count;     // (a distributed scan passes the all-gathered shard totals and its rank)
  ScanDesc16* desc;
  unsigned long long epoch;
  unsigned long long* counter;
  unsigned long long counter_base;
  T* total0;
  T* total1;
  int bulk_load, bulk_store;  // 16-byte alignment of x / y allows the TMA path
  int spin_sleep_ns;          // back-off between polls of an unpublished predecessor (0 = none)
  int dbg_flags;              // tools/sweep.py experiments only: 1 = skip look-back, 2 = skip the scan (pure copy)
  int prefetch_tiles;         // > 0: whoever takes tile t also asks the L2 for tile t + prefetch_tiles (HBM runs ahead of the stage ring)
  // ---- multi-GPU "rounds" (block-cyclic distributed scan; csrc/comm.cu).  tpr == 0: single GPU, nothing below is used.
  int64 tpr;                  // tiles per round; a round (tpr * TILE elements) is the block of the block-cyclic distribution
  int rank, world;
  unsigned rtag_base;         // LL tag of round k = rtag_base + k (never 0), identical on all ranks
  unsigned long long* rdesc;  // this GPU: [kRoundRing][4] LL words {base(k), running_before(k)}, written by the round leader
  unsigned long long* mbox;   // this GPU: [kRoundRing][kRoundMaxWorld][2] LL words: round aggregates of every rank
  unsigned long long* peer_mbox[8];  // the same array on every rank (peer-mapped over NVLink)
  unsigned long long* racc;   // this GPU: [kRoundRing][2] round accumulators (integral T): Sum(lo32) | count<<48, Sum(hi32) | count<<48
  unsigned* err;              // pinned host word, set before a time-out trap
  unsigned long long timeout_ns;
};
constexpr int kRoundRing = 64;      // rows of the round descriptor / mailbox rings: live rounds (<= 2 + grid*NSTAGE/tpr) + look-back reach
constexpr int kRoundLookback = 16;  // how many rounds back a round leader looks for a published running total
constexpr int kRoundMaxWorld = 8;

template <class T>
KB200_DEVICE_FUNCTION T scan_seed(const ScanContigParams<T>& p) {
  if (!p.seed_dev) return p.seed;
  T s = T(0);
  for (int k = 0; k < p.seed_count; ++k) s += p.seed_dev[k];
  return s;
}

// called by ONE thread per CTA once the CTA will take no more tile ids: the last CTA re-arms the counters
KB200_DEVICE_FUNCTION void scan_counter_release(unsigned long long* counter) {
  __threadfence();
  const unsigned long long done = atomicAdd(counter + 1, 1ull);
  if (done == (unsigned long long)gridDim.x - 1ull) {
    counter[0] = 0ull;
    counter[1] = 0ull;
    __threadfence();
  }
}

template <class T>
KB200_DEVICE_FUNCTION T warp_sum_all(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += ::kb200::Impl::shfl_xor(v, m);
  return v;
}
template <class T>
KB200_DEVICE_FUNCTION T warp_incl_scan(T v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T u = ::kb200::Impl::shfl_up(v, d);
    if (lane >= d) v += u;
  }
  return v;
}

// exclusive prefix of tile `tile` (> 0): sum of the aggregates of all earlier tiles.  Warp-collective.
// One step inspects LBW sub-windows of 32 consecutive descriptors (lane-contiguous, 512 B per request: coalesced);
// all LBW requests are in flight together, then the sub-windows are evaluated nearest first.
KB200_DEVICE_FUNCTION void ld_desc(const ScanDesc16* d, unsigned long long& a, unsigned long long& b, int weak) {
  if (weak) {  // L2-only weak load: may be reordered/overlapped freely; the descriptor is self-contained in its 16 bytes
    asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(d));
  } else {
    ptx::ld_relaxed_v2(d, a, b);
  }
}

template <class T, int LBW>
KB200_DEVICE_FUNCTION T lookback_sum(const ScanDesc16* desc, int64 tile, unsigned long long epoch, int lane, int sleep_ns, int weak KB200_STATS_ARG, int64 lo = 0) {
  T excl = T(0);
  int64 wbase = tile - 1;
#ifdef B200_SWEEP
  const long long t_begin = clock64();
  if (lane == 0) KB200_STAT_ADD(0, 1);
#endif
  while (true) {
    unsigned long long pay[LBW], st[LBW];
#pragma unroll
    for (int j = 0; j < LBW; ++j) {
      const int64 idx = wbase - ((int64)j * 32 + lane);
      if (idx >= lo) {
        ld_desc(desc + idx, pay[j], st[j], weak);
      } else {
        pay[j] = 0;
        st[j] = (epoch << 2) | kDescIncl;  // before the first tile (of the round): inclusive prefix = identity
      }
    }
    bool retry = false, done = false;
#pragma unroll
    for (int j = 0; j < LBW; ++j) {
      if (!retry && !done) {
        const bool valid = (st[j] >> 2) == epoch;
        const bool incl = valid && ((st[j] & 3ull) == kDescIncl);
        const unsigned term = __ballot_sync(kFullMask, incl);
        const unsigned inval = __ballot_sync(kFullMask, !valid);
        const int first_term = term ? (__ffs(term) - 1) : 32;
        const unsigned needed = first_term >= 31 ? kFullMask : ((2u << first_term) - 1u);
        if (inval & needed) {
          retry = true;  // a needed predecessor has not published yet; what was summed so far stays valid
          if (lane == 0) KB200_STAT_ADD(2, 1);
        } else {
          excl += warp_sum_all<T>(lane <= first_term ? from_bits<T>(pay[j]) : T(0));
          if (lane == 0) KB200_STAT_ADD(1, 1);
          if (term) done = true; else wbase -= 32;
        }
      }
    }
#ifdef B200_SWEEP
    if (done && lane == 0) KB200_STAT_ADD(3, clock64() - t_begin);
#endif
    if (done) return excl;
    if (retry && sleep_ns > 0) __nanosleep(sleep_ns);
  }
}

// Tried and dropped (profiles/r02_scan_assist_probe.log): an "assisted" look-back in which every look-back also publishes the
// inclusive prefixes of the tiles inside its window (suffix sums over the lanes) and owners first check whether their own
// descriptor was upgraded.  It shortens the chain but costs up to 31 extra 16-byte stores and one more load per tile:
// 5.89 TB/s against 6.18 TB/s for the plain look-back above at 2^30 int64.

// NV = 16-byte vectors per thread (odd).  ITEMS = NV*16/sizeof(T).
template <class T, int BLOCK, int NV, int NBUF, int LBW, bool INCLUSIVE>
__global__ void __launch_bounds__(BLOCK) contig_scan_kernel(const ScanContigParams<T> p) {
  static_assert(NV % 2 == 1, "odd vector count keeps blocked smem accesses conflict free");
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "");
  constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  constexpr int EPV = 16 / (int)sizeof(T);
  constexpr int TILE = BLOCK * ITEMS;
  constexpr unsigned TILE_BYTES = TILE * sizeof(T);
  constexpr int NWARPS = BLOCK / 32;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* const bufs = reinterpret_cast<T*>(smem_raw);
  __shared__ __align__(8) unsigned long long mbar[NBUF];
  __shared__ T s_warp[32];
  __shared__ T s_tile_prefix;
  __shared__ int64 s_tile[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KB200_STATS_DECL;
  const T seed = scan_seed(p);

  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NBUF; ++b) ptx::mbar_init(&mbar[b], 1);
    ptx::fence_mbar_init();
    s_tile[0] = (int64)(atomicAdd(p.counter, 1ull) - p.counter_base);
  }
  __syncthreads();
  int64 cur = s_tile[0];
  int stage = 0;
  unsigned parity = 0;  // bit b = parity to wait for on mbar[b]
  int it = 0;

  auto issue_load = [&](int64 tile, int st) {  // thread 0 only
    const int64 base = tile * TILE;
    if (p.bulk_load && base + TILE <= p.n) {
      ptx::mbar_expect_tx(&mbar[st], TILE_BYTES);
      ptx::bulk_g2s(bufs + (size_t)st * TILE, p.x + base, TILE_BYTES, &mbar[st]);
    }
  };
  if (tid == 0 && cur < p.ntiles) issue_load(cur, 0);

  while (cur < p.ntiles) {
    // ---- take the next tile id and start its load while this tile is processed
    const int nstage = (stage + 1 == NBUF) ? 0 : stage + 1;
    if (tid == 0) {
      const int64 nxt = (int64)(atomicAdd(p.counter, 1ull) - p.counter_base);
      s_tile[(it + 1) & 1] = nxt;
      if (nxt < p.ntiles) {
        ptx::bulk_wait_read<NBUF - 2>();  // the store that last read buffer `nstage` has drained
        issue_load(nxt, nstage);
      }
    }
    const int64 base = cur * TILE;
    const int64 remaining = p.n - base;
    const bool full = remaining >= TILE;
    T* const buf = bufs + (size_t)stage * TILE;

    if (full && p.bulk_load) {
      ptx::mbar_wait(&mbar[stage], (parity >> stage) & 1u);
      parity ^= 1u << stage;
    } else {
      // ragged or unaligned tile: cooperative coalesced loads, identity padding
      for (int i = tid; i < TILE; i += BLOCK) buf[i] = (i < remaining) ? p.x[base + i] : T(0);
      __syncthreads();
    }

    // ---- blocked read: NV conflict-free LDS.128 per thread
    T v[ITEMS];
    {
      const uint4* src = reinterpret_cast<const uint4*>(buf + (size_t)tid * ITEMS);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const uint4 q = src[k];
        memcpy(&v[k * EPV], &q, 16);
      }
    }
    T tsum = T(0);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) tsum += v[k];
    const T tincl = warp_incl_scan(tsum, lane);
    if (lane == 31) s_warp[warp] = tincl;
    __syncthreads();  // (A) also publishes s_tile[(it+1)&1]
    if (warp == 0) {
      const T w = lane < NWARPS ? s_warp[lane] : T(0);
      const T wi = warp_incl_scan(w, lane);
      if (lane < NWARPS) s_warp[lane] = wi - w;  // exclusive warp offsets
      const T agg = ::kb200::Impl::shfl_idx(wi, NWARPS - 1);
      ScanDesc16* const d = p.desc + cur;
      T excl = T(0);
      if (cur == 0) {
        if (lane == 0) ptx::st_relaxed_v2(d, to_bits(agg), (p.epoch << 2) | kDescIncl);
      } else {
        if (lane == 0) ptx::st_relaxed_v2(d, to_bits(agg), (p.epoch << 2) | kDescAgg);
        if (!(p.dbg_flags & 1)) excl = lookback_sum<T, LBW>(p.desc, cur, p.epoch, lane, p.spin_sleep_ns, p.dbg_flags & 8 KB200_STATS_PASS);
        if (lane == 0) ptx::st_relaxed_v2(d, to_bits((T)(excl + agg)), (p.epoch << 2) | kDescIncl);
      }
      if (lane == 0) {
        s_tile_prefix = excl;
        if (cur == p.ntiles - 1) {
          const T total = excl + agg;
          if (p.total0) *p.total0 = total;
          if (p.total1) *p.total1 = total;
        }
      }
    }
    __syncthreads();  // (B)
    T run = seed + s_tile_prefix + s_warp[warp] + (tincl - tsum);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const T in = v[k];
      if (INCLUSIVE) { run += in; v[k] = run; } else { v[k] = run; run += in; }
    }
    {
      uint4* dst = reinterpret_cast<uint4*>(buf + (size_t)tid * ITEMS);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        uint4 q;
        memcpy(&q, &v[k * EPV], 16);
        dst[k] = q;
      }
    }
    if (full && p.bulk_store) {
      ptx::fence_proxy_async_smem();
      __syncthreads();  // (C)
      if (tid == 0) {
        ptx::bulk_s2g(p.y + base, buf, TILE_BYTES);
        ptx::bulk_commit();
      }
    } else {
      __syncthreads();
      for (int i = tid; i < TILE && i < remaining; i += BLOCK) p.y[base + i] = buf[i];
      __syncthreads();  // buffer reusable
    }
    ++it;
    cur = s_tile[it & 1];
    stage = nstage;
  }
  if (tid == 0) {
    ptx::bulk_wait_read<0>();  // shared memory must outlive the last bulk store's reads
    scan_counter_release(p.counter);
  }
}

KB200_DEVICE_FUNCTION void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
KB200_DEVICE_FUNCTION void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
template <int MAXK>
KB200_DEVICE_FUNCTION void bulk_wait_read_dyn(int k) {
  if constexpr (MAXK <= 0) { ptx::bulk_wait_read<0>(); }
  else {
    if (k >= MAXK) ptx::bulk_wait_read<MAXK>();
    else bulk_wait_read_dyn<MAXK - 1>(k);
  }
}

// single-producer request ring (shared memory) from a service warp to the CTA's round warp; called by ONE lane
template <int RQ>
KB200_DEVICE_FUNCTION void rq_push(int64& nreq, unsigned long long* rq_full, unsigned long long* rq_empty, int64* s_rq, int64 kk) {
  const int slot = (int)(nreq % RQ);
  if (nreq >= RQ) ptx::mbar_wait(&rq_empty[slot], (unsigned)(((nreq / RQ) - 1) & 1));
  s_rq[slot] = kk;
  ptx::mbar_arrive(&rq_full[slot]);
  ++nreq;
}

// Round k of a block-cyclic distributed scan (ROUNDS kernels):
//   prefix of a tile = base(k) + (prefix inside the round),
//   base(k) = running_before(k) + aggregates of round k on the ranks below this one,
//   running_before(k) = total of rounds < k over ALL ranks.
// round_publish (warp-collective, run by the CTA's ROUND warp) computes base(kk) and publishes it in this GPU's round
// descriptor; it is requested by the look-back warp that finished the LAST tile of round kk-1 (or tile 0 for kk = 0), so
// it starts the moment this GPU's aggregate of round kk-1 exists and never queues behind parked tiles.  kk == nrounds
// writes the global total instead.  running_before comes from a round-level look-back: the nearest earlier round whose
// running_before is published plus the aggregates (all ranks) from there on -- publishers never wait for each other in
// a chain (a chain costs several L2 round trips per round: measured 4.6 us/round against a 1.5 us round time).
template <class T>
KB200_DEVICE_FUNCTION void round_publish(const ScanContigParams<T>& p, int64 kk, int64 nrounds, int lane, bool have_own, T own_prev) {
  T rb = T(0);
  if (kk > 0) {
    int m = 0;
    T rbm = T(0);
    {
      unsigned long long t0 = 0;
      for (unsigned spin = 0;; ++spin) {
        const int64 k2 = kk - 1 - lane;
        bool have = false;
        T v = T(0);
        if (k2 >= 0 && lane < kRoundLookback) {
          unsigned long long w0, w1;
          ll::ld_gpu(p.rdesc + (size_t)(k2 % kRoundRing) * 4 + 2, w0, w1);
          have = ll::ok(w0, w1, p.rtag_base + (unsigned)k2);
          v = ll::unpack<T>(w0, w1);
        }
        const unsigned bal = __ballot_sync(kFullMask, have);
        if (bal) {
          m = __ffs(bal) - 1;
          rbm = ::kb200::Impl::shfl_idx(v, m);
          break;
        }
        if ((spin & 1023u) == 1023u) {
          const unsigned long long t = ll::now_ns();
          if (t0 == 0) t0 = t;
          else if (t - t0 > p.timeout_ns) ll::give_up(p.err, 0xD2000000u);
        }
      }
    }
    T part = T(0);
    const int cnt = (m + 1) * p.world;
    for (int e = lane; e < cnt; e += 32) {
      const int64 k2 = kk - 1 - e / p.world;
      if (have_own && e / p.world == 0 && e % p.world == p.rank) { part += own_prev; continue; }  // just collected by this warp
      part += ll::wait_value<T, true>(p.mbox + ((size_t)(k2 % kRoundRing) * kRoundMaxWorld + (e % p.world)) * 2, p.rtag_base + (unsigned)k2,
                                      p.timeout_ns, p.err, 0xD1000000u | (unsigned)(e % p.world));
    }
    __syncwarp();
    rb = rbm + warp_sum_all<T>(part);
  }
  if (kk == nrounds) {  // past the last round: running_before is the global total
    if (lane == 0) {
      if (p.total0) *p.total0 = rb;
      if (p.total1) *p.total1 = rb;
    }
    return;
  }
  const int row = (int)(kk % kRoundRing);
  const unsigned tag = p.rtag_base + (unsigned)kk;
  if (lane == 0) {  // running_before first: later publishers may already build on it while this one waits for the lower ranks
    unsigned long long w0, w1;
    ll::pack(rb, tag, w0, w1);
    ll::st_gpu(p.rdesc + (size_t)row * 4 + 2, w0, w1);
  }
  T v = T(0);
  if (lane < p.rank) v = ll::wait_value<T, true>(p.mbox + (size_t)row * kRoundMaxWorld * 2 + 2 * lane, tag, p.timeout_ns, p.err, 0xD3000000u | (unsigned)lane);
  __syncwarp();
  const T before = warp_sum_all<T>(v);
  if (lane == 0) {
    unsigned long long w0, w1;
    ll::pack((T)(rb + before), tag, w0, w1);
    ll::st_gpu(p.rdesc + (size_t)row * 4, w0, w1);
#ifdef B200_SWEEP
    if (kk < 8192) g_round_ts[3][kk] = ll::now_ns();
#endif
  }
}

// ---------------------------------------------------------------------------------------------
// Fully decoupled variant: DMA warp + AGGREGATE warp + LOOK-BACK warp + compute warps.
//
// Measured on B200 (profiles/r01_scan_probe.log): with the look-back removed both kernels above run at
// 6.8 TB/s, with it 2.5-3.1 TB/s.  Cause: tile ids are taken NSTAGE tiles ahead (that is what lets the loads
// run ahead), but a tile's aggregate was only published when the compute warps reached it -- one full tile
// time after tiles with HIGHER ids (held by other CTAs) started to wait for it: a convoy.  Here a tile's
// aggregate is published by a dedicated warp as soon as its bytes land, and the chain resolution runs in a
// third warp, so publication never queues behind another tile's look-back or the compute pipeline:
//   DMA warp      : tile id (atomic) -> bulk load -> full[s];   outready[s] -> bulk store
//   AGGREGATE warp: full[s] -> sum the stage (lane-strided LDS.128) -> publish AGGREGATE -> aggready[s]
//   LOOK-BACK warp: aggready[s] -> look-back -> publish INCLUSIVE -> prefix[s] -> prefready[s]
//   compute warps : full[s] -> blocked scan in registers -> prefready[s] -> add prefix -> smem -> outready[s]
template <class T, int CBLOCK, int NV, int NSTAGE, int LBW, bool INCLUSIVE, bool ROUNDS = false>
__global__ void __launch_bounds__(CBLOCK + 96 + (ROUNDS ? 32 : 0)) contig_scan_ws2_kernel(const ScanContigParams<T> p) {
  static_assert(NV % 2 == 1, "odd vector count keeps blocked smem accesses conflict free");
  constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  constexpr int EPV = 16 / (int)sizeof(T);
  constexpr int TILE = CBLOCK * ITEMS;
  constexpr unsigned TILE_BYTES = TILE * sizeof(T);
  constexpr int NWARPS = CBLOCK / 32;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* const bufs = reinterpret_cast<T*>(smem_raw);
  __shared__ __align__(8) unsigned long long full[NSTAGE], aggready[NSTAGE], prefready[NSTAGE], outready[NSTAGE];
  __shared__ int64 s_tile_id[NSTAGE];
  __shared__ T s_agg[NSTAGE];
  __shared__ T s_prefix[NSTAGE];
  __shared__ T s_warp[32];
  __shared__ T s_base;  // ROUNDS: base of the round of the tile the compute warps are working on
  constexpr int RQ = 8;  // ROUNDS: requests "publish base(kk)" from a service warp to the round warp
  // integral T: tiles ADD their aggregate to a per-round accumulator (two fire-and-forget reds, no ordering needed: each
  // word carries its own arrival count), so the round aggregate exists ~1 us after the last tile landed instead of after
  // the last tile's look-back (measured 4.9 us).  Floating point keeps the look-back value (fixed summation order).
  constexpr bool ATOMIC_AGG = ROUNDS && std::is_integral<T>::value;
  __shared__ __align__(8) unsigned long long rq_full[RQ], rq_empty[RQ];
  __shared__ int64 s_rq[RQ];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KB200_STATS_DECL;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NSTAGE; ++b) {
      ptx::mbar_init(&full[b], 1); ptx::mbar_init(&aggready[b], 1);
      ptx::mbar_init(&prefready[b], 1); ptx::mbar_init(&outready[b], 1);
    }
    if constexpr (ROUNDS) {
#pragma unroll
      for (int b = 0; b < RQ; ++b) { ptx::mbar_init(&rq_full[b], 1); ptx::mbar_init(&rq_empty[b], 1); }
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if constexpr (ROUNDS) {
    if (warp == NWARPS + 3) {
      // ================= ROUND warp: publishes round bases on request (see round_publish) =================
      const int64 nrounds = p.ntiles / p.tpr;
      for (int64 c = 0;; ++c) {
        const int slot = (int)(c % RQ);
        ptx::mbar_wait(&rq_full[slot], (unsigned)((c / RQ) & 1));
        const int64 kk = s_rq[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&rq_empty[slot]);
        if (kk < 0) return;
        T own_prev = T(0);
        if constexpr (ATOMIC_AGG) {
          if (kk > 0) {  // collect this GPU's aggregate of round kk-1 from the accumulator, re-arm it, send it to every rank
            const int64 k1 = kk - 1;
            unsigned long long* const acc = p.racc + (size_t)(k1 % kRoundRing) * 2;
            unsigned long long w0 = 0, w1 = 0;
            if (lane == 0) {
              unsigned long long t0 = 0;
              for (unsigned spin = 0;; ++spin) {
                ll::ld_gpu(acc, w0, w1);
                if ((w0 >> 48) == (unsigned long long)p.tpr && (w1 >> 48) == (unsigned long long)p.tpr) break;
                if ((spin & 1023u) == 1023u) {
                  const unsigned long long t = ll::now_ns();
                  if (t0 == 0) t0 = t;
                  else if (t - t0 > p.timeout_ns) ll::give_up(p.err, 0xD7000000u);
                }
              }
              ll::st_gpu(acc, 0ull, 0ull);
            }
            const unsigned long long bits = (w0 & 0xffffffffffffull) + ((w1 & 0xffffffffffffull) << 32);
            own_prev = ::kb200::Impl::shfl_idx(from_bits<T>(bits), 0);
            if (lane < p.world) {
              unsigned long long a0, a1;
              ll::pack(own_prev, p.rtag_base + (unsigned)k1, a0, a1);
              ll::st_sys(p.peer_mbox[lane] + (size_t)(k1 % kRoundRing) * kRoundMaxWorld * 2 + 2 * p.rank, a0, a1);
            }
#ifdef B200_SWEEP
            if (lane == 0 && k1 < 8192) g_round_ts[2][k1] = ll::now_ns();
#endif
          }
        }
        round_publish<T>(p, kk, nrounds, lane, ATOMIC_AGG && kk > 0, own_prev);
      }
    }
  }

  if (warp == NWARPS) {
    // ================= DMA warp =================
    int64 jl = 0, js = 0, nvalid = 0;
    int64 tl[NSTAGE];
    bool more = true;
    while (more || js < nvalid) {
      if (more && jl - js < NSTAGE) {
        const int st = (int)(jl % NSTAGE);
        if (jl >= NSTAGE && lane == 0) bulk_wait_read_dyn<NSTAGE - 1>((int)(js - 1 - (jl - NSTAGE)));
        long long tile = 0;
        if (lane == 0) tile = (long long)(atomicAdd(p.counter, 1ull) - p.counter_base);
        tile = __shfl_sync(kFullMask, tile, 0);
        if (lane == 0) s_tile_id[st] = tile;
#pragma unroll
        for (int b = 0; b < NSTAGE; ++b) if (b == st) tl[b] = tile;
        if (tile >= p.ntiles) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[st]);
          more = false;
        } else {
          const int64 base = tile * TILE;
          T* const buf = bufs + (size_t)st * TILE;
          if (p.bulk_load && base + TILE <= p.n) {
            if (lane == 0) {
              ptx::mbar_expect_tx(&full[st], TILE_BYTES);
              ptx::bulk_g2s(buf, p.x + base, TILE_BYTES, &full[st]);
              // L2 is shared by all SMs: it does not matter WHICH CTA will own tile t + D, so the CTA that takes tile t asks for
              // it.  The HBM->L2 stream then runs D tiles ahead of the stage ring and the stage loads become L2 hits, which
              // shortens the time a stage is held for its load and leaves more of the ring to cover the prefix wait.
              const int64 ahead = base + (int64)p.prefetch_tiles * TILE;
              if (p.prefetch_tiles > 0 && ahead + TILE <= p.n) ptx::bulk_prefetch_l2(p.x + ahead, TILE_BYTES);
            }
          } else {
            const int64 remaining = p.n - base;
            for (int i = lane; i < TILE; i += 32) buf[i] = (i < remaining) ? p.x[base + i] : T(0);
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[st]);
          }
          ++nvalid;
        }
        ++jl;
      }
      if (js < nvalid) {
        const int st = (int)(js % NSTAGE);
        if (ptx::mbar_try_wait(&outready[st], (unsigned)((js / NSTAGE) & 1))) {
          long long tile = 0;
#pragma unroll
          for (int b = 0; b < NSTAGE; ++b) if (b == st) tile = tl[b];
          const int64 base = tile * TILE;
          T* const buf = bufs + (size_t)st * TILE;
          if (p.bulk_store && base + TILE <= p.n) {
            if (lane == 0) ptx::bulk_s2g(p.y + base, buf, TILE_BYTES);
          } else {
            const int64 remaining = p.n - base;
            for (int i = lane; i < TILE && i < remaining; i += 32) p.y[base + i] = buf[i];
            __syncwarp();
          }
          if (lane == 0) ptx::bulk_commit();
          ++js;
        }
      }
    }
    if (lane == 0) {
      ptx::bulk_wait_read<0>();
      scan_counter_release(p.counter);
    }
    { KB200_STATS_FLUSH(); return; }
  }

  if (warp == NWARPS + 1) {
    // ================= AGGREGATE warp =================
    int64 agg_nreq = 0;  // ATOMIC_AGG: requests handed to the round warp so far (lane 0)
    for (int64 j = 0;; ++j) {
      const int st = (int)(j % NSTAGE);
      ptx::mbar_wait(&full[st], (unsigned)((j / NSTAGE) & 1));
      const int64 cur = s_tile_id[st];
      if (cur >= p.ntiles) {
        if (lane == 0) {
          mbar_arrive(&aggready[st]);  // pass the end-of-work marker on
          if constexpr (ATOMIC_AGG) rq_push<RQ>(agg_nreq, rq_full, rq_empty, s_rq, -1);
        }
        { KB200_STATS_FLUSH(); return; }
      }
      const uint4* src = reinterpret_cast<const uint4*>(bufs + (size_t)st * TILE);
      T acc[4] = {T(0), T(0), T(0), T(0)};
      constexpr int NVEC = (int)(TILE_BYTES / 16);
#pragma unroll 4
      for (int i = lane; i < NVEC; i += 32) {
        const uint4 q = src[i];
        T e[EPV];
        memcpy(e, &q, 16);
#pragma unroll
        for (int k = 0; k < EPV; ++k) acc[k & 3] += e[k];
      }
      const T agg = warp_sum_all<T>((acc[0] + acc[1]) + (acc[2] + acc[3]));
      if (lane == 0) {
        const bool first = ROUNDS ? (cur % p.tpr == 0) : (cur == 0);  // nothing before it (in its round): the aggregate IS the inclusive prefix
#ifdef B200_SWEEP
        if constexpr (ROUNDS) {
          const int64 kq = cur / p.tpr;
          if (kq < 8192 && cur % p.tpr == 0) g_round_ts[0][kq] = ll::now_ns();
          if (kq < 8192 && cur % p.tpr == p.tpr - 1) g_round_ts[1][kq] = ll::now_ns();
        }
#endif
        ptx::st_relaxed_v2(p.desc + cur, to_bits(agg), (p.epoch << 2) | (first ? kDescIncl : kDescAgg));
        s_agg[st] = agg;
        mbar_arrive(&aggready[st]);
        if constexpr (ATOMIC_AGG) {
          const int64 kq = cur / p.tpr;
          unsigned long long* const acc = p.racc + (size_t)(kq % kRoundRing) * 2;
          const unsigned long long b = to_bits(agg);
          asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc), "l"((b & 0xffffffffull) | (1ull << 48)) : "memory");
          asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc + 1), "l"((b >> 32) | (1ull << 48)) : "memory");
          if (cur == 0) rq_push<RQ>(agg_nreq, rq_full, rq_empty, s_rq, 0);
          if (cur % p.tpr == p.tpr - 1) rq_push<RQ>(agg_nreq, rq_full, rq_empty, s_rq, kq + 1);  // the round's last tile has landed
        }
      }
    }
  }

  if (warp == NWARPS + 2) {
    // ================= LOOK-BACK warp =================
    int64 nreq = 0;  // ROUNDS without ATOMIC_AGG: requests handed to the round warp so far (lane 0)
    auto request_round = [&](int64 kk) {
      if (lane == 0) rq_push<RQ>(nreq, rq_full, rq_empty, s_rq, kk);
    };
    for (int64 j = 0;; ++j) {
      const int st = (int)(j % NSTAGE);
#ifdef B200_SWEEP
      const long long t_w2 = clock64();
#endif
      ptx::mbar_wait(&aggready[st], (unsigned)((j / NSTAGE) & 1));
#ifdef B200_SWEEP
      if (lane == 0) KB200_STAT_ADD(7, clock64() - t_w2);
#endif
      const int64 cur = s_tile_id[st];
      if (cur >= p.ntiles) {
        if constexpr (ROUNDS && !ATOMIC_AGG) request_round(-1);  // no more work for the round warp
        return;
      }
      const T agg = s_agg[st];
      T excl = T(0);
      const int64 lo = ROUNDS ? (cur / p.tpr) * p.tpr : 0;  // the look-back never leaves the tile's round
      if (cur > lo) {
        if (!(p.dbg_flags & 1)) excl = lookback_sum<T, LBW>(p.desc, cur, p.epoch, lane, p.spin_sleep_ns, p.dbg_flags & 8 KB200_STATS_PASS, lo);
        if (lane == 0) ptx::st_relaxed_v2(p.desc + cur, to_bits((T)(excl + agg)), (p.epoch << 2) | kDescIncl);
      }
      if constexpr (ROUNDS && !ATOMIC_AGG) {
        // the last tile of a round owns the round aggregate: store it into every rank's mailbox (NVLink), own rank included
        if (cur - lo == p.tpr - 1 && lane < p.world) {
          unsigned long long w0, w1;
          ll::pack((T)(excl + agg), p.rtag_base + (unsigned)(cur / p.tpr), w0, w1);
          ll::st_sys(p.peer_mbox[lane] + (size_t)((cur / p.tpr) % kRoundRing) * kRoundMaxWorld * 2 + 2 * p.rank, w0, w1);
#ifdef B200_SWEEP
          if (lane == 0 && cur / p.tpr < 8192) g_round_ts[2][cur / p.tpr] = ll::now_ns();
#endif
        }
        if (cur == 0) request_round(0);
        if (cur - lo == p.tpr - 1) request_round(cur / p.tpr + 1);  // this GPU's part of round k is known: the base of round k+1 can be built
      }
      if (lane == 0) {
        s_prefix[st] = excl;
        if (!ROUNDS && cur == p.ntiles - 1) {
          const T total = excl + agg;
          if (p.total0) *p.total0 = total;
          if (p.total1) *p.total1 = total;
        }
        mbar_arrive(&prefready[st]);
      }
    }
  }

  // ================= compute warps =================
  const T seed = scan_seed(p);
  for (int64 j = 0;; ++j) {
    const int st = (int)(j % NSTAGE);
    const unsigned par = (unsigned)((j / NSTAGE) & 1);
#ifdef B200_SWEEP
    const long long t_w0 = clock64();
#endif
    ptx::mbar_wait(&full[st], par);
#ifdef B200_SWEEP
    if (tid == 0) { KB200_STAT_ADD(5, clock64() - t_w0); KB200_STAT_ADD(6, 1); }
#endif
    const int64 cur = s_tile_id[st];
    if (cur >= p.ntiles) break;
    unsigned long long rd0 = 0, rd1 = 0;
    if constexpr (ROUNDS) {
      // the round's base is usually published long before this tile needs it: request it now, look at it after the local work
      if (tid == 0) ll::ld_gpu(p.rdesc + (size_t)((cur / p.tpr) % kRoundRing) * 4, rd0, rd1);
    }
    T* const buf = bufs + (size_t)st * TILE;
    T v[ITEMS];
    {
      const uint4* src = reinterpret_cast<const uint4*>(buf + (size_t)tid * ITEMS);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const uint4 q = src[k];
        memcpy(&v[k * EPV], &q, 16);
      }
    }
    T tsum = T(0);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) tsum += v[k];
    const T tincl = warp_incl_scan(tsum, lane);
    if (lane == 31) s_warp[warp] = tincl;
    if constexpr (ROUNDS) {
      // the tile parks HERE until its round's base is published (not in the look-back warp: that one keeps publishing
      // inclusive prefixes and round aggregates for the CTA's later tiles meanwhile)
      if (tid == 0) {
#ifdef B200_SWEEP
        const long long t_rb = clock64();
#endif
        const unsigned rtag = p.rtag_base + (unsigned)(cur / p.tpr);
        s_base = ll::ok(rd0, rd1, rtag) ? ll::unpack<T>(rd0, rd1)
                                        : ll::wait_value<T, false>(p.rdesc + (size_t)((cur / p.tpr) % kRoundRing) * 4, rtag, p.timeout_ns, p.err, 0xD4000000u);
#ifdef B200_SWEEP
        KB200_STAT_ADD(8, clock64() - t_rb);
        if (!ll::ok(rd0, rd1, rtag)) KB200_STAT_ADD(11, 1);
#endif
      }
    }
    named_bar_sync(1, CBLOCK);
    T woff = T(0);  // exclusive offset of this warp inside the tile: every warp folds the <=32 warp totals itself
    {
      const T w = lane < NWARPS ? s_warp[lane] : T(0);
      const T wi = warp_incl_scan(w, lane);
      woff = ::kb200::Impl::shfl_idx((T)(wi - w), warp);
    }
#ifdef B200_SWEEP
    const long long t_w1 = clock64();
#endif
    ptx::mbar_wait(&prefready[st], par);
#ifdef B200_SWEEP
    if (tid == 0) KB200_STAT_ADD(4, clock64() - t_w1);
#endif
    T run = seed + s_prefix[st] + woff + (tincl - tsum);
    if constexpr (ROUNDS) run += s_base;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const T in = v[k];
      if (INCLUSIVE) { run += in; v[k] = run; } else { v[k] = run; run += in; }
    }
    {
      uint4* dst = reinterpret_cast<uint4*>(buf + (size_t)tid * ITEMS);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        uint4 q;
        memcpy(&q, &v[k * EPV], 16);
        dst[k] = q;
      }
    }
    ptx::fence_proxy_async_smem();
    named_bar_sync(1, CBLOCK);  // also orders the s_warp / s_base reads above before the next tile's writes
    if (tid == 0) mbar_arrive(&outready[st]);
  }
  KB200_STATS_FLUSH();
}

}  // namespace Impl
}  // namespace kb200
#ifdef B200_SWEEP
#include "ScanContigSweep.hpp"
#endif
namespace kb200 {
namespace Impl {

// what a launch of the block-cyclic distributed scan needs from the communicator (csrc/comm.cu)
struct RoundPeers {
  int rank = 0, world = 1;
  unsigned long long* racc = nullptr;   // this GPU, [kRoundRing][2], zero between launches
  unsigned long long* rdesc = nullptr;  // this GPU, [kRoundRing][4]
  unsigned long long* mbox = nullptr;   // this GPU, [kRoundRing][kRoundMaxWorld][2]
  unsigned long long* peer_mbox[kRoundMaxWorld] = {};
  unsigned rtag_base = 1;
  unsigned* err = nullptr;
};

// WS = 2: the warp-specialised kernel above (16-byte aligned Views); 0: the uniform kernel; 1/3/4: sweep-only variants
template <class T, int BLOCK, int NV, int NBUF, int LBW, bool INCLUSIVE, int WS = 0>
struct ContigScanLaunch {
  static constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  static constexpr int TILE = BLOCK * ITEMS;
  static constexpr size_t SMEM = (size_t)NBUF * TILE * sizeof(T);
  static constexpr int THREADS = WS == 4 ? BLOCK + 96 + 32 * NBUF : WS == 3 ? BLOCK + 128 : (WS == 2 ? BLOCK + 96 : (WS == 1 ? BLOCK + 32 : BLOCK));

  static auto kernel() {
#ifdef B200_SWEEP
    if constexpr (WS == 4) return contig_scan_ws4_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
    else if constexpr (WS == 3) return contig_scan_ws3_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
    else if constexpr (WS == 1) return contig_scan_ws_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
    else
#endif
    if constexpr (WS == 2) return contig_scan_ws2_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
    else return contig_scan_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
  }
  static auto rounds_kernel() { return contig_scan_ws2_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE, true>; }
  static int resident_blocks_per_sm() {
    static PerDeviceInt cache;  // the shared-memory opt-in below is per device
    int& cached = cache.here();
    if (cached == 0) {
      auto k = kernel();
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, THREADS, SMEM);
      cached = nb > 0 ? nb : 1;
    }
    return cached;
  }

  // One rank of a block-cyclic distributed scan: x/y = this rank's rounds back to back (n elements), every rank runs
  // `nrounds` rounds of `tpr` tiles (tiles past n are empty but still take part in the exchange).  16-byte aligned Views.
  static int run_rounds(b200_instance* inst, const RoundPeers& peers, int64 tpr, int64 nrounds, const T* x, T* y, int64 n, T* total_host, T* total_dev,
                        int prefetch_tiles = 0) {
    static_assert(WS == 2, "the rounds kernel is the warp-specialised one");
    HostRuntime rt(inst);
    int rc;
    if (nrounds <= 0) {
      if (total_dev && (rc = b200_memset_async(inst, total_dev, 0, sizeof(T)))) return rc;
      if (total_host) {
        if ((rc = rt.fence("kb200::parallel_scan (empty)"))) return rc;
        *total_host = T(0);
      }
      return 0;
    }
    static PerDeviceInt cache;
    int& bps_cached = cache.here();
    if (bps_cached == 0) {
      auto k = rounds_kernel();
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, THREADS + 32, SMEM);
      bps_cached = nb > 0 ? nb : 1;
    }
    const int64 ntiles = nrounds * tpr;
    const int64 max_grid = (int64)rt.sm_count() * bps_cached;
    const int grid = (int)(ntiles < max_grid ? ntiles : max_grid);
    // live rounds <= 1 + grid * NBUF / tpr must fit the descriptor ring
    if ((int64)grid * NBUF / tpr + 2 + kRoundLookback > kRoundRing) return b200_report_error(B200_EINVAL, "kb200::parallel_scan (rounds): tiles per round too small for the round ring");
    ScanContigParams<T> p;
    memset(&p, 0, sizeof p);
    p.x = x; p.y = y; p.n = n; p.ntiles = ntiles; p.seed = T(0); p.seed_count = 0;
    void* desc = nullptr;
    if ((rc = b200_scratch_get(inst, B200_SCRATCH_SCAN_DESC, (size_t)ntiles * sizeof(ScanDesc16), &desc, nullptr))) return rc;
    p.desc = reinterpret_cast<ScanDesc16*>(desc);
    uint64_t epoch = 0, cbase = 0;
    if ((rc = b200_scan_begin(inst, (uint64_t)ntiles + (uint64_t)grid, &epoch, &cbase, &p.counter))) return rc;
    p.epoch = epoch; p.counter_base = cbase;
    void *slot_dev = nullptr, *slot_host = nullptr, *unused_p = nullptr;
    unsigned* unused_t = nullptr;
    if (total_host && (rc = rt.reduce_scratch(0, sizeof(T), true, &unused_p, &unused_t, &slot_dev, &slot_host))) return rc;
    p.total0 = total_host ? reinterpret_cast<T*>(slot_dev) : total_dev;
    p.total1 = total_host ? total_dev : nullptr;
    p.bulk_load = 1; p.bulk_store = 1;
    p.prefetch_tiles = prefetch_tiles < 0 ? grid : prefetch_tiles;  // auto: one wave of CTAs ahead
    p.tpr = tpr; p.rank = peers.rank; p.world = peers.world; p.rtag_base = peers.rtag_base;
    p.rdesc = peers.rdesc; p.mbox = peers.mbox; p.racc = peers.racc;
    if (tpr > 32768) return b200_report_error(B200_EINVAL, "kb200::parallel_scan (rounds): at most 32768 tiles per round");
    for (int q = 0; q < peers.world; ++q) p.peer_mbox[q] = peers.peer_mbox[q];
    p.err = peers.err;
    p.timeout_ns = 20ull * 1000000000ull;
    rounds_kernel()<<<grid, THREADS + 32, SMEM, rt.stream()>>>(p);
    if ((rc = rt.check_launch("kb200::contig_scan_ws2_kernel (rounds)"))) return rc;
    if (total_host) {
      if ((rc = rt.fence("kb200::parallel_scan: fence to hand the total to the host"))) return rc;
      memcpy(total_host, slot_host, sizeof(T));
    }
    return 0;
  }

  static int run(b200_instance* inst, const T* x, T* y, int64 n, T seed, const T* seed_dev, T* total_host, T* total_dev,
                 int blocks_per_sm_cap = 0, int spin_sleep_ns = 0, int dbg_flags = 0, int seed_count = 1, int prefetch_tiles = 0) {
    HostRuntime rt(inst);
    int rc;
    if (n == 0) {  // empty range: total = identity, nothing written
      if (total_dev && (rc = b200_memset_async(inst, total_dev, 0, sizeof(T)))) return rc;
      if (total_host) {
        if ((rc = rt.fence("kb200::parallel_scan (empty)"))) return rc;
        *total_host = T(0);
      }
      return 0;
    }
    int bps = resident_blocks_per_sm();
    if (blocks_per_sm_cap > 0 && blocks_per_sm_cap < bps) bps = blocks_per_sm_cap;
    const int64 ntiles = (n + TILE - 1) / TILE;
    const int64 max_grid = (int64)rt.sm_count() * bps;
    const int grid = (int)(ntiles < max_grid ? ntiles : max_grid);

    ScanContigParams<T> p;
    p.x = x; p.y = y; p.n = n; p.ntiles = ntiles; p.seed = seed; p.seed_dev = seed_dev; p.seed_count = seed_count;
    void* desc = nullptr;
    if ((rc = b200_scratch_get(inst, B200_SCRATCH_SCAN_DESC, (size_t)ntiles * sizeof(ScanDesc16), &desc, nullptr))) return rc;
    p.desc = reinterpret_cast<ScanDesc16*>(desc);
    uint64_t epoch = 0, cbase = 0;
    // every CTA takes exactly one id past the end before it stops: reserve ntiles + grid ids
    if ((rc = b200_scan_begin(inst, (uint64_t)ntiles + (uint64_t)grid, &epoch, &cbase, &p.counter))) return rc;
    p.epoch = epoch; p.counter_base = cbase;
    void *slot_dev = nullptr, *slot_host = nullptr, *unused_p = nullptr;
    unsigned* unused_t = nullptr;
    if (total_host && (rc = rt.reduce_scratch(0, sizeof(T), true, &unused_p, &unused_t, &slot_dev, &slot_host))) return rc;
    p.total0 = total_host ? reinterpret_cast<T*>(slot_dev) : total_dev;
    p.total1 = total_host ? total_dev : nullptr;
    p.bulk_load = (reinterpret_cast<uintptr_t>(x) % 16 == 0);
    p.bulk_store = (reinterpret_cast<uintptr_t>(y) % 16 == 0);
    p.spin_sleep_ns = spin_sleep_ns;
    p.dbg_flags = dbg_flags;
    // auto: one wave of CTAs ahead (B200, 2^30 int64, profiles/r02_scan_prefetch_probe.log: off 6.20, 148 tiles 6.85, 222-600 tiles
    // 6.97-7.00, 888 tiles 6.73, >= 1776 tiles (33 MB) 4.9 TB/s -- the prefetched lines are evicted before they are used)
    p.prefetch_tiles = prefetch_tiles < 0 ? grid : prefetch_tiles;
    kernel()<<<grid, THREADS, SMEM, rt.stream()>>>(p);
    if ((rc = rt.check_launch("kb200::contig_scan_kernel"))) return rc;
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
