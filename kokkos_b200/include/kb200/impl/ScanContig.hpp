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
// accumulated in registers, flushed once per warp at exit: per-event global atomics perturb the kernel badly
struct ScanStats {
  unsigned long long v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  KB200_DEVICE_FUNCTION void flush() {
    for (int i = 0; i < 8; ++i) if (v[i]) atomicAdd(&g_scan_stats[i], v[i]);
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
};

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
KB200_DEVICE_FUNCTION unsigned long long to_bits(T v) {
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
KB200_DEVICE_FUNCTION T from_bits(unsigned long long b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}

template <class T>
KB200_DEVICE_FUNCTION T warp_sum_all(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
  return v;
}
template <class T>
KB200_DEVICE_FUNCTION T warp_incl_scan(T v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T u = shfl_up(v, d);
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
KB200_DEVICE_FUNCTION T lookback_sum(const ScanDesc16* desc, int64 tile, unsigned long long epoch, int lane, int sleep_ns, int weak KB200_STATS_ARG) {
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
      if (idx >= 0) {
        ld_desc(desc + idx, pay[j], st[j], weak);
      } else {
        pay[j] = 0;
        st[j] = (epoch << 2) | kDescIncl;  // before the first tile: inclusive prefix = identity
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
      const T agg = shfl_idx(wi, NWARPS - 1);
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

// ---------------------------------------------------------------------------------------------
// Warp-specialised variant (the shipped one for 16-byte aligned Views).
//
// The kernel above keeps every latency on the CTA's critical path: the tile-id atomic's round trip,
// the wait for the previous bulk store to drain, the TMA issue and the look-back all happen between
// barriers that the 8 compute warps sit in (first B200 sweep: 2.99 TB/s, profiles/r01_sweep_v1.log).
// Here one extra warp is the DMA engine driver: it takes tile ids, issues the bulk loads NSTAGE
// tiles ahead, and issues the bulk stores when the compute warps hand a finished stage back; all
// hand-offs are mbarriers (full[s]: data landed; outready[s]: results are in smem).  The compute
// warps only ever wait for (a) a tile that was requested NSTAGE tiles ago and (b) the look-back.
// The look-back window is 32*LBW descriptors per step, wide enough that the distance to the nearest
// resolved predecessor (~ tile arrival rate x resolution latency) fits in one or two steps.
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

template <class T, int CBLOCK, int NV, int NSTAGE, int LBW, bool INCLUSIVE>
__global__ void __launch_bounds__(CBLOCK + 32) contig_scan_ws_kernel(const ScanContigParams<T> p) {
  static_assert(NV % 2 == 1, "odd vector count keeps blocked smem accesses conflict free");
  constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  constexpr int EPV = 16 / (int)sizeof(T);
  constexpr int TILE = CBLOCK * ITEMS;
  constexpr unsigned TILE_BYTES = TILE * sizeof(T);
  constexpr int NWARPS = CBLOCK / 32;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* const bufs = reinterpret_cast<T*>(smem_raw);
  __shared__ __align__(8) unsigned long long full[NSTAGE];
  __shared__ __align__(8) unsigned long long outready[NSTAGE];
  __shared__ int64 s_tile_id[NSTAGE];
  __shared__ T s_warp[32];
  __shared__ T s_tile_prefix;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KB200_STATS_DECL;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NSTAGE; ++b) { ptx::mbar_init(&full[b], 1); ptx::mbar_init(&outready[b], 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == NWARPS) {
    // ================= DMA warp =================
    int64 jl = 0, js = 0, nvalid = 0;
    int64 tl[NSTAGE];  // tile id held by each stage
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
          if (lane == 0) mbar_arrive(&full[st]);  // wake the consumers on the end-of-work marker
          more = false;
        } else {
          const int64 base = tile * TILE;
          T* const buf = bufs + (size_t)st * TILE;
          if (p.bulk_load && base + TILE <= p.n) {
            if (lane == 0) {
              ptx::mbar_expect_tx(&full[st], TILE_BYTES);
              ptx::bulk_g2s(buf, p.x + base, TILE_BYTES, &full[st]);
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
          if (lane == 0) ptx::bulk_commit();  // one group per stage hand-back keeps the drain accounting uniform
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

  // ================= compute warps =================
  const T seed = scan_seed(p);
  for (int64 j = 0;; ++j) {
    const int st = (int)(j % NSTAGE);
    ptx::mbar_wait(&full[st], (unsigned)((j / NSTAGE) & 1));
    const int64 cur = s_tile_id[st];
    if (cur >= p.ntiles) break;
    T* const buf = bufs + (size_t)st * TILE;
    if (p.dbg_flags & 2) {  // experiment: hand the stage straight back (measures the bulk-copy pipeline alone)
      ptx::fence_proxy_async_smem();
      named_bar_sync(1, CBLOCK);
      if (tid == 0) mbar_arrive(&outready[st]);
      continue;
    }
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
    named_bar_sync(1, CBLOCK);
    if (warp == 0) {
      const T w = lane < NWARPS ? s_warp[lane] : T(0);
      const T wi = warp_incl_scan(w, lane);
      if (lane < NWARPS) s_warp[lane] = wi - w;
      const T agg = shfl_idx(wi, NWARPS - 1);
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
    named_bar_sync(1, CBLOCK);
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
    ptx::fence_proxy_async_smem();
    named_bar_sync(1, CBLOCK);
    if (tid == 0) mbar_arrive(&outready[st]);
  }
  KB200_STATS_FLUSH();
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
template <class T, int CBLOCK, int NV, int NSTAGE, int LBW, bool INCLUSIVE>
__global__ void __launch_bounds__(CBLOCK + 96) contig_scan_ws2_kernel(const ScanContigParams<T> p) {
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

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KB200_STATS_DECL;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NSTAGE; ++b) {
      ptx::mbar_init(&full[b], 1); ptx::mbar_init(&aggready[b], 1);
      ptx::mbar_init(&prefready[b], 1); ptx::mbar_init(&outready[b], 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

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
    for (int64 j = 0;; ++j) {
      const int st = (int)(j % NSTAGE);
      ptx::mbar_wait(&full[st], (unsigned)((j / NSTAGE) & 1));
      const int64 cur = s_tile_id[st];
      if (cur >= p.ntiles) {
        if (lane == 0) mbar_arrive(&aggready[st]);  // pass the end-of-work marker on
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
        ptx::st_relaxed_v2(p.desc + cur, to_bits(agg), (p.epoch << 2) | (cur == 0 ? kDescIncl : kDescAgg));
        s_agg[st] = agg;
        mbar_arrive(&aggready[st]);
      }
    }
  }

  if (warp == NWARPS + 2) {
    // ================= LOOK-BACK warp =================
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
      if (cur >= p.ntiles) return;
      const T agg = s_agg[st];
      T excl = T(0);
      if (cur > 0) {
        if (!(p.dbg_flags & 1)) excl = lookback_sum<T, LBW>(p.desc, cur, p.epoch, lane, p.spin_sleep_ns, p.dbg_flags & 8 KB200_STATS_PASS);
        if (lane == 0) ptx::st_relaxed_v2(p.desc + cur, to_bits((T)(excl + agg)), (p.epoch << 2) | kDescIncl);
      }
      if (lane == 0) {
        s_prefix[st] = excl;
        if (cur == p.ntiles - 1) {
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
    named_bar_sync(1, CBLOCK);
    T woff = T(0);  // exclusive offset of this warp inside the tile: every warp folds the <=32 warp totals itself
    {
      const T w = lane < NWARPS ? s_warp[lane] : T(0);
      const T wi = warp_incl_scan(w, lane);
      woff = shfl_idx((T)(wi - w), warp);
    }
#ifdef B200_SWEEP
    const long long t_w1 = clock64();
#endif
    ptx::mbar_wait(&prefready[st], par);
#ifdef B200_SWEEP
    if (tid == 0) KB200_STAT_ADD(4, clock64() - t_w1);
#endif
    T run = seed + s_prefix[st] + woff + (tincl - tsum);
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
    named_bar_sync(1, CBLOCK);  // also orders the s_warp reads above before the next tile's writes
    if (tid == 0) mbar_arrive(&outready[st]);
  }
  KB200_STATS_FLUSH();
}

// ---------------------------------------------------------------------------------------------
// ws3 = ws2 with the DMA driver split in two warps.  Measured (profiles/r01_scan_probe_v3.log): one DMA thread
// per CTA serialises tile-id atomic round trip -> wait for the previous store to drain -> issue, ~1.5 us per
// tile, which caps a CTA at ~24 GB/s; configurations with one CTA per SM collapsed to 3 TB/s.  Here the LOAD warp
// keeps one tile-id atomic in flight ahead of its use and only waits for a free stage; the STORE warp issues
// the bulk store, waits for ITS reads, and recycles the stage (empty[s]).
template <class T, int CBLOCK, int NV, int NSTAGE, int LBW, bool INCLUSIVE>
__global__ void __launch_bounds__(CBLOCK + 128) contig_scan_ws3_kernel(const ScanContigParams<T> p) {
  static_assert(NV % 2 == 1, "odd vector count keeps blocked smem accesses conflict free");
  constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  constexpr int EPV = 16 / (int)sizeof(T);
  constexpr int TILE = CBLOCK * ITEMS;
  constexpr unsigned TILE_BYTES = TILE * sizeof(T);
  constexpr int NWARPS = CBLOCK / 32;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* const bufs = reinterpret_cast<T*>(smem_raw);
  __shared__ __align__(8) unsigned long long full[NSTAGE], aggready[NSTAGE], prefready[NSTAGE], outready[NSTAGE], empty[NSTAGE];
  __shared__ int64 s_tile_id[NSTAGE];
  __shared__ T s_agg[NSTAGE];
  __shared__ T s_prefix[NSTAGE];
  __shared__ T s_warp[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KB200_STATS_DECL;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NSTAGE; ++b) {
      ptx::mbar_init(&full[b], 1); ptx::mbar_init(&aggready[b], 1);
      ptx::mbar_init(&prefready[b], 1); ptx::mbar_init(&outready[b], 1); ptx::mbar_init(&empty[b], 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == NWARPS) {
    // ================= LOAD warp: tile ids (one atomic ahead) + bulk loads =================
    long long next_tile = 0;
    if (lane == 0) next_tile = (long long)(atomicAdd(p.counter, 1ull) - p.counter_base);
    for (int64 jl = 0;; ++jl) {
      const int st = (int)(jl % NSTAGE);
      if (jl >= NSTAGE) ptx::mbar_wait(&empty[st], (unsigned)(((jl / NSTAGE) - 1) & 1));  // freed by the STORE warp
      const long long tile = __shfl_sync(kFullMask, next_tile, 0);
      if (lane == 0) s_tile_id[st] = tile;
      if (tile >= p.ntiles) {
        __syncwarp();
        if (lane == 0) { mbar_arrive(&full[st]); scan_counter_release(p.counter); }
        { KB200_STATS_FLUSH(); return; }
      }
      if (lane == 0) next_tile = (long long)(atomicAdd(p.counter, 1ull) - p.counter_base);  // round trip overlaps the load
      const int64 base = tile * TILE;
      T* const buf = bufs + (size_t)st * TILE;
      if (p.bulk_load && base + TILE <= p.n) {
        if (lane == 0) {
          ptx::mbar_expect_tx(&full[st], TILE_BYTES);
          ptx::bulk_g2s(buf, p.x + base, TILE_BYTES, &full[st]);
        }
      } else {
        const int64 remaining = p.n - base;
        for (int i = lane; i < TILE; i += 32) buf[i] = (i < remaining) ? p.x[base + i] : T(0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
      }
    }
  }

  if (warp == NWARPS + 3) {
    // ================= STORE warp: bulk stores + stage recycling =================
    for (int64 js = 0;; ++js) {
      const int st = (int)(js % NSTAGE);
      const unsigned par = (unsigned)((js / NSTAGE) & 1);
      ptx::mbar_wait(&full[st], par);
      const int64 tile = s_tile_id[st];
      if (tile >= p.ntiles) {
        if (lane == 0) ptx::bulk_wait<0>();
        { KB200_STATS_FLUSH(); return; }
      }
      ptx::mbar_wait(&outready[st], par);
      const int64 base = tile * TILE;
      T* const buf = bufs + (size_t)st * TILE;
      if (p.bulk_store && base + TILE <= p.n) {
        if (lane == 0) {
          ptx::bulk_s2g(p.y + base, buf, TILE_BYTES);
          ptx::bulk_commit();
          ptx::bulk_wait_read<0>();  // the stage may be overwritten once its bytes have been read
        }
      } else {
        const int64 remaining = p.n - base;
        for (int i = lane; i < TILE && i < remaining; i += 32) p.y[base + i] = buf[i];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
  }

  if (warp == NWARPS + 1) {
    // ================= AGGREGATE warp =================
    for (int64 j = 0;; ++j) {
      const int st = (int)(j % NSTAGE);
      ptx::mbar_wait(&full[st], (unsigned)((j / NSTAGE) & 1));
      const int64 cur = s_tile_id[st];
      if (cur >= p.ntiles) {
        if (lane == 0) mbar_arrive(&aggready[st]);  // pass the end-of-work marker on
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
        ptx::st_relaxed_v2(p.desc + cur, to_bits(agg), (p.epoch << 2) | (cur == 0 ? kDescIncl : kDescAgg));
        s_agg[st] = agg;
        mbar_arrive(&aggready[st]);
      }
    }
  }

  if (warp == NWARPS + 2) {
    // ================= LOOK-BACK warp =================
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
      if (cur >= p.ntiles) return;
      const T agg = s_agg[st];
      T excl = T(0);
      if (cur > 0) {
        if (!(p.dbg_flags & 1)) excl = lookback_sum<T, LBW>(p.desc, cur, p.epoch, lane, p.spin_sleep_ns, p.dbg_flags & 8 KB200_STATS_PASS);
        if (lane == 0) ptx::st_relaxed_v2(p.desc + cur, to_bits((T)(excl + agg)), (p.epoch << 2) | kDescIncl);
      }
      if (lane == 0) {
        s_prefix[st] = excl;
        if (cur == p.ntiles - 1) {
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
    named_bar_sync(1, CBLOCK);
    T woff = T(0);  // exclusive offset of this warp inside the tile: every warp folds the <=32 warp totals itself
    {
      const T w = lane < NWARPS ? s_warp[lane] : T(0);
      const T wi = warp_incl_scan(w, lane);
      woff = shfl_idx((T)(wi - w), warp);
    }
#ifdef B200_SWEEP
    const long long t_w1 = clock64();
#endif
    ptx::mbar_wait(&prefready[st], par);
#ifdef B200_SWEEP
    if (tid == 0) KB200_STAT_ADD(4, clock64() - t_w1);
#endif
    T run = seed + s_prefix[st] + woff + (tincl - tsum);
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
    named_bar_sync(1, CBLOCK);  // also orders the s_warp reads above before the next tile's writes
    if (tid == 0) mbar_arrive(&outready[st]);
  }
  KB200_STATS_FLUSH();
}

// ---------------------------------------------------------------------------------------------
// ws4 = ws3 with one LOOK-BACK warp per stage that starts when the tile id is taken, not when the tile's own
// aggregate exists.  Measured (profiles/r01_scan_probe_v5_instrumented.log): one descriptor window costs ~1.3 us
// under full HBM load and a tile needs ~2.3 windows, so a single look-back warp resolved one tile per ~3.5 us and
// the compute warps waited 1.2-3 us per tile for the prefix.  Now the walk overlaps the load of the same tile and
// the walks of the CTA's other stages.
template <class T, int CBLOCK, int NV, int NSTAGE, int LBW, bool INCLUSIVE>
__global__ void __launch_bounds__(CBLOCK + 96 + 32 * NSTAGE) contig_scan_ws4_kernel(const ScanContigParams<T> p) {
  static_assert(NV % 2 == 1, "odd vector count keeps blocked smem accesses conflict free");
  constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  constexpr int EPV = 16 / (int)sizeof(T);
  constexpr int TILE = CBLOCK * ITEMS;
  constexpr unsigned TILE_BYTES = TILE * sizeof(T);
  constexpr int NWARPS = CBLOCK / 32;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* const bufs = reinterpret_cast<T*>(smem_raw);
  __shared__ __align__(8) unsigned long long full[NSTAGE], aggready[NSTAGE], prefready[NSTAGE], outready[NSTAGE], empty[NSTAGE], idready[NSTAGE];
  __shared__ int64 s_tile_id[NSTAGE];
  __shared__ T s_agg[NSTAGE];
  __shared__ T s_prefix[NSTAGE];
  __shared__ T s_warp[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KB200_STATS_DECL;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NSTAGE; ++b) {
      ptx::mbar_init(&full[b], 1); ptx::mbar_init(&aggready[b], 1);
      ptx::mbar_init(&prefready[b], 1); ptx::mbar_init(&outready[b], 1); ptx::mbar_init(&empty[b], 1); ptx::mbar_init(&idready[b], 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == NWARPS) {
    // ================= LOAD warp: tile ids (one atomic ahead) + bulk loads =================
    long long next_tile = 0;
    int n_end = 0;
    if (lane == 0) next_tile = (long long)(atomicAdd(p.counter, 1ull) - p.counter_base);
    for (int64 jl = 0;; ++jl) {
      const int st = (int)(jl % NSTAGE);
      if (jl >= NSTAGE) ptx::mbar_wait(&empty[st], (unsigned)(((jl / NSTAGE) - 1) & 1));  // freed by the STORE warp
      const long long tile = __shfl_sync(kFullMask, next_tile, 0);
      if (lane == 0) { s_tile_id[st] = tile; mbar_arrive(&idready[st]); }  // the look-back may start now
      if (tile >= p.ntiles) {
        // end of work: every stage gets the marker once (each per-stage look-back warp must see it), no more ids are taken
        __syncwarp();
        if (lane == 0) { mbar_arrive(&full[st]); if (n_end == 0) scan_counter_release(p.counter); }
        if (++n_end == NSTAGE) { KB200_STATS_FLUSH(); return; }
        continue;
      }
      if (lane == 0) next_tile = (long long)(atomicAdd(p.counter, 1ull) - p.counter_base);  // round trip overlaps the load
      const int64 base = tile * TILE;
      T* const buf = bufs + (size_t)st * TILE;
      if (p.bulk_load && base + TILE <= p.n) {
        if (lane == 0) {
          ptx::mbar_expect_tx(&full[st], TILE_BYTES);
          ptx::bulk_g2s(buf, p.x + base, TILE_BYTES, &full[st]);
        }
      } else {
        const int64 remaining = p.n - base;
        for (int i = lane; i < TILE; i += 32) buf[i] = (i < remaining) ? p.x[base + i] : T(0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
      }
    }
  }

  if (warp == NWARPS + 2) {
    // ================= STORE warp: bulk stores + stage recycling =================
    for (int64 js = 0;; ++js) {
      const int st = (int)(js % NSTAGE);
      const unsigned par = (unsigned)((js / NSTAGE) & 1);
      ptx::mbar_wait(&full[st], par);
      const int64 tile = s_tile_id[st];
      if (tile >= p.ntiles) {
        if (lane == 0) ptx::bulk_wait<0>();
        { KB200_STATS_FLUSH(); return; }
      }
      ptx::mbar_wait(&outready[st], par);
      const int64 base = tile * TILE;
      T* const buf = bufs + (size_t)st * TILE;
      if (p.bulk_store && base + TILE <= p.n) {
        if (lane == 0) {
          ptx::bulk_s2g(p.y + base, buf, TILE_BYTES);
          ptx::bulk_commit();
          ptx::bulk_wait_read<0>();  // the stage may be overwritten once its bytes have been read
        }
      } else {
        const int64 remaining = p.n - base;
        for (int i = lane; i < TILE && i < remaining; i += 32) p.y[base + i] = buf[i];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
  }

  if (warp == NWARPS + 1) {
    // ================= AGGREGATE warp =================
    for (int64 j = 0;; ++j) {
      const int st = (int)(j % NSTAGE);
      ptx::mbar_wait(&full[st], (unsigned)((j / NSTAGE) & 1));
      const int64 cur = s_tile_id[st];
      if (cur >= p.ntiles) {
        if (lane == 0) mbar_arrive(&aggready[st]);  // pass the end-of-work marker on
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
        ptx::st_relaxed_v2(p.desc + cur, to_bits(agg), (p.epoch << 2) | (cur == 0 ? kDescIncl : kDescAgg));
        s_agg[st] = agg;
        mbar_arrive(&aggready[st]);
      }
    }
  }

  if (warp >= NWARPS + 3) {
    // ================= LOOK-BACK warps: one per stage, started as soon as the tile id is known =================
    // The exclusive prefix of a tile depends on its predecessors only, so the walk overlaps the tile's own load;
    // consecutive tiles of this CTA resolve concurrently (one warp per stage).
    const int st = warp - (NWARPS + 3);
    for (int64 k = 0;; ++k) {
      const unsigned par = (unsigned)(k & 1);
      ptx::mbar_wait(&idready[st], par);
      const int64 cur = s_tile_id[st];
      if (cur >= p.ntiles) return;
      T excl = T(0);
      if (cur > 0 && !(p.dbg_flags & 1)) excl = lookback_sum<T, LBW>(p.desc, cur, p.epoch, lane, p.spin_sleep_ns, p.dbg_flags & 8 KB200_STATS_PASS);
#ifdef B200_SWEEP
      const long long t_w2 = clock64();
#endif
      ptx::mbar_wait(&aggready[st], par);
#ifdef B200_SWEEP
      if (lane == 0) KB200_STAT_ADD(7, clock64() - t_w2);
#endif
      if (lane == 0) {
        const T agg = s_agg[st];
        if (cur > 0) ptx::st_relaxed_v2(p.desc + cur, to_bits((T)(excl + agg)), (p.epoch << 2) | kDescIncl);
        s_prefix[st] = excl;
        if (cur == p.ntiles - 1) {
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
    named_bar_sync(1, CBLOCK);
    T woff = T(0);  // exclusive offset of this warp inside the tile: every warp folds the <=32 warp totals itself
    {
      const T w = lane < NWARPS ? s_warp[lane] : T(0);
      const T wi = warp_incl_scan(w, lane);
      woff = shfl_idx((T)(wi - w), warp);
    }
#ifdef B200_SWEEP
    const long long t_w1 = clock64();
#endif
    ptx::mbar_wait(&prefready[st], par);
#ifdef B200_SWEEP
    if (tid == 0) KB200_STAT_ADD(4, clock64() - t_w1);
#endif
    T run = seed + s_prefix[st] + woff + (tincl - tsum);
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
    named_bar_sync(1, CBLOCK);  // also orders the s_warp reads above before the next tile's writes
    if (tid == 0) mbar_arrive(&outready[st]);
  }
  KB200_STATS_FLUSH();
}

// WS = true: warp-specialised kernel (BLOCK compute threads + one DMA warp); false: the uniform kernel
template <class T, int BLOCK, int NV, int NBUF, int LBW, bool INCLUSIVE, int WS = 0>
struct ContigScanLaunch {
  static constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  static constexpr int TILE = BLOCK * ITEMS;
  static constexpr size_t SMEM = (size_t)NBUF * TILE * sizeof(T);
  static constexpr int THREADS = WS == 4 ? BLOCK + 96 + 32 * NBUF : WS == 3 ? BLOCK + 128 : (WS == 2 ? BLOCK + 96 : (WS == 1 ? BLOCK + 32 : BLOCK));

  static auto kernel() {
    if constexpr (WS == 4) return contig_scan_ws4_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
    else if constexpr (WS == 3) return contig_scan_ws3_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
    else if constexpr (WS == 2) return contig_scan_ws2_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
    else if constexpr (WS == 1) return contig_scan_ws_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
    else return contig_scan_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
  }
  static int resident_blocks_per_sm() {
    static int cached = 0;
    if (cached == 0) {
      auto k = kernel();
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, THREADS, SMEM);
      cached = nb > 0 ? nb : 1;
    }
    return cached;
  }

  static int run(b200_instance* inst, const T* x, T* y, int64 n, T seed, const T* seed_dev, T* total_host, T* total_dev,
                 int blocks_per_sm_cap = 0, int spin_sleep_ns = 0, int dbg_flags = 0, int seed_count = 1) {
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
