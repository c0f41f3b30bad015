// kb200/impl/ScanContigSweep.hpp -- scan kernel variants that were measured and NOT shipped (tools/sweep.py,
// tools/scan_probe.py; profiles/r01_scan_probe_v*.log).  Compiled only into libkokkos_b200_sweep.so (-DB200_SWEEP).
#ifndef KB200_IMPL_SCANCONTIGSWEEP_HPP
#define KB200_IMPL_SCANCONTIGSWEEP_HPP
#ifndef B200_SWEEP
#error "sweep-only header"
#endif

namespace kb200 {
namespace Impl {

// ---------------------------------------------------------------------------------------------
// Warp-specialised variant (first warp-specialised version; superseded by ws2).
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

}  // namespace Impl
}  // namespace kb200
#endif
