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
  const T* seed_dev;  // if non-null, read instead of `seed`
  ScanDesc16* desc;
  unsigned long long epoch;
  unsigned long long* counter;
  unsigned long long counter_base;
  T* total0;
  T* total1;
  int bulk_load, bulk_store;  // 16-byte alignment of x / y allows the TMA path
};

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
template <class T, int LBW>
KB200_DEVICE_FUNCTION T lookback_sum(const ScanDesc16* desc, int64 tile, unsigned long long epoch, int lane) {
  T excl = T(0);
  int64 wbase = tile - 1;
  while (true) {
    unsigned long long pay[LBW], st[LBW];
#pragma unroll
    for (int j = 0; j < LBW; ++j) {
      const int64 idx = wbase - ((int64)lane * LBW + j);
      if (idx >= 0) {
        ptx::ld_relaxed_v2(desc + idx, pay[j], st[j]);
      } else {
        pay[j] = 0;
        st[j] = (epoch << 2) | kDescIncl;  // before the first tile: inclusive prefix = identity
      }
    }
    T part = T(0);
    int state = 0;  // 0: only aggregates so far, 1: ended on an inclusive prefix, 2: hit a descriptor not yet published
#pragma unroll
    for (int j = 0; j < LBW; ++j) {
      if (state == 0) {
        if ((st[j] >> 2) != epoch) {
          state = 2;
        } else {
          part += from_bits<T>(pay[j]);
          if ((st[j] & 3ull) == kDescIncl) state = 1;
        }
      }
    }
    const unsigned term = __ballot_sync(kFullMask, state == 1);
    const unsigned inval = __ballot_sync(kFullMask, state == 2);
    const int first_term = term ? (__ffs(term) - 1) : 32;
    const unsigned needed = first_term >= 31 ? kFullMask : ((2u << first_term) - 1u);
    if (inval & needed) continue;  // a predecessor has not published yet: poll again
    excl += warp_sum_all<T>(lane <= first_term ? part : T(0));
    if (term) return excl;
    wbase -= 32 * LBW;
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
  const T seed = p.seed_dev ? *p.seed_dev : p.seed;

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
        excl = lookback_sum<T, LBW>(p.desc, cur, p.epoch, lane);
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
  if (tid == 0) ptx::bulk_wait_read<0>();  // shared memory must outlive the last bulk store's reads
}

template <class T, int BLOCK, int NV, int NBUF, int LBW, bool INCLUSIVE>
struct ContigScanLaunch {
  static constexpr int ITEMS = NV * 16 / (int)sizeof(T);
  static constexpr int TILE = BLOCK * ITEMS;
  static constexpr size_t SMEM = (size_t)NBUF * TILE * sizeof(T);

  static int resident_blocks_per_sm() {
    static int cached = 0;
    if (cached == 0) {
      auto k = contig_scan_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, BLOCK, SMEM);
      cached = nb > 0 ? nb : 1;
    }
    return cached;
  }

  static int run(b200_instance* inst, const T* x, T* y, int64 n, T seed, const T* seed_dev, T* total_host, T* total_dev,
                 int blocks_per_sm_cap = 0) {
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
    p.x = x; p.y = y; p.n = n; p.ntiles = ntiles; p.seed = seed; p.seed_dev = seed_dev;
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
    contig_scan_kernel<T, BLOCK, NV, NBUF, LBW, INCLUSIVE><<<grid, BLOCK, SMEM, rt.stream()>>>(p);
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
