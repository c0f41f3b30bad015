// kb200/impl/ScanGeneric.hpp -- parallel_scan over RangePolicy for an arbitrary Kokkos scan functor
// f(i, update, final) and an arbitrary static value_type with an associative (not necessarily commutative) join.
//
// Replaces ParallelScan / ParallelScanWithTotal<...,RangePolicy,Cuda>
// (core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:390-701,704-1047): two launches, the functor called three
// times per index, a block-wide shared-memory scan every 128 elements.
//
// One launch, functor called exactly TWICE per index (as on the OpenMP oracle):
//   pass 1  striped/coalesced: element e = j*BLOCK + tid ; c = identity ; f(i, c, false) -> contribution c
//           -> shared memory;
//   blocked each thread folds ITEMS consecutive contributions (ITEMS odd: conflict-free smem), ordered
//           warp scan of the thread totals (shuffles), one cross-warp hop;
//   chain   tile aggregate published, decoupled look-back over the predecessors (ordered fold, so
//           non-commutative joins see operands in index order), inclusive prefix published;
//   pass 2  exclusive prefixes written back to shared memory (blocked), read striped:
//           u = prefix ; f(i, u, true).
// Descriptors: value types of <= 8 bytes use ONE 16-byte descriptor {value bits, epoch<<2|state} per tile, written and
// read with a single relaxed 128-bit access (one L2 round trip per look-back window); larger value types use a status
// word (st.release / ld.acquire) + two value slots per tile.  The next tile id is fetched one tile ahead so the
// atomic's round trip overlaps pass 1.
#ifndef KB200_IMPL_SCANGENERIC_HPP
#define KB200_IMPL_SCANGENERIC_HPP

#include "Collectives.hpp"
#include "HostRuntime.hpp"
#include "Ptx.hpp"
#include "ScanContig.hpp"

namespace kb200 {
namespace Impl {

template <class V>
struct GenericScanScratch {
  ScanDesc16* desc;            // [ntiles], packed form (sizeof(V) <= 8)
  unsigned long long* status;  // [ntiles]
  V* agg;                      // [ntiles]
  V* incl;                     // [ntiles]
  unsigned long long epoch;
  unsigned long long* counter;
  unsigned long long counter_base;
};

template <class V>
KB200_DEVICE_FUNCTION void store_value(V* dst, const V& v) { *dst = v; }

// ordered inclusive warp scan: lane l ends with x0 (+) ... (+) xl
template <class Red>
KB200_DEVICE_FUNCTION void warp_incl_scan_ordered(const Red& red, typename Red::value_type& v, int lane) {
  using V = typename Red::value_type;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    V lo = ::kb200::Impl::shfl_up(v, d);
    if (lane >= d) { red.join(lo, v); v = lo; }
  }
}

// exclusive prefix of tile `tile` (>0), folded in tile order.  Warp-collective; result valid in every lane.
template <class Red>
KB200_DEVICE_FUNCTION typename Red::value_type lookback_ordered(const Red& red, const GenericScanScratch<typename Red::value_type>& s,
                                                                int64 tile, int lane) {
  using V = typename Red::value_type;
  V excl;
  red.init(excl);
  bool have = false;
  int64 wbase = tile - 1;
  while (true) {
    const int64 idx = wbase - lane;
    V val;
    red.init(val);
    int state;  // 1 = aggregate, 2 = inclusive prefix, 0 = not yet published
    if (idx >= 0) {
      const unsigned long long st = ptx::ld_acquire_u64(s.status + idx);
      state = ((st >> 2) == s.epoch) ? (int)(st & 3ull) : 0;
      if (state == 1) val = load_cg(s.agg + idx);
      if (state == 2) val = load_cg(s.incl + idx);
    } else {
      state = 2;  // before the first tile: the identity is an inclusive prefix
    }
    const unsigned term = __ballot_sync(kFullMask, state == 2);
    const unsigned inval = __ballot_sync(kFullMask, state == 0);
    const int first_term = term ? (__ffs(term) - 1) : 32;
    const unsigned needed = first_term >= 31 ? kFullMask : ((2u << first_term) - 1u);
    if (inval & needed) { __nanosleep(100); continue; }
    if (lane > first_term) red.init(val);  // older than the nearest resolved tile: not needed
    // fold lanes first_term .. 0 (older tile = higher lane = LEFT operand)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      V hi = ::kb200::Impl::shfl_down(val, d);
      if (lane + d < 32) { red.join(hi, val); val = hi; }
    }
    V window = ::kb200::Impl::shfl_idx(val, 0);
    if (have) { red.join(window, excl); }
    excl = window;
    have = true;
    if (term) return excl;
    wbase -= 32;
  }
}

// packed-descriptor form of the above (sizeof(V) <= 8): one 128-bit load per lane and window
template <class Red>
KB200_DEVICE_FUNCTION typename Red::value_type lookback_ordered_packed(const Red& red, const GenericScanScratch<typename Red::value_type>& s,
                                                                       int64 tile, int lane) {
  using V = typename Red::value_type;
  V excl;
  red.init(excl);
  bool have = false;
  int64 wbase = tile - 1;
  while (true) {
    const int64 idx = wbase - lane;
    V val;
    red.init(val);
    int state = 2;  // before the first tile: the identity is an inclusive prefix
    if (idx >= 0) {
      unsigned long long pay, st;
      ptx::ld_relaxed_v2(s.desc + idx, pay, st);
      state = ((st >> 2) == s.epoch) ? (int)(st & 3ull) : 0;
      if (state) memcpy(&val, &pay, sizeof(V));
    }
    const unsigned term = __ballot_sync(kFullMask, state == 2);
    const unsigned inval = __ballot_sync(kFullMask, state == 0);
    const int first_term = term ? (__ffs(term) - 1) : 32;
    const unsigned needed = first_term >= 31 ? kFullMask : ((2u << first_term) - 1u);
    if (inval & needed) { __nanosleep(100); continue; }
    if (lane > first_term) red.init(val);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      V hi = ::kb200::Impl::shfl_down(val, d);
      if (lane + d < 32) { red.join(hi, val); val = hi; }
    }
    V window = ::kb200::Impl::shfl_idx(val, 0);
    if (have) { red.join(window, excl); }
    excl = window;
    have = true;
    if (term) return excl;
    wbase -= 32;
  }
}
template <class V>
KB200_DEVICE_FUNCTION void publish_packed(ScanDesc16* d, const V& v, unsigned long long epoch, unsigned long long state) {
  unsigned long long pay = 0;
  memcpy(&pay, &v, sizeof(V));
  ptx::st_relaxed_v2(d, pay, (epoch << 2) | state);
}

// Software-pipelined across tiles: a CTA publishes the AGGREGATE of tile k as soon as its block scan is done and only then
// resolves tile k-1 (look-back + final functor call) -- while the global loads of tile k+1's first functor call are already
// in flight.  The look-back therefore never sits between a tile's loads and the publication successors wait for
// (measured without this: 49 barrier-stall cycles per issued instruction, 2.3 TB/s; profiles/r01_gscan_v1_ncu.txt).
template <class F, class Tag, class Index, class Red, int BLOCK, int ITEMS>
__global__ void __launch_bounds__(BLOCK)
    generic_scan_kernel(const __grid_constant__ F f, const __grid_constant__ Red red, const Index begin, const int64 n,
                        const int64 ntiles, const GenericScanScratch<typename Red::value_type> s,
                        typename Red::value_type* total0, typename Red::value_type* total1) {
  using V = typename Red::value_type;
  constexpr int TILE = BLOCK * ITEMS;
  constexpr int NWARPS = BLOCK / 32;
  constexpr bool PACKED = sizeof(V) <= 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // TILE values.  One buffer serves both tiles in flight: step (2) reads element e of the previous tile and step (3)
  // overwrites the same e with the new tile's contribution from the SAME thread (both striped), so no barrier is needed.
  V* const vals = reinterpret_cast<V*>(smem_raw);
  V* const s_warp = vals + TILE;                                         // 32 values
  V* const s_prefix = s_warp + 32;                                       // 1 value: exclusive prefix of the tile being finished
  V* const s_agg = s_prefix + 1;                                         // 1 value: aggregate of the tile published last
  __shared__ int64 s_tile[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  auto call = [&](int64 i, V& v, bool fin) {
    if constexpr (std::is_void<Tag>::value) f((Index)(begin + (Index)i), v, fin);
    else f(Tag{}, (Index)(begin + (Index)i), v, fin);
  };

  if (tid == 0) s_tile[0] = (int64)(atomicAdd(s.counter, 1ull) - s.counter_base);
  int64 prev = -1;  // tile whose final pass is still owed
  for (int it = 0;; it ^= 1) {
    __syncthreads();
    const int64 tile = s_tile[it];
    const bool have_new = tile < ntiles;
    if (!have_new && prev < 0) break;
    // next tile id, one tile ahead; every CTA takes exactly one id past the end
    if (have_new && tid == 0) s_tile[it ^ 1] = (int64)(atomicAdd(s.counter, 1ull) - s.counter_base);
    const int64 tbase = tile * TILE;

    // ---- (1) first functor call of the NEW tile: contributions into registers (loads go in flight)
    V c[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      red.init(c[j]);
      const int64 i = tbase + j * BLOCK + tid;
      if (have_new && i < n) call(i, c[j], false);
    }

    // ---- (2) finish the PREVIOUS tile: look-back, inclusive prefix, final functor call
    if (prev >= 0) {
      if (warp == 0) {
        V excl;
        red.init(excl);
        if (prev > 0) {
          if constexpr (PACKED) excl = lookback_ordered_packed(red, s, prev, lane);
          else excl = lookback_ordered(red, s, prev, lane);
          if (lane == 0) {
            V inc = excl;
            red.join(inc, *s_agg);
            if constexpr (PACKED) publish_packed(s.desc + prev, inc, s.epoch, 2ull);
            else { store_value(s.incl + prev, inc); ptx::st_release_u64(s.status + prev, (s.epoch << 2) | 2ull); }
          }
        }
        if (lane == 0) {
          *s_prefix = excl;
          if (prev == ntiles - 1 && (total0 || total1)) {
            V total = excl;
            red.join(total, *s_agg);
            // parallel_scan's total is the plain running value (no final(): Kokkos_Parallel.hpp:405-425)
            if (total0) *total0 = total;
            if (total1) *total1 = total;
          }
        }
      }
      __syncthreads();
      const V tprefix = *s_prefix;
      const int64 pbase = prev * TILE;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const int e = j * BLOCK + tid;
        const int64 i = pbase + e;
        if (i < n) {
          V u = tprefix;
          { V lp = vals[e]; red.join(u, lp); }
          call(i, u, true);
        }
      }
    }
    prev = -1;
    if (!have_new) break;  // ids are monotonic: nothing more will come

    // ---- (3) block scan of the new tile, aggregate published immediately
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) vals[j * BLOCK + tid] = c[j];
    __syncthreads();
    V loc[ITEMS];
    V tsum;
    red.init(tsum);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      loc[k] = vals[tid * ITEMS + k];
      red.join(tsum, loc[k]);
    }
    V tincl = tsum;
    warp_incl_scan_ordered(red, tincl, lane);
    if (lane == 31) s_warp[warp] = tincl;
    __syncthreads();
    if (warp == 0) {
      V w;
      red.init(w);
      if (lane < NWARPS) w = s_warp[lane];
      V wi = w;
      warp_incl_scan_ordered(red, wi, lane);
      V wex = ::kb200::Impl::shfl_up(wi, 1);  // exclusive warp prefix
      if (lane == 0) red.init(wex);
      if (lane < NWARPS) s_warp[lane] = wex;
      V agg = ::kb200::Impl::shfl_idx(wi, NWARPS - 1);
      if (lane == 0) {
        *s_agg = agg;
        const unsigned long long state = tile == 0 ? 2ull : 1ull;  // tile 0: its aggregate IS its inclusive prefix
        if constexpr (PACKED) publish_packed(s.desc + tile, agg, s.epoch, state);
        else {
          store_value((tile == 0 ? s.incl : s.agg) + tile, agg);
          ptx::st_release_u64(s.status + tile, (s.epoch << 2) | state);
        }
      }
    }
    __syncthreads();
    {  // tile-local exclusive prefix of every element, blocked write-back (the tile prefix is joined in step (2) next time)
      V run = s_warp[warp];
      V tex = ::kb200::Impl::shfl_up(tincl, 1);  // exclusive thread prefix inside the warp
      if (lane != 0) red.join(run, tex);
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        vals[tid * ITEMS + k] = run;
        red.join(run, loc[k]);
      }
    }
    prev = tile;
  }
  if (tid == 0) scan_counter_release(s.counter);
}

// BLOCK_ / ITEMS_ = 0: the shipped tile shape for this value size (tests/cxx/cases_perf.cu instantiates alternatives)
template <class Policy, class F, class Red, int BLOCK_ = 0, int ITEMS_ = 0>
struct GenericScan {
  using V = typename Red::value_type;
  using Index = typename Policy::index_type;
  using Tag = typename Policy::work_tag;
  // Large tiles: the per-tile costs (five barriers, one look-back) are what bounds this kernel -- B200 sweep at 2^30 int64
  // (profiles/r01_gscan_probe.log): 256x9 2.5, 512x13 3.8, 1024x9 4.3, 1024x13 5.07, 1024x17 5.15 TB/s.
  static constexpr int BLOCK = BLOCK_ ? BLOCK_ : (sizeof(V) <= 8 ? 1024 : (sizeof(V) <= 16 ? 512 : (sizeof(V) <= 64 ? 256 : 128)));
  static constexpr int ITEMS = ITEMS_ ? ITEMS_ : (sizeof(V) <= 8 ? 13 : (sizeof(V) <= 16 ? 9 : (sizeof(V) <= 32 ? 7 : 3)));
  static constexpr int TILE = BLOCK * ITEMS;
  static constexpr size_t SMEM = (size_t)(TILE + 34) * sizeof(V);
  static_assert(SMEM <= 200 * 1024, "parallel_scan value_type too large for the shared-memory tile");

  static int run(const Policy& policy, const F& f, const Red& red, V* total_host, V* total_dev) {
    b200_instance* inst = policy.space().impl_instance();
    HostRuntime rt(inst);
    int rc;
    int64 n = (int64)(policy.end() - policy.begin());
    if (n <= 0) {  // empty range: the total is the identity
      V t;
      red.init(t);
      if (total_host) *total_host = t;
      if (total_dev && (rc = b200_memcpy_h2d_async(inst, total_dev, &t, sizeof(V)))) return rc;
      if (total_dev) return rt.fence("kb200::parallel_scan (empty)");
      return 0;
    }
    auto k = generic_scan_kernel<F, Tag, Index, Red, BLOCK, ITEMS>;
    static PerDeviceInt cache;  // the shared-memory opt-in below is per device
    int& bps = cache.here();
    if (bps == 0) {
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, BLOCK, SMEM);
      if (bps < 1) bps = 1;
    }
    const int64 ntiles = (n + TILE - 1) / TILE;
    const int64 cap = (int64)rt.sm_count() * bps;
    const int grid = (int)(ntiles < cap ? ntiles : cap);

    GenericScanScratch<V> s;
    void *st = nullptr, *vals = nullptr, *desc = nullptr;
    s.desc = nullptr; s.status = nullptr; s.agg = nullptr; s.incl = nullptr;
    if constexpr (sizeof(V) <= 8) {
      if ((rc = b200_scratch_get(inst, B200_SCRATCH_SCAN_DESC, (size_t)ntiles * sizeof(ScanDesc16), &desc, nullptr))) return rc;
      s.desc = reinterpret_cast<ScanDesc16*>(desc);
    } else {
      if ((rc = b200_scratch_get(inst, B200_SCRATCH_SCAN_STATUS, (size_t)ntiles * 8, &st, nullptr))) return rc;
      if ((rc = b200_scratch_get(inst, B200_SCRATCH_SCAN_VALUES, (size_t)ntiles * 2 * sizeof(V) + 256, &vals, nullptr))) return rc;
      s.status = reinterpret_cast<unsigned long long*>(st);
      s.agg = reinterpret_cast<V*>(vals);
      s.incl = s.agg + ntiles;
    }
    uint64_t epoch = 0, cbase = 0;
    if ((rc = b200_scan_begin(inst, (uint64_t)ntiles + (uint64_t)grid, &epoch, &cbase, &s.counter))) return rc;
    s.epoch = epoch;
    s.counter_base = cbase;
    void *slot_dev = nullptr, *slot_host = nullptr, *unused_p = nullptr;
    unsigned* unused_t = nullptr;
    if (total_host && (rc = rt.reduce_scratch(0, sizeof(V), true, &unused_p, &unused_t, &slot_dev, &slot_host))) return rc;
    V* t0 = total_host ? reinterpret_cast<V*>(slot_dev) : total_dev;
    V* t1 = total_host ? total_dev : nullptr;
    k<<<grid, BLOCK, SMEM, rt.stream()>>>(f, red, policy.begin(), n, ntiles, s, t0, t1);
    if ((rc = rt.check_launch("kb200::generic_scan_kernel"))) return rc;
    if (total_host) {
      if ((rc = rt.fence("kb200::parallel_scan: fence to hand the total to the host"))) return rc;
      memcpy(total_host, slot_host, sizeof(V));
    }
    return 0;
  }
};

}  // namespace Impl
}  // namespace kb200
#endif
