// kb200/impl/ScanGeneric.hpp -- parallel_scan over RangePolicy for an arbitrary Kokkos scan functor
// f(i, update, final) and an arbitrary static value_type with an associative (not necessarily commutative) join.
//
// Replaces ParallelScan / ParallelScanWithTotal<...,RangePolicy,Cuda>
// (core/src/Cuda/Kokkos_Cuda_Parallel_Range.hpp:390-701,704-1047): two launches, the functor called three
// times per index, a block-wide shared-memory scan every 128 elements (measured on B200: 1.04 TB/s at 2^30 int64).
//
// One launch, functor called exactly TWICE per index (as on the OpenMP oracle), TWO block barriers per tile:
//   pass 1  each WARP owns 32*ITEMS consecutive elements; lane l calls f(i, c, false) for i = warp base + j*32 + l
//           (coalesced) -> contribution c, kept in registers while the loads are in flight;
//   scan    warp-private transposition through shared memory (striped write, blocked read; ITEMS odd: conflict free;
//           __syncwarp only), serial fold of ITEMS values per lane, ordered warp scan of the lane totals (shuffles);
//           warp totals meet in shared memory (barrier 1) and EVERY warp folds the <= 32 of them itself;
//   chain   tile aggregate published, decoupled look-back over the predecessors: LBW windows of 32 descriptors are
//           requested together (one L2 round trip usually covers the whole distance to the nearest inclusive prefix),
//           folded in tile order so non-commutative joins see operands in index order; inclusive prefix published;
//   pass 2  the tile prefix reaches the other warps (barrier 2): u = tile prefix (+) warp offset (+) local prefix ;
//           f(i, u, true), striped again (the functor's loads hit L1/L2).
// Software-pipelined across tiles: the look-back and pass 2 of tile k run while the pass-1 loads of tile k+1 are in flight.
// Descriptors: value types of <= 8 bytes use ONE 16-byte descriptor {value bits, epoch<<2|state} per tile, written and
// read with a single relaxed 128-bit access; larger value types use a status word (st.release / ld.acquire) + two value
// slots per tile.  The next tile id is fetched one tile ahead so the atomic's round trip overlaps pass 1.
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

// packed-descriptor form of the above (sizeof(V) <= 8): one 128-bit load per lane and window, LBW windows requested together
template <class Red, int LBW>
KB200_DEVICE_FUNCTION typename Red::value_type lookback_ordered_packed(const Red& red, const GenericScanScratch<typename Red::value_type>& s,
                                                                       int64 tile, int lane) {
  using V = typename Red::value_type;
  V excl;
  red.init(excl);
  bool have = false;
  int64 wbase = tile - 1;
  while (true) {
    unsigned long long pay[LBW], st[LBW];
#pragma unroll
    for (int j = 0; j < LBW; ++j) {
      const int64 idx = wbase - ((int64)j * 32 + lane);
      if (idx >= 0) ptx::ld_relaxed_v2(s.desc + idx, pay[j], st[j]);
      else { pay[j] = 0; st[j] = 0; }
    }
    bool retry = false, done = false;
#pragma unroll
    for (int j = 0; j < LBW; ++j) {
      if (!retry && !done) {
        const int64 idx = wbase - lane;  // wbase has moved past the windows already folded
        V val;
        red.init(val);
        int state = 2;  // before the first tile: the identity is an inclusive prefix
        if (idx >= 0) {
          state = ((st[j] >> 2) == s.epoch) ? (int)(st[j] & 3ull) : 0;
          if (state) memcpy(&val, &pay[j], sizeof(V));
        }
        const unsigned term = __ballot_sync(kFullMask, state == 2);
        const unsigned inval = __ballot_sync(kFullMask, state == 0);
        const int first_term = term ? (__ffs(term) - 1) : 32;
        const unsigned needed = first_term >= 31 ? kFullMask : ((2u << first_term) - 1u);
        if (inval & needed) {
          retry = true;  // a needed predecessor has not published yet; what was folded so far stays valid
        } else {
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
          if (term) done = true; else wbase -= 32;
        }
      }
    }
    if (done) return excl;
    if (retry) __nanosleep(64);
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
template <class F, class Tag, class Index, class Red, int BLOCK, int ITEMS, int LBW>
__global__ void __launch_bounds__(BLOCK)
    generic_scan_kernel(const __grid_constant__ F f, const __grid_constant__ Red red, const Index begin, const int64 n,
                        const int64 ntiles, const GenericScanScratch<typename Red::value_type> s,
                        typename Red::value_type* total0, typename Red::value_type* total1) {
  using V = typename Red::value_type;
  constexpr int TILE = BLOCK * ITEMS;
  constexpr int WTILE = 32 * ITEMS;  // elements owned by one warp
  constexpr int NWARPS = BLOCK / 32;
  constexpr bool PACKED = sizeof(V) <= 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // TILE values, one private region of WTILE per warp.  The region serves both tiles in flight: pass 2 reads element e of the
  // previous tile and the transposition overwrites the same e with the new tile's contribution from the SAME lane.
  V* const vals = reinterpret_cast<V*>(smem_raw);
  V* const s_warp = vals + TILE;     // 32 values: inclusive totals of the warps of the tile being scanned
  V* const s_woff = s_warp + 32;     // 32 values: exclusive offsets of the warps inside the tile awaiting pass 2
  V* const s_prefix = s_woff + 32;   // 1 value: exclusive prefix of the tile being finished
  __shared__ int64 s_tile[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  V* const wvals = vals + warp * WTILE;

  auto call = [&](int64 i, V& v, bool fin) {
    if constexpr (std::is_void<Tag>::value) f((Index)(begin + (Index)i), v, fin);
    else f(Tag{}, (Index)(begin + (Index)i), v, fin);
  };

  if (tid == 0) s_tile[0] = (int64)(atomicAdd(s.counter, 1ull) - s.counter_base);
  __syncthreads();
  int64 prev = -1;  // tile whose final pass is still owed
  V agg_prev;  // aggregate of tile `prev` (warp 0 only)
  red.init(agg_prev);
  for (int it = 0;; it ^= 1) {
    const int64 tile = s_tile[it];  // written before the last barrier this thread passed
    const bool have_new = tile < ntiles;
    if (!have_new && prev < 0) break;
    // next tile id, one tile ahead; every CTA takes exactly one id past the end
    if (have_new && tid == 0) s_tile[it ^ 1] = (int64)(atomicAdd(s.counter, 1ull) - s.counter_base);
    const int64 wbase = tile * TILE + (int64)warp * WTILE + lane;

    // ---- (1) first functor call of the NEW tile: contributions into registers (loads go in flight)
    V c[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      red.init(c[j]);
      const int64 i = wbase + j * 32;
      if (have_new && i < n) call(i, c[j], false);
    }

    // ---- (2) finish the PREVIOUS tile: look-back, inclusive prefix, final functor call
    if (prev >= 0) {
      if (warp == 0) {
        V excl;
        red.init(excl);
        if (prev > 0) {
          if constexpr (PACKED) excl = lookback_ordered_packed<Red, LBW>(red, s, prev, lane);
          else excl = lookback_ordered(red, s, prev, lane);
          if (lane == 0) {
            V inc = excl;
            red.join(inc, agg_prev);
            if constexpr (PACKED) publish_packed(s.desc + prev, inc, s.epoch, 2ull);
            else { store_value(s.incl + prev, inc); ptx::st_release_u64(s.status + prev, (s.epoch << 2) | 2ull); }
          }
        }
        if (lane == 0) {
          *s_prefix = excl;
          if (prev == ntiles - 1 && (total0 || total1)) {
            V total = excl;
            red.join(total, agg_prev);
            // parallel_scan's total is the plain running value (no final(): Kokkos_Parallel.hpp:405-425)
            if (total0) *total0 = total;
            if (total1) *total1 = total;
          }
        }
      }
      __syncthreads();  // barrier 2 (of tile prev): the tile prefix is visible; every warp's local prefixes are in place
      V tprefix = *s_prefix;
      { V wo = s_woff[warp]; red.join(tprefix, wo); }
      const int64 pbase = prev * TILE + (int64)warp * WTILE + lane;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const int64 i = pbase + j * 32;
        if (i < n) {
          V u = tprefix;
          { V lp = wvals[j * 32 + lane]; red.join(u, lp); }
          call(i, u, true);
        }
      }
    }
    prev = -1;
    if (!have_new) break;  // ids are monotonic: nothing more will come

    // ---- (3) scan of the new tile: warp-private transposition, aggregate published after ONE block barrier
    __syncwarp();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) wvals[j * 32 + lane] = c[j];
    __syncwarp();
    V loc[ITEMS];
    V tsum;
    red.init(tsum);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      loc[k] = wvals[lane * ITEMS + k];
      red.join(tsum, loc[k]);
    }
    V tincl = tsum;
    warp_incl_scan_ordered(red, tincl, lane);
    if (lane == 31) s_warp[warp] = tincl;
    __syncthreads();  // barrier 1: warp totals (and the next tile id) are visible
    if (warp == 0) {
      // ONE warp folds the warp totals and publishes the tile aggregate at once; the exclusive warp offsets are only needed in
      // pass 2, i.e. after the next barrier 2, so nobody waits for them here
      V w;
      red.init(w);
      if (lane < NWARPS) w = s_warp[lane];
      V wi = w;
      warp_incl_scan_ordered(red, wi, lane);
      V wex = ::kb200::Impl::shfl_up(wi, 1);  // exclusive warp prefix
      if (lane == 0) red.init(wex);
      if (lane < NWARPS) s_woff[lane] = wex;
      agg_prev = ::kb200::Impl::shfl_idx(wi, NWARPS - 1);
      if (lane == 0) {
        const unsigned long long state = tile == 0 ? 2ull : 1ull;  // tile 0: its aggregate IS its inclusive prefix
        if constexpr (PACKED) publish_packed(s.desc + tile, agg_prev, s.epoch, state);
        else {
          store_value((tile == 0 ? s.incl : s.agg) + tile, agg_prev);
          ptx::st_release_u64(s.status + tile, (s.epoch << 2) | state);
        }
      }
    }
    {  // warp-local exclusive prefix of every element, blocked write-back (warp offset and tile prefix are joined in pass 2)
      V run = ::kb200::Impl::shfl_up(tincl, 1);
      if (lane == 0) red.init(run);
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        wvals[lane * ITEMS + k] = run;
        red.join(run, loc[k]);
      }
    }
    prev = tile;
  }
  if (tid == 0) scan_counter_release(s.counter);
}

// BLOCK_ / ITEMS_ = 0: the shipped tile shape for this value size (tests/cxx/cases_perf.cu instantiates alternatives)
template <class Policy, class F, class Red, int BLOCK_ = 0, int ITEMS_ = 0, int LBW_ = 0>
struct GenericScan {
  using V = typename Red::value_type;
  using Index = typename Policy::index_type;
  using Tag = typename Policy::work_tag;
  // Large tiles: the per-tile fixed cost (look-back wait + two barriers + the cross-warp fold: ~2.4 us per tile, during which
  // nothing of this CTA is in flight) is what bounds this kernel; the streaming part runs at ~0.32 us per 1024-element row
  // (= 7.5 TB/s over 148 SMs).  B200 sweep at 2^30 int64 (profiles/r02_gscan_probe.log): 1024x13 4.78, 1024x15 4.90,
  // 1024x17 5.25 TB/s; 512-thread CTAs (two per SM) lose more to the longer look-back than they gain in overlap (512x17 4.1).
  static constexpr int BLOCK = BLOCK_ ? BLOCK_ : (sizeof(V) <= 8 ? 1024 : (sizeof(V) <= 16 ? 512 : (sizeof(V) <= 64 ? 256 : 128)));
  static constexpr int ITEMS = ITEMS_ ? ITEMS_ : (sizeof(V) <= 8 ? 17 : (sizeof(V) <= 16 ? 9 : (sizeof(V) <= 32 ? 7 : 3)));
  static constexpr int LBW = LBW_ ? LBW_ : 2;  // 64 predecessors per look-back round trip (4 costs registers: spills at 1024 threads)
  static constexpr int TILE = BLOCK * ITEMS;
  static constexpr size_t SMEM = (size_t)(TILE + 66) * sizeof(V);
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
    auto k = generic_scan_kernel<F, Tag, Index, Red, BLOCK, ITEMS, LBW>;
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
