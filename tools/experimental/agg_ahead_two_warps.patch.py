"""tools/experimental/agg_ahead_two_warps.patch.py <path to a COPY of ScanContig.hpp> -- NOT part of the build.\nThe "aggregates ahead" mode of the distributed scan with TWO aggregate warps per CTA, their own request rings and an ordered merge in\nthe round warp (see profiles/r02_cyclic_scan_probe.log).  Applied to a copy of kokkos_b200/include/kb200/impl/ScanContig.hpp\n(plus the comm.agd tune key passed to run_rounds in csrc/comm.cu) it builds, is bit-exact (tests/comm_worker.py ... big) and dead-lock free,\nbut it is SLOWER than the shipped kernel already at world 1 (5.0-5.2 vs 6.2 TB/s), so it cannot help at 8 GPUs either; kept as a record."""
import sys
p=sys.argv[1]
s=open(p).read()
def rep(old,new,count=None):
    global s
    assert s.count(old)>=1, old[:90]
    s = s.replace(old,new) if count is None else s.replace(old,new,count)
rep("  int prefetch_tiles;         // > 0: whoever takes tile t also asks the L2 for tile t + prefetch_tiles (HBM runs ahead of the stage ring)",
"  int prefetch_tiles;         // > 0: whoever takes tile t also asks the L2 for tile t + prefetch_tiles (HBM runs ahead of the stage ring)\n  int agg_ahead;              // > 0 (ROUNDS, integral T): whoever takes tile t computes and publishes the AGGREGATE of tile t + agg_ahead\n                              // from global memory (an L2 hit after the prefetch): aggregates, round accumulators and the exchange of\n                              // round aggregates between GPUs run that many tiles ahead of the data pipeline")
rep('#include "Collectives.hpp"\n#include "HostRuntime.hpp"\n#include "Ptx.hpp"\n','#include "Collectives.hpp"\n#include "ContigBody.hpp"\n#include "HostRuntime.hpp"\n#include "Ptx.hpp"\n',1)
rep("__global__ void __launch_bounds__(CBLOCK + 96 + (ROUNDS ? 32 : 0)) contig_scan_ws2_kernel(const ScanContigParams<T> p) {","__global__ void __launch_bounds__(CBLOCK + 96 + (ROUNDS ? 64 : 0)) contig_scan_ws2_kernel(const ScanContigParams<T> p) {")
# shared
rep("  __shared__ __align__(8) unsigned long long rq_full[RQ], rq_empty[RQ];\n  __shared__ int64 s_rq[RQ];\n\n  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;\n  KB200_STATS_DECL;\n  if (tid == 0) {\n#pragma unroll\n    for (int b = 0; b < NSTAGE; ++b) {",
"""  __shared__ __align__(8) unsigned long long rq_full[RQ], rq_empty[RQ];
  __shared__ int64 s_rq[RQ];
  // agg_ahead mode: a SECOND aggregate warp with its own request ring, the claimed tile ids handed over at claim time, and the
  // low-water marks that let the one round warp serve the two rings in increasing round order
  __shared__ __align__(8) unsigned long long rq2_full[RQ], rq2_empty[RQ];
  __shared__ int64 s_rq2[RQ];
  constexpr int CQ = 8;
  __shared__ __align__(8) unsigned long long cq_full[CQ], cq_empty[CQ];
  __shared__ int64 s_cq[CQ];
  __shared__ int64 s_lwm[2];      // in-progress item of aggregate warp 0 / 1: smallest round request it can still push; kIdle between items
  __shared__ int64 s_started[2];  // highest claim index that warp has picked up (-1: none yet; huge once it has finished)
  __shared__ int64 s_rqc[RQ], s_rq2c[RQ];  // claim index each queued request came from

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  KB200_STATS_DECL;
  if (tid == 0) {
    s_lwm[0] = 0x7ffffffffffffffell;  // kIdle
    s_lwm[1] = 0x7ffffffffffffffell;
    s_started[0] = -1;
    s_started[1] = -1;
#pragma unroll
    for (int b = 0; b < CQ; ++b) { ptx::mbar_init(&cq_full[b], 1); ptx::mbar_init(&cq_empty[b], 1); }
#pragma unroll
    for (int b = 0; b < RQ; ++b) { ptx::mbar_init(&rq2_full[b], 1); ptx::mbar_init(&rq2_empty[b], 1); }
#pragma unroll
    for (int b = 0; b < NSTAGE; ++b) {""")
# ROUND warp: ordered two-ring service
rep("""      const int64 nrounds = p.ntiles / p.tpr;
      for (int64 c = 0;; ++c) {
        const int slot = (int)(c % RQ);
        ptx::mbar_wait(&rq_full[slot], (unsigned)((c / RQ) & 1));
        const int64 kk = s_rq[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&rq_empty[slot]);
        if (kk < 0) return;
""","""      const int64 nrounds = p.ntiles / p.tpr;
      const bool two = p.agg_ahead > 0;  // two producers: requests are served in increasing round order (a handler may wait for
                                         // what the handlers of EARLIER rounds publish, so it must never run ahead of a smaller
                                         // request that is still queued, or still to be queued, in this CTA)
      int64 c0 = 0, c1 = 0;              // entries consumed from ring 0 / ring 1
      bool end0 = false, end1 = !two;
      for (;;) {
        long long kk = 0;
        if (lane == 0) {
          constexpr long long kNone = 0x7fffffffffffffffll, kIdle = 0x7ffffffffffffffell;
          for (;;) {
            long long h0 = kNone, h1 = kNone, q0 = 0, q1 = 0;
            bool v0 = false, v1 = false;
            if (!end0 && ptx::mbar_try_wait(&rq_full[c0 % RQ], (unsigned)((c0 / RQ) & 1))) { v0 = true; h0 = s_rq[c0 % RQ]; q0 = s_rqc[c0 % RQ]; }
            if (!end1 && ptx::mbar_try_wait(&rq2_full[c1 % RQ], (unsigned)((c1 / RQ) & 1))) { v1 = true; h1 = s_rq2[c1 % RQ]; q1 = s_rq2c[c1 % RQ]; }
            if (v0 && h0 < 0) { mbar_arrive(&rq_empty[c0 % RQ]); ++c0; end0 = true; continue; }
            if (v1 && h1 < 0) { mbar_arrive(&rq2_empty[c1 % RQ]); ++c1; end1 = true; continue; }
            if (end0 && end1) { kk = -1; break; }
            // May the head of ring a (request h, from claim q) be served now?  Yes if the other warp b cannot produce a smaller
            // request any more: its queued head (requests of one warp are increasing) must not be smaller; otherwise it must have
            // picked up every claim older than q (claims alternate: the newest older one is q - 1) and its item in progress, if
            // any, must not be able to ask for less.
            auto other_allows = [&](bool endb, bool vb, long long hb, int b, long long h, long long q) {
              if (endb) return true;
              if (vb) return h <= hb;
              const long long st = *reinterpret_cast<volatile long long*>(&s_started[b]);
              const long long lw = *reinterpret_cast<volatile long long*>(&s_lwm[b]);
              __threadfence_block();
              // the ring again, AFTER the marks: a warp pushes its request first and declares itself idle second, so "idle and
              // ring still empty" really means that nothing smaller is on its way
              unsigned long long* const fb = b ? rq2_full : rq_full;
              const int64 cb = b ? c1 : c0;
              if (ptx::mbar_try_wait(&fb[cb % RQ], (unsigned)((cb / RQ) & 1))) {
                const long long hb2 = b ? s_rq2[cb % RQ] : s_rq[cb % RQ];
                return hb2 >= 0 && h <= hb2;  // (an end marker is handled at the top of the loop)
              }
              if (st < q - 1) return false;
              return lw == kIdle || h <= lw;
            };
            if (v0 && other_allows(end1, v1, h1, 1, h0, q0)) { kk = h0; mbar_arrive(&rq_empty[c0 % RQ]); ++c0; break; }
            if (v1 && other_allows(end0, v0, h0, 0, h1, q1)) { kk = h1; mbar_arrive(&rq2_empty[c1 % RQ]); ++c1; break; }
          }
        }
        kk = __shfl_sync(kFullMask, kk, 0);
        if (kk < 0) return;
""")
# DMA
rep("    int64 jl = 0, js = 0, nvalid = 0;\n    int64 tl[NSTAGE];\n    bool more = true;\n    while (more || js < nvalid) {","    int64 jl = 0, js = 0, nvalid = 0, ncq = 0;\n    int64 tl[NSTAGE];\n    bool more = true;\n    while (more || js < nvalid) {",1)
rep("        if (lane == 0) s_tile_id[st] = tile;\n#pragma unroll\n        for (int b = 0; b < NSTAGE; ++b) if (b == st) tl[b] = tile;\n        if (tile >= p.ntiles) {",
    "        if (lane == 0) s_tile_id[st] = tile;\n        if (ROUNDS && p.agg_ahead > 0 && lane == 0) {\n          rq_push<CQ>(ncq, cq_full, cq_empty, s_cq, tile);\n          if (tile >= p.ntiles) rq_push<CQ>(ncq, cq_full, cq_empty, s_cq, tile);  // one end-of-work marker per aggregate warp\n        }\n#pragma unroll\n        for (int b = 0; b < NSTAGE; ++b) if (b == st) tl[b] = tile;\n        if (tile >= p.ntiles) {")
# AGG warps
rep("  if (warp == NWARPS + 1) {\n    // ================= AGGREGATE warp =================\n    int64 agg_nreq = 0;  // ATOMIC_AGG: requests handed to the round warp so far (lane 0)\n",
'''  if (warp == NWARPS + 1 || (ROUNDS && warp == NWARPS + 4)) {
    // ================= AGGREGATE warp(s) =================
    if (warp == NWARPS + 4 && p.agg_ahead <= 0) return;  // the second one only exists for the aggregate-ahead mode
    int64 agg_nreq = 0;  // ATOMIC_AGG: requests handed to the round warp so far (lane 0)
    if (ROUNDS && p.agg_ahead > 0) {
      if constexpr (ROUNDS && std::is_integral<T>::value) {
        // ---- aggregates AHEAD of the data pipeline (see ScanContigParams::agg_ahead).  Warp 0 first publishes the tiles nobody is
        // ahead of (u < D, dealt out by CTA index), then both warps take the claimed tile ids alternately and publish tile t + D.
        // Each warp's own sequence of tiles is increasing, its round requests go to its OWN ring, and it announces the smallest
        // request it can still make (s_lwm) before every tile, so the round warp can merge the two rings in round order.
        const int w = (warp == NWARPS + 4) ? 1 : 0;
        unsigned long long* const my_full = w ? rq2_full : rq_full;
        unsigned long long* const my_empty = w ? rq2_empty : rq_empty;
        int64* const my_rq = w ? s_rq2 : s_rq;
        int64* const my_rqc = w ? s_rq2c : s_rqc;
        constexpr long long kIdle = 0x7ffffffffffffffell;
        auto push = [&](int64 kk, int64 claim) {  // lane 0; like rq_push, plus the claim index
          const int slot = (int)(agg_nreq % RQ);
          if (agg_nreq >= RQ) ptx::mbar_wait(&my_empty[slot], (unsigned)(((agg_nreq / RQ) - 1) & 1));
          my_rq[slot] = kk;
          my_rqc[slot] = claim;
          ptx::mbar_arrive(&my_full[slot]);
          ++agg_nreq;
        };
        auto publish = [&](int64 u, int64 claim) {
          if (lane == 0) {  // first what this item can still ask for, then the fact that the claim has been picked up
            *reinterpret_cast<volatile long long*>(&s_lwm[w]) = (u == 0) ? 0ll : (long long)(u / p.tpr + 1);
            __threadfence_block();
            *reinterpret_cast<volatile long long*>(&s_started[w]) = (long long)claim;
            __threadfence_block();
          }
          const int64 base = u * TILE;
          T acc[4] = {T(0), T(0), T(0), T(0)};
          if (p.bulk_load && base + TILE <= p.n) {
            const char* src = reinterpret_cast<const char*>(p.x + base);
            constexpr int NVEC = (int)(TILE_BYTES / 16);
            constexpr int UN = 9;
            for (int i0 = lane; i0 < NVEC; i0 += 32 * UN) {
              RawVec<16> q[UN];
#pragma unroll
              for (int r = 0; r < UN; ++r)
                if (i0 + 32 * r < NVEC) q[r] = ld_stream<16>(src + (size_t)(i0 + 32 * r) * 16);
#pragma unroll
              for (int r = 0; r < UN; ++r)
                if (i0 + 32 * r < NVEC) {
                  T e[EPV];
                  memcpy(e, q[r].w, 16);
#pragma unroll
                  for (int k = 0; k < EPV; ++k) acc[k & 3] += e[k];
                }
            }
          } else {
            const int64 remaining = p.n - base;
            for (int i = lane; i < TILE && i < remaining; i += 32) acc[i & 3] += p.x[base + i];
          }
          const T agg = warp_sum_all<T>((acc[0] + acc[1]) + (acc[2] + acc[3]));
          if (lane == 0) {
            const bool first = (u % p.tpr == 0);
            ptx::st_relaxed_v2(p.desc + u, to_bits(agg), (p.epoch << 2) | (first ? kDescIncl : kDescAgg));
            if constexpr (ATOMIC_AGG) {
              const int64 kq = u / p.tpr;
              unsigned long long* const racc = p.racc + (size_t)(kq % kRoundRing) * 2;
              const unsigned long long b = to_bits(agg);
              asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(racc), "l"((b & 0xffffffffull) | (1ull << 48)) : "memory");
              asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(racc + 1), "l"((b >> 32) | (1ull << 48)) : "memory");
              if (u == 0) push(0, claim);
              if (u % p.tpr == p.tpr - 1) push(kq + 1, claim);
            }
            __threadfence_block();
            *reinterpret_cast<volatile long long*>(&s_lwm[w]) = kIdle;  // (after the pushes: they are visible as the ring's head)
          }
          __syncwarp();
        };
        if (w == 0)
          for (int64 u = blockIdx.x; u < p.agg_ahead && u < p.ntiles; u += gridDim.x) publish(u, -1);
        for (int64 c = w;; c += 2) {  // the two warps take alternate claims
          const int slot = (int)(c % CQ);
          ptx::mbar_wait(&cq_full[slot], (unsigned)((c / CQ) & 1));
          const int64 t = s_cq[slot];
          __syncwarp();
          if (lane == 0) mbar_arrive(&cq_empty[slot]);
          if (t >= p.ntiles) {
            if (lane == 0) {
              if constexpr (ATOMIC_AGG) push(-1, c);
              __threadfence_block();
              *reinterpret_cast<volatile long long*>(&s_started[w]) = 0x7fffffffffffffffll;
            }
            { KB200_STATS_FLUSH(); return; }
          }
          if (t + p.agg_ahead < p.ntiles) publish(t + p.agg_ahead, c);
          else if (lane == 0) {  // nothing to publish for this claim, but it HAS been picked up
            *reinterpret_cast<volatile long long*>(&s_started[w]) = (long long)c;
            __threadfence_block();
          }
        }
      }
    }
''')
# look-back warp
rep("      ptx::mbar_wait(&aggready[st], (unsigned)((j / NSTAGE) & 1));\n#ifdef B200_SWEEP\n      if (lane == 0) KB200_STAT_ADD(7, clock64() - t_w2);\n#endif\n      const int64 cur = s_tile_id[st];\n      if (cur >= p.ntiles) {\n        if constexpr (ROUNDS && !ATOMIC_AGG) request_round(-1);  // no more work for the round warp\n        return;\n      }\n      const T agg = s_agg[st];",
    "      if (ROUNDS && p.agg_ahead > 0) ptx::mbar_wait(&full[st], (unsigned)((j / NSTAGE) & 1));  // (the tile id is all this warp needs from the stage)\n      else ptx::mbar_wait(&aggready[st], (unsigned)((j / NSTAGE) & 1));\n#ifdef B200_SWEEP\n      if (lane == 0) KB200_STAT_ADD(7, clock64() - t_w2);\n#endif\n      const int64 cur = s_tile_id[st];\n      if (cur >= p.ntiles) {\n        if constexpr (ROUNDS && !ATOMIC_AGG) request_round(-1);  // no more work for the round warp\n        return;\n      }\n      T agg = s_agg[st];\n      if (ROUNDS && p.agg_ahead > 0) {  // published ahead of time by whoever took tile cur - D (or by a CTA's prologue)\n        unsigned long long pay, stw;\n        do { ptx::ld_relaxed_v2(p.desc + cur, pay, stw); } while ((stw >> 2) != p.epoch);\n        agg = from_bits<T>(pay);\n      }")
# host
rep("                        int prefetch_tiles = 0) {","                        int prefetch_tiles = 0, int agg_ahead = 0) {")
rep("    p.prefetch_tiles = prefetch_tiles < 0 ? grid : prefetch_tiles;  // auto: one wave of CTAs ahead\n    p.tpr = tpr;","    p.agg_ahead = std::is_integral<T>::value ? (agg_ahead < 0 ? grid : agg_ahead) : 0;\n    p.prefetch_tiles = prefetch_tiles < 0 ? grid + p.agg_ahead : prefetch_tiles;  // one wave of CTAs ahead of the aggregates\n    p.tpr = tpr;")
rep("    p.prefetch_tiles = prefetch_tiles < 0 ? grid : prefetch_tiles;\n","    p.prefetch_tiles = prefetch_tiles < 0 ? grid : prefetch_tiles;\n    p.agg_ahead = 0;\n")
rep("cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, THREADS + 32, SMEM);","cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, THREADS + 64, SMEM);")
rep("rounds_kernel()<<<grid, THREADS + 32, SMEM, rt.stream()>>>(p);","rounds_kernel()<<<grid, THREADS + 64, SMEM, rt.stream()>>>(p);  // + the round warp and the second aggregate warp")
open(p,'w').write(s)
print("patched")
