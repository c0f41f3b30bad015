"""tools/experimental/agg_ahead_two_warps.patch.py <path to ScanContig.hpp> -- NOT part of the build.\nThe unfinished "aggregates ahead" mode of the distributed scan with TWO aggregate warps per CTA (see profiles/r02_cyclic_scan_probe.log and\nDESIGN.md section 7): applied to a copy of kokkos_b200/include/kb200/impl/ScanContig.hpp it compiles; the default path (comm.agd=0) runs at\nthe shipped speed, the ahead mode (comm.agd=-1) still fails with a launch failure at world 1 -- kept as the starting point for the next round."""
import sys
p=sys.argv[1]
s=open(p).read()
def rep(old,new,count=None):
    global s
    assert old in s, old[:80]
    s = s.replace(old,new) if count is None else s.replace(old,new,count)
# ---- params
rep("  int prefetch_tiles;         // > 0: whoever takes tile t also asks the L2 for tile t + prefetch_tiles (HBM runs ahead of the stage ring)",
"  int prefetch_tiles;         // > 0: whoever takes tile t also asks the L2 for tile t + prefetch_tiles (HBM runs ahead of the stage ring)\n  int agg_ahead;              // > 0 (ROUNDS, integral T): whoever takes tile t computes and publishes the AGGREGATE of tile t + agg_ahead\n                              // from global memory (an L2 hit after the prefetch): aggregates, round accumulators and the exchange of\n                              // round aggregates between GPUs run that many tiles ahead of the data pipeline")
rep('#include "Collectives.hpp"\n#include "HostRuntime.hpp"\n#include "Ptx.hpp"\n','#include "Collectives.hpp"\n#include "ContigBody.hpp"\n#include "HostRuntime.hpp"\n#include "Ptx.hpp"\n',1)
# ---- multi-producer push
rep('''// Round k of a block-cyclic distributed scan (ROUNDS kernels):
//   prefix of a tile = base(k) + (prefix inside the round),''','''// the same ring with several producers (the two aggregate-ahead warps): positions are handed out by a shared counter
template <int RQ>
KB200_DEVICE_FUNCTION void rq_push_mp(unsigned long long* tail, unsigned long long* rq_full, unsigned long long* rq_empty, int64* s_rq, int64 kk) {
  const int64 nreq = (int64)atomicAdd(tail, 1ull);
  const int slot = (int)(nreq % RQ);
  if (nreq >= RQ) ptx::mbar_wait(&rq_empty[slot], (unsigned)(((nreq / RQ) - 1) & 1));
  s_rq[slot] = kk;
  ptx::mbar_arrive(&rq_full[slot]);
}

// Round k of a block-cyclic distributed scan (ROUNDS kernels):
//   prefix of a tile = base(k) + (prefix inside the round),''',1)
rep("__global__ void __launch_bounds__(CBLOCK + 96 + (ROUNDS ? 32 : 0)) contig_scan_ws2_kernel(const ScanContigParams<T> p) {","__global__ void __launch_bounds__(CBLOCK + 96 + (ROUNDS ? 64 : 0)) contig_scan_ws2_kernel(const ScanContigParams<T> p) {")
rep("  __shared__ __align__(8) unsigned long long rq_full[RQ], rq_empty[RQ];\n  __shared__ int64 s_rq[RQ];\n\n  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;\n  KB200_STATS_DECL;\n  if (tid == 0) {\n#pragma unroll\n    for (int b = 0; b < NSTAGE; ++b) {\n      ptx::mbar_init(&full[b], 1); ptx::mbar_init(&aggready[b], 1);",
    "  __shared__ __align__(8) unsigned long long rq_full[RQ], rq_empty[RQ];\n  __shared__ int64 s_rq[RQ];\n  constexpr int CQ = 8;  // agg_ahead: tile ids handed from the DMA warp to the aggregate-ahead warps at CLAIM time\n  __shared__ __align__(8) unsigned long long cq_full[CQ], cq_empty[CQ];\n  __shared__ int64 s_cq[CQ];\n  __shared__ unsigned long long s_rq_tail;  // ring position of the next round request (two producers in agg_ahead mode)\n  __shared__ unsigned s_agg_done;           // aggregate-ahead warps that have finished\n\n  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;\n  KB200_STATS_DECL;\n  if (tid == 0) {\n    s_rq_tail = 0ull;\n    s_agg_done = 0u;\n#pragma unroll\n    for (int b = 0; b < CQ; ++b) { ptx::mbar_init(&cq_full[b], 1); ptx::mbar_init(&cq_empty[b], 1); }\n#pragma unroll\n    for (int b = 0; b < NSTAGE; ++b) {\n      ptx::mbar_init(&full[b], 1); ptx::mbar_init(&aggready[b], 1);")
# ---- DMA warp
rep("    int64 jl = 0, js = 0, nvalid = 0;\n    int64 tl[NSTAGE];\n    bool more = true;\n    while (more || js < nvalid) {","    int64 jl = 0, js = 0, nvalid = 0, ncq = 0;\n    int64 tl[NSTAGE];\n    bool more = true;\n    while (more || js < nvalid) {",1)
rep("        if (lane == 0) s_tile_id[st] = tile;\n#pragma unroll\n        for (int b = 0; b < NSTAGE; ++b) if (b == st) tl[b] = tile;\n        if (tile >= p.ntiles) {",
    "        if (lane == 0) s_tile_id[st] = tile;\n        if (ROUNDS && p.agg_ahead > 0 && lane == 0) {\n          rq_push<CQ>(ncq, cq_full, cq_empty, s_cq, tile);\n          if (tile >= p.ntiles) rq_push<CQ>(ncq, cq_full, cq_empty, s_cq, tile);  // one end-of-work marker per aggregate-ahead warp\n        }\n#pragma unroll\n        for (int b = 0; b < NSTAGE; ++b) if (b == st) tl[b] = tile;\n        if (tile >= p.ntiles) {")
# ---- AGG warps
rep("  if (warp == NWARPS + 1) {\n    // ================= AGGREGATE warp =================\n    int64 agg_nreq = 0;  // ATOMIC_AGG: requests handed to the round warp so far (lane 0)\n",
'''  if (warp == NWARPS + 1 || (ROUNDS && warp == NWARPS + 4)) {
    // ================= AGGREGATE warp(s) =================
    if (warp == NWARPS + 4 && p.agg_ahead <= 0) return;  // the second one only exists for the aggregate-ahead mode
    if (ROUNDS && p.agg_ahead > 0) {
      if constexpr (ROUNDS && std::is_integral<T>::value) {
        auto publish = [&](int64 u) {
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
              if (u == 0) rq_push_mp<RQ>(&s_rq_tail, rq_full, rq_empty, s_rq, 0);
              if (u % p.tpr == p.tpr - 1) rq_push_mp<RQ>(&s_rq_tail, rq_full, rq_empty, s_rq, kq + 1);
            }
          }
          __syncwarp();
        };
        for (int64 c = (warp == NWARPS + 4) ? 1 : 0;; c += 2) {  // the two warps take alternate claims
          const int slot = (int)(c % CQ);
          ptx::mbar_wait(&cq_full[slot], (unsigned)((c / CQ) & 1));
          const int64 t = s_cq[slot];
          __syncwarp();
          if (lane == 0) mbar_arrive(&cq_empty[slot]);
          if (t >= p.ntiles) {
            if constexpr (ATOMIC_AGG) {  // the LAST of the two tells the round warp that nothing more will come
              if (lane == 0 && atomicAdd(&s_agg_done, 1u) == 1u) rq_push_mp<RQ>(&s_rq_tail, rq_full, rq_empty, s_rq, -1);
            }
            { KB200_STATS_FLUSH(); return; }
          }
          if (t < p.agg_ahead) publish(t);
          if (t + p.agg_ahead < p.ntiles) publish(t + p.agg_ahead);
        }
      }
    }
''')
s=s.replace("rq_push<RQ>(agg_nreq, rq_full, rq_empty, s_rq, ","rq_push_mp<RQ>(&s_rq_tail, rq_full, rq_empty, s_rq, ")
# ---- look-back warp
rep("      ptx::mbar_wait(&aggready[st], (unsigned)((j / NSTAGE) & 1));\n#ifdef B200_SWEEP\n      if (lane == 0) KB200_STAT_ADD(7, clock64() - t_w2);\n#endif\n      const int64 cur = s_tile_id[st];\n      if (cur >= p.ntiles) {\n        if constexpr (ROUNDS && !ATOMIC_AGG) request_round(-1);  // no more work for the round warp\n        return;\n      }\n      const T agg = s_agg[st];",
    "      if (ROUNDS && p.agg_ahead > 0) ptx::mbar_wait(&full[st], (unsigned)((j / NSTAGE) & 1));  // (the tile id is all this warp needs from the stage)\n      else ptx::mbar_wait(&aggready[st], (unsigned)((j / NSTAGE) & 1));\n#ifdef B200_SWEEP\n      if (lane == 0) KB200_STAT_ADD(7, clock64() - t_w2);\n#endif\n      const int64 cur = s_tile_id[st];\n      if (cur >= p.ntiles) {\n        if constexpr (ROUNDS && !ATOMIC_AGG) request_round(-1);  // no more work for the round warp\n        return;\n      }\n      T agg = s_agg[st];\n      if (ROUNDS && p.agg_ahead > 0) {  // published ahead of time by whoever took tile cur - D (or by this CTA during start-up)\n        unsigned long long pay, stw;\n        do { ptx::ld_relaxed_v2(p.desc + cur, pay, stw); } while ((stw >> 2) != p.epoch);\n        agg = from_bits<T>(pay);\n      }")
# ---- host
rep("                        int prefetch_tiles = 0) {","                        int prefetch_tiles = 0, int agg_ahead = 0) {")
rep("    p.prefetch_tiles = prefetch_tiles < 0 ? grid : prefetch_tiles;  // auto: one wave of CTAs ahead\n    p.tpr = tpr;","    p.agg_ahead = std::is_integral<T>::value ? (agg_ahead < 0 ? grid : agg_ahead) : 0;\n    p.prefetch_tiles = prefetch_tiles < 0 ? grid + p.agg_ahead : prefetch_tiles;  // one wave of CTAs ahead of the aggregates\n    p.tpr = tpr;")
rep("    p.prefetch_tiles = prefetch_tiles < 0 ? grid : prefetch_tiles;\n","    p.prefetch_tiles = prefetch_tiles < 0 ? grid : prefetch_tiles;\n    p.agg_ahead = 0;\n")
rep("cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, THREADS + 32, SMEM);","cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, THREADS + 64, SMEM);")
rep("rounds_kernel()<<<grid, THREADS + 32, SMEM, rt.stream()>>>(p);","rounds_kernel()<<<grid, THREADS + 64, SMEM, rt.stream()>>>(p);  // + the round warp and the second aggregate-ahead warp")
open(p,'w').write(s)
print("patched")
