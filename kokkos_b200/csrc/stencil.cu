// stencil.cu -- MDRangePolicy<Rank<3>> parallel_reduce fast path (C ABI): 7-point stencil with a
// MinMaxLoc reducer over the interior of a LayoutLeft View<double***> (config C4).
//
// Stands for   parallel_reduce(MDRangePolicy<B200,Rank<3>>({1,1,1},{n0-1,n1-1,n2-1}),
//                  KB200_LAMBDA(i,j,k, MinMaxLoc::value_type& r){ v = ...; update r with loc (i*n1+j)*n2+k },
//                  MinMaxLoc<double,int64>(result));
// replacing ParallelReduce<...,MDRangePolicy,Cuda> (core/src/Cuda/Kokkos_Cuda_Parallel_MDRange.hpp:248-497),
// whose reduce grid is capped at <=512 blocks with 64 of 256 threads active and a div/mod per element
// (impl/KokkosExp_IterateTileGPU.hpp:1206-1305).
//
// Two kernels:
//  (A) stencil7_tma_kernel -- the shipped path for even n0 <= 512 and 16-byte aligned Views (config C4).
//      2.5-D blocking: a CTA owns a tile of BJ rows x KC planes and marches along k.  One PRODUCER warp streams
//      whole (BJ+2)-row slabs of each plane -- contiguous in a LayoutLeft View -- into a ring of shared-memory
//      stages with ONE 1-D bulk async copy per plane (TMA engine, cp.async.bulk -> SASS UBLKCP, completion on an
//      mbarrier); the bytes in flight live in shared memory, not registers, so 8 consumer warps are enough to keep
//      ~120 KB per SM outstanding.  CONSUMER threads own 2 rows x 2 consecutive i and keep the k-1 / k planes of
//      their points in registers: per plane step a unit reads the k+1 centre (2 x LDS.128), the two j-halo rows
//      (2 x LDS.128) and four i-halo values (LDS.64); every input byte is fetched from DRAM once (the k/j halos of
//      neighbouring tiles are L2 hits).  Stages are recycled through empty[] mbarriers (one arrival per warp).
//  (B) stencil7_minmaxloc_kernel -- any shape: one warp per (j,k) row, lanes along i, seven scalar loads per
//      point served by L1/L2; one integer division per ROW, none per element; persistent grid.
// Ties: equal extrema keep the LOWEST location, i.e. the first one in the reference host iteration order
// (i slowest ... k fastest, impl/KokkosExp_Host_IterateTile.hpp) -- in-thread updates and all joins use that rule.
#include <kokkos_b200.h>
#include "runtime_internal.h"
#include <kb200/Reducers.hpp>
#include <kb200/impl/Collectives.hpp>
#include <kb200/impl/HostRuntime.hpp>
#include <kb200/impl/Ptx.hpp>

using namespace kb200;
using namespace kb200::Impl;

namespace {
using Red = MinMaxLoc<double, int64>;
using V = Red::value_type;

// MinMaxLoc with the lowest-location tie rule (see the header): commutative, so any combine order gives the same bits
struct StencilRed {
  using value_type = V;
  KB200_DEVICE_FUNCTION void init(V& v) const { V d; Red(d).init(v); }
  KB200_DEVICE_FUNCTION void join(V& d, const V& s) const {
    if (s.min_val < d.min_val || (s.min_val == d.min_val && s.min_loc < d.min_loc)) { d.min_val = s.min_val; d.min_loc = s.min_loc; }
    if (s.max_val > d.max_val || (s.max_val == d.max_val && s.max_loc < d.max_loc)) { d.max_val = s.max_val; d.max_loc = s.max_loc; }
  }
  KB200_DEVICE_FUNCTION void final(V&) const {}
};

// evaluation order as written in the header, no FMA contraction
KB200_DEVICE_FUNCTION double stencil_value(double ctr, double xm, double xp, double ym, double yp, double zm, double zp, double c0, double c1) {
  double s = __dadd_rn(xm, xp);
  s = __dadd_rn(s, ym);
  s = __dadd_rn(s, yp);
  s = __dadd_rn(s, zm);
  s = __dadd_rn(s, zp);
  return __dadd_rn(__dmul_rn(c0, ctr), __dmul_rn(c1, s));
}
// rare path: fold one point into the accumulator with the lowest-location tie rule
KB200_DEVICE_FUNCTION void stencil_update(V& acc, double v, int i, int j, int k, int n1, int n2) {
  const int64 loc = ((int64)i * n1 + j) * n2 + k;
  if (v < acc.min_val || (v == acc.min_val && loc < acc.min_loc)) { acc.min_val = v; acc.min_loc = loc; }
  if (v > acc.max_val || (v == acc.max_val && loc < acc.max_loc)) { acc.max_val = v; acc.max_loc = loc; }
}

// jstep/kstep > 1: only every jstep-th row of every kstep-th plane is visited (the sampling pre-pass of the TMA path)
template <int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK) stencil7_minmaxloc_kernel(const double* __restrict__ u, double* __restrict__ vout,
                                                                    int64 n0, int64 n1, int64 n2, double c0, double c1,
                                                                    ReduceScratch scratch, int jstep = 1, int kstep = 1) {
  __shared__ __align__(16) unsigned char smem[32 * sizeof(V)];
  const StencilRed red;
  V acc;
  red.init(acc);
  const int lane = threadIdx.x & 31;
  const int64 warps_per_grid = (int64)gridDim.x * (BLOCK / 32);
  const int64 m1 = (n1 - 2 + jstep - 1) / jstep, m2 = (n2 - 2 + kstep - 1) / kstep;
  const int64 rows = m1 * m2;
  const int64 sj = n0, sk = n0 * n1;
  // consecutive warps of a block take consecutive rows (j fastest) so j+-1 neighbours share L1
  for (int64 row = (int64)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); row < rows; row += warps_per_grid) {
    const int64 kq = row / m1;
    const int64 k = kq * kstep + 1;
    const int64 j = (row - kq * m1) * jstep + 1;
    const double* c = u + j * sj + k * sk;
    for (int64 i0 = 1; i0 < n0 - 1; i0 += 32 * UNROLL) {
      double ctr[UNROLL], xm[UNROLL], xp[UNROLL], ym[UNROLL], yp[UNROLL], zm[UNROLL], zp[UNROLL];
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) {
        const int64 i = i0 + q * 32 + lane;
        if (i < n0 - 1) {
          ctr[q] = c[i];
          xm[q] = c[i - 1];
          xp[q] = c[i + 1];
          ym[q] = c[i - sj];
          yp[q] = c[i + sj];
          zm[q] = c[i - sk];
          zp[q] = c[i + sk];
        }
      }
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) {
        const int64 i = i0 + q * 32 + lane;
        if (i < n0 - 1) {
          // evaluation order as written in the header, no FMA contraction
          double s = __dadd_rn(xm[q], xp[q]);
          s = __dadd_rn(s, ym[q]);
          s = __dadd_rn(s, yp[q]);
          s = __dadd_rn(s, zm[q]);
          s = __dadd_rn(s, zp[q]);
          const double v = __dadd_rn(__dmul_rn(c0, ctr[q]), __dmul_rn(c1, s));
          if (vout) vout[i + j * sj + k * sk] = v;
          if (v <= acc.min_val || v >= acc.max_val) stencil_update(acc, v, (int)i, (int)j, (int)k, (int)n1, (int)n2);
        }
      }
    }
  }
  block_reduce(red, acc, smem);
  __syncthreads();
  grid_reduce_and_store(red, acc, scratch, smem);
}

// hit = any of four values <= mn or >= mx.  Written as a predicate chain in PTX (8 x DSETP.xx.OR): the C++ form
// `(a <= mn) | (b <= mn) | ...` is rewritten by the compiler into fmin/fmax trees, ~10 instructions per min on sm_100
// (no DMNMX), which tripled the kernel's instruction count (profiles/r01_stencil_v2_ncu.txt).  NaN operands never hit.
KB200_DEVICE_FUNCTION bool stencil_hit(double a, double b, double c, double d, double mn, double mx) {
  unsigned r;
  asm("{\n"
      " .reg .pred p;\n"
      " setp.le.f64 p, %1, %5;\n"
      " setp.le.or.f64 p, %2, %5, p;\n"
      " setp.le.or.f64 p, %3, %5, p;\n"
      " setp.le.or.f64 p, %4, %5, p;\n"
      " setp.ge.or.f64 p, %1, %6, p;\n"
      " setp.ge.or.f64 p, %2, %6, p;\n"
      " setp.ge.or.f64 p, %3, %6, p;\n"
      " setp.ge.or.f64 p, %4, %6, p;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(r)
      : "d"(a), "d"(b), "d"(c), "d"(d), "d"(mn), "d"(mx));
  return r != 0;
}

// Per-unit constants (a unit = 2 rows x 2 consecutive i of the tile), in elements relative to a stage base.  Everything
// else (column, row, validity masks) is derived from these on the rare paths to keep the hot loop's register set small.
struct StencilUnit {
  int off;     // first point of the unit: (2*rg+1)*n0 + i0   (stage row 0 is the j0-1 halo)
  int xl, xr;  // i-1 / i+2 halo of row 0 (row 1: + n0); for i0 == 0 / i0+2 == n0 they address the stage's NaN pad,
               // so the two boundary columns evaluate to NaN and drop out of every comparison without a select
};

// One plane step of one unit: prev/cur are the k-1/k planes of the unit's points (registers), nxt receives the k+1 plane.
// P = stage of plane k (j- and i-halos), N = stage of plane k+1.  FULL: every row of the tile is interior (no row masks).
// vplane = vout + n0*(j0 + n1*k) (tile/plane base of the output).
template <bool STORE, bool FULL, int N0T, bool SLOW>
KB200_DEVICE_FUNCTION bool stencil_unit_step(const double* __restrict__ P, const double* __restrict__ N, const StencilUnit& un, int n0rt,
                                             int pad, int nrows, const double2 (&prev)[2], const double2 (&cur)[2], double2 (&nxt)[2],
                                             double c0, double c1, int j0, int k, int n1, int n2, double* __restrict__ vplane, V& acc, double tmin, double tmax) {
  const int n0 = N0T ? N0T : n0rt;
  const double* Pu = P + un.off;
  const double* Nu = N + un.off;
  if constexpr (!SLOW) {
    nxt[0] = *reinterpret_cast<const double2*>(Nu);
    nxt[1] = *reinterpret_cast<const double2*>(Nu + n0);
  }
  const double2 ym = *reinterpret_cast<const double2*>(Pu - n0);
  const double2 yp = *reinterpret_cast<const double2*>(Pu + 2 * n0);
  const double xm0 = P[un.xl], xp0 = P[un.xr];
  const double xm1 = P[un.xl + n0], xp1 = P[un.xr + n0];
  const double v00 = stencil_value(cur[0].x, xm0, cur[0].y, ym.x, cur[1].x, prev[0].x, nxt[0].x, c0, c1);
  const double v01 = stencil_value(cur[0].y, cur[0].x, xp0, ym.y, cur[1].y, prev[0].y, nxt[0].y, c0, c1);
  const double v10 = stencil_value(cur[1].x, xm1, cur[1].y, cur[0].x, yp.x, prev[1].x, nxt[1].x, c0, c1);
  const double v11 = stencil_value(cur[1].y, cur[1].x, xp1, cur[0].y, yp.y, prev[1].y, nxt[1].y, c0, c1);
  // validity: columns from the halo addresses, rows (partial tiles only) from the unit's row
  unsigned vm = (un.xl != pad ? 5u : 0u) | (un.xr != pad ? 10u : 0u);
  if constexpr (!FULL) {
    const int row = un.off / n0 - 1;
    vm &= (row < nrows ? 3u : 0u) | (row + 1 < nrows ? 12u : 0u);
  }
  if constexpr (STORE && !SLOW) {
    double* o = vplane + (un.off - n0);  // = i0 + n0*row
    if ((vm & 3u) == 3u) *reinterpret_cast<double2*>(o) = make_double2(v00, v01);
    else { if (vm & 1u) o[0] = v00; if (vm & 2u) o[1] = v01; }
    o += n0;
    if ((vm & 12u) == 12u) *reinterpret_cast<double2*>(o) = make_double2(v10, v11);
    else { if (vm & 4u) o[0] = v10; if (vm & 8u) o[1] = v11; }
  }
  if constexpr (SLOW) {  // rare re-evaluation of a step in which some lane saw a candidate: exact rule with locations
    const int row = un.off / n0 - 1, i0 = un.off - (row + 1) * n0, j = j0 + row;
    if (vm & 1u) stencil_update(acc, v00, i0, j, k, n1, n2);
    if (vm & 2u) stencil_update(acc, v01, i0 + 1, j, k, n1, n2);
    if (vm & 4u) stencil_update(acc, v10, i0, j + 1, k, n1, n2);
    if (vm & 8u) stencil_update(acc, v11, i0 + 1, j + 1, k, n1, n2);
    return false;
  }
  bool hit;
  if constexpr (FULL) {
    hit = stencil_hit(v00, v01, v10, v11, tmin, tmax);  // boundary columns are NaN already
  } else {  // partial tile: rows past the end hold stale data -> poison them
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    hit = stencil_hit((vm & 1u) ? v00 : qnan, (vm & 2u) ? v01 : qnan, (vm & 4u) ? v10 : qnan, (vm & 8u) ? v11 : qnan, tmin, tmax);
  }
  return hit;
}

// Warp-shared thresholds: a point can only be the global extremum if it reaches the best value ANY lane of the warp has
// seen (ties included, hence <= / >= in the hit test).  With per-thread thresholds each thread's ~2500 points trigger
// ~2 ln(2500) updates and some lane of a warp hit in 30 % of the steps (profiles/r01_stencil_v3_ncu.txt).
KB200_DEVICE_FUNCTION void stencil_refresh_thresholds(const V& acc, double& tmin, double& tmax) {
  double a = acc.min_val, b = acc.max_val;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double a2 = shfl_xor(a, d), b2 = shfl_xor(b, d);
    a = a2 < a ? a2 : a;
    b = b2 > b ? b2 : b;
  }
  tmin = a; tmax = b;
}

// probe modes exist only in the sweep build (tools/stencil_probe.py --sweep --dbg); the shipped kernel has no such branches
#ifdef B200_SWEEP
#define KB200_STENCIL_DBG(P, X) ((P).dbg == (X))
#else
#define KB200_STENCIL_DBG(P, X) false
#endif

struct StencilTmaParams {
  const double* u;
  double* vout;
  int n0, n1, n2;
  int kc, tiles_j, tiles_k;
  double c0, c1;
  const V* seed;  // MinMaxLoc over a sparse sample of the same points (or nullptr): every thread starts from it
  int dbg;  // tools/stencil_probe.py experiments only: 1 = consumers skip the arithmetic (pure data movement), 2 = no data movement (pure arithmetic on stale smem)
};

// stage = (BJ+2) rows of n0 doubles + a pad of n0+2 doubles whose elements [0,1] and [n0,n0+1] are NaN
KB200_FUNCTION constexpr int stencil_stage_elems(int bj, int n0) { return (bj + 2) * n0 + n0 + 2; }

// BJ rows per tile (even), NS stages, CT consumer threads (+ one producer warp); N0T = n0 when known at compile time (else 0)
// cursor over the block's sequence of plane slabs (tile-major, k inside): used by the dedicated producer warp, or -- when
// every warp computes (PW = false) -- by thread 0, which issues slab g+NS-1 right after it has released slab g.
template <int BJ, int NS>
struct StencilProducer {
  int tile, kk, ke, j0;
  unsigned x;  // index of the next slab to issue
  unsigned bytes;
  KB200_DEVICE_FUNCTION void open_tile(const StencilTmaParams& p, int n0) {
    const int tj = tile % p.tiles_j, tk = tile / p.tiles_j;
    const int k0 = 1 + tk * p.kc;
    j0 = 1 + tj * BJ;
    kk = k0 - 1;
    ke = min(k0 + p.kc, p.n2 - 1);                               // interior planes [k0, ke)
    const int rows = min(j0 + BJ, p.n1 - 1) - (j0 - 1) + 1;      // rows j0-1 .. min(j0+BJ, n1-1)
    bytes = (unsigned)rows * (unsigned)n0 * 8u;
  }
  KB200_DEVICE_FUNCTION void start(const StencilTmaParams& p, int n0) {
    tile = blockIdx.x; x = 0;
    if (tile < p.tiles_j * p.tiles_k) open_tile(p, n0);
  }
  // issue the next slab (waits until its stage's previous occupant was released by every consumer warp)
  KB200_DEVICE_FUNCTION void issue(const StencilTmaParams& p, int n0, double* stages, int stage_elems, unsigned long long* full,
                                   unsigned long long* empty) {
    if (tile >= p.tiles_j * p.tiles_k) return;
    if (KB200_STENCIL_DBG(p, 2)) { tile = p.tiles_j * p.tiles_k; return; }  // probe: no data movement
    const unsigned s = x % NS, use = x / NS;
    if (use > 0) ptx::mbar_wait(&empty[s], (use & 1u) ^ 1u);
    ptx::mbar_expect_tx(&full[s], bytes);
    ptx::bulk_g2s(stages + (size_t)s * stage_elems, p.u + (size_t)n0 * ((size_t)(j0 - 1) + (size_t)p.n1 * kk), bytes, &full[s]);
    ++x;
    if (++kk > ke) {
      tile += gridDim.x;
      if (tile < p.tiles_j * p.tiles_k) open_tile(p, n0);
    }
  }
};

template <int BJ, int NS, int CT, bool STORE, int N0T, bool PW>
__global__ void __launch_bounds__(CT + (PW ? 32 : 0), 1) stencil7_tma_kernel(const StencilTmaParams p, const ReduceScratch scratch) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(16) unsigned char red_smem[32 * sizeof(V)];
  constexpr int CW = CT / 32;
  constexpr int UPT = ((BJ / 2) * 256 + CT - 1) / CT;  // units per consumer thread at n0 = 512
  const int n0 = N0T ? N0T : p.n0, n1 = p.n1, n2 = p.n2;
  const int stage_elems = stencil_stage_elems(BJ, n0);
  const int pad = (BJ + 2) * n0;
  double* const stages = reinterpret_cast<double*>(dyn);
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(dyn + (size_t)NS * stage_elems * sizeof(double));
  unsigned long long* const empty = full + NS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], CW); }
    ptx::fence_mbar_init();
  }
  if (tid < NS) {
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    double* q = stages + (size_t)tid * stage_elems + pad;
    q[0] = qnan; q[1] = qnan; q[n0] = qnan; q[n0 + 1] = qnan;
  }
  if (KB200_STENCIL_DBG(p, 2))  // probe: arithmetic on resident pseudo-random data
    for (int q = tid; q < NS * stage_elems; q += blockDim.x) stages[q] = (double)((q * 2654435761u) >> 8) * (1.0 / 16777216.0);
  __syncthreads();
  const StencilRed red;
  V acc;
  red.init(acc);
  // Candidate thresholds from a sparse sample of the SAME values (true candidates with their locations, so the exact result is
  // unchanged): without them a smooth field, whose running extrema move with every plane, takes the re-evaluation path in
  // nearly every step (0.41 ms instead of 0.23 ms at 512^3).
  if (p.seed) acc = *p.seed;
  const int ntiles = p.tiles_j * p.tiles_k;

  StencilProducer<BJ, NS> prod;
  if (PW && warp == CW) {
    // ---------------- dedicated producer warp: one bulk copy per plane slab, as far ahead as the ring allows
    if (lane == 0) {
      prod.start(p, n0);
      while (prod.tile < ntiles) prod.issue(p, n0, stages, stage_elems, full, empty);
    }
  } else {
    // ---------------- consumers
    const int IV = n0 >> 1;
    const int nunits = (BJ / 2) * IV;
    StencilUnit un[UPT];
    unsigned actmask = 0;
#pragma unroll
    for (int m = 0; m < UPT; ++m) {
      const int q = tid + m * CT;
      const bool act = q < nunits;
      const int rg = act ? q / IV : 0;
      const int i0 = act ? 2 * (q - rg * IV) : 0;
      un[m].off = (2 * rg + 1) * n0 + i0;
      un[m].xl = i0 > 0 ? un[m].off - 1 : pad;
      un[m].xr = i0 + 2 < n0 ? un[m].off + 2 : pad;
      actmask |= act ? (1u << m) : 0u;
    }
    const double c0 = p.c0, c1 = p.c1;
    double tmin = acc.min_val, tmax = acc.max_val;  // warp-shared candidate thresholds (see stencil_unit_step)
    unsigned pos = 0;
    if (!PW && tid == 0) {  // every warp computes: thread 0 doubles as the producer, NS-1 slabs ahead of its own releases
      prod.start(p, n0);
      for (int q = 0; q < NS - 1; ++q) prod.issue(p, n0, stages, stage_elems, full, empty);
    }
#define KB200_STENCIL_RELEASE(S)                                                           \
    {                                                                                      \
      __syncwarp();                                                                        \
      if (lane == 0) ptx::mbar_arrive(&empty[S]);                                          \
      if (!PW && tid == 0) prod.issue(p, n0, stages, stage_elems, full, empty);            \
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tj = tile % p.tiles_j, tk = tile / p.tiles_j;
      const int j0 = 1 + tj * BJ, k0 = 1 + tk * p.kc;
      const int ke = min(k0 + p.kc, n2 - 1);
      const int nrows = min(BJ, n1 - 1 - j0);  // interior rows of this tile
      const int L = ke - k0 + 2;               // planes k0-1 .. ke
      const bool full_tile = nrows == BJ;
      double2 ra[UPT][2], rb[UPT][2], rc[UPT][2];  // three rotating plane buffers: no register moves between steps
      {  // k0-1 and k0 centres -> registers
        const unsigned s0 = pos % NS, s1 = (pos + 1) % NS;
        if (!KB200_STENCIL_DBG(p, 2)) ptx::mbar_wait(&full[s0], (pos / NS) & 1u);
        const double* P = stages + (size_t)s0 * stage_elems;
#pragma unroll
        for (int m = 0; m < UPT; ++m) {
          ra[m][0] = *reinterpret_cast<const double2*>(P + un[m].off);
          ra[m][1] = *reinterpret_cast<const double2*>(P + un[m].off + n0);
        }
        KB200_STENCIL_RELEASE(s0)
        if (!KB200_STENCIL_DBG(p, 2)) ptx::mbar_wait(&full[s1], ((pos + 1) / NS) & 1u);
        const double* Q = stages + (size_t)s1 * stage_elems;
#pragma unroll
        for (int m = 0; m < UPT; ++m) {
          rb[m][0] = *reinterpret_cast<const double2*>(Q + un[m].off);
          rb[m][1] = *reinterpret_cast<const double2*>(Q + un[m].off + n0);
        }
      }
      // one plane step: wait for plane k+1, run every unit, hand plane k's stage back
#define KB200_STENCIL_STEP(PREV, CUR, NXT)                                                                              \
      {                                                                                                                 \
        const int kk = k0 - 1 + t;                                                                                      \
        const unsigned scur = (pos + t) % NS, snext = (pos + t + 1) % NS;                                               \
        if (!KB200_STENCIL_DBG(p, 2)) ptx::mbar_wait(&full[snext], ((pos + t + 1) / NS) & 1u);                          \
        const double* P = stages + (size_t)scur * stage_elems;                                                          \
        const double* N = stages + (size_t)snext * stage_elems;                                                         \
        double* vplane = STORE ? p.vout + (size_t)n0 * ((size_t)j0 + (size_t)n1 * kk) : nullptr;                        \
        bool hit = false;                                                                                               \
        if (KB200_STENCIL_DBG(p, 1)) {                                                                                  \
        } else if (full_tile) {                                                                                         \
          _Pragma("unroll") for (int m = 0; m < UPT; ++m)                                                               \
            if ((actmask >> m) & 1u)                                                                                    \
              hit |= stencil_unit_step<STORE, true, N0T, false>(P, N, un[m], n0, pad, nrows, PREV[m], CUR[m], NXT[m], c0, c1, j0, kk, n1, n2, vplane, acc, tmin, tmax); \
        } else {                                                                                                        \
          _Pragma("unroll") for (int m = 0; m < UPT; ++m)                                                               \
            if ((actmask >> m) & 1u)                                                                                    \
              hit |= stencil_unit_step<STORE, false, N0T, false>(P, N, un[m], n0, pad, nrows, PREV[m], CUR[m], NXT[m], c0, c1, j0, kk, n1, n2, vplane, acc, tmin, tmax); \
        }                                                                                                               \
        /* one branch per STEP: the hot path stays a single basic block over all units (free interleaving) */         \
        if (__any_sync(kFullMask, hit)) {                                                                               \
          if (hit) {                                                                                                    \
            _Pragma("unroll") for (int m = 0; m < UPT; ++m)                                                             \
              if ((actmask >> m) & 1u)                                                                                  \
                stencil_unit_step<STORE, false, N0T, true>(P, N, un[m], n0, pad, nrows, PREV[m], CUR[m], NXT[m], c0, c1, j0, kk, n1, n2, vplane, acc, tmin, tmax); \
          }                                                                                                             \
          stencil_refresh_thresholds(acc, tmin, tmax);                                                                  \
        }                                                                                                               \
        KB200_STENCIL_RELEASE(scur)                                                                                     \
      }
      int t = 1;
      for (; t + 2 <= L - 2; t += 3) {
        KB200_STENCIL_STEP(ra, rb, rc)
        ++t;
        KB200_STENCIL_STEP(rb, rc, ra)
        ++t;
        KB200_STENCIL_STEP(rc, ra, rb)
        t -= 2;
      }
      if (t <= L - 2) {
        KB200_STENCIL_STEP(ra, rb, rc)
        ++t;
        if (t <= L - 2) KB200_STENCIL_STEP(rb, rc, ra)
      }
#undef KB200_STENCIL_STEP
      {  // the ke halo plane is done too
        const unsigned slast = (pos + L - 1) % NS;
        KB200_STENCIL_RELEASE(slast)
      }
      pos += (unsigned)L;
    }
#undef KB200_STENCIL_RELEASE
  }
  __syncthreads();
  block_reduce(red, acc, red_smem);
  __syncthreads();
  grid_reduce_and_store(red, acc, scratch, red_smem);
}

template <int BJ, int NS, int CT, int N0T, bool PW = (CT % 128 != 0)>
int launch_tma(b200_instance* I, const double* u, double* v_out, int n0, int n1, int n2, double c0, double c1,
               b200_minmaxloc_f64* rh, b200_minmaxloc_f64* rd) {
  const char* where = "b200_stencil7_minmaxloc_f64 (tma)";
  HostRuntime rt(I);
  auto kern = v_out ? stencil7_tma_kernel<BJ, NS, CT, true, N0T, PW> : stencil7_tma_kernel<BJ, NS, CT, false, N0T, PW>;
  const size_t smem = (size_t)NS * stencil_stage_elems(BJ, n0) * sizeof(double) + 2 * NS * sizeof(unsigned long long);
  static size_t smem_set_[64][2] = {};
  size_t& smem_set = smem_set_[I->device & 63][v_out ? 1 : 0];  // grow-only opt-in, like the reference's func-attr cache (KernelLaunch.hpp:131-145)
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return b200_report_error((int)e, where);
    smem_set = smem;
  }
  const int sms = rt.sm_count();
  const int tiles_j = (n1 - 2 + BJ - 1) / BJ;
  // planes per tile: minimise (waves of the persistent grid) x (planes loaded per tile, incl. the two k halos)
  int kc = b200_tune("stencil.kc", 0);
  if (kc <= 0) {
    long best = -1;
    for (int c = 4; c <= 256; ++c) {
      const long tk = (n2 - 2 + c - 1) / c, nt = tk * tiles_j;
      const long g = nt < sms ? nt : sms;
      const long cost = ((nt + g - 1) / g) * (long)((c < n2 - 2 ? c : n2 - 2) + 2);
      if (best < 0 || cost < best) { best = cost; kc = c; }
    }
  }
  const int tiles_k = (n2 - 2 + kc - 1) / kc;
  const long ntiles = (long)tiles_j * tiles_k;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  // sampling pre-pass: every 16th row of every 16th plane (5 rows read per sampled row: ~2 % of the field)
  constexpr int SBLOCK = 256, SUNROLL = 4, SSTEP = 16;
  const long srows = (long)((n1 - 2 + SSTEP - 1) / SSTEP) * ((n2 - 2 + SSTEP - 1) / SSTEP);
  const bool sample = b200_tune("stencil.seed", 1) && (long)(n1 - 2) * (n2 - 2) >= 16384;
  long sgrid = (srows + SBLOCK / 32 - 1) / (SBLOCK / 32);
  if (sgrid > 4L * sms) sgrid = 4L * sms;
  const size_t nparts = (size_t)(sample && sgrid > grid ? sgrid : grid);
  ReduceScratch s;
  void *slot_dev = nullptr, *slot_host = nullptr;
  int rc;
  if ((rc = rt.reduce_scratch((nparts + 1) * sizeof(V), sizeof(V), rh != nullptr, &s.partials, &s.ticket, &slot_dev, &slot_host))) return rc;
  V* const seed = reinterpret_cast<V*>(s.partials) + nparts;  // one slot behind the per-block partials
  if (sample) {
    ReduceScratch ss = s;
    ss.result0 = seed;
    ss.result1 = nullptr;
    stencil7_minmaxloc_kernel<SBLOCK, SUNROLL><<<(unsigned)sgrid, SBLOCK, 0, rt.stream()>>>(u, nullptr, n0, n1, n2, c0, c1, ss, SSTEP, SSTEP);
    if ((rc = rt.check_launch("b200_stencil7_minmaxloc_f64 (sample)"))) return rc;
  }
  s.result0 = rh ? slot_dev : (void*)rd;
  s.result1 = rh ? (void*)rd : nullptr;
  StencilTmaParams p{u, v_out, n0, n1, n2, kc, tiles_j, tiles_k, c0, c1, sample ? seed : nullptr, b200_tune("stencil.dbg", 0)};
  kern<<<grid, CT + (PW ? 32 : 0), smem, rt.stream()>>>(p, s);
  if ((rc = rt.check_launch(where))) return rc;
  if (rh) {
    if ((rc = rt.fence(where))) return rc;
    memcpy(rh, slot_host, sizeof(V));
  }
  return 0;
}
}  // namespace

extern "C" int b200_stencil7_minmaxloc_f64(b200_instance* I, const double* u, double* v_out, int64_t n0, int64_t n1, int64_t n2,
                                           double c0, double c1, b200_minmaxloc_f64* rh, b200_minmaxloc_f64* rd) {
  const char* where = "b200_stencil7_minmaxloc_f64";
  B200_CHECK_INST(I, where);
  if (n0 < 0 || n1 < 0 || n2 < 0) return b200_set_error(B200_EINVAL, where, "negative extent");
  if (!rh && !rd) return b200_set_error(B200_EINVAL, where, "no result destination");
  const bool empty = (n0 < 3 || n1 < 3 || n2 < 3);
  if (!empty && !u) return b200_set_error(B200_EINVAL, where, "u is NULL");
  HostRuntime rt(I);
  int rc;
  ReduceScratch s;
  void *slot_dev = nullptr, *slot_host = nullptr;
  const bool fast = !empty && b200_tune("stencil.tma", 1) && (n0 % 2 == 0) && n0 >= 4 && n0 <= 512 &&
                    (reinterpret_cast<uintptr_t>(u) % 16 == 0) && (reinterpret_cast<uintptr_t>(v_out) % 16 == 0);
  if (fast) {
    const int ct = b200_tune("stencil.ct", 352), ns = b200_tune("stencil.ns", 5);
    rc = B200_EUNSUPPORTED;
#define TMA_CFG(BJ, NS, CT)                                                                                              \
  if (ct == CT && ns == NS)                                                                                              \
    rc = n0 == 512 ? launch_tma<BJ, NS, CT, 512>(I, u, v_out, (int)n0, (int)n1, (int)n2, c0, c1, rh, rd)                 \
                   : launch_tma<BJ, NS, CT, 0>(I, u, v_out, (int)n0, (int)n1, (int)n2, c0, c1, rh, rd);
    TMA_CFG(8, 5, 512) TMA_CFG(8, 5, 352) TMA_CFG(8, 5, 224)
#ifdef B200_SWEEP
    TMA_CFG(8, 4, 512) TMA_CFG(8, 4, 352) TMA_CFG(8, 5, 384) TMA_CFG(8, 5, 256)
#endif
#undef TMA_CFG
    if (rc == B200_EUNSUPPORTED) return b200_set_error(rc, where, "tuning combination not compiled in");
    return rc;
  }
  constexpr int BLOCK = 256, UNROLL = 4;
  static int bps = 0;
  if (!bps) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, stencil7_minmaxloc_kernel<BLOCK, UNROLL>, BLOCK, 0);
    if (bps < 1) bps = 1;
  }
  const int64 rows = empty ? 0 : (int64)(n1 - 2) * (n2 - 2);
  int64 blocks = (rows + BLOCK / 32 - 1) / (BLOCK / 32);
  const int cap = b200_tune("stencil.bps", 0);
  const int64 max_grid = (int64)rt.sm_count() * ((cap > 0 && cap < bps) ? cap : bps);
  const int grid = (int)(blocks < 1 ? 1 : (blocks < max_grid ? blocks : max_grid));
  if ((rc = rt.reduce_scratch((size_t)grid * sizeof(V), sizeof(V), rh != nullptr, &s.partials, &s.ticket, &slot_dev, &slot_host))) return rc;
  s.result0 = rh ? slot_dev : (void*)rd;
  s.result1 = rh ? (void*)rd : nullptr;
  stencil7_minmaxloc_kernel<BLOCK, UNROLL><<<grid, BLOCK, 0, rt.stream()>>>(u, v_out, empty ? 2 : n0, empty ? 2 : n1, empty ? 2 : n2, c0, c1, s);
  if ((rc = rt.check_launch(where))) return rc;
  if (rh) {
    if ((rc = rt.fence(where))) return rc;
    memcpy(rh, slot_host, sizeof(V));
  }
  return 0;
}
