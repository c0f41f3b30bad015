"""kokkos_b200.sharded -- one execution-space instance per GPU, index ranges sharded across the ranks of a box.

The reference has no multi-GPU collectives of its own: a Kokkos program gets one ``Kokkos::Cuda`` instance per device
(core/src/Cuda/Kokkos_Cuda_Instance.cpp:268-283, core/unit_test/cuda/TestCuda_MultiGPU / TestMultiGPU.hpp) and combines
per-device results itself.  This module is that combine step for the hot path, one process per
GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) as plumbing only:

  parallel_for  (stream / GUPS on local tables / SpMV rows)   shard [k*N/g, (k+1)*N/g); no exchange
  parallel_reduce  Sum / Min / Max / MinMax                    local partial (device scalar) -> ONE all_reduce
  parallel_reduce  MinLoc / MaxLoc / MinMaxLoc                 local partial -> all_gather of the value struct -> join in
                                                               RANK ORDER with the reference's join rule
                                                               (core/src/Kokkos_Parallel_Reduce.hpp:441-449,628-644), so
                                                               equal extrema keep the lowest-ranked (= lowest index) location
  parallel_scan  (contiguous shards)                           local total -> all_gather (8 B per rank) -> the seeded local
                                                               scan sums the lower ranks' totals itself on the device
                                                               (b200_scan_excl_i64_seeds_dev): reduce-then-scan, 24 B/element
  parallel_scan  (block-cyclic shards, needs `comm`)           ONE fused kernel per rank (b200_comm_scan_*): round aggregates
                                                               travel over peer-mapped NVLink mailboxes, 16 B/element at any N

With `comm` (kokkos_b200.Comm, the library's own one-box communicator) the few-byte combines run as device-side,
stream-ordered kernels of this library (no host round trip); without it they go through torch.distributed.

The local work is done by a *local executor*: on a GPU box that is always ``kokkos_b200.B200`` (there is no CPU
fallback in this package); the world_size-2 gloo tests inject a stand-in built on the test oracle to exercise the
partitioning / combine logic without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

INDEX_IDENTITY = np.iinfo(np.int64).max  # reduction_identity<int64_t>::min() (Kokkos_ReductionIdentity.hpp:357-380)


def shard_bounds(n: int, world: int, rank: int, align: int = 1) -> tuple[int, int]:
    """Contiguous shard [begin, end) of [0, n) owned by `rank`: cut points k*n/world rounded DOWN to `align` elements
    (align keeps every shard start 16/32-byte aligned for the vector/TMA kernels).  Shards are ordered by rank, cover
    [0, n) exactly, and differ in length by less than `align + 1`."""
    if world < 1 or not (0 <= rank < world) or n < 0 or align < 1:
        raise ValueError("shard_bounds: bad arguments")

    def cut(k: int) -> int:
        if k >= world:
            return n
        return ((k * n) // world) // align * align
    return cut(rank), cut(rank + 1)


def cyclic_blocks(n_global: int, block: int, world: int, rank: int) -> list[tuple[int, int]]:
    """Global [begin, end) ranges owned by `rank` under the block-cyclic distribution of the fused scan
    (kb200/impl/ScanChunked.hpp): global block c lives on rank c % world as local block c // world; local blocks are stored
    back to back in that order.  Only the last global block may be short."""
    if world < 1 or not (0 <= rank < world) or n_global < 0 or block < 1:
        raise ValueError("cyclic_blocks: bad arguments")
    nblocks = -(-n_global // block)
    return [(c * block, min(n_global, (c + 1) * block)) for c in range(rank, nblocks, world)]


def cyclic_take(global_array: np.ndarray, block: int, world: int, rank: int) -> np.ndarray:
    """The local part (owned blocks, concatenated) of a global array."""
    parts = [global_array[b:e] for b, e in cyclic_blocks(global_array.size, block, world, rank)]
    return np.concatenate(parts) if parts else global_array[:0].copy()


def cyclic_put(global_out: np.ndarray, local: np.ndarray, block: int, world: int, rank: int) -> None:
    """Inverse of cyclic_take: scatter a rank's local part into the global array."""
    off = 0
    for b, e in cyclic_blocks(global_out.size, block, world, rank):
        global_out[b:e] = local[off:off + (e - b)]
        off += e - b


def join_minloc(dest: tuple, src: tuple) -> tuple:
    """MinLoc::join (Kokkos_Parallel_Reduce.hpp:441-449) on (val, loc) pairs."""
    if src[0] < dest[0]:
        return src
    if src[0] == dest[0] and dest[1] == INDEX_IDENTITY:
        return (dest[0], src[1])
    return dest


def join_maxloc(dest: tuple, src: tuple) -> tuple:
    """MaxLoc::join (Kokkos_Parallel_Reduce.hpp:501-509)."""
    if src[0] > dest[0]:
        return src
    if src[0] == dest[0] and dest[1] == INDEX_IDENTITY:
        return (dest[0], src[1])
    return dest


@dataclass
class MinMaxLocResult:
    min_val: float
    max_val: float
    min_loc: int
    max_loc: int


class ShardedB200:
    """Rank-local handle of a range-sharded execution over `world` B200s (one process per GPU)."""

    def __init__(self, local, group=None, coll_device=None, comm=None):
        """`local`: the rank's execution-space instance (kokkos_b200.B200).  `group`: torch.distributed process group
        (default: WORLD; None with no initialised backend = single GPU).  `coll_device`: where the few-byte collective
        buffers live (the GPU for NCCL, CPU for gloo).  `comm`: optional kokkos_b200.Comm over the same ranks."""
        self.local = local
        self.group = group
        self.comm = comm
        # Stream order: the local kernels run on the instance's stream, torch.distributed collectives and `.cpu()` on torch's
        # CURRENT stream.  Nothing else orders them, so the two must be the same stream (build the instance with
        # kb.B200(device, stream=torch.cuda.current_stream().cuda_stream), as bench.py does).
        if (dist.is_available() and dist.is_initialized() and torch.cuda.is_available() and getattr(local, "stream", None)
                and coll_device is not None and torch.device(coll_device).type == "cuda"):
            cur = torch.cuda.current_stream(torch.device(coll_device)).cuda_stream
            if int(local.stream) != int(cur):
                raise ValueError("ShardedB200: the execution-space instance must be created on torch's current CUDA stream "
                                 "(its kernels and the collectives are ordered by that stream only)")
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if coll_device is None:
            coll_device = torch.device("cuda", local.device) if getattr(local, "device", None) is not None and torch.cuda.is_available() else torch.device("cpu")
        self.dev = coll_device
        # persistent few-byte buffers: no allocation on the hot path
        self._f64 = torch.zeros(8, dtype=torch.float64, device=self.dev)
        self._i64 = torch.zeros(8, dtype=torch.int64, device=self.dev)
        self._gather_i64 = torch.zeros(self.world, dtype=torch.int64, device=self.dev)

    # ------------------------------------------------------------------ partition
    def shard(self, n_global: int, align: int = 4) -> tuple[int, int]:
        return shard_bounds(n_global, self.world, self.rank, align)

    # ------------------------------------------------------------------ parallel_reduce
    def _scalar_buf(self, dtype):
        return self._f64 if np.dtype(dtype).kind == "f" else self._i64

    def reduce_sum_async(self, view, out: torch.Tensor | None = None) -> torch.Tensor:
        """Global Sum of a sharded View: local kernel writes its partial into a device scalar, one all_reduce combines.
        Nothing blocks the host; the returned 1-element tensor is valid in stream order."""
        buf = out if out is not None else self._scalar_buf(view.dtype)[:1]
        self.local.parallel_reduce_sum(view, result_dev=buf.data_ptr(), blocking=False)
        if self.world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf

    def parallel_reduce_sum(self, view):
        return self.reduce_sum_async(view).cpu()[0].item()  # scalar result => fence (Kokkos_Parallel_Reduce.hpp:1592-1638)

    def _reduce_minmax(self, view, which: str):
        buf = self._scalar_buf(view.dtype)[:1]
        getattr(self.local, f"parallel_reduce_{which}")(view, result_dev=buf.data_ptr(), blocking=False)
        if self.world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.MIN if which == "min" else dist.ReduceOp.MAX, group=self.group)
        return buf.cpu()[0].item()

    def parallel_reduce_min(self, view):
        return self._reduce_minmax(view, "min")

    def parallel_reduce_max(self, view):
        return self._reduce_minmax(view, "max")

    def parallel_reduce_minmaxloc(self, view, index_base: int) -> MinMaxLocResult:
        """MinMaxLoc over a sharded View<double*>; `index_base` = global index of the shard's first element.
        The 32-byte partials are all-gathered and joined in rank order on every rank (deterministic; ties keep the
        lowest-ranked location, i.e. the lowest index -- what the OpenMP reference produces with its thread-ordered
        joins, core/src/OpenMP/Kokkos_OpenMP_Parallel_Reduce.hpp:147-151)."""
        r = self.local.parallel_reduce_minmaxloc(view, index_base)
        mine = np.zeros(1, dtype=[("min_val", "<f8"), ("max_val", "<f8"), ("min_loc", "<i8"), ("max_loc", "<i8")])
        mine[0] = (r.min_val, r.max_val, r.min_loc, r.max_loc)
        parts = self._allgather_struct(mine)
        mn = (float(parts[0]["min_val"]), int(parts[0]["min_loc"]))
        mx = (float(parts[0]["max_val"]), int(parts[0]["max_loc"]))
        for p in parts[1:]:
            mn = join_minloc(mn, (float(p["min_val"]), int(p["min_loc"])))
            mx = join_maxloc(mx, (float(p["max_val"]), int(p["max_loc"])))
        return MinMaxLocResult(mn[0], mx[0], mn[1], mx[1])

    def parallel_reduce_minloc(self, view, index_base: int) -> tuple:
        r = self.local.parallel_reduce_minloc(view, index_base)
        mine = np.zeros(1, dtype=[("val", "<f8"), ("loc", "<i8")])
        mine[0] = (r.val, r.loc)
        parts = self._allgather_struct(mine)
        acc = (float(parts[0]["val"]), int(parts[0]["loc"]))
        for p in parts[1:]:
            acc = join_minloc(acc, (float(p["val"]), int(p["loc"])))
        return acc

    def parallel_reduce_maxloc(self, view, index_base: int) -> tuple:
        r = self.local.parallel_reduce_maxloc(view, index_base)
        mine = np.zeros(1, dtype=[("val", "<f8"), ("loc", "<i8")])
        mine[0] = (r.val, r.loc)
        parts = self._allgather_struct(mine)
        acc = (float(parts[0]["val"]), int(parts[0]["loc"]))
        for p in parts[1:]:
            acc = join_maxloc(acc, (float(p["val"]), int(p["loc"])))
        return acc

    def _allgather_struct(self, mine: np.ndarray) -> np.ndarray:
        """all_gather of one small POD struct per rank, returned in rank order (bytes travel as uint8)."""
        if self.world == 1:
            return mine
        nb = mine.dtype.itemsize
        src = torch.from_numpy(mine.view(np.uint8).copy()).to(self.dev)
        dst = torch.empty(nb * self.world, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(dst, src, group=self.group)
        return dst.cpu().numpy().view(mine.dtype)

    # ------------------------------------------------------------------ parallel_scan
    def scan_exclusive_async(self, x, y, total_out: torch.Tensor | None = None, around_scan_kernel=None) -> torch.Tensor:
        """Range-sharded exclusive prefix sum of int64 Views (config C3): local total -> all_gather -> seeded local scan.
        Fully stream-ordered (no host synchronisation).  Returns the all-gathered shard totals (length `world`); their
        sum is the global total.  DRAM traffic per element: 8 (totals pass) + 16 (scan) = 24 B on world > 1.
        `around_scan_kernel` = (before, after) callables invoked immediately around the scan-kernel launch (bench.py
        records its roofline events there)."""
        before, after = around_scan_kernel if around_scan_kernel else (None, None)
        tot = total_out if total_out is not None else self._i64[1:2]
        if self.world == 1:
            tot = total_out if total_out is not None else self._gather_i64
            if before:
                before()
            self.local.parallel_scan(x, y, total_dev=tot.data_ptr(), blocking=False)
            if after:
                after()
            return tot
        mine = self._i64[:1]
        self.local.parallel_reduce_sum(x, result_dev=mine.data_ptr(), blocking=False)
        dist.all_gather_into_tensor(self._gather_i64, mine, group=self.group)
        if before:
            before()
        self.local.parallel_scan_seeds_dev(x, y, self._gather_i64.data_ptr(), self.rank, tot.data_ptr())
        if after:
            after()
        return self._gather_i64

    def parallel_scan(self, x, y) -> int:
        """Blocking form: returns the GLOBAL total (ParallelScanWithTotal semantics, Kokkos_Parallel.hpp:405-425)."""
        totals = self.scan_exclusive_async(x, y)
        return int(totals.cpu().sum().item())

    # ------------------------------------------------------------------ MDRange stencil (k-slab sharding)
    def stencil7_minmaxloc(self, u_slab, n0: int, n1: int, n2_local: int, n2_global: int, k_offset: int, c0: float, c1: float) -> MinMaxLocResult:
        """Config C4 sharded along the slowest dimension: this rank holds planes [k_offset, k_offset + n2_local) of a
        LayoutLeft n0 x n1 x n2_global field INCLUDING one halo plane on each side that has a neighbour.  Locations are
        re-based from the slab's (i*n1+j)*n2_local+k to the global (i*n1+j)*n2_global+(k+k_offset) before the rank-ordered
        join."""
        r = self.local.stencil7_minmaxloc(u_slab, n0, n1, n2_local, c0, c1)

        def rebase(loc: int) -> int:
            if loc == INDEX_IDENTITY:
                return loc
            ij, k = divmod(loc, n2_local)
            return ij * n2_global + k + k_offset
        mine = np.zeros(1, dtype=[("min_val", "<f8"), ("max_val", "<f8"), ("min_loc", "<i8"), ("max_loc", "<i8")])
        mine[0] = (r.min_val, r.max_val, rebase(r.min_loc), rebase(r.max_loc))
        parts = self._allgather_struct(mine)
        # slabs are ordered by k, but loc order is (i, j, k): a tie between ranks must keep the LOWEST loc, which a rank-ordered
        # fold does not give here -- resolve ties explicitly by location
        mn = min(((float(p["min_val"]), int(p["min_loc"])) for p in parts), key=lambda t: (t[0], t[1]))
        mx = min(((-float(p["max_val"]), int(p["max_loc"])) for p in parts), key=lambda t: (t[0], t[1]))
        return MinMaxLocResult(mn[0], -mx[0], mn[1], mx[1])

    def barrier(self) -> None:
        if self.world > 1:
            dist.barrier(group=self.group)

    # ------------------------------------------------------------------ device-resident (asynchronous) forms
    def minmaxloc_async(self, view, index_base: int, out: torch.Tensor) -> torch.Tensor:
        """MinMaxLoc over a contiguous-sharded View<double*>: the local kernel writes {min_val, max_val, min_loc, max_loc}
        (32 bytes; `out` = 4-element float64 tensor, the locations are int64 bit patterns) and the rank partials are joined
        on the device in rank order, lowest location on ties (b200_allreduce_minmaxloc_f64).  Needs `comm` when world > 1."""
        self.local.parallel_reduce_minmaxloc_dev(view, index_base, out.data_ptr())
        if self.world > 1:
            self._need_comm("minmaxloc_async")
            self.comm.allreduce_loc("minmaxloc", out.data_ptr())
        return out

    def stencil7_minmaxloc_async(self, u_slab, n0: int, n1: int, n2_local: int, n2_global: int, k_offset: int, c0: float, c1: float,
                                 out: torch.Tensor) -> torch.Tensor:
        """Asynchronous k-slab form of stencil7_minmaxloc: result and re-basing stay on the device (`out`: 4 x float64 as above)."""
        self.local.stencil7_minmaxloc_dev(u_slab, n0, n1, n2_local, c0, c1, out.data_ptr())
        if self.world > 1 or n2_local != n2_global or k_offset:
            loc = out[2:4].view(torch.int64)
            ident = loc == INDEX_IDENTITY
            reb = torch.div(loc, n2_local, rounding_mode="floor") * n2_global + loc % n2_local + k_offset
            loc.copy_(torch.where(ident, loc, reb))
        if self.world > 1:
            self._need_comm("stencil7_minmaxloc_async")
            self.comm.allreduce_loc("minmaxloc", out.data_ptr())
        return out

    def cyclic_scan_async(self, x_local, y_local, n_global: int, total_out: torch.Tensor, inclusive: bool = False) -> None:
        """Fused distributed parallel_scan over block-cyclic Views (see cyclic_blocks): one kernel per rank, 16 B/element."""
        if self.world == 1 and self.comm is None:
            self.local.parallel_scan(x_local, y_local, inclusive=inclusive, total_dev=total_out.data_ptr(), blocking=False)
            return
        self._need_comm("cyclic_scan_async")
        self.comm.parallel_scan(x_local, y_local, n_global, inclusive=inclusive, total_dev=total_out.data_ptr(), blocking=False)

    def _need_comm(self, what: str) -> None:
        if self.comm is None:
            raise ValueError(f"ShardedB200.{what}: world > 1 needs the library communicator (pass comm=kokkos_b200.Comm(...))")

    # ------------------------------------------------------------------ parallel_for: shard-local, no exchange
    def stream_copy(self, a, c) -> None:
        """benchmarks/stream copy on this rank's shard: c = a (no exchange; stream-kokkos.cpp:217-222)."""
        self.local.stream_copy(a, c)

    def stream_triad(self, a, b, c, scalar: float) -> None:
        """benchmarks/stream triad on this rank's shard: a = b + scalar * c (stream-kokkos.cpp:226-231)."""
        self.local.stream_triad(a, b, c, scalar)

    def gups(self, table_shard, indices_local, datum: int, op: str = "add") -> None:
        """GUPS on a table sharded by index range: this rank owns table[rank*L, (rank+1)*L) and applies the updates whose
        targets fall in its shard (`indices_local` are shard-relative).  The update stream is generated per owner (see
        gups_owner_indices), so no update crosses ranks: weak scaling with no exchange (benchmarks/gups/gups-kokkos.cpp)."""
        self.local.gups(table_shard, indices_local, datum, op)

    def spmv_rows(self, row_map_local, col_idx, values, x_full, y_local) -> None:
        """CRS SpMV sharded by rows: this rank holds rows [r0, r1) (row_map re-based to 0) and a full copy of x; y stays
        shard-local (no exchange)."""
        self.local.spmv_crs(row_map_local, col_idx, values, x_full, y_local)
