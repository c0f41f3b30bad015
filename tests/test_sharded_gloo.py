"""world_size-2 (and 3) gloo tests of the range-sharded multi-GPU layer (kokkos_b200/sharded.py), run on CPU.

The product's local executor is the CUDA library (no CPU fallback), so here the ranks get a stand-in executor built on
the TEST ORACLE (oracle/liboracle_port.so) with the same method names: what is under test is the partitioning, the
collective plumbing and the rank-ordered joins -- the code that runs unchanged over NCCL on the GPU box
(bench.py --gpus N).  Expected values come from the oracle run over the WHOLE (unsharded) input.
"""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class HostView:
    def __init__(self, arr):
        self.array = np.ascontiguousarray(arr)
        self.dtype = self.array.dtype
        self.n = self.array.size


def _store(ptr, ctype, value):
    ctypes.cast(ptr, ctypes.POINTER(ctype))[0] = value


class OracleLocal:
    """Stand-in for kokkos_b200.B200 in CPU tests: same call surface, computed by the oracle port."""
    device = None

    def __init__(self):
        from oracle.bindings import Port
        self.port = Port()

    def _reduce(self, op, v, result_dev, blocking):
        r = self.port.reduce(op, v.array, 1)
        if result_dev:
            _store(result_dev, ctypes.c_double if v.dtype.kind == "f" else ctypes.c_int64, r)
        return r if blocking else None

    def parallel_reduce_sum(self, v, result_dev=0, blocking=True):
        return self._reduce("sum", v, result_dev, blocking)

    def parallel_reduce_min(self, v, result_dev=0, blocking=True):
        return self._reduce("min", v, result_dev, blocking)

    def parallel_reduce_max(self, v, result_dev=0, blocking=True):
        return self._reduce("max", v, result_dev, blocking)

    def parallel_reduce_minmaxloc(self, v, index_base=0):
        return self.port.reduce_loc("minmaxloc", v.array, index_base, 1)

    def parallel_reduce_minloc(self, v, index_base=0):
        return self.port.reduce_loc("minloc", v.array, index_base, 1)

    def parallel_reduce_maxloc(self, v, index_base=0):
        return self.port.reduce_loc("maxloc", v.array, index_base, 1)

    def parallel_scan(self, x, y, inclusive=False, seed=0, total_dev=0, blocking=True):
        out, total = self.port.scan(x.array, inclusive, seed, 1)
        y.array[:] = out
        if total_dev:
            _store(total_dev, ctypes.c_int64, total)
        return total if blocking else None

    def parallel_scan_seeds_dev(self, x, y, seeds_dev, nseeds, total_dev=0):
        seeds = ctypes.cast(seeds_dev, ctypes.POINTER(ctypes.c_int64))
        seed = int(np.sum(np.array([seeds[k] for k in range(nseeds)], dtype=np.int64)))  # wraps mod 2^64 like the kernel
        out, total = self.port.scan(x.array, False, seed, 1)
        y.array[:] = out
        if total_dev:
            _store(total_dev, ctypes.c_int64, total)

    def stencil7_minmaxloc(self, u, n0, n1, n2, c0, c1, v_out=None):
        r, _ = self.port.stencil7(u.array, n0, n1, n2, c0, c1)
        return r

    # ---- device-result (asynchronous) forms: "device memory" is host memory here
    @staticmethod
    def _store_mml(ptr, r):
        d = ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double))
        q = ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64))
        d[0], d[1], q[2], q[3] = r.min_val, r.max_val, r.min_loc, r.max_loc

    def parallel_reduce_minmaxloc_dev(self, v, index_base, result_dev):
        self._store_mml(result_dev, self.port.reduce_loc("minmaxloc", v.array, index_base, 1))

    def stencil7_minmaxloc_dev(self, u, n0, n1, n2, c0, c1, result_dev, v_out=None):
        self._store_mml(result_dev, self.port.stencil7(u.array, n0, n1, n2, c0, c1)[0])

    # ---- shard-local parallel_for bodies
    def stream_copy(self, a, b):
        b.array[:] = a.array

    def stream_triad(self, a, b, c, s):
        a.array[:] = b.array + s * c.array

    def gups(self, table, indices, datum, op="add"):
        assert op == "add"
        np.add.at(table.array, indices.array, datum)

    def spmv_crs(self, row_map, col_idx, values, x, y):
        y.array[:] = self.port.spmv(row_map.array, col_idx.array, values.array, x.array)


class GlooComm:
    """Stand-in for kokkos_b200.Comm over gloo: the same collective semantics (rank-ordered joins, lowest location on ties;
    block-cyclic distributed scan) computed on the host, so ShardedB200's partition / re-basing logic runs without a GPU."""
    BLOCK = 96  # elements per block of the block-cyclic distribution (the real one is tiles-per-round x tile)

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def allreduce_loc(self, kind, buf_ptr):
        assert kind == "minmaxloc"
        mine = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(buf_ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(32,)).copy())
        parts = [torch.empty(32, dtype=torch.uint8) for _ in range(self.world)]
        dist.all_gather(parts, mine)
        rec = np.concatenate([p.numpy() for p in parts]).view([("mn", "<f8"), ("mx", "<f8"), ("lmn", "<i8"), ("lmx", "<i8")])
        mn = min(((float(r["mn"]), int(r["lmn"])) for r in rec), key=lambda t: (t[0], t[1]))
        mx = min(((-float(r["mx"]), int(r["lmx"])) for r in rec), key=lambda t: (t[0], t[1]))
        d = ctypes.cast(buf_ptr, ctypes.POINTER(ctypes.c_double))
        q = ctypes.cast(buf_ptr, ctypes.POINTER(ctypes.c_int64))
        d[0], d[1], q[2], q[3] = mn[0], -mx[0], mn[1], mx[1]

    def cyclic_layout(self, n_global, dtype):
        from kokkos_b200.sharded import cyclic_blocks
        blocks = cyclic_blocks(n_global, self.BLOCK, self.world, self.rank)
        return self.BLOCK, sum(e - b for b, e in blocks), -(-n_global // (self.BLOCK * self.world))

    def parallel_scan(self, x, y, n_global, inclusive=False, total_dev=0, blocking=True):
        from kokkos_b200.sharded import cyclic_put, cyclic_take
        nl_max = -(-n_global // (self.BLOCK * self.world)) * self.BLOCK
        mine = torch.zeros(nl_max, dtype=torch.int64)
        mine[: x.n] = torch.from_numpy(x.array)
        parts = [torch.empty(nl_max, dtype=torch.int64) for _ in range(self.world)]
        dist.all_gather(parts, mine)
        g = np.zeros(n_global, dtype=np.int64)
        for q in range(self.world):
            from kokkos_b200.sharded import cyclic_blocks
            nq = sum(e - b for b, e in cyclic_blocks(n_global, self.BLOCK, self.world, q))
            cyclic_put(g, parts[q].numpy()[:nq], self.BLOCK, self.world, q)
        with np.errstate(over="ignore"):
            incl = np.cumsum(g, dtype=np.int64)
        res = incl if inclusive else incl - g
        y.array[:] = cyclic_take(res, self.BLOCK, self.world, self.rank)
        total = int(incl[-1]) if n_global else 0
        if total_dev:
            _store(total_dev, ctypes.c_int64, total)
        return total if blocking else None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import workloads as W
        from kokkos_b200.sharded import ShardedB200
        from oracle.bindings import Port
        sp = ShardedB200(OracleLocal(), coll_device=torch.device("cpu"))
        whole = Port()
        out = {}
        if case == "reduce":
            n = 100003
            x = W.c1_exact(n)
            b, e = sp.shard(n)
            out["sum"] = (sp.parallel_reduce_sum(HostView(x[b:e])), whole.reduce("sum", x, 1))
            xi = W.c3_wrap(n)
            out["isum"] = (sp.parallel_reduce_sum(HostView(xi[b:e])), whole.reduce("sum", xi, 1))
            g = W.c1_uniform(n)
            out["min"] = (sp.parallel_reduce_min(HostView(g[b:e])), whole.reduce("min", g, 1))
            out["max"] = (sp.parallel_reduce_max(HostView(g[b:e])), whole.reduce("max", g, 1))
            r = sp.parallel_reduce_minmaxloc(HostView(g[b:e]), b)
            w = whole.reduce_loc("minmaxloc", g, 0, 1)
            out["mml"] = ((r.min_val, r.max_val, r.min_loc, r.max_loc), (w.min_val, w.max_val, w.min_loc, w.max_loc))
            # ties across ranks: constant array -> both extrema are everywhere; lowest index (rank 0's first) must win
            c = np.full(n, 2.5)
            r = sp.parallel_reduce_minmaxloc(HostView(c[b:e]), b)
            out["ties"] = ((r.min_val, r.max_val, r.min_loc, r.max_loc), (2.5, 2.5, 0, 0))
            r = sp.parallel_reduce_minloc(HostView(g[b:e]), b)
            w = whole.reduce_loc("minloc", g, 0, 1)
            out["minloc"] = (tuple(r), (w.val, w.loc))
            r = sp.parallel_reduce_maxloc(HostView(g[b:e]), b)
            w = whole.reduce_loc("maxloc", g, 0, 1)
            out["maxloc"] = (tuple(r), (w.val, w.loc))
        elif case == "scan":
            for n in (100003, 7, 1 << 16):
                for gen in (W.c3_small, W.c3_wrap):
                    x = gen(n)
                    b, e = sp.shard(n, align=2)
                    y = HostView(np.zeros(e - b, dtype=np.int64))
                    total = sp.parallel_scan(HostView(x[b:e]), y)
                    wy, wt = whole.scan(x, False, 0, 1)
                    out[f"scan{n}{gen.__name__}"] = ((total, y.array.tobytes()), (wt, wy[b:e].tobytes()))
        elif case == "empty_shards":
            n = 1  # fewer elements than ranks: some shards are empty and must contribute identities
            x = np.array([5.0])
            b, e = sp.shard(n, align=1)
            out["sum"] = (sp.parallel_reduce_sum(HostView(x[b:e])), 5.0)
            r = sp.parallel_reduce_minmaxloc(HostView(x[b:e]), b)
            out["mml"] = ((r.min_val, r.max_val, r.min_loc, r.max_loc), (5.0, 5.0, 0, 0))
            xi = np.array([9], dtype=np.int64)
            y = HostView(np.zeros(e - b, dtype=np.int64))
            out["scan_total"] = (sp.parallel_scan(HostView(xi[b:e]), y), 9)
        elif case == "stencil":
            n0, n1, n2 = 20, 12, 19
            u, _, _ = W.c4_field(n0, n1, n2)
            u3 = u.reshape((n0, n1, n2), order="F")
            # interior planes 1..n2-2 split by rank; each slab carries one halo plane per side
            kb_, ke_ = [(1 + (k * (n2 - 2)) // world) for k in (rank, rank + 1)]
            slab = np.asfortranarray(u3[:, :, kb_ - 1:ke_ + 1]).reshape(-1, order="F").copy()
            r = sp.stencil7_minmaxloc(HostView(slab), n0, n1, ke_ - kb_ + 2, n2, kb_ - 1, 0.5, 0.125)
            w, _ = whole.stencil7(u, n0, n1, n2, 0.5, 0.125)
            out["stencil"] = ((r.min_val, r.max_val, r.min_loc, r.max_loc), (w.min_val, w.max_val, w.min_loc, w.max_loc))
        elif case == "async_and_for":
            from kokkos_b200.sharded import cyclic_take
            spc = ShardedB200(OracleLocal(), coll_device=torch.device("cpu"), comm=GlooComm(rank, world))
            # MinMaxLoc, device-result form, joined by the communicator (ties: constant array -> location 0 everywhere)
            n = 50021
            g = W.c1_uniform(n)
            b, e = spc.shard(n)
            res = torch.zeros(4, dtype=torch.float64)
            spc.minmaxloc_async(HostView(g[b:e]), b, res)
            w = whole.reduce_loc("minmaxloc", g, 0, 1)
            out["mml_async"] = ((float(res[0]), float(res[1]), int(res[2:].view(torch.int64)[0]), int(res[2:].view(torch.int64)[1])),
                                (w.min_val, w.max_val, w.min_loc, w.max_loc))
            spc.minmaxloc_async(HostView(np.full(e - b, -1.5)), b, res)
            out["mml_async_ties"] = ((float(res[0]), float(res[1]), int(res[2:].view(torch.int64)[0]), int(res[2:].view(torch.int64)[1])), (-1.5, -1.5, 0, 0))
            # k-slab stencil, device-result form: locations re-based on the "device", joined by the communicator
            n0, n1, n2 = 18, 11, 23
            u, _, _ = W.c4_field(n0, n1, n2)
            u3 = u.reshape((n0, n1, n2), order="F")
            kb_, ke_ = [(1 + (k * (n2 - 2)) // world) for k in (rank, rank + 1)]
            slab = np.asfortranarray(u3[:, :, kb_ - 1:ke_ + 1]).reshape(-1, order="F").copy()
            spc.stencil7_minmaxloc_async(HostView(slab), n0, n1, ke_ - kb_ + 2, n2, kb_ - 1, 0.5, 0.125, res)
            w, _ = whole.stencil7(u, n0, n1, n2, 0.5, 0.125)
            out["stencil_async"] = ((float(res[0]), float(res[1]), int(res[2:].view(torch.int64)[0]), int(res[2:].view(torch.int64)[1])),
                                    (w.min_val, w.max_val, w.min_loc, w.max_loc))
            # fused distributed scan over block-cyclic Views (ragged and empty global sizes)
            for n in (0, 5, 96 * world, 96 * world * 3 + 41):
                xg = W.c3_wrap(n)
                blk, nl, _ = spc.comm.cyclic_layout(n, np.int64)
                xl = cyclic_take(xg, blk, world, rank)
                assert xl.size == nl
                y = HostView(np.zeros(nl, dtype=np.int64))
                tot = torch.zeros(1, dtype=torch.int64)
                spc.cyclic_scan_async(HostView(xl), y, n, tot)
                wy, wt = whole.scan(xg, False, 0, 1) if n else (xg, 0)
                out[f"cyclic{n}"] = ((int(tot[0]), y.array.tobytes()), (wt, cyclic_take(wy, blk, world, rank).tobytes()))
            # STREAM shards: no exchange
            n = 4099
            bb, cc = W.c1_general(n), W.c1_uniform(n)
            b, e = spc.shard(n)
            a = HostView(np.zeros(e - b))
            spc.stream_triad(a, HostView(bb[b:e]), HostView(cc[b:e]), 3.0)
            out["triad"] = (a.array.tobytes(), (bb + 3.0 * cc)[b:e].tobytes())
            spc.stream_copy(HostView(bb[b:e]), a)
            out["copy"] = (a.array.tobytes(), bb[b:e].tobytes())
            # GUPS on a table sharded by index range: updates are generated per owner, shard-relative
            L, m = 1 << 10, 1 << 13
            table = HostView(np.zeros(L, dtype=np.int64))
            idx_local = W.c5_indices(m, L, seed=100 + rank)
            spc.gups(table, HostView(idx_local), 3)
            gt = np.zeros(L * world, dtype=np.int64)
            for rk in range(world):  # the global update stream: rank rk's updates target its shard [rk*L, (rk+1)*L)
                np.add.at(gt, W.c5_indices(m, L, seed=100 + rk) + rk * L, 3)
            out["gups"] = (table.array.tobytes(), gt[rank * L:(rank + 1) * L].tobytes())
            # SpMV sharded by rows, x replicated
            R = 64 * world
            rm, ci, va, x = W.c5_crs(R, 8, integer_valued=True)
            b, e = rank * 64, (rank + 1) * 64
            yl = HostView(np.zeros(64))
            spc.spmv_rows(HostView(rm[b:e + 1] - rm[b]), HostView(ci[rm[b]:rm[e]]), HostView(va[rm[b]:rm[e]]), HostView(x), yl)
            out["spmv"] = (yl.array.tobytes(), whole.spmv(rm, ci, va, x)[b:e].tobytes())
        q.put((rank, out))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _run(case, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in results) == list(range(world))
    for rank, out in results:
        assert out, "worker produced nothing"
        for key, (got, exp) in out.items():
            assert got == exp, (case, world, rank, key, got, exp)


def test_shard_bounds_cover_and_align():
    from kokkos_b200.sharded import shard_bounds
    for n in (0, 1, 7, 100003, 1 << 30):
        for world in (1, 2, 3, 8):
            for align in (1, 4):
                cuts = [shard_bounds(n, world, r, align) for r in range(world)]
                assert cuts[0][0] == 0 and cuts[-1][1] == n
                for (b0, e0), (b1, e1) in zip(cuts, cuts[1:]):
                    assert e0 == b1 and b0 <= e0
                assert all(b % align == 0 for b, _ in cuts)


def test_join_rules_match_reference_tie_semantics():
    from kokkos_b200.sharded import INDEX_IDENTITY, join_maxloc, join_minloc
    assert join_minloc((1.0, 5), (1.0, 9)) == (1.0, 5)               # equal values keep the first-combined location
    assert join_minloc((1.0, INDEX_IDENTITY), (1.0, 9)) == (1.0, 9)  # ... unless it is the identity
    assert join_minloc((2.0, 5), (1.0, 9)) == (1.0, 9)
    assert join_maxloc((1.0, 5), (1.0, 9)) == (1.0, 5)
    assert join_maxloc((1.0, 5), (3.0, 9)) == (3.0, 9)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_reductions_world(world):
    _run("reduce", world)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_world(world):
    _run("scan", world)


def test_sharded_empty_shards_world2():
    _run("empty_shards", 2)


def test_sharded_stencil_slabs_world2():
    _run("stencil", 2)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_async_forms_cyclic_scan_and_parallel_for_shards(world):
    _run("async_and_for", world)
