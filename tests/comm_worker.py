"""One rank of the multi-process communicator tests (tests/test_gpu_comm.py starts `world` of these, one per GPU).

usage: comm_worker.py RANK WORLD UNIQUE_ID [big]
Checks, against numpy on the same seeded inputs (bit-exact): all-gather, rank-ordered all-reduces, loc joins with ties,
and the fused block-cyclic scan at ragged / empty / multi-round sizes.  Prints "ok rank R" on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import kokkos_b200 as kb  # noqa: E402
from kokkos_b200.sharded import cyclic_take  # noqa: E402
import workloads as W  # noqa: E402


def main():
    rank, world, uid = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    big = len(sys.argv) > 4 and sys.argv[4] == "big"
    space = kb.B200(rank)
    comm = kb.Comm(space, rank, world, uid)
    for kv in os.environ.get("KB200_TUNE", "").split(","):  # e.g. KB200_TUNE=comm.agd=-1 (aggregates ahead of the data)
        if "=" in kv:
            kb.tune_set(kv.split("=")[0], int(kv.split("=")[1]))
    if os.environ.get("KB200_COMM_ALGO"):  # 1 = the lock-step kernel (ScanChunked.hpp) instead of the rounds kernel
        kb.tune_set("comm.scan_algo", int(os.environ["KB200_COMM_ALGO"]))

    # ---- all-gather: 24 bytes per rank
    src = space.view_from_host(np.array([rank * 10 + 1, rank * 10 + 2, -rank], dtype=np.int64))
    dst = space.view(3 * world, np.int64)
    for _ in range(6):  # more calls than ring slots
        comm.allgather(src.ptr, dst.ptr, 24)
    exp = np.concatenate([np.array([q * 10 + 1, q * 10 + 2, -q], dtype=np.int64) for q in range(world)])
    assert np.array_equal(dst.to_host(), exp), (rank, dst.to_host())

    # ---- all-reduce, folded in rank order (double sum is NOT associative: the order is part of the contract)
    vals = lambda q: np.array([0.1 * (q + 1), 1e16 if q == 0 else 1.0, -3.5 + q], dtype=np.float64)  # noqa: E731
    buf = space.view_from_host(vals(rank))
    comm.allreduce("sum", buf.ptr, 3, np.float64)
    acc = vals(0).copy()
    for q in range(1, world):
        acc = acc + vals(q)
    assert np.array_equal(buf.to_host(), acc), (rank, buf.to_host(), acc)
    for op, fn in (("min", np.minimum), ("max", np.maximum)):
        buf = space.view_from_host(vals(rank))
        comm.allreduce(op, buf.ptr, 3, np.float64)
        acc = vals(0).copy()
        for q in range(1, world):
            acc = fn(acc, vals(q))
        assert np.array_equal(buf.to_host(), acc)
    ib = space.view_from_host(np.array([(1 << 62) + rank, -rank], dtype=np.int64))
    comm.allreduce("sum", ib.ptr, 2, np.int64)
    with np.errstate(over="ignore"):
        e0 = np.int64(0)
        for q in range(world):
            e0 = np.int64(e0 + np.int64((1 << 62) + q))  # wraps mod 2^64 for world >= 2
    assert ib.to_host()[0] == e0 and ib.to_host()[1] == -sum(range(world))

    # ---- MinMaxLoc join: every rank holds the same extrema (a tie) -> the lowest location wins on every rank
    mm = np.zeros(1, dtype=[("a", "<f8"), ("b", "<f8"), ("c", "<i8"), ("d", "<i8")])
    mm[0] = (-2.0, 7.0, 1000 - rank, 500 + rank)
    mv = space.view_from_host(mm.view(np.int64))
    comm.allreduce_loc("minmaxloc", mv.ptr)
    got = mv.to_host().view(mm.dtype)[0]
    assert (got["a"], got["b"], got["c"], got["d"]) == (-2.0, 7.0, 1000 - (world - 1), 500), got

    # ---- fused block-cyclic scan
    block, _, _ = comm.cyclic_layout(1, np.int64)
    sizes = [0, 1, block - 1, block, block + 1, world * block, world * block + 17, 3 * world * block - 5, 2 * world * block + block // 2]
    if big:
        sizes.append((1 << 27) * world + 12345)
    for n in sizes:
        for gen in (W.c3_small, W.c3_wrap):
            if n > (1 << 26) and gen is W.c3_wrap:
                continue
            xg = gen(n)
            with np.errstate(over="ignore"):
                incl = np.cumsum(xg, dtype=np.int64)
            excl = incl - xg
            b, n_local, nsteps = comm.cyclic_layout(n, np.int64)
            xl = cyclic_take(xg, b, world, rank)
            assert xl.size == n_local, (n, xl.size, n_local)
            vx = space.view_from_host(xl) if n_local else space.view(2, np.int64)
            vy = space.view(max(n_local, 2), np.int64)
            vx.n = vy.n = n_local
            for inclusive, expg in ((False, excl), (True, incl)):
                total = comm.parallel_scan(vx, vy, n, inclusive=inclusive)
                exp_total = int(incl[-1]) if n else 0
                assert total == exp_total, ("total", n, inclusive, total, exp_total)
                if n_local:
                    vy.n = n_local
                    got = vy.to_host()
                    ok = np.array_equal(got, cyclic_take(expg, b, world, rank))
                    assert ok, ("scan", rank, n, inclusive, int(np.argmax(got != cyclic_take(expg, b, world, rank))))
            del vx, vy
    assert comm.error() == 0
    comm.barrier()
    space.fence()
    comm.finalize()
    space.finalize()
    print(f"ok rank {rank}", flush=True)


if __name__ == "__main__":
    main()
