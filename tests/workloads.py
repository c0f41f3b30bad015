"""Seeded synthetic inputs of SURVEY.md section 8(d) / BASELINE.md section 4, shared by tests and bench.

Everything is a pure function of (n, seed) built with vectorised numpy so the same arrays can be
regenerated on any machine; nothing here touches the oracle or the product."""
from __future__ import annotations

import numpy as np

GUPS_SEED = 20230913  # benchmarks/gups/gups.cpp:103


def hash_u32(i: np.ndarray) -> np.ndarray:
    """(i * 2654435761) >> 7 on 64-bit integers, the mixer SURVEY 8(d) C1 names."""
    return ((i.astype(np.uint64) * np.uint64(2654435761)) >> np.uint64(7))


def hash64(i: np.ndarray, seed: int = 0) -> np.ndarray:
    """splitmix64 finaliser: full-range 64-bit values (exercises wrap-around in integer sums/scans)."""
    z = i.astype(np.uint64) + np.uint64((0x9E3779B97F4A7C15 * (seed + 1)) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def c1_exact(n: int) -> np.ndarray:
    """C1 (i): integer-valued doubles hash(i) % 100 -- any summation order gives the same bits."""
    return (hash_u32(np.arange(n, dtype=np.uint64)) % np.uint64(100)).astype(np.float64)


def c1_general(n: int) -> np.ndarray:
    """C1 (ii): 1/((i % 1000)+1)."""
    return 1.0 / ((np.arange(n, dtype=np.int64) % 1000) + 1).astype(np.float64)


def c1_uniform(n: int, seed: int = GUPS_SEED) -> np.ndarray:
    """C1 (iii): uniform(-1,1) from a counter-based generator."""
    return (hash64(np.arange(n, dtype=np.uint64), seed) >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def c3_small(n: int) -> np.ndarray:
    """C3: hash(i) % 7 - 3 as int64."""
    return (hash_u32(np.arange(n, dtype=np.uint64)) % np.uint64(7)).astype(np.int64) - 3


def c3_wrap(n: int, seed: int = 1) -> np.ndarray:
    """C3 large-magnitude variant: full 64-bit range, sums wrap mod 2^64 (still bit-exact)."""
    return hash64(np.arange(n, dtype=np.uint64), seed).view(np.int64)


def c4_field(n0: int, n1: int, n2: int, seed: int = 7):
    """C4: smooth field on a LayoutLeft n0 x n1 x n2 grid (i fastest) with one planted unique max and
    min at seeded interior positions.  Returns (flat array, (imax,jmax,kmax), (imin,jmin,kmin))."""
    i = np.arange(n0, dtype=np.float64)[:, None, None]
    j = np.arange(n1, dtype=np.float64)[None, :, None]
    k = np.arange(n2, dtype=np.float64)[None, None, :]
    u = np.sin(0.011 * i + 0.3) * np.cos(0.017 * j) + 0.5 * np.sin(0.013 * k + 0.1 * np.sin(0.02 * i))
    rng = np.random.default_rng(seed)
    pmax = tuple(int(rng.integers(2, d - 2)) for d in (n0, n1, n2))
    pmin = tuple(int(rng.integers(2, d - 2)) for d in (n0, n1, n2))
    while max(abs(a - b) for a, b in zip(pmax, pmin)) < 3:
        pmin = tuple(int(rng.integers(2, d - 2)) for d in (n0, n1, n2))
    u[pmax] = 64.0
    u[pmin] = -64.0
    return np.asfortranarray(u).reshape(-1, order="F").copy(), pmax, pmin


def c5_indices(m: int, table_len: int, seed: int = GUPS_SEED) -> np.ndarray:
    """C5a: uniform indices in [0, table_len)."""
    return (hash64(np.arange(m, dtype=np.uint64), seed) % np.uint64(table_len)).astype(np.int64)


def c5_crs(nrows: int, nnz_per_row: int = 32, ncols: int | None = None, seed: int = 11, integer_valued: bool = False):
    """C5b: CRS matrix, `nnz_per_row` entries per row: half banded around the diagonal, half random."""
    ncols = ncols or nrows
    r = np.arange(nrows, dtype=np.int64)[:, None]
    q = np.arange(nnz_per_row, dtype=np.int64)[None, :]
    band = (r + q - nnz_per_row // 4) % ncols
    rnd = (hash64((r * nnz_per_row + q).astype(np.uint64), seed) % np.uint64(ncols)).astype(np.int64)
    col = np.where(q < nnz_per_row // 2, band, rnd).astype(np.int32).reshape(-1)
    row_map = np.arange(nrows + 1, dtype=np.int64) * nnz_per_row
    h = hash64(np.arange(nrows * nnz_per_row, dtype=np.uint64), seed + 1)
    if integer_valued:
        values = ((h % np.uint64(17)).astype(np.float64) - 8.0)
        x = ((hash64(np.arange(ncols, dtype=np.uint64), seed + 2) % np.uint64(13)).astype(np.float64) - 6.0)
    else:
        values = (h >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0
        x = (hash64(np.arange(ncols, dtype=np.uint64), seed + 2) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    return row_map, col, values, x


def exact_sum(x: np.ndarray) -> float:
    """Correctly rounded sum (math.fsum) -- the 'exact' the 1e-12 tolerance is stated against."""
    import math
    return math.fsum(x.tolist()) if x.size <= (1 << 22) else float(np.sum(x.astype(np.longdouble)))


def checksum64(a: np.ndarray) -> str:
    """Position-dependent 64-bit checksum of an array's bytes: sum_i word_i * (2i+1) mod 2^64 (hex)."""
    w = np.frombuffer(np.ascontiguousarray(a).tobytes(), dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = np.arange(w.size, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        return hex(int(np.sum(w * k, dtype=np.uint64)))
