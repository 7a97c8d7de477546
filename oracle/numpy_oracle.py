"""Independent numpy restatement of the reference search path (TEST INFRASTRUCTURE ONLY).

Written separately from bsr_oracle.c so the two can be checked against each other
(tests/test_oracle.py).  Sequential f32 sums use ``np.add.accumulate`` on float32,
which is a strict left-to-right fold (unlike ``np.sum``, which is pairwise); the
pure-Python ``*_loop`` variants use np.float32 scalars and are for tiny cases.

Citations are relative to /root/reference/.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
_EPS = F32(1e-10)  # src/metrics.rs:13, an f32 literal


def _seq_sum_f32(x: np.ndarray) -> np.float32:
    """Iterator::sum::<f32>() -- left-to-right fold starting from 0.0."""
    x = np.asarray(x, dtype=F32)
    if x.size == 0:
        return F32(0.0)
    return np.add.accumulate(x, dtype=F32)[-1]


def vectors_are_identical(a, b) -> bool:
    """src/metrics.rs:7-19"""
    a = np.asarray(a, dtype=F32)
    b = np.asarray(b, dtype=F32)
    if a.size != b.size:
        return False
    with np.errstate(invalid="ignore"):
        return not bool(np.any(np.abs(a - b) > _EPS))  # NaN > eps is False


def cosine_distance(a, b) -> np.float32:
    """src/metrics.rs:143-165"""
    a = np.asarray(a, dtype=F32).ravel()
    b = np.asarray(b, dtype=F32).ravel()
    if a.size != b.size or a.size == 0:
        return F32(1.0)
    if vectors_are_identical(a, b):
        return F32(0.0)
    with np.errstate(all="ignore"):
        dot = _seq_sum_f32(a * b)
        mag_a = np.sqrt(_seq_sum_f32(a * a), dtype=F32)
        mag_b = np.sqrt(_seq_sum_f32(b * b), dtype=F32)
        if mag_a == 0.0 or mag_b == 0.0:
            return F32(1.0)
        sim = F32(dot / F32(mag_a * mag_b))
        # f32::max / f32::min ignore a NaN operand -> np.fmax / np.fmin
        sim = np.fmin(np.fmax(sim, F32(-1.0)), F32(1.0))
        return F32(F32(1.0) - sim)


def cosine_distance_loop(a, b) -> np.float32:
    """Same as cosine_distance, element-by-element np.float32 scalar arithmetic."""
    a = [F32(x) for x in np.asarray(a, dtype=F32).ravel()]
    b = [F32(x) for x in np.asarray(b, dtype=F32).ravel()]
    if len(a) != len(b) or not a:
        return F32(1.0)
    with np.errstate(all="ignore"):
        if all(not (abs(F32(x - y)) > _EPS) for x, y in zip(a, b)):
            return F32(0.0)
        dot = F32(0.0)
        for x, y in zip(a, b):
            dot = F32(dot + F32(x * y))
        sa = F32(0.0)
        for x in a:
            sa = F32(sa + F32(x * x))
        sb = F32(0.0)
        for y in b:
            sb = F32(sb + F32(y * y))
        ma, mb = F32(np.sqrt(sa)), F32(np.sqrt(sb))
        if ma == 0.0 or mb == 0.0:
            return F32(1.0)
        sim = F32(dot / F32(ma * mb))
        sim = np.fmin(np.fmax(sim, F32(-1.0)), F32(1.0))
        return F32(F32(1.0) - sim)


def interval_by_rank(rank: int, size: int, count: int) -> tuple[int, int]:
    """src/mpi_helpers/load_balance.rs:24-42 (start is NOT clamped to count)."""
    per_rank = 1 if size > count else -(-count // size)
    start = per_rank * rank
    end = count if rank == size - 1 else min(start + per_rank, count)
    return start, end


def effective_block(rank: int, size: int, count: int) -> tuple[int, int]:
    """Rows get_many(Some(SliceArgs)) really returns: polars clamps the slice
    (src/vectorstore/polars.rs:121-129), so start >= count gives an empty block even when
    get_count() (load_balance.rs:14-16) wrapped."""
    start, end = interval_by_rank(rank, size, count)
    if start >= count:
        return count, 0
    length = (end - start) % (1 << 64)
    return start, min(length, count - start)


def _distances(rows: np.ndarray, target: np.ndarray) -> np.ndarray:
    """cosine_distance of every row against target, vectorised over rows but with the
    reference's sequential column order (accumulate along axis 1)."""
    rows = np.asarray(rows, dtype=F32)
    target = np.asarray(target, dtype=F32).ravel()
    n, d = rows.shape
    if d != target.size or d == 0:
        return np.full(n, 1.0, dtype=F32)
    with np.errstate(all="ignore"):
        ident = ~np.any(np.abs(rows - target[None, :]) > _EPS, axis=1)
        dot = np.add.accumulate(rows * target[None, :], axis=1, dtype=F32)[:, -1]
        mag_a = np.sqrt(np.add.accumulate(rows * rows, axis=1, dtype=F32)[:, -1], dtype=F32)
        mag_b = np.sqrt(_seq_sum_f32(target * target), dtype=F32)
        sim = (dot / (mag_a * mag_b).astype(F32)).astype(F32)
        sim = np.fmin(np.fmax(sim, F32(-1.0)), F32(1.0))
        dist = (F32(1.0) - sim).astype(F32)
        dist[(mag_a == 0.0) | (mag_b == 0.0)] = F32(1.0)
        dist[ident] = F32(0.0)
    return dist


def dot_distance(a, b) -> np.float32:
    """BSR_DOT (include/bsr.h): 1 - the fold of src/metrics.rs:153 alone."""
    a = np.asarray(a, dtype=F32).ravel()
    b = np.asarray(b, dtype=F32).ravel()
    if a.size != b.size or a.size == 0:
        return F32(1.0)
    with np.errstate(all="ignore"):
        return F32(F32(1.0) - _seq_sum_f32(a * b))


def _dot_distances(rows: np.ndarray, target: np.ndarray) -> np.ndarray:
    rows = np.asarray(rows, dtype=F32)
    target = np.asarray(target, dtype=F32).ravel()
    n, d = rows.shape
    if d != target.size or d == 0:
        return np.full(n, 1.0, dtype=F32)
    with np.errstate(all="ignore"):
        dot = np.add.accumulate(rows * target[None, :], axis=1, dtype=F32)[:, -1]
        return (F32(1.0) - dot).astype(F32)


def _ordered_bits(d: np.ndarray) -> np.ndarray:
    """total order of BSR_DOT distances: numbers ascending, -0 == +0, every NaN last"""
    d = (np.asarray(d, dtype=F32) + F32(0.0)).astype(F32)
    b = d.view(np.uint32).copy()
    b[np.isnan(d)] = np.uint32(0x7FC00000)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def compute_local_top_k(rows, rank: int, size: int, top_k: int, target, metric: int = 0):
    """src/mpi_helpers/metrics.rs:16-53 (metric 1: BSR_DOT distances in their total order)"""
    rows = np.asarray(rows, dtype=F32)
    first, n = effective_block(rank, size, rows.shape[0])
    if n == 0:
        return np.zeros(0, np.uint64), np.zeros(0, F32)
    if metric:
        dist = _dot_distances(rows[first:first + n], target)
        order = np.argsort(_ordered_bits(dist), kind="stable")[:top_k]
    else:
        dist = _distances(rows[first:first + n], target)
        order = np.argsort(dist, kind="stable")[:top_k]  # stable sort + truncate (:47-50)
    return (order + first).astype(np.uint64), dist[order]


def mean_reciprocal_rank(queries, results) -> np.float32:
    """src/metrics.rs:21-58 (ids_map only feeds an unused binding there)."""
    if len(queries) == 0 or len(results) == 0 or len(queries) != len(results):
        return F32(0.0)
    total = F32(0.0)
    for q, res in zip(queries, results):
        rank = 0
        for i, r in enumerate(res):
            if vectors_are_identical(q, r):
                rank = i + 1
                break
        if rank == 0:
            rank = len(res) + 1
        total = F32(total + F32(F32(1.0) / F32(rank)))
    return F32(total / F32(len(queries)))


def recall_at_k(queries, results, k: int) -> np.float32:
    """src/metrics.rs:60-96"""
    if len(queries) == 0 or len(results) == 0 or len(queries) != len(results):
        return F32(0.0)
    total = F32(0.0)
    for q, res in zip(queries, results):
        found = any(vectors_are_identical(q, res[i]) for i in range(min(k, len(res))))
        total = F32(total + F32(1.0 if found else 0.0))
    return F32(total / F32(len(queries)))


def top_k_overlap(sequential_results, parallel_results, k: int) -> np.float32:
    """src/metrics.rs:98-141"""
    if (len(sequential_results) == 0 or len(parallel_results) == 0
            or len(sequential_results) != len(parallel_results)):
        return F32(0.0)
    total = F32(0.0)
    for seq, par in zip(sequential_results, parallel_results):
        actual_k = min(k, len(seq), len(par))
        if actual_k == 0:
            continue
        overlap = 0
        for i in range(actual_k):
            if any(vectors_are_identical(seq[i], par[j]) for j in range(actual_k)):
                overlap += 1
        total = F32(total + F32(F32(overlap) / F32(actual_k)))
    return F32(total / F32(len(sequential_results)))


def compute_global_top_k(global_idx, global_dist, top_k: int):
    """src/mpi_helpers/metrics.rs:141-171"""
    gi = np.asarray(global_idx, dtype=np.uint64).ravel()
    gd = np.asarray(global_dist, dtype=F32).ravel()
    order = np.argsort(gd, kind="stable")
    out_i, out_d, seen = [], [], set()
    for j in order:
        i = int(gi[j])
        if i in seen:
            continue
        seen.add(i)
        out_i.append(i)
        out_d.append(gd[j])
        if len(out_i) >= top_k:
            break
    return np.asarray(out_i, dtype=np.uint64), np.asarray(out_d, dtype=F32)


def parallel_top_k_similarity_search(rows, size: int, top_k: int, target):
    """src/mpi_helpers/metrics.rs:174-206; ranks run one after another, gathered in rank order."""
    parts = [compute_local_top_k(rows, r, size, top_k, target) for r in range(size)]
    gi = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.uint64)
    gd = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, F32)
    return compute_global_top_k(gi, gd, top_k)


def calculate_accuracy_metrics(idx, query_idx: int, top_k: int):
    """src/mpi_helpers/metrics.rs:217-249"""
    position = 0
    for i, v in enumerate(np.asarray(idx).ravel()):
        if int(v) == query_idx:
            position = i + 1
            break
    mrr = float(F32(1.0) / F32(position)) if position > 0 else 0.0
    recall = 1.0 if 0 < position <= top_k else 0.0
    overlap = 1.0 if position > 0 else 0.0
    return mrr, recall, overlap


# ---------------------------------------------------------------------------
# synthetic rows (same counter-based generator as bsr_oracle.c / the CUDA kernel)
# ---------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return z ^ (z >> np.uint64(31))


def generate_rows(seed: int, first_row: int, n_rows: int, dim: int = 768) -> np.ndarray:
    with np.errstate(over="ignore"):
        rows = (np.arange(n_rows, dtype=np.uint64) + np.uint64(first_row))
        rk = _splitmix64(np.uint64(seed) ^ (rows * np.uint64(0x9E3779B97F4A7C15)))
        h = _splitmix64(rk[:, None] + np.arange(dim, dtype=np.uint64)[None, :])
    u = (h >> np.uint64(40)).astype(np.uint32)
    return (u.astype(F32) * F32(2.0 ** -23) - F32(1.0)).astype(F32)
