"""Result check / print -- mirrors src/mpi_helpers/metrics.rs:209-249 and the offline IR metrics
of src/metrics.rs:7-141 of the reference (north_star: "metrics.rs, which is reused to check
results").  Host-side bookkeeping over a handful of result vectors; no device work here."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

TopK = List[Tuple[int, float]]
Vectors = Sequence[Sequence[float]]

_F32 = np.float32
_EPS = np.float32(1e-10)  # src/metrics.rs:13 (an f32 literal)


def vectors_are_identical(a, b) -> bool:
    """src/metrics.rs:7-19: same length and no element further apart than 1e-10 (a NaN
    difference does not count as a difference)."""
    a = np.asarray(a, dtype=_F32).ravel()
    b = np.asarray(b, dtype=_F32).ravel()
    if a.size != b.size:
        return False
    with np.errstate(invalid="ignore", over="ignore"):
        return not bool(np.any(np.abs(a - b) > _EPS))


def mean_reciprocal_rank(queries: Vectors, results: Sequence[Vectors], ids_map=None) -> float:
    """src/metrics.rs:21-58.  `ids_map` is accepted for signature parity: the reference only binds
    an unused variable from it (:38-41)."""
    if len(queries) == 0 or len(results) == 0 or len(queries) != len(results):
        return 0.0
    total = _F32(0.0)
    for query, query_results in zip(queries, results):
        rank = 0
        for i, result in enumerate(query_results):
            if vectors_are_identical(query, result):
                rank = i + 1
                break
        if rank == 0:
            rank = len(query_results) + 1
        total = _F32(total + _F32(_F32(1.0) / _F32(rank)))
    return float(_F32(total / _F32(len(queries))))


def recall_at_k(queries: Vectors, results: Sequence[Vectors], k: int, ids_map=None) -> float:
    """src/metrics.rs:60-96."""
    if len(queries) == 0 or len(results) == 0 or len(queries) != len(results):
        return 0.0
    total = _F32(0.0)
    for query, query_results in zip(queries, results):
        actual_k = min(k, len(query_results))
        found = any(vectors_are_identical(query, query_results[i]) for i in range(actual_k))
        total = _F32(total + _F32(1.0 if found else 0.0))
    return float(_F32(total / _F32(len(queries))))


def top_k_overlap(sequential_results: Sequence[Vectors], parallel_results: Sequence[Vectors],
                  k: int) -> float:
    """src/metrics.rs:98-141: per query, the share of the first k sequential results that also
    appear among the first k parallel results; averaged over all queries (a query with no results
    contributes 0 but still counts, :117-119)."""
    if (len(sequential_results) == 0 or len(parallel_results) == 0
            or len(sequential_results) != len(parallel_results)):
        return 0.0
    total = _F32(0.0)
    for seq_results, par_results in zip(sequential_results, parallel_results):
        actual_k = min(k, len(seq_results), len(par_results))
        if actual_k == 0:
            continue
        overlap = 0
        for i in range(actual_k):
            if any(vectors_are_identical(seq_results[i], par_results[j]) for j in range(actual_k)):
                overlap += 1
        total = _F32(total + _F32(_F32(overlap) / _F32(actual_k)))
    return float(_F32(total / _F32(len(sequential_results))))


def result_vectors(store, top_k_results: Sequence[Tuple[int, float]]) -> List[np.ndarray]:
    """The stored rows behind a top-k list: what the offline metrics above compare.  `store` is
    any object with `get(local_index)` and `row_base` (a GpuVectorstore shard)."""
    base = int(getattr(store, "row_base", 0))
    return [np.asarray(store.get(int(idx) - base), dtype=_F32) for idx, _ in top_k_results]


def agreement_report(queries: Vectors, sequential_results: Sequence[Vectors],
                     parallel_results: Sequence[Vectors], top_k: int) -> str:
    """GPU-vs-CPU agreement in the layout of src/main.rs:151-154: MRR and recall@k of the parallel
    (GPU) results against the queries, and their top-k overlap with the sequential (CPU) results."""
    return format_accuracy_metrics(mean_reciprocal_rank(queries, parallel_results),
                                   recall_at_k(queries, parallel_results, top_k),
                                   top_k_overlap(sequential_results, parallel_results, top_k), top_k)


def _fmt_f32(x: float) -> str:
    """Rust's `{}` for an f32: the shortest decimal that round-trips."""
    x = np.float32(x)
    if x == np.floor(x) and abs(float(x)) < 1e16:
        return str(int(x))
    return np.format_float_positional(x, unique=True, trim="-")


def format_top_k_results(top_k_results: TopK) -> str:
    lines = [f"Global top-{len(top_k_results)} results:"]
    for i, (idx, dist) in enumerate(top_k_results):
        lines.append(f"  {i + 1}. Index: {idx}, Distance: {_fmt_f32(dist)}")
    return "\n".join(lines)


def print_top_k_results(top_k_results: TopK) -> None:
    """metrics.rs:209-214."""
    print(format_top_k_results(top_k_results))


def calculate_accuracy_metrics(top_k_results: Sequence[Tuple[int, float]], query_idx: int,
                               top_k: int) -> Tuple[float, float, float]:
    """metrics.rs:217-249 -> (mrr, recall, overlap) for the single self-query."""
    position = 0
    for i, (idx, _) in enumerate(top_k_results):
        if idx == query_idx:
            position = i + 1
            break
    mrr = float(np.float32(1.0) / np.float32(position)) if position > 0 else 0.0
    recall = 1.0 if (position > 0 and position <= top_k) else 0.0
    overlap = 1.0 if position > 0 else 0.0
    return mrr, recall, overlap


def format_accuracy_metrics(mrr: float, recall: float, overlap: float, top_k: int) -> str:
    """The block src/main.rs:151-154 prints."""
    return ("Accuracy Metrics:\n"
            f"  Mean Reciprocal Rank (MRR): {mrr:.4f}\n"
            f"  Recall@{top_k}: {recall:.4f}\n"
            f"  Top-k Overlap: {overlap:.4f}")
