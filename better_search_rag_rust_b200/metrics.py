"""Result check / print -- mirrors src/mpi_helpers/metrics.rs:209-249 of the reference
(north_star: "metrics.rs, which is reused to check results")."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

TopK = List[Tuple[int, float]]


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
