"""Row-block partitioning -- mirrors src/mpi_helpers/load_balance.rs:8-49 of the reference.

The same ceil-div block distribution the reference uses to give each MPI rank its rows now
decides which rows each GPU shard keeps resident in HBM.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Generic, List, Sequence, TypeVar

T = TypeVar("T")
_U64 = 1 << 64


@dataclass(frozen=True)
class RankInterval:
    """load_balance.rs:8-17.  `start_index` is NOT clamped to the row count (load_balance.rs:31)."""

    start_index: int
    end_index: int

    def get_count(self) -> int:
        """load_balance.rs:14-16 as a release build computes it: wrapping usize subtraction."""
        return (self.end_index - self.start_index) % _U64

    def rows(self, count: int) -> tuple[int, int]:
        """(first_row, n_rows) this interval really selects from a store of `count` rows.

        The reference passes (start_index, get_count()) to a polars slice
        (src/mpi_helpers/metrics.rs:30-33), which clamps: a start at or past the end selects
        nothing even when get_count() wrapped.
        """
        if self.start_index >= count:
            return count, 0
        return self.start_index, min(self.get_count(), count - self.start_index)


def interval_by_rank(rank: int, size: int, count: int) -> RankInterval:
    """load_balance.rs:24-42."""
    if size <= 0:
        raise ValueError("size must be positive")
    per_rank = 1 if size > count else -(-count // size)
    start_index = per_rank * rank
    end_index = count if rank == size - 1 else min(start_index + per_rank, count)
    return RankInterval(start_index, end_index)


@dataclass
class RankSlice(Generic[T]):
    """load_balance.rs:19-22."""

    interval: RankInterval
    slice: List[T]


def slice_by_rank(rank: int, size: int, arr: Sequence[T]) -> RankSlice[T]:
    """load_balance.rs:44-49.  Rust's `arr[start..end]` panics when start > end or end > len;
    so does this (IndexError)."""
    interval = interval_by_rank(rank, size, len(arr))
    if interval.start_index > interval.end_index or interval.end_index > len(arr):
        raise IndexError(
            f"slice index starts at {interval.start_index} but ends at {interval.end_index}")
    return RankSlice(interval, list(arr[interval.start_index:interval.end_index]))
