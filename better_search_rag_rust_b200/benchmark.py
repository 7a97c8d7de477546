"""Timing report -- mirrors src/mpi_helpers/benchmark.rs of the reference (BenchmarkTiming,
BenchmarkTimer, BenchmarkManager::generate_report, :10-438) so a run of the GPU search phase prints
the same "PARALLEL PERFORMANCE REPORT" block as the reference binary, operation names included
(`similarity_search`, `metrics_calculation`, ... src/main.rs:36,53,61,86,114,138).  Host-side only."""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Dict, List, Optional


@dataclass
class BenchmarkTiming:
    """benchmark.rs:10-41."""

    name: str
    duration: float  # seconds
    items_processed: Optional[int] = None

    def seconds(self) -> float:
        return self.duration

    def throughput(self) -> Optional[float]:
        return None if self.items_processed is None else self.items_processed / self.duration


class BenchmarkTimer:
    """benchmark.rs:44-77."""

    def __init__(self, name: str, items: Optional[int] = None):
        self.name, self.items, self._t0 = name, items, time.perf_counter()

    @classmethod
    def start(cls, name: str) -> "BenchmarkTimer":
        return cls(name)

    @classmethod
    def start_with_items(cls, name: str, items: int) -> "BenchmarkTimer":
        return cls(name, items)

    def stop(self) -> BenchmarkTiming:
        return BenchmarkTiming(self.name, time.perf_counter() - self._t0, self.items)


class BenchmarkManager:
    """benchmark.rs:80-439.  `world` is a torch.distributed group (None = default group, or no
    process group at all for a single rank)."""

    def __init__(self, rank: int):
        self.rank = rank
        self.timings: Dict[str, BenchmarkTiming] = {}

    def record(self, timing: BenchmarkTiming) -> None:
        self.timings[timing.name] = timing

    def get(self, name: str) -> Optional[BenchmarkTiming]:
        return self.timings.get(name)

    def all_timings(self) -> List[BenchmarkTiming]:
        return list(self.timings.values())

    def gather_timings(self, world=None) -> Dict[int, List[BenchmarkTiming]]:
        """benchmark.rs:131-292: every rank's timings at root, keyed by rank."""
        try:
            import torch.distributed as dist
            multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(world) > 1
        except Exception:
            multi = False
        if not multi:
            return {self.rank: self.all_timings()} if self.rank == 0 else {}
        size = dist.get_world_size(world)
        gathered: List[Optional[list]] = [None] * size
        dist.all_gather_object(gathered, [(t.name, t.duration, t.items_processed) for t in self.all_timings()],
                               group=world)
        if self.rank != 0:
            return {}
        return {r: [BenchmarkTiming(*t) for t in (gathered[r] or [])] for r in range(size)}

    def generate_report(self, world=None, sequential_times: Optional[Dict[str, float]] = None) -> str:
        """benchmark.rs:296-438: text on root, empty string elsewhere."""
        all_timings = self.gather_timings(world)
        if self.rank != 0:
            return ""
        report = "==== PARALLEL PERFORMANCE REPORT ====\n\n"
        names = sorted({t.name for ts in all_timings.values() for t in ts})
        for op_name in names:
            report += f"Operation: {op_name}\n"
            op = [(r, t) for r, ts in all_timings.items() for t in ts if t.name == op_name]
            if not op:
                report += "  No timing data available\n\n"
                continue
            secs = [t.seconds() for _, t in op]
            min_time, max_time, avg_time = min(secs), max(secs), sum(secs) / len(secs)
            min_rank = min(op, key=lambda x: x[1].seconds())[0]
            max_rank = max(op, key=lambda x: x[1].seconds())[0]
            report += f"  Min time: {min_time:.4f} sec (Rank {min_rank})\n"
            report += f"  Max time: {max_time:.4f} sec (Rank {max_rank})\n"
            report += f"  Avg time: {avg_time:.4f} sec\n"
            if sequential_times and op_name in sequential_times:
                report += f"  Speedup: {sequential_times[op_name] / min_time:.2f}x\n"
        return report


def time_operation(name: str, fn, *args, **kwargs):
    """benchmark.rs:441-466: run fn, return (result, BenchmarkTiming)."""
    timer = BenchmarkTimer.start(name)
    result = fn(*args, **kwargs)
    return result, timer.stop()
