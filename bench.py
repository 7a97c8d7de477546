#!/usr/bin/env python
"""bench.py -- headline benchmark of the exact top-k search path (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 1 M x 768 fp32 rows resident in each GPU's HBM, one query,
exact top-10 cosine distance.  One step = one query.  At N > 1 every rank holds its own 1 M-row
shard (weak scaling: the store is N x 1 M rows, row-sharded as
src/mpi_helpers/load_balance.rs:24-42 would), a step is scan -> NCCL all-gather of the k candidate
slots -> on-device merge on root.  `value` counts 1 M-row shard scans per second (= queries/s x N,
the whole-job aggregate the contract asks for; at N = 1 it is plain queries/s); the queries/s
against the whole N x 1 M-row store is reported beside it as `queries_per_s_whole_store`.

`--impl reference` times the reference's own CPU path (the C restatement in oracle/, since the
Rust reference cannot be built here) on the host cores; see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

DIM = 768
TOP_K = 10
SEED = 0x5EED0001
ROWS_PER_GPU = 1_000_000
METRIC = "queries/sec (exact top-10 cosine, 768-d fp32)"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--top-k", type=int, default=TOP_K)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(n_gpus: int, rows_per_gpu: int, top_k: int) -> dict:
    return {
        "workload": f"{rows_per_gpu} x {DIM} fp32 rows per GPU, HBM-resident, single query, "
                    f"exact top-{top_k} cosine (BASELINE.json configs[1] shard)",
        "rows_per_gpu": rows_per_gpu,
        "rows_total": rows_per_gpu * n_gpus,
        "dim": DIM,
        "top_k": top_k,
        "queries_per_step": 1,
        "sharding": f"row blocks over {n_gpus} GPU(s), one process per GPU" +
                    (", NCCL all-gather of k candidate slots + on-device merge" if n_gpus > 1 else ""),
        "value_counts": "shard scans per second = queries/s x n_gpus (one query scans every shard)",
        "l2": "inputs larger than L2 (3.07 GB streamed per step per GPU vs 126 MB L2); no flush needed",
        "query": "row 0 of the store (the reference's self-query, src/main.rs:111,120)",
    }


# ---------------------------------------------------------------------------------------------
# clocks (NVML), sampled while the GPU is under load
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "app_clocks_setting",
               0x10: "sync_boost", 0x100: "display_clock_setting"}

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples = []  # (t, sm_mhz, reasons_mask)
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:
                h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = device_index
                if vis:
                    try:
                        idx = int(vis.split(",")[device_index])
                    except Exception:
                        idx = device_index
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.h = h
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop_evt.set()

    def summary(self, t0: float, t1: float) -> dict:
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        scope = "timed region"
        if len(inside) < 3:
            inside = list(self.samples)
            scope = "whole loaded run (timed region too short for 3 samples)"
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = sorted(name for bit, name in self.REASONS.items() if mask & bit)
        mhz = float(np.median([s[1] for s in inside])) if inside else None
        return {"sm_mhz": mhz, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "scope": scope}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, ranks = OS threads
# ---------------------------------------------------------------------------------------------
def _host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def _generate_host_rows(n_rows: int, threads: int) -> np.ndarray:
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle as co
    rows = np.empty((n_rows, DIM), dtype=np.float32)
    chunk = max(1, (n_rows + threads - 1) // threads)

    def fill(i):
        lo, hi = i * chunk, min(n_rows, (i + 1) * chunk)
        if lo < hi:
            co.lib().bsr_oracle_generate_rows(SEED, lo, hi - lo, DIM, rows[lo:hi])

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(fill, range(threads)))
    return rows


def cpu_rate(rows: np.ndarray, threads: int, top_k: int, n_queries: int) -> float:
    """seconds per query: reference path (scan + full stable sort + gather + merge), in memory."""
    from oracle import c_oracle as co
    q = rows[0].copy()
    t0 = time.perf_counter()
    for _ in range(n_queries):
        co.parallel_top_k_similarity_search(rows, threads, top_k, q, n_threads=threads)
    return (time.perf_counter() - t0) / n_queries


def cpu_baseline(rows_per_query: int, top_k: int, budget_s: float = 12.0) -> dict:
    threads = _host_threads()
    sample_rows = min(rows_per_query, 1_000_000)
    rows = _generate_host_rows(sample_rows, threads)
    t1 = cpu_rate(rows, threads, top_k, 1)
    n_q = int(max(2, min(50, budget_s / max(t1, 1e-6))))
    sec = cpu_rate(rows, threads, top_k, n_q)
    value = (1.0 / sec) * (sample_rows / rows_per_query)
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_q} queries over {sample_rows} x {DIM} fp32 rows held in host memory, "
                      f"{threads} ranks as threads (oracle/bsr_oracle.c: scan + full stable sort + "
                      "rank-ordered gather + merge; Parquet decode excluded)",
            "ms_per_query": sec * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from oracle import c_oracle as co
    co.build()
    threads = _host_threads()
    steps, warmup = args.steps, args.warmup
    # calibrate on a small block, then size the per-step sample so the run ends in ~2 minutes
    probe = _generate_host_rows(65_536, threads)
    sec_probe = cpu_rate(probe, threads, args.top_k, 2)
    rows_per_s = 65_536 / sec_probe
    budget = 100.0
    sample_rows = int(min(args.rows_per_gpu, max(4096, rows_per_s * budget / max(1, steps + warmup))))
    rows = probe if sample_rows <= 65_536 else _generate_host_rows(sample_rows, threads)
    rows = rows[:sample_rows]
    q = rows[0].copy()
    for _ in range(warmup):
        co.parallel_top_k_similarity_search(rows, threads, args.top_k, q, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        co.parallel_top_k_similarity_search(rows, threads, args.top_k, q, n_threads=threads)
    dt = time.perf_counter() - t0
    # one step scanned sample_rows rows; a whole 1 M-row shard scan takes rows_per_gpu/sample_rows steps
    value = (steps / dt) * (sample_rows / args.rows_per_gpu)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus, args.rows_per_gpu, args.top_k),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"each step = one query over {sample_rows} of the {args.rows_per_gpu} rows "
                      f"(value scaled by {sample_rows}/{args.rows_per_gpu}), {threads} ranks as "
                      "threads, oracle/bsr_oracle.c restatement of the reference (Rust toolchain "
                      "absent); rows in host memory, Parquet decode excluded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def _peak_gbs() -> tuple[float, str]:
    p = os.path.join(HERE, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy benchmark)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def _traffic_per_launch(rows_per_gpu: int):
    """dram bytes per scan launch from the committed ncu capture, if it matches this workload."""
    p = os.path.join(HERE, "profiles", "r01_scan_summary.json")
    try:
        with open(p) as f:
            j = json.load(f)
        if int(j.get("rows")) == rows_per_gpu:
            return float(j["dram_bytes_read"]) + float(j["dram_bytes_write"])
    except Exception:
        pass
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    import better_search_rag_rust_b200 as bsr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the search path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, k, steps, warmup = args.rows_per_gpu, args.top_k, args.steps, max(3, args.warmup)

    store = bsr.GpuVectorstore(None, True, device=local, row_base=rank * n, dim=DIM, capacity_rows=n)
    store.append_synthetic(SEED, rank * n, n)  # global store = generator rows [0, world*n)
    eng = bsr.CudaShardEngine(store)
    slots = bsr.candidate_slots(k)

    # the query: row 0 of the global store (lives on rank 0), broadcast like src/main.rs:123-125
    q_host = torch.zeros((1, DIM), dtype=torch.float32).pin_memory()
    if rank == 0:
        q_host.copy_(torch.from_numpy(store.get(0)).reshape(1, -1))
    q_dev = q_host.to(dev)
    if world > 1:
        dist.broadcast(q_dev, src=0)
        q_host.copy_(q_dev.cpu())
    q_np = q_host.numpy()

    cand = torch.empty((1, slots), dtype=torch.int64, device=dev)
    gathered = torch.empty((world, 1, slots), dtype=torch.int64, device=dev)
    merged = torch.empty((1, slots), dtype=torch.int64, device=dev)

    def step_device():
        eng.search_candidates(q_dev, k, out=cand)
        if world > 1:
            dist.all_gather_into_tensor(gathered, cand)
            if rank == 0:
                eng.merge(gathered, k, out=merged)

    comm = None
    if world > 1:
        # the library's own NCCL communicator: the id travels over the host channel (here torch's)
        ids = [bsr.BsrComm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = bsr.BsrComm(ids[0], world, rank, local)

    def step_e2e():
        if world == 1:
            return eng.local_top_k(q_np, k)  # bsr_search: host buffers in, host buffers out
        return comm.search(eng, q_np, k)     # bsr_search_parallel: one C-ABI call per rank

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident region: `value` and the roofline ------------------------------------
    for _ in range(warmup):
        step_device()
    fence()
    store.set_profiling(True)
    launches0 = store.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    fence()
    t_wall1 = time.perf_counter()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    launches = sum_over_ranks(store.launch_count() - launches0)
    scan_ms, n_scans = store.scan_profile()
    store.set_profiling(False)

    # ---- end-to-end region: host query in, host top-k out, every step ----------------------------
    for _ in range(warmup):
        step_e2e()
    fence()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):
        res = step_e2e()
    g1.record()
    fence()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.stop()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return None

    # sanity: the self-query must come back first with distance exactly 0 (src/main.rs:148-154)
    idx, dst, cnt = res
    assert int(idx[0, 0]) == 0 and float(dst[0, 0]) == 0.0 and int(cnt[0]) == k, (idx, dst, cnt)

    peak, peak_src = _peak_gbs()
    scan_bytes = store.scan_bytes()
    avg_scan_ms = scan_ms / max(1, n_scans)
    achieved = scan_bytes / (avg_scan_ms * 1e-3) / 1e9
    q_per_s = steps / (ms_dev * 1e-3)
    line = {
        "metric": METRIC, "value": q_per_s * world, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_dev / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, n, k),
        "queries_per_s_whole_store": q_per_s,
        "wall_ms_per_step": (t_wall1 - t_wall0) / steps * 1e3,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": _traffic_per_launch(n),
            "kernel": "scan_topk_f32 (fused cosine scan + top-k), rank 0",
            "algorithmic_bytes_per_launch": scan_bytes,
            "avg_launch_ms": avg_scan_ms, "launches_timed": n_scans,
            "kernel_share_of_step": avg_scan_ms / (ms_dev / steps),
            "peak_source": peak_src,
            "frac_of_nominal_8TBs": achieved / 8000.0,
        },
        "e2e": {"value": steps / e2e_s * world, "unit": UNIT,
                "h2d_bytes_per_step": DIM * 4 * world,
                "d2h_bytes_per_step": k * 12 + 4,
                "ms_per_step": e2e_s / steps * 1e3,
                "path": "bsr_search (C ABI, host buffers)" if world == 1 else
                        "bsr_search_parallel (C ABI, host buffers: scan -> ncclAllGather -> merge on root)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(t_wall0, t_wall1),
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import c_oracle as co
            co.build()
            line["cpu_baseline"] = cpu_baseline(n, k)
        except Exception as e:  # the GPU numbers stand even if the host leg cannot run
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": _host_threads(),
                                    "kind": "port", "sample": f"failed: {e!r}"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    args = parse_args()
    # The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun notices)
    # also write to fd 1, so while the benchmark runs fd 1 points at stderr; the line is written to
    # the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
