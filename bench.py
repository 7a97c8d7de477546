#!/usr/bin/env python
"""bench.py -- headline benchmark of the exact top-k search path (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the north-star target, one shard of it per GPU): 12.5 M x 768 fp32
rows resident in each GPU's HBM (38.4 GB), one query, exact top-10 cosine distance.  At N = 8 the
store is the full 100 M x 768 rows, row-sharded as src/mpi_helpers/load_balance.rs:24-42 would.  One
step = one query against the whole N-shard store: scan of every shard -> exchange of the k candidate
slots -> on-device merge on root (weak scaling: rows per GPU fixed).

`value` counts 12.5 M-row shard scans per second (= queries/s x N: the whole-job aggregate the
contract asks for, proportional to rows scanned per second; at N = 1 it is plain queries/s).  The
queries/s against the whole N-shard store is reported beside it as `queries_per_s_whole_store`.
`extra_1m` repeats the measurement on BASELINE.json configs[1] (1 M x 768 per GPU);
`extra_small_stores` is the latency of one search on 2 305 and 50 000 rows (rank 0's GPU).

After the timed regions every rank checks its shard's COMPLETE top-k against the CPU oracle run
block-wise over the regenerated rows, and root checks the merged list (`parity_checked`).

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
ROWS_PER_GPU = 12_500_000      # 100 M rows / 8 GPUs (BASELINE.json configs[3]); 38.4 GB of fp32 per GPU
ROWS_EXTRA = 1_000_000         # BASELINE.json configs[1]
METRIC = "shard-queries/sec: queries/sec x row shards scanned per query (exact top-10 cosine, 768-d fp32)"
UNIT = "shard-queries/s"


def small_store_latency(bsr, device: int) -> dict:
    """One search on a small resident store, end to end through the C ABI with host buffers (`bsr_search`
    called through ctypes on preallocated arrays: what a Rust / C++ caller pays), query = row 0 of the store
    (the reference's self-query, src/main.rs:111-120): median / 10th percentile of the per-call wall clock
    and the launch's own duration (bsr_store_last_scan_ms).  2 305 rows is the reference's own store
    (README.md:159), k = 50 its own top_k (src/main.rs:110); 50 000 rows is BASELINE.json configs[0]."""
    from better_search_rag_rust_b200 import _lib
    L = _lib.lib()
    cases = []
    for rows, kk in ((2305, 50), (2305, 10), (50_000, 10), (50_000, 50)):
        st = bsr.GpuVectorstore(None, True, device=device, dim=DIM, capacity_rows=rows)
        try:
            st.append_synthetic(SEED, 0, rows)
            q = np.ascontiguousarray(st.get(0).reshape(1, -1), dtype=np.float32)
            oi, od = np.zeros((1, kk), np.uint64), np.zeros((1, kk), np.float32)
            oc = np.zeros(1, np.uint32)
            cargs = (st.handle, q.ctypes.data, 1, kk, oi.ctypes.data, od.ctypes.data, oc.ctypes.data)
            for _ in range(50):
                _lib.check(L.bsr_search(*cargs))
            t = np.empty(1000)
            for i in range(t.size):
                t0 = time.perf_counter_ns()
                L.bsr_search(*cargs)
                t[i] = time.perf_counter_ns() - t0
            ok = bool(oi[0, 0] == 0 and od[0, 0] == 0.0 and oc[0] == kk)
            cases.append({"rows": rows, "dim": DIM, "k": kk, "e2e_us_median": float(np.median(t)) / 1e3,
                          "e2e_us_p10": float(np.percentile(t, 10)) / 1e3,
                          "queries_per_s": 1e9 / float(np.median(t)),
                          "kernel_us": st.last_scan_ms() * 1e3, "self_query_first_at_distance_0": ok})
        finally:
            st.close()
    return {"path": "bsr_search (C ABI, host buffers), per-call wall clock on rank 0, 1 000 calls per case",
            "cases": cases}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--top-k", type=int, default=TOP_K)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run oracle check (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the 1 M-row extra measurement")
    ap.add_argument("--no-sweeps", action="store_true", help="skip the k / batch sweeps on the 1 M-row extra store")
    ap.add_argument("--exchange", choices=["auto", "nccl"], default="auto",
                    help="cross-rank candidate exchange of the e2e leg: library default, or ncclAllGather only")
    return ap.parse_args()


def workload_config(n_gpus: int, rows_per_gpu: int, top_k: int) -> dict:
    which = ("BASELINE.json configs[3] shard: 100 M x 768 over 8 GPUs" if rows_per_gpu == ROWS_PER_GPU
             else "BASELINE.json configs[1]" if rows_per_gpu == ROWS_EXTRA else "custom size")
    return {
        "workload": f"{rows_per_gpu} x {DIM} fp32 rows per GPU, HBM-resident, single query, "
                    f"exact top-{top_k} cosine ({which})",
        "rows_per_gpu": rows_per_gpu,
        "rows_total": rows_per_gpu * n_gpus,
        "dim": DIM,
        "top_k": top_k,
        "queries_per_step": 1,
        "sharding": f"row blocks over {n_gpus} GPU(s), one process per GPU" +
                    (", exchange of k candidate slots + on-device merge on root" if n_gpus > 1 else ""),
        "value_counts": "row shards scanned per second = queries/s x n_gpus (one query scans every shard); "
                        "queries/s against the whole store is queries_per_s_whole_store",
        "l2": f"inputs larger than L2 ({rows_per_gpu * DIM * 4 / 1e9:.2f} GB streamed per step per GPU vs 126 MB L2); no flush needed",
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
    # at most 2 M rows (6 GB of host memory) per step: the rate is flat in the row count
    sample_rows = int(min(args.rows_per_gpu, 2_000_000, max(4096, rows_per_s * budget / max(1, steps + warmup))))
    rows = probe if sample_rows <= 65_536 else _generate_host_rows(sample_rows, threads)
    rows = rows[:sample_rows]
    q = rows[0].copy()
    for _ in range(warmup):
        co.parallel_top_k_similarity_search(rows, threads, args.top_k, q, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        co.parallel_top_k_similarity_search(rows, threads, args.top_k, q, n_threads=threads)
    dt = time.perf_counter() - t0
    # one step scanned sample_rows rows; a whole shard scan takes rows_per_gpu/sample_rows steps
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
    """dram bytes per scan launch.  DRAM counters cannot be read inside a timed run; the figure comes
    from the committed ncu capture of this same command (profiles/r02_scan_summary.json) when its
    row count matches, and the line says so in `traffic_source`."""
    p = os.path.join(HERE, "profiles", "r02_scan_summary.json")
    try:
        with open(p) as f:
            j = json.load(f)
        for cap in j.get("captures", []):
            if int(cap.get("rows")) == rows_per_gpu and cap.get("dtype", "f32") == "f32":
                return (float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"]),
                        f"profiles/r02_scan_summary.json ({cap.get('source', 'ncu --set full')})")
    except Exception:
        pass
    return None, "no committed ncu capture for this row count"


def oracle_shard_top_k(rank: int, n: int, k: int, q: np.ndarray, threads: int):
    """The reference's compute_local_top_k (src/mpi_helpers/metrics.rs:36-50) for generator rows
    [rank*n, (rank+1)*n), run block-wise on `threads` host threads -> (global idx, dist)."""
    from oracle import c_oracle as co
    return co.scan_generated_parallel(SEED, rank * n, n, k, q, threads=threads, dim=DIM)


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
    k, steps, warmup = args.top_k, args.steps, max(3, args.warmup)

    comm = None
    if world > 1:
        # the library's own communicator: the id travels over the host channel (here torch's)
        ids = [bsr.BsrComm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = bsr.BsrComm(ids[0], world, rank, local)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()

    def measure(n: int, n_steps: int, check: bool, sweep: bool = False) -> dict:
        """Both timed regions (device-resident and end to end) on an n-row shard per GPU, then the
        parity check of the complete top-k against the oracle."""
        store = bsr.GpuVectorstore(None, True, device=local, row_base=rank * n, dim=DIM, capacity_rows=n)
        store.append_synthetic(SEED, rank * n, n)  # global store = generator rows [0, world*n)
        if args.exchange == "nccl":
            store.set_option("no_mailbox", 1)
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
            if world == 1:
                eng.search_candidates(q_dev, k, out=cand)      # bsr_search_device
            else:
                comm.search_device(eng, q_dev, k, out=merged)  # bsr_search_parallel_device (asynchronous)

        def step_e2e():
            if world == 1:
                return eng.local_top_k(q_np, k)  # bsr_search: host buffers in, host buffers out
            return comm.search(eng, q_np, k)     # bsr_search_parallel: one C-ABI call per rank

        # ---- device-resident region: `value` and the roofline --------------------------------
        fence()  # every rank's shard is resident before the first collective search (ranks build at their own pace)
        for _ in range(warmup):
            step_device()
        fence()
        store.set_profiling(True)
        launches0 = store.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        e0.record()
        for _ in range(n_steps):
            step_device()
        e1.record()
        fence()
        t_wall1 = time.perf_counter()
        ms_dev = reduce_ranks(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None)
        launches = reduce_ranks(store.launch_count() - launches0, dist.ReduceOp.SUM if world > 1 else None)
        scan_ms, n_scans = store.scan_profile()
        store.set_profiling(False)

        # ---- end-to-end region: host query in, host top-k out, every step ------------------------
        for _ in range(warmup):
            step_e2e()
        fence()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            res = step_e2e()
        fence()
        e2e_s = reduce_ranks(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)

        exchange = "none (one rank)"
        if world > 1:
            exchange = ("NVLink mailbox: peer stores from each scan's last CTA into root's HBM, folded in by root's "
                        "last CTA (one kernel per rank and query)" if comm.uses_mailbox() and args.exchange != "nccl"
                        else "ncclAllGather of the candidate slots + merge launch on root")
        # ---- BASELINE.json configs[4] in small: k sweep and query-batch sweep, end to end ---------------
        sweeps = None
        if sweep:
            sweeps = {"path": "bsr_search / bsr_search_parallel (C ABI, host buffers), host-timed, max over ranks; "
                              "queries/s against the whole store", "k_sweep_single_query": [], "batch_sweep_k10": []}
            rng = np.random.default_rng(11)

            def timed_e2e(fn, reps):
                fn()
                fence()
                t0 = time.perf_counter()
                for _ in range(reps):
                    r = fn()
                s_ = reduce_ranks(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
                return s_ / reps, r

            def search(qs, kk):
                return eng.local_top_k(qs, kk) if world == 1 else comm.search(eng, qs, kk)

            for kk in (1, 50, 100, 1024):
                sec, r = timed_e2e(lambda: search(q_np, kk), 20)
                if rank == 0:
                    assert int(r[0][0, 0]) == 0 and float(r[1][0, 0]) == 0.0 and int(r[2][0]) == kk
                sweeps["k_sweep_single_query"].append({"k": kk, "ms_per_query": sec * 1e3, "queries_per_s": 1.0 / sec})
            for nq in (8, 64, 256, 1024):
                qs = rng.uniform(-1, 1, size=(nq, DIM)).astype(np.float32)
                qs[0] = q_np[0]
                if world > 1:  # every rank must pass the same queries
                    t = torch.from_numpy(qs).to(dev)
                    dist.broadcast(t, src=0)
                    qs = t.cpu().numpy()
                sec, r = timed_e2e(lambda: search(qs, k), 3)
                if rank == 0:
                    assert int(r[0][0, 0]) == 0 and float(r[1][0, 0]) == 0.0
                sweeps["batch_sweep_k10"].append({"queries": nq, "ms_per_batch": sec * 1e3, "queries_per_s": nq / sec,
                                                  "path": "tcgen05 filter + exact rescoring (fp32 store: bf16 shadow)"})

        out = {"exchange": exchange, "sweeps": sweeps, "n": n, "steps": n_steps, "ms_dev": ms_dev, "launches": launches, "scan_ms": scan_ms,
               "n_scans": n_scans, "e2e_s": e2e_s, "scan_bytes": store.scan_bytes(),
               "t_wall": (t_wall0, t_wall1), "parity": "skipped (--no-parity)"}

        # ---- parity: the COMPLETE top-k against the oracle, outside the timed regions -----------
        if check:
            threads = max(1, _host_threads() // world)
            t_chk = time.perf_counter()
            li, ld, lc = eng.local_top_k(q_np, k)                      # this shard, through bsr_search
            wi, wd = oracle_shard_top_k(rank, n, k, q_np[0], threads)  # compute_local_top_k on the same rows
            ok_local = (int(lc[0]) == len(wi) and li[0, :lc[0]].tolist() == wi.tolist()
                        and ld[0, :lc[0]].view(np.uint32).tolist() == wd.view(np.uint32).tolist())
            lists = [(wi, wd)]
            if world > 1:
                lists = [None] * world if rank == 0 else None
                dist.gather_object((wi, wd), lists, dst=0)
            ok_global = True
            if rank == 0:
                from oracle import c_oracle as co
                gi = np.concatenate([p[0] for p in lists])             # rank order (metrics.rs:86-89)
                gd = np.concatenate([p[1] for p in lists])
                fi, fd = co.compute_global_top_k(gi, gd, k)            # metrics.rs:141-171
                idx, dst, cnt = res
                if world > 1:  # the device-resident leg's last result, too
                    mi, md, mc = eng.unpack(merged, k)
                    ok_local = ok_local and (int(mc[0]) == len(fi) and mi[0, :mc[0]].tolist() == fi.tolist()
                                             and md[0, :mc[0]].view(np.uint32).tolist() == fd.view(np.uint32).tolist())
                ok_global = (int(cnt[0]) == len(fi) and idx[0, :cnt[0]].tolist() == fi.tolist()
                             and dst[0, :cnt[0]].view(np.uint32).tolist() == fd.view(np.uint32).tolist()
                             # the self-query comes back first with distance exactly 0 (src/main.rs:148-154)
                             and int(idx[0, 0]) == 0 and float(dst[0, 0]) == 0.0)
            n_bad = reduce_ranks(0.0 if (ok_local and ok_global) else 1.0, dist.ReduceOp.SUM if world > 1 else None)
            if n_bad:
                raise SystemExit(f"PARITY FAILURE on {int(n_bad)} rank(s): GPU top-{k} differs from the oracle "
                                 f"(rank {rank}: local ok={ok_local}, merged ok={ok_global})")
            out["parity"] = "full"
            out["parity_detail"] = (
                f"every rank: its shard's complete top-{k} (bsr_search) == oracle compute_local_top_k over the "
                f"{n} regenerated rows, indices and distance bits; root: merged list of the e2e leg == oracle "
                f"compute_global_top_k of the rank-ordered oracle lists; {threads} host thread(s) per rank, "
                f"{time.perf_counter() - t_chk:.1f} s")
        store.close()
        return out

    main = measure(args.rows_per_gpu, steps, not args.no_parity)
    extra = None
    if not args.no_extra and args.rows_per_gpu != ROWS_EXTRA:
        extra = measure(ROWS_EXTRA, max(steps, 200), not args.no_parity, sweep=not args.no_sweeps)
    sampler.stop()
    if comm is not None:
        comm.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return None

    peak, peak_src = _peak_gbs()

    def summarize(m: dict) -> dict:
        avg_scan_ms = m["scan_ms"] / max(1, m["n_scans"])
        q_per_s = m["steps"] / (m["ms_dev"] * 1e-3)
        achieved = m["scan_bytes"] / (avg_scan_ms * 1e-3) / 1e9
        traffic, traffic_src = _traffic_per_launch(m["n"])
        return {
            "value": q_per_s * world, "queries_per_s_whole_store": q_per_s,
            "ms_per_step": m["ms_dev"] / m["steps"],
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "scan_topk_bulk_kernel<1,false,8,true> (fused cosine scan + top-k, k <= 32 instantiation), rank 0",
                "algorithmic_bytes_per_launch": m["scan_bytes"],
                "avg_launch_ms": avg_scan_ms, "launches_timed": m["n_scans"],
                "kernel_share_of_step": avg_scan_ms / (m["ms_dev"] / m["steps"]),
                "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,
            },
            "e2e": {"value": m["steps"] / m["e2e_s"] * world, "unit": UNIT,
                    "queries_per_s_whole_store": m["steps"] / m["e2e_s"],
                    "h2d_bytes_per_step": DIM * 4 * world,
                    "d2h_bytes_per_step": k * 12 + 4,
                    "ms_per_step": m["e2e_s"] / m["steps"] * 1e3,
                    "path": "bsr_search (C ABI, host buffers)" if world == 1 else
                            "bsr_search_parallel (C ABI, host buffers: scan -> candidate exchange -> merge on root)"},
            "gpu_launches": int(m["launches"]), "exchange": m["exchange"],
            "parity_checked": m["parity"], "parity_detail": m.get("parity_detail"),
        }

    sm = summarize(main)
    line = {
        "metric": METRIC, "value": sm["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": sm["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, args.rows_per_gpu, k),
        "queries_per_s_whole_store": sm["queries_per_s_whole_store"],
        "wall_ms_per_step": (main["t_wall"][1] - main["t_wall"][0]) / steps * 1e3,
        "roofline": sm["roofline"], "e2e": sm["e2e"], "gpu_launches": sm["gpu_launches"],
        "exchange": sm["exchange"],
        "parity_checked": sm["parity_checked"], "parity_detail": sm["parity_detail"],
        "clocks": sampler.summary(*main["t_wall"]),
    }
    if extra is not None:
        se = summarize(extra)
        line["extra_1m"] = {"config": workload_config(world, ROWS_EXTRA, k), "steps": extra["steps"], **se}
        if extra.get("sweeps"):
            line["extra_1m"]["sweeps"] = extra["sweeps"]
    if not args.no_extra:
        # BASELINE.json configs[0] and the reference's own store under the driver's clock: the latency of one
        # search on a small store, rank 0's GPU only (the other ranks wait at the final barrier).  Never
        # allowed to cost the line above: any failure is recorded, not raised.
        try:
            line["extra_small_stores"] = small_store_latency(bsr, local)
        except Exception as e:
            line["extra_small_stores"] = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import c_oracle as co
            co.build()
            line["cpu_baseline"] = cpu_baseline(args.rows_per_gpu, k)
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
