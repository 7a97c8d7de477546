"""BASELINE.json configs[4]: k sweep 1-1024 and query-batch sweep 1-4096 on N x B200 (one process per
GPU, launched with torch.distributed.run).  The store is 100 M x 768 rows at N = 8 (12.5 M rows per
GPU; --rows-per-gpu to change), resident once as fp32 and once as bf16.  Batches of 32 queries and
more run on the tcgen05 filter + exact rescoring (fp32 stores through their bf16 shadow), smaller ones
on the exact scan; every result is the exact scan's, bit for bit.

Every measurement is one `bsr_search_parallel` call per rank (C ABI, host query buffers in, the
root's host top-k out: scan -> candidate exchange (NVLink mailbox for one query with k <= 128, else
ncclAllGather + merge launch) -> final list on root), timed on the host around the collective call
after a barrier, max over ranks.  Rank 0 prints one JSON object and writes it to
gpurun_out/r02_sweeps_n<N>.json."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import better_search_rag_rust_b200 as bsr  # noqa: E402

SEED = 0x5EED0001
DIM = 768


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--max-batch", type=int, default=4096)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n = args.rows_per_gpu

    ids = [bsr.BsrComm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = bsr.BsrComm(ids[0], world, rank, local)

    def timed(fn, reps):
        fn()  # warm-up (buffers, NCCL channels)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            res = fn()
        s = time.perf_counter() - t0
        t = torch.tensor([s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / reps, res

    out = {"n_gpus": world, "rows_per_gpu": n, "rows_total": n * world, "dim": DIM,
           "path": "bsr_search_parallel (C ABI, host buffers; scan -> candidate exchange -> final list on root)",
           "mailbox": comm.uses_mailbox()}
    rng = np.random.default_rng(1)

    def queries(nq, first):
        q = rng.uniform(-1, 1, size=(nq, DIM)).astype(np.float32)
        q[0] = first
        t = torch.from_numpy(q).to(dev)
        dist.broadcast(t, src=0)  # every rank must pass the same queries
        return t.cpu().numpy()

    for dtype in ("f32", "bf16"):
        store = bsr.GpuVectorstore(None, True, device=local, row_base=rank * n, dim=DIM, capacity_rows=n,
                                   dtype=dtype)
        store.append_synthetic(SEED, rank * n, n)
        eng = bsr.CudaShardEngine(store)
        row0 = torch.zeros(DIM, device=dev)
        if rank == 0:
            row0.copy_(torch.from_numpy(store.get(0)))
        dist.broadcast(row0, src=0)
        row0 = row0.cpu().numpy()
        bytes_total = store.scan_bytes() * world

        if dtype == "f32":
            sweep = []
            for k in (1, 10, 50, 100, 256, 1024):
                q = queries(1, row0)
                s, res = timed(lambda: comm.search(eng, q, k), 20)
                if rank == 0:
                    idx, dst, cnt = res
                    assert int(idx[0, 0]) == 0 and float(dst[0, 0]) == 0.0 and int(cnt[0]) == k
                sweep.append({"k": k, "ms_per_query": s * 1e3, "queries_per_s": 1.0 / s,
                              "aggregate_GBps": bytes_total / s / 1e9})
            out["k_sweep_fp32_single_query"] = sweep

        sweep = []
        sizes = (1, 2, 8, 64, 256, 1024, 4096)
        for nq in sizes:
            if nq > args.max_batch:
                continue
            q = queries(nq, row0)
            reps = max(1, min(20, 64 // nq))
            s, res = timed(lambda: comm.search(eng, q, 10), reps)
            if rank == 0:
                idx, dst, cnt = res
                assert int(idx[0, 0]) == 0 and float(dst[0, 0]) == 0.0
            e = {"queries": nq, "ms_per_batch": s * 1e3, "queries_per_s": nq / s}
            if nq >= 32:
                e["TFLOPs_aggregate"] = 2.0 * n * world * DIM * nq / s / 1e12
            else:
                e["aggregate_GBps_per_query_pass"] = bytes_total * nq / s / 1e9
            sweep.append(e)
        out[f"batch_sweep_{dtype}_k10_exact_scan_below_32_tcgen05_filter_above"] = sweep
        if dtype == "bf16":
            sweep = []
            for k in (100, 1024):
                q = queries(1024, row0)
                s, res = timed(lambda: comm.search(eng, q, k), 1)
                sweep.append({"k": k, "queries": 1024, "ms_per_batch": s * 1e3, "queries_per_s": 1024 / s})
            out["k_sweep_bf16_1024_queries"] = sweep
        store.close()

    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/r02_sweeps_n{world}.json", "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
