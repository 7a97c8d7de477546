#!/usr/bin/env python
"""The search phase of the reference's driver (src/main.rs:107-186, SKIP_PROCESS=true) on GPUs.

    python examples/search_main.py [VSTORE_DIR] [--check]                     # one rank
    torchrun --nproc-per-node N examples/search_main.py [VSTORE_DIR] [--check]  # one rank per GPU

VSTORE_DIR holds global.parquet and rank_0.parquet as the reference writes them
(src/mpi_helpers/vectorstore.rs:5-20); without an argument a synthetic JabRef-sized store
(2 305 x 768, README.md:159 of the reference) is written to a temporary directory first.
Prints what the reference prints: the top-k list, the accuracy metrics and the timing report.
With --check root also runs the CPU restatement of the reference (oracle/, test infrastructure) over
the same Parquet file and prints the GPU-vs-CPU agreement in the same layout: MRR and recall@k of the
GPU results, and their top-k overlap with the CPU results (src/metrics.rs:21-141)."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import better_search_rag_rust_b200 as bsr  # noqa: E402
from better_search_rag_rust_b200.benchmark import BenchmarkManager, BenchmarkTimer  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    manager = BenchmarkManager(rank)
    overall = BenchmarkTimer.start("total_execution")                      # main.rs:36

    check = "--check" in sys.argv
    argv = [a for a in sys.argv[1:] if a != "--check"]
    if argv:
        vstore_dir = argv[0]
    else:
        vstore_dir = os.path.join(tempfile.gettempdir(), "bsr_example_vstore")
        if bsr.is_root(rank):
            rows = np.random.default_rng(0).uniform(-1, 1, (2305, 768)).astype(np.float32)
            for name, part in (("global.parquet", rows), ("rank_0.parquet", rows[:400])):
                st = bsr.GpuVectorstore(os.path.join(vstore_dir, name), True, device=local)
                st.append_many(part)
                st.persist()
                st.close()
        if size > 1:
            dist.barrier()

    top_k, query_idx = 50, 0                                               # main.rs:110-111
    search_timer = BenchmarkTimer.start("similarity_search")               # main.rs:114
    target = torch.zeros(768, device="cuda")
    if bsr.is_root(rank):
        print(f"[Rank {rank}] Generating target vector")
        local_vstore = bsr.get_local_vstore(vstore_dir, rank, False, device=local)
        target.copy_(torch.from_numpy(local_vstore.get(query_idx)))        # main.rs:120
    if size > 1:
        dist.broadcast(target, src=bsr.ROOT)                               # main.rs:123-125
    top = bsr.parallel_top_k_similarity_search(None, rank, size, vstore_dir, top_k,
                                               target.cpu().numpy(), device=local)
    manager.record(search_timer.stop())                                    # main.rs:134

    metrics_timer = BenchmarkTimer.start("metrics_calculation")            # main.rs:138
    if bsr.is_root(rank):
        if top is not None:
            bsr.print_top_k_results(top)                                   # main.rs:145
            mrr, recall, overlap = bsr.calculate_accuracy_metrics(top, query_idx, top_k)
            print(bsr.format_accuracy_metrics(mrr, recall, overlap, top_k))  # main.rs:151-154
            if check:
                from better_search_rag_rust_b200.vectorstore import _read_parquet_rows
                from oracle import c_oracle as co
                rows = _read_parquet_rows(os.path.join(vstore_dir, "global.parquet"))
                q = target.cpu().numpy()
                ci, cd = co.parallel_top_k_similarity_search(rows, size, top_k, q, n_threads=os.cpu_count() or 1)
                same = [p[0] for p in top] == ci.tolist() and \
                    np.asarray([p[1] for p in top], np.float32).view(np.uint32).tolist() == cd.view(np.uint32).tolist()
                print(f"GPU vs CPU reference path ({size} rank(s)): "
                      f"{'identical lists (indices and distance bits)' if same else 'LISTS DIFFER'}")
                print(bsr.agreement_report([q], [[rows[int(i)] for i in ci]], [[rows[p[0]] for p in top]], top_k))
        else:
            print("Error: Failed to compute global top-k results")
    manager.record(metrics_timer.stop())
    manager.record(overall.stop())
    report = manager.generate_report(None)                                 # main.rs:174
    if bsr.is_root(rank):
        print(f"[Rank {rank}] Benchmark report:")
        print("----------------------------------------")
        print(report)
        print("----------------------------------------")
    if size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
