"""CPU, world_size 2 and 3 over gloo: the multi-rank orchestration of
parallel_top_k_similarity_search / gather_top_k_results (src/mpi_helpers/metrics.rs:56-206).
The per-rank device work is replaced by a test double backed by the oracle (tests/_oracle_engine.py);
the real CUDA engine is exercised by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, n_rows, dim, k, q_out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import better_search_rag_rust_b200 as bsr
        from _oracle_engine import OracleShardEngine
        from oracle import c_oracle as co

        rows = co.generate_rows(0x5EED0001, 0, n_rows, dim)
        query = rows[0] if n_rows else np.zeros(dim, np.float32)
        eng = OracleShardEngine(rows, rank, size)
        got = bsr.parallel_top_k_similarity_search(None, rank, size, None, k, query, engine=eng)
        local = bsr.compute_local_top_k(None, rank, size, k, query, engine=eng)
        g_idx, g_dist = bsr.gather_top_k_results(None, rank, local)
        q_out.put((rank, got, local, g_idx, g_dist))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,n_rows,k", [(2, 1000, 10), (2, 1001, 50), (3, 5, 4), (3, 2, 10), (2, 0, 5)])
def test_parallel_search_over_gloo(size, n_rows, k):
    from oracle import c_oracle as co
    dim = 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, n_rows, dim, k, q)) for r in range(size)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(size):
        rank, got, local, g_idx, g_dist = q.get(timeout=120)
        results[rank] = (got, local, g_idx, g_dist)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    rows = co.generate_rows(0x5EED0001, 0, n_rows, dim)
    query = rows[0] if n_rows else np.zeros(dim, np.float32)
    want_i, want_d = co.parallel_top_k_similarity_search(rows, size, k, query)
    # Some(top_k) on root, None elsewhere (metrics.rs:200-205)
    got = results[0][0]
    assert [p[0] for p in got] == want_i.tolist()
    assert np.asarray([p[1] for p in got], np.float32).view(np.uint32).tolist() == want_d.view(np.uint32).tolist()
    for r in range(1, size):
        assert results[r][0] is None
        assert results[r][2] == [] and results[r][3] == []          # gather: empty off root
    # gather_top_k_results: rank-ordered concatenation of the local lists (metrics.rs:86-126)
    cat_i = [p[0] for r in range(size) for p in results[r][1]]
    cat_d = [p[1] for r in range(size) for p in results[r][1]]
    assert results[0][2] == cat_i
    assert np.asarray(results[0][3], np.float32).view(np.uint32).tolist() == \
        np.asarray(cat_d, np.float32).view(np.uint32).tolist()
    for r in range(size):
        li, ld = co.compute_local_top_k(rows, r, size, k, query)
        assert [p[0] for p in results[r][1]] == li.tolist()
