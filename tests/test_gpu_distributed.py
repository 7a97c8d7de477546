"""GPU, >= 2 devices: parallel_top_k_similarity_search over NCCL (one process per GPU) against the
oracle's parallel search.  Skipped on a single-GPU box; the single-GPU suite covers the same
pipeline with shards run back to back (tests/test_gpu_parity.py::_sharded_search)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, n_rows, k, q_out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=size, device_id=torch.device("cuda", rank))
    try:
        import better_search_rag_rust_b200 as bsr
        from better_search_rag_rust_b200 import search as S
        seed = 0x5EED0001
        iv = bsr.interval_by_rank(rank, size, n_rows)
        first, cnt = iv.rows(n_rows)
        store = bsr.GpuVectorstore(None, True, device=rank, dim=768, row_base=min(iv.start_index, 2**32 - 2))
        if cnt:
            store.append_synthetic(seed, first, cnt)
        eng = bsr.CudaShardEngine(store)
        q = torch.zeros(768, device="cuda")
        if rank == 0:
            q.copy_(torch.from_numpy(store.get(0)))
        dist.broadcast(q, src=0)                       # src/main.rs:123-125
        queries = (q.cpu().numpy(), np.linspace(-1, 1, 768, dtype=np.float32))
        out = {}
        out["torch"] = [bsr.parallel_top_k_similarity_search(None, rank, size, None, k, query, engine=eng)
                        for query in queries]
        # the same search through the library's own communicator (bsr_comm_*, bsr_search_parallel):
        # NVLink mailbox where it applies (top_k <= 32), then forced through ncclAllGather
        ids = [bsr.BsrComm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = bsr.BsrComm(ids[0], size, rank, rank)
        out["uses_mailbox"] = comm.uses_mailbox()
        out["comm"] = [bsr.parallel_top_k_similarity_search(comm, rank, size, None, k, query, engine=eng)
                       for query in queries]
        store.set_option("no_mailbox", 1)
        out["comm_nccl"] = [bsr.parallel_top_k_similarity_search(comm, rank, size, None, k, query, engine=eng)
                            for query in queries]
        store.set_option("no_mailbox", 0)
        # device-resident, asynchronous: twelve searches in flight back to back (peers run ahead of root
        # through the mailbox's four slots and then wait for its acknowledgements), alternating queries
        qd = [torch.from_numpy(x).cuda().reshape(1, -1) for x in queries]
        outs = [comm.search_device(eng, qd[i % 2], k) for i in range(12)]
        torch.cuda.synchronize()
        dev_lists = []
        if rank == 0:
            for o in outs:
                gi, gd = S.unpack_candidates(o.cpu().numpy().view(np.uint64)[0], k)
                dev_lists.append(list(zip(gi.tolist(), gd.tolist())))
        out["device"] = dev_lists
        comm.close()
        q_out.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,n_rows", [(10, 100_003), (32, 20_011), (50, 2_305), (100, 100_003), (200, 30_000), (10, 5), (50, 5)])
def test_nccl_parallel_search(k, n_rows):
    """Every GPU of the box as one rank: torch's NCCL group, the library's communicator with the NVLink
    mailbox (k <= 128) and with ncclAllGather, host buffers and device-resident -- all against the
    oracle's parallel_top_k_similarity_search.  n_rows = 5 leaves ranks with an empty shard."""
    import torch
    import torch.multiprocessing as mp
    from oracle import c_oracle as co
    size = torch.cuda.device_count()          # every GPU of the box (8 on the scaling box)
    if size < 2:
        pytest.skip("needs at least 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, n_rows, k, q)) for r in range(size)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(size))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rows = co.generate_rows(0x5EED0001, 0, n_rows, 768)
    queries = (rows[0], np.linspace(-1, 1, 768, dtype=np.float32))
    assert res[0]["uses_mailbox"], "separate processes on separate NVLink-connected GPUs: the mailbox must be up"

    def same(got, wi, wd):
        assert [p[0] for p in got] == wi.tolist()
        assert np.asarray([p[1] for p in got], np.float32).view(np.uint32).tolist() == wd.view(np.uint32).tolist()

    for qi, query in enumerate(queries):
        wi, wd = co.parallel_top_k_similarity_search(rows, size, k, query, n_threads=8)
        for path in ("torch", "comm", "comm_nccl"):
            same(res[0][path][qi], wi, wd)
            for r in range(1, size):
                assert res[r][path][qi] is None
        for i in range(qi, 12, 2):
            same(res[0]["device"][i], wi, wd)
