"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the committed
fixtures, bit for bit.

The bar for this path is exact equality of the (index, distance-bits) lists: the scan keeps the
reference's sequential binary32 arithmetic (src/metrics.rs:153-164) and the candidate keys
reproduce the stable-sort order (src/mpi_helpers/metrics.rs:47,153), so no tolerance is needed.
"""
import os

import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from better_search_rag_rust_b200 import search as S
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

SEED = 0x5EED0001
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
F = np.float32


def _bits(x):
    return np.asarray(x, dtype=F).view(np.uint32)


def _assert_pairs(got, want_idx, want_bits):
    assert [p[0] for p in got] == np.asarray(want_idx).tolist()
    assert _bits([p[1] for p in got]).tolist() == np.asarray(want_bits).tolist()


def _store_from(rows, **kw):
    s = bsr.GpuVectorstore(None, True, **kw)
    s.append_many(rows)
    return s


def _sharded_search(rows, size, k, q):
    """The multi-rank pipeline on ONE GPU: `size` shards as separate stores (row blocks of
    interval_by_rank), their candidate lists concatenated in rank order exactly as the all-gather
    lays them out, merged on device.  (Ranks that wait on each other must not be emulated as
    concurrent processes on one GPU; this runs them back to back.)"""
    import torch
    n = rows.shape[0]
    cands, engines = [], []
    slots = bsr.candidate_slots(k)
    for r in range(size):
        iv = bsr.interval_by_rank(r, size, n)
        first, cnt = iv.rows(n)
        st = bsr.GpuVectorstore(None, True, dim=rows.shape[1], row_base=min(iv.start_index, 2**32 - 2))
        if cnt:
            st.append_many(rows[first:first + cnt])
        eng = bsr.CudaShardEngine(st)
        engines.append(eng)
        cands.append(eng.search_candidates(q.reshape(1, -1), k))
    gathered = torch.stack(cands)                      # [size, 1, slots]
    assert gathered.shape == (size, 1, slots)
    merged = engines[0].merge(gathered, k)
    idx, dist, count = engines[0].unpack(merged, k)
    for e in engines:
        e.store.close()
    return [(int(idx[0, j]), float(dist[0, j])) for j in range(int(count[0]))]


# ---- against the oracle on seeded inputs ---------------------------------------------------------
@pytest.mark.parametrize("n,k", [(1000, 10), (4097, 50), (33, 10), (31, 50), (1, 1), (50000, 10),
                                 (50000, 50), (20000, 1024), (3000, 100), (100, 1024)])
def test_local_top_k_matches_oracle(n, k):
    rows = co.generate_rows(SEED, 0, n, 768)
    store = _store_from(rows)
    eng = bsr.CudaShardEngine(store)
    for q in (rows[0], co.generate_rows(SEED + 0x1000, 0, 1, 768)[0]):
        got = bsr.compute_local_top_k(None, 0, 1, k, q, engine=eng)
        wi, wd = co.compute_local_top_k(rows, 0, 1, k, q)
        _assert_pairs(got, wi, _bits(wd))
    store.close()


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 31, 77, 100, 770, 1024])
def test_other_dimensions(dim):
    rows = co.generate_rows(SEED, 0, 777, dim)
    store = _store_from(rows)
    eng = bsr.CudaShardEngine(store)
    for q in (rows[5], co.generate_rows(9, 0, 1, dim)[0]):
        got = bsr.compute_local_top_k(None, 0, 1, 20, q, engine=eng)
        wi, wd = co.compute_local_top_k(rows, 0, 1, 20, q)
        _assert_pairs(got, wi, _bits(wd))
    store.close()


@pytest.mark.parametrize("variant", ["bulk", "ldg"])
def test_both_scan_kernels(variant, monkeypatch):
    """The bulk-async (TMA ring) kernel is the default; the register-staged one must agree."""
    import subprocess, sys, textwrap
    code = textwrap.dedent(f"""
        import numpy as np, better_search_rag_rust_b200 as bsr
        from oracle import c_oracle as co
        rows = co.generate_rows({SEED}, 0, 40000, 768)
        st = bsr.GpuVectorstore(None, True); st.append_many(rows)
        eng = bsr.CudaShardEngine(st)
        for k in (10, 200):
            got = bsr.compute_local_top_k(None, 0, 1, k, rows[0], engine=eng)
            wi, wd = co.compute_local_top_k(rows, 0, 1, k, rows[0])
            assert [p[0] for p in got] == wi.tolist()
            assert np.asarray([p[1] for p in got], np.float32).view(np.uint32).tolist() == wd.view(np.uint32).tolist()
        print("ok")
    """)
    env = dict(os.environ, BSR_SCAN_VARIANT=variant)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


# ---- against the committed fixtures --------------------------------------------------------------
@pytest.mark.parametrize("name", ["synthetic_2k_768", "synthetic_ragged_1027_100",
                                  "synthetic_odd_dim_300_77"])
def test_synthetic_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n, dim = int(g["n"]), int(g["dim"])
    store = bsr.GpuVectorstore(None, True, dim=dim)
    store.append_synthetic(int(g["seed"]), 0, n)          # generated on device
    eng = bsr.CudaShardEngine(store)
    rows = store.get_many_array(None)
    assert np.array_equal(rows.view(np.uint32), co.generate_rows(int(g["seed"]), 0, n, dim).view(np.uint32))
    for qi, q in enumerate(g["queries"]):
        for k in g["ks"]:
            got = bsr.compute_local_top_k(None, 0, 1, int(k), q, engine=eng)
            _assert_pairs(got, g[f"idx_q{qi}_k{k}_s1"], g[f"dist_bits_q{qi}_k{k}_s1"])
            for size in g["sizes"]:
                if size == 1:
                    continue
                got = _sharded_search(rows, int(size), int(k), q)
                _assert_pairs(got, g[f"idx_q{qi}_k{k}_s{size}"], g[f"dist_bits_q{qi}_k{k}_s{size}"])
    store.close()


def test_edge_case_golden():
    """zero rows, duplicates (ties), NaN / inf, denormal products, identity threshold, k > rows,
    more ranks than rows."""
    g = np.load(os.path.join(GOLDEN, "edge_cases.npz"))
    rows = g["rows"]
    store = _store_from(rows)
    eng = bsr.CudaShardEngine(store)
    for qi, q in enumerate(g["queries"]):
        # every distance: ask for all rows and compare per index
        got = bsr.compute_local_top_k(None, 0, 1, 64, q, engine=eng)
        assert len(got) == len(rows)
        by_idx = {i: d for i, d in got}
        assert _bits([by_idx[i] for i in range(len(rows))]).tolist() == g[f"all_dist_bits_q{qi}"].tolist()
        for k in g["ks"]:
            got = bsr.compute_local_top_k(None, 0, 1, int(k), q, engine=eng)
            _assert_pairs(got, g[f"idx_q{qi}_k{k}_s1"], g[f"dist_bits_q{qi}_k{k}_s1"])
            for size in g["sizes"]:
                if size == 1:
                    continue
                got = _sharded_search(rows, int(size), int(k), q)
                _assert_pairs(got, g[f"idx_q{qi}_k{k}_s{size}"], g[f"dist_bits_q{qi}_k{k}_s{size}"])
    store.close()


# ---- the reference's own functions, multi-rank -----------------------------------------------------
@pytest.mark.parametrize("size", [2, 3, 8])
def test_sharded_matches_parallel_oracle(size):
    rows = co.generate_rows(SEED, 0, 10007, 768)
    for k in (10, 50):
        for q in (rows[0], co.generate_rows(SEED + 0x1001, 0, 1, 768)[0]):
            got = _sharded_search(rows, size, k, q)
            wi, wd = co.parallel_top_k_similarity_search(rows, size, k, q, n_threads=4)
            _assert_pairs(got, wi, _bits(wd))


def test_compute_global_top_k_on_device():
    rows = co.generate_rows(SEED, 0, 5000, 768)
    q = rows[3]
    parts = [co.compute_local_top_k(rows, r, 4, 50, q) for r in range(4)]
    gi = np.concatenate([p[0] for p in parts])
    gd = np.concatenate([p[1] for p in parts])
    got = bsr.compute_global_top_k(gi.tolist(), gd.tolist(), 50, device=0)
    wi, wd = co.compute_global_top_k(gi, gd, 50)
    _assert_pairs(got, wi, _bits(wd))
    assert bsr.compute_global_top_k([], [], 5, device=0) == []


def test_self_query_invariant_and_metrics():
    """src/main.rs:110-154: k = 50, query = row 0 -> index 0 first, distance 0.0, metrics 1/1/1."""
    store = bsr.GpuVectorstore(None, True, dim=768)
    store.append_synthetic(SEED, 0, 20000)
    eng = bsr.CudaShardEngine(store)
    target = store.get(0)
    top = bsr.parallel_top_k_similarity_search(None, 0, 1, None, 50, target, engine=eng)
    assert top[0] == (0, 0.0) and len(top) == 50
    assert bsr.calculate_accuracy_metrics(top, 0, 50) == (1.0, 1.0, 1.0)
    assert bsr.parallel_top_k_similarity_search(None, 1, 1, None, 50, target, engine=eng) is None
    store.close()


def test_dimension_mismatch_and_empty_shard():
    rows = co.generate_rows(SEED, 0, 100, 768)
    store = _store_from(rows, row_base=1000)
    eng = bsr.CudaShardEngine(store)
    # lengths differ -> every distance is 1.0 (src/metrics.rs:144-146): first rows of the block win
    got = bsr.compute_local_top_k(None, 0, 1, 5, rows[0][:700], engine=eng)
    assert got == [(1000 + i, 1.0) for i in range(5)]
    got = bsr.compute_local_top_k(None, 0, 1, 5, rows[0], engine=eng)
    assert got[0] == (1000, 0.0)                          # global index = row_base + local
    store.reset()
    assert bsr.compute_local_top_k(None, 0, 1, 5, rows[0], engine=eng) == []
    store.close()


def test_query_batch_through_the_abi():
    rows = co.generate_rows(SEED, 0, 3000, 768)
    store = _store_from(rows)
    eng = bsr.CudaShardEngine(store)
    qs = np.stack([rows[0], rows[77], co.generate_rows(1, 0, 1, 768)[0]])
    idx, dist, count = eng.local_top_k(qs, 10)
    for i, q in enumerate(qs):
        wi, wd = co.compute_local_top_k(rows, 0, 1, 10, q)
        assert count[i] == 10 and idx[i].tolist() == wi.tolist()
        assert _bits(dist[i]).tolist() == _bits(wd).tolist()
    store.close()


# ---- full size (BASELINE.json configs[1]): the complete list against the oracle ----------------------
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_one_million_rows_full_compare(dtype):
    """1 M x 768, the whole top-k against the oracle run over the regenerated rows (block-wise on the
    host threads, oracle/c_oracle.py::scan_generated_parallel -- about a second of CPU time)."""
    n = 1_000_000
    store = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n, dtype=dtype)
    store.append_synthetic(SEED, 0, n)
    store.set_tensor_min_batch(0)
    eng = bsr.CudaShardEngine(store)
    threads = max(1, len(os.sched_getaffinity(0)))
    q = co.generate_rows(SEED + 0x1000, 0, 1, 768)[0]
    for query, self_query in ((store.get(0), True), (q, False)):
        for k in (10, 50):
            top = bsr.compute_local_top_k(None, 0, 1, k, query, engine=eng)
            wi, wd = co.scan_generated_parallel(SEED, 0, n, k, query, threads=threads, bf16=(dtype == "bf16"))
            _assert_pairs(top, wi, _bits(wd))
            if self_query:
                assert top[0] == (0, 0.0)
    # ... and as one batch of four through the shared pass
    qs = np.stack([store.get(0), q, store.get(999_999), co.generate_rows(SEED + 0x1001, 0, 1, 768)[0]])
    idx, dist, cnt = eng.local_top_k(qs, 10)
    for i in range(4):
        wi, wd = co.scan_generated_parallel(SEED, 0, n, 10, qs[i], threads=threads, bf16=(dtype == "bf16"))
        assert idx[i].tolist() == wi.tolist() and _bits(dist[i]).tolist() == _bits(wd).tolist()
    store.close()
