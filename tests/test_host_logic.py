"""CPU: host-side mirror of the reference interface (no GPU, no compute through the library)."""
import os

import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from better_search_rag_rust_b200 import search as S
from better_search_rag_rust_b200 import vectorstore as V
from oracle import c_oracle as co

F = np.float32


def test_interval_by_rank_mirrors_reference_arithmetic():
    for count in list(range(0, 30)) + [1000, 1027, 100_000_000]:
        for size in (1, 2, 3, 4, 7, 8, 41):
            for rank in range(size):
                iv = bsr.interval_by_rank(rank, size, count)
                assert (iv.start_index, iv.end_index) == co.interval_by_rank(rank, size, count)
    iv = bsr.interval_by_rank(3, 4, 5)                   # start 6 > end 5
    assert iv.get_count() == (5 - 6) % (1 << 64)         # release-build wrap (load_balance.rs:14-16)
    assert iv.rows(5) == (5, 0)                          # polars' clamping slice selects nothing
    assert bsr.interval_by_rank(7, 8, 100_000_000).rows(100_000_000) == (87_500_000, 12_500_000)


def test_slice_by_rank():
    arr = list(range(10))
    parts = [bsr.slice_by_rank(r, 4, arr).slice for r in range(4)]
    assert parts == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9]]
    with pytest.raises(IndexError):                      # Rust would panic on arr[6..5]
        bsr.slice_by_rank(3, 4, list(range(5)))


def test_candidate_keys_order_like_the_reference_sort():
    rng = np.random.default_rng(0)
    dist = rng.uniform(0, 2, 500).astype(F)
    dist[10:20] = dist[10]                               # ties
    dist[30] = 0.0
    dist[31] = -0.0
    idx = rng.permutation(500).astype(np.uint64)
    keys = S.pack_candidates(idx, dist, 512)
    order = np.argsort(keys[:500], kind="stable")
    want = sorted(range(500), key=lambda i: (float(dist[i]), int(idx[i])))
    assert order.tolist() == want
    gi, gd = S.unpack_candidates(np.sort(keys), 500)
    assert gi.tolist() == [int(idx[i]) for i in want]
    assert np.array_equal(gd, np.abs(dist[want]))        # -0.0 canonicalised to +0.0
    assert np.all(keys[500:] == S.KEY_MAX)
    assert S.unpack_candidates(keys[500:], 12)[0].size == 0


def test_candidate_slots():
    assert [bsr.candidate_slots(k) for k in (1, 10, 32, 33, 100, 1024)] == [32, 32, 32, 64, 128, 1024]
    for bad in (0, 1025, -1):
        with pytest.raises(bsr.BsrError):
            bsr.candidate_slots(bad)


def test_accuracy_metrics_and_printing(capsys):
    top = [(0, 0.0), (60, 0.8749089), (1230, 0.88296163)]
    assert bsr.calculate_accuracy_metrics(top, 0, 50) == (1.0, 1.0, 1.0)
    assert bsr.calculate_accuracy_metrics(top, 1230, 50) == co.calculate_accuracy_metrics([0, 60, 1230], 1230, 50)
    assert bsr.calculate_accuracy_metrics(top, 1230, 2) == (pytest.approx(1 / 3), 0.0, 1.0)
    assert bsr.calculate_accuracy_metrics(top, 99, 50) == (0.0, 0.0, 0.0)
    bsr.print_top_k_results(top)
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "Global top-3 results:"
    assert out[1] == "  1. Index: 0, Distance: 0"
    assert out[2] == "  2. Index: 60, Distance: 0.8749089"
    assert bsr.format_accuracy_metrics(1.0, 1.0, 1.0, 50).splitlines()[2] == "  Recall@50: 1.0000"


def test_parquet_helpers_roundtrip(tmp_path):
    rows = co.generate_rows(3, 0, 123, 768)
    p = str(tmp_path / "sub" / "global.parquet")
    os.makedirs(os.path.dirname(p))
    V._write_parquet_rows(p, rows)
    back = V._read_parquet_rows(p)
    assert back.dtype == F and np.array_equal(back.view(np.uint32), rows.view(np.uint32))
    # a plain list<float> column (what another writer might produce) decodes the same
    import pyarrow as pa
    import pyarrow.parquet as pq
    p2 = str(tmp_path / "plain.parquet")
    pq.write_table(pa.table({"embeddings": pa.array(rows.tolist(), type=pa.list_(pa.float32()))}), p2)
    assert np.array_equal(V._read_parquet_rows(p2), rows)
    # a missing file is created empty, as read_parquet does (polars.rs:64-73)
    p3 = str(tmp_path / "new" / "rank_0.parquet")
    assert V._read_parquet_rows(p3).shape[0] == 0 and os.path.exists(p3)


def test_store_paths():
    # src/mpi_helpers/vectorstore.rs:5-20 naming contract, checked without opening a device
    import inspect
    assert 'f"rank_{rank}.parquet"' in inspect.getsource(bsr.get_local_vstore)
    assert '"global.parquet"' in inspect.getsource(bsr.get_global_vstore)


def test_timing_report_format():
    """BenchmarkManager::generate_report text (src/mpi_helpers/benchmark.rs:313-416)."""
    from better_search_rag_rust_b200.benchmark import BenchmarkManager, BenchmarkTiming, BenchmarkTimer
    m = BenchmarkManager(0)
    m.record(BenchmarkTiming("similarity_search", 0.00123, None))
    m.record(BenchmarkTiming("metrics_calculation", 2.5, 7))
    t = BenchmarkTimer.start("total_execution").stop()
    assert t.name == "total_execution" and t.seconds() >= 0.0
    rep = m.generate_report(None, {"similarity_search": 1.23})
    assert rep.splitlines() == [
        "==== PARALLEL PERFORMANCE REPORT ====", "",
        "Operation: metrics_calculation",
        "  Min time: 2.5000 sec (Rank 0)", "  Max time: 2.5000 sec (Rank 0)", "  Avg time: 2.5000 sec",
        "Operation: similarity_search",
        "  Min time: 0.0012 sec (Rank 0)", "  Max time: 0.0012 sec (Rank 0)", "  Avg time: 0.0012 sec",
        "  Speedup: 1000.00x"]
    assert BenchmarkManager(1).generate_report(None) == ""
    assert m.get("metrics_calculation").throughput() == pytest.approx(2.8)


def test_parquet_row_source_reads_only_the_row_groups_it_needs(tmp_path, monkeypatch):
    """open_shard's loader: footers give the row counts, only the row groups overlapping the rank's
    block are decoded (the reference decodes the whole file on every rank, metrics.rs:23)."""
    from better_search_rag_rust_b200 import vectorstore as V
    rows = np.arange(1000 * 6, dtype=np.float32).reshape(1000, 6)
    path = str(tmp_path / "global.parquet")
    V._write_parquet_rows(path, rows, row_group_size=128)           # 8 row groups
    src = V.ParquetRowSource([path])
    assert src.count == 1000 and not src.has_nulls
    reads = []
    pf = src._files[0]
    real = pf.read_row_group
    monkeypatch.setattr(pf, "read_row_group", lambda rg, **kw: (reads.append(rg), real(rg, **kw))[1])
    for size in (1, 3, 8, 1001):
        for rank in range(min(size, 9)):
            iv = bsr.interval_by_rank(rank, size, 1000)
            first, n = iv.rows(1000)
            reads.clear()
            got = list(src.blocks(first, n))
            cat = np.concatenate(got) if got else np.zeros((0, 6), np.float32)
            assert np.array_equal(cat, rows[first:first + n])
            want_groups = sorted({g for g in range(8) if n and g * 128 < first + n and (g + 1) * 128 > first})
            assert sorted(reads) == want_groups
    assert list(src.blocks(1000, 0)) == []


def test_rank_files_concatenate_in_merge_order(tmp_path):
    """rank_{r}.parquet taken in rank order are the rows merge_vector_stores writes to global.parquet
    (tasks.rs:186-197; empty rank files are skipped there and contribute nothing here)."""
    from better_search_rag_rust_b200 import vectorstore as V
    rng = np.random.default_rng(2)
    parts = [rng.uniform(-1, 1, (n, 16)).astype(np.float32) for n in (300, 0, 17, 450)]
    for r, p in enumerate(parts):
        V._write_parquet_rows(str(tmp_path / f"rank_{r}.parquet"), p if len(p) else np.zeros((0, 0), np.float32),
                              row_group_size=100)
    glob = np.concatenate([p for p in parts if len(p)])
    assert [os.path.basename(p) for p in bsr.rank_file_paths(tmp_path)] == [f"rank_{r}.parquet" for r in range(4)]
    src = V.ParquetRowSource(bsr.rank_file_paths(tmp_path))
    assert src.count == len(glob)
    for size in (1, 2, 5):
        for rank in range(size):
            first, n = bsr.interval_by_rank(rank, size, len(glob)).rows(len(glob))
            cat = np.concatenate(list(src.blocks(first, n)) or [np.zeros((0, 16), np.float32)])
            assert np.array_equal(cat, glob[first:first + n])


def test_parquet_row_source_with_null_rows_falls_back_to_whole_file_semantics(tmp_path):
    """get_many drops null rows (polars.rs:142), so indices are positions among the non-null rows."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from better_search_rag_rust_b200 import vectorstore as V
    vals = [[1.0, 2.0], None, [3.0, 4.0], [5.0, 6.0], None, [7.0, 8.0]]
    path = str(tmp_path / "global.parquet")
    pq.write_table(pa.table({"embeddings": pa.array(vals, type=pa.large_list(pa.float32()))}), path, row_group_size=2)
    src = V.ParquetRowSource([path])
    assert src.has_nulls
    got = np.concatenate(list(src.blocks(1, 2)))
    assert got.tolist() == [[3.0, 4.0], [5.0, 6.0]]
