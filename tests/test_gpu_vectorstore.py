"""GPU: GpuVectorstore against the behaviour the reference's own store tests pin
(src/vectorstore/polars.rs:249-395): append / append_many / slice / persist+reload round trips
return bit-identical rows in insertion order; reload of an empty file errors unless forced."""
import os

import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _rows(n, dim=768, seed=21):
    return co.generate_rows(seed, 0, n, dim)


def test_append_and_get(tmp_path):
    s = bsr.get_global_vstore(tmp_path, True)
    assert s.get_count() == 0 and s.get_many() == []
    v = _rows(1)[0]
    s.append(v)
    assert s.get_count() == 1 and np.array_equal(s.get(0), v)
    with pytest.raises(bsr.NoDataError):
        s.get(1)
    s.close()


def test_append_many_and_order(tmp_path):
    s = bsr.get_local_vstore(tmp_path, 3, True)
    assert s.path.endswith("rank_3.parquet")
    a, b = _rows(10), _rows(7, seed=22)
    s.append_many(a)
    s.append_many([])                                     # no-op (polars.rs:104-105)
    s.append_many(b)
    got = s.get_many()
    assert len(got) == 17
    assert np.array_equal(np.stack(got), np.concatenate([a, b]))
    with pytest.raises(ValueError):
        s.append_many(np.zeros((2, 5), np.float32))       # wrong row length
    s.close()


def test_slicing_1000_rows(tmp_path):
    s = bsr.get_global_vstore(tmp_path, True)
    rows = _rows(1000)
    s.append_many(rows)
    for off, ln in [(0, 10), (990, 10), (500, 1), (0, 1000), (123, 456)]:
        got = s.get_many(bsr.SliceArgs(off, ln))
        assert np.array_equal(np.stack(got), rows[off:off + ln])
    # DataFrame::slice clamps (polars.rs:125): past-the-end and negative offsets
    assert len(s.get_many(bsr.SliceArgs(995, 100))) == 5
    assert s.get_many(bsr.SliceArgs(5000, 10)) == []
    assert np.array_equal(np.stack(s.get_many(bsr.SliceArgs(-3, 10))), rows[-3:])
    s.close()


def test_persist_reload_roundtrip(tmp_path):
    s = bsr.get_global_vstore(tmp_path / "vstore", True)
    rows = _rows(321)
    s.append_many(rows)
    s.persist()
    assert os.path.exists(tmp_path / "vstore" / "global.parquet")
    t = bsr.get_global_vstore(tmp_path / "vstore", False)     # new(path, empty=false) reads it
    assert t.get_count() == 321
    assert np.array_equal(t.get_many_array().view(np.uint32), rows.view(np.uint32))
    t.append_many(_rows(4, seed=5))
    t.reload(False)                                            # back to what is on disk
    assert t.get_count() == 321
    s.reset()
    assert s.get_count() == 0
    s.close()
    t.close()


def test_reload_empty_errors_unless_forced(tmp_path):
    s = bsr.get_global_vstore(tmp_path, False)                 # missing file -> created empty
    assert os.path.exists(tmp_path / "global.parquet") and s.get_count() == 0
    with pytest.raises(bsr.NoDataError):
        s.reload(False)
    s.reload(True)
    assert s.get_count() == 0
    s.close()


def test_search_from_reference_style_parquet(tmp_path):
    """SKIP_PROCESS=true flow (src/main.rs:42-57,110-134): global.parquet + rank_0.parquet on
    disk, query = row 0 of rank_0.parquet, every rank scans its block of global.parquet."""
    rows = _rows(2305)                                         # JabRef-sized store
    g = bsr.get_global_vstore(tmp_path, True)
    g.append_many(rows)
    g.persist()
    l0 = bsr.get_local_vstore(tmp_path, 0, True)
    l0.append_many(rows[:400])
    l0.persist()
    g.close()
    l0.close()
    local = bsr.get_local_vstore(tmp_path, 0, False)
    target = local.get(0)
    size = 6                                                   # makefile:1 N ?= 6
    parts = [bsr.compute_local_top_k(tmp_path, r, size, 50, target, device=0) for r in range(size)]
    gi = [p[0] for part in parts for p in part]
    gd = [p[1] for part in parts for p in part]
    top = bsr.compute_global_top_k(gi, gd, 50, device=0)
    wi, wd = co.parallel_top_k_similarity_search(rows, size, 50, target)
    assert [p[0] for p in top] == wi.tolist()
    assert np.asarray([p[1] for p in top], np.float32).view(np.uint32).tolist() == wd.view(np.uint32).tolist()
    assert bsr.calculate_accuracy_metrics(top, 0, 50) == (1.0, 1.0, 1.0)
    bsr.release_shards()
    local.close()
