"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

The bar for this path is exact equality of (index, distance-bits) lists: the scan keeps the
reference's sequential binary32 arithmetic, so there is no tolerance to state.
"""
import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

SEED = 0x5EED0001


def _bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def _assert_same(got, want_idx, want_dist):
    gi = np.array([p[0] for p in got], dtype=np.uint64)
    gd = np.array([p[1] for p in got], dtype=np.float32)
    assert gi.tolist() == want_idx.tolist()
    assert _bits(gd).tolist() == _bits(want_dist).tolist()


def _store_from(rows, **kw):
    s = bsr.GpuVectorstore(None, True, **kw)
    s.append_many(rows)
    return s


@pytest.mark.parametrize("n,k", [(1000, 10), (4097, 50), (33, 10), (31, 50), (1, 1), (50000, 10)])
def test_local_top_k_matches_oracle(n, k):
    rows = co.generate_rows(SEED, 0, n, 768)
    store = _store_from(rows)
    eng = bsr.CudaShardEngine(store)
    for q in (rows[0], co.generate_rows(SEED + 0x1000, 0, 1, 768)[0]):
        got = bsr.compute_local_top_k(None, 0, 1, k, q, engine=eng)
        wi, wd = co.compute_local_top_k(rows, 0, 1, k, q)
        _assert_same(got, wi, wd)
    store.close()


def test_synthetic_rows_match_generator():
    store = bsr.GpuVectorstore(None, True, dim=768)
    store.append_synthetic(SEED, 0, 1000)
    store.append_synthetic(SEED, 1000, 77)
    got = store.get_many_array(None)
    want = co.generate_rows(SEED, 0, 1077, 768)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    store.close()


def test_roundtrip_rows_bit_identical():
    rows = co.generate_rows(7, 0, 203, 768)
    store = _store_from(rows[:100])
    store.append_many(rows[100:101])
    store.append_many(rows[101:])
    assert store.get_count() == 203
    assert np.array_equal(store.get_many_array(None).view(np.uint32), rows.view(np.uint32))
    assert np.array_equal(store.get(57), rows[57])
    store.close()
