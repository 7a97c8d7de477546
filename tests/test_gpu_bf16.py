"""GPU: BF16 stores.  Rows are rounded (RNE) to bfloat16 at load; the exact scan then applies the
reference arithmetic to the rounded rows, so results equal the oracle run on
oracle.round_bf16(rows) bit for bit (the query stays binary32)."""
import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001


def _bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def test_bf16_roundtrip_and_synthetic():
    rows = co.generate_rows(SEED, 0, 333, 768)
    rows[5, 7] = np.nan
    rows[6, 0] = np.inf
    rows[7] = 1e-41                                    # denormal: rounds within bf16's range
    st = bsr.GpuVectorstore(None, True, dtype="bf16")
    st.append_many(rows[:100])
    st.append_many(rows[100:])
    want = co.round_bf16(rows)
    got = st.get_many_array(None)
    assert np.array_equal(_bits(got), _bits(want))
    st.close()
    st = bsr.GpuVectorstore(None, True, dim=768, dtype="bf16")
    st.append_synthetic(SEED, 0, 1000)
    st.append_synthetic(SEED, 1000, 77)
    assert np.array_equal(_bits(st.get_many_array(None)), _bits(co.round_bf16(co.generate_rows(SEED, 0, 1077, 768))))
    assert st.scan_bytes() == 1077 * 768 * 2
    st.close()


@pytest.mark.parametrize("n,k,dim", [(1000, 10, 768), (4097, 50, 768), (65, 10, 768), (63, 100, 768),
                                     (20000, 1024, 768), (777, 20, 100), (500, 10, 5), (3000, 10, 1000)])
def test_bf16_exact_scan_matches_oracle_on_rounded_rows(n, k, dim):
    rows = co.generate_rows(SEED, 0, n, dim)
    st = bsr.GpuVectorstore(None, True, dtype="bf16")
    st.append_many(rows)
    eng = bsr.CudaShardEngine(st)
    rr = co.round_bf16(rows)
    for q in (rr[0], rows[1], co.generate_rows(9, 0, 1, dim)[0]):   # rr[0]: identity shortcut
        got = bsr.compute_local_top_k(None, 0, 1, k, q, engine=eng)
        wi, wd = co.compute_local_top_k(rr, 0, 1, k, q)
        assert [p[0] for p in got] == wi.tolist()
        assert _bits([p[1] for p in got]).tolist() == _bits(wd).tolist()
    st.close()


def test_bf16_query_batch():
    rows = co.generate_rows(SEED, 0, 5000, 768)
    st = bsr.GpuVectorstore(None, True, dtype="bf16")
    st.append_many(rows)
    eng = bsr.CudaShardEngine(st)
    rr = co.round_bf16(rows)
    qs = np.concatenate([rr[:2], co.generate_rows(3, 0, 9, 768)])
    idx, dist, cnt = eng.local_top_k(qs, 10)
    for i, q in enumerate(qs):
        wi, wd = co.compute_local_top_k(rr, 0, 1, 10, q)
        assert idx[i].tolist() == wi.tolist() and _bits(dist[i]).tolist() == _bits(wd).tolist()
    st.close()
