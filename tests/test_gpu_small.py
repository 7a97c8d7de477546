"""GPU: stores of at most one 32-row panel per warp (up to ~38 k fp32 rows), one host query -- the
latency path (`scan_small_kernel`: every variant of ring depth and epilogue) against the oracle, bit for
bit, and against the streaming kernel (`no_small` option) on the same store.

Variants by size on a 148-SM part (W = 8 warps per CTA):
  rows <= 4 736          one panel per CTA, the whole panel in shared memory (WHOLE), |b| by an idle warp
  rows <= 9 472          two panels per CTA, WHOLE
  rows <= 33 152         3-7 panels per CTA, a ring of 4 KiB chunks per warp
  rows <= 37 888         8 panels per CTA: no idle warp, |b| on the host
  rows <= 75 776         9-16 panels per CTA: 16 warps per CTA (above: the streaming kernel)
and by top_k: k <= 32 / 64 / 128 register epilogues (NB = 1 / 2 / 4), k > 128 the generic one."""
import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001


def _bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def _check(eng, ref_rows, q, k, metric=0):
    idx, dist, cnt = eng.local_top_k(q.reshape(1, -1), k)
    wi, wd = co.compute_local_top_k(ref_rows, 0, 1, k, q, metric=metric)
    assert cnt[0] == len(wi)
    assert idx[0, :cnt[0]].tolist() == wi.tolist()
    assert _bits(dist[0, :cnt[0]]).tolist() == _bits(wd).tolist()
    return idx[0, :cnt[0]].copy(), dist[0, :cnt[0]].copy()


@pytest.mark.parametrize("n", [1, 31, 33, 2305, 4736, 4737, 9472, 9500, 20011, 33152, 33153, 37888, 37889, 50000, 75776, 75777])
@pytest.mark.parametrize("k", [1, 10, 32, 50, 100, 128, 200])
def test_small_store_matches_the_oracle_and_the_streaming_kernel(n, k):
    rows = co.generate_rows(SEED, 0, n, 768)
    if n > 40:
        rows[n // 2] = rows[7]                       # a duplicate row: two identical rows for query 7
        rows[n - 1, 5] = np.nan                      # a NaN row
    st = bsr.GpuVectorstore(None, True)
    st.append_many(rows)
    eng = bsr.CudaShardEngine(st)
    queries = [rows[min(7, n - 1)].copy(),                          # a store row (identity: distance 0.0)
               co.generate_rows(SEED + 99, 0, 1, 768)[0],           # not a store row
               np.zeros(768, np.float32)]                           # |b| = 0
    got = [_check(eng, rows, q, k) for q in queries]
    if n > 40 and k >= 2:
        assert got[0][0][:2].tolist() == [7, n // 2] and got[0][1][0] == 0.0 and got[0][1][1] == 0.0
    # the same kernel with the query in device memory (bsr_search_device; above 8 panels per SM with no idle
    # warp for |b| that is the streaming kernel): same candidates
    import torch
    from better_search_rag_rust_b200 import search as S
    for q, (gi, gd) in zip(queries, got):
        cand = eng.search_candidates(torch.from_numpy(q).cuda().reshape(1, -1), k)
        ci, cd = S.unpack_candidates(cand.cpu().numpy().view(np.uint64)[0], k)
        assert ci.tolist() == gi.tolist() and _bits(cd).tolist() == _bits(gd).tolist()
    st.set_option("no_small", 1)                     # the streaming kernel on the same store: same answers
    for q, (gi, gd) in zip(queries, got):
        idx, dist, cnt = eng.local_top_k(q.reshape(1, -1), k)
        assert idx[0, :cnt[0]].tolist() == gi.tolist()
        assert _bits(dist[0, :cnt[0]]).tolist() == _bits(gd).tolist()
    st.close()


@pytest.mark.parametrize("n,k", [(2305, 50), (20011, 10), (37888, 64)])
def test_small_store_dot_metric_other_dim_and_append(n, k):
    """BSR_DOT through the same kernel; a dimension that is not a multiple of 32 columns falls back to
    the streaming kernel (the chunked ring needs whole 4 KiB chunks) with the same answers; a store that
    grows past the small-store limit changes kernels between two searches."""
    rows = co.generate_rows(SEED, 0, n, 768)
    st = bsr.GpuVectorstore(None, True)
    st.append_many(rows)
    st.set_metric("dot")
    eng = bsr.CudaShardEngine(st)
    q = co.generate_rows(SEED + 5, 0, 1, 768)[0]
    _check(eng, rows, q, k, metric=1)
    st.set_metric("cosine")
    more = co.generate_rows(SEED + 6, 0, 40000 - n, 768)
    st.append_many(more)
    _check(eng, np.concatenate([rows, more]), q, k)
    st.close()
    for dim in (200, 256, 32):                       # 200: streaming kernel; 256, 32: 8 chunks / 1 chunk per panel
        rows_d = co.generate_rows(SEED, 0, n, dim)
        st = bsr.GpuVectorstore(None, True, dim=dim)
        st.append_many(rows_d)
        eng = bsr.CudaShardEngine(st)
        _check(eng, rows_d, rows_d[3].copy(), k)
        _check(eng, rows_d, co.generate_rows(SEED + 8, 0, 1, dim)[0], k)
        st.close()


def test_small_store_reports_its_own_scan_time():
    """One-launch searches carry no timing events; the launch measures itself and
    bsr_store_last_scan_ms reports that."""
    st = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=2305)
    st.append_synthetic(SEED, 0, 2305)
    eng = bsr.CudaShardEngine(st)
    q = st.get(0)
    for _ in range(5):
        eng.local_top_k(q.reshape(1, -1), 10)
    ms = st.last_scan_ms()
    assert 0.001 < ms < 1.0
    st.close()
