"""GPU: the tcgen05 / TMEM batched path (batches >= tensor_min_batch; bf16 stores, and fp32 stores
through their bf16 shadow).

The tensor cores are a FILTER: they work on bf16-rounded queries (and rows), certify a bound with a
margin of twice the worst-case rounding gap, and every survivor's distance is then recomputed with the
reference's own arithmetic over the stored row and the query as given.  The bar is therefore the exact
scan's: index lists and distance bit patterns equal the oracle's (oracle on the stored rows -- rounded
to bf16 for a bf16 store -- and the UNROUNDED queries)."""
import os

import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001


def _bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def _check(idx, dist, cnt, ref_rows, qs, k, sample=None, metric=0):
    for i in (range(len(qs)) if sample is None else sample):
        wi, wd = co.compute_local_top_k(ref_rows, 0, 1, k, qs[i], metric=metric)
        assert cnt[i] == len(wi), i
        assert idx[i, :cnt[i]].tolist() == wi.tolist(), (i, idx[i, :cnt[i]][:8], wi[:8])
        assert _bits(dist[i, :cnt[i]]).tolist() == _bits(wd).tolist(), i


def _store(rows, dtype, min_batch=32):
    st = bsr.GpuVectorstore(None, True, dtype=dtype)
    st.append_many(rows)
    st.set_tensor_min_batch(min_batch)
    return st, (co.round_bf16(rows) if dtype == "bf16" else rows)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("n,nq,k", [(20000, 200, 10), (4099, 128, 100), (70, 40, 10), (9000, 33, 50),
                                    (30000, 64, 300), (12000, 40, 1024)])
def test_tensor_path_is_bit_exact(n, nq, k, dtype):
    rows = co.generate_rows(SEED, 0, n, 768)
    st, ref = _store(rows, dtype)
    eng = bsr.CudaShardEngine(st)
    qs = co.generate_rows(SEED + 0x1000, 0, nq, 768)
    qs[0] = ref[0]                                   # self-query: identity shortcut -> exactly 0.0
    qs[1] = ref[min(n - 1, 77)] * np.float32(0.5)    # same direction: sim ~ 1 but not identical
    launches0 = st.launch_count()
    idx, dist, cnt = eng.local_top_k(qs, k)
    assert idx[0, 0] == 0 and dist[0, 0] == 0.0
    _check(idx, dist, cnt, ref, qs, k)
    # it really was the tensor path: a handful of launches, not one exact scan per query group
    assert st.launch_count() - launches0 <= 16
    # below the batch threshold the same store answers with the exact scan: same bits
    idx2, dist2, cnt2 = eng.local_top_k(qs[:5], k)
    assert np.array_equal(idx2, idx[:5]) and np.array_equal(_bits(dist2), _bits(dist[:5]))
    st.close()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_tensor_path_other_dim_edge_rows_and_degenerate_queries(dtype):
    n, dim = 3000, 200
    rows = co.generate_rows(SEED, 0, n, dim)
    rows[5] = 0.0
    rows[9] = rows[8]
    rows[11, 7] = np.nan
    rows[13] = np.float32(3e19)                      # squares overflow -> |a| = inf
    st, ref = _store(rows, dtype)
    eng = bsr.CudaShardEngine(st)
    qs = co.generate_rows(3, 0, 64, dim)
    qs[2] = ref[8]
    qs[3] = 0.0                                      # zero query: every distance 1.0 (row 5 is identical: 0.0)
    qs[4, 17] = np.nan                               # NaN query
    qs[5] = np.float32(1e-30)                        # denormal-scale query
    qs[6] = np.float32(3e19)                         # |b| overflows
    idx, dist, cnt = eng.local_top_k(qs, 10)
    _check(idx, dist, cnt, ref, qs, 10)
    assert idx[2, :2].tolist() == [8, 9] and dist[2, 0] == 0.0 and dist[2, 1] == 0.0
    assert idx[3].tolist() == [5, 0, 1, 2, 3, 4, 6, 7, 8, 9]
    assert dist[3, 0] == 0.0 and np.all(dist[3, 1:] == 1.0)
    # a batch of nothing but degenerate queries
    bad = np.zeros((40, dim), np.float32)
    bad[1::2, 3] = np.nan
    idx, dist, cnt = eng.local_top_k(bad, 10)
    _check(idx, dist, cnt, ref, bad, 10, sample=[0, 1, 39])
    st.close()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("nq,k,n", [(40, 10, 20000), (200, 10, 20000), (40, 100, 6000)])
def test_tensor_path_adversarial_row_order_overflows_the_survivor_buffers(nq, k, n, dtype):
    """Rows sorted by *increasing* similarity to the queries: every new row beats the running bound,
    so nothing is pruned and -- with the grid capped to one cluster (`max_grid`, a test hook) --
    each (cluster, query) survivor list overflows its slots many times.  The overflow path
    (`compact_survivors`: keep exactly the best k of the buffer, by exact key) must still give the
    oracle's list."""
    dim = 768
    base = co.generate_rows(SEED + 11, 0, 1, dim)[0]
    noise = co.generate_rows(SEED + 12, 0, n, dim)
    w = np.linspace(0.05, 0.95, n, dtype=np.float32)[:, None]
    rows = (noise * (1 - w) + base[None, :] * w).astype(np.float32)
    st, ref = _store(rows, dtype)
    st.set_option("max_grid", 2)
    eng = bsr.CudaShardEngine(st)
    qs = (base[None, :] + 0.02 * co.generate_rows(SEED + 13, 0, nq, dim)).astype(np.float32)
    idx, dist, cnt = eng.local_top_k(qs, k)
    _check(idx, dist, cnt, ref, qs, k, sample=range(0, nq, max(1, nq // 12)))
    st.close()


def test_tensor_path_dot_metric_and_append():
    """BSR_DOT through the filter (scores not divided by |a|: the margin scales with the largest norm),
    and an fp32 store's shadow following appends."""
    n = 30000
    rows = co.generate_rows(SEED, 0, n, 768)
    rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    rows[100] *= np.float32(3.0)                     # not every row is unit length
    qs = co.generate_rows(SEED + 0x2000, 0, 96, 768)
    for dtype in ("bf16", "f32"):
        st, ref = _store(rows[:20000], dtype)
        st.set_metric("dot")
        eng = bsr.CudaShardEngine(st)
        idx, dist, cnt = eng.local_top_k(qs, 20)
        _check(idx, dist, cnt, ref[:20000], qs, 20, sample=range(0, 96, 7), metric=1)
        st.append_many(rows[20000:])                 # the shadow (fp32) / tiles (bf16) grow
        ref_all = co.round_bf16(rows) if dtype == "bf16" else rows
        idx, dist, cnt = eng.local_top_k(qs, 20)
        _check(idx, dist, cnt, ref_all, qs, 20, sample=range(0, 96, 7), metric=1)
        st.set_metric("cosine")
        idx, dist, cnt = eng.local_top_k(qs, 20)
        _check(idx, dist, cnt, ref_all, qs, 20, sample=range(0, 96, 7))
        st.close()


@pytest.mark.parametrize("dtype,nq", [("bf16", 512), ("bf16", 1024), ("f32", 512)])
def test_tensor_path_at_the_benchmark_shape(dtype, nq):
    """BASELINE.json configs[2]'s own shape class: CTA pairs (MODE 2, cta_group::2), top-100 (two class
    groups, kc = 128), several back-to-back 256-query launches sharing the class-maxima and survivor
    scratch -- on 1 M rows, a sample of 20 queries spread over all launches against the threaded oracle
    (regenerated rows; rounded to bf16 for the bf16 store), bit for bit."""
    n, k = 1_000_000, 100
    st = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n, dtype=dtype)
    st.append_synthetic(SEED, 0, n)
    st.set_tensor_min_batch(32)
    eng = bsr.CudaShardEngine(st)
    qs = co.generate_rows(SEED + 0x3000, 0, nq, 768)
    qs[0] = st.get(0)                                # self-query, first launch
    qs[nq - 1] = st.get(n - 1)                       # self-query, last launch, last row of the store
    idx, dist, cnt = eng.local_top_k(qs, k)
    assert idx[0, 0] == 0 and dist[0, 0] == 0.0
    assert idx[nq - 1, 0] == n - 1 and dist[nq - 1, 0] == 0.0
    threads = max(1, len(os.sched_getaffinity(0)))
    sample = sorted(set([0, 1, 127, 128, 255, 256, 257, nq // 2, nq - 2, nq - 1] +
                        list(range(3, nq, max(1, nq // 10)))))
    assert len(sample) >= 16
    for i in sample:
        wi, wd = co.scan_generated_parallel(SEED, 0, n, k, qs[i], threads=threads, bf16=(dtype == "bf16"))
        assert cnt[i] == k
        assert idx[i].tolist() == wi.tolist(), (i, idx[i][:6], wi[:6])
        assert _bits(dist[i]).tolist() == _bits(wd).tolist(), i
    st.close()
