"""GPU: the dot-product metric (BSR_DOT, include/bsr.h) against the oracle's restatement
(oracle/bsr_oracle.c::bsr_oracle_dot_distance) and the committed fixture tests/golden/dot_metric.npz.

Every path -- the exact scan and the tensor-core filter with exact rescoring -- is compared bit for bit."""
import os

import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
F = np.float32


def _bits(x):
    return np.asarray(x, dtype=F).view(np.uint32)


def _store(rows, **kw):
    st = bsr.GpuVectorstore(None, True, **kw)
    st.append_many(rows)
    st.set_metric("dot")
    assert st.metric() == "dot"
    return st


def _sharded(rows, size, k, q):
    import torch
    n = rows.shape[0]
    cands, engines = [], []
    for r in range(size):
        iv = bsr.interval_by_rank(r, size, n)
        first, cnt = iv.rows(n)
        st = bsr.GpuVectorstore(None, True, dim=rows.shape[1], row_base=min(iv.start_index, 2**32 - 2))
        if cnt:
            st.append_many(rows[first:first + cnt])
        st.set_metric("dot")
        eng = bsr.CudaShardEngine(st)
        engines.append(eng)
        cands.append(eng.search_candidates(q.reshape(1, -1), k))
    merged = engines[0].merge(torch.stack(cands), k)
    idx, dist, count = engines[0].unpack(merged, k)
    for e in engines:
        e.store.close()
    return idx[0, :count[0]], dist[0, :count[0]]


def test_dot_golden_fixture():
    g = np.load(os.path.join(GOLDEN, "dot_metric.npz"))
    rows = g["rows"]
    st = _store(rows)
    eng = bsr.CudaShardEngine(st)
    for qi, q in enumerate(g["queries"]):
        idx, dist, cnt = eng.local_top_k(q, 1024)
        assert cnt[0] == len(rows)
        by_idx = dict(zip(idx[0, :cnt[0]].tolist(), _bits(dist[0, :cnt[0]]).tolist()))
        want = g[f"all_dist_bits_q{qi}"]
        for i in range(len(rows)):
            # NaN distances: any NaN payload is a NaN (the order treats them alike)
            if np.isnan(want[i:i + 1].view(F)[0]):
                assert np.isnan(np.array([by_idx[i]], np.uint32).view(F)[0])
            else:
                assert by_idx[i] == want[i], (qi, i)
        for k in g["ks"]:
            idx, dist, cnt = eng.local_top_k(q, int(k))
            wi, wb = g[f"idx_q{qi}_k{k}_s1"], g[f"dist_bits_q{qi}_k{k}_s1"]
            assert idx[0, :cnt[0]].tolist() == wi.tolist(), (qi, k)
            ok = ~np.isnan(wb.view(F))
            assert np.array_equal(_bits(dist[0, :cnt[0]])[ok], wb[ok])
            for size in g["sizes"]:
                if size == 1:
                    continue
                gi, gd = _sharded(rows, int(size), int(k), q)
                assert gi.tolist() == g[f"idx_q{qi}_k{k}_s{size}"].tolist(), (qi, k, size)
    st.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_dot_exact_paths_match_oracle(dtype):
    n = 20011
    rows = co.generate_rows(SEED, 0, n, 768)
    rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(F)
    st = _store(rows, dtype=dtype)
    st.set_tensor_min_batch(0)
    eng = bsr.CudaShardEngine(st)
    ref = co.round_bf16(rows) if dtype == "bf16" else rows
    qs = np.stack([ref[0], ref[123] * F(0.5), co.generate_rows(SEED + 9, 0, 1, 768)[0],
                   np.zeros(768, F)] + [co.generate_rows(SEED + 10 + i, 0, 1, 768)[0] for i in range(7)])
    for k in (10, 50, 200):
        # one by one (QB = 1) and as a batch (QB = 8 + 2 + 1): every query keeps its own list
        idx_b, dist_b, cnt_b = eng.local_top_k(qs, k)
        for i, q in enumerate(qs):
            wi, wd = co.compute_local_top_k(ref, 0, 1, k, q, metric=1)
            idx, dist, cnt = eng.local_top_k(q, k)
            assert idx[0, :cnt[0]].tolist() == wi.tolist(), (k, i)
            assert _bits(dist[0, :cnt[0]]).tolist() == _bits(wd).tolist(), (k, i)
            assert idx_b[i, :cnt_b[i]].tolist() == wi.tolist(), (k, i)
            assert _bits(dist_b[i, :cnt_b[i]]).tolist() == _bits(wd).tolist(), (k, i)
    # the metric is a property of the handle: back to cosine, the reference's distances return
    st.set_metric("cosine")
    wi, wd = co.compute_local_top_k(ref, 0, 1, 10, qs[2])
    idx, dist, cnt = eng.local_top_k(qs[2], 10)
    assert idx[0].tolist() == wi.tolist() and _bits(dist[0]).tolist() == _bits(wd).tolist()
    st.close()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("nq,k", [(64, 10), (300, 100)])
def test_dot_tensor_path_is_bit_exact(nq, k, dtype):
    n = 30000
    rows = co.generate_rows(SEED, 0, n, 768)
    rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(F)
    st = _store(rows, dtype=dtype)
    st.set_tensor_min_batch(32)
    eng = bsr.CudaShardEngine(st)
    ref = co.round_bf16(rows) if dtype == "bf16" else rows
    qs = co.generate_rows(SEED + 0x2000, 0, nq, 768)
    qs = (qs / np.linalg.norm(qs.astype(np.float64), axis=1, keepdims=True)).astype(F)
    qs[0] = ref[5]
    idx, dist, cnt = eng.local_top_k(qs, k)
    for i in range(0, nq, max(1, nq // 24)):
        wi, wd = co.compute_local_top_k(ref, 0, 1, k, qs[i], metric=1)
        assert cnt[i] == k and idx[i].tolist() == wi.tolist(), i
        assert _bits(dist[i]).tolist() == _bits(wd).tolist(), i
    st.close()
