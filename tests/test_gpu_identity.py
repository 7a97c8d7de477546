"""GPU: the identity shortcut (src/metrics.rs:7-19, :148-151) tracked inline by the scan kernels.

A row is "identical" to the query when every |a_i - b_i| <= 1e-10 (NaN passes); its distance is
then exactly 0.0 whatever the cosine formula would give.  The kernels seed a per-lane bit from the
first columns of a panel and clear it chunk by chunk, so rows that differ from the query only in
a late column (any chunk, any code path: single query, narrow and wide batches, the ragged last
chunk, BF16 tiles, 8- and 16-warp CTAs) are the cases to pin against the oracle, bit for bit."""
import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001


def _bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def _rows_with_near_copies(n, dim, rng, with_nan):
    """Random rows plus near copies of the query q.  q is zero at a few probe columns (one per code
    path / chunk position) so that a 5e-11 perturbation there stays within the 1e-10 threshold.
    Per probe column c: a copy with +0.5 at c (not identical: formula distance ~5e-4), a copy with
    5e-11 at c (identical -> exactly 0.0) and, for fp32 stores, a copy with NaN at c (NaN passes
    the identity test, src/metrics.rs:12-16, so the reference says 0.0 where the formula says 2.0)."""
    rows = co.generate_rows(SEED, 0, n, dim)
    q = co.generate_rows(SEED + 7, 0, 1, dim)[0]
    cols = sorted({c for c in (0, 3, 4, 31, 32, 255, 256, dim - 2, dim - 1) if 0 <= c < dim})
    q[cols] = 0.0
    slots = iter(rng.permutation(n)[: 3 * len(cols) + 1].tolist())
    for c in cols:
        i = next(slots)
        rows[i] = q
        rows[i, c] = np.float32(0.5)
        j = next(slots)
        rows[j] = q
        rows[j, c] = np.float32(5e-11)
        if with_nan:
            m = next(slots)
            rows[m] = q
            rows[m, c] = np.nan
    rows[next(slots)] = q        # exact copy
    return rows, q


@pytest.mark.parametrize("n,dim", [(300, 768), (5000, 768), (50000, 768), (777, 770), (400, 100), (90, 36)])
def test_identity_in_any_column_fp32(n, dim):
    rng = np.random.default_rng(n + dim)
    rows, q = _rows_with_near_copies(n, dim, rng, with_nan=True)
    st = bsr.GpuVectorstore(None, True)
    st.append_many(rows)
    eng = bsr.CudaShardEngine(st)
    k = 64
    got = bsr.compute_local_top_k(None, 0, 1, k, q, engine=eng)
    wi, wd = co.compute_local_top_k(rows, 0, 1, k, q)
    assert [p[0] for p in got] == wi.tolist()
    assert _bits([p[1] for p in got]).tolist() == _bits(wd).tolist()
    # sanity of the fixture itself: the NaN copies are in the list at exactly 0.0
    nan_rows = set(np.nonzero(np.isnan(rows).any(axis=1))[0].tolist())
    assert nan_rows and nan_rows <= {p[0] for p in got if p[1] == 0.0}
    # batches: 2 (narrow path), 8 and 13 (wide path + leftovers); the probe query sits at varying slots
    others = co.generate_rows(SEED + 9, 0, 13, dim)
    for nq, at in ((2, 1), (8, 5), (13, 12)):
        qs = others[:nq].copy()
        qs[at] = q
        idx, dist, cnt = eng.local_top_k(qs, 20)
        for i in range(nq):
            wi, wd = co.compute_local_top_k(rows, 0, 1, 20, qs[i])
            assert idx[i, : cnt[i]].tolist() == wi.tolist(), (nq, i)
            assert _bits(dist[i, : cnt[i]]).tolist() == _bits(wd).tolist(), (nq, i)
    st.close()


@pytest.mark.parametrize("n,dim", [(300, 768), (60000, 768), (500, 200)])
def test_identity_in_any_column_bf16(n, dim):
    rng = np.random.default_rng(n + dim + 1)
    rows, q = _rows_with_near_copies(n, dim, rng, with_nan=False)
    st = bsr.GpuVectorstore(None, True, dtype="bf16")
    st.append_many(rows)
    st.set_tensor_min_batch(0)          # exact scan only
    eng = bsr.CudaShardEngine(st)
    rr = co.round_bf16(rows)
    # the exact bf16 scan compares the *rounded* rows with the binary32 query
    for qq in (q, co.round_bf16(q.reshape(1, -1))[0]):
        got = bsr.compute_local_top_k(None, 0, 1, 64, qq, engine=eng)
        wi, wd = co.compute_local_top_k(rr, 0, 1, 64, qq)
        assert [p[0] for p in got] == wi.tolist()
        assert _bits([p[1] for p in got]).tolist() == _bits(wd).tolist()
    qs = co.generate_rows(SEED + 9, 0, 8, dim)
    qs[3] = q
    idx, dist, cnt = eng.local_top_k(qs, 20)
    for i in range(8):
        wi, wd = co.compute_local_top_k(rr, 0, 1, 20, qs[i])
        assert idx[i, : cnt[i]].tolist() == wi.tolist()
        assert _bits(dist[i, : cnt[i]]).tolist() == _bits(wd).tolist()
    st.close()
