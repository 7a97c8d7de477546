"""GPU: the tcgen05 / TMEM batched path (BF16 stores, batches >= tensor_min_batch).

Oracle: the reference arithmetic applied to the bf16-rounded rows AND bf16-rounded queries.
Products of bf16 values are exact in binary32; only the accumulation order differs on the tensor
core, so the stated bound for this path is |distance - oracle| <= 2e-6 (absolute; north_star
allows a looser bound than the fp32 path's for bf16), and indices must match except where two
oracle distances are closer than that bound."""
import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001
TOL = 2e-6


def _check(idx, dist, cnt, rr, qr, k):
    for i in range(len(qr)):
        wi, wd = co.compute_local_top_k(rr, 0, 1, k, qr[i])
        assert cnt[i] == len(wi)
        gi, gd = idx[i, :cnt[i]], dist[i, :cnt[i]]
        assert np.all(np.diff(gd) >= 0)
        assert np.max(np.abs(gd - wd)) <= TOL, (i, np.max(np.abs(gd - wd)))
        for j in np.nonzero(gi != wi)[0]:
            # a swap is only acceptable between near-ties: the row the GPU reports must be within
            # TOL of the oracle's j-th distance
            d = co.cosine_distance(rr[int(gi[j])], qr[i])
            assert abs(float(d) - float(wd[j])) <= TOL, (i, j, gi[j], wi[j])
        assert len(set(gi.tolist())) == len(gi)


@pytest.mark.parametrize("n,nq,k", [(20000, 200, 10), (4099, 128, 100), (70, 40, 10), (9000, 33, 50)])
def test_tensor_path_matches_oracle_within_bound(n, nq, k):
    rows = co.generate_rows(SEED, 0, n, 768)
    st = bsr.GpuVectorstore(None, True, dtype="bf16")
    st.append_many(rows)
    st.set_tensor_min_batch(32)
    eng = bsr.CudaShardEngine(st)
    rr = co.round_bf16(rows)
    qs = co.generate_rows(SEED + 0x1000, 0, nq, 768)
    qs[0] = rr[0]                                   # self-query: identity shortcut -> exactly 0.0
    qs[1] = rr[min(n - 1, 77)] * np.float32(0.5)    # same direction: sim ~ 1 but not identical
    qr = co.round_bf16(qs)
    idx, dist, cnt = eng.local_top_k(qs, k)
    assert idx[0, 0] == 0 and dist[0, 0] == 0.0
    _check(idx, dist, cnt, rr, qr, k)
    # below the batch threshold the same store answers with the exact scan (bit-exact, query not rounded)
    idx2, dist2, cnt2 = eng.local_top_k(qs[:5], k)
    for i in range(5):
        wi, wd = co.compute_local_top_k(rr, 0, 1, k, qs[i])
        assert idx2[i, :cnt2[i]].tolist() == wi.tolist()
        assert dist2[i, :cnt2[i]].view(np.uint32).tolist() == wd.view(np.uint32).tolist()
    st.close()


def test_tensor_path_other_dim_and_edge_rows():
    n, dim = 3000, 200
    rows = co.generate_rows(SEED, 0, n, dim)
    rows[5] = 0.0
    rows[9] = rows[8]
    st = bsr.GpuVectorstore(None, True, dtype="bf16")
    st.append_many(rows)
    eng = bsr.CudaShardEngine(st)
    rr = co.round_bf16(rows)
    qs = co.generate_rows(3, 0, 64, dim)
    qs[2] = rr[8]
    qs[3] = 0.0                                      # zero query: every distance 1.0
    idx, dist, cnt = eng.local_top_k(qs, 10)
    _check(idx, dist, cnt, rr, co.round_bf16(qs), 10)
    assert idx[2, :2].tolist() == [8, 9] and dist[2, 0] == 0.0 and dist[2, 1] == 0.0
    # zero query: row 5 (all zeros) is *identical* -> 0.0 (the identity test precedes the zero-magnitude
    # rule, src/metrics.rs:148-159); every other row is at 1.0, in index order
    assert idx[3].tolist() == [5, 0, 1, 2, 3, 4, 6, 7, 8, 9]
    assert dist[3, 0] == 0.0 and np.all(dist[3, 1:] == 1.0)
    st.close()
