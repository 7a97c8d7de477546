"""GPU: the per-warp running top-k *after* its fill phase.

On stores small enough for the CPU oracle a full grid gives every warp at most a few panels, so
its K-slot list never fills and the steady-state path (threshold test, K == 32: in-place insert,
K > 32: batched block merge, `WarpTopK::flush`) would go untested.  The `max_grid` option (a
per-handle test hook, bsr_store_set_option) caps the grid so that each warp scans thousands of rows."""
import os

import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001


def _bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


@pytest.fixture
def small_grid(request):
    yield request.param


@pytest.mark.parametrize("small_grid", [1, 3], indirect=True)
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_steady_state_topk_matches_oracle(small_grid, dtype):
    n = 30000
    rows = co.generate_rows(SEED, 0, n, 768)
    # descending quality: later rows are better matches, so every warp keeps inserting
    q = co.generate_rows(SEED + 3, 0, 1, 768)[0]
    w = np.linspace(0.0, 0.9, n, dtype=np.float32)[:, None]
    rows = (rows * (1 - w) + q[None, :] * w).astype(np.float32)
    st = bsr.GpuVectorstore(None, True, dtype=dtype)
    st.append_many(rows)
    st.set_tensor_min_batch(0)
    st.set_option("max_grid", small_grid)
    eng = bsr.CudaShardEngine(st)
    ref_rows = co.round_bf16(rows) if dtype == "bf16" else rows
    for k in (10, 33, 50, 100, 300, 1024):
        for qq in (q, ref_rows[n // 2]):
            got = bsr.compute_local_top_k(None, 0, 1, k, qq, engine=eng)
            wi, wd = co.compute_local_top_k(ref_rows, 0, 1, k, qq)
            assert [p[0] for p in got] == wi.tolist(), (k,)
            assert _bits([p[1] for p in got]).tolist() == _bits(wd).tolist(), (k,)
    # batches share the pass; every query keeps its own list
    qs = np.concatenate([q[None, :], co.generate_rows(SEED + 5, 0, 4, 768)])
    for k in (50, 200):
        idx, dist, cnt = eng.local_top_k(qs, k)
        for i in range(len(qs)):
            wi, wd = co.compute_local_top_k(ref_rows, 0, 1, k, qs[i])
            assert idx[i, : cnt[i]].tolist() == wi.tolist(), (k, i)
            assert _bits(dist[i, : cnt[i]]).tolist() == _bits(wd).tolist(), (k, i)
    st.close()


@pytest.mark.parametrize("small_grid", [3, 148], indirect=True)
def test_fused_final_merge_when_one_cta_holds_the_whole_top_k(small_grid):
    """k <= 128: the last CTA folds every CTA's list, fetching a list's later 32-key blocks only
    while it can still contribute.  Here the best 128 rows sit in panels 0, 3, 6 and 9, which a
    grid of 3 CTAs deals to CTA 0 alone, so its list supplies all four blocks."""
    n = 2000
    rows = co.generate_rows(SEED, 0, n, 768)
    q = co.generate_rows(SEED + 3, 0, 1, 768)[0]
    rng = np.random.default_rng(5)
    for p in (0, 3, 6, 9):
        for r in range(32):
            rows[p * 32 + r] = q + rng.uniform(-0.05, 0.05, 768).astype(np.float32)
    st = bsr.GpuVectorstore(None, True)
    st.append_many(rows)
    eng = bsr.CudaShardEngine(st)
    for k in (33, 50, 64, 100, 128):
        got = bsr.compute_local_top_k(None, 0, 1, k, q, engine=eng)
        wi, wd = co.compute_local_top_k(rows, 0, 1, k, q)
        assert [p[0] for p in got] == wi.tolist(), (k,)
        assert _bits([p[1] for p in got]).tolist() == _bits(wd).tolist(), (k,)
    qs = np.stack([q, rows[7], co.generate_rows(SEED + 5, 0, 1, 768)[0]])
    idx, dist, cnt = eng.local_top_k(qs, 50)
    for i in range(len(qs)):
        wi, wd = co.compute_local_top_k(rows, 0, 1, 50, qs[i])
        assert idx[i, : cnt[i]].tolist() == wi.tolist() and _bits(dist[i, : cnt[i]]).tolist() == _bits(wd).tolist()
    st.close()
