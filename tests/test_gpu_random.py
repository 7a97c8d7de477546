"""GPU: randomized parity sweep of the exact path -- random store sizes, dimensions, k, shard counts,
value scales and injected special rows -- every result compared bit for bit with the oracle."""
import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


@pytest.mark.parametrize("seed", range(12))
def test_random_configurations(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 255, 1000, 4099, 20011]))
    dim = int(rng.choice([1, 3, 4, 7, 64, 100, 257, 768, 1536]))
    k = int(rng.choice([1, 2, 10, 32, 33, 50, 128, 1000]))
    dtype = "bf16" if seed % 3 == 2 else "f32"
    scale = float(rng.choice([1.0, 1e-3, 1e3, 1e-20, 1e18]))
    rows = (rng.uniform(-1, 1, (n, dim)) * scale).astype(np.float32)
    # special rows: zero, duplicates (ties), NaN, inf, sign flips
    for _ in range(min(n, 6)):
        i, j = int(rng.integers(n)), int(rng.integers(n))
        what = int(rng.integers(6))
        if what == 0:
            rows[i] = 0
        elif what == 1:
            rows[i] = rows[j]
        elif what == 2:
            rows[i, int(rng.integers(dim))] = np.nan
        elif what == 3:
            rows[i, int(rng.integers(dim))] = np.inf
        elif what == 4:
            rows[i] = -rows[j]
        else:
            rows[i] = rows[j] * np.float32(2.0)
    st = bsr.GpuVectorstore(None, True, dtype=dtype, row_base=int(rng.integers(0, 1000)))
    st.append_many(rows)
    if dtype == "bf16":
        st.set_tensor_min_batch(0)          # exact scan only
    eng = bsr.CudaShardEngine(st)
    ref_rows = co.round_bf16(rows) if dtype == "bf16" else rows
    queries = [ref_rows[int(rng.integers(n))], (rng.uniform(-1, 1, dim) * scale).astype(np.float32),
               np.zeros(dim, np.float32)]
    qs = np.stack(queries)
    idx, dist, cnt = eng.local_top_k(qs, k)                # one batched call (groups of 8/4/2/1)
    for qi, q in enumerate(queries):
        wi, wd = co.compute_local_top_k(ref_rows, 0, 1, k, q)
        assert cnt[qi] == len(wi), (seed, n, dim, k, dtype)
        assert (idx[qi, :cnt[qi]] - np.uint64(st.row_base)).tolist() == wi.tolist(), (seed, n, dim, k, dtype, qi)
        assert _bits(dist[qi, :cnt[qi]]).tolist() == _bits(wd).tolist(), (seed, n, dim, k, dtype, qi)
    st.close()
