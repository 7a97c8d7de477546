"""GPU: contract details of the C ABI that the parity tests do not reach -- published candidate
slots (include/bsr.h), calls on different streams sharing one handle, search -> append -> search,
bsr_merge_host's de-dup, and the reference's own store size (2 305 x 768, k = 50)."""
import ctypes as C

import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from better_search_rag_rust_b200 import _lib
from better_search_rag_rust_b200 import search as S
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
SEED = 0x5EED0001
F = np.float32
KEY_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)


def _bits(x):
    return np.asarray(x, dtype=F).view(np.uint32)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("k", [1, 10, 32, 33, 50, 100, 129, 300])
def test_published_slots_past_top_k_are_empty(k, dtype):
    """include/bsr.h: the first min(top_k, rows) slots hold the top-k, every other slot UINT64_MAX --
    for every merge route (k <= 32 fused, k <= 128 fused fold, k > 128 merge tree)."""
    import torch
    n = 5000
    rows = co.generate_rows(SEED, 0, n, 768)
    st = bsr.GpuVectorstore(None, True, dtype=dtype)
    st.append_many(rows)
    st.set_tensor_min_batch(0)
    eng = bsr.CudaShardEngine(st)
    ref = co.round_bf16(rows) if dtype == "bf16" else rows
    qs = np.stack([ref[0], co.generate_rows(SEED + 1, 0, 1, 768)[0], co.generate_rows(SEED + 2, 0, 1, 768)[0]])
    cand = eng.search_candidates(qs, k)
    keys = cand.cpu().numpy().view(np.uint64)
    slots = bsr.candidate_slots(k)
    assert keys.shape == (3, slots)
    for i, q in enumerate(qs):
        wi, wd = co.compute_local_top_k(ref, 0, 1, k, q)
        assert np.array_equal(keys[i, :k], S.pack_candidates(wi, wd, k))
        assert np.all(keys[i, k:] == KEY_MAX)
    # the cross-rank merge publishes the same way
    merged = eng.merge(torch.stack([cand, torch.full_like(cand, -1)]), k).cpu().numpy().view(np.uint64)
    assert np.array_equal(merged[:, :k], keys[:, :k])
    assert np.all(merged[:, k:] == KEY_MAX)
    st.close()


def test_calls_on_different_streams_share_one_handle_safely():
    """All calls on a handle share one scratch set; an asynchronous call on a caller's stream followed
    at once by a call on another stream (the handle's own, inside bsr_search) must not overlap it."""
    import torch
    n = 400_000
    st = bsr.GpuVectorstore(None, True, dim=768, capacity_rows=n)
    st.append_synthetic(SEED, 0, n)
    eng = bsr.CudaShardEngine(st)
    q1 = co.generate_rows(SEED + 0x1000, 0, 1, 768)[0]
    q2 = co.generate_rows(SEED + 0x1001, 0, 1, 768)[0]
    w1 = co.scan_generated_parallel(SEED, 0, n, 10, q1, threads=8)
    w2 = co.scan_generated_parallel(SEED, 0, n, 10, q2, threads=8)
    side = torch.cuda.Stream()
    for _ in range(20):
        with torch.cuda.stream(side):
            c1 = eng.search_candidates(q1.reshape(1, -1), 10)       # asynchronous, on `side`
        idx2, dist2, cnt2 = eng.local_top_k(q2, 10)                 # handle's own stream, synchronous
        side.synchronize()
        i1, d1 = S.unpack_candidates(c1.cpu().numpy().view(np.uint64)[0], 10)
        assert i1.tolist() == w1[0].tolist() and _bits(d1).tolist() == _bits(w1[1]).tolist()
        assert idx2[0].tolist() == w2[0].tolist() and _bits(dist2[0]).tolist() == _bits(w2[1]).tolist()
    st.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_search_append_search(dtype):
    """append_many against a store that has already been searched: the allocation grows (new panel,
    norm and scratch buffers), earlier rows keep their indices, new rows are found."""
    dim = 768
    first = co.generate_rows(SEED, 0, 3001, dim)
    st = bsr.GpuVectorstore(None, True, dtype=dtype)          # no capacity hint: every append reallocates
    st.append_many(first)
    st.set_tensor_min_batch(0)
    eng = bsr.CudaShardEngine(st)
    rnd = (lambda a: co.round_bf16(a)) if dtype == "bf16" else (lambda a: a)
    q = co.generate_rows(SEED + 0x1000, 0, 1, dim)[0]
    rows = first
    for step, extra_n in enumerate((1, 31, 4096, 20000)):
        for k in (10, 50, 200):
            idx, dist, cnt = eng.local_top_k(q, k)
            wi, wd = co.compute_local_top_k(rnd(rows), 0, 1, k, q)
            assert idx[0, :cnt[0]].tolist() == wi.tolist(), (step, k)
            assert _bits(dist[0, :cnt[0]]).tolist() == _bits(wd).tolist(), (step, k)
        extra = co.generate_rows(SEED + 50 + step, 0, extra_n, dim)
        extra[0] = q * F(0.5 + 0.1 * step)                    # a row that must enter the top-k at once
        st.append_many(extra)
        rows = np.concatenate([rows, extra])
        assert st.get_count() == len(rows)
        idx, dist, cnt = eng.local_top_k(q, 10)
        assert len(rows) - extra_n in idx[0].tolist()
    wi, wd = co.compute_local_top_k(rnd(rows), 0, 1, 50, q)
    idx, dist, cnt = eng.local_top_k(q, 50)
    assert idx[0].tolist() == wi.tolist() and _bits(dist[0]).tolist() == _bits(wd).tolist()
    assert np.array_equal(st.get_many_array(None).view(np.uint32), rnd(rows).view(np.uint32))
    st.close()


def test_merge_host_dedups_like_the_reference():
    """compute_global_top_k's HashSet pass (src/mpi_helpers/metrics.rs:156-168): an index seen twice
    keeps its first occurrence in sorted order."""
    gi = np.array([7, 3, 7, 9, 3, 11, 7], np.uint64)
    gd = np.array([0.5, 0.25, 0.125, 0.75, 0.25, 0.0625, 0.125], F)
    for k in (1, 3, 4, 10):
        got = bsr.compute_global_top_k(gi.tolist(), gd.tolist(), k, device=0)
        wi, wd = co.compute_global_top_k(gi, gd, k)
        assert [p[0] for p in got] == wi.tolist() and _bits([p[1] for p in got]).tolist() == _bits(wd).tolist()
    rng = np.random.default_rng(3)
    gi = rng.integers(0, 300, 2000).astype(np.uint64)
    gd = rng.uniform(0, 2, 2000).astype(F)
    got = bsr.compute_global_top_k(gi.tolist(), gd.tolist(), 100, device=0)
    wi, wd = co.compute_global_top_k(gi, gd, 100)
    assert [p[0] for p in got] == wi.tolist() and _bits([p[1] for p in got]).tolist() == _bits(wd).tolist()


def test_the_references_own_store_size():
    """src/main.rs:110-131 on the JabRef store the reference ships numbers for: 2 305 x 768, k = 50,
    query = row 0; every rank count the reference's makefile/SLURM script uses (1, 2, 4, 6)."""
    import torch
    n, k = 2305, 50
    rows = co.generate_rows(SEED, 0, n, 768)
    for size in (1, 2, 4, 6):
        cands, engines = [], []
        for r in range(size):
            iv = bsr.interval_by_rank(r, size, n)
            first, cnt = iv.rows(n)
            st = bsr.GpuVectorstore(None, True, dim=768, row_base=iv.start_index)
            st.append_many(rows[first:first + cnt])
            engines.append(bsr.CudaShardEngine(st))
            cands.append(engines[-1].search_candidates(rows[0].reshape(1, -1), k))
        merged = engines[0].merge(torch.stack(cands), k)
        idx, dist, cnt = engines[0].unpack(merged, k)
        wi, wd = co.parallel_top_k_similarity_search(rows, size, k, rows[0])
        assert idx[0].tolist() == wi.tolist() and _bits(dist[0]).tolist() == _bits(wd).tolist()
        top = list(zip(idx[0].tolist(), dist[0].tolist()))
        assert bsr.calculate_accuracy_metrics(top, 0, k) == (1.0, 1.0, 1.0)
        for e in engines:
            e.store.close()


def test_unknown_option_is_rejected():
    st = bsr.GpuVectorstore(None, True, dim=8)
    with pytest.raises(bsr.BsrError):
        st.set_option("no_such_option", 1)
    with pytest.raises(bsr.BsrError):
        _lib.check(_lib.lib().bsr_store_set_metric(st.handle, 7))
    st.close()
