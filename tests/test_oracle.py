"""CPU: the oracle (oracle/bsr_oracle.c) against hand-derived known answers, the reference's
run-time self-query invariant, an independent numpy restatement, and the committed fixtures.

The reference has no test for this path, so these are the known answers
that pin the restatement; parity with the Rust reference itself stays "unpinned"."""
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import numpy_oracle as no

F = np.float32
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def bits(x):
    return np.asarray(x, dtype=F).view(np.uint32)


# ---- cosine_distance: src/metrics.rs:143-165 ------------------------------------------------
def test_known_answers():
    a = np.array([1, 2, 3, 4], F)
    assert co.cosine_distance(a, a) == 0.0                      # identity shortcut (:148-151)
    assert co.cosine_distance([1, 0], [0, 1]) == 1.0            # orthogonal
    assert co.cosine_distance([1, 0], [-1, 0]) == 2.0           # opposite
    assert co.cosine_distance([0, 0, 0], [1, 2, 3]) == 1.0      # zero magnitude (:157-159)
    assert co.cosine_distance([1, 2, 3], [0, 0, 0]) == 1.0
    assert co.cosine_distance([1, 2], [1, 2, 3]) == 1.0         # length mismatch (:144-146)
    assert co.cosine_distance([], []) == 1.0                    # empty
    # [1,0] vs [1,1]: dot 1, |a| 1, |b| sqrt(2) in f32, sim = 1/sqrt(2) rounded, dist = 1 - sim
    sim = F(1.0) / (F(1.0) * np.sqrt(F(2.0)))
    assert bits(co.cosine_distance([1, 0], [1, 1])) == bits(F(1.0) - sim)
    assert float(co.cosine_distance([1, 0], [1, 1])) == pytest.approx(0.29289323, abs=1e-7)


def test_nan_and_clamp_semantics():
    nan, inf = F("nan"), F("inf")
    # NaN elements at the same position count as identical (|NaN| > 1e-10 is false, :13)
    assert co.vectors_are_identical([1, nan, 3], [1, nan, 3])
    assert co.vectors_are_identical([1, nan, 3], [1, 5, 3])
    assert co.cosine_distance([1, nan, 3], [1, 5, 3]) == 0.0
    # a NaN similarity is clamped by f32::max(-1.0) to -1.0 -> distance 2.0 (:161-164)
    assert co.cosine_distance([inf, 1], [1, 1]) == 2.0
    assert co.cosine_distance([3e19, 3e19], [2e19, 1e19]) == 2.0  # dot and squares overflow: inf/inf
    assert co.cosine_distance([3e19, 3e19], [1, 2]) == 1.0        # finite dot / inf magnitude -> sim 0
    # similarity slightly above 1 from rounding is clamped to 1 -> distance 0, never negative
    rng = np.random.default_rng(0)
    for _ in range(200):
        v = rng.uniform(-1, 1, 768).astype(F)
        d = co.cosine_distance(v, v * F(3.0))
        assert 0.0 <= d <= 2.0
    # 1e-10 threshold: tiny vectors that differ by less are "identical"
    assert co.cosine_distance([1e-12, 2e-12], [3e-12, -1e-12]) == 0.0


def test_sequential_sum_order_matters():
    """The restatement must keep the left-to-right f32 fold: a pairwise sum gives another value."""
    rng = np.random.default_rng(1)
    found = False
    for _ in range(50):
        a, b = rng.uniform(-1, 1, (2, 768)).astype(F)
        seq = co.cosine_distance(a, b)
        assert bits(seq) == bits(no.cosine_distance(a, b)) == bits(no.cosine_distance_loop(a, b))
        pair = F(1.0) - F(np.dot(a, b)) / (np.sqrt(F(np.dot(a, a))) * np.sqrt(F(np.dot(b, b))))
        found |= bits(seq) != bits(F(pair))
    assert found


@pytest.mark.parametrize("dim", [1, 3, 4, 77, 768])
def test_c_matches_numpy_restatement(dim):
    rng = np.random.default_rng(dim)
    rows = rng.uniform(-1, 1, (300, dim)).astype(F)
    rows[7] = 0
    rows[9] = rows[8]
    q = rows[8].copy()
    want = no._distances(rows, q)
    got = np.array([co.cosine_distance(r, q) for r in rows], F)
    assert np.array_equal(bits(got), bits(want))


# ---- interval_by_rank: src/mpi_helpers/load_balance.rs:24-42 ---------------------------------
def test_interval_by_rank_known():
    assert co.interval_by_rank(0, 4, 10) == (0, 3)
    assert co.interval_by_rank(3, 4, 10) == (9, 10)
    assert co.interval_by_rank(3, 4, 9) == (9, 9)
    assert co.interval_by_rank(3, 4, 5) == (6, 5)        # start past the end (the survey's quirk)
    assert co.interval_by_rank(2, 8, 3) == (2, 3)        # size > count -> one row each
    assert co.interval_by_rank(7, 8, 3) == (7, 3)
    assert co.interval_by_rank(0, 1, 0) == (0, 0)
    assert co.interval_by_rank(7, 8, 100_000_000) == (87_500_000, 100_000_000)


def test_interval_by_rank_covers_exactly_once():
    for count in list(range(0, 40)) + [100, 1000, 1027]:
        for size in (1, 2, 3, 4, 5, 7, 8, 16, 41):
            seen = np.zeros(count, int)
            for r in range(size):
                assert co.interval_by_rank(r, size, count) == no.interval_by_rank(r, size, count)
                first, n = no.effective_block(r, size, count)
                seen[first:first + n] += 1
            assert np.all(seen == 1), (count, size)


# ---- top-k: src/mpi_helpers/metrics.rs:16-206 ------------------------------------------------
def test_ties_resolve_by_ascending_index():
    rng = np.random.default_rng(3)
    base = rng.uniform(-1, 1, (8, 16)).astype(F)
    rows = np.concatenate([base, base, base])            # every distance appears three times
    q = rng.uniform(-1, 1, 16).astype(F)
    for size in (1, 2, 3, 5):
        idx, dist = co.parallel_top_k_similarity_search(rows, size, 24, q)
        assert np.all(np.diff(dist) >= 0)
        for a, b in zip(range(0, 24, 3), range(3, 27, 3)):
            grp = idx[a:b]
            assert bits(dist[a:b]).tolist() == [bits(dist[a])] * 3
            assert grp.tolist() == sorted(grp.tolist()) and len({int(g) % 8 for g in grp}) == 1


def test_global_top_k_dedups_and_truncates():
    gi = np.array([5, 1, 9, 5, 2, 1], np.uint64)
    gd = np.array([0.5, 0.1, 0.9, 0.5, 0.1, 0.3], F)
    idx, dist = co.compute_global_top_k(gi, gd, 10)
    assert idx.tolist() == [1, 2, 5, 9] and dist.tolist() == [F(0.1), F(0.1), F(0.5), F(0.9)]
    idx, _ = co.compute_global_top_k(gi, gd, 2)
    assert idx.tolist() == [1, 2]
    assert co.compute_global_top_k([], [], 5)[0].size == 0


def test_self_query_invariant():
    """src/main.rs:111,120,148-154: querying with row 0 puts index 0 first at distance 0.0 and
    calculate_accuracy_metrics returns (1, 1, 1)."""
    rows = co.generate_rows(0x5EED0001, 0, 3000, 768)
    for size in (1, 6):
        idx, dist = co.parallel_top_k_similarity_search(rows, size, 50, rows[0], n_threads=4)
        assert idx[0] == 0 and dist[0] == 0.0 and len(idx) == 50
        assert co.calculate_accuracy_metrics(idx, 0, 50) == (1.0, 1.0, 1.0)
    assert co.calculate_accuracy_metrics([4, 7, 0], 0, 50) == (pytest.approx(1 / 3), 1.0, 1.0)
    assert co.calculate_accuracy_metrics([4, 7], 0, 50) == (0.0, 0.0, 0.0)


def test_more_ranks_than_rows_and_k_larger_than_store():
    rows = co.generate_rows(5, 0, 5, 32)
    q = co.generate_rows(6, 0, 1, 32)[0]
    ref = co.parallel_top_k_similarity_search(rows, 1, 64, q)
    assert len(ref[0]) == 5
    for size in (4, 8, 41):
        got = co.parallel_top_k_similarity_search(rows, size, 64, q)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(bits(got[1]), bits(ref[1]))


def test_threads_do_not_change_the_result():
    rows = co.generate_rows(11, 0, 5000, 64)
    a = co.parallel_top_k_similarity_search(rows, 8, 100, rows[17], n_threads=1)
    b = co.parallel_top_k_similarity_search(rows, 8, 100, rows[17], n_threads=8)
    assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))


# ---- generator --------------------------------------------------------------------------------
def test_generator_matches_numpy_and_is_counter_based():
    a = co.generate_rows(0x5EED0001, 0, 257, 768)
    assert np.array_equal(bits(a), bits(no.generate_rows(0x5EED0001, 0, 257, 768)))
    assert np.array_equal(a[100:], co.generate_rows(0x5EED0001, 100, 157, 768))
    assert a.min() >= -1.0 and a.max() < 1.0 and abs(float(a.mean())) < 0.01
    assert abs(float(a.std()) - 1 / np.sqrt(3)) < 0.01           # uniform(-1,1), src/utils.rs:19


# ---- committed fixtures ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["synthetic_2k_768", "synthetic_ragged_1027_100",
                                  "synthetic_odd_dim_300_77"])
def test_oracle_reproduces_synthetic_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    rows = co.generate_rows(int(g["seed"]), 0, int(g["n"]), int(g["dim"]))
    for qi, q in enumerate(g["queries"]):
        for k in g["ks"]:
            for size in g["sizes"]:
                idx, dist = co.parallel_top_k_similarity_search(rows, int(size), int(k), q, n_threads=2)
                assert np.array_equal(idx, g[f"idx_q{qi}_k{k}_s{size}"])
                assert np.array_equal(bits(dist), g[f"dist_bits_q{qi}_k{k}_s{size}"])


def test_oracle_reproduces_edge_golden():
    g = np.load(os.path.join(GOLDEN, "edge_cases.npz"))
    rows = g["rows"]
    for qi, q in enumerate(g["queries"]):
        all_d = np.array([no.cosine_distance(r, q) for r in rows], F)
        assert np.array_equal(bits(all_d), g[f"all_dist_bits_q{qi}"])
        for k in g["ks"]:
            for size in g["sizes"]:
                idx, dist = co.parallel_top_k_similarity_search(rows, int(size), int(k), q)
                assert np.array_equal(idx, g[f"idx_q{qi}_k{k}_s{size}"])
                assert np.array_equal(bits(dist), g[f"dist_bits_q{qi}_k{k}_s{size}"])


# ---- BSR_DOT (an extension: the reference has cosine only) ---------------------------------------
def test_dot_distance_known_answers_and_order():
    assert co.dot_distance([1, 2, 3], [4, 5, 6]) == F(1.0) - F(32.0)
    assert co.dot_distance([1, 0], [0, 1]) == 1.0
    assert co.dot_distance([1, 2], [1, 2, 3]) == 1.0 and co.dot_distance([], []) == 1.0
    # sequential fold, separately rounded products: same value as the numpy restatement, bit for bit
    rng = np.random.default_rng(5)
    for _ in range(50):
        a, b = rng.uniform(-1, 1, 768).astype(F), rng.uniform(-1, 1, 768).astype(F)
        assert bits(co.dot_distance(a, b)) == bits(no.dot_distance(a, b))
    # on unit vectors it is the cosine distance up to the rounding of the norms
    a /= np.linalg.norm(a); b /= np.linalg.norm(b)
    assert abs(float(co.dot_distance(a, b)) - float(co.cosine_distance(a, b))) < 1e-6
    # total order: numbers ascending, NaN after +inf, ties by index
    rows = np.array([[1, 0], [np.nan, 0], [-3e38, -3e38], [3e38, 3e38], [1, 0], [0, 0]], F)
    idx, dist = co.compute_local_top_k(rows, 0, 1, 6, np.array([3e38, 3e38], F), metric=1)
    assert idx.tolist() == [3, 0, 4, 5, 2, 1]
    assert dist[0] == -np.inf and dist[4] == np.inf and np.isnan(dist[5])


def test_oracle_reproduces_dot_golden():
    g = np.load(os.path.join(GOLDEN, "dot_metric.npz"))
    rows, queries = g["rows"], g["queries"]
    for qi, q in enumerate(queries):
        assert np.array_equal(bits([no.dot_distance(r, q) for r in rows]), g[f"all_dist_bits_q{qi}"])
        for k in g["ks"]:
            for size in g["sizes"]:
                parts = [no.compute_local_top_k(rows, r, int(size), int(k), q, metric=1) for r in range(int(size))]
                gi, gd = co.compute_global_top_k(np.concatenate([p[0] for p in parts]),
                                                 np.concatenate([p[1] for p in parts]), int(k), metric=1)
                assert np.array_equal(gi, g[f"idx_q{qi}_k{k}_s{size}"])
                assert np.array_equal(bits(gd), g[f"dist_bits_q{qi}_k{k}_s{size}"])


def test_scan_generated_equals_materialised_scan():
    """bench.py's full-size parity check walks generator rows block-wise; on a size the plain oracle
    also handles, the block-wise list must be the plain one (every block size, threads, bf16)."""
    n, k = 5000, 50
    rows = co.generate_rows(0x5EED0001, 100, n, 768)
    q = co.generate_rows(0x5EED0001 + 0x1000, 0, 1, 768)[0]
    wi, wd = co.compute_local_top_k(rows, 0, 1, k, q)
    for block, threads in ((n, 1), (777, 1), (1024, 4)):
        gi, gd = co.scan_generated_parallel(0x5EED0001, 100, n, k, q, threads=threads, block_rows=block)
        assert (gi - 100).tolist() == wi.tolist() and np.array_equal(bits(gd), bits(wd))
    wi, wd = co.compute_local_top_k(co.round_bf16(rows), 0, 1, k, q, metric=1)
    gi, gd = co.scan_generated_parallel(0x5EED0001, 100, n, k, q, threads=2, block_rows=999, bf16=True, metric=1)
    assert (gi - 100).tolist() == wi.tolist() and np.array_equal(bits(gd), bits(wd))
