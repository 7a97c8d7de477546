"""CPU: the offline IR metrics of src/metrics.rs:21-141 -- product port (metrics.py), C oracle and
numpy oracle -- against the known answers the reference's own tests hold (src/metrics.rs:172-317).

These are the only known-answer tests the reference has on this side of the code base; inputs are
built the way the reference builds them (mock embeddings uniform(-1,1), DIMENSION = 768,
src/utils.rs:8,14-23) and the expected values are the reference's: 1.0, ~0.611 (+-0.01), 0.25 / 0.5 /
0.75 / 1.0 (+-1e-5), 0.8 (+-0.01)."""
import numpy as np
import pytest

import better_search_rag_rust_b200 as bsr
from oracle import c_oracle as co
from oracle import numpy_oracle as no

IMPLS = {
    "product": (bsr.mean_reciprocal_rank, bsr.recall_at_k, bsr.top_k_overlap),
    "c_oracle": (co.mean_reciprocal_rank, co.recall_at_k, co.top_k_overlap),
    "numpy_oracle": (no.mean_reciprocal_rank, no.recall_at_k, no.top_k_overlap),
}


def mock(n, seed):
    """generate_many_mock_embeddings (src/utils.rs:25-27): n uniform(-1,1) vectors of 768 floats."""
    return [v for v in co.generate_rows(0xABCD00 + seed, 0, n, 768)]


@pytest.mark.parametrize("impl", IMPLS)
def test_mrr_perfect_match(impl):  # src/metrics.rs:172-190
    mrr = IMPLS[impl][0]
    queries = [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]]
    results = [[q] for q in queries]
    assert abs(mrr(queries, results) - 1.0) < 1e-5


@pytest.mark.parametrize("impl", IMPLS)
def test_mrr_mixed_ranks(impl):  # src/metrics.rs:192-221
    mrr = IMPLS[impl][0]
    e = mock(3, 1)
    o = mock(6, 2)
    results = [[e[0], o[0], o[1]], [o[2], e[1], o[3]], [o[4], o[5], e[2]]]
    got = mrr(e, results)
    assert abs(got - 0.611) < 0.01
    assert abs(got - np.float32((1 + 0.5 + np.float32(1) / np.float32(3)) / 3)) < 1e-6


@pytest.mark.parametrize("impl", IMPLS)
def test_recall_at_k(impl):  # src/metrics.rs:223-268
    recall = IMPLS[impl][1]
    e = mock(4, 3)
    results = [
        [e[0]] + mock(5, 4),
        mock(1, 5) + [e[1]] + mock(4, 6),
        mock(2, 7) + [e[2]] + mock(3, 8),
        mock(4, 9) + [e[3]] + mock(1, 10),
    ]
    assert abs(recall(e, results, 1) - 0.25) < 1e-5
    assert abs(recall(e, results, 2) - 0.5) < 1e-5
    assert abs(recall(e, results, 3) - 0.75) < 1e-5
    assert abs(recall(e, results, 5) - 1.0) < 1e-5


@pytest.mark.parametrize("impl", IMPLS)
def test_top_k_overlap(impl):  # src/metrics.rs:270-317
    overlap = IMPLS[impl][2]
    e = mock(10, 11)
    seq = [[e[0], e[1], e[2], e[3], e[4]], [e[0], e[1], e[2], e[3], e[4]]]
    par = [[e[0], e[1], e[5], e[3], e[6]], [e[0], e[1], e[2], e[3], e[4]]]
    assert abs(overlap(seq, par, 5) - 0.8) < 0.01


@pytest.mark.parametrize("impl", IMPLS)
def test_degenerate_inputs(impl):
    mrr, recall, overlap = IMPLS[impl]
    e = mock(2, 12)
    # empty / mismatched outer lists -> 0.0 (src/metrics.rs:26-28, 66-68, 103-108)
    assert mrr([], []) == 0.0 and recall([], [], 3) == 0.0 and overlap([], [], 3) == 0.0
    assert mrr(e, [[e[0]]]) == 0.0 and recall(e, [[e[0]]], 1) == 0.0
    assert overlap([[e[0]]], [[e[0]], [e[1]]], 1) == 0.0
    # a query whose own vector is absent ranks one past the end (:50-52)
    assert abs(mrr([e[0]], [[e[1]]]) - 0.5) < 1e-6
    # a query with no results contributes 0 to the overlap but still counts (:117-119)
    assert abs(overlap([[e[0]], []], [[e[0]], []], 3) - 0.5) < 1e-6
    # k larger than the lists is clamped (:75, :115)
    assert recall([e[0]], [[e[1], e[0]]], 10) == 1.0


def test_vectors_are_identical_rules():  # src/metrics.rs:7-19
    a = np.array([1.0, 2.0, 3.0], np.float32)
    for f in (bsr.vectors_are_identical, co.vectors_are_identical, no.vectors_are_identical):
        assert f(a, a.copy())
        assert not f(a, a[:2])                                   # length (:8-10)
        assert not f(a, a + np.float32(1e-6))
        assert f(np.array([np.nan, 1.0], np.float32), np.array([0.0, 1.0], np.float32))  # NaN > 1e-10 is false
        assert f(np.array([1e-11], np.float32), np.array([0.0], np.float32))             # within 1e-10


def test_agreement_report_layout():
    e = mock(3, 13)
    text = bsr.agreement_report([e[0]], [[e[0], e[1]]], [[e[0], e[2]]], 2)
    assert text.splitlines() == ["Accuracy Metrics:", "  Mean Reciprocal Rank (MRR): 1.0000",
                                 "  Recall@2: 1.0000", "  Top-k Overlap: 0.5000"]
