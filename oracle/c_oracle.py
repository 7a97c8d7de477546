"""ctypes binding of oracle/bsr_oracle.c (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Every function here restates one reference function; citations are relative to
/root/reference/ and repeated from the C source.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "libbsr_oracle.so")

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "bsr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.bsr_oracle_vectors_are_identical.restype = C.c_int
        L.bsr_oracle_vectors_are_identical.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64]
        L.bsr_oracle_cosine_distance.restype = C.c_float
        L.bsr_oracle_cosine_distance.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64]
        L.bsr_oracle_interval_by_rank.restype = None
        L.bsr_oracle_interval_by_rank.argtypes = [
            C.c_int32, C.c_int32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.bsr_oracle_interval_count.restype = C.c_uint64
        L.bsr_oracle_interval_count.argtypes = [C.c_uint64, C.c_uint64]
        L.bsr_oracle_compute_local_top_k.restype = C.c_uint64
        L.bsr_oracle_compute_local_top_k.argtypes = [
            _f32p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int32, C.c_uint64, _f32p, C.c_uint32,
            _u64p, _f32p]
        L.bsr_oracle_compute_global_top_k.restype = C.c_uint64
        L.bsr_oracle_compute_global_top_k.argtypes = [
            _u64p, _f32p, C.c_uint64, C.c_uint64, _u64p, _f32p]
        L.bsr_oracle_parallel_top_k_similarity_search.restype = C.c_uint64
        L.bsr_oracle_parallel_top_k_similarity_search.argtypes = [
            _f32p, C.c_uint64, C.c_uint32, C.c_int32, C.c_uint64, _f32p, C.c_uint32, C.c_int32,
            _u64p, _f32p]
        L.bsr_oracle_calculate_accuracy_metrics.restype = None
        L.bsr_oracle_calculate_accuracy_metrics.argtypes = [
            _u64p, C.c_uint64, C.c_uint64, C.c_uint64,
            C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.bsr_oracle_generate_rows.restype = None
        L.bsr_oracle_generate_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _f32p]
        L.bsr_oracle_dot_distance.restype = C.c_float
        L.bsr_oracle_dot_distance.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64]
        L.bsr_oracle_compute_local_top_k_metric.restype = C.c_uint64
        L.bsr_oracle_compute_local_top_k_metric.argtypes = [
            _f32p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int32, C.c_uint64, _f32p, C.c_uint32,
            C.c_int, _u64p, _f32p]
        L.bsr_oracle_compute_global_top_k_metric.restype = C.c_uint64
        L.bsr_oracle_compute_global_top_k_metric.argtypes = [
            _u64p, _f32p, C.c_uint64, C.c_uint64, C.c_int, _u64p, _f32p]
        L.bsr_oracle_scan_generated.restype = C.c_uint64
        L.bsr_oracle_scan_generated.argtypes = [
            C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_uint64, _f32p,
            C.c_uint32, _u64p, _f32p]
        L.bsr_oracle_mean_reciprocal_rank.restype = C.c_float
        L.bsr_oracle_mean_reciprocal_rank.argtypes = [_f32p, C.c_uint64, _f32p, _u64p, C.c_uint64, C.c_uint32]
        L.bsr_oracle_recall_at_k.restype = C.c_float
        L.bsr_oracle_recall_at_k.argtypes = [_f32p, C.c_uint64, _f32p, _u64p, C.c_uint64, C.c_uint32,
                                             C.c_uint64]
        L.bsr_oracle_top_k_overlap.restype = C.c_float
        L.bsr_oracle_top_k_overlap.argtypes = [_f32p, _u64p, C.c_uint64, _f32p, _u64p, C.c_uint64,
                                               C.c_uint32, C.c_uint64]
        L.bsr_oracle_round_bf16.restype = None
        L.bsr_oracle_round_bf16.argtypes = [_f32p, C.c_uint64, _f32p]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def vectors_are_identical(a, b) -> bool:
    """src/metrics.rs:7-19"""
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return bool(lib().bsr_oracle_vectors_are_identical(a, a.size, b, b.size))


def cosine_distance(a, b) -> np.float32:
    """src/metrics.rs:143-165"""
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return np.float32(lib().bsr_oracle_cosine_distance(a, a.size, b, b.size))


def interval_by_rank(rank: int, size: int, count: int) -> tuple[int, int]:
    """src/mpi_helpers/load_balance.rs:24-42 -> (start_index, end_index), start unclamped."""
    s, e = C.c_uint64(), C.c_uint64()
    lib().bsr_oracle_interval_by_rank(rank, size, count, C.byref(s), C.byref(e))
    return int(s.value), int(e.value)


def dot_distance(a, b) -> np.float32:
    """BSR_DOT (include/bsr.h): 1 - the sequential dot product of src/metrics.rs:153."""
    a, b = _f32(a).ravel(), _f32(b).ravel()
    return np.float32(lib().bsr_oracle_dot_distance(a, a.size, b, b.size))


def compute_local_top_k(rows, rank: int, size: int, top_k: int, target, metric: int = 0):
    """src/mpi_helpers/metrics.rs:16-53 over the whole row-major store `rows`."""
    rows = _f32(rows)
    assert rows.ndim == 2
    target = _f32(target).ravel()
    idx = np.zeros(max(top_k, 1), dtype=np.uint64)
    dist = np.zeros(max(top_k, 1), dtype=np.float32)
    n = lib().bsr_oracle_compute_local_top_k_metric(
        rows, rows.shape[0], rows.shape[1], rank, size, top_k, target, target.size, metric, idx, dist)
    return idx[:n].copy(), dist[:n].copy()


def compute_global_top_k(global_idx, global_dist, top_k: int, metric: int = 0):
    """src/mpi_helpers/metrics.rs:141-171"""
    gi = np.ascontiguousarray(global_idx, dtype=np.uint64).ravel()
    gd = _f32(global_dist).ravel()
    assert gi.size == gd.size
    idx = np.zeros(max(top_k, 1), dtype=np.uint64)
    dist = np.zeros(max(top_k, 1), dtype=np.float32)
    if gi.size == 0:
        return idx[:0], dist[:0]
    n = lib().bsr_oracle_compute_global_top_k_metric(gi, gd, gi.size, top_k, metric, idx, dist)
    return idx[:n].copy(), dist[:n].copy()


def scan_generated(seed: int, first_row: int, n_rows: int, top_k: int, target, *, dim: int = 768,
                   bf16: bool = False, metric: int = 0):
    """compute_local_top_k (metrics.rs:36-50) over generator rows [first_row, first_row + n_rows),
    produced on the fly (bounded memory); indices are generator row numbers."""
    target = _f32(target).ravel()
    idx = np.zeros(max(top_k, 1), dtype=np.uint64)
    dist = np.zeros(max(top_k, 1), dtype=np.float32)
    n = lib().bsr_oracle_scan_generated(seed, first_row, n_rows, dim, int(bf16), metric, top_k, target,
                                        target.size, idx, dist)
    return idx[:n].copy(), dist[:n].copy()


def scan_generated_parallel(seed: int, first_row: int, n_rows: int, top_k: int, target, *,
                            threads: int = 1, dim: int = 768, bf16: bool = False, metric: int = 0,
                            block_rows: int = 65536):
    """The same list as scan_generated, computed block-wise on `threads` OS threads: every block's
    stable-sorted first top_k, concatenated in row order, go through compute_global_top_k (stable
    sort + first top_k) -- identical to one stable sort over all rows."""
    from concurrent.futures import ThreadPoolExecutor
    blocks = [(lo, min(block_rows, n_rows - lo)) for lo in range(0, n_rows, block_rows)]
    if not blocks:
        return np.zeros(0, np.uint64), np.zeros(0, np.float32)

    def one(b):
        return scan_generated(seed, first_row + b[0], b[1], top_k, target, dim=dim, bf16=bf16, metric=metric)

    if threads <= 1:
        parts = [one(b) for b in blocks]
    else:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(one, blocks))
    gi = np.concatenate([p[0] for p in parts])
    gd = np.concatenate([p[1] for p in parts])
    return compute_global_top_k(gi, gd, top_k, metric)


def _flatten_lists(lists, dim):
    counts = np.asarray([len(x) for x in lists], dtype=np.uint64)
    flat = [np.asarray(v, dtype=np.float32).reshape(-1) for x in lists for v in x]
    for v in flat:
        assert v.size == dim, "every vector must have the same length"
    data = np.concatenate(flat) if flat else np.zeros(1, np.float32)
    return np.ascontiguousarray(data, dtype=np.float32), (counts if counts.size else np.zeros(1, np.uint64))


def mean_reciprocal_rank(queries, results) -> float:
    """src/metrics.rs:21-58"""
    q = _f32(queries)
    dim = q.shape[1] if q.ndim == 2 and q.size else 1
    r, nr = _flatten_lists(results, dim)
    return float(lib().bsr_oracle_mean_reciprocal_rank(q.reshape(-1) if q.size else np.zeros(1, np.float32),
                                                       len(queries), r, nr, len(results), dim))


def recall_at_k(queries, results, k: int) -> float:
    """src/metrics.rs:60-96"""
    q = _f32(queries)
    dim = q.shape[1] if q.ndim == 2 and q.size else 1
    r, nr = _flatten_lists(results, dim)
    return float(lib().bsr_oracle_recall_at_k(q.reshape(-1) if q.size else np.zeros(1, np.float32),
                                              len(queries), r, nr, len(results), dim, k))


def top_k_overlap(sequential_results, parallel_results, k: int) -> float:
    """src/metrics.rs:98-141"""
    dim = 1
    for lists in (sequential_results, parallel_results):
        for x in lists:
            for v in x:
                dim = np.asarray(v).size
                break
    s, ns = _flatten_lists(sequential_results, dim)
    p, npar = _flatten_lists(parallel_results, dim)
    return float(lib().bsr_oracle_top_k_overlap(s, ns, len(sequential_results), p, npar,
                                                len(parallel_results), dim, k))


def parallel_top_k_similarity_search(rows, size: int, top_k: int, target, n_threads: int = 1):
    """src/mpi_helpers/metrics.rs:174-206 with `size` ranks run as OS threads; root's result."""
    rows = _f32(rows)
    assert rows.ndim == 2
    target = _f32(target).ravel()
    idx = np.zeros(max(top_k, 1), dtype=np.uint64)
    dist = np.zeros(max(top_k, 1), dtype=np.float32)
    n = lib().bsr_oracle_parallel_top_k_similarity_search(
        rows, rows.shape[0], rows.shape[1], size, top_k, target, target.size, n_threads, idx, dist)
    return idx[:n].copy(), dist[:n].copy()


def calculate_accuracy_metrics(idx, query_idx: int, top_k: int):
    """src/mpi_helpers/metrics.rs:217-249 -> (mrr, recall, overlap)"""
    idx = np.ascontiguousarray(idx, dtype=np.uint64).ravel()
    m, r, o = C.c_float(), C.c_float(), C.c_float()
    lib().bsr_oracle_calculate_accuracy_metrics(
        idx if idx.size else np.zeros(1, np.uint64), idx.size, query_idx, top_k,
        C.byref(m), C.byref(r), C.byref(o))
    return float(m.value), float(r.value), float(o.value)


def generate_rows(seed: int, first_row: int, n_rows: int, dim: int = 768) -> np.ndarray:
    """Synthetic uniform(-1,1) rows (distribution of src/utils.rs:14-23), counter-based."""
    out = np.empty((n_rows, dim), dtype=np.float32)
    if n_rows:
        lib().bsr_oracle_generate_rows(seed, first_row, n_rows, dim, out)
    return out


def round_bf16(a) -> np.ndarray:
    a = _f32(a)
    out = np.empty_like(a)
    if a.size:
        lib().bsr_oracle_round_bf16(a.ravel(), a.size, out.reshape(-1))
    return out
