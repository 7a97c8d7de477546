"""Top-k similarity search -- mirrors src/mpi_helpers/metrics.rs:16-206 of the reference.

Reference                                   here
------------------------------------------  -----------------------------------------------------
compute_local_top_k       (:16-53)          one fused CUDA scan + top-k over the rank's HBM shard
gather_top_k_results      (:56-138)         one all-gather of fixed-size candidate slots
compute_global_top_k      (:141-171)        on-device merge of the gathered lists
parallel_top_k_similarity_search (:174-206) the three above, device-resident end to end

One process per GPU; `world` is a torch.distributed process group (NCCL over NVLink on a GPU box,
gloo in the CPU tests) instead of an rsmpi communicator.  All arithmetic runs in libbsr_b200.so;
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import BsrError
from .load_balance import interval_by_rank
from .vectorstore import GpuVectorstore, ParquetRowSource, _read_parquet_rows, rank_file_paths

ROOT = 0  # src/mpi_helpers/mod.rs
KEY_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)

TopK = List[Tuple[int, float]]


def is_root(rank: int) -> bool:
    return rank == ROOT


# ---------------------------------------------------------------------------------------------
# candidate keys (host side of the format documented in include/bsr.h)
# ---------------------------------------------------------------------------------------------
def candidate_slots(top_k: int) -> int:
    if not 1 <= top_k <= _lib.BSR_MAX_TOP_K:
        raise BsrError(_lib.BSR_ERR_INVALID, f"top_k {top_k} outside [1, {_lib.BSR_MAX_TOP_K}]")
    p = 32
    while p < top_k:
        p *= 2
    return p


def pack_candidates(idx, dist, slots: int) -> np.ndarray:
    """(index, distance) pairs -> sorted-order-preserving uint64 keys, padded with KEY_MAX."""
    idx = np.asarray(idx, dtype=np.uint64).ravel()
    d = np.asarray(dist, dtype=np.float32).ravel() + np.float32(0.0)  # -0.0 -> +0.0
    b = d.view(np.uint32)
    o = np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint64)
    keys = np.full(slots, KEY_MAX, dtype=np.uint64)
    n = min(idx.size, slots)
    keys[:n] = ((o << np.uint64(32)) | (idx & np.uint64(0xFFFFFFFF)))[:n]
    return keys


def unpack_candidates(keys, top_k: int) -> Tuple[np.ndarray, np.ndarray]:
    keys = np.asarray(keys, dtype=np.uint64).ravel()[:top_k]
    keys = keys[keys != KEY_MAX]
    o = (keys >> np.uint64(32)).astype(np.uint32)
    b = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    return (keys & np.uint64(0xFFFFFFFF)).astype(np.uint64), b.view(np.float32)


# ---------------------------------------------------------------------------------------------
# one rank's device side
# ---------------------------------------------------------------------------------------------
class CudaShardEngine:
    """The CUDA work of one rank, over one HBM-resident shard (a GpuVectorstore)."""

    def __init__(self, store: GpuVectorstore):
        self.store = store

    # -- host-buffer path: compute_local_top_k -------------------------------------------------
    def local_top_k(self, queries: np.ndarray, top_k: int):
        """-> idx [nq, top_k] u64 (global), dist [nq, top_k] f32, count [nq] u32."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        idx = np.zeros((nq, top_k), dtype=np.uint64)
        dist = np.zeros((nq, top_k), dtype=np.float32)
        count = np.zeros(nq, dtype=np.uint32)
        n = self.store.get_count()
        if n == 0:
            return idx, dist, count
        if q.shape[1] != self.store.dim:
            # cosine_distance returns 1.0 for every row when the lengths differ
            # (src/metrics.rs:144-146); the stable sort then keeps the first rows of the block.
            m = min(top_k, n)
            idx[:, :m] = self.store.row_base + np.arange(m, dtype=np.uint64)
            dist[:, :m] = 1.0
            count[:] = m
            return idx, dist, count
        _lib.check(_lib.lib().bsr_search(self.store.handle, q.ctypes.data, nq, top_k,
                                         idx.ctypes.data, dist.ctypes.data, count.ctypes.data))
        return idx, dist, count

    # -- device-resident path used by parallel_top_k_similarity_search --------------------------
    def _torch(self):
        import torch
        return torch

    def search_candidates(self, queries, top_k: int, out=None):
        """queries [nq, dim] (numpy or torch) -> int64 tensor [nq, slots] of candidate keys on the
        shard's GPU (bit pattern of the uint64 keys).  `out` reuses a preallocated tensor."""
        torch = self._torch()
        dev = torch.device("cuda", self.store.device)
        slots = candidate_slots(top_k)
        q = torch.as_tensor(np.ascontiguousarray(queries, dtype=np.float32)
                            if not torch.is_tensor(queries) else queries)
        if q.dim() == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        n = self.store.get_count()
        if n == 0 or q.shape[1] != self.store.dim:
            idx, dist, count = self.local_top_k(q.cpu().numpy(), top_k)
            keys = np.stack([pack_candidates(idx[i, :count[i]], dist[i, :count[i]], slots)
                             for i in range(nq)])
            res = torch.from_numpy(keys.view(np.int64)).to(dev)
            if out is not None:
                out.copy_(res)
                return out
            return res
        q = q.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
        if out is None:
            out = torch.empty((nq, slots), dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().bsr_search_device(self.store.handle, q.data_ptr(), nq, top_k,
                                                out.data_ptr(), C.c_void_p(stream)))
        return out

    def merge(self, lists, top_k: int, out=None):
        """lists [n_lists, nq, slots] int64 (any device) -> [nq, slots] on the shard's GPU."""
        torch = self._torch()
        dev = torch.device("cuda", self.store.device)
        lists = lists.to(dev).contiguous()
        n_lists, nq, slots = lists.shape
        assert slots == candidate_slots(top_k)
        if out is None:
            out = torch.empty((nq, slots), dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().bsr_merge_device(self.store.handle, lists.data_ptr(), n_lists, nq,
                                               top_k, out.data_ptr(), C.c_void_p(stream)))
        return out

    def unpack(self, cand, top_k: int):
        """[nq, slots] candidates -> (idx [nq,k] u64, dist [nq,k] f32, count [nq] u32) on host."""
        torch = self._torch()
        dev = torch.device("cuda", self.store.device)
        cand = cand.to(dev).contiguous()
        nq = cand.shape[0]
        idx = torch.empty((nq, top_k), dtype=torch.int64, device=dev)
        dist = torch.empty((nq, top_k), dtype=torch.float32, device=dev)
        count = torch.empty((nq,), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().bsr_unpack_device(self.store.handle, cand.data_ptr(), nq, top_k,
                                                idx.data_ptr(), dist.data_ptr(), count.data_ptr(),
                                                C.c_void_p(stream)))
        return (idx.cpu().numpy().view(np.uint64), dist.cpu().numpy(),
                count.cpu().numpy().view(np.uint32))


class BsrComm:
    """A communicator owned by libbsr_b200.so (NCCL resolved inside the library): the `world` for
    hosts that have no GPU communicator of their own.  Rank 0 calls `BsrComm.unique_id()` and ships
    the 128 bytes to every rank over any host channel (the reference would MPI_Bcast them)."""

    def __init__(self, unique_id: bytes, n_ranks: int, rank: int, device: int):
        if len(unique_id) != _lib.BSR_COMM_ID_BYTES:
            raise ValueError("unique_id must be 128 bytes")
        self.n_ranks, self.rank, self.device = n_ranks, rank, device
        self._id = (C.c_ubyte * _lib.BSR_COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().bsr_comm_create(self._id, n_ranks, rank, device, C.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_ubyte * _lib.BSR_COMM_ID_BYTES)()
        _lib.check(_lib.lib().bsr_comm_unique_id(buf))
        return bytes(buf)

    def search(self, engine: "CudaShardEngine", queries, top_k: int):
        """bsr_search_parallel: collective; (idx, dist, count) on rank 0, None elsewhere."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        if q.shape[1] != engine.store.dim:
            raise BsrError(_lib.BSR_ERR_INVALID, "query length differs from the store's dimension")
        if self.rank == 0:
            idx = np.zeros((nq, top_k), dtype=np.uint64)
            dist = np.zeros((nq, top_k), dtype=np.float32)
            count = np.zeros(nq, dtype=np.uint32)
            args = (idx.ctypes.data, dist.ctypes.data, count.ctypes.data)
        else:
            idx = dist = count = None
            args = (None, None, None)
        _lib.check(_lib.lib().bsr_search_parallel(engine.store.handle, self._h, q.ctypes.data, nq, top_k, *args))
        return (idx, dist, count) if self.rank == 0 else None

    def search_device(self, engine: "CudaShardEngine", queries, top_k: int, out=None):
        """bsr_search_parallel_device: the same collective, device-resident and asynchronous on the
        current torch stream.  `queries` is a CUDA tensor [nq, dim]; returns the int64 tensor
        [nq, slots] of merged candidate keys (meaningful on rank 0)."""
        import torch
        dev = torch.device("cuda", engine.store.device)
        q = queries.to(device=dev, dtype=torch.float32).contiguous()
        if q.dim() == 1:
            q = q.reshape(1, -1)
        if q.shape[1] != engine.store.dim:
            raise BsrError(_lib.BSR_ERR_INVALID, "query length differs from the store's dimension")
        if out is None:
            out = torch.empty((q.shape[0], candidate_slots(top_k)), dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().bsr_search_parallel_device(engine.store.handle, self._h, q.data_ptr(), q.shape[0],
                                                         top_k, out.data_ptr(), C.c_void_p(stream)))
        return out

    def uses_mailbox(self) -> bool:
        """True when single-query top_k <= 128 searches exchange candidates through NVLink peer stores
        fused into the scan kernel instead of ncclAllGather + a merge launch."""
        yes = C.c_int()
        _lib.check(_lib.lib().bsr_comm_uses_mailbox(self._h, C.byref(yes)))
        return bool(yes.value)

    def close(self) -> None:
        if self._h:
            _lib.lib().bsr_comm_destroy(self._h)
            self._h = C.c_void_p()


# ---------------------------------------------------------------------------------------------
# resident shards: the reference re-reads global.parquet inside every compute_local_top_k call
# (metrics.rs:23); here a (dir, rank, size) block is decoded once and stays in HBM.
# ---------------------------------------------------------------------------------------------
_RESIDENT: Dict[tuple, Tuple[tuple, CudaShardEngine]] = {}


def _default_device(rank: int) -> int:
    local = os.environ.get("LOCAL_RANK")
    if local is not None:
        return int(local)
    n = C.c_int()
    _lib.check(_lib.lib().bsr_device_count(C.byref(n)))
    if n.value <= 0:
        raise BsrError(_lib.BSR_ERR_CUDA, "no CUDA device is visible; the search path has no CPU fallback")
    return rank % n.value


def open_shard(vstore_dir, rank: int, size: int, *, device: Optional[int] = None,
               source: str = "global", n_files: Optional[int] = None, dtype: str = "f32") -> CudaShardEngine:
    """Rank `rank` of `size`'s row block of the store under `vstore_dir`, resident in HBM.

    get_global_vstore + get_count + interval_by_rank + get_many of metrics.rs:23-33, done once; the
    shard is reloaded when a file changes.  Only the Parquet row groups that overlap the rank's
    block are decoded (the footers give the row counts), one at a time, and uploaded through the
    library's pinned double-buffered pipeline -- host memory is bounded by one row group where the
    reference holds the whole store on every rank.

    source = "global": <dir>/global.parquet (what the reference searches).
    source = "ranks" : the ingest output <dir>/rank_0.parquet ... rank_{n_files-1}.parquet taken in
      merge order (tasks.rs:186-197), i.e. the same rows global.parquet would hold, without needing
      the merged file to exist."""
    d = os.fspath(vstore_dir)
    if source == "global":
        paths = [os.path.join(d, "global.parquet")]
        if not os.path.exists(paths[0]):
            _read_parquet_rows(paths[0])  # read_parquet creates a missing file, empty (polars.rs:64-73)
    elif source == "ranks":
        paths = rank_file_paths(d, n_files)
        if not paths:
            raise FileNotFoundError(f"no rank_*.parquet under {d}")
    else:
        raise ValueError("source must be 'global' or 'ranks'")
    device = _default_device(rank) if device is None else device
    key = (tuple(os.path.abspath(p) for p in paths), rank, size, device, dtype)

    def stamp_of():
        return tuple((os.stat(p).st_mtime_ns, os.stat(p).st_size) for p in paths)

    stamp = stamp_of()
    hit = _RESIDENT.get(key)
    if hit and hit[0] == stamp:
        return hit[1]
    src = ParquetRowSource(paths)
    count = src.count
    iv = interval_by_rank(rank, size, count)
    first, n = iv.rows(count)
    store = None
    for block in src.blocks(first, n):
        if store is None:
            store = GpuVectorstore(None, True, device=device, row_base=min(iv.start_index, 0xFFFFFFFE),
                                   dim=block.shape[1], capacity_rows=n, dtype=dtype)
        store.append_many(block)
    if store is None:  # empty block (more ranks than rows, or an empty store)
        store = GpuVectorstore(None, True, device=device, row_base=min(iv.start_index, 0xFFFFFFFE), dtype=dtype)
    eng = CudaShardEngine(store)
    if hit:
        hit[1].store.close()
    _RESIDENT[key] = (stamp_of(), eng)
    return eng


def release_shards() -> None:
    for _, eng in _RESIDENT.values():
        eng.store.close()
    _RESIDENT.clear()


def _pairs(idx: np.ndarray, dist: np.ndarray, count: int) -> TopK:
    return [(int(idx[j]), float(dist[j])) for j in range(int(count))]


# ---------------------------------------------------------------------------------------------
# the reference's functions
# ---------------------------------------------------------------------------------------------
def compute_local_top_k(vstore_dir, rank: int, size: int, top_k: int,
                        target_vector: Sequence[float], *, device: Optional[int] = None,
                        engine=None) -> TopK:
    """metrics.rs:16-53 -> [(global_index, distance)] ascending, at most top_k long."""
    eng = engine or open_shard(vstore_dir, rank, size, device=device)
    idx, dist, count = eng.local_top_k(np.asarray(target_vector, dtype=np.float32), top_k)
    return _pairs(idx[0], dist[0], count[0])


def _dist():
    import torch.distributed as dist
    return dist


def _all_gather(world, t):
    """All-gather `t` over `world`, staging through the host when the backend is gloo."""
    import torch
    dist = _dist()
    size = dist.get_world_size(world)
    backend = dist.get_backend(world)
    src = t if backend == "nccl" or not t.is_cuda else t.cpu()
    src = src.contiguous()
    flat = torch.empty(size * src.numel(), dtype=src.dtype, device=src.device)
    dist.all_gather_into_tensor(flat, src.reshape(-1), group=world)
    return flat.reshape((size,) + tuple(src.shape))


def gather_top_k_results(world, rank: int, local_top_k: TopK):
    """metrics.rs:56-138 -> (global_indices, global_distances): on root the rank-ordered
    concatenation of every rank's list (displacements, :86-89), empty lists elsewhere.

    The reference's allgather of counts + per-rank send/recv pairs become one all-gather of
    BSR_MAX_TOP_K-independent fixed slots: [count | keys...]."""
    import torch
    dist = _dist()
    size = dist.get_world_size(world) if (dist.is_available() and dist.is_initialized()) else 1
    idx = np.asarray([p[0] for p in local_top_k], dtype=np.uint64)
    dst = np.asarray([p[1] for p in local_top_k], dtype=np.float32)
    if size == 1:
        return (idx.tolist(), [float(x) for x in dst]) if is_root(rank) else ([], [])
    # every rank needs the same slot count: agree on the longest list first (:68)
    n_local = torch.tensor([idx.size], dtype=torch.int64)
    backend = dist.get_backend(world)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    counts = _all_gather(world, n_local.to(dev)).cpu().numpy().ravel()
    slots = int(max(1, counts.max()))
    buf = np.zeros(2 * slots, dtype=np.int64)
    buf[:idx.size] = idx.view(np.int64)
    buf[slots:slots + dst.size] = dst.view(np.int32).astype(np.int64)
    gathered = _all_gather(world, torch.from_numpy(buf).to(dev)).cpu().numpy()
    if not is_root(rank):
        return [], []
    g_idx: List[int] = []
    g_dist: List[float] = []
    for r in range(size):
        c = int(counts[r])
        g_idx.extend(gathered[r, :c].view(np.uint64).tolist())
        g_dist.extend(float(x) for x in gathered[r, slots:slots + c].astype(np.int32).view(np.float32))
    return g_idx, g_dist


_MERGE_ENGINES: Dict[int, CudaShardEngine] = {}


def _merge_engine(device: int) -> CudaShardEngine:
    eng = _MERGE_ENGINES.get(device)
    if eng is None:
        eng = CudaShardEngine(GpuVectorstore(None, True, device=device, dim=4))
        _MERGE_ENGINES[device] = eng
    return eng


def compute_global_top_k(global_indices: Sequence[int], global_distances: Sequence[float],
                         top_k: int, *, device: Optional[int] = None, engine=None) -> TopK:
    """metrics.rs:141-171: sort the gathered candidates on the GPU and keep the first top_k.

    One bsr_merge_host call: H2D of the pairs, key packing + per-list sort, merge tree, unpack,
    D2H.  Indices are unique (disjoint row blocks), which makes the reference's HashSet pass
    (:156-168) a no-op."""
    gi = np.ascontiguousarray(global_indices, dtype=np.uint64).ravel()
    gd = np.ascontiguousarray(global_distances, dtype=np.float32).ravel()
    if gi.size != gd.size:
        raise ValueError("global_indices and global_distances differ in length")
    if gi.size == 0:
        return []
    eng = engine or _merge_engine(_default_device(0) if device is None else device)
    if not hasattr(eng, "store"):  # test double
        cand = eng.merge(_torch_keys(gi, gd, top_k), top_k)
        idx, dist, count = eng.unpack(cand, top_k)
        return _pairs(idx[0], dist[0], count[0])
    idx = np.zeros(top_k, dtype=np.uint64)
    dist = np.zeros(top_k, dtype=np.float32)
    count = C.c_uint32()
    _lib.check(_lib.lib().bsr_merge_host(eng.store.handle, gi.ctypes.data, gd.ctypes.data, gi.size,
                                         top_k, idx.ctypes.data, dist.ctypes.data, C.byref(count)))
    return _pairs(idx, dist, count.value)


def _torch_keys(gi, gd, top_k):
    import torch
    slots = candidate_slots(top_k)
    n_lists = -(-gi.size // slots)
    keys = np.full(n_lists * slots, KEY_MAX, dtype=np.uint64)
    keys[:gi.size] = pack_candidates(gi, gd, gi.size)
    keys = np.sort(keys.reshape(n_lists, slots), axis=1)
    return torch.from_numpy(keys.view(np.int64)).reshape(n_lists, 1, slots)


def parallel_top_k_similarity_search(world, rank: int, size: int, vstore_dir, top_k: int,
                                     target_vector: Sequence[float], *,
                                     device: Optional[int] = None, engine=None,
                                     verbose: bool = False) -> Optional[TopK]:
    """metrics.rs:174-206 -> Some(top_k list) on root, None on the other ranks.

    Device-resident pipeline: fused scan+select on the rank's shard -> all-gather of
    `slots` candidates per rank (NCCL over NVLink) -> on-device merge on root -> unpack."""
    if verbose:
        print(f"[Rank {rank}] Calculating top-{top_k} distances")
    if isinstance(world, BsrComm):
        # the library's own communicator: scan, ncclAllGather, merge and unpack in one C-ABI call
        eng = engine or open_shard(vstore_dir, rank, size, device=device)
        res = world.search(eng, np.asarray(target_vector, dtype=np.float32), top_k)
        return _pairs(res[0][0], res[1][0], res[2][0]) if res is not None else None
    slots = candidate_slots(top_k)
    target = np.asarray(target_vector, dtype=np.float32).reshape(1, -1)
    try:
        eng = engine or open_shard(vstore_dir, rank, size, device=device)
        cand = eng.search_candidates(target, top_k)
    except (OSError, ValueError) as err:
        # metrics.rs:185-191: a failed local scan contributes an empty list.  Only store/decode
        # errors (the reference's PolarsError) degrade; a missing library or CUDA failure raises.
        print(f"[Rank {rank}] Error in similarity search: {err!r}")
        import torch
        cand = torch.from_numpy(np.full((1, slots), KEY_MAX, dtype=np.uint64).view(np.int64))
        eng = engine or _merge_engine(_default_device(rank) if device is None else device)
    if size > 1:
        gathered = _all_gather(world, cand)  # [size, 1, slots], rank-ordered (:86-89)
        if not is_root(rank):
            return None
        cand = eng.merge(gathered, top_k)
    elif not is_root(rank):
        return None
    idx, dist, count = eng.unpack(cand, top_k)
    return _pairs(idx[0], dist[0], count[0])
