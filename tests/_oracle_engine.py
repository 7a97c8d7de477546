"""TEST DOUBLE, never shipped: a shard engine with the product's CudaShardEngine interface whose
device work is done by the CPU oracle, so the host orchestration (row blocks, all-gather layout,
root-only result) can run under gloo on a machine without a GPU."""
import numpy as np
import torch

from better_search_rag_rust_b200 import search as S
from oracle import c_oracle as co


class OracleShardEngine:
    def __init__(self, global_rows: np.ndarray, rank: int, size: int):
        self.rows, self.rank, self.size = np.ascontiguousarray(global_rows, np.float32), rank, size

    def local_top_k(self, queries, top_k):
        q = np.asarray(queries, np.float32).reshape(-1, self.rows.shape[1])
        idx = np.zeros((len(q), top_k), np.uint64)
        dist = np.zeros((len(q), top_k), np.float32)
        count = np.zeros(len(q), np.uint32)
        for i, qq in enumerate(q):
            gi, gd = co.compute_local_top_k(self.rows, self.rank, self.size, top_k, qq)
            idx[i, :len(gi)], dist[i, :len(gi)], count[i] = gi, gd, len(gi)
        return idx, dist, count

    def search_candidates(self, queries, top_k, out=None):
        q = queries.cpu().numpy() if torch.is_tensor(queries) else np.asarray(queries, np.float32)
        idx, dist, count = self.local_top_k(q, top_k)
        slots = S.candidate_slots(top_k)
        keys = np.stack([S.pack_candidates(idx[i, :count[i]], dist[i, :count[i]], slots)
                         for i in range(len(count))])
        return torch.from_numpy(keys.view(np.int64))

    def merge(self, lists, top_k, out=None):
        keys = lists.cpu().numpy().view(np.uint64)       # [n_lists, nq, slots]
        n_lists, nq, slots = keys.shape
        merged = np.sort(keys.transpose(1, 0, 2).reshape(nq, -1), axis=1)[:, :slots]
        return torch.from_numpy(np.ascontiguousarray(merged).view(np.int64))

    def unpack(self, cand, top_k):
        keys = cand.cpu().numpy().view(np.uint64)
        nq = keys.shape[0]
        idx = np.zeros((nq, top_k), np.uint64)
        dist = np.zeros((nq, top_k), np.float32)
        count = np.zeros(nq, np.uint32)
        for i in range(nq):
            gi, gd = S.unpack_candidates(keys[i], top_k)
            idx[i, :len(gi)], dist[i, :len(gi)], count[i] = gi, gd, len(gi)
        return idx, dist, count
