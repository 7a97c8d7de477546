"""HBM-resident vector store -- mirrors src/vectorstore/polars.rs (PolarsVectorstore) and
src/mpi_helpers/vectorstore.rs (path helpers) of the reference.

The reference keeps a polars DataFrame (one column ``embeddings: List(Float32)``) on the host and
re-decodes ``global.parquet`` on every rank for every search.  Here the rows live in one GPU's HBM
behind a ``bsr_store`` handle (include/bsr.h); Parquet is only touched by new()/reload()/persist().
Method names, argument meaning and error behaviour follow the reference.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import BsrError

EMBEDDINGS_COLUMN = "embeddings"  # polars.rs:17-27


class NoDataError(LookupError):
    """PolarsError::NoData (polars.rs:167,176)."""


@dataclass
class SliceArgs:
    """polars.rs:12-15 (offset is an i32 there; negative counts from the end, as in polars)."""

    offset: int
    length: int


def _as_rows(vectors, dim: Optional[int]) -> np.ndarray:
    rows = np.ascontiguousarray(vectors, dtype=np.float32)
    if rows.ndim == 1:
        rows = rows.reshape(1, -1)
    if rows.ndim != 2:
        raise ValueError("vectors must be a 2-D array / list of equal-length rows")
    if dim is not None and rows.shape[1] != dim:
        raise ValueError(f"row length {rows.shape[1]} does not match the store's dimension {dim}")
    return rows


def _read_parquet_rows(path: str) -> np.ndarray:
    """read_parquet (polars.rs:50-77): decode the `embeddings` list column to [n, dim] f32.
    A missing file is created empty as the reference does (:64-73)."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    parent = os.path.dirname(path)
    if parent and not os.path.exists(parent):
        os.makedirs(parent, exist_ok=True)
    if not os.path.exists(path):
        _write_parquet_rows(path, np.zeros((0, 0), dtype=np.float32))
        return np.zeros((0, 0), dtype=np.float32)
    table = pq.read_table(path, columns=[EMBEDDINGS_COLUMN])
    return _decode_embeddings(table.column(EMBEDDINGS_COLUMN), path)


def _decode_embeddings(col, where: str) -> np.ndarray:
    """One Arrow `embeddings` array (list / large_list / fixed_size_list of float) -> [n, dim] f32."""
    import pyarrow as pa

    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks()
        if isinstance(col, pa.ChunkedArray):
            col = col.chunk(0) if col.num_chunks else pa.array([], type=pa.list_(pa.float32()))
    if len(col) == 0:
        return np.zeros((0, 0), dtype=np.float32)
    if col.null_count:
        col = col.drop_null()  # get_many's filter_map drops null rows (polars.rs:142)
    if pa.types.is_fixed_size_list(col.type):
        dim = col.type.list_size
    else:
        lens = np.diff(np.asarray(col.offsets))
        if len(lens) and not np.all(lens == lens[0]):
            raise ValueError(f"{where}: embeddings rows have different lengths")
        dim = int(lens[0]) if len(lens) else 0
    values = np.asarray(col.flatten().cast(pa.float32()).to_numpy(zero_copy_only=False), dtype=np.float32)
    return np.ascontiguousarray(values.reshape(len(col), dim))


class ParquetRowSource:
    """Row-addressable view of one or several Parquet files holding the `embeddings` column, decoded
    one ROW GROUP at a time.

    The reference decodes the whole of global.parquet on every rank for every search
    (src/mpi_helpers/metrics.rs:23) -- 307 GB of host memory per rank at 100 M x 768.  Here a rank
    reads the footers (row counts per row group, no data) and then only the row groups that overlap
    its row block; host memory is bounded by one row group.

    Several files are the concatenation in list order: for `rank_{r}.parquet` files that is the order
    merge_vector_stores appends them to the global store (src/mpi_helpers/tasks.rs:186-197), so row g
    of the concatenation is row g of the global.parquet the reference would have written."""

    def __init__(self, paths: Sequence[str]):
        import pyarrow.parquet as pq

        self.paths = [os.fspath(p) for p in paths]
        self._groups = []  # (file index, row group, first global row, rows)
        self._files = []
        self.has_nulls = False
        at = 0
        for fi, path in enumerate(self.paths):
            pf = pq.ParquetFile(path)
            self._files.append(pf)
            md = pf.metadata
            if EMBEDDINGS_COLUMN not in pf.schema_arrow.names and md.num_rows:
                raise ValueError(f"{path}: no `{EMBEDDINGS_COLUMN}` column")
            for rg in range(md.num_row_groups):
                g = md.row_group(rg)
                for ci in range(g.num_columns):
                    st = g.column(ci).statistics
                    if g.column(ci).path_in_schema.split(".")[0] == EMBEDDINGS_COLUMN and st is not None \
                            and st.has_null_count and st.null_count:
                        self.has_nulls = True
                self._groups.append((fi, rg, at, g.num_rows))
                at += g.num_rows
        self.count = at

    def blocks(self, first: int, n: int):
        """Yield [m, dim] f32 arrays covering rows [first, first + n) in order."""
        if self.has_nulls:
            # null rows are dropped by get_many (polars.rs:142), which shifts every later index: the
            # row-group arithmetic above does not hold, fall back to the reference's whole-file decode
            rows = np.concatenate([r for r in (_read_parquet_rows(p) for p in self.paths) if r.shape[0]] or
                                  [np.zeros((0, 0), np.float32)])
            if n:
                yield rows[first:first + n]
            return
        end = first + n
        for fi, rg, at, m in self._groups:
            if at + m <= first or at >= end or m == 0:
                continue
            tbl = self._files[fi].read_row_group(rg, columns=[EMBEDDINGS_COLUMN])
            rows = _decode_embeddings(tbl.column(EMBEDDINGS_COLUMN), self.paths[fi])
            if rows.shape[0] != m:  # nulls the footer statistics did not announce
                raise ValueError(f"{self.paths[fi]}: row group {rg} holds null embeddings; "
                                 "re-write the file without nulls or load it with get_many")
            lo, hi = max(first, at) - at, min(end, at + m) - at
            yield rows[lo:hi]


def rank_file_paths(vstore_dir, n_files: Optional[int] = None) -> List[str]:
    """<dir>/rank_0.parquet, rank_1.parquet, ... (get_local_vstore, src/mpi_helpers/vectorstore.rs:5-13):
    `n_files` of them, or as many consecutive ones as exist."""
    d = os.fspath(vstore_dir)
    if n_files is not None:
        return [os.path.join(d, f"rank_{r}.parquet") for r in range(n_files)]
    out = []
    while os.path.exists(os.path.join(d, f"rank_{len(out)}.parquet")):
        out.append(os.path.join(d, f"rank_{len(out)}.parquet"))
    return out


def _write_parquet_rows(path: str, rows: np.ndarray, row_group_size: int = 512 * 512) -> None:
    """ParquetWriter with the options of polars.rs:210-216 (polars' defaults: zstd, row groups of
    512 x 512 rows)."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    n, dim = rows.shape if rows.ndim == 2 else (0, 0)
    flat = pa.array(np.ascontiguousarray(rows, dtype=np.float32).reshape(-1), type=pa.float32())
    offsets = pa.array(np.arange(n + 1, dtype=np.int64) * dim, type=pa.int64())
    col = pa.LargeListArray.from_arrays(offsets, flat)  # polars' List is Arrow large_list
    pq.write_table(pa.table({EMBEDDINGS_COLUMN: col}), path, compression="zstd",
                   row_group_size=max(1, int(row_group_size)))


class GpuVectorstore:
    """PolarsVectorstore (polars.rs:7-10,79-247) with the rows resident in one GPU's HBM.

    Extra keyword arguments (no reference counterpart): ``device`` -- CUDA ordinal that holds the
    shard; ``row_base`` -- global index of local row 0 (the rank's RankInterval.start_index);
    ``dim`` -- fix the dimension before the first append; ``capacity_rows`` -- pre-size HBM.
    """

    def __init__(self, path: Optional[str], empty: bool, *, device: int = 0, row_base: int = 0,
                 dim: Optional[int] = None, capacity_rows: int = 0, dtype: str = "f32"):
        """PolarsVectorstore::new (polars.rs:80-91)."""
        self.path = path
        self.device = int(device)
        self.row_base = int(row_base)
        self._dtype = {"f32": _lib.BSR_F32, "bf16": _lib.BSR_BF16}[dtype]
        self._capacity_rows = int(capacity_rows)
        self._handle: Optional[C.c_void_p] = None
        self._dim: Optional[int] = None
        if dim is not None:
            self._open(int(dim))
        if not empty:
            if path is None:
                raise ValueError("a path is required unless empty=True")
            rows = _read_parquet_rows(path)
            if rows.shape[0]:
                self.append_many(rows)

    # -- handle management ------------------------------------------------------------------
    def _open(self, dim: int) -> None:
        h = C.c_void_p()
        _lib.check(_lib.lib().bsr_store_create(self.device, dim, self._dtype, self.row_base,
                                               self._capacity_rows, C.byref(h)))
        self._handle, self._dim = h, dim

    @property
    def handle(self) -> C.c_void_p:
        if self._handle is None:
            raise BsrError(_lib.BSR_ERR_INVALID, "the store has no rows and no dimension yet")
        return self._handle

    @property
    def dim(self) -> Optional[int]:
        return self._dim

    def close(self) -> None:
        if self._handle is not None:
            _lib.lib().bsr_store_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # -- PolarsVectorstore API --------------------------------------------------------------
    def reset(self) -> None:
        """polars.rs:93-95."""
        if self._handle is not None:
            _lib.check(_lib.lib().bsr_store_reset(self._handle))

    def append(self, vector: Sequence[float]) -> None:
        """polars.rs:97-99."""
        self.append_many([vector])

    def append_many(self, vectors) -> None:
        """polars.rs:101-119: rows are uploaded to HBM; an empty list is a no-op."""
        if len(vectors) == 0:
            return
        rows = _as_rows(vectors, self._dim)
        if self._handle is None:
            self._open(rows.shape[1])
        _lib.check(_lib.lib().bsr_store_append(self._handle, rows.ctypes.data, rows.shape[0]))

    def append_synthetic(self, seed: int, first_row: int, n_rows: int) -> None:
        """Test fixture sample_vstore / generate_mock_embeddings (src/utils.rs:14-35), generated
        on device; see bsr_store_append_synthetic in include/bsr.h."""
        _lib.check(_lib.lib().bsr_store_append_synthetic(self.handle, seed, first_row, n_rows))

    def get_count(self) -> int:
        """polars.rs:243-246."""
        if self._handle is None:
            return 0
        n = C.c_uint64()
        _lib.check(_lib.lib().bsr_store_count(self._handle, C.byref(n)))
        return int(n.value)

    def get_many(self, slice: Optional[SliceArgs] = None) -> List[np.ndarray]:
        """polars.rs:121-156.  A slice is clamped to the store like DataFrame::slice."""
        rows = self.get_many_array(slice)
        return [rows[i] for i in range(rows.shape[0])]

    def get_many_array(self, slice: Optional[SliceArgs] = None) -> np.ndarray:
        count = self.get_count()
        if slice is None:
            start, length = 0, count
        else:
            off = int(slice.offset)
            start = max(0, count + off) if off < 0 else min(off, count)
            length = max(0, min(int(slice.length), count - start))
        out = np.empty((length, self._dim or 0), dtype=np.float32)
        if length:
            _lib.check(_lib.lib().bsr_store_get_many(self._handle, start, length, out.ctypes.data))
        return out

    def get(self, index: int) -> np.ndarray:
        """polars.rs:158-169."""
        rows = self.get_many_array(SliceArgs(offset=index, length=1))
        if rows.shape[0] == 0:
            raise NoDataError("Index not found")
        return rows[0]

    def reload(self, force: bool) -> None:
        """polars.rs:171-181."""
        rows = _read_parquet_rows(self.path)
        if rows.shape[0] == 0 and not force:
            raise NoDataError("Found a empty or invalid file")
        self.reset()
        if rows.shape[0]:
            if self._dim is not None and rows.shape[1] != self._dim:
                self.close()
                self._dim = None
            self.append_many(rows)

    def persist(self) -> None:
        """polars.rs:183-241: write every row back to `path` as a Parquet file with the single
        `embeddings: list<float>` column."""
        parent = os.path.dirname(self.path)
        if parent and not os.path.exists(parent):
            os.makedirs(parent, exist_ok=True)
        rows = self.get_many_array(None)
        _write_parquet_rows(self.path, rows)
        if not os.path.exists(self.path):
            raise OSError(f"File was not created: {self.path}")

    # -- introspection ------------------------------------------------------------------------
    def scan_bytes(self) -> int:
        b = C.c_uint64()
        _lib.check(_lib.lib().bsr_store_scan_bytes(self.handle, C.byref(b)))
        return int(b.value)

    def launch_count(self) -> int:
        n = C.c_uint64()
        _lib.check(_lib.lib().bsr_store_launch_count(self.handle, C.byref(n)))
        return int(n.value)

    def set_tensor_min_batch(self, min_batch: int) -> None:
        """BF16 stores: smallest query batch that uses the tcgen05 path (0 = always exact scan)."""
        _lib.check(_lib.lib().bsr_store_set_tensor_min_batch(self.handle, int(min_batch)))

    def set_metric(self, metric: str) -> None:
        """"cosine" (the reference's cosine_distance, default) or "dot" (1 - dot product)."""
        m = {"cosine": _lib.BSR_COSINE, "dot": _lib.BSR_DOT}[metric]
        _lib.check(_lib.lib().bsr_store_set_metric(self.handle, m))

    def metric(self) -> str:
        m = C.c_int()
        _lib.check(_lib.lib().bsr_store_metric(self.handle, C.byref(m)))
        return {_lib.BSR_COSINE: "cosine", _lib.BSR_DOT: "dot"}[m.value]

    def set_option(self, name: str, value: int) -> None:
        """Ablation switch / test hook of this handle (bsr_store_set_option, include/bsr.h)."""
        _lib.check(_lib.lib().bsr_store_set_option(self.handle, name.encode(), int(value)))

    def set_profiling(self, enabled: bool) -> None:
        _lib.check(_lib.lib().bsr_store_set_profiling(self.handle, 1 if enabled else 0))

    def scan_profile(self) -> tuple[float, int]:
        """(summed scan-kernel device time in ms, number of scan launches) since set_profiling."""
        ms, n = C.c_double(), C.c_uint64()
        _lib.check(_lib.lib().bsr_store_scan_profile(self.handle, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def last_scan_ms(self) -> float:
        ms = C.c_float()
        _lib.check(_lib.lib().bsr_store_last_scan_ms(self.handle, C.byref(ms)))
        return float(ms.value)


def get_local_vstore(vstore_dir, rank: int, empty: bool, **kw) -> GpuVectorstore:
    """src/mpi_helpers/vectorstore.rs:5-13 -> <dir>/rank_{rank}.parquet."""
    return GpuVectorstore(os.path.join(os.fspath(vstore_dir), f"rank_{rank}.parquet"), empty, **kw)


def get_global_vstore(vstore_dir, empty: bool, **kw) -> GpuVectorstore:
    """src/mpi_helpers/vectorstore.rs:16-20 -> <dir>/global.parquet."""
    return GpuVectorstore(os.path.join(os.fspath(vstore_dir), "global.parquet"), empty, **kw)
