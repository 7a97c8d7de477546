"""better_search_rag_rust_b200 -- B200-native exact top-k search path.

A from-scratch replacement for ONE path of nichmorgan/better-search-rag-rust: the
block-distributed cosine-distance scan and top-k merge (src/mpi_helpers/metrics.rs:16-206,
src/metrics.rs:143-165).  The compute lives in hand-written sm_100a CUDA kernels behind the C ABI
of include/bsr.h (libbsr_b200.so); this package is the host-side mirror of the reference's
vectorstore / search interface.  There is no CPU fallback.
"""
from ._lib import BsrError, BSR_MAX_TOP_K, LIB_PATH  # noqa: F401
from .load_balance import RankInterval, RankSlice, interval_by_rank, slice_by_rank  # noqa: F401
from .metrics import (agreement_report, calculate_accuracy_metrics, format_accuracy_metrics,  # noqa: F401
                      format_top_k_results, mean_reciprocal_rank, print_top_k_results, recall_at_k,
                      result_vectors, top_k_overlap, vectors_are_identical)
from .search import (ROOT, BsrComm, CudaShardEngine, candidate_slots, compute_global_top_k,  # noqa: F401
                     compute_local_top_k, gather_top_k_results, is_root, open_shard,
                     pack_candidates, parallel_top_k_similarity_search, release_shards,
                     unpack_candidates)
from .vectorstore import (GpuVectorstore, NoDataError, ParquetRowSource, SliceArgs,  # noqa: F401
                          get_global_vstore, get_local_vstore, rank_file_paths)

__version__ = "0.2.0"
