/*
 * bsr.h -- C ABI of the B200-native exact top-k search path (libbsr_b200.so).
 *
 * Drop-in boundary for ONE path of nichmorgan/better-search-rag-rust: the block-distributed
 * cosine-distance scan and top-k merge.  The reference has no FFI of its own (100 % safe
 * Rust); every entry point below names the Rust item a maintainer would re-implement on top
 * of it (citations relative to the reference tree).  INTEGRATION.md shows the Rust binding.
 *
 * Conventions
 *   - every function returns an int status: BSR_OK (0) or a negative bsr_status; nothing
 *     unwinds across the boundary; bsr_last_error() gives a thread-local message.
 *   - the library owns all device memory behind the opaque bsr_store handle; host buffers are
 *     caller-owned and only borrowed for the duration of the call.
 *   - one handle is used from one host thread at a time (the reference's caller is
 *     single-threaded per rank, src/main.rs:25-189); distinct handles are independent.
 *   - one handle = one shard = one GPU = one rank (reference: one MPI process,
 *     src/mpi_helpers/metrics.rs:16-53).  Row indices reported by searches are GLOBAL:
 *     row_base + local row, as the reference tags them (src/mpi_helpers/metrics.rs:41).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     BSR_ERR_CUDA.
 *
 * Results are bit-identical to the reference's CPU arithmetic (src/metrics.rs:143-165) on every
 * path -- single queries, query batches, tensor-core batches, one rank or many: the reference's
 * strict left-to-right binary32 sums are kept (no FMA contraction), so distances match bit for bit
 * and ties resolve as the reference's stable sorts do (ascending global row index).
 */
#ifndef BSR_B200_H
#define BSR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BSR_API __attribute__((visibility("default")))
#else
#define BSR_API
#endif

#define BSR_VERSION 200 /* 0.2.0 */

typedef enum bsr_status {
    BSR_OK = 0,
    BSR_ERR_INVALID = -1,     /* bad argument (null pointer, dim == 0, top_k out of range ...) */
    BSR_ERR_CUDA = -2,        /* CUDA runtime error, or no usable device */
    BSR_ERR_OOM = -3,         /* device or pinned-host allocation failed */
    BSR_ERR_RANGE = -4,       /* row index / slice outside the store ("Index not found") */
    BSR_ERR_UNSUPPORTED = -5  /* valid request this build does not implement */
} bsr_status;

typedef enum bsr_dtype {
    BSR_F32 = 0, /* rows kept as binary32: the reference's own element type */
    BSR_BF16 = 1 /* rows rounded (RNE) to bfloat16 at load: half the HBM traffic; results equal
                    the reference arithmetic applied to the rounded rows */
} bsr_dtype;

/* Score a search ranks by.  The reference has cosine distance only (src/metrics.rs:143-165);
 * BASELINE.json's north_star adds the dot product for stores that are normalised at ingest. */
typedef enum bsr_metric {
    BSR_COSINE = 0, /* cosine_distance, src/metrics.rs:143-165 (identity shortcut, zero-norm rule, clamp) */
    BSR_DOT = 1     /* distance = 1 - dot(a, b): the reference's sequential binary32 fold of src/metrics.rs:153
                       alone -- no norms, no clamp, no identity shortcut.  Equals the cosine distance on
                       unit-norm rows and queries up to the rounding of the norms.  A NaN distance (NaN or
                       inf*0 in the inputs) sorts after every number. */
} bsr_metric;

/* Largest top_k a search accepts (per-warp candidate lists live in shared memory). */
#define BSR_MAX_TOP_K 1024u

typedef struct bsr_store bsr_store;

/* ---- library ---------------------------------------------------------------------------- */
BSR_API int bsr_version(void);
BSR_API const char *bsr_last_error(void);
BSR_API int bsr_device_count(int *n_devices);

/* ---- store: mirrors PolarsVectorstore (src/vectorstore/polars.rs:79-247) ------------------ */

/* PolarsVectorstore::new(path, empty = true) (polars.rs:80-91): an empty shard on `device`.
 * row_base is the global index of local row 0, i.e. RankInterval::start_index of this rank
 * (src/mpi_helpers/load_balance.rs:24-42).  capacity_rows pre-sizes the HBM allocation
 * (0 = grow on demand). */
BSR_API int bsr_store_create(int device, uint32_t dim, bsr_dtype dtype, uint64_t row_base,
                     uint64_t capacity_rows, bsr_store **out);
BSR_API int bsr_store_destroy(bsr_store *store);

/* PolarsVectorstore::reset (polars.rs:93-95): drop all rows, keep the allocation. */
BSR_API int bsr_store_reset(bsr_store *store);

/* PolarsVectorstore::append_many (polars.rs:101-119): n_rows row-major rows of `dim` floats
 * from host memory are appended to the HBM-resident shard; row norms are computed on device.
 * append (polars.rs:97-99) is n_rows = 1. */
BSR_API int bsr_store_append(bsr_store *store, const float *rows, uint64_t n_rows);

/* Same, rows already in device memory on the store's device (row-major, 16-byte aligned). */
BSR_API int bsr_store_append_device(bsr_store *store, const float *d_rows, uint64_t n_rows);

/* Test fixture generate_mock_embeddings / sample_vstore (src/utils.rs:14-35): append n_rows
 * uniform(-1,1) rows generated ON DEVICE by the counter-based generator documented in
 * oracle/bsr_oracle.c (bsr_oracle_generate_rows); row r of the call is generator row
 * first_row + r. */
BSR_API int bsr_store_append_synthetic(bsr_store *store, uint64_t seed, uint64_t first_row,
                               uint64_t n_rows);

/* Metric used by every search on this handle (default BSR_COSINE).  Takes effect from the next
 * search; no data is touched (row norms are kept either way). */
BSR_API int bsr_store_set_metric(bsr_store *store, bsr_metric metric);
BSR_API int bsr_store_metric(const bsr_store *store, bsr_metric *metric);

/* PolarsVectorstore::get_count (polars.rs:243-246). */
BSR_API int bsr_store_count(const bsr_store *store, uint64_t *n_rows);
BSR_API int bsr_store_dim(const bsr_store *store, uint32_t *dim);
BSR_API int bsr_store_row_base(const bsr_store *store, uint64_t *row_base);
BSR_API int bsr_store_device(const bsr_store *store, int *device);

/* Bytes of embedding data one scan of this shard streams from HBM (rows * padded dim *
 * sizeof(element)): the "algorithmic bytes" of DESIGN.md, excluding the norm side array. */
BSR_API int bsr_store_scan_bytes(const bsr_store *store, uint64_t *bytes);

/* PolarsVectorstore::get_many(Some(SliceArgs{offset,length})) / get (polars.rs:121-169):
 * copy local rows [offset, offset+length) to host, row-major, bit-identical to what was
 * appended (BSR_F32) or to its bf16 rounding (BSR_BF16).  Unlike polars' clamping slice the
 * range must lie inside the store (BSR_ERR_RANGE otherwise: get()'s "Index not found"). */
BSR_API int bsr_store_get_many(const bsr_store *store, uint64_t offset, uint64_t length, float *out_rows);

/* ---- search: mirrors src/mpi_helpers/metrics.rs ------------------------------------------- */

/* compute_local_top_k (metrics.rs:16-53) for this shard, host buffers in and out.
 *   queries   [n_queries][dim] row-major (the reference has n_queries == 1)
 *   out_idx   [n_queries][top_k] global row indices
 *   out_dist  [n_queries][top_k] cosine distances (src/metrics.rs:143-165)
 *   out_count [n_queries]        min(top_k, rows in this shard)
 * Ascending distance, ties by ascending index (what the reference's stable sort yields).
 * 1 <= top_k <= BSR_MAX_TOP_K. */
BSR_API int bsr_search(bsr_store *store, const float *queries, uint32_t n_queries, uint32_t top_k,
               uint64_t *out_idx, float *out_dist, uint32_t *out_count);

/* ---- device-level pieces, for the multi-rank exchange (gather_top_k_results,
 *      metrics.rs:56-138, becomes an all-gather of candidate slots on the host's communicator,
 *      compute_global_top_k, metrics.rs:141-171, becomes bsr_merge_device) ------------------- */

/* A candidate is one uint64: (order-preserving bits of the f32 distance) << 32 | global row
 * index.  Unsigned ascending order of candidates == the reference's (distance, index) order.
 * Each query owns bsr_candidate_slots(top_k) slots (a power of two >= 32); the first
 * min(top_k, rows) slots hold the shard's top-k in ascending order, every other slot holds
 * UINT64_MAX. */
BSR_API uint32_t bsr_candidate_slots(uint32_t top_k);

/* Scan this shard for n_queries device-resident queries; writes the shard's sorted candidates
 * to d_candidates[n_queries][slots] on `stream` (a cudaStream_t; NULL = the CUDA default stream, as in the runtime API).
 * Asynchronous with respect to the host. */
BSR_API int bsr_search_device(bsr_store *store, const float *d_queries, uint32_t n_queries,
                      uint32_t top_k, uint64_t *d_candidates, void *stream);

/* Merge n_lists candidate lists per query, laid out [n_lists][n_queries][slots] (the layout an
 * all-gather of per-rank [n_queries][slots] buffers produces, rank-ordered as metrics.rs:86-89),
 * into d_out[n_queries][slots].  Shards own disjoint row blocks, so candidate indices are
 * unique by construction and the reference's HashSet de-dup (metrics.rs:156-168) is vacuous;
 * the merge does not re-check it.  Scratch is taken from `store`. */
BSR_API int bsr_merge_device(bsr_store *store, const uint64_t *d_lists, uint32_t n_lists,
                     uint32_t n_queries, uint32_t top_k, uint64_t *d_out, void *stream);

/* Unpack candidates[n_queries][slots] into d_idx[n_queries][top_k], d_dist[n_queries][top_k],
 * d_count[n_queries] (all device pointers). */
BSR_API int bsr_unpack_device(bsr_store *store, const uint64_t *d_candidates, uint32_t n_queries,
                      uint32_t top_k, uint64_t *d_idx, float *d_dist, uint32_t *d_count,
                      void *stream);

/* compute_global_top_k (metrics.rs:141-171) with host buffers: the n gathered (index, distance)
 * pairs -- the rank-ordered concatenation gather_top_k_results builds at root -- are sorted on the
 * device by (distance, index) and the first min(top_k, distinct indices) written to
 * out_idx/out_dist; *out_count receives how many.  An index that occurs more than once keeps its
 * first occurrence in sorted order, as the reference's HashSet pass does (metrics.rs:156-168).
 * Indices must be < 2^32 - 1. */
BSR_API int bsr_merge_host(bsr_store *store, const uint64_t *idx, const float *dist, uint64_t n,
                           uint32_t top_k, uint64_t *out_idx, float *out_dist, uint32_t *out_count);

/* ---- multi-rank search inside the library (for hosts without their own GPU communicator) ------
 * One rank = one process = one GPU = one bsr_store shard, as in the reference (one MPI process per
 * rank).  The candidate exchange (gather_top_k_results, metrics.rs:56-138) is an ncclAllGather of
 * the per-shard candidate slots over NVLink; NCCL is resolved at run time (dlopen of libnccl.so.2,
 * reusing the copy already loaded in the process if there is one). */
#define BSR_COMM_ID_BYTES 128
typedef struct bsr_comm bsr_comm;

/* Rank 0 creates the id and hands it to every rank over the host's own channel (in the reference's
 * world: one MPI_Bcast, like the query at src/main.rs:123-125). */
BSR_API int bsr_comm_unique_id(uint8_t id[BSR_COMM_ID_BYTES]);
/* Collective: every rank calls it with the same id and n_ranks; `device` is the rank's GPU. */
BSR_API int bsr_comm_create(const uint8_t id[BSR_COMM_ID_BYTES], int n_ranks, int rank, int device,
                            bsr_comm **out);
BSR_API int bsr_comm_destroy(bsr_comm *comm);

/* parallel_top_k_similarity_search (metrics.rs:174-206).  Collective: every rank passes its shard
 * and the same queries (host memory).  Scan of the local shard -> all-gather of the candidate slots
 * -> on root: on-device merge, unpack, copy to the host outputs (as bsr_search).  On the other
 * ranks the outputs are not written (the reference returns None there); they may be NULL.
 * A rank whose local scan fails still joins the exchange with an empty list (the reference's
 * behaviour, metrics.rs:185-191) and returns its error afterwards; the other ranks complete.
 * Single queries with top_k <= 128 skip the collective call altogether when the communicator has a
 * mailbox (bsr_comm_uses_mailbox). */
BSR_API int bsr_search_parallel(bsr_store *shard, bsr_comm *comm, const float *queries,
                                uint32_t n_queries, uint32_t top_k, uint64_t *out_idx,
                                float *out_dist, uint32_t *out_count);

/* The same collective with device-resident queries and candidates, asynchronous on `stream` (what a
 * host keeps in flight back to back).  On root d_candidates[n_queries][slots] receives the merged
 * candidates of the whole store; on the other ranks its contents are unspecified. */
BSR_API int bsr_search_parallel_device(bsr_store *shard, bsr_comm *comm, const float *d_queries,
                                       uint32_t n_queries, uint32_t top_k, uint64_t *d_candidates,
                                       void *stream);

/* 1 when this communicator exchanges single-query, top_k <= 128 candidates through the NVLink mailbox
 * (peer stores from the scan's last CTA into root's HBM, folded in by root's last CTA: one kernel per
 * rank and search, no collective call); 0 when every search uses ncclAllGather + a merge launch
 * (ranks in one process, two ranks on one GPU, no peer access, BSR_NO_MAILBOX set at create time).
 * The per-handle option "no_mailbox" selects the ncclAllGather path for one shard's searches; it must
 * be set identically on every rank. */
BSR_API int bsr_comm_uses_mailbox(const bsr_comm *comm, int *yes);

/* Query batches of at least `min_batch` queries (default 8; 0 = never) are scanned on the tensor
 * cores (tcgen05, queries in tensor memory) -- as a FILTER: bf16-rounded queries against the bf16
 * rows (an fp32 store gets a bf16 shadow copy, +50 % HBM, built on the first such batch), a certified
 * bound with a margin of twice the worst-case rounding gap, and every survivor's distance recomputed
 * with the reference's arithmetic over the stored row and the query as given.  Results are
 * bit-identical to the exact scan's; only the time differs.  Needs dim <= 768; other shapes always
 * use the exact scan.  Queries without a usable direction (zero, denormal-small, NaN or inf magnitude)
 * are answered by the exact scan (bsr_search sorts them out; with device-resident queries they stay
 * correct but slow their launch). */
BSR_API int bsr_store_set_tensor_min_batch(bsr_store *store, uint32_t min_batch);

/* Ablation switches and test hooks, per handle (DESIGN.md lists them).  Defaults are the measured
 * optimum; the same names upper-cased with a BSR_ prefix are read from the environment ONCE, when
 * the library is first used, and seed every new handle.  Unknown name: BSR_ERR_INVALID.
 *   "max_grid" n      cap the scan grid at n CTAs (test hook: long per-warp row streams on small stores)
 *   "bulk_stages" n   ring depth of the bulk-async scan        "max_qb" n     widest exact query group
 *   "scan_ldg" 0/1    register-staged scan kernel              "no_wide" 0/1  no 16-warp CTAs
 *   "no_fuse_wide", "tensor_mode", "tensor_stages", "tensor_stage_kb", "tc_stats", "no_pdl", "no_mailbox",
 *   "no_small" (small stores through the streaming kernel), "no_list_ll", "no_l2_keep" */
BSR_API int bsr_store_set_option(bsr_store *store, const char *name, int64_t value);

/* ---- introspection for the benchmark harness ---------------------------------------------- */

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
BSR_API int bsr_store_launch_count(const bsr_store *store, uint64_t *n_launches);

/* Device time of the most recent scan stage on this handle, in milliseconds.  A search that is one
 * fused launch answered through pinned host memory (bsr_search / bsr_search_parallel, one query,
 * top_k <= 128) carries no events in its stream: the launch measures itself (%globaltimer of its first
 * CTA's first instruction to the store of the result) and that duration is reported.  Every other
 * search: CUDA events recorded on the stream the kernels were launched on (synchronises that stream). */
BSR_API int bsr_store_last_scan_ms(bsr_store *store, float *ms);

/* Per-launch timing of the scan kernel: while enabled, every scan launch is bracketed by CUDA
 * events on the stream it is launched on (no host synchronisation at launch time).
 * bsr_store_scan_profile synchronises, then returns the summed device time and the number of
 * scan launches since profiling was (re-)enabled. */
BSR_API int bsr_store_set_profiling(bsr_store *store, int enabled);
BSR_API int bsr_store_scan_profile(bsr_store *store, double *total_ms, uint64_t *n_scans);

#ifdef __cplusplus
}
#endif
#endif /* BSR_B200_H */
