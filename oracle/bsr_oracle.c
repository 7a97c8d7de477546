/*
 * bsr_oracle.c -- CPU restatement of the reference's exact top-k search path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or as the timed CPU
 * baseline.  The product path (better_search_rag_rust_b200 + libbsr_b200.so)
 * never links, imports or calls it.
 *
 * PARITY UNPINNED for the search path proper: the reference
 * (nichmorgan/better-search-rag-rust) holds no golden vector, known-answer test
 * or fixture for it (no test calls cosine_distance or any mpi_helpers::metrics
 * function) and cannot be built here (no cargo/rustc, no MPI).  This file is
 * therefore a line-by-line restatement of the reference source, checked against
 * (i) hand-derived known answers, (ii) the run-time self-query invariant of
 * src/main.rs:111,120,148-154 and (iii) an independent numpy restatement
 * (oracle/numpy_oracle.py).
 * PINNED by the reference's own tests: vectors_are_identical and the three IR
 * metrics built on it (mean_reciprocal_rank, recall_at_k, top_k_overlap) -- the
 * known answers of src/metrics.rs:172-317 (1.0; ~0.611; 0.25/0.5/0.75/1.0; 0.8)
 * are reproduced by tests/test_ir_metrics.py on inputs of the same construction.
 *
 * All arithmetic lives in the reference tree itself plus Rust core
 * (Iterator::sum::<f32>, f32::sqrt/max/min, slice::sort_by -- stable); the
 * third-party crates on the path (polars =0.47.0, mpi 0.8.0) only move data.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  With
 * those flags every f32 operation below is a single correctly rounded IEEE-754
 * binary32 operation in source order, which is what rustc emits for the
 * reference (Rust never contracts a*b+c into an FMA and never reassociates
 * float sums).
 *
 * Reference citations are relative to /root/reference/.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

#define BSR_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* src/metrics.rs:7-19  vectors_are_identical                                 */
/* ------------------------------------------------------------------------- */
BSR_EXPORT int bsr_oracle_vectors_are_identical(const float *a, uint64_t na,
                                                const float *b, uint64_t nb)
{
    if (na != nb) /* :8-10 */
        return 0;
    for (uint64_t i = 0; i < na; i++) { /* :12-16 */
        /* `(a[i] - b[i]).abs() > 1e-10` with an f32 literal; a NaN difference
         * compares false, i.e. NaN elements count as identical. */
        if (fabsf(a[i] - b[i]) > 1e-10f)
            return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* src/metrics.rs:143-165  cosine_distance                                    */
/* ------------------------------------------------------------------------- */
BSR_EXPORT float bsr_oracle_cosine_distance(const float *a, uint64_t na,
                                            const float *b, uint64_t nb)
{
    if (na != nb || na == 0) /* :144-146 */
        return 1.0f;
    if (bsr_oracle_vectors_are_identical(a, na, b, nb)) /* :148-151 */
        return 0.0f;

    /* :153-155 -- three separate Iterator::sum::<f32>() folds: strict
     * left-to-right adds of individually rounded products. */
    float dot = 0.0f;
    for (uint64_t i = 0; i < na; i++)
        dot = dot + a[i] * b[i];
    float sq_a = 0.0f;
    for (uint64_t i = 0; i < na; i++)
        sq_a = sq_a + a[i] * a[i];
    float sq_b = 0.0f;
    for (uint64_t i = 0; i < nb; i++)
        sq_b = sq_b + b[i] * b[i];
    float mag_a = sqrtf(sq_a);
    float mag_b = sqrtf(sq_b);

    if (mag_a == 0.0f || mag_b == 0.0f) /* :157-159 */
        return 1.0f;

    float sim = dot / (mag_a * mag_b); /* :161 */
    /* :162 f32::max / f32::min return the non-NaN operand, as fmaxf/fminf do:
     * a NaN similarity becomes -1.0 and the distance 2.0. */
    sim = fminf(fmaxf(sim, -1.0f), 1.0f);
    return 1.0f - sim; /* :164 */
}

/* ------------------------------------------------------------------------- */
/* BSR_DOT (include/bsr.h; an extension, the reference has cosine only):      */
/* distance = 1 - dot, dot being the fold of src/metrics.rs:153 alone.        */
/* Length mismatch / empty input follow cosine_distance's first rule (:144).  */
/* ------------------------------------------------------------------------- */
BSR_EXPORT float bsr_oracle_dot_distance(const float *a, uint64_t na, const float *b,
                                         uint64_t nb)
{
    if (na != nb || na == 0)
        return 1.0f;
    float dot = 0.0f;
    for (uint64_t i = 0; i < na; i++)
        dot = dot + a[i] * b[i];
    return 1.0f - dot;
}

/* total order used for BSR_DOT distances: numbers ascending (-0 == +0), every
 * NaN after every number (the order include/bsr.h documents) */
static uint32_t ordered_bits(float d)
{
    uint32_t b;
    d = d + 0.0f;
    memcpy(&b, &d, 4);
    if (d != d)
        b = 0x7FC00000u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

/* ------------------------------------------------------------------------- */
/* src/mpi_helpers/load_balance.rs:8-42  RankInterval / interval_by_rank      */
/* ------------------------------------------------------------------------- */
BSR_EXPORT void bsr_oracle_interval_by_rank(int32_t rank, int32_t size, uint64_t count,
                                            uint64_t *start_index, uint64_t *end_index)
{
    uint64_t usize = (uint64_t)size;
    uint64_t per_rank = (usize > count) ? 1 : (count + usize - 1) / usize; /* :25-29 */
    uint64_t start = per_rank * (uint64_t)rank;                             /* :31 (unclamped) */
    uint64_t end;
    if (rank == size - 1) /* :32-36 */
        end = count;
    else
        end = (start + per_rank < count) ? start + per_rank : count;
    *start_index = start;
    *end_index = end;
}

/* RankInterval::get_count (:14-16) in a release build (makefile:9): wrapping
 * usize subtraction. */
BSR_EXPORT uint64_t bsr_oracle_interval_count(uint64_t start_index, uint64_t end_index)
{
    return end_index - start_index;
}

/* The rows `PolarsVectorstore::get_many(Some(SliceArgs{offset,length}))`
 * (src/vectorstore/polars.rs:121-129) actually yields for that interval:
 * DataFrame::slice clamps to the frame, so an offset at or past the end gives an
 * empty block whatever the (possibly wrapped) length is. */
static void effective_block(uint64_t start, uint64_t end, uint64_t count,
                            uint64_t *first, uint64_t *n)
{
    uint64_t len = end - start; /* wrapped like get_count() */
    if (start >= count) {
        *first = count;
        *n = 0;
        return;
    }
    uint64_t avail = count - start;
    *first = start;
    *n = len < avail ? len : avail;
}

/* ------------------------------------------------------------------------- */
/* stable ascending sort by distance only (slice::sort_by is stable;          */
/* comparator is partial_cmp on the f32, src/mpi_helpers/metrics.rs:47,153)   */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t idx;
    float dist;
} pair_t;

static inline int pair_less(const pair_t *x, const pair_t *y, int total_order)
{
    if (total_order)
        return ordered_bits(x->dist) < ordered_bits(y->dist);
    return x->dist < y->dist; /* partial_cmp: the reference's comparator */
}

static void stable_sort_pairs_ord(pair_t *v, uint64_t n, int total_order)
{
    if (n < 2)
        return;
    pair_t *tmp = (pair_t *)malloc(n * sizeof(pair_t));
    pair_t *src = v, *dst = tmp;
    for (uint64_t width = 1; width < n; width *= 2) {
        for (uint64_t lo = 0; lo < n; lo += 2 * width) {
            uint64_t mid = lo + width < n ? lo + width : n;
            uint64_t hi = lo + 2 * width < n ? lo + 2 * width : n;
            uint64_t i = lo, j = mid, o = lo;
            while (i < mid && j < hi) {
                /* take the right element only when strictly smaller: stability */
                if (pair_less(&src[j], &src[i], total_order))
                    dst[o++] = src[j++];
                else
                    dst[o++] = src[i++];
            }
            while (i < mid)
                dst[o++] = src[i++];
            while (j < hi)
                dst[o++] = src[j++];
        }
        pair_t *t = src;
        src = dst;
        dst = t;
    }
    if (src != v)
        memcpy(v, src, n * sizeof(pair_t));
    free(tmp);
}

/* ------------------------------------------------------------------------- */
/* src/mpi_helpers/metrics.rs:16-53  compute_local_top_k                      */
/*   rows: the WHOLE global store, row-major [count][dim] (every rank opens    */
/*   global.parquet, :23).  Returns the number of pairs written (<= top_k).    */
/* ------------------------------------------------------------------------- */
BSR_EXPORT uint64_t bsr_oracle_compute_local_top_k_metric(const float *rows, uint64_t count,
                                                          uint32_t dim, int32_t rank, int32_t size,
                                                          uint64_t top_k, const float *target,
                                                          uint32_t target_len, int metric,
                                                          uint64_t *out_idx, float *out_dist);

BSR_EXPORT uint64_t bsr_oracle_compute_local_top_k(const float *rows, uint64_t count,
                                                   uint32_t dim, int32_t rank, int32_t size,
                                                   uint64_t top_k, const float *target,
                                                   uint32_t target_len, uint64_t *out_idx,
                                                   float *out_dist)
{
    return bsr_oracle_compute_local_top_k_metric(rows, count, dim, rank, size, top_k, target,
                                                 target_len, 0, out_idx, out_dist);
}

/* metric: 0 = cosine_distance (the reference), 1 = BSR_DOT */
BSR_EXPORT uint64_t bsr_oracle_compute_local_top_k_metric(const float *rows, uint64_t count,
                                                          uint32_t dim, int32_t rank, int32_t size,
                                                          uint64_t top_k, const float *target,
                                                          uint32_t target_len, int metric,
                                                          uint64_t *out_idx, float *out_dist)
{
    uint64_t start, end, first, n;
    bsr_oracle_interval_by_rank(rank, size, count, &start, &end); /* :27 */
    effective_block(start, end, count, &first, &n);               /* :30-33 */
    if (n == 0)
        return 0;
    pair_t *d = (pair_t *)malloc(n * sizeof(pair_t));
    for (uint64_t i = 0; i < n; i++) { /* :36-44 */
        d[i].idx = start + i;          /* :41 global index */
        const float *row = rows + (first + i) * (uint64_t)dim;
        d[i].dist = metric ? bsr_oracle_dot_distance(row, dim, target, target_len)
                           : bsr_oracle_cosine_distance(row, dim, target, target_len);
    }
    stable_sort_pairs_ord(d, n, metric != 0); /* :47 */
    uint64_t keep = n > top_k ? top_k : n; /* :48-50 */
    for (uint64_t i = 0; i < keep; i++) {
        out_idx[i] = d[i].idx;
        out_dist[i] = d[i].dist;
    }
    free(d);
    return keep;
}

/* ------------------------------------------------------------------------- */
/* src/mpi_helpers/metrics.rs:141-171  compute_global_top_k                   */
/*   input: the rank-ordered concatenation gather_top_k_results builds at     */
/*   root (:86-126).                                                          */
/* ------------------------------------------------------------------------- */
BSR_EXPORT uint64_t bsr_oracle_compute_global_top_k_metric(const uint64_t *global_idx,
                                                           const float *global_dist, uint64_t n,
                                                           uint64_t top_k, int metric,
                                                           uint64_t *out_idx, float *out_dist);

BSR_EXPORT uint64_t bsr_oracle_compute_global_top_k(const uint64_t *global_idx,
                                                    const float *global_dist, uint64_t n,
                                                    uint64_t top_k, uint64_t *out_idx,
                                                    float *out_dist)
{
    return bsr_oracle_compute_global_top_k_metric(global_idx, global_dist, n, top_k, 0, out_idx,
                                                  out_dist);
}

BSR_EXPORT uint64_t bsr_oracle_compute_global_top_k_metric(const uint64_t *global_idx,
                                                           const float *global_dist, uint64_t n,
                                                           uint64_t top_k, int metric,
                                                           uint64_t *out_idx, float *out_dist)
{
    if (n == 0 || top_k == 0)
        return 0;
    pair_t *g = (pair_t *)malloc(n * sizeof(pair_t));
    for (uint64_t i = 0; i < n; i++) { /* :147-150 zip */
        g[i].idx = global_idx[i];
        g[i].dist = global_dist[i];
    }
    stable_sort_pairs_ord(g, n, metric != 0); /* :153 */

    /* :156-168 HashSet de-dup on the index, first top_k survivors. */
    uint64_t cap = 16;
    while (cap < 2 * n)
        cap *= 2;
    uint64_t *set = (uint64_t *)malloc(cap * sizeof(uint64_t));
    uint8_t *used = (uint8_t *)calloc(cap, 1);
    uint64_t kept = 0;
    for (uint64_t i = 0; i < n && kept < top_k; i++) {
        uint64_t h = (g[i].idx * 0x9E3779B97F4A7C15ull) & (cap - 1);
        int seen = 0;
        while (used[h]) {
            if (set[h] == g[i].idx) {
                seen = 1;
                break;
            }
            h = (h + 1) & (cap - 1);
        }
        if (seen)
            continue;
        used[h] = 1;
        set[h] = g[i].idx;
        out_idx[kept] = g[i].idx;
        out_dist[kept] = g[i].dist;
        kept++;
    }
    free(set);
    free(used);
    free(g);
    return kept;
}

/* ------------------------------------------------------------------------- */
/* src/mpi_helpers/metrics.rs:174-206  parallel_top_k_similarity_search       */
/*   The reference's N MPI processes (makefile:13-14) become N OS threads;     */
/*   gather_top_k_results (:56-138) becomes the rank-ordered concatenation     */
/*   below.  Returns root's Some(top_k) list.                                  */
/* ------------------------------------------------------------------------- */
typedef struct {
    const float *rows;
    uint64_t count;
    uint32_t dim;
    int32_t rank, size;
    uint64_t top_k;
    const float *target;
    uint32_t target_len;
    uint64_t *idx;
    float *dist;
    uint64_t n_out;
} rank_job_t;

static void *rank_main(void *p)
{
    rank_job_t *j = (rank_job_t *)p;
    j->n_out = bsr_oracle_compute_local_top_k(j->rows, j->count, j->dim, j->rank, j->size,
                                              j->top_k, j->target, j->target_len, j->idx,
                                              j->dist);
    return NULL;
}

BSR_EXPORT uint64_t bsr_oracle_parallel_top_k_similarity_search(
    const float *rows, uint64_t count, uint32_t dim, int32_t size, uint64_t top_k,
    const float *target, uint32_t target_len, int32_t n_threads, uint64_t *out_idx,
    float *out_dist)
{
    if (size < 1)
        return 0;
    rank_job_t *jobs = (rank_job_t *)calloc((size_t)size, sizeof(rank_job_t));
    uint64_t slot = top_k ? top_k : 1;
    uint64_t *all_idx = (uint64_t *)malloc((size_t)size * slot * sizeof(uint64_t));
    float *all_dist = (float *)malloc((size_t)size * slot * sizeof(float));
    for (int32_t r = 0; r < size; r++) {
        jobs[r] = (rank_job_t){rows, count, dim, r, size, top_k, target, target_len,
                               all_idx + (uint64_t)r * slot, all_dist + (uint64_t)r * slot, 0};
    }
    if (n_threads <= 1) {
        for (int32_t r = 0; r < size; r++)
            rank_main(&jobs[r]);
    } else {
        /* at most n_threads ranks in flight at a time */
        pthread_t *th = (pthread_t *)malloc((size_t)size * sizeof(pthread_t));
        for (int32_t base = 0; base < size; base += n_threads) {
            int32_t lim = base + n_threads < size ? base + n_threads : size;
            for (int32_t r = base; r < lim; r++)
                pthread_create(&th[r], NULL, rank_main, &jobs[r]);
            for (int32_t r = base; r < lim; r++)
                pthread_join(th[r], NULL);
        }
        free(th);
    }
    /* gather: rank-ordered concatenation (displacements, :86-89) */
    uint64_t total = 0;
    for (int32_t r = 0; r < size; r++)
        total += jobs[r].n_out;
    uint64_t *g_idx = (uint64_t *)malloc((total ? total : 1) * sizeof(uint64_t));
    float *g_dist = (float *)malloc((total ? total : 1) * sizeof(float));
    uint64_t o = 0;
    for (int32_t r = 0; r < size; r++) {
        memcpy(g_idx + o, jobs[r].idx, jobs[r].n_out * sizeof(uint64_t));
        memcpy(g_dist + o, jobs[r].dist, jobs[r].n_out * sizeof(float));
        o += jobs[r].n_out;
    }
    uint64_t kept = bsr_oracle_compute_global_top_k(g_idx, g_dist, total, top_k, out_idx, out_dist);
    free(g_idx);
    free(g_dist);
    free(all_idx);
    free(all_dist);
    free(jobs);
    return kept;
}

/* ------------------------------------------------------------------------- */
/* src/metrics.rs:21-141  mean_reciprocal_rank / recall_at_k / top_k_overlap   */
/*   A Vec<Vec<Vec<f32>>> arrives flattened: results[q] is n_results[q]        */
/*   vectors of `dim` floats, queries back to back.  (The reference's ids_map   */
/*   argument only feeds an unused binding, :38-41,79-82, so it has no effect   */
/*   and no counterpart here.)                                                 */
/* ------------------------------------------------------------------------- */
BSR_EXPORT float bsr_oracle_mean_reciprocal_rank(const float *queries, uint64_t n_queries,
                                                 const float *results,
                                                 const uint64_t *n_results,
                                                 uint64_t n_result_lists, uint32_t dim)
{
    if (n_queries == 0 || n_result_lists == 0 || n_queries != n_result_lists) /* :26-28 */
        return 0.0f;
    float sum = 0.0f;
    float total = (float)n_queries;
    const float *r = results;
    for (uint64_t q = 0; q < n_queries; q++) { /* :33-55 */
        uint64_t rank = 0;
        for (uint64_t i = 0; i < n_results[q]; i++) {
            if (bsr_oracle_vectors_are_identical(queries + q * (uint64_t)dim, dim,
                                                 r + i * (uint64_t)dim, dim)) {
                rank = i + 1;
                break;
            }
        }
        if (rank == 0)
            rank = n_results[q] + 1; /* :50-52 */
        sum += 1.0f / (float)rank;
        r += n_results[q] * (uint64_t)dim;
    }
    return sum / total;
}

BSR_EXPORT float bsr_oracle_recall_at_k(const float *queries, uint64_t n_queries,
                                        const float *results, const uint64_t *n_results,
                                        uint64_t n_result_lists, uint32_t dim, uint64_t k)
{
    if (n_queries == 0 || n_result_lists == 0 || n_queries != n_result_lists) /* :66-68 */
        return 0.0f;
    float sum = 0.0f;
    float total = (float)n_queries;
    const float *r = results;
    for (uint64_t q = 0; q < n_queries; q++) { /* :73-93 */
        uint64_t actual_k = k < n_results[q] ? k : n_results[q];
        int found = 0;
        for (uint64_t i = 0; i < actual_k; i++) {
            if (bsr_oracle_vectors_are_identical(queries + q * (uint64_t)dim, dim,
                                                 r + i * (uint64_t)dim, dim)) {
                found = 1;
                break;
            }
        }
        sum += found ? 1.0f : 0.0f;
        r += n_results[q] * (uint64_t)dim;
    }
    return sum / total;
}

BSR_EXPORT float bsr_oracle_top_k_overlap(const float *seq, const uint64_t *n_seq,
                                          uint64_t n_seq_lists, const float *par,
                                          const uint64_t *n_par, uint64_t n_par_lists,
                                          uint32_t dim, uint64_t k)
{
    if (n_seq_lists == 0 || n_par_lists == 0 || n_seq_lists != n_par_lists) /* :103-108 */
        return 0.0f;
    float sum = 0.0f;
    float total = (float)n_seq_lists;
    const float *s = seq, *p = par;
    for (uint64_t q = 0; q < n_seq_lists; q++) { /* :113-138 */
        uint64_t m = n_seq[q] < n_par[q] ? n_seq[q] : n_par[q];
        uint64_t actual_k = k < m ? k : m;
        if (actual_k != 0) {
            uint64_t overlap = 0;
            for (uint64_t i = 0; i < actual_k; i++) {
                for (uint64_t j = 0; j < actual_k; j++) {
                    if (bsr_oracle_vectors_are_identical(s + i * (uint64_t)dim, dim,
                                                         p + j * (uint64_t)dim, dim)) {
                        overlap++;
                        break;
                    }
                }
            }
            sum += (float)overlap / (float)actual_k;
        }
        s += n_seq[q] * (uint64_t)dim;
        p += n_par[q] * (uint64_t)dim;
    }
    return sum / total;
}

/* ------------------------------------------------------------------------- */
/* src/mpi_helpers/metrics.rs:217-249  calculate_accuracy_metrics             */
/* ------------------------------------------------------------------------- */
BSR_EXPORT void bsr_oracle_calculate_accuracy_metrics(const uint64_t *idx, uint64_t n,
                                                      uint64_t query_idx, uint64_t top_k,
                                                      float *mrr, float *recall, float *overlap)
{
    uint64_t position = 0;
    for (uint64_t i = 0; i < n; i++) { /* :223-229 */
        if (idx[i] == query_idx) {
            position = i + 1;
            break;
        }
    }
    *mrr = position > 0 ? 1.0f / (float)position : 0.0f;        /* :232-236 */
    *recall = (position > 0 && position <= top_k) ? 1.0f : 0.0f; /* :239-243 */
    *overlap = position > 0 ? 1.0f : 0.0f;                       /* :246 */
}

/* ------------------------------------------------------------------------- */
/* Synthetic store: the value distribution of the reference's fixture         */
/* generate_mock_embeddings (src/utils.rs:14-23, uniform(-1,1), DIMENSION=768) */
/* from a counter-based generator so the GPU can regenerate any row.          */
/*   elem(seed,row,col) = u24(splitmix64(splitmix64(seed ^ row*PHI) + col))    */
/*                        * 2^-23 - 1   (exact in binary32, in [-1,1))         */
/* The same function is implemented on the device in csrc/bsr_kernels.cu.     */
/* ------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

BSR_EXPORT void bsr_oracle_generate_rows(uint64_t seed, uint64_t first_row, uint64_t n_rows,
                                         uint32_t dim, float *out)
{
    for (uint64_t r = 0; r < n_rows; r++) {
        uint64_t rk = splitmix64(seed ^ ((first_row + r) * 0x9E3779B97F4A7C15ull));
        float *o = out + r * (uint64_t)dim;
        for (uint32_t c = 0; c < dim; c++) {
            uint32_t u = (uint32_t)(splitmix64(rk + c) >> 40);
            o[c] = (float)u * 0x1p-23f - 1.0f;
        }
    }
}

/* compute_local_top_k (src/mpi_helpers/metrics.rs:36-50) over generator rows
 * [first_row, first_row + n_rows) WITHOUT materialising them: rows are produced
 * 2048 at a time, every row's distance is kept, then the same stable sort and
 * truncate.  Indices are first_row + i, i.e. global generator rows.  This is what
 * lets bench.py check a 12.5 M-row shard in full with a few MB of host memory.
 * bf16 != 0: rows are rounded to bfloat16 first (the BSR_BF16 store's content). */
BSR_EXPORT void bsr_oracle_round_bf16(const float *in, uint64_t n, float *out);

BSR_EXPORT uint64_t bsr_oracle_scan_generated(uint64_t seed, uint64_t first_row, uint64_t n_rows,
                                              uint32_t dim, int bf16, int metric, uint64_t top_k,
                                              const float *target, uint32_t target_len,
                                              uint64_t *out_idx, float *out_dist)
{
    if (n_rows == 0)
        return 0;
    const uint64_t chunk = 2048;
    float *buf = (float *)malloc(chunk * (uint64_t)dim * sizeof(float));
    pair_t *d = (pair_t *)malloc(n_rows * sizeof(pair_t));
    for (uint64_t lo = 0; lo < n_rows; lo += chunk) {
        uint64_t n = n_rows - lo < chunk ? n_rows - lo : chunk;
        bsr_oracle_generate_rows(seed, first_row + lo, n, dim, buf);
        if (bf16)
            bsr_oracle_round_bf16(buf, n * (uint64_t)dim, buf);
        for (uint64_t i = 0; i < n; i++) {
            const float *row = buf + i * (uint64_t)dim;
            d[lo + i].idx = first_row + lo + i;
            d[lo + i].dist = metric ? bsr_oracle_dot_distance(row, dim, target, target_len)
                                    : bsr_oracle_cosine_distance(row, dim, target, target_len);
        }
    }
    stable_sort_pairs_ord(d, n_rows, metric != 0);
    uint64_t keep = n_rows > top_k ? top_k : n_rows;
    for (uint64_t i = 0; i < keep; i++) {
        out_idx[i] = d[i].idx;
        out_dist[i] = d[i].dist;
    }
    free(d);
    free(buf);
    return keep;
}

/* round-to-nearest-even f32 -> bf16 -> f32, the quantisation the bf16 store applies */
BSR_EXPORT void bsr_oracle_round_bf16(const float *in, uint64_t n, float *out)
{
    for (uint64_t i = 0; i < n; i++) {
        uint32_t b;
        memcpy(&b, &in[i], 4);
        if ((b & 0x7F800000u) == 0x7F800000u && (b & 0x007FFFFFu)) {
            b |= 0x00400000u; /* quiet NaN */
            b &= 0xFFFF0000u;
        } else {
            b += 0x7FFFu + ((b >> 16) & 1u);
            b &= 0xFFFF0000u;
        }
        memcpy(&out[i], &b, 4);
    }
}

BSR_EXPORT int bsr_oracle_version(void) { return 2; }
