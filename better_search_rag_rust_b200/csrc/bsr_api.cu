// bsr_api.cu -- host side of the C ABI declared in include/bsr.h.
// Plain CUDA runtime; no torch types cross this boundary.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/bsr.h"
#include "bsr_kernels.cuh"
#include "bsr_tensor.cuh"

// the scan kernels are compiled in bsr_inst.cu (one translation unit per group, in parallel)
namespace bsr {
#define BSR_INST(...) extern template __global__ __VA_ARGS__;
#include "bsr_instances.inc"
#undef BSR_INST
}  // namespace bsr


using namespace bsr;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(expr)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(e_ == cudaErrorMemoryAllocation ? BSR_ERR_OOM : BSR_ERR_CUDA,     \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                        \
    } while (0)

#define CHECK(cond, code, ...) \
    do {                       \
        if (!(cond)) return fail(code, __VA_ARGS__); \
    } while (0)

#define TRY(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_ != BSR_OK) return rc_; \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Ablation switches and test hooks (last section).  Read ONCE, when the library is first
// used; none is needed in production and every default is the measured optimum.
// ---------------------------------------------------------------------------------------------
struct Knobs {
    int scan_ldg = 0;        // BSR_SCAN_VARIANT=ldg: register-staged scan kernel instead of the bulk-async ring
    int ldg_u = 0;           // BSR_LDG_U
    int bulk_stages = 0;     // BSR_BULK_STAGES (0 = default)
    int max_qb = 8;          // BSR_MAX_QB
    int no_wide = 0;         // BSR_NO_WIDE
    int no_fuse_wide = 0;    // BSR_NO_FUSE_WIDE
    int max_grid = 0;        // BSR_MAX_GRID (test hook, 0 = no cap)
    int tensor_min_batch = -1;  // BSR_TENSOR_MIN_BATCH
    int tensor_mode = -1;    // BSR_TENSOR_MODE
    int tensor_stages = 0;   // BSR_TENSOR_STAGES
    int tensor_stage_kb = 0; // BSR_TENSOR_STAGE_KB
    int tc_stats = 0;        // BSR_TC_STATS
    int wide = -1;           // BSR_WIDE: 16-warp CTAs: -1 auto (small f32 stores, every bf16 store), 0 never, 1 always
    int no_tickets = 0;      // BSR_NO_TICKETS: static panel deal on large stores too
    int no_f32_tensor = 0;   // BSR_NO_F32_TENSOR: fp32 stores answer every batch with the exact scan (no bf16 shadow)
    int no_param_query = 0;  // BSR_NO_PARAM_QUERY: bsr_search always copies the query and waits on the stream
    int no_pdl = 0;          // BSR_NO_PDL: plain stream-ordered launches for back-to-back scans
    int no_mailbox = 0;      // BSR_NO_MAILBOX: cross-rank exchange through ncclAllGather only
    int no_small = 0;        // BSR_NO_SMALL: stores of one panel per warp through the streaming kernel too
    int no_list_ll = 0;      // BSR_NO_LIST_LL: scan_small_kernel reduces its CTA lists behind a last-CTA ticket
    int no_l2_keep = 0;      // BSR_NO_L2_KEEP: small stores are read with evict-first like large ones
};
static const Knobs& knobs() {
    static const Knobs k = [] {
        Knobs v;
        auto geti = [](const char* name, int dflt) {
            const char* e = getenv(name);
            return e ? atoi(e) : dflt;
        };
        if (const char* e = getenv("BSR_SCAN_VARIANT")) v.scan_ldg = strcmp(e, "ldg") == 0;
        v.ldg_u = geti("BSR_LDG_U", 0);
        v.bulk_stages = geti("BSR_BULK_STAGES", 0);
        v.max_qb = std::max(1, geti("BSR_MAX_QB", 8));
        v.no_wide = getenv("BSR_NO_WIDE") != nullptr;
        v.no_fuse_wide = getenv("BSR_NO_FUSE_WIDE") != nullptr;
        v.max_grid = std::max(0, geti("BSR_MAX_GRID", 0));
        v.tensor_min_batch = geti("BSR_TENSOR_MIN_BATCH", -1);
        v.tensor_mode = geti("BSR_TENSOR_MODE", -1);
        v.tensor_stages = geti("BSR_TENSOR_STAGES", 0);
        v.tensor_stage_kb = geti("BSR_TENSOR_STAGE_KB", 0);
        v.tc_stats = getenv("BSR_TC_STATS") != nullptr;
        v.wide = geti("BSR_WIDE", -1);
        v.no_tickets = getenv("BSR_NO_TICKETS") != nullptr;
        v.no_f32_tensor = getenv("BSR_NO_F32_TENSOR") != nullptr;
        v.no_param_query = getenv("BSR_NO_PARAM_QUERY") != nullptr;
        v.no_pdl = getenv("BSR_NO_PDL") != nullptr;
        v.no_mailbox = getenv("BSR_NO_MAILBOX") != nullptr;
        v.no_small = getenv("BSR_NO_SMALL") != nullptr;
        v.no_list_ll = getenv("BSR_NO_LIST_LL") != nullptr;
        v.no_l2_keep = getenv("BSR_NO_L2_KEEP") != nullptr;
        return v;
    }();
    return k;
}

// ---------------------------------------------------------------------------------------------
// the store handle
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct bsr_store {
    int device = 0;
    u32 dim = 0, dq = 0;  // dq: query quads (f32: ceil(dim/4); bf16: 16 * kb)
    u32 kb = 0;          // bf16 only: 64-column blocks per row
    bsr_dtype dtype = BSR_F32;
    bsr_metric metric = BSR_COSINE;
    Knobs kn = knobs();  // environment defaults, overridable per handle with bsr_store_set_option
    u64 row_base = 0;
    u64 n_rows = 0;
    u64 cap_rows = 0;  // multiple of 32
    float4* panels = nullptr;
    float* norms = nullptr;
    float* inv_norms = nullptr;  // 1/|a| (0 where |a| is 0/inf/NaN): the tensor path's cheap score scale
    cudaStream_t stream = nullptr;
    int n_sm = 0;
    size_t smem_optin = 0;
    size_t l2_bytes = 0;

    // per-search scratch (grown on demand)
    DevBuf cta_ll;  // scan_small_kernel's self-validating CTA lists
    u32 list_seq = 0;
    DevBuf cta_keys, merge_a, merge_b, d_queries, d_cand, d_result, d_counter, d_gtau, d_surv, d_gather, staging, ones,
        shadow, d_maxnorm, d_query1;
    // fp32 stores: bf16 copy of the rows in the tensor path's tile layout, the FILTER operand of batched
    // searches (built on the first such search, extended after appends); rows [0, shadow_rows) are valid
    u64 shadow_rows = 0;
    float max_norm = -1.0f;  // largest |a| of the shard (BSR_DOT filter margin); < 0: not computed yet
    // batches at least this large use the tcgen05 filter (0 = never).  One pass of it reads 1.5 KB per row
    // whatever the batch (<= 128 queries); the exact scan reads 3 KB (fp32) per row per group of <= 8
    // queries and is issue-bound from 4 queries up: measured on 1 M x 768, 8 queries 0.94 ms exact vs
    // 0.36 ms through the filter.
    u32 tensor_min_batch = 8;
    void* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    // upload pipeline of bsr_store_append: two pinned host blocks + the two halves of `staging`, so the
    // host-side copy of block i+1 overlaps the H2D transfer and the re-layout kernel of block i
    void* h_stage[2] = {nullptr, nullptr};
    size_t h_stage_bytes = 0;
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};

    // set by bsr_search around ONE search_device_impl call: the host query to pass in the kernel's
    // parameter block, and the pinned-host flag the last CTA publishes behind the results
    const float* param_query = nullptr;
    u64* host_ll = nullptr;   // FinalOut::host_ll of the launch being set up
    u32 ll_k = 0;
    u32 ll_seq = 0;
    const Mailbox* mailbox = nullptr;  // set by bsr_search_parallel around one fused launch

    u64 launches = 0;
    // All calls on a handle share one scratch set (cta_keys, merge buffers, the last-CTA ticket).  The
    // asynchronous entry points (bsr_search_device / bsr_merge_device / bsr_unpack_device) run on a
    // caller-supplied stream: each records ev_busy when it has queued its work, and the next call on
    // ANOTHER stream (the handle's own included) waits on it first, so two calls never overlap.
    cudaEvent_t ev_busy = nullptr;
    cudaStream_t busy_stream = nullptr;
    bool busy = false;
    cudaEvent_t ev_scan0 = nullptr, ev_scan1 = nullptr;
    bool scan_timed = false;
    bool scan_time_in_ll = false;  // the last scan published its own duration (FinalOut::t_start)
    u32 last_scan_ns = 0;

    // optional per-launch timing of the scan kernel (bsr_store_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;  // pairs: [2i] before, [2i+1] after
    size_t prof_used = 0;              // pairs recorded and not yet folded into the totals
    double prof_ms = 0.0;
    u64 prof_scans = 0;
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static int ensure(DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return BSR_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
    size_t want = std::max(bytes, (size_t)256);
    CU(cudaMalloc(&b.p, want));
    b.bytes = want;
    return BSR_OK;
}

// results up to this size are written by the kernels directly into pinned host memory (bsr_search)
constexpr size_t kDirectResultBytes = 4096;

static int ensure_pinned(bsr_store* s, size_t bytes) {
    if (s->h_pinned_bytes >= bytes) return BSR_OK;
    if (s->h_pinned) CU(cudaFreeHost(s->h_pinned));
    s->h_pinned = nullptr;
    s->h_pinned_bytes = 0;
    CU(cudaMallocHost(&s->h_pinned, bytes));
    memset(s->h_pinned, 0, bytes);
    s->h_pinned_bytes = bytes;
    return BSR_OK;
}

// see bsr_store::ev_busy
static int order_after_previous(bsr_store* s, cudaStream_t st) {
    if (s->busy && s->busy_stream != st) CU(cudaStreamWaitEvent(st, s->ev_busy, 0));
    return BSR_OK;
}
static int mark_busy(bsr_store* s, cudaStream_t st) {
    CU(cudaEventRecord(s->ev_busy, st));
    s->busy_stream = st;
    s->busy = true;
    return BSR_OK;
}
static void mark_idle(bsr_store* s) { s->busy = false; }  // after a host synchronisation of the stream last used

static u32 next_pow2(u32 v) {
    u32 p = 1;
    while (p < v) p <<= 1;
    return p;
}

static bool is_bf16(const bsr_store* s) { return s->dtype == BSR_BF16; }
// allocation unit: one 32-row panel (f32) or one 64-row tile (bf16)
static u32 unit_rows(const bsr_store* s) { return is_bf16(s) ? 64u : 32u; }
static size_t panel_bytes(const bsr_store* s) {
    return is_bf16(s) ? (size_t)s->kb * kTileBytes : (size_t)s->dq * 32 * sizeof(float4);
}
static size_t used_units(const bsr_store* s, u64 rows) { return (size_t)((rows + unit_rows(s) - 1) / unit_rows(s)); }

// grow the panel/norm allocation to hold at least `rows` rows, preserving contents
static int reserve_rows(bsr_store* s, u64 rows) {
    if (rows <= s->cap_rows) return BSR_OK;
    u64 want = std::max(rows, s->cap_rows + s->cap_rows / 2);
    const u64 ur = unit_rows(s);
    want = (want + ur - 1) / ur * ur;
    const size_t pb = panel_bytes(s);
    float4* np = nullptr;
    float* nn = nullptr;
    CU(cudaMalloc((void**)&np, (size_t)(want / ur) * pb));
    float* ni = nullptr;
    cudaError_t e = cudaMalloc((void**)&nn, (size_t)want * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ni, (size_t)want * sizeof(float));
    if (e != cudaSuccess) {
        cudaFree(np);
        if (nn) cudaFree(nn);
        return fail(BSR_ERR_OOM, "cudaMalloc(norms) failed: %s", cudaGetErrorString(e));
    }
    CU(cudaMemsetAsync(nn, 0, (size_t)want * sizeof(float), s->stream));
    CU(cudaMemsetAsync(ni, 0, (size_t)want * sizeof(float), s->stream));
    const size_t used_panels = used_units(s, s->n_rows);
    if (used_panels) {
        CU(cudaMemcpyAsync(np, s->panels, used_panels * pb, cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaMemcpyAsync(nn, s->norms, (size_t)s->n_rows * sizeof(float),
                           cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaMemcpyAsync(ni, s->inv_norms, (size_t)s->n_rows * sizeof(float),
                           cudaMemcpyDeviceToDevice, s->stream));
    }
    // rows past n_rows inside a used panel are already zero; zero everything newly allocated
    CU(cudaMemsetAsync((char*)np + used_panels * pb, 0, (size_t)(want / ur - used_panels) * pb,
                       s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (s->panels) CU(cudaFree(s->panels));
    if (s->norms) CU(cudaFree(s->norms));
    if (s->inv_norms) CU(cudaFree(s->inv_norms));
    s->panels = np;
    s->norms = nn;
    s->inv_norms = ni;
    s->cap_rows = want;
    return BSR_OK;
}

// ---------------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------------
extern "C" int bsr_version(void) { return BSR_VERSION; }
extern "C" const char* bsr_last_error(void) { return g_err; }

extern "C" int bsr_device_count(int* n) {
    CHECK(n, BSR_ERR_INVALID, "bsr_device_count: null output");
    *n = 0;
    CU(cudaGetDeviceCount(n));
    return BSR_OK;
}

// ---------------------------------------------------------------------------------------------
// store
// ---------------------------------------------------------------------------------------------
extern "C" int bsr_store_create(int device, uint32_t dim, bsr_dtype dtype, uint64_t row_base,
                                uint64_t capacity_rows, bsr_store** out) {
    CHECK(out, BSR_ERR_INVALID, "bsr_store_create: null output");
    *out = nullptr;
    CHECK(dim > 0, BSR_ERR_INVALID, "bsr_store_create: dim must be > 0");
    CHECK(dtype == BSR_F32 || dtype == BSR_BF16, BSR_ERR_INVALID, "bsr_store_create: bad dtype");
    CHECK(row_base < 0xFFFFFFFFull, BSR_ERR_INVALID, "bsr_store_create: row_base must fit 32 bits");
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    CHECK(device >= 0 && device < n_dev, BSR_ERR_CUDA, "bsr_store_create: no CUDA device %d (%d visible)",
          device, n_dev);
    DeviceGuard g(device);
    bsr_store* s = new (std::nothrow) bsr_store();
    CHECK(s, BSR_ERR_OOM, "bsr_store_create: host allocation failed");
    s->device = device;
    s->dim = dim;
    s->dtype = dtype;
    s->kb = (dim + 63) / 64;
    s->dq = dtype == BSR_BF16 ? s->kb * 16 : (dim + 3) / 4;
    s->row_base = row_base;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_busy, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev_scan0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev_scan1);
    if (e != cudaSuccess) {
        delete s;
        return fail(BSR_ERR_CUDA, "bsr_store_create: %s", cudaGetErrorString(e));
    }
    if (s->kn.tensor_min_batch >= 0) s->tensor_min_batch = (u32)s->kn.tensor_min_batch;
    s->n_sm = prop.multiProcessorCount;
    s->smem_optin = prop.sharedMemPerBlockOptin;
    s->l2_bytes = (size_t)prop.l2CacheSize;
    if (capacity_rows) {
        int rc = reserve_rows(s, capacity_rows);
        if (rc != BSR_OK) {
            bsr_store_destroy(s);
            return rc;
        }
    }
    *out = s;
    return BSR_OK;
}

extern "C" int bsr_store_destroy(bsr_store* s) {
    if (!s) return BSR_OK;
    DeviceGuard g(s->device);
    if (s->busy) cudaEventSynchronize(s->ev_busy);
    cudaStreamSynchronize(s->stream);
    DevBuf* bufs[] = {&s->cta_ll, &s->cta_keys, &s->merge_a, &s->merge_b, &s->d_queries,
                      &s->d_cand,   &s->d_result, &s->d_counter, &s->d_gtau, &s->d_surv, &s->d_gather, &s->staging, &s->ones,
                      &s->shadow,   &s->d_maxnorm, &s->d_query1};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (s->panels) cudaFree(s->panels);
    if (s->norms) cudaFree(s->norms);
    if (s->inv_norms) cudaFree(s->inv_norms);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    for (int i = 0; i < 2; i++) {
        if (s->h_stage[i]) cudaFreeHost(s->h_stage[i]);
        if (s->ev_stage[i]) cudaEventDestroy(s->ev_stage[i]);
    }
    for (cudaEvent_t e : s->prof_ev) cudaEventDestroy(e);
    if (s->ev_busy) cudaEventDestroy(s->ev_busy);
    if (s->ev_scan0) cudaEventDestroy(s->ev_scan0);
    if (s->ev_scan1) cudaEventDestroy(s->ev_scan1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return BSR_OK;
}

extern "C" int bsr_store_reset(bsr_store* s) {
    CHECK(s, BSR_ERR_INVALID, "bsr_store_reset: null store");
    DeviceGuard g(s->device);
    if (s->n_rows) {
        CU(cudaMemsetAsync(s->panels, 0, used_units(s, s->n_rows) * panel_bytes(s), s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    s->n_rows = 0;
    s->shadow_rows = 0;
    s->max_norm = -1.0f;
    return BSR_OK;
}

static int check_room(bsr_store* s, u64 n_rows) {
    CHECK(s->row_base + s->n_rows + n_rows < 0xFFFFFFFFull, BSR_ERR_RANGE,
          "append: global row index would exceed 32 bits (row_base %llu + %llu rows)",
          (unsigned long long)s->row_base, (unsigned long long)(s->n_rows + n_rows));
    return BSR_OK;
}

static int finish_append(bsr_store* s, u64 n_rows) {
    const u32 lo = (u32)s->n_rows, hi = (u32)(s->n_rows + n_rows);
    const u32 span = hi - (lo & ~31u);
    if (is_bf16(s))
        row_norms_bf16_kernel<<<(hi - lo + 127) / 128, 128, 0, s->stream>>>(
            (const unsigned char*)s->panels, s->norms, s->inv_norms, lo, hi, s->kb);
    else
        row_norms_kernel<<<(span + 255) / 256, 256, 0, s->stream>>>(s->panels, s->norms, lo, hi, s->dq);
    s->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    s->n_rows += n_rows;
    s->max_norm = -1.0f;
    return BSR_OK;
}

static int interleave_from_device(bsr_store* s, const float* d_rows, u64 row_lo, u64 n) {
    if (is_bf16(s)) {
        const u32 first_tile = (u32)(row_lo >> 6), last_tile = (u32)((row_lo + n - 1) >> 6);
        const size_t total = (size_t)(last_tile - first_tile + 1) * s->kb * 512;
        interleave_rows_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
            d_rows, (uint4*)s->panels, (u32)row_lo, (u32)n, s->dim, s->kb);
        s->launches++;
        CU(cudaGetLastError());
        return BSR_OK;
    }
    const u32 first_panel = (u32)(row_lo >> 5);
    const u32 last_panel = (u32)((row_lo + n - 1) >> 5);
    const size_t total = (size_t)(last_panel - first_panel + 1) * s->dq * 128;
    const unsigned blocks = (unsigned)((total + 1023) / 1024);
    interleave_rows_kernel<<<blocks, 1024, 0, s->stream>>>(d_rows, (float*)s->panels, (u32)row_lo,
                                                           (u32)n, s->dim, s->dq);
    s->launches++;
    CU(cudaGetLastError());
    return BSR_OK;
}

extern "C" int bsr_store_append_device(bsr_store* s, const float* d_rows, uint64_t n_rows) {
    CHECK(s, BSR_ERR_INVALID, "bsr_store_append_device: null store");
    if (n_rows == 0) return BSR_OK;
    CHECK(d_rows, BSR_ERR_INVALID, "bsr_store_append_device: null rows");
    TRY(check_room(s, n_rows));
    DeviceGuard g(s->device);
    TRY(reserve_rows(s, s->n_rows + n_rows));
    // chunk so the interleave grid stays within 32-bit thread counts
    const u64 chunk = std::max<u64>(32, ((1ull << 30) / ((u64)s->dq * 128)) * 32);
    for (u64 done = 0; done < n_rows; done += chunk) {
        const u64 n = std::min(chunk, n_rows - done);
        TRY(interleave_from_device(s, d_rows + done * s->dim, s->n_rows + done, n));
    }
    return finish_append(s, n_rows);
}

// Upload block: large enough to run PCIe at full rate, small enough to keep the pinned footprint modest.
constexpr size_t kUploadBlockBytes = 32ull << 20;

static int ensure_upload_pipeline(bsr_store* s, size_t block_bytes) {
    if (s->h_stage_bytes >= block_bytes) return BSR_OK;
    for (int i = 0; i < 2; i++) {
        if (s->h_stage[i]) CU(cudaFreeHost(s->h_stage[i]));
        s->h_stage[i] = nullptr;
    }
    s->h_stage_bytes = 0;
    for (int i = 0; i < 2; i++) {
        CU(cudaMallocHost(&s->h_stage[i], block_bytes));
        if (!s->ev_stage[i]) CU(cudaEventCreateWithFlags(&s->ev_stage[i], cudaEventDisableTiming));
    }
    s->h_stage_bytes = block_bytes;
    return BSR_OK;
}

extern "C" int bsr_store_append(bsr_store* s, const float* rows, uint64_t n_rows) {
    CHECK(s, BSR_ERR_INVALID, "bsr_store_append: null store");
    if (n_rows == 0) return BSR_OK;
    CHECK(rows, BSR_ERR_INVALID, "bsr_store_append: null rows");
    TRY(check_room(s, n_rows));
    DeviceGuard g(s->device);
    TRY(order_after_previous(s, s->stream));
    TRY(reserve_rows(s, s->n_rows + n_rows));
    // caller's (pageable) rows -> pinned block -> device staging -> re-layout kernel, two blocks in
    // flight: while block i crosses PCIe and is interleaved into the panels, the host fills block i+1
    const size_t row_bytes = (size_t)s->dim * sizeof(float);
    const u64 block_rows = std::max<u64>(1, std::min<u64>(n_rows, kUploadBlockBytes / row_bytes));
    const size_t block_bytes = (size_t)block_rows * row_bytes;
    TRY(ensure_upload_pipeline(s, block_bytes));
    TRY(ensure(s->staging, 2 * block_bytes));
    bool used[2] = {false, false};
    int b = 0;
    for (u64 done = 0; done < n_rows; done += block_rows, b ^= 1) {
        const u64 n = std::min(block_rows, n_rows - done);
        if (used[b]) CU(cudaEventSynchronize(s->ev_stage[b]));  // block b's previous contents are on the device
        memcpy(s->h_stage[b], rows + done * s->dim, (size_t)n * row_bytes);
        float* d_block = (float*)((char*)s->staging.p + (size_t)b * block_bytes);
        CU(cudaMemcpyAsync(d_block, s->h_stage[b], (size_t)n * row_bytes, cudaMemcpyHostToDevice, s->stream));
        CU(cudaEventRecord(s->ev_stage[b], s->stream));  // the pinned block may be refilled from here on
        TRY(interleave_from_device(s, d_block, s->n_rows + done, n));
        used[b] = true;
    }
    return finish_append(s, n_rows);  // norms of the new rows; synchronises the stream
}

extern "C" int bsr_store_append_synthetic(bsr_store* s, uint64_t seed, uint64_t first_row,
                                          uint64_t n_rows) {
    CHECK(s, BSR_ERR_INVALID, "bsr_store_append_synthetic: null store");
    if (n_rows == 0) return BSR_OK;
    TRY(check_room(s, n_rows));
    DeviceGuard g(s->device);
    TRY(reserve_rows(s, s->n_rows + n_rows));
    const u64 chunk = std::max<u64>(32, ((1ull << 30) / ((u64)s->dq * 32)) * 32);
    for (u64 done = 0; done < n_rows; done += chunk) {
        const u64 n = std::min(chunk, n_rows - done);
        const u64 lo = s->n_rows + done;
        if (is_bf16(s)) {
            const u32 first_tile = (u32)(lo >> 6), last_tile = (u32)((lo + n - 1) >> 6);
            const size_t total = (size_t)(last_tile - first_tile + 1) * s->kb * 512;
            synth_rows_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
                (uint4*)s->panels, (u32)lo, (u32)n, s->dim, s->kb, seed, first_row + done);
            s->launches++;
            CU(cudaGetLastError());
            continue;
        }
        const u32 first_panel = (u32)(lo >> 5), last_panel = (u32)((lo + n - 1) >> 5);
        const size_t total = (size_t)(last_panel - first_panel + 1) * s->dq * 32;
        synth_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
            s->panels, (u32)lo, (u32)n, s->dim, s->dq, seed, first_row + done);
        s->launches++;
        CU(cudaGetLastError());
    }
    return finish_append(s, n_rows);
}

extern "C" int bsr_store_count(const bsr_store* s, uint64_t* n) {
    CHECK(s && n, BSR_ERR_INVALID, "bsr_store_count: null argument");
    *n = s->n_rows;
    return BSR_OK;
}
extern "C" int bsr_store_dim(const bsr_store* s, uint32_t* dim) {
    CHECK(s && dim, BSR_ERR_INVALID, "bsr_store_dim: null argument");
    *dim = s->dim;
    return BSR_OK;
}
extern "C" int bsr_store_row_base(const bsr_store* s, uint64_t* row_base) {
    CHECK(s && row_base, BSR_ERR_INVALID, "bsr_store_row_base: null argument");
    *row_base = s->row_base;
    return BSR_OK;
}
extern "C" int bsr_store_device(const bsr_store* s, int* device) {
    CHECK(s && device, BSR_ERR_INVALID, "bsr_store_device: null argument");
    *device = s->device;
    return BSR_OK;
}
extern "C" int bsr_store_scan_bytes(const bsr_store* s, uint64_t* bytes) {
    CHECK(s && bytes, BSR_ERR_INVALID, "bsr_store_scan_bytes: null argument");
    *bytes = is_bf16(s) ? s->n_rows * (u64)s->kb * 128 : s->n_rows * (u64)s->dq * sizeof(float4);
    return BSR_OK;
}

extern "C" int bsr_store_get_many(const bsr_store* cs, uint64_t offset, uint64_t length,
                                  float* out_rows) {
    bsr_store* s = const_cast<bsr_store*>(cs);
    CHECK(s, BSR_ERR_INVALID, "bsr_store_get_many: null store");
    CHECK(offset <= s->n_rows && length <= s->n_rows - offset, BSR_ERR_RANGE,
          "Index not found: rows [%llu, %llu) outside a store of %llu rows",
          (unsigned long long)offset, (unsigned long long)(offset + length),
          (unsigned long long)s->n_rows);
    if (length == 0) return BSR_OK;
    CHECK(out_rows, BSR_ERR_INVALID, "bsr_store_get_many: null output");
    DeviceGuard g(s->device);
    const size_t row_bytes = (size_t)s->dim * sizeof(float);
    const u64 chunk_rows = std::max<u64>(1, std::min<u64>(length, (256ull << 20) / row_bytes));
    TRY(order_after_previous(s, s->stream));
    TRY(ensure(s->staging, (size_t)chunk_rows * row_bytes));
    for (u64 done = 0; done < length; done += chunk_rows) {
        const u64 n = std::min(chunk_rows, length - done);
        const size_t total = (size_t)n * s->dim;
        if (is_bf16(s))
            gather_rows_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
                (const unsigned char*)s->panels, (float*)s->staging.p, (u32)(offset + done), (u32)n,
                s->dim, s->kb);
        else
            gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>(
                (const float*)s->panels, (float*)s->staging.p, (u32)(offset + done), (u32)n, s->dim,
                s->dq);
        s->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out_rows + done * s->dim, s->staging.p, total * sizeof(float),
                           cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return BSR_OK;
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
extern "C" uint32_t bsr_candidate_slots(uint32_t top_k) {
    if (top_k == 0 || top_k > BSR_MAX_TOP_K) return 0;
    return std::max<u32>(32, next_pow2(top_k));
}

template <int U>
static int launch_scan(bsr_store* s, const ScanArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
    static thread_local size_t configured[64] = {0};  // per device
    if (smem > 48 * 1024 && configured[s->device & 63] < smem) {
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, scan_topk_f32_kernel<U>));
        CU(cudaFuncSetAttribute(scan_topk_f32_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(s->smem_optin - fa.sharedSizeBytes)));
        configured[s->device & 63] = s->smem_optin;
    }
    scan_topk_f32_kernel<U><<<grid, kScanThreads, smem, st>>>(a);
    s->launches++;
    CU(cudaGetLastError());
    return BSR_OK;
}

// |b| of a host query exactly as the reference computes it (src/metrics.rs:155,159): a left-to-right
// binary32 fold of separately rounded squares, then sqrt.  `volatile` keeps every intermediate a
// binary32 value and the multiply and the add two operations whatever the host compiler's
// contraction / excess-precision settings are.
static float host_query_magnitude(const float* q, u32 dim) {
    volatile float acc = 0.0f;
    for (u32 i = 0; i < dim; i++) {
        volatile float sq = q[i] * q[i];
        acc = acc + sq;
    }
    return sqrtf(acc);
}

#ifdef BSR_TC_TIMELINE
// host-side stamps of one-launch searches (debug build, BSR_HOST_TL=1): averages every 1000 calls
#include <time.h>
static double g_htl[8];
static unsigned g_htl_n;
static inline double htl_now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static double g_htl_t[8];
#define HTL(i) do { g_htl_t[i] = htl_now(); } while (0)
static void htl_commit() {
    static const bool on = getenv("BSR_HOST_TL") != nullptr;
    if (!on) return;
    for (int i = 1; i < 6; i++) g_htl[i] += g_htl_t[i] - g_htl_t[i - 1];
    if (++g_htl_n == 1000) {
        fprintf(stderr, "host timeline (us, mean of 1000): entry->impl %.2f, impl->launch %.2f, launch call %.2f, launch->result %.2f, result->return %.2f\n",
                g_htl[1] / 1000, g_htl[2] / 1000, g_htl[3] / 1000, g_htl[4] / 1000, g_htl[5] / 1000);
        for (double& v : g_htl) v = 0;
        g_htl_n = 0;
    }
}
#else
#define HTL(i) do { } while (0)
static void htl_commit() {}
#endif

static const size_t kProfPairs = 1024;

static int prof_fold(bsr_store* s) {
    for (size_t i = 0; i < s->prof_used; i++) {
        float ms = 0.f;
        CU(cudaEventSynchronize(s->prof_ev[2 * i + 1]));
        CU(cudaEventElapsedTime(&ms, s->prof_ev[2 * i], s->prof_ev[2 * i + 1]));
        s->prof_ms += ms;
        s->prof_scans++;
    }
    s->prof_used = 0;
    return BSR_OK;
}

static int prof_begin(bsr_store* s, cudaStream_t st) {
    if (!s->profiling) return BSR_OK;
    if (s->prof_used == kProfPairs) TRY(prof_fold(s));
    while (s->prof_ev.size() < 2 * (s->prof_used + 1)) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        s->prof_ev.push_back(e);
    }
    CU(cudaEventRecord(s->prof_ev[2 * s->prof_used], st));
    return BSR_OK;
}

static int prof_end(bsr_store* s, cudaStream_t st) {
    if (!s->profiling) return BSR_OK;
    CU(cudaEventRecord(s->prof_ev[2 * s->prof_used + 1], st));
    s->prof_used++;
    return BSR_OK;
}

static int scan_variant(const bsr_store* s) {
    // BSR_SCAN_VARIANT=ldg selects the register-staged kernel (ablation); default: bulk-async ring
    return s->kn.scan_ldg ? 0 : 1;
}

static size_t bulk_smem_bytes(const bsr_store* s, u32 K, u32 stages, u32 qb, u32 active_warps = kBulkWarps,
                              u32 W = kBulkWarps) {
    return (size_t)active_warps * stages * kChunkBytes + (size_t)qb * s->dq * sizeof(float4) +
           (size_t)qb * W * K * sizeof(u64) + (size_t)W * stages * sizeof(u64);
}

// Ring depth per warp.  Large stores: 3 stages x 4 KiB x 8 warps = 96 KiB in flight per SM measured
// fastest on B200 (deeper rings lose bandwidth: 6 stages -7 %).  Small stores (at most two panels
// per warp) are latency-bound -- a warp's panel is 24 chunks, i.e. 8 dependent round trips through a
// 3-stage ring -- so they get as deep a ring as fits, up to a whole panel in flight per warp; the
// ring is only allocated for the warps that own a panel.  BSR_BULK_STAGES overrides for tuning.
static u32 bulk_stages(const bsr_store* s, u32 K, u32 qb, u32 active_warps = kBulkWarps, bool small = false,
                       u32 W = kBulkWarps) {
    u32 st = small ? 24 : 3;
    if (s->kn.bulk_stages) st = (u32)std::max(2, s->kn.bulk_stages);
    while (st >= 2 && bulk_smem_bytes(s, K, st, qb, active_warps, W) + 1024 > s->smem_optin) st--;
    return st;  // < 2: does not fit
}

// largest query-batch width (8, 4, 2, 1) whose lists fit next to a >= 2-stage ring
static u32 bulk_max_qb(const bsr_store* s, u32 K) {
    const u32 cap = (u32)s->kn.max_qb;
    for (u32 qb = 8; qb >= 2; qb >>= 1)
        if (qb <= cap && bulk_stages(s, K, qb) >= 2) return qb;
    return 1;
}

// The part of a store that asks to stay in L2 between searches (make_keep_fraction_policy): all of it up
// to 60 % of L2, that many bytes beyond.  Measured on B200 (126 MB of L2), 50 k x 768 fp32 = 153 MB, kernel
// time per search: 30 % of L2 22.9 us, 50 % 20.9, 65 % 20.6, 75 % 26.6, 85 % 28.9, 100 % 29.5, none
// (evict-first) 29.0.  Stores above 1 GB gain less than the noise and keep the plain evict-first stream.
static float l2_keep_fraction(const bsr_store* s, u32 n_panels) {
    if (s->kn.no_l2_keep || !s->l2_bytes) return 0.0f;
    // (n_panels counts 32-row panels, also for a bf16 store whose allocation unit is a 64-row tile)
    const double store_bytes = (double)n_panels * 32.0 * (is_bf16(s) ? (double)s->kb * 128.0 : (double)s->dq * 16.0);
    const double keep = std::min(1.0, 0.60 * (double)s->l2_bytes / store_bytes);
    return store_bytes > 1e9 ? 0.0f : (float)keep;
}

template <int QB, bool BF16, int W = kBulkWarps, bool K32 = false>
static int launch_scan_bulk_t(bsr_store* s, const ScanArgs& a, unsigned grid, u32 stages, u32 active_warps,
                              cudaStream_t st, u32* panel_ticket = nullptr) {
    if constexpr (QB == 1 && !K32) {  // top_k <= 32, one query: the K-specialised instantiation
        if (a.k_pad == 32) return launch_scan_bulk_t<1, BF16, W, true>(s, a, grid, stages, active_warps, st, panel_ticket);
    }
    static thread_local bool configured[64] = {false};
    if (!configured[s->device & 63]) {
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, scan_topk_bulk_kernel<QB, BF16, W, K32>));
        CU(cudaFuncSetAttribute(scan_topk_bulk_kernel<QB, BF16, W, K32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(s->smem_optin - fa.sharedSizeBytes)));
        configured[s->device & 63] = true;
    }
    BulkScanArgs ba;
    ba.s = a;
    ba.stages = stages;
    ba.active_warps = active_warps;
    ba.interleave = 1u;  // panels dealt over CTAs first: the ring is sized for active_warps on that assumption
    ba.panel_ticket = panel_ticket;
    ba.l2_keep = l2_keep_fraction(s, a.n_panels);
    ba.query_in_params = 0;
    if (QB == 1 && s->param_query) {  // single host query: rides in the parameter block (no H2D copy)
        ba.query_in_params = 1;
        memcpy(ba.qparam, s->param_query, (size_t)s->dim * sizeof(float));
        ba.qparam_mag = host_query_magnitude(s->param_query, s->dim);
    }
    scan_topk_bulk_kernel<QB, BF16, W, K32>
        <<<grid, W * 32, bulk_smem_bytes(s, a.k_pad, stages, QB, active_warps, W), st>>>(ba);
    s->launches++;
    CU(cudaGetLastError());
    return BSR_OK;
}

// Stores of at most one panel per warp, fp32, one host query in the parameter block: scan_small_kernel,
// 8 warps per CTA up to 8 panels per SM, 16 beyond.  Ring depth: the whole panel per live warp when that
// fits, else as deep as fits.
static bool small_scan_applies(const bsr_store* s, u32 n_panels, unsigned grid, u32 K, u32* stages, u32* live,
                               u32* warps) {
    if (s->kn.no_small || is_bf16(s) || (s->param_query && s->dim > kParamQueryMax)) return false;
    if (s->dq % kChunkQuads != 0 || n_panels > grid * 16u) return false;
    const u32 W = n_panels <= grid * (u32)kBulkWarps ? (u32)kBulkWarps : 16u;
    const u32 cpp = s->dq / kChunkQuads;
    const u32 lv = (n_panels + grid - 1) / grid;
    if (!s->param_query && lv == W) return false;  // a device query needs an idle warp for |b|
    u32 st = cpp;
    while (st >= 2 && bulk_smem_bytes(s, K, st, 1, lv, W) + 1024 > s->smem_optin) st--;
    if (st < 2 || lv * st > W * 32u) return false;
    *stages = st;
    *live = lv;
    *warps = W;
    return true;
}
template <int W, bool WHOLE, int NB>
static int launch_scan_small_t(bsr_store* s, const ScanArgs& a, unsigned grid, u32 stages, u32 live, cudaStream_t st) {
    static thread_local bool configured[64] = {false};
    if (!configured[s->device & 63]) {
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, scan_small_kernel<W, WHOLE, NB>));
        CU(cudaFuncSetAttribute(scan_small_kernel<W, WHOLE, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(s->smem_optin - fa.sharedSizeBytes)));
        configured[s->device & 63] = true;
    }
    BulkScanArgs ba;
    ba.s = a;
    ba.stages = stages;
    ba.active_warps = live;
    ba.interleave = 1u;
    ba.panel_ticket = nullptr;
    // |b|: by the CTAs' idle warp when there is one (every CTA has at most `live` panels), else here
    const bool mag_on_device = live < (u32)W;
    ba.query_in_params = !s->param_query ? 0 : mag_on_device ? 2 : 1;
    // self-validating CTA lists (no last-CTA ticket): register epilogue whose fold is a tree
    ba.l2_keep = l2_keep_fraction(s, a.n_panels);
    ba.list_seq = 0;
    ba.cta_ll = nullptr;
    if (NB != 0 && NB <= 2 && !s->kn.no_list_ll) {  // (NB = 4: measured no faster than the ticket)
        const size_t need = (size_t)grid * a.k_pad * sizeof(ulonglong2);
        if (s->cta_ll.bytes < need) {  // fresh memory must not look like a published list
            TRY(ensure(s->cta_ll, need));
            CU(cudaMemsetAsync(s->cta_ll.p, 0, s->cta_ll.bytes, st));
        }
        s->list_seq = s->list_seq + 1 ? s->list_seq + 1 : 1;
        ba.list_seq = s->list_seq;
        ba.cta_ll = (ulonglong2*)s->cta_ll.p;
    }
    if (s->param_query) {
        memset(ba.qparam, 0, sizeof(ba.qparam));
        memcpy(ba.qparam, s->param_query, (size_t)s->dim * sizeof(float));
    }
    ba.qparam_mag = mag_on_device ? 0.0f : host_query_magnitude(s->param_query, s->dim);
    HTL(2);
    scan_small_kernel<W, WHOLE, NB><<<grid, W * 32, bulk_smem_bytes(s, a.k_pad, stages, 1, live, W), st>>>(ba);
    HTL(3);
    s->launches++;
    CU(cudaGetLastError());
    return BSR_OK;
}
static int launch_scan_small(bsr_store* s, const ScanArgs& a, unsigned grid, u32 stages, u32 live, u32 warps,
                             cudaStream_t st) {
    // register epilogue: k_pad <= 128 with the final stage fused into the launch, at most 20 CTA lists per
    // warp of the last CTA (16-warp CTAs: 10, and k_pad <= 64 -- registers)
    const bool regs = a.fuse_final && a.k_pad <= 128 && grid <= 160u;
    if (warps == 16) {
        if (regs && a.k_pad == 32) return launch_scan_small_t<16, false, 1>(s, a, grid, stages, live, st);
        if (regs && a.k_pad == 64) return launch_scan_small_t<16, false, 2>(s, a, grid, stages, live, st);
        return launch_scan_small_t<16, false, 0>(s, a, grid, stages, live, st);
    }
    const bool whole = stages == s->dq / kChunkQuads;
    if (regs && a.k_pad == 32)
        return whole ? launch_scan_small_t<8, true, 1>(s, a, grid, stages, live, st) : launch_scan_small_t<8, false, 1>(s, a, grid, stages, live, st);
    if (regs && a.k_pad == 64)
        return whole ? launch_scan_small_t<8, true, 2>(s, a, grid, stages, live, st) : launch_scan_small_t<8, false, 2>(s, a, grid, stages, live, st);
    if (regs && a.k_pad == 128)
        return whole ? launch_scan_small_t<8, true, 4>(s, a, grid, stages, live, st) : launch_scan_small_t<8, false, 4>(s, a, grid, stages, live, st);
    return whole ? launch_scan_small_t<8, true, 0>(s, a, grid, stages, live, st) : launch_scan_small_t<8, false, 0>(s, a, grid, stages, live, st);
}

template <int QB>
static int launch_scan_bulk(bsr_store* s, const ScanArgs& a, unsigned grid, u32 stages, u32 active_warps,
                            cudaStream_t st, u32* panel_ticket) {
    return is_bf16(s) ? launch_scan_bulk_t<QB, true>(s, a, grid, stages, active_warps, st, panel_ticket)
                      : launch_scan_bulk_t<QB, false>(s, a, grid, stages, active_warps, st, panel_ticket);
}

// reduce n_lists lists per query to one; returns the buffer holding [nq][K] in *result
static int merge_tree(bsr_store* s, const u64* in, u32 n_lists, u32 nq, u32 K, u32 k, size_t list_stride,
                      size_t query_stride, u64* final_out, cudaStream_t st) {
    static thread_local bool configured[64] = {false};
    if (!configured[s->device & 63]) {
        CU(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kMergeMaxKeys * sizeof(u64))));
        configured[s->device & 63] = true;
    }
    const u32 max_per_cta = std::max<u32>(2, kMergeMaxKeys / K);
    u64* ping = (u64*)s->merge_a.p;
    u64* pong = (u64*)s->merge_b.p;
    const u64* cur = in;
    while (true) {
        const u32 per_cta = std::min(max_per_cta, std::max<u32>(2, next_pow2(n_lists)));
        const u32 groups = (n_lists + per_cta - 1) / per_cta;
        u64* dst = (groups == 1) ? final_out : ping;
        merge_lists_kernel<<<dim3(groups, nq), kMergeThreads, (size_t)per_cta * K * sizeof(u64), st>>>(
            cur, n_lists, K, per_cta, list_stride, query_stride, dst, groups == 1 ? k : K);
        s->launches++;
        CU(cudaGetLastError());
        if (groups == 1) break;
        cur = dst;
        n_lists = groups;
        list_stride = K;
        query_stride = (size_t)groups * K;
        std::swap(ping, pong);
    }
    return BSR_OK;
}

static unsigned scan_grid(const bsr_store* s, u32 n_panels, int warps_per_cta) {
    const unsigned need = (n_panels + warps_per_cta - 1) / warps_per_cta;
    return std::max(1u, std::min<unsigned>((unsigned)s->n_sm, need));
}

// ---- tcgen05 path (BF16 stores, large batches) -------------------------------------------------
// Ring shape (measured on B200, 10 M x 768 bf16, top-100): every stage costs a fixed wait/arm/copy +
// wait/issue/commit handshake, so stages are large.  Single CTA: 3 x 48 KiB.  Cluster of 2: one
// whole tile (96 KiB) per stage, 2 stages.  BSR_TENSOR_STAGE_KB / BSR_TENSOR_STAGES override.
static u32 tensor_stage_kb(const bsr_store* s, u32 mode) {
    u32 v = mode == 0 ? kTcStageKB / 2 : kTcStageKB;  // pair mode: 12 half-height blocks = 48 KiB
    if (s->kn.tensor_stage_kb) v = (u32)std::min<int>(kTcStageKB, std::max(1, s->kn.tensor_stage_kb));
    return v;
}
static size_t tensor_smem_bytes(u32 stage_kb, u32 stages, u32 mode = 0) {
    const size_t blk = mode == 2 ? kTileBytes / 2 : kTileBytes;
    return (size_t)stages * stage_kb * blk + (size_t)kTcNormSlots * 2 * kTcN * sizeof(float) +
           (size_t)(3 * stages + 5 + 2 * kTcNormSlots) * sizeof(u64) + 16 + 1024;  // + ring alignment slack
}
static u32 tensor_stages(const bsr_store* s, u32 mode) {
    u32 st = mode == 1 ? 2 : 3;
    if (s->kn.tensor_stages) st = (u32)std::max(2, s->kn.tensor_stages);
    while (st >= 2 && tensor_smem_bytes(tensor_stage_kb(s, mode), st, mode) + 2048 > s->smem_optin) st--;
    return st;
}
static bool tensor_eligible(const bsr_store* s, u32 nq, u32 K) {
    return (is_bf16(s) || !s->kn.no_f32_tensor) && s->tensor_min_batch > 0 && nq >= s->tensor_min_batch &&
           s->kb <= kTcMaxKB && K <= BSR_MAX_TOP_K && tensor_stages(s, 0) >= 2 && tensor_stages(s, 1) >= 2 &&
           tensor_stages(s, 2) >= 2;
}

// fp32 store: bring the bf16 shadow (and 1/|a|) up to date with the rows
static int ensure_shadow(bsr_store* s, cudaStream_t st) {
    if (is_bf16(s) || s->shadow_rows == s->n_rows) return BSR_OK;
    const size_t tiles_cap = (size_t)((s->cap_rows + 63) / 64);
    const size_t need = tiles_cap * s->kb * kTileBytes;
    if (s->shadow.bytes < need) {  // (re)allocated: everything is converted again
        TRY(ensure(s->shadow, need));
        s->shadow_rows = 0;
    }
    const u32 lo = (u32)s->shadow_rows, hi = (u32)s->n_rows;
    const u32 first_tile = lo >> 6, last_tile = (hi - 1) >> 6;
    const size_t total = (size_t)(last_tile - first_tile + 1) * s->kb * 512;
    const size_t per_tile = (size_t)s->kb * 512;
    const size_t step = (((size_t)1 << 30) / per_tile) * per_tile;  // whole tiles, grid within 32-bit thread counts
    for (size_t done = 0; done < total; done += step) {
        const size_t n = std::min(total - done, step);
        const u32 tile0 = first_tile + (u32)(done / per_tile);
        shadow_from_panels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            s->panels, (uint4*)s->shadow.p, s->norms, s->inv_norms, done ? tile0 * 64 : lo, hi, s->dim, s->dq, s->kb);
        s->launches++;
        CU(cudaGetLastError());
    }
    s->shadow_rows = s->n_rows;
    return BSR_OK;
}

// largest |a| of the shard (finite norms only), for the BSR_DOT filter margin
static int ensure_max_norm(bsr_store* s, cudaStream_t st) {
    if (s->max_norm >= 0.0f) return BSR_OK;
    TRY(ensure(s->d_maxnorm, 256));
    CU(cudaMemsetAsync(s->d_maxnorm.p, 0, 4, st));
    max_norm_kernel<<<std::min<unsigned>(1024, (unsigned)((s->n_rows + 255) / 256)), 256, 0, st>>>(
        s->norms, (u32)s->n_rows, (u32*)s->d_maxnorm.p);
    s->launches++;
    CU(cudaGetLastError());
    float m = 0.0f;
    CU(cudaMemcpyAsync(&m, s->d_maxnorm.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    s->max_norm = m;
    return BSR_OK;
}

template <int MODE>
static int launch_tensor(bsr_store* s, const TensorScanArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
    static thread_local bool configured[64] = {false};
    if (!configured[s->device & 63]) {
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, scan_topk_bf16_tensor_kernel<MODE>));
        CU(cudaFuncSetAttribute(scan_topk_bf16_tensor_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(s->smem_optin - fa.sharedSizeBytes)));
        configured[s->device & 63] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MODE == 0 ? 1 : 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CU(cudaLaunchKernelEx(&cfg, scan_topk_bf16_tensor_kernel<MODE>, a));
    s->launches++;
    return BSR_OK;
}

static int search_tensor(bsr_store* s, const float* d_queries, u32 nq, u32 k, u64* d_cand, u64* d_idx,
                         float* d_dist, u32* d_count, cudaStream_t st) {
    const u32 K = bsr_candidate_slots(k);
    const u32 n_tiles = (u32)((s->n_rows + kTcN - 1) / kTcN);
    TRY(ensure_shadow(s, st));
    // Filter margin (TensorScanArgs::margin): twice the worst-case gap between the filter's dot product
    // and the exact one, relative to |a| |b|.  Query rounded to bf16: 2^-9 per element; an fp32 store's
    // shadow rows another 2^-9; 1e-4 covers both accumulations (768 * 2^-24 for the sequential fold,
    // less for the tensor core) and the second-order terms.
    float margin = 2.0f * ((is_bf16(s) ? 1.0f : 2.0f) * 0.001953125f + 1e-4f);
    if (s->metric == BSR_DOT) {  // scores are not divided by |a|: scale by the largest one
        TRY(ensure_max_norm(s, st));
        margin *= s->max_norm;
    }
    const u32 surv_cap = std::max<u32>(kTcSurvivorCap, 2 * K);
    // more than one 128-query group: clusters of 2 CTAs share every store tile (TMA multicast)
    // mode 0: single CTAs; 1: clusters of 2 with multicast tiles; 2: CTA pairs (cta_group::2)
    u32 mode = (nq > kTcQ && s->n_sm >= 2) ? 2u : 0u;  // measured: pairs (2) beat multicast (1) by ~6 %
    if (s->kn.tensor_mode >= 0) mode = (u32)std::min(2, s->kn.tensor_mode);
    const u32 C = mode == 0 ? 1u : 2u;
    unsigned n_clusters = std::max(1u, std::min<unsigned>((unsigned)s->n_sm / C, n_tiles));
    // BSR_MAX_GRID: test hook (see search_device_impl) -- few clusters make every survivor list long
    if (s->kn.max_grid) n_clusters = std::max(1u, std::min<unsigned>(n_clusters, (unsigned)s->kn.max_grid / C));
    const unsigned grid = n_clusters * C;
    const u32 QS = C * kTcQ;  // queries per launch
    const u32 stages = tensor_stages(s, mode), stage_kb = tensor_stage_kb(s, mode);
    const size_t smem = tensor_smem_bytes(stage_kb, stages, mode);
    const size_t gm_bytes = (size_t)QS * ((k + kTcN - 1) / kTcN) * kTcN * sizeof(u32);
    TRY(ensure(s->d_gtau, gm_bytes));
    const size_t surv_bytes = (size_t)QS * n_clusters * 2 * surv_cap * sizeof(u64);
    TRY(ensure(s->d_surv, surv_bytes + (size_t)QS * n_clusters * sizeof(u32) + 2 * QS * sizeof(float)));
    if (s->metric == BSR_DOT && !s->ones.p) {
        TRY(ensure(s->ones, kTcN * sizeof(float)));
        float h_ones[kTcN];
        for (u32 i = 0; i < kTcN; i++) h_ones[i] = 1.0f;
        CU(cudaMemcpyAsync(s->ones.p, h_ones, sizeof h_ones, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));  // h_ones is a stack buffer
    }
    CU(cudaEventRecord(s->ev_scan0, st));
    for (u32 q0 = 0; q0 < nq; q0 += QS) {
        TensorScanArgs a;
        a.tiles = is_bf16(s) ? (const unsigned char*)s->panels : (const unsigned char*)s->shadow.p;
        a.exact_panels = is_bf16(s) ? nullptr : s->panels;
        a.exact_dq = s->dq;
        a.margin = margin;
        a.surv_cap = surv_cap;
        a.norms = s->norms;
        a.inv_norms = s->inv_norms;
        a.queries = d_queries + (size_t)q0 * s->dim;
        a.cta_keys = nullptr;
        a.g_gm = (u32*)s->d_gtau.p;
        a.survivors = (u64*)s->d_surv.p;
        a.surv_count = (u32*)((char*)s->d_surv.p + surv_bytes);
        a.q_mag = (float*)(a.surv_count + (size_t)QS * n_clusters);
        a.nq_valid = std::min(QS, nq - q0);
        a.n_rows = (u32)s->n_rows;
        a.n_tiles = n_tiles;
        a.KB = s->kb;
        a.dim = s->dim;
        a.k = k;
        a.k_pad = K;
        a.row_base = (u32)s->row_base;
        a.stages = stages;
        a.stage_kb = stage_kb;
        a.metric = (u32)s->metric;
        a.norm_stride = kTcN;
        if (s->metric == BSR_DOT) {  // scores are ranked as they are: a constant block of ones stands in for 1/|a|
            a.inv_norms = (const float*)s->ones.p;
            a.norm_stride = 0;
        }
        a.timeline = nullptr;
        a.tl_start = 0;
#ifdef BSR_TC_TIMELINE
        static long long* d_tl = nullptr;
        const size_t tl_bytes = (size_t)2 * kTlTiles * 8 * sizeof(long long);
        const bool tl_on = getenv("BSR_TC_TL") && q0 == 0;
        if (tl_on) {
            if (!d_tl) CU(cudaMalloc(&d_tl, tl_bytes));
            CU(cudaMemsetAsync(d_tl, 0, tl_bytes, st));
            a.timeline = d_tl;
            a.tl_start = (u32)atoi(getenv("BSR_TC_TL"));
        }
#endif
        CU(cudaMemsetAsync(s->d_gtau.p, 0, gm_bytes, st));  // 0 < ordered(-inf): "nobody published yet"
        TRY(prof_begin(s, st));
        if (mode == 2) TRY(launch_tensor<2>(s, a, grid, smem, st));
        else if (mode == 1) TRY(launch_tensor<1>(s, a, grid, smem, st));
        else TRY(launch_tensor<0>(s, a, grid, smem, st));
        {
            // per query: final threshold, exact keys of the survivors near it, the query's final list
            const size_t sel_smem = (size_t)K * sizeof(u64) + std::max((size_t)kThrSmemValues * sizeof(u32), (size_t)kSelChunk * sizeof(u64));
            static thread_local bool sel_configured[64] = {false};
            if (!sel_configured[s->device & 63]) {
                CU(cudaFuncSetAttribute(select_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(BSR_MAX_TOP_K * sizeof(u64) + kThrSmemValues * sizeof(u32))));
                sel_configured[s->device & 63] = true;
            }
            select_rescore_kernel<<<a.nq_valid, kSelThreads, sel_smem, st>>>(a, n_clusters, d_cand + (size_t)q0 * K);
            s->launches++;
            CU(cudaGetLastError());
        }
        TRY(prof_end(s, st));
        if (s->kn.tc_stats && q0 == 0) {  // debug: survivors per (query, cluster) list
            std::vector<u32> h((size_t)QS * n_clusters);
            CU(cudaStreamSynchronize(st));
            CU(cudaMemcpy(h.data(), a.surv_count, h.size() * sizeof(u32), cudaMemcpyDeviceToHost));
            unsigned long long sum = 0;
            u32 mx = 0;
            const size_t cnt = (size_t)a.nq_valid * n_clusters;
            for (size_t i = 0; i < cnt; i++) {
                sum += h[i];
                mx = std::max(mx, h[i]);
            }
            fprintf(stderr, "tensor stats: %zu lists, survivors mean %.1f max %u\n", cnt, (double)sum / cnt, mx);
        }
#ifdef BSR_TC_TIMELINE
        if (tl_on) {  // debug only: dump the pipeline stamps of CTAs 0 and 1 (cycles since CTA 0's first stamp)
            std::vector<long long> h(2 * kTlTiles * 8);
            CU(cudaStreamSynchronize(st));
            CU(cudaMemcpy(h.data(), d_tl, tl_bytes, cudaMemcpyDeviceToHost));
            for (int c = 0; c < 2; c++) {
                const long long t0 = h[(size_t)c * kTlTiles * 8];
                fprintf(stderr, "cta %d tile: prod_slot_free mma_acc_free mma_data_ok mma_issued | epi_wait epi_acc_ready "
                                "epi_ld_done epi_done\n", c);
                for (u32 t = 0; t < kTlTiles; t++) {
                    fprintf(stderr, "%3u:", t);
                    for (int e = 0; e < 8; e++) {
                        const long long v = h[((size_t)c * kTlTiles + t) * 8 + e];
                        fprintf(stderr, " %8lld", v ? v - t0 : -1);
                    }
                    fprintf(stderr, "\n");
                }
            }
        }
#endif
    }
    CU(cudaEventRecord(s->ev_scan1, st));
    s->scan_timed = true;
    if (d_idx) {
        unpack_kernel<<<nq, 128, 0, st>>>(d_cand, K, k, d_idx, d_dist, d_count);
        s->launches++;
        CU(cudaGetLastError());
    }
    return BSR_OK;
}

// Scan + merge for nq queries.  Always writes the merged candidates to d_cand[nq][K]; when d_idx
// is non-null also the unpacked idx[nq][k], dist[nq][k], count[nq].
static int search_device_impl(bsr_store* s, const float* d_queries, u32 nq, u32 k, u64* d_cand,
                              u64* d_idx, float* d_dist, u32* d_count, cudaStream_t st) {
    const u32 K = bsr_candidate_slots(k);
    if (tensor_eligible(s, nq, K)) {
        // in slices that keep the per-cluster candidate lists (slice x clusters x K keys) within 1 GiB
        const u32 per = std::max<u32>(256, (u32)((((size_t)1 << 30) / ((size_t)s->n_sm * K * sizeof(u64))) / 256 * 256));
        for (u32 q0 = 0; q0 < nq; q0 += per) {
            const u32 n = std::min(per, nq - q0);
            TRY(search_tensor(s, d_queries + (size_t)q0 * s->dim, n, k, d_cand + (size_t)q0 * K,
                              d_idx ? d_idx + (size_t)q0 * k : nullptr, d_idx ? d_dist + (size_t)q0 * k : nullptr,
                              d_idx ? d_count + q0 : nullptr, st));
        }
        return BSR_OK;
    }
    const u32 n_panels = (u32)((s->n_rows + 31) / 32);
    const bool bulk = (scan_variant(s) == 1 || is_bf16(s)) && bulk_stages(s, K, 1) >= 2;
    CHECK(bulk || !is_bf16(s), BSR_ERR_UNSUPPORTED, "search: BF16 stores need the bulk-async scan kernel");
    // bulk kernel: one CTA per SM, or one per panel when the store has fewer panels than SMs
    // BSR_MAX_GRID: test hook -- fewer CTAs give every warp many panels, which is what exercises the
    // steady-state (post-fill) top-k insertion on stores small enough for the CPU oracle
    unsigned grid_cap = (unsigned)s->n_sm;
    if (s->kn.max_grid) grid_cap = std::max(1u, std::min<unsigned>(grid_cap, (unsigned)s->kn.max_grid));
    const unsigned grid = bulk ? std::max(1u, std::min<unsigned>(grid_cap, n_panels))
                               : std::min<unsigned>(grid_cap, scan_grid(s, n_panels, kScanWarps));
    // 16 warps per CTA (single queries): small fp32 stores (more than one, at most four rounds of
    // panels at 8 warps per CTA), where a second round on a few warps would double the latency-bound
    // run time; and EVERY bf16 store -- a bf16 warp needs ~2.7x the instructions per byte (widening) and
    // is bound by its dependent-issue latency at two warps per scheduler (ncu: 0.47 IPC per scheduler,
    // 4.3 cycles between a warp's instructions, 5.0 of 7.2 TB/s): four warps per scheduler hide it.
    const u32 per_cta = (n_panels + grid - 1) / grid;
    const bool wide_fits = bulk && per_cta > (u32)kBulkWarps &&
                           bulk_stages(s, K, 1, std::min<u32>(16, per_cta), per_cta <= 32u, 16) >= 2;
    const bool wide_auto = per_cta <= 4u * kBulkWarps || is_bf16(s);
    const bool wide = wide_fits && !s->kn.no_wide && (s->kn.wide < 0 ? wide_auto : s->kn.wide != 0);
    const u32 active_warps = std::min<u32>(wide ? 16u : (u32)kBulkWarps, per_cta);
    const bool small_store = n_panels <= 2u * grid * (wide ? 16u : (u32)kBulkWarps);
    const size_t smem = (size_t)s->dq * sizeof(float4) + (size_t)kScanWarps * K * sizeof(u64);
    CHECK(bulk || smem + 1024 <= s->smem_optin, BSR_ERR_UNSUPPORTED,
          "search: dim %u with top_k %u needs %zu bytes of shared memory (limit %zu)", s->dim, k,
          smem, s->smem_optin);
    TRY(ensure(s->cta_keys, (size_t)nq * grid * K * sizeof(u64)));
    // k <= 128: the scan's last CTA merges the per-CTA lists and unpacks -> one kernel per launch
    // (the reference's own k is 50, src/main.rs:110).  BSR_NO_FUSE_WIDE=1: k > 32 through the merge tree.
    const bool fuse = (K == 32) || (K <= 128 && !s->kn.no_fuse_wide);
    if (!s->d_counter.p) {
        TRY(ensure(s->d_counter, 256));
        CU(cudaMemsetAsync(s->d_counter.p, 0, 256, st));
    }
    // large stores: panels after a warp's first are dealt by ticket (BulkScanArgs::panel_ticket); the
    // kernel's single-panel lookahead needs a ring no deeper than a panel has chunks
    const u32 cpp = is_bf16(s) ? s->kb : (s->dq + kChunkQuads - 1) / kChunkQuads;
    u32* const ticket = (u32*)s->d_counter.p + 16;
    const u32 max_per_cta = std::max<u32>(2, kMergeMaxKeys / K);
    const size_t inter = (size_t)nq * ((grid + max_per_cta - 1) / max_per_cta) * K * sizeof(u64);
    TRY(ensure(s->merge_a, inter));
    TRY(ensure(s->merge_b, inter));
    const u32 max_qb = bulk ? bulk_max_qb(s, K) : 1;

    // one-launch searches time themselves (FinalOut::t_start): no events in their stream
    const bool ll_timed = fuse && nq == 1 && s->host_ll != nullptr;
    if (!ll_timed) CU(cudaEventRecord(s->ev_scan0, st));
    for (u32 q = 0; q < nq;) {
        // queries are scanned in groups that share one pass over the store
        u32 qb = max_qb;
        while (qb > nq - q) qb >>= 1;
        ScanArgs a;
        a.panels = s->panels;
        a.norms = s->norms;
        a.query = d_queries + (size_t)q * s->dim;
        a.cta_keys = (u64*)s->cta_keys.p + (size_t)q * grid * K;
        a.n_rows = (u32)s->n_rows;
        a.n_panels = n_panels;
        a.dq = s->dq;
        a.dim = s->dim;
        a.k = k;
        a.k_pad = K;
        a.row_base = (u32)s->row_base;
        a.fuse_final = fuse ? 1u : 0u;
        a.metric = (u32)s->metric;
        a.fin.done_counter = (u32*)s->d_counter.p;
        a.fin.keys = d_cand + (size_t)q * K;
        a.fin.idx = d_idx ? d_idx + (size_t)q * k : nullptr;
        a.fin.dist = d_idx ? d_dist + (size_t)q * k : nullptr;
        a.fin.count = d_idx ? d_count + q : nullptr;
        a.fin.host_ll = (fuse && nq == 1) ? s->host_ll : nullptr;
        a.fin.ll_k = s->ll_k;
        a.fin.ll_seq = s->ll_seq;
        a.fin.t_start = a.fin.host_ll ? (unsigned long long*)((char*)s->d_counter.p + 128) : nullptr;
        a.fin.mb = Mailbox{};
        if (s->mailbox && fuse && nq == 1) a.fin.mb = *s->mailbox;
        a.dbg = a.fin.dbg = nullptr;
#ifdef BSR_TC_TIMELINE
        static unsigned long long* d_dbg = nullptr;
        static unsigned long long* h_dbg = nullptr;  // pinned, also stamped by the host
        const bool dbg_on = getenv("BSR_SCAN_TL") != nullptr;
        if (dbg_on) {
            if (!d_dbg) {
                CU(cudaMalloc(&d_dbg, 64 * sizeof(unsigned long long)));
                CU(cudaMallocHost(&h_dbg, 64 * sizeof(unsigned long long)));
            }
            CU(cudaMemsetAsync(d_dbg, 0, 64 * sizeof(unsigned long long), st));
            CU(cudaStreamSynchronize(st));
            a.dbg = a.fin.dbg = d_dbg;
        }
#endif
        TRY(prof_begin(s, st));
        u32 small_stages = 0, small_live = 0, small_warps = 0;
        if (bulk && qb == 1 && small_scan_applies(s, n_panels, grid, K, &small_stages, &small_live, &small_warps)) {
            TRY(launch_scan_small(s, a, grid, small_stages, small_live, small_warps, st));
        } else if (bulk) {
            if (qb == 1 && wide) {
                const u32 stages = bulk_stages(s, K, 1, active_warps, small_store, 16);
                u32* const tk = (!small_store && !s->kn.no_tickets && stages <= cpp) ? ticket : nullptr;
                if (is_bf16(s)) TRY((launch_scan_bulk_t<1, true, 16>(s, a, grid, stages, active_warps, st, tk)));
                else TRY((launch_scan_bulk_t<1, false, 16>(s, a, grid, stages, active_warps, st, tk)));
                TRY(prof_end(s, st));
                q += qb;
                continue;
            }
            const u32 aw8 = std::min<u32>(kBulkWarps, active_warps);
            const bool small8 = n_panels <= 2u * grid * kBulkWarps;
            const u32 stages = bulk_stages(s, K, qb, aw8, small8);
            u32* const tk = (!small8 && !s->kn.no_tickets && stages <= cpp) ? ticket : nullptr;
            if (qb == 8) TRY(launch_scan_bulk<8>(s, a, grid, stages, aw8, st, tk));
            else if (qb == 4) TRY(launch_scan_bulk<4>(s, a, grid, stages, aw8, st, tk));
            else if (qb == 2) TRY(launch_scan_bulk<2>(s, a, grid, stages, aw8, st, tk));
            else TRY(launch_scan_bulk<1>(s, a, grid, stages, aw8, st, tk));
        } else if (s->dq % 16 == 0 && s->kn.ldg_u == 16)
            TRY(launch_scan<16>(s, a, grid, smem, st));
        else if (s->dq % 8 == 0 && s->kn.ldg_u != 4)
            TRY(launch_scan<8>(s, a, grid, smem, st));
        else if (s->dq % 4 == 0) TRY(launch_scan<4>(s, a, grid, smem, st));
        else if (s->dq % 2 == 0) TRY(launch_scan<2>(s, a, grid, smem, st));
        else TRY(launch_scan<1>(s, a, grid, smem, st));
        TRY(prof_end(s, st));
#ifdef BSR_TC_TIMELINE
        if (dbg_on) {
            CU(cudaStreamSynchronize(st));
            CU(cudaMemcpy(h_dbg, d_dbg, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            static const char* names[] = {"cta0 start", "cta0 staged", "cta0 sync", "cta0 first chunk", "cta0 panels done",
                                          "cta0 topk finish", "cta0 lists published", "cta0 ticket", "last: start",
                                          "last: lists folded", "last: done"};
            fprintf(stderr, "scan timeline (ns since cta0 start), grid %u:", grid);
            for (int i = 0; i < 11; i++) fprintf(stderr, " [%s %lld]", names[i], h_dbg[i] ? (long long)(h_dbg[i] - h_dbg[0]) : -1LL);
            if (h_dbg[14]) fprintf(stderr, " [cta0 requests issued %lld] [cta0 query staged %lld]", (long long)(h_dbg[14] - h_dbg[0]), (long long)(h_dbg[15] - h_dbg[0]));
            if (h_dbg[13] > h_dbg[12] && h_dbg[4] > h_dbg[0])
                fprintf(stderr, " [SM clock %.0f MHz over the panel loop]", (double)(h_dbg[13] - h_dbg[12]) * 1e3 / (double)(h_dbg[4] - h_dbg[0]));
            fprintf(stderr, "\n  cta0 warp0 chunks, SM cycles since the first wait (wait->ready):");
            for (int i = 0; i < 24; i++)
                fprintf(stderr, " %u->%u", (unsigned)(h_dbg[16 + 2 * i] - h_dbg[16]), (unsigned)(h_dbg[17 + 2 * i] - h_dbg[16]));
            fprintf(stderr, "\n");
        }
#endif
        q += qb;
    }
    if (!ll_timed) CU(cudaEventRecord(s->ev_scan1, st));
    s->scan_timed = true;
    s->scan_time_in_ll = ll_timed;
    if (fuse) return BSR_OK;
    TRY(merge_tree(s, (const u64*)s->cta_keys.p, grid, nq, K, k, K, (size_t)grid * K, d_cand, st));
    if (d_idx) {
        unpack_kernel<<<nq, 128, 0, st>>>(d_cand, K, k, d_idx, d_dist, d_count);
        s->launches++;
        CU(cudaGetLastError());
    }
    return BSR_OK;
}

static int check_search_args(const bsr_store* s, const void* q, u32 nq, u32 k, const char* fn) {
    CHECK(s, BSR_ERR_INVALID, "%s: null store", fn);
    CHECK(nq == 0 || q, BSR_ERR_INVALID, "%s: null queries", fn);
    CHECK(k >= 1 && k <= BSR_MAX_TOP_K, BSR_ERR_INVALID, "%s: top_k %u outside [1, %u]", fn, k,
          BSR_MAX_TOP_K);
    return BSR_OK;
}

extern "C" int bsr_search_device(bsr_store* s, const float* d_queries, uint32_t nq, uint32_t k,
                                 uint64_t* d_candidates, void* stream) {
    TRY(check_search_args(s, d_queries, nq, k, "bsr_search_device"));
    if (nq == 0) return BSR_OK;
    CHECK(d_candidates, BSR_ERR_INVALID, "bsr_search_device: null output");
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    TRY(order_after_previous(s, st));
    TRY(search_device_impl(s, d_queries, nq, k, (u64*)d_candidates, nullptr, nullptr, nullptr, st));
    return mark_busy(s, st);
}

extern "C" int bsr_merge_device(bsr_store* s, const uint64_t* d_lists, uint32_t n_lists,
                                uint32_t nq, uint32_t k, uint64_t* d_out, void* stream) {
    TRY(check_search_args(s, d_lists, nq, k, "bsr_merge_device"));
    CHECK(n_lists >= 1, BSR_ERR_INVALID, "bsr_merge_device: n_lists must be >= 1");
    if (nq == 0) return BSR_OK;
    CHECK(d_out, BSR_ERR_INVALID, "bsr_merge_device: null output");
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    TRY(order_after_previous(s, st));
    const u32 K = bsr_candidate_slots(k);
    if (K == 32 && n_lists <= 4096) {  // one warp per query folds the lists in registers
        merge32_kernel<<<nq, 32, 0, st>>>((const u64*)d_lists, n_lists, (size_t)nq * K, K, k, (u64*)d_out, nullptr,
                                          nullptr, nullptr);
        s->launches++;
        CU(cudaGetLastError());
        return mark_busy(s, st);
    }
    const u32 max_per_cta = std::max<u32>(2, kMergeMaxKeys / K);
    const size_t inter = (size_t)nq * ((n_lists + max_per_cta - 1) / max_per_cta) * K * sizeof(u64);
    TRY(ensure(s->merge_a, inter));
    TRY(ensure(s->merge_b, inter));
    TRY(merge_tree(s, (const u64*)d_lists, n_lists, nq, K, k, (size_t)nq * K, K, (u64*)d_out, st));
    return mark_busy(s, st);
}

extern "C" int bsr_unpack_device(bsr_store* s, const uint64_t* d_cand, uint32_t nq, uint32_t k,
                                 uint64_t* d_idx, float* d_dist, uint32_t* d_count, void* stream) {
    TRY(check_search_args(s, d_cand, nq, k, "bsr_unpack_device"));
    if (nq == 0) return BSR_OK;
    CHECK(d_idx && d_dist && d_count, BSR_ERR_INVALID, "bsr_unpack_device: null output");
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    TRY(order_after_previous(s, st));
    unpack_kernel<<<nq, 128, 0, st>>>((const u64*)d_cand, bsr_candidate_slots(k), k, (u64*)d_idx,
                                      d_dist, d_count);
    s->launches++;
    CU(cudaGetLastError());
    return mark_busy(s, st);
}

// Host side of FinalOut::host_ll: poll the self-validating result words of search `seq` in the pinned
// block and unpack them.  The count word is written last by the kernel but every word is checked on its
// own.  Every ~1 000 polls the stream is asked as well, so a failed launch or a faulting kernel
// surfaces as an error instead of a hang.
static int wait_host_ll(bsr_store* s, const volatile u64* ll, u32 seq, u32 k, uint64_t* out_idx, float* out_dist,
                        uint32_t* out_count, bool timed) {
    auto wait_word = [&](u32 i, u32* value) -> int {
        for (u32 spins = 0;; spins++) {
            const u64 w = ll[i];
            if ((u32)(w >> 32) == seq) {
                *value = (u32)w;
                return BSR_OK;
            }
            if ((spins & 0x3FF) == 0x3FF) {
                cudaError_t e = cudaStreamQuery(s->stream);
                if (e == cudaSuccess) {
                    const u64 w2 = ll[i];
                    if ((u32)(w2 >> 32) == seq) {
                        *value = (u32)w2;
                        return BSR_OK;
                    }
                    return fail(BSR_ERR_CUDA, "search: the scan finished without publishing its result");
                }
                if (e != cudaErrorNotReady) return fail(BSR_ERR_CUDA, "search: %s", cudaGetErrorString(e));
            }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
    };
    u32 count = 0;
    TRY(wait_word(2 * k, &count));
    HTL(4);
    for (u32 j = 0; j < k; j++) {
        u32 v = 0;
        TRY(wait_word(j, &v));
        out_idx[j] = v;
        TRY(wait_word(k + j, &v));
        memcpy(&out_dist[j], &v, sizeof(float));
    }
    if (out_count) *out_count = count;
    if (timed) TRY(wait_word(2 * k + 1, &s->last_scan_ns));
    return BSR_OK;
}

static u32 next_ll_seq(bsr_store* s) {
    s->ll_seq = s->ll_seq + 1 ? s->ll_seq + 1 : 1;  // never 0: a fresh pinned block reads 0
    return s->ll_seq;
}

extern "C" int bsr_search(bsr_store* s, const float* queries, uint32_t nq, uint32_t k,
                          uint64_t* out_idx, float* out_dist, uint32_t* out_count) {
    HTL(0);
    TRY(check_search_args(s, queries, nq, k, "bsr_search"));
    if (nq == 0) return BSR_OK;
    CHECK(out_idx && out_dist && out_count, BSR_ERR_INVALID, "bsr_search: null output");
    DeviceGuard g(s->device);
    const u32 K = bsr_candidate_slots(k);
    const size_t qbytes = (size_t)nq * s->dim * sizeof(float);
    const size_t ibytes = (size_t)nq * k * sizeof(u64), dbytes = (size_t)nq * k * sizeof(float),
                 cbytes = (size_t)nq * sizeof(u32);
    // one packed result block [idx | dist | count | flag] on the device and in pinned host memory, so a
    // search is at most one H2D copy, the kernels, one D2H copy
    const size_t rbytes = ibytes + dbytes + cbytes;
    const size_t off_r = (qbytes + 15) & ~(size_t)15;
    const size_t off_ll = (off_r + rbytes + 63) & ~(size_t)63;  // one-launch searches: (2k + 1) self-validating words
    TRY(ensure_pinned(s, off_ll + (2 * (size_t)k + 2) * sizeof(u64)));
    TRY(ensure(s->d_queries, qbytes));
    TRY(ensure(s->d_cand, (size_t)nq * K * sizeof(u64)));
    TRY(ensure(s->d_result, rbytes));
    char* h = (char*)s->h_pinned;
    if (s->n_rows == 0) {
        // empty shard: compute_local_top_k returns an empty list (metrics.rs:36-52)
        memset(out_count, 0, cbytes);
        return BSR_OK;
    }
    TRY(order_after_previous(s, s->stream));
    // Small results (a few queries): the kernel that unpacks writes them straight into the pinned
    // host block (mapped into the device's address space under UVA) -- the D2H transfer still
    // happens every call, as posted PCIe writes instead of a separate copy operation.
    const bool direct = rbytes <= kDirectResultBytes;
    // One query, one fused launch (k <= 128, bulk kernel, not the tensor path): the query travels in the
    // kernel's parameter block -- the H2D transfer of its 4*dim bytes is part of the launch -- and
    // the host waits on a flag the last CTA writes behind the results, not on the stream.
    const bool one_launch = direct && nq == 1 && K <= 128 && !s->kn.no_fuse_wide && !s->kn.scan_ldg &&
                            s->dim <= kParamQueryMax && !s->kn.no_param_query && bulk_stages(s, K, 1) >= 2 &&
                            !tensor_eligible(s, 1, K);
    char* r = direct ? h + off_r : (char*)s->d_result.p;
    volatile u64* ll = (volatile u64*)(h + off_ll);
    if (one_launch) {
        s->param_query = queries;
        s->host_ll = (u64*)ll;
        s->ll_k = k;
        next_ll_seq(s);
        for (u32 i = 0; i < 2 * k + 2; i++) ll[i] = 0;  // whatever an earlier call left in these words of the shared block
    } else {
        memcpy(h, queries, qbytes);
    }
    // Batches on the tensor path: a query without a usable direction (zero, denormal-small, NaN or inf
    // magnitude) cannot be bounded by the filter -- every row would survive it and stall its whole
    // launch -- so such queries ride along as a copy of a well-formed one and are answered afterwards,
    // one by one, by the exact scan (same results: both paths end in the reference's arithmetic).
    std::vector<u32> degenerate;
    const u32 saved_min_batch = s->tensor_min_batch;
    if (!one_launch && tensor_eligible(s, nq, K)) {
        u32 good = nq;
        for (u32 q = 0; q < nq; q++) {
            double ss = 0.0;
            for (u32 c = 0; c < s->dim; c++) ss += (double)queries[(size_t)q * s->dim + c] * (double)queries[(size_t)q * s->dim + c];
            if (!(ss >= 1e-10) || !(ss < 1e37)) degenerate.push_back(q);  // zero / tiny / NaN / inf / |b|^2 near FLT_MAX
            else if (good == nq) good = q;
        }
        if (good == nq) {
            s->tensor_min_batch = 0;  // nothing but degenerate queries: the exact scan answers the batch
            degenerate.clear();
        } else {
            for (u32 q : degenerate)
                memcpy(h + (size_t)q * s->dim * sizeof(float), queries + (size_t)good * s->dim, (size_t)s->dim * sizeof(float));
        }
    }
    if (!one_launch) CU(cudaMemcpyAsync(s->d_queries.p, h, qbytes, cudaMemcpyHostToDevice, s->stream));
    HTL(1);
    int rc = search_device_impl(s, (const float*)s->d_queries.p, nq, k, (u64*)s->d_cand.p,
                                one_launch ? nullptr : (u64*)r, (float*)(r + ibytes), (u32*)(r + ibytes + dbytes),
                                s->stream);
    s->param_query = nullptr;
    s->host_ll = nullptr;
    for (size_t i = 0; i < degenerate.size() && rc == BSR_OK; i++) {
        const u32 q = degenerate[i];
        rc = ensure(s->d_query1, (size_t)s->dim * sizeof(float));
        if (rc != BSR_OK) break;
        if (cudaMemcpyAsync(s->d_query1.p, queries + (size_t)q * s->dim, (size_t)s->dim * sizeof(float),
                            cudaMemcpyHostToDevice, s->stream) != cudaSuccess) {
            rc = fail(BSR_ERR_CUDA, "bsr_search: query upload failed");
            break;
        }
        s->tensor_min_batch = 0;
        rc = search_device_impl(s, (const float*)s->d_query1.p, 1, k, (u64*)s->d_cand.p + (size_t)q * K,
                                (u64*)r + (size_t)q * k, (float*)(r + ibytes) + (size_t)q * k,
                                (u32*)(r + ibytes + dbytes) + q, s->stream);
    }
    s->tensor_min_batch = saved_min_batch;
    TRY(rc);
    if (one_launch) {
        TRY(wait_host_ll(s, ll, s->ll_seq, k, out_idx, out_dist, out_count, s->scan_time_in_ll));
        HTL(5);
        htl_commit();
        // the kernel may still be retiring; the next call on this handle is stream-ordered behind it
        mark_idle(s);
        return BSR_OK;
    }
    if (!direct) CU(cudaMemcpyAsync(h + off_r, r, rbytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    mark_idle(s);
    memcpy(out_idx, h + off_r, ibytes);
    memcpy(out_dist, h + off_r + ibytes, dbytes);
    memcpy(out_count, h + off_r + ibytes + dbytes, cbytes);
    return BSR_OK;
}

extern "C" int bsr_merge_host(bsr_store* s, const uint64_t* idx, const float* dist, uint64_t n,
                              uint32_t k, uint64_t* out_idx, float* out_dist, uint32_t* out_count) {
    CHECK(s, BSR_ERR_INVALID, "bsr_merge_host: null store");
    CHECK(k >= 1 && k <= BSR_MAX_TOP_K, BSR_ERR_INVALID, "bsr_merge_host: top_k %u outside [1, %u]", k,
          BSR_MAX_TOP_K);
    CHECK(out_count, BSR_ERR_INVALID, "bsr_merge_host: null output");
    *out_count = 0;
    if (n == 0) return BSR_OK;
    CHECK(idx && dist && out_idx && out_dist, BSR_ERR_INVALID, "bsr_merge_host: null argument");
    CHECK(n <= (1ull << 26), BSR_ERR_RANGE, "bsr_merge_host: too many candidates (%llu)",
          (unsigned long long)n);
    for (u64 i = 0; i < n; i++)
        CHECK(idx[i] < 0xFFFFFFFFull, BSR_ERR_RANGE, "bsr_merge_host: index %llu does not fit 32 bits",
              (unsigned long long)idx[i]);
    // HashSet de-dup of metrics.rs:156-168: after the stable sort the first occurrence of an index
    // wins, i.e. its entry with the smallest (distance, position).  Shards own disjoint row blocks, so
    // in the engine's own use there are no duplicates and this pass only looks; when there are, the
    // losing entries are dropped before the pairs go to the device.
    std::vector<uint64_t> dd_idx;
    std::vector<float> dd_dist;
    {
        std::vector<u64> order(n);
        for (u64 i = 0; i < n; i++) order[i] = i;
        // partial_cmp order of the reference's sort_by (metrics.rs:153); NaN distances compare "equal" there,
        // here they are ordered last so the choice is deterministic
        auto less_dist = [&](u64 a, u64 b) {
            const float x = dist[a], y = dist[b];
            if (x != x || y != y) return (y != y) && !(x != x);
            return x < y;
        };
        std::stable_sort(order.begin(), order.end(), [&](u64 a, u64 b) {
            if (idx[a] != idx[b]) return idx[a] < idx[b];
            return less_dist(a, b);
        });
        bool dup = false;
        for (u64 i = 1; i < n && !dup; i++) dup = idx[order[i]] == idx[order[i - 1]];
        if (dup) {
            std::vector<char> keep(n, 0);
            for (u64 i = 0; i < n; i++)
                if (i == 0 || idx[order[i]] != idx[order[i - 1]]) keep[order[i]] = 1;
            for (u64 i = 0; i < n; i++)
                if (keep[i]) {
                    dd_idx.push_back(idx[i]);
                    dd_dist.push_back(dist[i]);
                }
            idx = dd_idx.data();
            dist = dd_dist.data();
            n = dd_idx.size();
        }
    }
    DeviceGuard g(s->device);
    const u32 K = bsr_candidate_slots(k);
    const u32 n_lists = (u32)((n + K - 1) / K);
    const size_t in_bytes = (size_t)n * (sizeof(u64) + sizeof(float));
    const size_t ibytes = (size_t)k * sizeof(u64), dbytes = (size_t)k * sizeof(float);
    const size_t rbytes = ibytes + dbytes + sizeof(u32);
    TRY(ensure_pinned(s, std::max(in_bytes, rbytes)));
    TRY(ensure(s->staging, in_bytes));
    TRY(ensure(s->cta_keys, (size_t)n_lists * K * sizeof(u64)));
    TRY(ensure(s->d_cand, (size_t)K * sizeof(u64)));
    TRY(ensure(s->d_result, rbytes));
    const u32 max_per_cta = std::max<u32>(2, kMergeMaxKeys / K);
    const size_t inter = (size_t)((n_lists + max_per_cta - 1) / max_per_cta) * K * sizeof(u64);
    TRY(ensure(s->merge_a, inter));
    TRY(ensure(s->merge_b, inter));
    char* h = (char*)s->h_pinned;
    memcpy(h, idx, (size_t)n * sizeof(u64));
    memcpy(h + (size_t)n * sizeof(u64), dist, (size_t)n * sizeof(float));
    TRY(order_after_previous(s, s->stream));
    CU(cudaMemcpyAsync(s->staging.p, h, in_bytes, cudaMemcpyHostToDevice, s->stream));
    const u64* d_i = (const u64*)s->staging.p;
    const float* d_d = (const float*)((char*)s->staging.p + (size_t)n * sizeof(u64));
    pack_sort_lists_kernel<<<n_lists, 256, (size_t)K * sizeof(u64), s->stream>>>(d_i, d_d, n, K,
                                                                                 (u64*)s->cta_keys.p);
    s->launches++;
    CU(cudaGetLastError());
    TRY(merge_tree(s, (const u64*)s->cta_keys.p, n_lists, 1, K, k, K, (size_t)n_lists * K,
                   (u64*)s->d_cand.p, s->stream));
    char* r = (char*)s->d_result.p;
    unpack_kernel<<<1, 128, 0, s->stream>>>((const u64*)s->d_cand.p, K, k, (u64*)r, (float*)(r + ibytes),
                                            (u32*)(r + ibytes + dbytes));
    s->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h, r, rbytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    mark_idle(s);
    u32 cnt;
    memcpy(&cnt, h + ibytes + dbytes, sizeof(u32));
    memcpy(out_idx, h, (size_t)cnt * sizeof(u64));
    memcpy(out_dist, h + ibytes, (size_t)cnt * sizeof(float));
    *out_count = cnt;
    return BSR_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-rank search: NCCL resolved at run time
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclId {
    char internal[BSR_COMM_ID_BYTES];
};
struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
const int kNcclUint64 = 5;  // ncclDataType_t

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // prefer the copy already mapped into the process (e.g. the one torch.distributed uses)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
    api.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy;
    return api;
}
}  // namespace

struct bsr_comm {
    void* comm = nullptr;
    int n_ranks = 0, rank = 0, device = 0;
    // NVLink mailbox (see Mailbox in bsr_kernels.cuh): every rank owns one device block
    // [mailbox entries | ack word]; root reads its own entries and writes the peers' ack words, the
    // peers write root's entries.  Mapped across processes with CUDA IPC at bsr_comm_create.
    bool mb_ok = false;
    void* mb_local = nullptr;                // my block (cudaMalloc)
    void* mb_root = nullptr;                 // non-root: root's block, mapped into this process
    std::vector<void*> mb_peers;             // root: every peer's block (index = rank; [0] unused)
    void* d_ack_ptrs = nullptr;              // root: device array of the peers' ack-word addresses
    u32* h_status = nullptr;                 // pinned: set to 1 by a kernel that gave up waiting
    u32 seq = 0;                             // sequence number of the last mailbox search
    cudaStream_t stream = nullptr;
};

static size_t mailbox_entries_bytes(int n_ranks) {
    return (size_t)kMbDepth * n_ranks * kMbSlots * sizeof(ulonglong2);
}

#define NCCL(expr)                                                                              \
    do {                                                                                        \
        int r_ = (expr);                                                                        \
        if (r_ != 0)                                                                            \
            return fail(BSR_ERR_CUDA, "%s failed: %s", #expr,                                   \
                        nccl().GetErrorString ? nccl().GetErrorString(r_) : "NCCL error");      \
    } while (0)

extern "C" int bsr_comm_unique_id(uint8_t id[BSR_COMM_ID_BYTES]) {
    CHECK(id, BSR_ERR_INVALID, "bsr_comm_unique_id: null output");
    CHECK(nccl().ok, BSR_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "not found");
    NcclId nid;
    NCCL(nccl().GetUniqueId(&nid));
    memcpy(id, nid.internal, BSR_COMM_ID_BYTES);
    return BSR_OK;
}

// What every rank tells the others at bsr_comm_create (one all-gather of this record).
struct MailboxHello {
    cudaIpcMemHandle_t handle;  // of mb_local
    char uuid[16];              // the GPU: two ranks on one GPU must not wait on each other inside kernels
    int ok;                     // local allocation + handle export worked
    int pad[3];
};

// Collective.  Any failure simply leaves mb_ok false on EVERY rank (they all look at the same gathered
// records and exchange their mapping results), and searches use ncclAllGather.
static int mailbox_setup(bsr_comm* c) {
    c->mb_ok = false;
    if (c->n_ranks < 2 || c->n_ranks > 32 || getenv("BSR_NO_MAILBOX")) return BSR_OK;  // decided identically everywhere
    const size_t ebytes = mailbox_entries_bytes(c->n_ranks), bytes = ebytes + 256;
    MailboxHello me;
    memset(&me, 0, sizeof me);
    cudaDeviceProp prop;
    bool ok = cudaGetDeviceProperties(&prop, c->device) == cudaSuccess;
    if (ok) memcpy(me.uuid, &prop.uuid, 16);
    ok = ok && cudaMalloc(&c->mb_local, bytes) == cudaSuccess;
    ok = ok && cudaMemset(c->mb_local, 0, bytes) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&me.handle, c->mb_local) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&c->h_status, 64) == cudaSuccess;
    if (c->h_status) *c->h_status = 0;
    cudaGetLastError();
    me.ok = ok ? 1 : 0;
    // gather the records (device staging: NCCL moves device memory)
    MailboxHello* d_all = nullptr;
    std::vector<MailboxHello> all((size_t)c->n_ranks);
    CU(cudaMalloc((void**)&d_all, sizeof(MailboxHello) * (c->n_ranks + 1)));
    CU(cudaMemcpyAsync(d_all + c->n_ranks, &me, sizeof me, cudaMemcpyHostToDevice, c->stream));
    NCCL(nccl().AllGather(d_all + c->n_ranks, d_all, sizeof(MailboxHello), 0 /* ncclInt8 */, c->comm, c->stream));
    CU(cudaMemcpyAsync(all.data(), d_all, sizeof(MailboxHello) * c->n_ranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    bool all_ok = true;
    for (int r = 0; r < c->n_ranks; r++) {
        all_ok = all_ok && all[r].ok;
        for (int q = 0; q < r; q++) all_ok = all_ok && memcmp(all[r].uuid, all[q].uuid, 16) != 0;
    }
    // map: peers -> root's block; root -> every peer's block
    int mapped = all_ok ? 1 : 0;
    if (all_ok) {
        if (c->rank == 0) {
            c->mb_peers.assign((size_t)c->n_ranks, nullptr);
            std::vector<u32*> acks((size_t)c->n_ranks, nullptr);
            for (int r = 1; r < c->n_ranks && mapped; r++) {
                if (cudaIpcOpenMemHandle(&c->mb_peers[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    c->mb_peers[r] = nullptr;
                    mapped = 0;
                } else {
                    acks[r] = (u32*)((char*)c->mb_peers[r] + ebytes);
                }
            }
            if (mapped && (cudaMalloc(&c->d_ack_ptrs, sizeof(u32*) * c->n_ranks) != cudaSuccess ||
                           cudaMemcpy(c->d_ack_ptrs, acks.data(), sizeof(u32*) * c->n_ranks, cudaMemcpyHostToDevice) != cudaSuccess))
                mapped = 0;
        } else if (cudaIpcOpenMemHandle(&c->mb_root, all[0].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            c->mb_root = nullptr;
            mapped = 0;
        }
        cudaGetLastError();
    }
    // everyone must have succeeded
    int* d_flags = (int*)d_all;
    CU(cudaMemcpyAsync(d_flags + c->n_ranks, &mapped, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NCCL(nccl().AllGather(d_flags + c->n_ranks, d_flags, sizeof(int), 0, c->comm, c->stream));
    std::vector<int> flags((size_t)c->n_ranks);
    CU(cudaMemcpyAsync(flags.data(), d_flags, sizeof(int) * c->n_ranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(d_all));
    bool every = true;
    for (int f : flags) every = every && f;
    c->mb_ok = every;
    return BSR_OK;
}

extern "C" int bsr_comm_create(const uint8_t id[BSR_COMM_ID_BYTES], int n_ranks, int rank, int device,
                               bsr_comm** out) {
    CHECK(out, BSR_ERR_INVALID, "bsr_comm_create: null output");
    *out = nullptr;
    CHECK(id && n_ranks >= 1 && rank >= 0 && rank < n_ranks, BSR_ERR_INVALID, "bsr_comm_create: bad arguments");
    CHECK(nccl().ok, BSR_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    DeviceGuard g(device);
    bsr_comm* c = new (std::nothrow) bsr_comm();
    CHECK(c, BSR_ERR_OOM, "bsr_comm_create: host allocation failed");
    c->n_ranks = n_ranks;
    c->rank = rank;
    c->device = device;
    NcclId nid;
    memcpy(nid.internal, id, BSR_COMM_ID_BYTES);
    int r = nccl().CommInitRank(&c->comm, n_ranks, nid, rank);
    if (r != 0) {
        delete c;
        return fail(BSR_ERR_CUDA, "ncclCommInitRank failed: %s",
                    nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        nccl().CommDestroy(c->comm);
        delete c;
        return fail(BSR_ERR_CUDA, "bsr_comm_create: cudaStreamCreate failed");
    }
    int rc = mailbox_setup(c);
    if (rc != BSR_OK) {
        bsr_comm_destroy(c);
        return rc;
    }
    *out = c;
    return BSR_OK;
}

extern "C" int bsr_comm_destroy(bsr_comm* c) {
    if (!c) return BSR_OK;
    DeviceGuard g(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void* p : c->mb_peers)
        if (p) cudaIpcCloseMemHandle(p);
    if (c->mb_root) cudaIpcCloseMemHandle(c->mb_root);
    if (c->d_ack_ptrs) cudaFree(c->d_ack_ptrs);
    if (c->mb_local) cudaFree(c->mb_local);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BSR_OK;
}

extern "C" int bsr_comm_uses_mailbox(const bsr_comm* c, int* yes) {
    CHECK(c && yes, BSR_ERR_INVALID, "bsr_comm_uses_mailbox: null argument");
    *yes = c->mb_ok ? 1 : 0;
    return BSR_OK;
}

// The Mailbox record of search `seq` for this rank.
static Mailbox mailbox_for(const bsr_comm* c, u32 seq) {
    Mailbox mb{};
    const size_t ebytes = mailbox_entries_bytes(c->n_ranks);
    const size_t slot = (size_t)(seq % kMbDepth) * c->n_ranks * kMbSlots;
    if (c->rank == 0) {
        mb.recv = (const ulonglong2*)c->mb_local + slot;
        mb.ack_remote = (u32* const*)c->d_ack_ptrs;
    } else {
        mb.send = (ulonglong2*)c->mb_root + slot + (size_t)c->rank * kMbSlots;
        mb.ack_local = (const u32*)((const char*)c->mb_local + ebytes);
    }
    mb.status = c->h_status;
    mb.n_ranks = (u32)c->n_ranks;
    mb.seq = seq;
    return mb;
}

// Can this search exchange its candidates through the NVLink mailbox?  Every input is identical on all
// ranks (arguments, dimension, option), so all ranks take the same path.
static bool mailbox_usable(const bsr_store* s, const bsr_comm* c, u32 nq, u32 K) {
    // K <= 128: the scan's last CTA merges (fused final stage); 32 < K: its shared-memory lists must hold one
    // list per rank (8 warps' worth at least)
    return c->mb_ok && !s->kn.no_mailbox && nq == 1 && K <= kMbSlots && (K == 32 || c->n_ranks <= kBulkWarps) &&
           !s->kn.scan_ldg && !s->kn.no_fuse_wide && bulk_stages(s, K, 1) >= 2 && !tensor_eligible(s, 1, K);
}

static u32 next_mailbox_seq(bsr_comm* c) {
    c->seq = c->seq + 1 ? c->seq + 1 : 1;  // never 0: a fresh mailbox reads 0
    return c->seq;
}

// A rank with nothing to scan (empty shard, or a local failure) still takes part: one warp sends an
// empty list / folds the peers' lists and publishes.
static int launch_mailbox_only(bsr_store* s, const Mailbox& mb, u32 k, u64* d_cand, u64* host_ll, u32 ll_k, u32 ll_seq,
                               cudaStream_t st) {
    FinalOut fo{};
    fo.keys = d_cand;
    fo.host_ll = host_ll;
    fo.ll_k = ll_k;
    fo.ll_seq = ll_seq;
    fo.mb = mb;
    mailbox_only_kernel<<<1, 256, 0, st>>>(fo, k, bsr_candidate_slots(k));
    s->launches++;
    CU(cudaGetLastError());
    return BSR_OK;
}

extern "C" int bsr_search_parallel(bsr_store* s, bsr_comm* c, const float* queries, uint32_t nq, uint32_t k,
                                   uint64_t* out_idx, float* out_dist, uint32_t* out_count) {
    // Arguments every rank passes identically (or that are local programming errors) are rejected
    // before any collective work.  Past that point a rank ALWAYS joins the exchange: a local failure
    // (allocation, scan) turns its contribution into an empty list -- what the reference does with a
    // failed local scan (metrics.rs:185-191) -- and the error is returned after the collective, so the
    // other ranks are never left waiting.
    TRY(check_search_args(s, queries, nq, k, "bsr_search_parallel"));
    CHECK(c && c->comm, BSR_ERR_INVALID, "bsr_search_parallel: null communicator");
    CHECK(c->device == s->device, BSR_ERR_INVALID, "bsr_search_parallel: shard and communicator are on different devices");
    const bool root = c->rank == 0;
    if (nq == 0) return BSR_OK;
    CHECK(!root || (out_idx && out_dist && out_count), BSR_ERR_INVALID, "bsr_search_parallel: null output on root");
    DeviceGuard g(s->device);
    const u32 K = bsr_candidate_slots(k);
    const size_t qbytes = (size_t)nq * s->dim * sizeof(float);
    const size_t ibytes = (size_t)nq * k * sizeof(u64), dbytes = (size_t)nq * k * sizeof(float),
                 cbytes = (size_t)nq * sizeof(u32);
    const size_t rbytes = ibytes + dbytes + cbytes, off_r = (qbytes + 15) & ~(size_t)15;
    const size_t off_flag = (off_r + rbytes + 63) & ~(size_t)63;  // mailbox path: (2k + 1) self-validating result words
    const size_t cand_bytes = (size_t)nq * K * sizeof(u64);
    // the two buffers the exchange itself needs; without them this rank cannot take part at all
    TRY(ensure(s->d_cand, cand_bytes));
    TRY(ensure(s->d_gather, cand_bytes * c->n_ranks + cand_bytes));
    TRY(ensure_pinned(s, off_flag + (2 * (size_t)k + 2) * sizeof(u64)));
    TRY(order_after_previous(s, s->stream));
    char* h = (char*)s->h_pinned;

    if (mailbox_usable(s, c, nq, K)) {
        // ---- one fused launch per rank: scan -> (NVLink stores into root's mailbox | root folds them in) ----
        const u32 seq = next_mailbox_seq(c);
        const Mailbox mb = mailbox_for(c, seq);
        *c->h_status = 0;
        volatile u64* ll = (volatile u64*)(h + off_flag);
        const u32 ll_k = root ? k : 0;  // the other ranks only signal completion (the count word)
        const u32 ll_seq = next_ll_seq(s);
        for (u32 i = 0; i < 2 * ll_k + 2; i++) ll[i] = 0;
        int local_rc = BSR_OK;
        if (s->n_rows != 0) {
            const bool in_params = s->dim <= kParamQueryMax && !s->kn.no_param_query;
            if (!in_params) {
                local_rc = ensure(s->d_queries, qbytes);
                if (local_rc == BSR_OK) {
                    memcpy(h, queries, qbytes);
                    if (cudaMemcpyAsync(s->d_queries.p, h, qbytes, cudaMemcpyHostToDevice, s->stream) != cudaSuccess)
                        local_rc = fail(BSR_ERR_CUDA, "bsr_search_parallel: query upload failed");
                }
            }
            if (local_rc == BSR_OK) {
                s->param_query = in_params ? queries : nullptr;
                s->host_ll = (u64*)ll;
                s->ll_k = ll_k;
                s->mailbox = &mb;
                local_rc = search_device_impl(s, (const float*)s->d_queries.p, nq, k, (u64*)s->d_cand.p, nullptr, nullptr,
                                              nullptr, s->stream);
                s->param_query = nullptr;
                s->host_ll = nullptr;
                s->mailbox = nullptr;
            }
        }
        char local_err[sizeof g_err];
        if (local_rc != BSR_OK) memcpy(local_err, g_err, sizeof g_err);
        const bool scanned = s->n_rows != 0 && local_rc == BSR_OK;
        if (!scanned)
            TRY(launch_mailbox_only(s, mb, k, (u64*)s->d_cand.p, (u64*)ll, ll_k, ll_seq, s->stream));
        u32 cnt = 0;
        TRY(wait_host_ll(s, ll, ll_seq, ll_k, out_idx, out_dist, &cnt, scanned && s->scan_time_in_ll));
        mark_idle(s);
        if (local_rc != BSR_OK) {
            memcpy(g_err, local_err, sizeof g_err);
            return local_rc;
        }
        CHECK(*c->h_status == 0, BSR_ERR_CUDA,
              "bsr_search_parallel: a rank did not deliver its candidates (mailbox wait gave up)");
        if (root) *out_count = cnt;
        return BSR_OK;
    }

    // ---- ncclAllGather of the candidate slots + merge launch on root ----
    auto local_scan = [&]() -> int {
        TRY(ensure(s->d_queries, qbytes));
        TRY(ensure(s->d_result, rbytes));
        memcpy(h, queries, qbytes);
        CU(cudaMemcpyAsync(s->d_queries.p, h, qbytes, cudaMemcpyHostToDevice, s->stream));
        if (s->n_rows == 0) {
            // an empty shard contributes an empty list (metrics.rs:36-52): all slots UINT64_MAX
            CU(cudaMemsetAsync(s->d_cand.p, 0xFF, cand_bytes, s->stream));
            return BSR_OK;
        }
        return search_device_impl(s, (const float*)s->d_queries.p, nq, k, (u64*)s->d_cand.p, nullptr, nullptr,
                                  nullptr, s->stream);
    };
    int local_rc = local_scan();
    char local_err[sizeof g_err];
    if (local_rc != BSR_OK) {
        memcpy(local_err, g_err, sizeof g_err);
        // whatever the failed scan left behind must not reach the other ranks
        if (cudaMemsetAsync(s->d_cand.p, 0xFF, cand_bytes, s->stream) != cudaSuccess) cudaGetLastError();
    }
    // rank-ordered concatenation of every rank's [nq][K] slots (displacements, metrics.rs:86-89)
    NCCL(nccl().AllGather(s->d_cand.p, s->d_gather.p, (size_t)nq * K, kNcclUint64, c->comm, s->stream));
    auto root_merge = [&]() -> int {
        u64* merged = (u64*)((char*)s->d_gather.p + cand_bytes * c->n_ranks);
        const u32 max_per_cta = std::max<u32>(2, kMergeMaxKeys / K);
        const size_t inter = (size_t)nq * ((c->n_ranks + max_per_cta - 1) / max_per_cta) * K * sizeof(u64);
        TRY(ensure(s->merge_a, inter));
        TRY(ensure(s->merge_b, inter));
        const bool direct = rbytes <= kDirectResultBytes;  // see bsr_search
        char* r = direct ? h + off_r : (char*)s->d_result.p;
        if (K == 32) {  // merge + unpack in one launch, one warp per query
            merge32_kernel<<<nq, 32, 0, s->stream>>>((const u64*)s->d_gather.p, (u32)c->n_ranks, (size_t)nq * K, K, k,
                                                     nullptr, (u64*)r, (float*)(r + ibytes),
                                                     (u32*)(r + ibytes + dbytes));
        } else {
            TRY(merge_tree(s, (const u64*)s->d_gather.p, (u32)c->n_ranks, nq, K, k, (size_t)nq * K, K, merged,
                           s->stream));
            unpack_kernel<<<nq, 128, 0, s->stream>>>(merged, K, k, (u64*)r, (float*)(r + ibytes),
                                                     (u32*)(r + ibytes + dbytes));
            s->launches++;
        }
        s->launches++;
        CU(cudaGetLastError());
        if (!direct) CU(cudaMemcpyAsync(h + off_r, r, rbytes, cudaMemcpyDeviceToHost, s->stream));
        return BSR_OK;
    };
    int merge_rc = BSR_OK;
    if (root && local_rc == BSR_OK) merge_rc = root_merge();
    CU(cudaStreamSynchronize(s->stream));
    mark_idle(s);
    if (local_rc != BSR_OK) {
        memcpy(g_err, local_err, sizeof g_err);
        return local_rc;
    }
    TRY(merge_rc);
    if (root) {
        memcpy(out_idx, h + off_r, ibytes);
        memcpy(out_dist, h + off_r + ibytes, dbytes);
        memcpy(out_count, h + off_r + ibytes + dbytes, cbytes);
    }
    return BSR_OK;
}

// The same collective with device-resident queries and candidates, asynchronous on `stream`: what a
// host keeps in flight back to back (the mailbox lets a rank run kMbDepth searches ahead of root).
extern "C" int bsr_search_parallel_device(bsr_store* s, bsr_comm* c, const float* d_queries, uint32_t nq,
                                          uint32_t k, uint64_t* d_candidates, void* stream) {
    TRY(check_search_args(s, d_queries, nq, k, "bsr_search_parallel_device"));
    CHECK(c && c->comm, BSR_ERR_INVALID, "bsr_search_parallel_device: null communicator");
    CHECK(c->device == s->device, BSR_ERR_INVALID, "bsr_search_parallel_device: shard and communicator are on different devices");
    if (nq == 0) return BSR_OK;
    CHECK(d_candidates, BSR_ERR_INVALID, "bsr_search_parallel_device: null output");
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    const u32 K = bsr_candidate_slots(k);
    const size_t cand_bytes = (size_t)nq * K * sizeof(u64);
    TRY(order_after_previous(s, st));
    if (mailbox_usable(s, c, nq, K)) {
        const u32 seq = next_mailbox_seq(c);
        const Mailbox mb = mailbox_for(c, seq);
        int rc = BSR_OK;
        if (s->n_rows != 0) {
            s->mailbox = &mb;
            rc = search_device_impl(s, d_queries, nq, k, (u64*)d_candidates, nullptr, nullptr, nullptr, st);
            s->mailbox = nullptr;
        }
        char local_err[sizeof g_err];
        if (rc != BSR_OK) memcpy(local_err, g_err, sizeof g_err);
        if (s->n_rows == 0 || rc != BSR_OK)
            TRY(launch_mailbox_only(s, mb, k, (u64*)d_candidates, nullptr, 0, 0, st));
        TRY(mark_busy(s, st));
        if (rc != BSR_OK) memcpy(g_err, local_err, sizeof g_err);
        return rc;
    }
    TRY(ensure(s->d_cand, cand_bytes));
    TRY(ensure(s->d_gather, cand_bytes * c->n_ranks));
    int rc = BSR_OK;
    if (s->n_rows != 0) rc = search_device_impl(s, d_queries, nq, k, (u64*)s->d_cand.p, nullptr, nullptr, nullptr, st);
    char local_err[sizeof g_err];
    if (rc != BSR_OK) memcpy(local_err, g_err, sizeof g_err);
    if (s->n_rows == 0 || rc != BSR_OK) CU(cudaMemsetAsync(s->d_cand.p, 0xFF, cand_bytes, st));
    NCCL(nccl().AllGather(s->d_cand.p, s->d_gather.p, (size_t)nq * K, kNcclUint64, c->comm, st));
    if (c->rank == 0) {
        if (K == 32) {
            merge32_kernel<<<nq, 32, 0, st>>>((const u64*)s->d_gather.p, (u32)c->n_ranks, (size_t)nq * K, K, k,
                                              (u64*)d_candidates, nullptr, nullptr, nullptr);
            s->launches++;
            CU(cudaGetLastError());
        } else {
            const u32 max_per_cta = std::max<u32>(2, kMergeMaxKeys / K);
            const size_t inter = (size_t)nq * ((c->n_ranks + max_per_cta - 1) / max_per_cta) * K * sizeof(u64);
            TRY(ensure(s->merge_a, inter));
            TRY(ensure(s->merge_b, inter));
            TRY(merge_tree(s, (const u64*)s->d_gather.p, (u32)c->n_ranks, nq, K, k, (size_t)nq * K, K,
                           (u64*)d_candidates, st));
        }
    }
    TRY(mark_busy(s, st));
    if (rc != BSR_OK) memcpy(g_err, local_err, sizeof g_err);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
extern "C" int bsr_store_launch_count(const bsr_store* s, uint64_t* n) {
    CHECK(s && n, BSR_ERR_INVALID, "bsr_store_launch_count: null argument");
    *n = s->launches;
    return BSR_OK;
}

extern "C" int bsr_store_last_scan_ms(bsr_store* s, float* ms) {
    CHECK(s && ms, BSR_ERR_INVALID, "bsr_store_last_scan_ms: null argument");
    CHECK(s->scan_timed, BSR_ERR_INVALID, "bsr_store_last_scan_ms: no scan has run on this store");
    if (s->scan_time_in_ll) {  // a one-launch search: measured by the launch itself (FinalOut::t_start)
        *ms = (float)s->last_scan_ns * 1e-6f;
        return BSR_OK;
    }
    DeviceGuard g(s->device);
    CU(cudaEventSynchronize(s->ev_scan1));
    CU(cudaEventElapsedTime(ms, s->ev_scan0, s->ev_scan1));
    return BSR_OK;
}

extern "C" int bsr_store_set_profiling(bsr_store* s, int enabled) {
    CHECK(s, BSR_ERR_INVALID, "bsr_store_set_profiling: null store");
    DeviceGuard g(s->device);
    TRY(prof_fold(s));
    s->profiling = enabled != 0;
    s->prof_ms = 0.0;
    s->prof_scans = 0;
    return BSR_OK;
}

extern "C" int bsr_store_scan_profile(bsr_store* s, double* total_ms, uint64_t* n_scans) {
    CHECK(s && total_ms && n_scans, BSR_ERR_INVALID, "bsr_store_scan_profile: null argument");
    DeviceGuard g(s->device);
    TRY(prof_fold(s));
    *total_ms = s->prof_ms;
    *n_scans = s->prof_scans;
    return BSR_OK;
}

extern "C" int bsr_store_set_metric(bsr_store* s, bsr_metric metric) {
    CHECK(s, BSR_ERR_INVALID, "bsr_store_set_metric: null store");
    CHECK(metric == BSR_COSINE || metric == BSR_DOT, BSR_ERR_INVALID, "bsr_store_set_metric: bad metric %d", (int)metric);
    s->metric = metric;
    return BSR_OK;
}

extern "C" int bsr_store_metric(const bsr_store* s, bsr_metric* metric) {
    CHECK(s && metric, BSR_ERR_INVALID, "bsr_store_metric: null argument");
    *metric = s->metric;
    return BSR_OK;
}

extern "C" int bsr_store_set_option(bsr_store* s, const char* name, int64_t value) {
    CHECK(s && name, BSR_ERR_INVALID, "bsr_store_set_option: null argument");
    const int v = (int)value;
    struct { const char* name; int* field; } table[] = {
        {"scan_ldg", &s->kn.scan_ldg},         {"ldg_u", &s->kn.ldg_u},
        {"bulk_stages", &s->kn.bulk_stages},   {"max_qb", &s->kn.max_qb},
        {"no_wide", &s->kn.no_wide},           {"no_fuse_wide", &s->kn.no_fuse_wide},
        {"max_grid", &s->kn.max_grid},         {"tensor_mode", &s->kn.tensor_mode},
        {"tensor_stages", &s->kn.tensor_stages}, {"tensor_stage_kb", &s->kn.tensor_stage_kb},
        {"tc_stats", &s->kn.tc_stats},         {"no_pdl", &s->kn.no_pdl},
        {"wide", &s->kn.wide},                 {"no_tickets", &s->kn.no_tickets},
        {"no_param_query", &s->kn.no_param_query}, {"no_f32_tensor", &s->kn.no_f32_tensor},
        {"no_mailbox", &s->kn.no_mailbox},     {"no_small", &s->kn.no_small},
        {"no_list_ll", &s->kn.no_list_ll},     {"no_l2_keep", &s->kn.no_l2_keep},
    };
    for (auto& t : table)
        if (strcmp(t.name, name) == 0) {
            *t.field = v;
            if (t.field == &s->kn.max_qb) s->kn.max_qb = std::max(1, v);
            return BSR_OK;
        }
    return fail(BSR_ERR_INVALID, "bsr_store_set_option: unknown option '%s'", name);
}

extern "C" int bsr_store_set_tensor_min_batch(bsr_store* s, uint32_t min_batch) {
    CHECK(s, BSR_ERR_INVALID, "bsr_store_set_tensor_min_batch: null store");
    s->tensor_min_batch = min_batch;
    return BSR_OK;
}
