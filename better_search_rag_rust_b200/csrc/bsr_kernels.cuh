// bsr_kernels.cuh -- sm_100a device code of the exact top-k search path.
//
// Data layout in HBM ("panels"): the shard's rows are stored in groups of 32 rows (one panel per
// warp pass), interleaved at 16-byte granularity:
//
//     panel p, column-quad j (columns 4j..4j+3), row r = 32p + l   ->   float4 at
//     panels[(p * dq + j) * 32 + l]                      (dq = ceil(dim / 4))
//
// so that lane l of a warp owns row 32p+l, walks its columns in the reference's order, and every
// warp-wide 128-bit load is one contiguous 512-byte segment (a panel is one contiguous
// dq*512-byte stream).  This is what lets the scan keep the reference's *sequential* binary32
// dot product (src/metrics.rs:153, Iterator::sum) -- one thread per row, one dependent
// add chain per row -- and still read HBM fully coalesced.
//
// Reference citations are relative to the reference tree (nichmorgan/better-search-rag-rust).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bsr {

typedef unsigned long long u64;
typedef uint32_t u32;

constexpr int kPanelRows = 32;
constexpr int kScanWarps = 16;
constexpr int kScanThreads = kScanWarps * 32;
constexpr u64 kKeyMax = ~0ull;
constexpr int kMergeThreads = 512;
constexpr u32 kMergeMaxKeys = 16384;  // keys one merge CTA holds in shared memory (128 KiB)

// ------------------------------------------------------------------------------------------
// candidate keys: (order-preserving distance bits) << 32 | global row index.
// Unsigned ascending key order == ascending distance, ties by ascending index, which is the
// order the reference's stable sorts produce (src/mpi_helpers/metrics.rs:47,153).
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 ordered_from_bits(u32 b) {
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ u32 bits_from_ordered(u32 o) {
    return (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
}
__device__ __forceinline__ u64 make_key(float dist, u32 gidx) {
    // partial_cmp treats -0.0 and +0.0 as equal: canonicalise to +0.0 (cosine distances are
    // never -0.0 anyway).
    u32 b = __float_as_uint(__fadd_rn(dist, 0.0f));
    return ((u64)ordered_from_bits(b) << 32) | (u64)gidx;
}

// streaming 128-bit load: read-only path, do not keep the line in L1 (each byte is used once)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "l"(p));
    return v;
}

// One reference step `acc = acc + a_i * b_i` for four consecutive columns: separately rounded
// multiply and add, never contracted to FMA (rustc does not contract; __fmul_rn/__fadd_rn are
// exempt from nvcc's -fmad).
__device__ __forceinline__ float dot4_seq(float acc, const float4& a, const float4& q) {
    acc = __fadd_rn(acc, __fmul_rn(a.x, q.x));
    acc = __fadd_rn(acc, __fmul_rn(a.y, q.y));
    acc = __fadd_rn(acc, __fmul_rn(a.z, q.z));
    acc = __fadd_rn(acc, __fmul_rn(a.w, q.w));
    return acc;
}

// src/metrics.rs:13 -- `(a[i] - b[i]).abs() > 1e-10`; false for NaN.
__device__ __forceinline__ bool differs(float a, float b) {
    return fabsf(__fsub_rn(a, b)) > 1e-10f;
}
__device__ __forceinline__ bool differs4(const float4& a, const float4& q) {
    return differs(a.x, q.x) || differs(a.y, q.y) || differs(a.z, q.z) || differs(a.w, q.w);
}

// src/metrics.rs:157-164 given dot, |a|, |b| and the identity flag (:148-151).
__device__ __forceinline__ float finish_distance(float dot, float mag_a, float mag_b,
                                                 bool identical) {
    if (identical) return 0.0f;
    if (mag_a == 0.0f || mag_b == 0.0f) return 1.0f;
    float sim = __fdiv_rn(dot, __fmul_rn(mag_a, mag_b));
    sim = fminf(fmaxf(sim, -1.0f), 1.0f);  // fmaxf/fminf drop a NaN operand like f32::max/min
    return __fsub_rn(1.0f, sim);
}

// ------------------------------------------------------------------------------------------
// per-warp sorted candidate list in shared memory: insert x (x < list[k-1]) keeping order.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_list_insert(u64* list, int k, u64 x, int lane, u64& tau) {
    for (int base = (k - 1) & ~31; base >= 0; base -= 32) {
        const int j = base + lane;
        u64 cur = kKeyMax, prev = 0;
        if (j < k) {
            cur = list[j];
            if (j > 0) prev = list[j - 1];
        }
        __syncwarp();
        if (j < k && cur > x) list[j] = (prev > x) ? prev : x;
        __syncwarp();
        // lane 0 holds list[base-1] in prev (0 when base == 0): if it is <= x the insertion
        // point is inside this chunk and everything below is untouched.
        const u64 below = __shfl_sync(0xFFFFFFFFu, prev, 0);
        if (below <= x) break;
    }
    tau = list[k - 1];
}

__device__ __forceinline__ void warp_offer(u64* list, int k, u64 key, int lane, u64& tau) {
    unsigned m = __ballot_sync(0xFFFFFFFFu, key < tau);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const u64 x = __shfl_sync(0xFFFFFFFFu, key, src);
        if (x < tau) warp_list_insert(list, k, x, lane, tau);
    }
}

// ------------------------------------------------------------------------------------------
// CTA-wide: reduce L (power of two) ascending lists of K (power of two) keys, contiguous in
// shared memory, to the K smallest keys, ascending, in buf[0..K).
//   min(A[j], B[K-1-j]) holds the K smallest of A u B as a bitonic sequence; log2(K) half-cleaner
//   steps sort it.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cta_merge_lists(u64* buf, u32 L, u32 K) {
    const u32 tid = threadIdx.x, nt = blockDim.x;
    for (u32 s = L >> 1; s >= 1; s >>= 1) {
        for (u32 t = tid; t < s * K; t += nt) {
            const u32 i = t / K, j = t - i * K;
            const u64 a = buf[i * K + j];
            const u64 b = buf[(i + s) * K + (K - 1 - j)];
            buf[i * K + j] = a < b ? a : b;
        }
        __syncthreads();
        for (u32 d = K >> 1; d >= 1; d >>= 1) {
            for (u32 t = tid; t < s * (K >> 1); t += nt) {
                const u32 i = t / (K >> 1), jj = t - i * (K >> 1);
                const u32 lo = i * K + ((jj / d) * 2 * d) + (jj % d);
                const u32 hi = lo + d;
                const u64 x = buf[lo], y = buf[hi];
                if (x > y) {
                    buf[lo] = y;
                    buf[hi] = x;
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// K1: fused cosine scan + per-warp top-k + per-CTA merge  (fp32 rows, one query)
//     restates compute_local_top_k's two hot loops (src/mpi_helpers/metrics.rs:36-50) without
//     ever writing the distance vector to HBM.
// ------------------------------------------------------------------------------------------
struct ScanArgs {
    const float4* panels;  // [n_panels][dq][32]
    const float* norms;    // [n_rows]  |a| = sqrt(seq-sum a_i^2)  (src/metrics.rs:154)
    const float* query;    // [dim]
    const float* qnorm;    // [1]      |b|                        (src/metrics.rs:155)
    u64* cta_keys;         // [gridDim.x][k_pad]
    u32 n_rows, n_panels, dq, dim, k, k_pad;
    u32 row_base;          // global index of local row 0
};

template <int U>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_f32_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sq = reinterpret_cast<float4*>(smem_raw);
    u64* lists = reinterpret_cast<u64*>(smem_raw + (size_t)a.dq * sizeof(float4));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 dq = a.dq;

    // stage the query in shared memory, zero-padded to a multiple of 4 columns (0*0 terms leave
    // the sequential sum unchanged)
    for (u32 j = tid; j < dq; j += kScanThreads) {
        const u32 c = 4 * j;
        float4 v;
        v.x = (c + 0 < a.dim) ? a.query[c + 0] : 0.0f;
        v.y = (c + 1 < a.dim) ? a.query[c + 1] : 0.0f;
        v.z = (c + 2 < a.dim) ? a.query[c + 2] : 0.0f;
        v.w = (c + 3 < a.dim) ? a.query[c + 3] : 0.0f;
        sq[j] = v;
    }
    u64* wl = lists + (size_t)warp * a.k_pad;
    for (u32 j = lane; j < a.k_pad; j += 32) wl[j] = kKeyMax;
    __syncthreads();

    const float mag_b = *a.qnorm;
    const int k = (int)a.k;
    u64 tau = kKeyMax;

    const u32 total_warps = gridDim.x * kScanWarps;
    u32 panel = blockIdx.x * kScanWarps + warp;
    if (panel < a.n_panels) {
        const size_t panel_stride = (size_t)dq * 32;
        const float4* p = a.panels + (size_t)panel * panel_stride + lane;
        u32 j0 = 0;
        float4 cur[U], nxt[U];
#pragma unroll
        for (int u = 0; u < U; u++) cur[u] = ldg_stream(p + (size_t)u * 32);

        float acc = 0.0f, mag_a = 0.0f;
        bool maybe_identical = false;
        while (true) {
            // where the next chunk of U column-quads lives (possibly in this warp's next panel)
            u32 nj0 = j0 + U, npanel = panel;
            if (nj0 == dq) {
                nj0 = 0;
                npanel = panel + total_warps;
            }
            const bool has_next = npanel < a.n_panels;
            if (has_next) {
                const float4* np = a.panels + (size_t)npanel * panel_stride + (size_t)nj0 * 32 + lane;
#pragma unroll
                for (int u = 0; u < U; u++) nxt[u] = ldg_stream(np + (size_t)u * 32);
            }
            if (j0 == 0) {
                const u32 row = panel * kPanelRows + lane;
                mag_a = (row < a.n_rows) ? __ldg(a.norms + row) : 0.0f;
                acc = 0.0f;
                // vectors_are_identical exits at the first differing element
                // (src/metrics.rs:12-16): pre-filter on columns 0..3, full check only if needed
                maybe_identical = !differs4(cur[0], sq[0]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) acc = dot4_seq(acc, cur[u], sq[j0 + u]);

            if (j0 + U == dq) {
                const u32 row = panel * kPanelRows + lane;
                u64 key = kKeyMax;
                if (row < a.n_rows) {
                    bool identical = false;
                    if (maybe_identical) {
                        identical = true;
                        const float4* pr = a.panels + (size_t)panel * panel_stride + lane;
                        for (u32 j = 1; j < dq; j++) {
                            if (differs4(pr[(size_t)j * 32], sq[j])) {
                                identical = false;
                                break;
                            }
                        }
                    }
                    const float dist = finish_distance(acc, mag_a, mag_b, identical);
                    key = make_key(dist, a.row_base + row);
                }
                warp_offer(wl, k, key, lane, tau);
            }
            if (!has_next) break;
#pragma unroll
            for (int u = 0; u < U; u++) cur[u] = nxt[u];
            j0 = nj0;
            panel = npanel;
        }
    }
    __syncthreads();
    cta_merge_lists(lists, kScanWarps, a.k_pad);
    u64* out = a.cta_keys + (size_t)blockIdx.x * a.k_pad;
    for (u32 j = tid; j < a.k_pad; j += kScanThreads) out[j] = lists[j];
}

// ------------------------------------------------------------------------------------------
// K3: merge candidate lists (cross-CTA after the scan; cross-rank after the all-gather).
//     restates compute_global_top_k's sort + first-k (src/mpi_helpers/metrics.rs:147-170).
//   list l of query q starts at in + l*list_stride + q*query_stride; CTA (g, q) reduces lists
//   [g*lists_per_cta, (g+1)*lists_per_cta) to out[(q*gridDim.x + g) * K].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMergeThreads) merge_lists_kernel(
    const u64* __restrict__ in, u32 n_lists, u32 K, u32 lists_per_cta, size_t list_stride,
    size_t query_stride, u64* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* buf = reinterpret_cast<u64*>(smem_raw);
    const u32 g = blockIdx.x, q = blockIdx.y;
    const u32 first = g * lists_per_cta;
    for (u32 t = threadIdx.x; t < lists_per_cta * K; t += blockDim.x) {
        const u32 l = t / K, j = t - l * K;
        u64 v = kKeyMax;
        if (first + l < n_lists) v = in[(size_t)(first + l) * list_stride + (size_t)q * query_stride + j];
        buf[t] = v;
    }
    __syncthreads();
    cta_merge_lists(buf, lists_per_cta, K);
    u64* o = out + ((size_t)q * gridDim.x + g) * K;
    for (u32 j = threadIdx.x; j < K; j += blockDim.x) o[j] = buf[j];
}

// candidates[q][K] -> idx[q][k] (u64, as the reference's usize), dist[q][k], count[q]
__global__ void unpack_kernel(const u64* __restrict__ cand, u32 K, u32 k, u64* __restrict__ idx,
                              float* __restrict__ dist, u32* __restrict__ count) {
    const u32 q = blockIdx.x;
    int valid = 0;
    for (u32 j = threadIdx.x; j < k; j += blockDim.x) {
        const u64 key = cand[(size_t)q * K + j];
        const bool ok = key != kKeyMax;
        valid += ok ? 1 : 0;
        idx[(size_t)q * k + j] = ok ? (key & 0xFFFFFFFFull) : 0ull;
        dist[(size_t)q * k + j] = ok ? __uint_as_float(bits_from_ordered((u32)(key >> 32))) : 0.0f;
    }
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (valid) atomicAdd(&total, valid);
    __syncthreads();
    if (threadIdx.x == 0) count[q] = (u32)total;
}

// ------------------------------------------------------------------------------------------
// query preparation: |b| = sqrt(sequential sum b_i^2)  (src/metrics.rs:155), one thread per query
// ------------------------------------------------------------------------------------------
__global__ void query_norm_kernel(const float* __restrict__ queries, u32 n_queries, u32 dim,
                                  float* __restrict__ qnorm) {
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const float* b = queries + (size_t)q * dim;
    float s = 0.0f;
    for (u32 i = 0; i < dim; i++) {
        const float v = b[i];
        s = __fadd_rn(s, __fmul_rn(v, v));
    }
    qnorm[q] = __fsqrt_rn(s);
}

// ------------------------------------------------------------------------------------------
// K4: load-time kernels
// ------------------------------------------------------------------------------------------

// |a| for local rows [row_lo, row_hi): one thread per row, sequential (src/metrics.rs:154)
__global__ void row_norms_kernel(const float4* __restrict__ panels, float* __restrict__ norms,
                                 u32 row_lo, u32 row_hi, u32 dq) {
    const u32 row = (row_lo & ~31u) + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row_hi) return;
    const float4* p = panels + (size_t)(row >> 5) * dq * 32 + (row & 31);
    float s = 0.0f;
#pragma unroll 8
    for (u32 j = 0; j < dq; j++) {
        const float4 v = p[(size_t)j * 32];
        s = __fadd_rn(s, __fmul_rn(v.x, v.x));
        s = __fadd_rn(s, __fmul_rn(v.y, v.y));
        s = __fadd_rn(s, __fmul_rn(v.z, v.z));
        s = __fadd_rn(s, __fmul_rn(v.w, v.w));
    }
    if (row >= row_lo) norms[row] = __fsqrt_rn(s);
}

// row-major src[n][dim] -> panels, for local rows [row_lo, row_lo + n).  Thread t of the grid
// owns destination float t of the panel range, so stores are fully coalesced.
__global__ void interleave_rows_kernel(const float* __restrict__ src, float* __restrict__ panels,
                                       u32 row_lo, u32 n, u32 dim, u32 dq) {
    const size_t per_panel = (size_t)dq * 128;
    const u32 first_panel = row_lo >> 5;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 panel = first_panel + (u32)(t / per_panel);
    const u32 d = (u32)(t % per_panel);
    const u32 quad = d >> 7, rr = (d & 127) >> 2, e = d & 3;
    const u32 row = panel * 32 + rr, col = quad * 4 + e;
    if (row < row_lo || row >= row_lo + n) return;
    const float v = (col < dim) ? src[(size_t)(row - row_lo) * dim + col] : 0.0f;
    panels[(size_t)panel * per_panel + d] = v;
}

// panels -> row-major out[length][dim] for local rows [offset, offset+length)
__global__ void gather_rows_kernel(const float* __restrict__ panels, float* __restrict__ out,
                                   u32 offset, u32 length, u32 dim, u32 dq) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)length * dim) return;
    const u32 r = offset + (u32)(t / dim), c = (u32)(t % dim);
    const size_t per_panel = (size_t)dq * 128;
    out[t] = panels[(size_t)(r >> 5) * per_panel + ((size_t)(c >> 2) * 32 + (r & 31)) * 4 + (c & 3)];
}

// counter-based uniform(-1,1) generator, identical to bsr_oracle_generate_rows (oracle/) and
// mirroring the value distribution of the reference fixture (src/utils.rs:14-23).
__host__ __device__ __forceinline__ u64 splitmix64(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ float synth_elem(u64 row_key, u32 col) {
    const u32 u = (u32)(splitmix64(row_key + col) >> 40);
    return (float)u * 1.1920928955078125e-07f - 1.0f;  // u * 2^-23 - 1, exact
}
__global__ void synth_rows_kernel(float4* __restrict__ panels, u32 row_lo, u32 n, u32 dim, u32 dq,
                                  u64 seed, u64 gen_first_row) {
    const size_t per_panel = (size_t)dq * 32;  // float4 units
    const u32 first_panel = row_lo >> 5;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 panel = first_panel + (u32)(t / per_panel);
    const u32 d = (u32)(t % per_panel);
    const u32 quad = d >> 5, rr = d & 31;
    const u32 row = panel * 32 + rr;
    if (row < row_lo || row >= row_lo + n) return;
    const u64 rk = splitmix64(seed ^ ((gen_first_row + (row - row_lo)) * 0x9E3779B97F4A7C15ull));
    const u32 c = quad * 4;
    float4 v;
    v.x = (c + 0 < dim) ? synth_elem(rk, c + 0) : 0.0f;
    v.y = (c + 1 < dim) ? synth_elem(rk, c + 1) : 0.0f;
    v.z = (c + 2 < dim) ? synth_elem(rk, c + 2) : 0.0f;
    v.w = (c + 3 < dim) ? synth_elem(rk, c + 3) : 0.0f;
    panels[(size_t)panel * per_panel + d] = v;
}

}  // namespace bsr
