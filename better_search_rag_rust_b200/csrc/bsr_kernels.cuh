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
    // a NaN distance (BSR_DOT on NaN / inf*0 inputs; cosine distances are never NaN) sorts after
    // every number whatever its sign bit: one canonical quiet NaN
    if (dist != dist) b = 0x7FC00000u;
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

// src/metrics.rs:13 -- `(a[i] - b[i]).abs() > 1e-10`; false for NaN.
__device__ __forceinline__ bool differs(float a, float b) {
    return fabsf(__fsub_rn(a, b)) > 1e-10f;
}
__device__ __forceinline__ bool differs4(const float4& a, const float4& q) {
    return differs(a.x, q.x) || differs(a.y, q.y) || differs(a.z, q.z) || differs(a.w, q.w);
}
// Bitwise-equal quads never "differ" (a - a is 0, or NaN for NaN / inf, and neither is > 1e-10): the
// common case on the one row that IS the query (the reference's self-query) costs three integer ops
// per quad instead of four subtract-compare pairs.
// (a ^ b) | c as ONE lop3.  Written in PTX on purpose: from C the compiler rewrites "an OR of XORs is
// zero" into a chain of predicated integer compares (ISETP.EQ.AND P0, .., P0), one dependent
// predicate round trip per column -- ~13 cycles each on the lone lane every other lane waits for.
__device__ __forceinline__ u32 xor_or(float a, float b, u32 c) {
    u32 d;
    asm("lop3.b32 %0, %1, %2, %3, 0xBE;" : "=r"(d) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(c));
    return d;
}
__device__ __forceinline__ bool differs4_fast(const float4& a, const float4& q) {
    const u32 x = xor_or(a.x, q.x, xor_or(a.y, q.y, 0u)) | xor_or(a.z, q.z, xor_or(a.w, q.w, 0u));
    return x != 0u && differs4(a, q);
}

// ------------------------------------------------------------------------------------------
// BF16 stores use a second layout, shared by the exact scan and the tensor-core path: tiles of
// 64 rows x 64 columns (8 KiB), stored in HBM as the shared-memory image tcgen05.mma expects for
// a K-major operand with 128-byte swizzle (8-row x 128-byte atoms, 16-byte chunks XOR-ed with the
// row index inside the atom):
//
//   byte offset of element (row r, column c) =
//       ((r/64) * KB + c/64) * 8192 + ((r%64)/8) * 1024 + (r%8) * 128
//       + ((((c%64)/8) ^ (r%8)) * 16) + (c%8) * 2                      KB = ceil(dim/64)
//
// One cp.async.bulk of 8 KiB brings a ready-to-multiply operand tile; the exact scan copies the
// 4 KiB half (32 rows) its warp owns and reads it conflict-free with the same XOR.
// ------------------------------------------------------------------------------------------
constexpr u32 kTileRows = 64;
constexpr u32 kTileBytes = 8192;

__host__ __device__ __forceinline__ size_t bf16_chunk_offset(u32 row, u32 kb, u32 chunk, u32 KB) {
    return ((size_t)(row >> 6) * KB + kb) * kTileBytes + ((row & 63u) >> 3) * 1024u + (row & 7u) * 128u +
           ((chunk ^ (row & 7u)) << 4);
}
__device__ __forceinline__ unsigned short f32_to_bf16_rne(float f) {
    u32 b = __float_as_uint(f);
    if ((b & 0x7F800000u) == 0x7F800000u && (b & 0x007FFFFFu)) return (unsigned short)((b >> 16) | 0x0040u);
    b += 0x7FFFu + ((b >> 16) & 1u);
    return (unsigned short)(b >> 16);
}
__device__ __forceinline__ float bf16_lo(u32 packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi(u32 packed) { return __uint_as_float(packed & 0xFFFF0000u); }

// eight consecutive columns held as packed bf16 (uint4) against two query quads, in column order
__device__ __forceinline__ float dot8_seq_bf16(float acc, const uint4& a, const float4& q0, const float4& q1) {
    acc = __fadd_rn(acc, __fmul_rn(bf16_lo(a.x), q0.x));
    acc = __fadd_rn(acc, __fmul_rn(bf16_hi(a.x), q0.y));
    acc = __fadd_rn(acc, __fmul_rn(bf16_lo(a.y), q0.z));
    acc = __fadd_rn(acc, __fmul_rn(bf16_hi(a.y), q0.w));
    acc = __fadd_rn(acc, __fmul_rn(bf16_lo(a.z), q1.x));
    acc = __fadd_rn(acc, __fmul_rn(bf16_hi(a.z), q1.y));
    acc = __fadd_rn(acc, __fmul_rn(bf16_lo(a.w), q1.z));
    acc = __fadd_rn(acc, __fmul_rn(bf16_hi(a.w), q1.w));
    return acc;
}
__device__ __forceinline__ bool differs8_bf16(const uint4& a, const float4& q0, const float4& q1) {
    return differs(bf16_lo(a.x), q0.x) || differs(bf16_hi(a.x), q0.y) || differs(bf16_lo(a.y), q0.z) ||
           differs(bf16_hi(a.y), q0.w) || differs(bf16_lo(a.z), q1.x) || differs(bf16_hi(a.z), q1.y) ||
           differs(bf16_lo(a.w), q1.z) || differs(bf16_hi(a.w), q1.w);
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

// |b|^2 accumulation step for the query (src/metrics.rs:155), same sequential order
__device__ __forceinline__ float sq4_seq(float acc, const float4& q) {
    acc = __fadd_rn(acc, __fmul_rn(q.x, q.x));
    acc = __fadd_rn(acc, __fmul_rn(q.y, q.y));
    acc = __fadd_rn(acc, __fmul_rn(q.z, q.z));
    acc = __fadd_rn(acc, __fmul_rn(q.w, q.w));
    return acc;
}

// src/metrics.rs:157-164 given dot, |a|, |b| and the identity flag (:148-151).
// metric != 0 (BSR_DOT, include/bsr.h): 1 - dot, nothing else.
__device__ __forceinline__ float finish_distance(float dot, float mag_a, float mag_b,
                                                 bool identical, u32 metric = 0) {
    if (metric) return __fsub_rn(1.0f, dot);
    if (identical) return 0.0f;
    if (mag_a == 0.0f || mag_b == 0.0f) return 1.0f;
    float sim = __fdiv_rn(dot, __fmul_rn(mag_a, mag_b));
    sim = fminf(fmaxf(sim, -1.0f), 1.0f);  // fmaxf/fminf drop a NaN operand like f32::max/min
    return __fsub_rn(1.0f, sim);
}

// Full identity test (src/metrics.rs:7-19) for the rare rows whose first four columns already
// match the query: the whole warp checks the flagged lane's row, 32 column-quads at a time, so
// the test costs a couple of memory round trips instead of dq dependent ones.
__device__ __forceinline__ bool warp_confirm_identical(bool maybe, const float4* panel_base,
                                                       const float4* sq, u32 dq, int lane) {
    unsigned m = __ballot_sync(0xFFFFFFFFu, maybe);
    bool identical = false;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        bool diff = false;
        for (u32 j = lane; j < dq; j += 32) diff |= differs4(panel_base[(size_t)j * 32 + src], sq[j]);
        const bool any_diff = __any_sync(0xFFFFFFFFu, diff);
        if (lane == src) identical = !any_diff;
    }
    return identical;
}

// same for a BF16 store: `tiles` is the store base, `row` the flagged lane's local row
__device__ __forceinline__ bool warp_confirm_identical_bf16(bool maybe, const unsigned char* tiles,
                                                            u32 row_of_lane0, u32 KB, const float4* sq,
                                                            int lane) {
    unsigned m = __ballot_sync(0xFFFFFFFFu, maybe);
    bool identical = false;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const u32 row = row_of_lane0 + src;
        bool diff = false;
        for (u32 g = lane; g < KB * 8; g += 32) {  // g = 16-byte chunk index along the row
            const uint4 a = *reinterpret_cast<const uint4*>(tiles + bf16_chunk_offset(row, g >> 3, g & 7u, KB));
            diff |= differs8_bf16(a, sq[2 * g], sq[2 * g + 1]);
        }
        const bool any_diff = __any_sync(0xFFFFFFFFu, diff);
        if (lane == src) identical = !any_diff;
    }
    return identical;
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

// ascending bitonic sort of K (power of two >= 32) keys in shared memory by one warp
__device__ __forceinline__ void warp_sort_list(u64* list, u32 K, int lane) {
    for (u32 size = 2; size <= K; size <<= 1) {
        for (u32 d = size >> 1; d >= 1; d >>= 1) {
            const u32 ld = 31 - __clz(d);  // d is a power of two: no integer division in the index arithmetic
            for (u32 t = lane; t < (K >> 1); t += 32) {
                const u32 lo = ((t >> ld) << (ld + 1)) + (t & (d - 1)), hi = lo + d;
                const bool asc = (lo & size) == 0;
                const u64 x = list[lo], y = list[hi];
                if ((x > y) == asc) {
                    list[lo] = y;
                    list[hi] = x;
                }
            }
            __syncwarp();
        }
    }
}

// ---- register-resident 32-key blocks (lane i holds the i-th key) ---------------------------
// ascending bitonic sort of 32 keys, one per lane
__device__ __forceinline__ u64 warp_sort32(u64 v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int d = size >> 1; d >= 1; d >>= 1) {
            const u64 p = __shfl_xor_sync(0xFFFFFFFFu, v, d);
            const bool asc = (lane & size) == 0;          // size == 32: every lane ascending
            const bool keep_min = ((lane & d) == 0) == asc;
            v = (keep_min == (v < p)) ? v : p;
        }
    }
    return v;
}
// a, c: ascending 32-key blocks.  Returns the 32 smallest of the 64 (ascending) and leaves the 32
// largest (ascending) in c.
__device__ __forceinline__ u64 warp_merge_split32(u64 a, u64& c, int lane) {
    const u64 rev = __shfl_sync(0xFFFFFFFFu, c, 31 - lane);
    u64 lo = a < rev ? a : rev, hi = a < rev ? rev : a;  // both bitonic
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const u64 pl = __shfl_xor_sync(0xFFFFFFFFu, lo, d);
        const u64 ph = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
        const bool keep_min = (lane & d) == 0;
        lo = (keep_min == (lo < pl)) ? lo : pl;
        hi = (keep_min == (hi < ph)) ? hi : ph;
    }
    c = hi;
    return lo;
}

// Per-warp running top-k.  Fill phase: the first K keys are appended unsorted and sorted once;
// after that only keys below the current k-th best (tau) are taken.
//   K == 32 (k <= 32): a taken key is inserted at once (rare: ~k ln(rows/k) per warp).
//   K  > 32: taken keys collect in a register block (one per lane); when 32 are pending the block
//   is sorted with shuffles and rippled through the list's 32-key blocks (merge-split: the block
//   keeps the smaller half, the larger half carries on), skipping the blocks that lie entirely
//   below it.  ~K/64 shuffle merges per 32 candidates instead of a ~K/64-step shared-memory
//   shift per candidate (k = 1024 on a 12.5 M-row shard: 2.7 ms of inserts per scan before).
struct WarpTopK {
    u64* list;
    int k;
    u32 K, cnt;
    u64 tau;
    u64 pend;  // K > 32: my pending candidate (lanes < pc)
    u32 pc;    // K > 32: number of pending candidates (warp-uniform)

    __device__ __forceinline__ void init(u64* l, u32 k_, u32 K_, int lane) {
        list = l;
        k = (int)k_;
        K = K_;
        cnt = 0;
        tau = kKeyMax;
        pend = kKeyMax;
        pc = 0;
        for (u32 j = lane; j < K; j += 32) list[j] = kKeyMax;
        __syncwarp();
    }
    __device__ __forceinline__ void flush(int lane) {
        u64 carry = warp_sort32((u32)lane < pc ? pend : kKeyMax, lane);
        for (u32 b = 0; b < K; b += 32) {
            const u64 cmin = __shfl_sync(0xFFFFFFFFu, carry, 0);
            if (cmin == kKeyMax) break;  // nothing left to place
            const u64 a = list[b + lane];
            const u64 amax = __shfl_sync(0xFFFFFFFFu, a, 31);
            if (cmin > amax) continue;   // the whole carry belongs above this block
            list[b + lane] = warp_merge_split32(a, carry, lane);
        }
        __syncwarp();
        tau = list[k - 1];
        pend = kKeyMax;
        pc = 0;
    }
    __device__ __forceinline__ void offer(u64 key, int lane) {
        if (cnt < K) {  // warp-uniform
            list[cnt + lane] = key;
            cnt += 32;
            __syncwarp();
            if (cnt == K) {
                warp_sort_list(list, K, lane);
                tau = list[k - 1];
            }
            return;
        }
        unsigned m = __ballot_sync(0xFFFFFFFFu, key < tau);
        if (K == 32) {
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const u64 x = __shfl_sync(0xFFFFFFFFu, key, src);
                if (x < tau) warp_list_insert(list, k, x, lane, tau);
            }
            return;
        }
        if (!m) return;
        const u32 c = (u32)__popc(m);
        if (pc + c > 32) flush(lane);
        // lane d in [pc, pc + c) takes the (d - pc)-th candidate of this offer
        const u32 nth = (u32)lane - pc;
        const unsigned src = __fns(m, 0, (nth < c ? nth : 0u) + 1);
        const u64 v = __shfl_sync(0xFFFFFFFFu, key, (int)(src & 31u));
        if (nth < c) pend = v;
        pc += c;
        if (pc == 32) flush(lane);
    }
    __device__ __forceinline__ void finish(int lane) {
        if (cnt < K) {
            // fill phase not completed: slots [cnt, K) still hold KEY_MAX.  One offer (a single panel, the
            // small-store case): the 32 keys are sorted in registers; otherwise the filled prefix, rounded
            // up to a power of two, in shared memory.
            __syncwarp();
            if (cnt <= 32) {
                if (cnt) list[lane] = warp_sort32(list[lane], lane);
                __syncwarp();
            } else {
                u32 n2 = 64;
                while (n2 < cnt) n2 <<= 1;
                warp_sort_list(list, n2, lane);
            }
        } else if (pc) {
            flush(lane);
        }
    }
};

// ------------------------------------------------------------------------------------------
// CTA-wide: reduce L (power of two) ascending lists of K (power of two) keys, contiguous in
// shared memory, to the K smallest keys, ascending, in buf[0..K).
//   min(A[j], B[K-1-j]) holds the K smallest of A u B as a bitonic sequence; log2(K) half-cleaner
//   steps sort it.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cta_merge_lists(u64* buf, u32 L, u32 K) {
    const u32 tid = threadIdx.x, nt = blockDim.x;
    const u32 logK = 31 - __clz(K);
    for (u32 s = L >> 1; s >= 1; s >>= 1) {
        for (u32 t = tid; t < (s << logK); t += nt) {
            const u32 i = t >> logK, j = t & (K - 1);
            const u64 a = buf[(i << logK) + j];
            const u64 b = buf[((i + s) << logK) + (K - 1 - j)];
            buf[(i << logK) + j] = a < b ? a : b;
        }
        __syncthreads();
        for (u32 ld = logK; ld-- > 0;) {
            const u32 d = 1u << ld;
            for (u32 t = tid; t < (s << (logK - 1)); t += nt) {
                const u32 i = t >> (logK - 1), jj = t & ((K >> 1) - 1);
                const u32 lo = (i << logK) + ((jj >> ld) << (ld + 1)) + (jj & (d - 1));
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

// Register version for K == 32: lane i holds the i-th key of an ascending list.  Returns (lane i
// of) the 32 smallest keys of the two lists, ascending: min(a[i], b[31-i]) is bitonic, five
// shuffle half-cleaners sort it.
__device__ __forceinline__ u64 warp_merge32(u64 mine, u64 other, int lane) {
    const u64 rev = __shfl_sync(0xFFFFFFFFu, other, 31 - lane);
    u64 v = mine < rev ? mine : rev;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const u64 p = __shfl_xor_sync(0xFFFFFFFFu, v, d);
        const bool keep_min = (lane & d) == 0;
        v = (keep_min == (v < p)) ? v : p;
    }
    return v;
}

__device__ __forceinline__ u64 ldcg_u64(const u64* p) {
    u64 v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// ------------------------------------------------------------------------------------------
// Cross-rank candidate exchange over NVLink peer memory ("mailbox"), fused into the scan's last CTA.
// Replaces gather_top_k_results' messages (src/mpi_helpers/metrics.rs:56-138) and the separate merge
// launch: every non-root rank stores its shard's 32 candidate slots straight into root's HBM; root's
// last CTA waits for them, folds them into its own list and writes the final top-k.
//
//   entry = 16 bytes {key, seq}, written with ONE 16-byte store: a peer store is delivered whole, so an
//   entry whose seq field equals the current search's sequence number carries that search's key -- no
//   fence, no separate flag (the "LL" protocol of collective libraries).
//   root's mailbox: [kMbDepth][n_ranks][kMbSlots] entries; search `seq` uses slot seq % kMbDepth, so a peer
//   may run up to kMbDepth searches ahead of root before it has to wait for root's acknowledgement
//   (root stores the last sequence number it has consumed into every peer's `ack` word).
// ------------------------------------------------------------------------------------------
constexpr u32 kMbDepth = 4;
// A wait inside a kernel ends after this many SM cycles (~35 s at 1.95 GHz) with an error status: long
// enough for any skew between the ranks' calls (a collective waits for its slowest caller), short enough
// that a dead rank becomes an error on the others and never an unbounded spin on the GPU.
constexpr long long kMbGiveUpCycles = 1ll << 36;
constexpr u32 kMbSlots = 128;  // entries per rank and search: top_k <= 128 (the reference's own k is 50, src/main.rs:110)

struct Mailbox {
    ulonglong2* send;         // non-root: root's entries for (seq % kMbDepth, my rank) -- PEER memory; else nullptr
    const ulonglong2* recv;   // root: my entries for seq % kMbDepth, [n_ranks][kMbSlots] -- local memory; else nullptr
    const u32* ack_local;     // non-root: where root's acknowledgements land (local memory)
    u32* const* ack_remote;   // root: [n_ranks] pointers to the peers' ack words (PEER memory)
    u32* status;              // pinned host word, set to 1 when a wait below gave up
    u32 n_ranks, seq;
};

__device__ __forceinline__ ulonglong2 ld_volatile_v2(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_volatile_u64(const unsigned long long* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_v2(ulonglong2* p, u64 x, u64 y) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}

// One warp; lane i holds the i-th key of this rank's final (ascending) list.  Non-root: send it.
// Root: fold every peer's list in as it arrives and acknowledge.  Returns the list to publish.
// A rank that never delivers (it died, or the caller broke the collective contract) ends the wait after
// kMbGiveUpCycles with *status = 1; the host turns that into an error.
__device__ __forceinline__ u64 mailbox_exchange32(u64 v, const Mailbox& mb, u32 k, int lane) {
    constexpr long long kGiveUp = kMbGiveUpCycles;
    if (mb.send) {
        if (mb.seq > kMbDepth) {  // flow control: root must be done with the slot's previous use
            const long long t0 = clock64();
            bool ok = true;
            if (lane == 0) {
                while (*reinterpret_cast<const volatile u32*>(mb.ack_local) + kMbDepth < mb.seq) {
                    if (clock64() - t0 > kGiveUp) {
                        ok = false;
                        break;
                    }
                }
                if (!ok) *reinterpret_cast<volatile u32*>(mb.status) = 1;
            }
            __syncwarp();
        }
        st_volatile_v2(mb.send + lane, (u32)lane < k ? v : kKeyMax, (u64)mb.seq);
        return v;
    }
    if (!mb.recv) return v;
    const long long t0 = clock64();
    for (u32 r = 1; r < mb.n_ranks; r++) {
        ulonglong2 e;
        bool have = false;
        while (true) {
            if (!have) {
                e = ld_volatile_v2(mb.recv + (size_t)r * kMbSlots + lane);
                have = e.y == (u64)mb.seq;
            }
            if (__all_sync(0xFFFFFFFFu, have)) break;
            // lane 0 decides for the warp, so that every lane leaves in the same iteration
            if (__shfl_sync(0xFFFFFFFFu, (int)(clock64() - t0 > kGiveUp), 0)) {
                if (lane == 0) *reinterpret_cast<volatile u32*>(mb.status) = 1;
                e.x = have ? e.x : kKeyMax;
                break;
            }
        }
        const u64 their_best = __shfl_sync(0xFFFFFFFFu, e.x, 0);
        const u64 my_kth = __shfl_sync(0xFFFFFFFFu, v, (int)k - 1);
        if (their_best < my_kth) v = warp_merge32(v, e.x, lane);
    }
    // acknowledge: the slot may be overwritten by search seq + kMbDepth
    if (lane > 0 && (u32)lane < mb.n_ranks) *reinterpret_cast<volatile u32*>(mb.ack_remote[lane]) = mb.seq;
    return v;
}

// The same exchange for 32 < K <= 128, by the whole CTA: `ql[0..K)` (shared memory) holds this rank's
// final ascending list; `ql` has room for `room` lists of K keys (the per-warp lists it was reduced from).
// Non-root: send.  Root: the peers' lists land in ql[K..n_ranks*K) as they arrive and the lot is reduced
// again with cta_merge_lists.  Needs n_ranks <= room (host-checked).  Ends with __syncthreads().
__device__ __forceinline__ void mailbox_exchange_cta(u64* ql, u32 K, u32 k, const Mailbox& mb) {
    constexpr long long kGiveUp = kMbGiveUpCycles;
    const u32 tid = threadIdx.x, nt = blockDim.x;
    if (mb.send) {
        if (mb.seq > kMbDepth && tid == 0) {  // flow control, see mailbox_exchange32
            const long long t0 = clock64();
            while (*reinterpret_cast<const volatile u32*>(mb.ack_local) + kMbDepth < mb.seq) {
                if (clock64() - t0 > kGiveUp) {
                    *reinterpret_cast<volatile u32*>(mb.status) = 1;
                    break;
                }
            }
        }
        __syncthreads();
        for (u32 j = tid; j < K; j += nt) st_volatile_v2(mb.send + j, j < k ? ql[j] : kKeyMax, (u64)mb.seq);
        __syncthreads();
        return;
    }
    if (!mb.recv) return;
    const long long t0 = clock64();
    for (u32 t = tid; t < (mb.n_ranks - 1) * K; t += nt) {
        const u32 r = 1 + t / K, j = t % K;
        ulonglong2 e;
        while (true) {
            e = ld_volatile_v2(mb.recv + (size_t)r * kMbSlots + j);
            if (e.y == (u64)mb.seq) break;
            if (clock64() - t0 > kGiveUp) {
                *reinterpret_cast<volatile u32*>(mb.status) = 1;
                e.x = kKeyMax;
                break;
            }
        }
        ql[(size_t)r * K + j] = e.x;
    }
    u32 L = 1;
    while (L < mb.n_ranks) L <<= 1;
    for (u32 t = tid + mb.n_ranks * K; t < L * K; t += nt) ql[t] = kKeyMax;  // pad to a power of two of lists
    __syncthreads();
    cta_merge_lists(ql, L, K);  // ends with __syncthreads()
    if (tid > 0 && tid < mb.n_ranks) *reinterpret_cast<volatile u32*>(mb.ack_remote[tid]) = mb.seq;
}

// Optional outputs of the fused final stage (see scan_epilogue); all indexed by query
struct FinalOut {
    u32* done_counter;  // zero before the launch; the last CTA to finish resets it
    u64* keys;          // [nq][k_pad] merged candidates of the whole shard
    u64* idx;           // [nq][k]  or nullptr
    float* dist;        // [nq][k]
    u32* count;         // [nq]
    // One-launch searches (bsr_search / bsr_search_parallel, one query): the result goes to pinned host
    // memory as self-validating 8-byte words -- (sequence number << 32) | value -- which the host polls:
    //   host_ll[0 .. ll_k)        global row indices (they fit 32 bits, include/bsr.h)
    //   host_ll[ll_k .. 2 ll_k)   distance bit patterns
    //   host_ll[2 ll_k]           count
    // An 8-byte store arrives whole, so a word whose upper half is this search's sequence number carries
    // this search's value: no system-scope fence (a ~2 us stall in the last CTA: it waits for every
    // posted PCIe write to be acknowledged), no waiting for the grid to retire and the stream to signal.
    // ll_k == 0 (a non-root rank of a multi-rank search): only the count word, as a completion signal.
    // nullptr: results go to idx / dist / count.
    //   host_ll[2 ll_k + 1]       (t_start != nullptr) nanoseconds from the first CTA's first instruction to
    //                             this word's store: what bsr_store_last_scan_ms reports for such a launch.
    //                             Two timing events around the launch would cost a small store's search
    //                             ~1.1 us of its ~20 (measured); the first CTA leaves %globaltimer in
    //                             *t_start (device memory), the last CTA reads it back after the ticket.
    u64* host_ll;
    u32 ll_k;
    u32 ll_seq;
    unsigned long long* t_start;
    Mailbox mb;         // cross-rank exchange fused into the last CTA (k <= 128, one query); all-null: none
    unsigned long long* dbg;  // debug stamps (see ScanArgs::dbg)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void ll_store(u64* p, u32 seq, u32 value) {
    *reinterpret_cast<volatile u64*>(p) = ((u64)seq << 32) | (u64)value;
}
// slot j of the final list (key v) and, from one thread, the count
__device__ __forceinline__ void ll_publish_entry(const FinalOut& fo, u32 j, u64 v) {
    if (j >= fo.ll_k) return;
    const bool ok = v != kKeyMax;
    ll_store(fo.host_ll + j, fo.ll_seq, ok ? (u32)v : 0u);
    ll_store(fo.host_ll + fo.ll_k + j, fo.ll_seq, ok ? bits_from_ordered((u32)(v >> 32)) : 0u);
}
__device__ __forceinline__ void ll_publish_count(const FinalOut& fo, u32 count) {
    if (fo.t_start) {  // the launch's own duration (see FinalOut::t_start), ns, saturating
        const unsigned long long ns = globaltimer_ns() - *reinterpret_cast<const volatile unsigned long long*>(fo.t_start);
        ll_store(fo.host_ll + 2 * fo.ll_k + 1, fo.ll_seq, ns < 0xFFFFFFFFull ? (u32)ns : 0xFFFFFFFFu);
    }
    ll_store(fo.host_ll + 2 * fo.ll_k, fo.ll_seq, count);
}
// first thing a scan kernel does (see FinalOut::t_start)
__device__ __forceinline__ void ll_stamp_start(const FinalOut& fo) {
    if (fo.host_ll && fo.t_start && blockIdx.x == 0 && threadIdx.x == 0) *fo.t_start = globaltimer_ns();
}

#ifdef BSR_TC_TIMELINE
// CTA 0 thread 0 stamps slots 0-7; the last CTA to finish stamps 8-11 (see scan_epilogue)
#define BSR_STAMP(dbg, cond, slot) do { if ((dbg) && (cond)) (dbg)[slot] = globaltimer_ns(); } while (0)
#else
#define BSR_STAMP(dbg, cond, slot) do { } while (0)
#endif

// Last-CTA fold for 32 < K <= 128 (NB = K/32 register blocks per lane): this warp's running best K
// keys live in `best[NB]` (ascending blocks).  A list's first block is fetched with all the other
// lists' (one L2 round trip); its later blocks are fetched only while the list can still
// contribute (its previous block's largest key is below the running K-th best), which on data in
// no particular order is almost never.
template <int NB>
__device__ __forceinline__ void fold_block(u64 (&best)[NB], u64 in, int lane) {
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const u64 cmin = __shfl_sync(0xFFFFFFFFu, in, 0);
        const u64 bmax = __shfl_sync(0xFFFFFFFFu, best[j], 31);
        if (cmin < bmax) best[j] = warp_merge_split32(best[j], in, lane);  // else: all of `in` lies above block j
    }
}
template <int NB, int W>
__device__ __forceinline__ void fold_cta_lists(const u64* src, u32 n_lists, u64 (&best)[NB], int warp, int lane) {
    constexpr u32 K = NB * 32;
    constexpr u32 B = 10;  // lists per batch: their first two blocks are fetched together (one L2 round trip)
    for (u32 l0 = warp; l0 < n_lists; l0 += B * W) {
        u64 t0[B], t1[B];
#pragma unroll
        for (u32 i = 0; i < B; i++) {
            const u32 l = l0 + i * W;
            t0[i] = l < n_lists ? ldcg_u64(src + (size_t)l * K + lane) : kKeyMax;
            t1[i] = l < n_lists ? ldcg_u64(src + (size_t)l * K + 32 + lane) : kKeyMax;
        }
#pragma unroll
        for (u32 i = 0; i < B; i++) {
            const u32 l = l0 + i * W;
            if (l >= n_lists) break;  // warp-uniform
            u64 in = t0[i];
            for (int b = 0;; b++) {
                const u64 in_max = __shfl_sync(0xFFFFFFFFu, in, 31);
                fold_block<NB>(best, in, lane);
                const u64 kth = __shfl_sync(0xFFFFFFFFu, best[NB - 1], 31);
                // the list is ascending: once a block's largest key is not below the running K-th
                // best, nothing after it can enter
                if (b + 1 == NB || !(in_max < kth)) break;
                in = (b == 0) ? t1[i] : ldcg_u64(src + (size_t)l * K + (size_t)(b + 1) * 32 + lane);
            }
        }
    }
}

// End of a scan CTA that handled NQ queries: for each query reduce the W per-warp lists
// (lists[q][w][K] in shared memory) to the CTA's best K and publish them to
// cta_keys[q][cta][K]; when the launch fuses the final stage (K == 32) the last CTA to finish
// merges every CTA's list and unpacks the result, so a whole search is one kernel.
// `live` = warps of this CTA that scanned at least one panel (the others' lists are still all KEY_MAX):
// only those lists take part in the CTA-level reduce -- on a small store that is one or two of them.
template <int W, int NQ>
__device__ __forceinline__ void scan_epilogue(u64* lists, u32 K, u32 k, u64* cta_keys,
                                              const FinalOut& fo, bool fuse_final, u32* panel_ticket = nullptr,
                                              u32 live = W) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ u32 s_is_last;
    const size_t q_stride = (size_t)gridDim.x * K;  // between queries in cta_keys
    if (K == 32) {
        // warp w reduces queries w, w+W, ... entirely in registers
        for (int q = warp; q < NQ; q += W) {
            u64 v = lists[((size_t)q * W) * 32 + lane];
            for (u32 w = 1; w < live; w++) v = warp_merge32(v, lists[((size_t)q * W + w) * 32 + lane], lane);
            cta_keys[q * q_stride + (size_t)blockIdx.x * 32 + lane] = v;
        }
    } else {
        u32 live2 = 1;  // lists to reduce, rounded up to a power of two (the extra ones are empty)
        while (live2 < live) live2 <<= 1;
        for (int q = 0; q < NQ; q++) {
            u64* ql = lists + (size_t)q * W * K;
            cta_merge_lists(ql, live2, K);
            for (u32 j = tid; j < K; j += W * 32) cta_keys[q * q_stride + (size_t)blockIdx.x * K + j] = ql[j];
        }
    }
    BSR_STAMP(fo.dbg, blockIdx.x == 0 && tid == 0, 6);
    if (!fuse_final && !panel_ticket) return;

    __threadfence();
    __syncthreads();
    if (tid == 0) s_is_last = (atomicAdd(fo.done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    BSR_STAMP(fo.dbg, blockIdx.x == 0 && tid == 0, 7);
    if (!s_is_last) return;
    if (panel_ticket && tid == 0) *panel_ticket = 0;  // every CTA has drawn its last ticket: ready for the next launch
    if (!fuse_final) {
        if (tid == 0) *fo.done_counter = 0;
        return;
    }
    BSR_STAMP(fo.dbg, tid == 0, 8);
    __threadfence();
    if (K != 32) {
        // 32 < K <= 128: per query every warp folds its share of the CTA lists into K/32 register
        // blocks; the W partial results meet in shared memory and are reduced like the per-warp lists
        for (int q = 0; q < NQ; q++) {
            const u64* src = cta_keys + q * q_stride;
            u64* mine = lists + ((size_t)q * W + warp) * K;
            if (K == 64) {
                u64 best[2] = {kKeyMax, kKeyMax};
                fold_cta_lists<2, W>(src, gridDim.x, best, warp, lane);
                mine[lane] = best[0];
                mine[32 + lane] = best[1];
            } else {
                u64 best[4] = {kKeyMax, kKeyMax, kKeyMax, kKeyMax};
                fold_cta_lists<4, W>(src, gridDim.x, best, warp, lane);
#pragma unroll
                for (int j = 0; j < 4; j++) mine[j * 32 + lane] = best[j];
            }
        }
        __syncthreads();
        BSR_STAMP(fo.dbg, tid == 0, 9);
        for (int q = 0; q < NQ; q++) {
            u64* ql = lists + (size_t)q * W * K;
            cta_merge_lists(ql, W, K);  // ends with __syncthreads()
            if (NQ == 1) mailbox_exchange_cta(ql, K, k, fo.mb);  // multi-rank: send to root / fold the peers in
            u32 valid = 0;
            for (u32 j = tid; j < K; j += W * 32) {
                const u64 v = ql[j];
                fo.keys[(size_t)q * K + j] = j < k ? v : kKeyMax;  // slots past top_k are published empty
                if (fo.idx && j < k) {
                    const bool ok = v != kKeyMax;
                    fo.idx[(size_t)q * k + j] = ok ? (v & 0xFFFFFFFFull) : 0ull;
                    fo.dist[(size_t)q * k + j] = ok ? __uint_as_float(bits_from_ordered((u32)(v >> 32))) : 0.0f;
                    valid += ok ? 1u : 0u;
                }
            }
            if (fo.idx) {  // K <= 128 <= blockDim.x: at most one key per thread; warps 0-3 hold them all
                const u32 wsum = __reduce_add_sync(0xFFFFFFFFu, valid);
                __shared__ u32 s_valid[4];
                if (warp < 4 && lane == 0) s_valid[warp] = wsum;
                __syncthreads();
                if (tid == 0) fo.count[q] = s_valid[0] + s_valid[1] + s_valid[2] + s_valid[3];
                __syncthreads();
            }
            if (NQ == 1 && fo.host_ll) {
                u32 v_ok = 0;
                for (u32 j = tid; j < K; j += W * 32) {
                    const u64 v = ql[j];
                    ll_publish_entry(fo, j, v);
                    v_ok += (j < fo.ll_k && v != kKeyMax) ? 1u : 0u;
                }
                const u32 wsum = __reduce_add_sync(0xFFFFFFFFu, v_ok);
                __shared__ u32 s_ll_valid[4];
                if (warp < 4 && lane == 0) s_ll_valid[warp] = wsum;
                __syncthreads();
                if (tid == 0) ll_publish_count(fo, s_ll_valid[0] + s_ll_valid[1] + s_ll_valid[2] + s_ll_valid[3]);
            }
        }
        if (tid == 0) *fo.done_counter = 0;
        BSR_STAMP(fo.dbg, tid == 0, 10);
        return;
    }
    // K == 32: every warp folds a strided subset of the CTA lists in registers, per query
    for (int q = 0; q < NQ; q++) {
        u64 v = kKeyMax;
        const u64* src = cta_keys + q * q_stride;
        // this warp's lists are l = warp, warp+W, ...: fetch up to 20 at a time (independent L2
        // reads, one round trip for a grid of <= 160 CTAs at W = 8) and only then run the dependent
        // merges; a list whose best key is not below the current 32nd best is skipped
        for (u32 l0 = warp; l0 < gridDim.x; l0 += 20 * W) {
            u64 t[20];
#pragma unroll
            for (u32 i = 0; i < 20; i++) {
                const u32 l = l0 + i * W;
                t[i] = l < gridDim.x ? ldcg_u64(src + (size_t)l * 32 + lane) : kKeyMax;
            }
#pragma unroll
            for (u32 i = 0; i < 20; i++) {
                // only the first k keys of the result are ever read: a list whose best key is not below
                // my k-th best cannot change them
                const u64 their_best = __shfl_sync(0xFFFFFFFFu, t[i], 0);
                const u64 my_kth = __shfl_sync(0xFFFFFFFFu, v, (int)k - 1);
                if (their_best < my_kth) v = warp_merge32(v, t[i], lane);
            }
        }
        lists[((size_t)q * W + warp) * 32 + lane] = v;
    }
    __syncthreads();
    BSR_STAMP(fo.dbg, tid == 0, 9);
    for (int q = warp; q < NQ; q += W) {
        u64 v = lists[((size_t)q * W) * 32 + lane];
#pragma unroll
        for (int w = 1; w < W; w++) v = warp_merge32(v, lists[((size_t)q * W + w) * 32 + lane], lane);
        if (NQ == 1) v = mailbox_exchange32(v, fo.mb, k, lane);  // multi-rank: send to root / fold the peers in
        fo.keys[(size_t)q * 32 + lane] = (u32)lane < k ? v : kKeyMax;  // slots past top_k are published empty
        if (fo.idx) {
            const bool ok = (u32)lane < k && v != kKeyMax;
            if ((u32)lane < k) {
                fo.idx[(size_t)q * k + lane] = ok ? (v & 0xFFFFFFFFull) : 0ull;
                fo.dist[(size_t)q * k + lane] = ok ? __uint_as_float(bits_from_ordered((u32)(v >> 32))) : 0.0f;
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
            if (lane == 0) fo.count[q] = (u32)__popc(m);
        }
        if (NQ == 1 && fo.host_ll) {
            ll_publish_entry(fo, (u32)lane, v);
            const unsigned m = __ballot_sync(0xFFFFFFFFu, (u32)lane < fo.ll_k && v != kKeyMax);
            if (lane == 0) ll_publish_count(fo, (u32)__popc(m));
        }
    }
    if (tid == 0) *fo.done_counter = 0;
    BSR_STAMP(fo.dbg, tid == 0, 10);
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
    u64* cta_keys;         // [gridDim.x][k_pad]
    u32 n_rows, n_panels, dq, dim, k, k_pad;
    u32 row_base;          // global index of local row 0
    u32 fuse_final;        // 1: the last CTA also merges all CTA lists and unpacks (requires k_pad <= 128)
    u32 metric;            // bsr_metric: 0 cosine distance, 1 dot product (1 - dot)
    FinalOut fin;
    unsigned long long* dbg;  // debug builds (-DBSR_TC_TIMELINE): globaltimer stamps, else unused
};

template <int U>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_f32_kernel(const ScanArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ll_stamp_start(a.fin);
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
    WarpTopK top;
    top.init(lists + (size_t)warp * a.k_pad, a.k, a.k_pad, lane);
    __syncthreads();

    // |b| is the same for every row; each thread accumulates it (sequentially, as
    // src/metrics.rs:155 does) alongside its first panel, from the query values it reads anyway.
    float mag_b = 0.0f, qsq = 0.0f;
    bool need_qnorm = true;

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
            if (need_qnorm) {
#pragma unroll
                for (int u = 0; u < U; u++) qsq = sq4_seq(qsq, sq[j0 + u]);
            }

            if (j0 + U == dq) {
                if (need_qnorm) {
                    mag_b = __fsqrt_rn(qsq);
                    need_qnorm = false;
                }
                const u32 row = panel * kPanelRows + lane;
                const bool identical = warp_confirm_identical(
                    maybe_identical && row < a.n_rows, a.panels + (size_t)panel * panel_stride, sq,
                    dq, lane);
                u64 key = kKeyMax;
                if (row < a.n_rows)
                    key = make_key(finish_distance(acc, mag_a, mag_b, identical, a.metric), a.row_base + row);
                top.offer(key, lane);
            }
            if (!has_next) break;
#pragma unroll
            for (int u = 0; u < U; u++) cur[u] = nxt[u];
            j0 = nj0;
            panel = npanel;
        }
    }
    top.finish(lane);
    __syncthreads();
    scan_epilogue<kScanWarps, 1>(lists, a.k_pad, a.k, a.cta_keys, a.fin, a.fuse_final != 0);
}

// ------------------------------------------------------------------------------------------
// K1 (bulk-async variant): the same scan, but every warp streams its panels through a private
// shared-memory ring filled by cp.async.bulk (the TMA engine, SASS UBLKCP) and tracked by
// mbarriers.  No registers are tied up by loads in flight, so each warp keeps `stages` x 4 KiB
// outstanding (~190 KiB per SM): enough to hold HBM bandwidth even in the last, partially
// filled round of panels, which is what limits small stores.
// ------------------------------------------------------------------------------------------
constexpr int kBulkWarps = 8;
constexpr int kBulkThreads = kBulkWarps * 32;
constexpr int kChunkQuads = 8;                       // column-quads per bulk copy
constexpr int kChunkBytes = kChunkQuads * 32 * 16;   // 4 KiB

__device__ __forceinline__ u32 smem_u32(const void* p) {
    return (u32)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u32 bar, u32 parity) {
    u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ u32 mbar_try_wait_u32(u32 bar, u32 parity) {  // 1: the phase has completed
    u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ u64 make_evict_first_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// A store a few times the size of L2, scanned again and again: `keep` of its lines (chosen by address,
// so the same ones every launch) ask to stay in L2, the others to leave first -- instead of every scan
// streaming the whole store from HBM through a cache that never holds what the next scan reads first.
__device__ __forceinline__ u64 make_keep_fraction_policy(float keep) {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(keep));
    return pol;
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar, u64 policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}

// Relaxed ticket draw whose result is first needed a whole panel later.  Around an atomic add on a
// provably warp-uniform address ptxas builds its warp-aggregation pattern (vote, ATOMG by one lane,
// SHFL of the result -- also for inline PTX, also for atom.inc), and that shuffle waits for the atomic's
// L2 round trip on the spot.  Callers therefore pass `counter + z`, z being a zero the compiler cannot
// prove uniform (read from shared memory at a per-lane address).
__device__ __forceinline__ u32 draw_ticket(u32* counter) {
    u32 r;
    asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(r) : "l"(counter) : "memory");
    return r;
}

struct BulkScanArgs {
    ScanArgs s;
    u32 stages;        // ring depth per warp (>= 2)
    u32 active_warps;  // warps per CTA that own at least one panel (the ring is sized for these)
    u32 interleave;    // 1: panels dealt round-robin over CTAs first (small stores); 0: over a CTA's warps first
    // Dynamic panel dealing (large stores): every warp's first panel is static, every later one is
    // the next ticket of this counter (ticket t = panel gridDim.x * W + t).  SMs do not all see the
    // same HBM bandwidth; with a static deal the launch ends when the slowest SM does (a ~3 % spread
    // of the scan time), with tickets it ends within one panel time (~16 us).  nullptr: static deal.
    // Zero before the launch; the last CTA to finish resets it (needs ScanArgs::fin.done_counter).
    u32* panel_ticket;
    // Single host query (bsr_search, dim <= kParamQueryMax): the query travels in the kernel's
    // parameter block (constant bank) instead of a separate host-to-device copy.  scan_small_kernel only:
    // 2: ... and qparam_mag is not set, a warp without a panel computes |b|; 0: the query is ScanArgs::query
    // (device memory), |b| likewise.
    u32 query_in_params;
    // ... and its magnitude |b| (src/metrics.rs:155, the same sequential fold, done once on the host)
    // so that no warp runs the second 768-step chain next to its first panel's dot products.
    float qparam_mag;
    // scan_small_kernel, register epilogue: CTA lists as {key, list_seq} pairs (small_epilogue); 0: ticket
    u32 list_seq;
    ulonglong2* cta_ll;
    // fraction of the store's lines that ask to stay in L2 (make_keep_fraction_policy); 0: plain evict-first
    float l2_keep;
    alignas(16) float qparam[768];  // zero-padded by the host (read as float4 by scan_small_kernel)
};
constexpr u32 kParamQueryMax = 768;

// QB queries share one pass over the store: the row chunk is read from shared memory once and
// feeds QB independent sequential dot products (a.query -> [QB][dim], outputs indexed by query).
// BF16 selects the tiled bf16 layout (a.panels is then the tile base, a.dq = 16 * KB); the
// arithmetic is the same sequential binary32 fold over the widened bf16 values.
// W = warps per CTA: 8 (default), or 16 for small stores, where a second round of panels on a few
// warps would otherwise double the (latency-bound) run time.
// K32: the launch has k_pad == 32 (top_k <= 32, the headline case): every K > 32 path -- batched
// inserts, shared-memory list merges, the wide last-CTA fold -- is compiled out, which cuts the kernel
// to a fraction of its size (a small store's launch is largely instruction-cache misses otherwise).
template <int QB, bool BF16, int W = kBulkWarps, bool K32 = false>
__global__ void __launch_bounds__(W * 32, 1) scan_topk_bulk_kernel(const __grid_constant__ BulkScanArgs ba) {
    const ScanArgs& a = ba.s;
    constexpr int CH = kChunkQuads;                    // fp32: column-quads per 4 KiB chunk
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ll_stamp_start(a.fin);
    BSR_STAMP(a.dbg, blockIdx.x == 0 && threadIdx.x == 0, 0);
#ifdef BSR_TC_TIMELINE
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[12] = (unsigned long long)clock64();
#endif
    const u32 dq = a.dq, S = ba.stages, K = K32 ? 32u : a.k_pad;
    // [ring: active_warps*S*4KiB | queries: dq*QB*16, sq[j*QB + b] = column-quad j of query b (one
    //  address register per quad, the query index is an immediate offset) | lists: QB*W*K*8 |
    //  mbarriers: W*S*8]
    unsigned char* ring_all = smem_raw;
    float4* sq = reinterpret_cast<float4*>(ring_all + (size_t)ba.active_warps * S * kChunkBytes);
    u64* lists = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(sq) + (size_t)QB * dq * 16);
    u64* bars = lists + (size_t)QB * W * K;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned char* ring_b = ring_all + (size_t)warp * S * kChunkBytes;
    const u32 ring = smem_u32(ring_b);
    const u32 bar0 = smem_u32(bars + (size_t)warp * S);
    const unsigned char* store_b = reinterpret_cast<const unsigned char*>(a.panels);

    // panels are dealt round-robin over CTAs first, then over the warps of a CTA: when the panel
    // count is not a multiple of the warp count the extra panels spread over all SMs (each SM's
    // share of HBM bandwidth is what bounds it), and a small store uses as many SMs as it has panels
    const u32 total_warps = gridDim.x * W;
    const u32 panel0 = ba.interleave ? warp * gridDim.x + blockIdx.x : blockIdx.x * W + warp;
    const u32 KB = dq >> 4;                                      // bf16: 64-column blocks
    const u32 cpp = BF16 ? KB : (dq + CH - 1) / CH;              // chunks per 32-row panel
    const size_t panel_stride = (size_t)dq * 32;                 // fp32 layout, float4 units
    const u64 policy = ba.l2_keep > 0.0f ? make_keep_fraction_policy(ba.l2_keep) : make_evict_first_policy();
    const bool dyn = ba.panel_ticket != nullptr;                 // host: only with S <= cpp (see below)
    const u32 first_ticket_panel = gridDim.x * W;

    // Request cursor, warp-uniform (every lane keeps it; lane 0 alone talks to the copy engine): the
    // panel being requested, the chunk inside it, the ring slot.  The cursor moves to the next panel
    // lazily, on the first request that needs it, so after S <= cpp requests beyond the chunk being
    // consumed it is at most ONE panel ahead of the consumer -- which is how the consumer learns its
    // next panel in the dynamic deal.  Lane 0 draws the ticket for the panel after `ipanel` as soon as
    // `ipanel` is known: the atomic's round trip is hidden behind a whole panel of requests.
#ifdef BSR_TC_TIMELINE
    __shared__ u32 s_tl[48];
    if (threadIdx.x < 48) s_tl[threadIdx.x] = 0;
#endif
    __shared__ u32 s_zero[32];  // see draw_ticket
    if (tid < 32) s_zero[tid] = 0;
    u32 ipanel = panel0, ici = 0, islot = 0, inext = 0;
    u32* my_ticket = ba.panel_ticket;  // + s_zero[lane] once shared memory is initialised (below)
    auto issue = [&]() {
        if (ici == cpp) {
            ici = 0;
            if (dyn) {
                ipanel = __shfl_sync(0xFFFFFFFFu, inext, 0);
                if (lane == 0 && ipanel < a.n_panels) inext = first_ticket_panel + draw_ticket(my_ticket);
            } else {
                ipanel += total_warps;
            }
        }
        if (ipanel >= a.n_panels) return;  // nothing left to request
        if (lane == 0) {
            const void* src;
            u32 bytes;
            if (BF16) {  // this warp's 32-row half of tile (ipanel/2), column block ici
                src = store_b + ((size_t)(ipanel >> 1) * KB + ici) * kTileBytes + (ipanel & 1u) * 4096u;
                bytes = 4096;
            } else {
                src = a.panels + (size_t)ipanel * panel_stride + (size_t)ici * CH * 32;
                bytes = min((u32)CH, dq - ici * CH) * 512;
            }
            mbar_expect_tx(bar0 + islot * 8, bytes);
            bulk_g2s(ring + islot * kChunkBytes, src, bytes, bar0 + islot * 8, policy);
        }
        ici++;
        if (++islot == S) islot = 0;
    };
    if (lane == 0) {
        for (u32 st = 0; st < S; st++) mbar_init(bar0 + st * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    for (u32 i = 0; i < S; i++) issue();

    // stage the queries, zero-padded to the padded column count
    for (u32 t = tid; t < (u32)QB * dq; t += (W * 32)) {
        const u32 b = t / dq, j = t - b * dq, c = 4 * j;
        // (parameter-block queries: one query, dim <= kParamQueryMax -- the host guarantees it)
        const float* qp = ba.query_in_params ? ba.qparam : a.query + (size_t)b * a.dim;
        float4 v;
        v.x = (c + 0 < a.dim) ? qp[c + 0] : 0.0f;
        v.y = (c + 1 < a.dim) ? qp[c + 1] : 0.0f;
        v.z = (c + 2 < a.dim) ? qp[c + 2] : 0.0f;
        v.w = (c + 3 < a.dim) ? qp[c + 3] : 0.0f;
        sq[(size_t)j * QB + b] = v;  // interleaved by query: column-quad j of all QB queries is contiguous
    }
    WarpTopK top[QB];
#pragma unroll
    for (int b = 0; b < QB; b++) top[b].init(lists + ((size_t)b * W + warp) * K, a.k, K, lane);
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 1);
    __syncthreads();
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 2);
    if (dyn) {  // first draw: the ticket for this warp's second panel (the requests so far are all in the first)
        my_ticket = ba.panel_ticket + s_zero[lane];
        if (lane == 0 && panel0 < a.n_panels) inext = first_ticket_panel + draw_ticket(my_ticket);
    }

    // |b| per query: accumulated by every thread during its first panel (src/metrics.rs:155)
    float mag_b[QB], qsq[QB], acc[QB];
    bool need_qnorm = true;
#pragma unroll
    for (int b = 0; b < QB; b++) mag_b[b] = qsq[b] = acc[b] = 0.0f;
    if (QB == 1 && ba.query_in_params) {  // |b| came with the query
        mag_b[0] = ba.qparam_mag;
        need_qnorm = false;
    }

    u32 stage = 0, parity = 0, ci = 0, panel = panel0;
    float mag_a = 0.0f;
    // bit b: every column of my row seen so far is within 1e-10 of query b's.  Seeded from the first
    // columns of a panel, cleared chunk by chunk (only lanes with a live bit run the compare)
    unsigned maybe_mask = 0;
    for (u32 c = 0; panel < a.n_panels; c++) {
        if (ci == 0) {
            const u32 row = panel * kPanelRows + lane;
            mag_a = (row < a.n_rows) ? __ldg(a.norms + row) : 0.0f;
#pragma unroll
            for (int b = 0; b < QB; b++) acc[b] = 0.0f;
        }
#ifdef BSR_TC_TIMELINE  // per-chunk SM-clock stamps of CTA 0 / warp 0, kept in shared memory until the loop is over
        if (a.dbg && blockIdx.x == 0 && tid == 0 && c < 24) s_tl[2 * c] = (u32)clock64();
#endif
        mbar_wait(bar0 + stage * 8, parity);
#ifdef BSR_TC_TIMELINE
        if (a.dbg && blockIdx.x == 0 && tid == 0 && c < 24) s_tl[2 * c + 1] = (u32)clock64();
#endif
        BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0 && c == 0, 3);
        if (BF16) {
            // my row inside the 4 KiB half tile: atom lane/8, row lane%8; chunk c at (c ^ row%8)
            const unsigned char* rp = ring_b + (size_t)stage * kChunkBytes + (lane >> 3) * 1024 + (lane & 7) * 128;
            const u32 sw = lane & 7;
            const u32 q0i = ci * 16;  // first query quad of this column block
            if (ci == 0) {
                const uint4 v0 = *reinterpret_cast<const uint4*>(rp + (sw << 4));
                maybe_mask = 0;
#pragma unroll
                for (int b = 0; b < QB; b++)
                    maybe_mask |= differs8_bf16(v0, sq[b], sq[(1) * QB + b]) ? 0u : (1u << b);
            }
            if (need_qnorm) {  // first panel only: the |b| chains run next to the dot chains
#pragma unroll 2
                for (u32 cc = 0; cc < 8; cc++) {
                    const uint4 v = *reinterpret_cast<const uint4*>(rp + ((cc ^ sw) << 4));
#pragma unroll
                    for (int b = 0; b < QB; b++) {
                        const float4 qa = sq[(q0i + 2 * cc) * QB + b], qb2 = sq[(q0i + 2 * cc + 1) * QB + b];
                        acc[b] = dot8_seq_bf16(acc[b], v, qa, qb2);
                        qsq[b] = sq4_seq(sq4_seq(qsq[b], qa), qb2);
                    }
                }
            } else {
                // steady state: the whole 64-column block unrolled -- the eight swizzled chunk
                // addresses are loop-invariant per lane and the query offsets immediates
#pragma unroll
                for (u32 cc = 0; cc < 8; cc++) {
                    const uint4 v = *reinterpret_cast<const uint4*>(rp + ((cc ^ sw) << 4));
#pragma unroll
                    for (int b = 0; b < QB; b++) {
                        const float4 qa = sq[(q0i + 2 * cc) * QB + b], qb2 = sq[(q0i + 2 * cc + 1) * QB + b];
                        if (QB == 1) {
                            // one query: the eight products first, then the eight dependent adds (see the
                            // fp32 path: same values, better use of the adds' latency slots)
                            const float p0 = __fmul_rn(bf16_lo(v.x), qa.x), p1 = __fmul_rn(bf16_hi(v.x), qa.y);
                            const float p2 = __fmul_rn(bf16_lo(v.y), qa.z), p3 = __fmul_rn(bf16_hi(v.y), qa.w);
                            const float p4 = __fmul_rn(bf16_lo(v.z), qb2.x), p5 = __fmul_rn(bf16_hi(v.z), qb2.y);
                            const float p6 = __fmul_rn(bf16_lo(v.w), qb2.z), p7 = __fmul_rn(bf16_hi(v.w), qb2.w);
                            float t = acc[b];
                            t = __fadd_rn(t, p0); t = __fadd_rn(t, p1); t = __fadd_rn(t, p2); t = __fadd_rn(t, p3);
                            t = __fadd_rn(t, p4); t = __fadd_rn(t, p5); t = __fadd_rn(t, p6); t = __fadd_rn(t, p7);
                            acc[b] = t;
                        } else {
                            acc[b] = dot8_seq_bf16(acc[b], v, qa, qb2);
                        }
                    }
                }
            }
            if (maybe_mask) {
                // rare: my row has matched a query on every column so far -- compare this block too
                // (out of the hot loop: the stage is still in shared memory).  Bitwise-equal columns never
                // differ: four lop3 chains over the block first (see xor_or), the subtract-compare pass
                // only for a query some bit of which differs
#pragma unroll
                for (int b = 0; b < QB; b++) {
                    if (!((maybe_mask >> b) & 1u)) continue;
                    u32 x0 = 0, x1 = 0, x2 = 0, x3 = 0;
#pragma unroll
                    for (u32 cc = 0; cc < 8; cc++) {
                        const uint4 v = *reinterpret_cast<const uint4*>(rp + ((cc ^ sw) << 4));
                        const float4 qa = sq[(q0i + 2 * cc) * QB + b], qb2 = sq[(q0i + 2 * cc + 1) * QB + b];
                        x0 = xor_or(bf16_lo(v.x), qa.x, xor_or(bf16_hi(v.x), qa.y, x0));
                        x1 = xor_or(bf16_lo(v.y), qa.z, xor_or(bf16_hi(v.y), qa.w, x1));
                        x2 = xor_or(bf16_lo(v.z), qb2.x, xor_or(bf16_hi(v.z), qb2.y, x2));
                        x3 = xor_or(bf16_lo(v.w), qb2.z, xor_or(bf16_hi(v.w), qb2.w, x3));
                    }
                    if (!((x0 | x1) | (x2 | x3))) continue;
                    for (u32 cc = 0; cc < 8; cc++) {
                        const uint4 v = *reinterpret_cast<const uint4*>(rp + ((cc ^ sw) << 4));
                        if (differs8_bf16(v, sq[(q0i + 2 * cc) * QB + b], sq[(q0i + 2 * cc + 1) * QB + b]))
                            maybe_mask &= ~(1u << b);
                    }
                }
            }
        } else {
            const u32 j0 = ci * CH;
            const float4* sp = reinterpret_cast<const float4*>(ring_b + (size_t)stage * kChunkBytes) + lane;
            if (QB <= 2 && j0 + CH <= dq) {
                // narrow batches: pull the whole chunk into registers, then run the chains
                float4 v[CH];
#pragma unroll
                for (int u = 0; u < CH; u++) v[u] = sp[u * 32];
                if (ci == 0) {
                    maybe_mask = 0;
#pragma unroll
                    for (int b = 0; b < QB; b++) maybe_mask |= differs4(v[0], sq[b]) ? 0u : (1u << b);
                }
                if (need_qnorm) {  // first panel: the |b| chain runs interleaved with the dot chain
#pragma unroll
                    for (int u = 0; u < CH; u++) {
#pragma unroll
                        for (int b = 0; b < QB; b++) {
                            const float4 q = sq[(j0 + u) * QB + b];
                            acc[b] = dot4_seq(acc[b], v[u], q);
                            qsq[b] = sq4_seq(qsq[b], q);
                        }
                    }
                } else if (QB == 1) {
                    // One query: the only chain is the row's own 32 dependent adds.  All 32 products first
                    // (independent, separately rounded as the reference's are), then the adds in column
                    // order: the scheduler fills the adds' latency slots with the multiplies and the
                    // shared-memory reads instead of alternating dependent FMUL/FADD pairs -- what bounds a
                    // warp that has its scheduler to itself (small stores) or shares it with one other.
                    float4 pr[CH];
#pragma unroll
                    for (int u = 0; u < CH; u++) {
                        const float4 q = sq[j0 + u];
                        pr[u].x = __fmul_rn(v[u].x, q.x);
                        pr[u].y = __fmul_rn(v[u].y, q.y);
                        pr[u].z = __fmul_rn(v[u].z, q.z);
                        pr[u].w = __fmul_rn(v[u].w, q.w);
                    }
#pragma unroll
                    for (int u = 0; u < CH; u++) {
                        acc[0] = __fadd_rn(acc[0], pr[u].x);
                        acc[0] = __fadd_rn(acc[0], pr[u].y);
                        acc[0] = __fadd_rn(acc[0], pr[u].z);
                        acc[0] = __fadd_rn(acc[0], pr[u].w);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < CH; u++) {
#pragma unroll
                        for (int b = 0; b < QB; b++) acc[b] = dot4_seq(acc[b], v[u], sq[(j0 + u) * QB + b]);
                    }
                }
                if (maybe_mask) {  // rare: my row has matched a query on every column so far
#pragma unroll
                    for (int b = 0; b < QB; b++) {
                        if ((maybe_mask >> b) & 1u) {
                            // bitwise-equal columns never differ (differs4_fast): one OR-reduction over the
                            // chunk, branch-free; the subtract-compare pass only if some bit differs
                            u32 x0 = 0, x1 = 0, x2 = 0, x3 = 0;  // four independent lop3 chains (see xor_or)
#pragma unroll
                            for (int u = 0; u < CH; u++) {
                                const float4 q = sq[(j0 + u) * QB + b];
                                x0 = xor_or(v[u].x, q.x, x0);
                                x1 = xor_or(v[u].y, q.y, x1);
                                x2 = xor_or(v[u].z, q.z, x2);
                                x3 = xor_or(v[u].w, q.w, x3);
                            }
                            if ((x0 | x1) | (x2 | x3)) {
                                bool d = false;
#pragma unroll
                                for (int u = 0; u < CH; u++) d |= differs4(v[u], sq[(j0 + u) * QB + b]);
                                if (d) maybe_mask &= ~(1u << b);
                            }
                        }
                    }
                }
            } else if (j0 + CH <= dq) {
                // wide batches: QB independent chains already give the ILP; keep register use flat
                if (ci == 0) {
                    const float4 v0 = sp[0];
                    maybe_mask = 0;
#pragma unroll
                    for (int b = 0; b < QB; b++) maybe_mask |= differs4(v0, sq[b]) ? 0u : (1u << b);
                }
#pragma unroll 2
                for (int u = 0; u < CH; u++) {
                    const float4 v = sp[u * 32];
#pragma unroll
                    for (int b = 0; b < QB; b++) {
                        const float4 q = sq[(j0 + u) * QB + b];
                        acc[b] = dot4_seq(acc[b], v, q);
                        if (need_qnorm) qsq[b] = sq4_seq(qsq[b], q);
                    }
                }
                if (maybe_mask) {  // rare; out of the hot loop, the stage is still in shared memory
                    for (int u = 0; u < CH; u++) {
                        const float4 v = sp[u * 32];
#pragma unroll
                        for (int b = 0; b < QB; b++)
                            if (((maybe_mask >> b) & 1u) && differs4_fast(v, sq[(j0 + u) * QB + b])) maybe_mask &= ~(1u << b);
                    }
                }
            } else {
                for (u32 u = 0; j0 + u < dq; u++) {
                    const float4 v = sp[u * 32];
#pragma unroll
                    for (int b = 0; b < QB; b++) {
                        if (ci == 0 && u == 0) {
                            if (b == 0) maybe_mask = 0;
                            maybe_mask |= differs4(v, sq[b]) ? 0u : (1u << b);
                        }
                        acc[b] = dot4_seq(acc[b], v, sq[(j0 + u) * QB + b]);
                        if (need_qnorm) qsq[b] = sq4_seq(qsq[b], sq[(j0 + u) * QB + b]);
                        if (((maybe_mask >> b) & 1u) && differs4(v, sq[(j0 + u) * QB + b])) maybe_mask &= ~(1u << b);
                    }
                }
            }
        }
        __syncwarp();  // every lane has consumed this stage: hand it back to the copy engine
        issue();

        if (++ci == cpp) {
            if (need_qnorm) {
#pragma unroll
                for (int b = 0; b < QB; b++) mag_b[b] = __fsqrt_rn(qsq[b]);
                need_qnorm = false;
            }
            const u32 row = panel * kPanelRows + lane;
            const bool valid = row < a.n_rows;
#pragma unroll
            for (int b = 0; b < QB; b++) {
                // identity (src/metrics.rs:7-19, :148-151): every column of my row matched query b
                const bool identical = valid && ((maybe_mask >> b) & 1u);
                u64 key = kKeyMax;
                if (valid)
                    key = make_key(finish_distance(acc[b], mag_a, mag_b[b], identical, a.metric), a.row_base + row);
                top[b].offer(key, lane);
            }
            ci = 0;
            // dynamic deal: the request cursor is exactly one panel ahead by now (or past the end)
            panel = dyn ? ipanel : panel + total_warps;
        }
        if (++stage == S) {
            stage = 0;
            parity ^= 1;
        }
    }
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 4);
#ifdef BSR_TC_TIMELINE
    if (a.dbg && blockIdx.x == 0 && tid == 0) {
        a.dbg[13] = (unsigned long long)clock64();
        for (int i = 0; i < 48; i++) a.dbg[16 + i] = s_tl[i];
    }
#endif
#pragma unroll
    for (int b = 0; b < QB; b++) top[b].finish(lane);
    __syncthreads();
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 5);
    // warps that scanned a panel: with the CTA-first deal those are warps 0 .. live-1
    const u32 cta_panels = blockIdx.x < a.n_panels ? (a.n_panels - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    scan_epilogue<W, QB>(lists, K, a.k, a.cta_keys, a.fin, a.fuse_final != 0, ba.panel_ticket,
                         min(cta_panels, (u32)W));
}

// One warp's 32 keys (one per lane) -> slot[0..32) ascending, through the slots themselves: every lane
// reads all 32 keys back (broadcast reads) and counts the ones ahead of its own.  ~200 cycles of
// independent loads and compares; a bitonic sort of the same keys is 15 dependent shuffle stages (~800).
// Equal keys (only the empty key repeats: rows past the end of the store) are ordered by lane.
__device__ __forceinline__ void warp_rank_sort32(u64* slot, u64 key, int lane) {
    slot[lane] = key;
    __syncwarp();
    u32 rank = 0;
    const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(slot);  // 16-byte aligned (host layout)
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const ulonglong2 p = s2[i];
        rank += (p.x < key || (p.x == key && 2 * i < lane)) ? 1u : 0u;
        rank += (p.y < key || (p.y == key && 2 * i + 1 < lane)) ? 1u : 0u;
    }
    __syncwarp();
    slot[rank] = key;
}

// ------------------------------------------------------------------------------------------
// Register-resident epilogue of scan_small_kernel (top_k <= 128: NB = k_pad / 32 blocks of 32 keys per
// lane).  A small store's epilogue is a chain of list merges, each ~6 shuffle stages of ~55 cycles on a
// warp that has nothing else to issue -- so the merges are arranged as TREES whose independent nodes the
// compiler interleaves (straight-line code, no data-dependent skips), instead of one running list that
// takes every other list in turn: 73 CTA lists are 4 + 3 merge latencies deep instead of 10 + 7.
// ------------------------------------------------------------------------------------------
// a, b: ascending lists of NB*32 keys (block j = keys 32j .. 32j+31, lane i holds key 32j+i).
// Leaves the NB*32 smallest of the two in a, ascending.
template <int NB>
__device__ __forceinline__ void merge_keep(u64 (&a)[NB], const u64 (&b)[NB], int lane) {
    // a_i against b_(K-1-i): the K minima are the K smallest of the 2K and form a bitonic sequence
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const u64 r = __shfl_sync(0xFFFFFFFFu, b[NB - 1 - j], 31 - lane);
        a[j] = a[j] < r ? a[j] : r;
    }
    // bitonic merge: distances >= 32 are between registers, 16 .. 1 between lanes
#pragma unroll
    for (int dj = NB / 2; dj >= 1; dj >>= 1) {
#pragma unroll
        for (int j = 0; j < NB; j++) {
            if ((j & dj) == 0) {
                const u64 x = a[j], y = a[j + dj];
                a[j] = x < y ? x : y;
                a[j + dj] = x < y ? y : x;
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const u64 p = __shfl_xor_sync(0xFFFFFFFFu, a[j], d);
            const bool keep_min = (lane & d) == 0;
            a[j] = (keep_min == (a[j] < p)) ? a[j] : p;
        }
    }
}
// N lists -> t[0], pairwise: every level's merges are independent of one another
template <int NB, int N>
__device__ __forceinline__ void merge_tree_regs(u64 (&t)[N][NB], int lane) {
#pragma unroll
    for (int s = 1; s < N; s <<= 1) {
#pragma unroll
        for (int i = 0; i + s < N; i += 2 * s) merge_keep<NB>(t[i], t[i + s], lane);
    }
}
// this warp's share (lists warp, warp + W, ...) of the n_lists CTA lists in global memory -> best
template <int NB, int W, int N>
__device__ __forceinline__ void fold_cta_lists_tree(const u64* src, u32 n_lists, u64 (&best)[NB], int warp, int lane,
                                                    u32 first = 0) {
    constexpr u32 K = NB * 32;
    u64 t[N][NB];
#pragma unroll
    for (int i = 0; i < N; i++) {  // every load first: one L2 round trip
        const u32 l = (u32)warp + (first + (u32)i) * W;
#pragma unroll
        for (int j = 0; j < NB; j++) t[i][j] = l < n_lists ? ldcg_u64(src + (size_t)l * K + j * 32 + lane) : kKeyMax;
    }
    merge_tree_regs<NB, N>(t, lane);
#pragma unroll
    for (int j = 0; j < NB; j++) best[j] = t[0][j];
}

// The same over self-validating lists ({key, sequence number} pairs, see small_epilogue): every entry is
// polled until it carries this search's number -- no ticket, no fence, the data IS the signal.
template <int NB, int W, int N>
__device__ __forceinline__ void fold_cta_lists_tree_ll(const ulonglong2* src, u32 n_lists, u32 seq, u64 (&best)[NB],
                                                       int warp, int lane) {
    constexpr u32 K = NB * 32;
    u64 t[N][NB];
    // phase 1: every sequence word of my lists, all loads in flight together, until all of them match (a
    // 16-byte store arrives whole: a matching sequence word means the key beside it is this search's)
    while (true) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < N; i++) {
#pragma unroll
            for (int j = 0; j < NB; j++) {
                const u32 l = (u32)warp + (u32)i * W;
                // (a list this warp does not have: CTA 0's own, which always exists, stands in for it)
                const u64 y = ld_volatile_u64(&src[(size_t)(l < n_lists ? l : 0u) * K + j * 32 + lane].y);
                ok &= y == (u64)seq;
            }
        }
        if (__all_sync(0xFFFFFFFFu, ok)) break;
    }
    // phase 2: the keys
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const u32 l = (u32)warp + (u32)i * W;
            t[i][j] = l < n_lists ? ld_volatile_u64(&src[(size_t)l * K + j * 32 + lane].x) : kKeyMax;
        }
    }
    merge_tree_regs<NB, N>(t, lane);
#pragma unroll
    for (int j = 0; j < NB; j++) best[j] = t[0][j];
}

// The same tree across the warps of a CTA, through shared memory (16-warp CTAs: sixteen lists per lane
// do not fit their 128 registers): `n` ascending NB*32-key lists at base + i*stride -> warp 0's registers
// (and base).  Level by level, one merge per warp, a CTA barrier between levels.  Every thread of the
// CTA must call it.
template <int NB>
__device__ __forceinline__ void merge_tree_smem(u64* base, u32 stride, u32 n, u64 (&out)[NB], int warp, int lane) {
    for (u32 s = 1; s < n; s <<= 1) {
        __syncthreads();
        if ((u32)warp < (n + 2 * s - 1) / (2 * s) && (2 * s * warp + s) < n) {
            u64* pa = base + (size_t)(2 * s * warp) * stride;
            const u64* pb = base + (size_t)(2 * s * warp + s) * stride;
            u64 a[NB], b[NB];
#pragma unroll
            for (int j = 0; j < NB; j++) {
                a[j] = pa[j * 32 + lane];
                b[j] = pb[j * 32 + lane];
            }
            merge_keep<NB>(a, b, lane);
#pragma unroll
            for (int j = 0; j < NB; j++) pa[j * 32 + lane] = a[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NB; j++) out[j] = base[j * 32 + lane];
}

// lists: shared memory, on entry [W][32] -- warp w's 32 keys, ascending (warps >= live: unused) -- and W*K
// keys of room.  Same contract as scan_epilogue<W, 1> for K = NB*32 <= 128 with the final stage fused.
// cta_ll / list_seq (list_seq != 0, final stage fused): the CTA lists are published as {key, list_seq}
// pairs and CTA 0 -- every CTA of the grid is resident, one per SM -- polls them: the last-CTA ticket costs
// a store -> fence -> atomic -> load chain of three L2 round trips between the slowest CTA's list and its
// first use, this costs one.
template <int W, int NB>
__device__ __forceinline__ void small_epilogue(u64* lists, u32 live, u32 k, u64* cta_keys, const FinalOut& fo,
                                               bool fuse_final, ulonglong2* cta_ll, u32 list_seq) {
    constexpr u32 K = NB * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ u32 s_is_last;
    u64 mine[NB];  // (warp 0) the CTA's best K: this CTA's per-warp lists, one block each
    if constexpr (W > 8) {
        // pairs of 32-key lists first (in place: a pair's two lists make room for one K-key list, K <= 64),
        // then the tree over the W/2 results
        static_assert(W <= 8 || NB <= 2, "16-warp CTAs: top_k <= 64 in the register epilogue");
        if (warp < W / 2) {
            const u64 a = (u32)(2 * warp) < live ? lists[(2 * warp) * 32 + lane] : kKeyMax;
            u64 b = (u32)(2 * warp + 1) < live ? lists[(2 * warp + 1) * 32 + lane] : kKeyMax;
            __syncwarp();
            const u64 lo = warp_merge_split32(a, b, lane);  // b: the upper 32
            lists[(size_t)warp * 64 + lane] = lo;
            if (NB > 1) lists[(size_t)warp * 64 + 32 + lane] = b;
        }
        merge_tree_smem<NB>(lists, 64, W / 2, mine, warp, lane);
    } else if (warp == 0) {
        u64 t[W][NB];
#pragma unroll
        for (int w = 0; w < W; w++) {
            t[w][0] = (u32)w < live ? lists[w * 32 + lane] : kKeyMax;
#pragma unroll
            for (int j = 1; j < NB; j++) t[w][j] = kKeyMax;
        }
        if (live > 1) merge_tree_regs<NB, W>(t, lane);
#pragma unroll
        for (int j = 0; j < NB; j++) mine[j] = t[0][j];
    }
    if (warp == 0) {
        if (list_seq) {
#pragma unroll
            for (int j = 0; j < NB; j++) st_volatile_v2(cta_ll + (size_t)blockIdx.x * K + j * 32 + lane, mine[j], (u64)list_seq);
        } else {
#pragma unroll
            for (int j = 0; j < NB; j++) cta_keys[(size_t)blockIdx.x * K + j * 32 + lane] = mine[j];
        }
        if (fuse_final && !list_seq) {
            // release: the stores above (ordered before lane 0's atomic by the warp barrier) are visible to
            // whoever reads the counter's final value; acquire: the last CTA sees every other CTA's list
            __syncwarp();
            if (lane == 0) {
                u32 ticket;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(fo.done_counter) : "memory");
                s_is_last = (ticket == gridDim.x - 1) ? 1u : 0u;
            }
        }
    }
    BSR_STAMP(fo.dbg, blockIdx.x == 0 && tid == 0, 6);
    if (!fuse_final) return;
    if (list_seq) {
        if (blockIdx.x != 0) return;
    } else {
        __syncthreads();
        BSR_STAMP(fo.dbg, blockIdx.x == 0 && tid == 0, 7);
        if (!s_is_last) return;
    }
    BSR_STAMP(fo.dbg, tid == 0, 8);
    // last CTA (or CTA 0): every warp folds its share of the CTA lists, the W partial results meet in shared memory
    u64 best[NB];
    if (NB <= 2 && list_seq) {  // host: only with NB <= 2
        if constexpr (NB <= 2) {
            if (gridDim.x <= 10u * W) fold_cta_lists_tree_ll<NB, W, 10>(cta_ll, gridDim.x, list_seq, best, warp, lane);
            else if constexpr (W <= 8) fold_cta_lists_tree_ll<NB, W, 20>(cta_ll, gridDim.x, list_seq, best, warp, lane);
        }
        __syncthreads();  // warp 0 has read the per-warp lists this CTA's partial results overwrite
    } else if (gridDim.x <= 10u * W) {
        fold_cta_lists_tree<NB, W, 10>(cta_keys, gridDim.x, best, warp, lane);
    } else if constexpr (NB <= 2 && W <= 8) {
        fold_cta_lists_tree<NB, W, 20>(cta_keys, gridDim.x, best, warp, lane);  // host: gridDim.x <= 20 W
    } else if constexpr (W <= 8) {
        // 20 lists of 4 blocks are too many registers for one tree: two trees of 10, then one more merge
        u64 more[NB];
        fold_cta_lists_tree<NB, W, 10>(cta_keys, gridDim.x, best, warp, lane);
        fold_cta_lists_tree<NB, W, 10>(cta_keys, gridDim.x, more, warp, lane, 10);
        merge_keep<NB>(best, more, lane);
    } else {
#pragma unroll
        for (int j = 0; j < NB; j++) best[j] = kKeyMax;
        fold_cta_lists<NB, W>(cta_keys, gridDim.x, best, warp, lane);
    }
#pragma unroll
    for (int j = 0; j < NB; j++) lists[(size_t)warp * K + j * 32 + lane] = best[j];
    __syncthreads();
    BSR_STAMP(fo.dbg, tid == 0, 9);
    u64 fin[NB];  // (warp 0) the shard's best K
    if constexpr (W > 8) {
        merge_tree_smem<NB>(lists, K, W, fin, warp, lane);
    } else if (warp == 0) {
        u64 t[W][NB];
#pragma unroll
        for (int w = 0; w < W; w++) {
#pragma unroll
            for (int j = 0; j < NB; j++) t[w][j] = lists[(size_t)w * K + j * 32 + lane];
        }
        merge_tree_regs<NB, W>(t, lane);
#pragma unroll
        for (int j = 0; j < NB; j++) fin[j] = t[0][j];
    }
    if (warp == 0) {
        if (NB == 1) {
            const u64 v = mailbox_exchange32(fin[0], fo.mb, k, lane);  // multi-rank: send to root / fold the peers in
            fo.keys[lane] = (u32)lane < k ? v : kKeyMax;  // slots past top_k are published empty
            if (fo.idx) {
                const bool ok = (u32)lane < k && v != kKeyMax;
                if ((u32)lane < k) {
                    fo.idx[lane] = ok ? (v & 0xFFFFFFFFull) : 0ull;
                    fo.dist[lane] = ok ? __uint_as_float(bits_from_ordered((u32)(v >> 32))) : 0.0f;
                }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
                if (lane == 0) fo.count[0] = (u32)__popc(m);
            }
            if (fo.host_ll) {
                ll_publish_entry(fo, (u32)lane, v);
                const unsigned m = __ballot_sync(0xFFFFFFFFu, (u32)lane < fo.ll_k && v != kKeyMax);
                if (lane == 0) ll_publish_count(fo, (u32)__popc(m));
            }
            if (lane == 0 && !list_seq) *fo.done_counter = 0;
        } else {
            // 32 < K <= 128: the final list goes on through shared memory (the cross-rank exchange is
            // CTA-wide); it overwrites partial list 0, which only this warp has read
#pragma unroll
            for (int j = 0; j < NB; j++) lists[j * 32 + lane] = fin[j];
        }
    }
    if (NB == 1) {
        BSR_STAMP(fo.dbg, tid == 0, 10);
        return;
    }
    __syncthreads();
    u64* ql = lists;
    mailbox_exchange_cta(ql, K, k, fo.mb);  // multi-rank: send to root / fold the peers in
    u32 valid = 0, v_ok = 0;
    if ((u32)tid < K) {  // K <= 128 <= blockDim.x: one key per thread, warps 0-3 hold them all
        const u64 v = ql[tid];
        const bool ok = v != kKeyMax;
        fo.keys[tid] = (u32)tid < k ? v : kKeyMax;  // slots past top_k are published empty
        if (fo.idx && (u32)tid < k) {
            fo.idx[tid] = ok ? (v & 0xFFFFFFFFull) : 0ull;
            fo.dist[tid] = ok ? __uint_as_float(bits_from_ordered((u32)(v >> 32))) : 0.0f;
            valid = ok ? 1u : 0u;
        }
        if (fo.host_ll) {
            ll_publish_entry(fo, (u32)tid, v);
            v_ok = ((u32)tid < fo.ll_k && ok) ? 1u : 0u;
        }
    }
    __shared__ u32 s_cnt[2][4];
    const u32 wsum = __reduce_add_sync(0xFFFFFFFFu, valid), lsum = __reduce_add_sync(0xFFFFFFFFu, v_ok);
    if (warp < 4 && lane == 0) {
        s_cnt[0][warp] = wsum;
        s_cnt[1][warp] = lsum;
    }
    __syncthreads();
    if (tid == 0) {
        if (fo.idx) fo.count[0] = s_cnt[0][0] + s_cnt[0][1] + s_cnt[0][2] + s_cnt[0][3];
        if (fo.host_ll) ll_publish_count(fo, s_cnt[1][0] + s_cnt[1][1] + s_cnt[1][2] + s_cnt[1][3]);
        if (!list_seq) *fo.done_counter = 0;
    }
    BSR_STAMP(fo.dbg, tid == 0, 10);
}

// ------------------------------------------------------------------------------------------
// K1s: the scan for stores of at most ONE panel per warp (n_panels <= W * gridDim.x: up to ~38 k
//      fp32 rows), one host query in the parameter block.  Same arithmetic, same layout, same
//      epilogue as scan_topk_bulk_kernel<1, false, W>; what differs is that nothing here is built to
//      stream.  A small store is a latency chain -- launch, first bytes, the reference's 768
//      dependent adds per row (src/metrics.rs:152-156), the last-CTA fold -- and a warp that has its
//      scheduler to itself pays every stall of the streaming loop in full (~800 cycles per 4 KiB chunk
//      measured, 128 of them the add chain).  So:
//        * every request of the CTA's first ring fill is issued by a different thread, spread over all
//          W warps (the streaming kernel's lane 0 needs ~2 us for a panel's 24 requests);
//        * a warp knows its one panel: no cursor, no ticket, no panel loop;
//        * chunk c+1 is waited for and read into registers BEFORE chunk c's chain runs, so the
//          barrier's try_wait and the shared-memory latency sit under the adds;
//        * WHOLE: the ring holds the whole panel -- no refills, no slot arithmetic;
//        * the warp's 32 keys are sorted in registers (one panel: nothing to insert into), and for
//          top_k <= 128 the whole epilogue stays in registers (small_epilogue).
// ------------------------------------------------------------------------------------------
template <int W, bool WHOLE, int NB>
__global__ void __launch_bounds__(W * 32, 1) scan_small_kernel(const __grid_constant__ BulkScanArgs ba) {
    const ScanArgs& a = ba.s;
    constexpr int CH = kChunkQuads;
    constexpr u32 kChunkF4 = kChunkBytes / 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ll_stamp_start(a.fin);
    BSR_STAMP(a.dbg, blockIdx.x == 0 && threadIdx.x == 0, 0);
#ifdef BSR_TC_TIMELINE
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[12] = (unsigned long long)clock64();
    __shared__ u32 s_tl[48];
    if (threadIdx.x < 48) s_tl[threadIdx.x] = 0;
#endif
    // NB: k_pad / 32 for k_pad <= 128 (register epilogue, final stage fused); 0: any k_pad, scan_epilogue
    const u32 dq = a.dq, S = ba.stages, K = NB ? 32u * NB : a.k_pad, cpp = dq / CH;  // host: dq % CH == 0, S <= cpp
    // [ring: active_warps*S*4KiB | query: dq*16 | lists: W*K*8 | mbarriers: W*S*8] (as the streaming kernel)
    unsigned char* ring_all = smem_raw;
    float4* sq = reinterpret_cast<float4*>(ring_all + (size_t)ba.active_warps * S * kChunkBytes);
    u64* lists = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(sq) + (size_t)dq * 16);
    u64* bars = lists + (size_t)W * K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // this CTA's panels are blockIdx.x, blockIdx.x + gridDim.x, ...: warp w owns the w-th of them
    const u32 live = min((a.n_panels - 1 - blockIdx.x) / gridDim.x + 1, (u32)W);
    const u64 policy = ba.l2_keep > 0.0f ? make_keep_fraction_policy(ba.l2_keep) : make_evict_first_policy();

    // First ring fill: request r = (slot, warp), slot-major so that every warp's chunk 0 is asked for
    // first; thread (warp r % W, lane r / W) initialises the request's barrier, arms it and issues the copy
    {
        const u32 r = (u32)lane * W + (u32)warp;
        if (r < live * S) {
            const u32 slot = r / live, w = r - slot * live;
            const u32 bar = smem_u32(bars + (size_t)w * S + slot);
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const u32 p = w * gridDim.x + blockIdx.x;
            mbar_expect_tx(bar, kChunkBytes);
            bulk_g2s(smem_u32(ring_all + ((size_t)w * S + slot) * kChunkBytes),
                     a.panels + (size_t)p * dq * 32 + (size_t)slot * kChunkF4, kChunkBytes, bar, policy);
        }
    }
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 14);
    // The query into shared memory, zero-padded to the padded column count (the host pads the parameter
    // block).  Reading a parameter block that is cold every launch is slow -- per-lane addresses in the
    // constant bank are served one by one: 1.3 us with four scalar loads per thread, 0.8 us with one float4
    // per thread -- so with enough idle warps (`prog`) only the first four chunks' columns are staged before
    // the scan warps start (32/W lanes of every warp: no warp waits for more than one constant-cache line)
    // and the warps without a panel stage the rest next to the first chunks' adds (named barrier 1, see the
    // loop).  query_in_params == 0: a query in device memory (bsr_search_device), read column by column --
    // it need not be 16-byte aligned; |b| then always comes from the idle warp (the host makes sure of one).
    auto query_quad = [&](u32 j) -> float4 {
        if (ba.query_in_params) return reinterpret_cast<const float4*>(ba.qparam)[j];
        const u32 c = 4 * j;
        float4 v;
        v.x = (c + 0 < a.dim) ? __ldg(a.query + c + 0) : 0.0f;
        v.y = (c + 1 < a.dim) ? __ldg(a.query + c + 1) : 0.0f;
        v.z = (c + 2 < a.dim) ? __ldg(a.query + c + 2) : 0.0f;
        v.w = (c + 3 < a.dim) ? __ldg(a.query + c + 3) : 0.0f;
        return v;
    };
    const bool has = (u32)warp < live;
    const bool prog = live <= 3 && cpp > 4;
    if (prog) {
        if (lane < 32 / W) sq[warp * (32 / W) + lane] = query_quad(warp * (32 / W) + lane);
    } else {
        for (u32 j = tid; j < dq; j += (W * 32)) sq[j] = query_quad(j);
    }
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 15);
    const u32 panel = (u32)warp * gridDim.x + blockIdx.x;
    const u32 row = panel * kPanelRows + lane;
    const bool valid = has && row < a.n_rows;
    const float mag_a = valid ? __ldg(a.norms + row) : 0.0f;
    // a warp's result is its panel's 32 keys, sorted: lists[warp][32] for the register epilogue, a K-slot
    // list padded with empty slots for scan_epilogue
    const size_t my_list = (size_t)warp * (NB ? 32u : K);
    if (!NB) {
        for (u32 j = 32 + lane; j < K; j += 32) lists[my_list + j] = kKeyMax;
        if (!has) lists[my_list + lane] = kKeyMax;
    }
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 1);
    __syncthreads();
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 2);

    __shared__ float s_mag_b;
    float acc_out = 0.0f;
    bool maybe_out = false;
    if (prog && !has) {  // the rest of the query, by the warps without a panel
        for (u32 j = 32 + ((u32)warp - live) * 32 + lane; j < dq; j += (W - live) * 32) sq[j] = query_quad(j);
        if (warp == W - 1) asm volatile("bar.sync 1, %0;" ::"n"(W * 32) : "memory");  // it reads the whole query next
        else asm volatile("bar.arrive 1, %0;" ::"n"(W * 32) : "memory");
    }
    if (!has && ba.query_in_params != 1 && warp == W - 1) {
        // |b| (src/metrics.rs:155, the same sequential fold) by a warp that has no panel: the host would
        // spend ~1.3 us on these 768 dependent adds before it could launch; here they run next to the scan
        float qs = 0.0f;
        for (u32 j = 0; j < dq; j++) qs = sq4_seq(qs, sq[j]);  // the zero padding adds +0.0: exact
        if (lane == 0) s_mag_b = __fsqrt_rn(qs);
    }
    if (has) {
        const float4* sp = reinterpret_cast<const float4*>(ring_all + (size_t)warp * S * kChunkBytes) + lane;
        const u32 ring = smem_u32(ring_all + (size_t)warp * S * kChunkBytes);
        const u32 bar0 = smem_u32(bars + (size_t)warp * S);
        const float4* src = a.panels + (size_t)panel * dq * 32;
        // (slot, phase parity) of chunks c, c+1 and c+2; past the last chunk they stay on it (its phase stays
        // complete), which keeps the loop body free of end-of-panel branches
        auto adv = [&](u32& sl, u32& pa) {
            if (++sl == S) {  // WHOLE: S == cpp, never reached
                sl = 0;
                pa ^= 1u;
            }
        };
        u32 c = 0, s0 = 0, p0 = 0, s1 = 0, p1 = 0;
        if (cpp > 1) adv(s1, p1);
        u32 s2 = s1, p2 = p1;
        if (cpp > 2) adv(s2, p2);
        bool maybe = false;  // every column of my row seen so far is within 1e-10 of the query's (src/metrics.rs:7-19)
        float acc = 0.0f;
        // One chunk into registers and straight into its 32 products (independent, each rounded as the
        // reference's); lanes whose row has matched the query so far compare it while it is there.
        auto fetch = [&](float4(&pr)[CH], const float4* vp, const float4* qp, bool first) {
            float4 v[CH], q[CH];
#pragma unroll
            for (int u = 0; u < CH; u++) v[u] = vp[u * 32];
#pragma unroll
            for (int u = 0; u < CH; u++) q[u] = qp[u];
#pragma unroll
            for (int u = 0; u < CH; u++) {
                pr[u].x = __fmul_rn(v[u].x, q[u].x);
                pr[u].y = __fmul_rn(v[u].y, q[u].y);
                pr[u].z = __fmul_rn(v[u].z, q[u].z);
                pr[u].w = __fmul_rn(v[u].w, q[u].w);
            }
            if (first) maybe = !differs4(v[0], q[0]);
            // WHOLE: the panel stays in shared memory -- flagged rows are verified once, after the loop
            if (!WHOLE && maybe) {  // rare (normally one lane of one warp: the row that IS the query)
                u32 x0 = 0, x1 = 0, x2 = 0, x3 = 0;  // bitwise-equal columns never differ: four lop3 chains
#pragma unroll
                for (int u = 0; u < CH; u++) {
                    x0 = xor_or(v[u].x, q[u].x, x0);
                    x1 = xor_or(v[u].y, q[u].y, x1);
                    x2 = xor_or(v[u].z, q[u].z, x2);
                    x3 = xor_or(v[u].w, q[u].w, x3);
                }
                if ((x0 | x1) | (x2 | x3)) {
                    bool d = false;
#pragma unroll
                    for (int u = 0; u < CH; u++) d |= differs4(v[u], q[u]);
                    if (d) maybe = false;
                }
            }
        };
        float4 pa[CH], pb[CH];
        mbar_wait(bar0, 0);
        BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 3);
        fetch(pa, sp, sq, true);
        // The barrier of chunk c+1 is polled one step early (non-blocking try_wait whose result is read a
        // step later), so the poll's latency sits under a chain as well.
        u32 ok = mbar_try_wait_u32(bar0 + s1 * 8, p1);
        // Step c: chunk c+1 is fetched and multiplied while chunk c's products are added in column order --
        // the adds are a pure register chain (no scoreboard in it), everything that waits on shared memory
        // runs in its latency slots.  The chain is made to depend on the wait's result, a constant 1 the
        // compiler cannot know: otherwise ptxas hoists the whole chain above the wait and the reads end up
        // behind it, every latency exposed.
        auto step = [&](float4(&pr)[CH], float4(&pn)[CH]) {
#ifdef BSR_TC_TIMELINE
            if (a.dbg && blockIdx.x == 0 && tid == 0 && c < 24) s_tl[2 * c] = s_tl[2 * c + 1] = (u32)clock64();
#endif
            // (the columns of chunks 4 and up: staged by the idle warps meanwhile)
            if (prog && c == 3) asm volatile("bar.sync 1, %0;" ::"n"(W * 32) : "memory");
            while (!ok) ok = mbar_try_wait_u32(bar0 + s1 * 8, p1);
            acc = __uint_as_float(__float_as_uint(acc) | (ok ^ 1u));
            const u32 ok2 = mbar_try_wait_u32(bar0 + s2 * 8, p2);
            fetch(pn, sp + (size_t)s1 * kChunkF4, sq + (size_t)min(c + 1, cpp - 1) * CH, false);
#pragma unroll
            for (int u = 0; u < CH; u++) {
                acc = __fadd_rn(acc, pr[u].x);
                acc = __fadd_rn(acc, pr[u].y);
                acc = __fadd_rn(acc, pr[u].z);
                acc = __fadd_rn(acc, pr[u].w);
            }
            if (!WHOLE) {
                __syncwarp();  // chunk c's slot was read a step ago: hand it back
                if (lane == 0 && c + S < cpp) {
                    mbar_expect_tx(bar0 + s0 * 8, kChunkBytes);
                    bulk_g2s(ring + s0 * kChunkBytes, src + (size_t)(c + S) * kChunkF4, kChunkBytes, bar0 + s0 * 8, policy);
                }
            }
            c++;
            s0 = s1;
            p0 = p1;
            s1 = s2;
            p1 = p2;
            if (c + 2 < cpp) adv(s2, p2);
            ok = ok2;
        };
        while (c < cpp) {
            step(pa, pb);
            if (c < cpp) step(pb, pa);
        }
        BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 4);
#ifdef BSR_TC_TIMELINE
        if (a.dbg && blockIdx.x == 0 && tid == 0) {
            a.dbg[13] = (unsigned long long)clock64();
            for (int i = 0; i < 48; i++) a.dbg[16 + i] = s_tl[i];
        }
#endif
        if (WHOLE) {
            // identity (src/metrics.rs:7-19) of the rows whose first four columns matched: the whole warp
            // checks one flagged row at a time against the query, 32 column-quads per step
            unsigned m = __ballot_sync(0xFFFFFFFFu, maybe);
            maybe = false;
            while (m) {
                const int r = __ffs(m) - 1;
                m &= m - 1;
                bool diff = false;
                for (u32 j = lane; j < dq; j += 32) diff |= differs4_fast(sp[(size_t)j * 32 + r - lane], sq[j]);
                if (!__any_sync(0xFFFFFFFFu, diff) && lane == r) maybe = true;
            }
        }
        acc_out = acc;
        maybe_out = maybe;
    }
    float mag_b = ba.qparam_mag;
    if (ba.query_in_params != 1) {  // |b| from this CTA's idle warp (above)
        __syncthreads();
        mag_b = s_mag_b;
    }
    if (has) {
        u64 key = kKeyMax;
        if (valid) key = make_key(finish_distance(acc_out, mag_a, mag_b, maybe_out, a.metric), a.row_base + row);
        warp_rank_sort32(lists + my_list, key, lane);
    }
    __syncthreads();
    BSR_STAMP(a.dbg, blockIdx.x == 0 && tid == 0, 5);
    if constexpr (NB != 0) small_epilogue<W, NB>(lists, live, a.k, a.cta_keys, a.fin, a.fuse_final != 0, ba.cta_ll, ba.list_seq);
    else scan_epilogue<W, 1>(lists, K, a.k, a.cta_keys, a.fin, a.fuse_final != 0, nullptr, live);
}

#ifndef BSR_TEMPLATES_ONLY  // the plain (non-template) kernels are compiled once, in bsr_api.cu
// ------------------------------------------------------------------------------------------
// K3: merge candidate lists (cross-CTA after the scan; cross-rank after the all-gather).
//     restates compute_global_top_k's sort + first-k (src/mpi_helpers/metrics.rs:147-170).
//   list l of query q starts at in + l*list_stride + q*query_stride; CTA (g, q) reduces lists
//   [g*lists_per_cta, (g+1)*lists_per_cta) to out[(q*gridDim.x + g) * K].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMergeThreads) merge_lists_kernel(
    const u64* __restrict__ in, u32 n_lists, u32 K, u32 lists_per_cta, size_t list_stride,
    size_t query_stride, u64* __restrict__ out, u32 k_keep) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
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
    // k_keep < K on the last level of a tree: slots past top_k are published empty (include/bsr.h)
    for (u32 j = threadIdx.x; j < K; j += blockDim.x) o[j] = j < k_keep ? buf[j] : kKeyMax;
}

// (idx, dist) pairs -> candidate keys, cut into lists of K slots; a list must be ascending for the
// merge, so each list is sorted here by one CTA (bitonic, K <= 1024 keys in shared memory).
__global__ void pack_sort_lists_kernel(const u64* __restrict__ idx, const float* __restrict__ dist,
                                       u64 n, u32 K, u64* __restrict__ lists) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u64* buf = reinterpret_cast<u64*>(smem_raw);
    const u64 base = (u64)blockIdx.x * K;
    for (u32 j = threadIdx.x; j < K; j += blockDim.x) {
        const u64 i = base + j;
        buf[j] = (i < n) ? make_key(dist[i], (u32)idx[i]) : kKeyMax;
    }
    __syncthreads();
    for (u32 size = 2; size <= K; size <<= 1) {
        for (u32 d = size >> 1; d >= 1; d >>= 1) {
            const u32 ld = 31 - __clz(d);
            for (u32 t = threadIdx.x; t < (K >> 1); t += blockDim.x) {
                const u32 lo = ((t >> ld) << (ld + 1)) + (t & (d - 1)), hi = lo + d;
                const bool asc = (lo & size) == 0;
                const u64 x = buf[lo], y = buf[hi];
                if ((x > y) == asc) {
                    buf[lo] = y;
                    buf[hi] = x;
                }
            }
            __syncthreads();
        }
    }
    for (u32 j = threadIdx.x; j < K; j += blockDim.x) lists[base + j] = buf[j];
}

// K3 for 32-slot lists (k <= 32), e.g. the cross-rank merge after the all-gather: one warp per
// query folds all lists in registers (lane i holds a list's i-th key) and, when asked, unpacks the
// result in the same launch.  list l of query q starts at in + l*list_stride + q*query_stride.
__global__ void __launch_bounds__(32) merge32_kernel(const u64* __restrict__ in, u32 n_lists, size_t list_stride,
                                                     size_t query_stride, u32 k, u64* __restrict__ out_keys,
                                                     u64* __restrict__ idx, float* __restrict__ dist,
                                                     u32* __restrict__ count) {
    const u32 q = blockIdx.x;
    const int lane = threadIdx.x;
    const u64* src = in + (size_t)q * query_stride;
    u64 v = kKeyMax;
    for (u32 l0 = 0; l0 < n_lists; l0 += 16) {
        u64 t[16];
#pragma unroll
        for (u32 i = 0; i < 16; i++) t[i] = (l0 + i < n_lists) ? src[(size_t)(l0 + i) * list_stride + lane] : kKeyMax;
#pragma unroll
        for (u32 i = 0; i < 16; i++) {
            const u64 their_best = __shfl_sync(0xFFFFFFFFu, t[i], 0);
            const u64 my_worst = __shfl_sync(0xFFFFFFFFu, v, 31);
            if (their_best < my_worst) v = warp_merge32(v, t[i], lane);
        }
    }
    if (out_keys) out_keys[(size_t)q * 32 + lane] = (u32)lane < k ? v : kKeyMax;
    if (idx) {
        const bool ok = (u32)lane < k && v != kKeyMax;
        if ((u32)lane < k) {
            idx[(size_t)q * k + lane] = ok ? (v & 0xFFFFFFFFull) : 0ull;
            dist[(size_t)q * k + lane] = ok ? __uint_as_float(bits_from_ordered((u32)(v >> 32))) : 0.0f;
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
        if (lane == 0) count[q] = (u32)__popc(m);
    }
}

// The mailbox exchange for a rank that has nothing of its own to contribute (empty shard, or a local
// scan that failed: the reference's empty list, metrics.rs:185-191): send an empty list / fold the
// peers' lists, then publish like the scan's last CTA does.  One warp.
__global__ void __launch_bounds__(256) mailbox_only_kernel(const FinalOut fo, u32 k, u32 K) {
    __shared__ u64 ql[8 * 128];  // up to 8 ranks' lists of K <= 128 keys (more ranks: K == 32 only, host-checked)
    const u32 tid = threadIdx.x;
    if (K == 32) {
        if (tid >= 32) return;  // one warp
        const int lane = (int)tid;
        const u64 v = mailbox_exchange32(kKeyMax, fo.mb, k, lane);
        if (fo.keys) fo.keys[lane] = (u32)lane < k ? v : kKeyMax;
        if (fo.idx) {
            const bool ok = (u32)lane < k && v != kKeyMax;
            if ((u32)lane < k) {
                fo.idx[lane] = ok ? (v & 0xFFFFFFFFull) : 0ull;
                fo.dist[lane] = ok ? __uint_as_float(bits_from_ordered((u32)(v >> 32))) : 0.0f;
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
            if (lane == 0) fo.count[0] = (u32)__popc(m);
        }
        if (fo.host_ll) {
            ll_publish_entry(fo, (u32)lane, v);
            const unsigned m = __ballot_sync(0xFFFFFFFFu, (u32)lane < fo.ll_k && v != kKeyMax);
            if (lane == 0) ll_publish_count(fo, (u32)__popc(m));
        }
        return;
    }
    for (u32 j = tid; j < K; j += blockDim.x) ql[j] = kKeyMax;
    __syncthreads();
    mailbox_exchange_cta(ql, K, k, fo.mb);
    u32 valid = 0;
    for (u32 j = tid; j < K; j += blockDim.x) {
        const u64 v = ql[j];
        if (fo.keys) fo.keys[j] = j < k ? v : kKeyMax;
        if (fo.idx && j < k) {
            const bool ok = v != kKeyMax;
            fo.idx[j] = ok ? (v & 0xFFFFFFFFull) : 0ull;
            fo.dist[j] = ok ? __uint_as_float(bits_from_ordered((u32)(v >> 32))) : 0.0f;
            valid += ok ? 1u : 0u;
        }
    }
    if (fo.idx) {
        __shared__ u32 s_valid;
        if (tid == 0) s_valid = 0;
        __syncthreads();
        if (valid) atomicAdd(&s_valid, valid);
        __syncthreads();
        if (tid == 0) fo.count[0] = s_valid;
    }
    if (fo.host_ll) {
        __shared__ u32 s_ll;
        if (tid == 0) s_ll = 0;
        __syncthreads();
        u32 v_ok = 0;
        for (u32 j = tid; j < K; j += blockDim.x) {
            ll_publish_entry(fo, j, ql[j]);
            v_ok += (j < fo.ll_k && ql[j] != kKeyMax) ? 1u : 0u;
        }
        if (v_ok) atomicAdd(&s_ll, v_ok);
        __syncthreads();
        if (tid == 0) ll_publish_count(fo, s_ll);
    }
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

// ---- BF16 store (tile layout above) ------------------------------------------------------------
// thread t owns destination 16-byte chunk t of the tile range starting at tile row_lo/64
__device__ __forceinline__ void bf16_chunk_coords(size_t t, u32 first_tile, u32 KB, u32& row, u32& col) {
    const u32 per_tile = KB * 512;  // chunks per 64-row tile
    const u32 tile = first_tile + (u32)(t / per_tile);
    const u32 rem = (u32)(t % per_tile);
    const u32 kb = rem >> 9, in = rem & 511u;
    const u32 atom = in >> 6, i = (in >> 3) & 7u, pos = in & 7u;
    row = tile * 64 + atom * 8 + i;
    col = kb * 64 + ((pos ^ i) << 3);
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
    uint4 o;
    o.x = (u32)f32_to_bf16_rne(v[0]) | ((u32)f32_to_bf16_rne(v[1]) << 16);
    o.y = (u32)f32_to_bf16_rne(v[2]) | ((u32)f32_to_bf16_rne(v[3]) << 16);
    o.z = (u32)f32_to_bf16_rne(v[4]) | ((u32)f32_to_bf16_rne(v[5]) << 16);
    o.w = (u32)f32_to_bf16_rne(v[6]) | ((u32)f32_to_bf16_rne(v[7]) << 16);
    return o;
}
__global__ void interleave_rows_bf16_kernel(const float* __restrict__ src, uint4* __restrict__ tiles,
                                            u32 row_lo, u32 n, u32 dim, u32 KB) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 row, col;
    bf16_chunk_coords(t, row_lo >> 6, KB, row, col);
    if (row < row_lo || row >= row_lo + n) return;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = (col + e < dim) ? src[(size_t)(row - row_lo) * dim + col + e] : 0.0f;
    tiles[(size_t)(row_lo >> 6) * KB * 512 + t] = pack8_bf16(v);
}
__global__ void synth_rows_bf16_kernel(uint4* __restrict__ tiles, u32 row_lo, u32 n, u32 dim, u32 KB,
                                       u64 seed, u64 gen_first_row);
__global__ void gather_rows_bf16_kernel(const unsigned char* __restrict__ tiles, float* __restrict__ out,
                                        u32 offset, u32 length, u32 dim, u32 KB) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)length * dim) return;
    const u32 r = offset + (u32)(t / dim), c = (u32)(t % dim);
    const unsigned short h = *reinterpret_cast<const unsigned short*>(
        tiles + bf16_chunk_offset(r, c >> 6, (c & 63u) >> 3, KB) + (c & 7u) * 2);
    out[t] = __uint_as_float((u32)h << 16);
}
// |a| over the bf16-rounded values, sequential (src/metrics.rs:154 on the rounded row)
__global__ void row_norms_bf16_kernel(const unsigned char* __restrict__ tiles, float* __restrict__ norms,
                                      float* __restrict__ inv_norms, u32 row_lo, u32 row_hi, u32 KB) {
    const u32 row = row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row_hi) return;
    float s = 0.0f;
    for (u32 g = 0; g < KB * 8; g++) {
        const uint4 a = *reinterpret_cast<const uint4*>(tiles + bf16_chunk_offset(row, g >> 3, g & 7u, KB));
        const u32 w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
            s = __fadd_rn(s, __fmul_rn(lo, lo));
            s = __fadd_rn(s, __fmul_rn(hi, hi));
        }
    }
    const float na = __fsqrt_rn(s);
    norms[row] = na;
    // scale for the tensor path's bound test only (exact distances always divide by |a|)
    inv_norms[row] = (na > 0.0f && na < INFINITY) ? __fdiv_rn(1.0f, na) : 0.0f;
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

__global__ void synth_rows_bf16_kernel(uint4* __restrict__ tiles, u32 row_lo, u32 n, u32 dim, u32 KB,
                                       u64 seed, u64 gen_first_row) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 row, col;
    bf16_chunk_coords(t, row_lo >> 6, KB, row, col);
    if (row < row_lo || row >= row_lo + n) return;
    const u64 rk = splitmix64(seed ^ ((gen_first_row + (row - row_lo)) * 0x9E3779B97F4A7C15ull));
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = (col + e < dim) ? synth_elem(rk, col + e) : 0.0f;
    tiles[(size_t)(row_lo >> 6) * KB * 512 + t] = pack8_bf16(v);
}

#endif  // BSR_TEMPLATES_ONLY

}  // namespace bsr
