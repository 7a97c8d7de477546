// bsr_tensor.cuh -- K2: batched cosine scan on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// Used only for BF16 stores and query batches large enough to make the scan a dense contraction
//     D[q, r] = sum_c Q[q, c] * R[r, c]          q: 128 queries (UMMA M), r: 64 store rows (UMMA N)
//
// * the 128 queries of a CTA are rounded to bf16 once and live in TENSOR MEMORY as the A operand
//   (128 lanes x dim/2 32-bit columns = 384 of the 512 TMEM columns for dim 768), written with
//   tcgen05.st; they are never re-read from shared memory;
// * store tiles (64 rows x 64 columns, 8 KiB, already in the K-major SWIZZLE_128B image, see
//   bsr_kernels.cuh) stream HBM -> shared memory through a ring of cp.async.bulk copies (TMA) and
//   are the B operand via a shared-memory descriptor;
// * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 N64 K16, fp32 accumulate)
//   into one of two 64-column TMEM accumulators; tcgen05.commit releases ring slots and hands the
//   finished accumulator to the epilogue through mbarriers;
// * epilogue: TMEM lane q belongs to thread q, which reads its query's 64 scores with tcgen05.ld
//   and selects that query's candidates privately -- no cross-thread reduction.  Selection is
//   bound-based so the common case is branch-free:
//     - rows are split into k residue classes (row mod k); per query the best normalised score of
//       every class is tracked with one LDS + FMNMX + STS per score (gm[class]).  The minimum over
//       the k classes certifies k DISTINCT rows at least that good, so any row below it (minus a
//       1e-6 safety margin) cannot be in the top-k.  Class maxima are merged across CTAs through a
//       global array (atomicMax), so the bound reflects every row the whole grid has seen.
//     - a score is tested with one multiply + compare; only survivors (a fraction ~k ln(n)/n) get
//       the reference's distance formula (src/metrics.rs:157-164) and are appended, unsorted, to
//       the (CTA, query) survivor buffer in global memory.
//     - the first kTcLearnTiles tiles of a CTA only feed the class maxima ("learn") and are scanned
//       again at the end, when the bound is mature.
//     - at the end each warp bitonic-sorts its queries' survivors and writes the best k_pad.
//
// Numerics: products of bf16 values are exact in binary32; the tensor core's accumulation order
// differs from the reference's left-to-right fold, so distances agree with the oracle (run on the
// bf16-rounded rows and queries) to ~1e-6, not bit for bit -- the documented "looser bound" of
// this path.  Everything after the dot product (norms, division, clamp, 1 - sim, identity test,
// tie order) is the same code as the exact path.
#pragma once

#include "bsr_kernels.cuh"

namespace bsr {

constexpr int kTcThreads = 192;  // warps 0-3: query load + epilogue, warp 4: TMA producer, warp 5: MMA
constexpr u32 kTcQ = 128;        // queries per CTA  (UMMA M)
constexpr u32 kTcN = 64;         // store rows per tile (UMMA N)
constexpr u32 kTcAccCols = 128;  // two 64-column fp32 accumulators at TMEM columns [0, 128)
constexpr u32 kTcMaxKB = 12;     // A operand: 32 columns per 64-element block, 384 columns at most
constexpr u32 kTcLearnTiles = 8; // tiles scanned bound-only first, then again at the end
constexpr u32 kTcSurvivorCap = 1024;  // survivor slots per (CTA, query)
constexpr u32 kTcRefresh = 8;    // tiles between merges of the class maxima with the global array

struct TensorScanArgs {
    const unsigned char* tiles;  // BF16 store (tile layout)
    const float* norms;          // |a| per row
    const float* queries;        // [nq_valid][dim] binary32, device
    u64* cta_keys;               // [nq_valid][gridDim.x][k_pad]
    u32* g_gm;                   // [128][k] global class maxima (ordered-float bits), 0 before the launch
    u64* survivors;              // [128][gridDim.x][kTcSurvivorCap]
    u32 nq_valid;                // <= 128
    u32 n_rows, n_tiles, KB, dim, k, k_pad, row_base, stages;
};

// ---- tcgen05 wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit(u32 bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// D[tmem_d] (+)= A[tmem_a] * B[smem desc]   (A from tensor memory, B from shared memory)
__device__ __forceinline__ void tc_mma_ts(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand tile starting at shared address `addr` (1024-byte aligned tile,
// optionally advanced by k*32 bytes inside the swizzle atom): 8-row groups 1024 bytes apart.
__device__ __forceinline__ u64 make_sw128_desc(u32 addr) {
    return (u64)((addr >> 4) & 0x3FFF) | ((u64)1 << 16) | ((u64)(1024 >> 4) << 32) | ((u64)1 << 46) |
           ((u64)2 << 61);
}
// kind::f16, A = B = BF16, D = F32, both K-major, M = 128, N = 64
constexpr u32 kIdescBf16M128N64 = (1u << 4) | (1u << 7) | (1u << 10) | ((kTcN >> 3) << 17) | ((kTcQ >> 4) << 24);

__device__ __forceinline__ void tmem_st16(u32 taddr, const u32* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(u32 taddr, float* v) {
    u32* r = reinterpret_cast<u32*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// order-preserving u32 image of a float (for atomicMax on class maxima); NaN never gets here
__device__ __forceinline__ u32 f32_ordered(float f) { return ordered_from_bits(__float_as_uint(f)); }
__device__ __forceinline__ float f32_from_ordered(u32 o) { return __uint_as_float(bits_from_ordered(o)); }

// exact identity test of a store row against the bf16-rounded query (rare: only when sim ~ 1)
__device__ __noinline__ bool row_equals_query_bf16(const unsigned char* tiles, u32 row, u32 KB, u32 dim,
                                                   const float* q) {
    for (u32 g = 0; g < KB * 8; g++) {
        const uint4 a = *reinterpret_cast<const uint4*>(tiles + bf16_chunk_offset(row, g >> 3, g & 7u, KB));
        const u32 w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const u32 c = g * 8 + e;
            const float qa = c < dim ? __uint_as_float((u32)f32_to_bf16_rne(q[c]) << 16) : 0.0f;
            const float ra = (e & 1) ? bf16_hi(w[e >> 1]) : bf16_lo(w[e >> 1]);
            if (differs(ra, qa)) return false;
        }
    }
    return true;
}

// Rare path, deliberately out of line (the unrolled score loop must stay small enough for the
// instruction cache): exact distance of one surviving (row, query) pair, appended to the survivor
// buffer.  An overflowing buffer (adversarial row order only) is compacted to its best k, exactly.
struct SurvivorState {
    u64* buf;       // survivor slots of this (CTA, query)
    u32 cnt;
    u64 tau_exact;  // k-th best after a compaction, else UINT64_MAX
};
__device__ __noinline__ void tc_survivor(SurvivorState& st, const TensorScanArgs& a, const float* qp, u32 row,
                                         float score, float na, float mag_b, bool tiny_q) {
    float dist = 1.0f;  // zero magnitude (src/metrics.rs:157-159)
    bool maybe_identical = tiny_q;
    if (na != 0.0f && mag_b != 0.0f) {
        float sim = __fdiv_rn(score, __fmul_rn(na, mag_b));
        sim = fminf(fmaxf(sim, -1.0f), 1.0f);
        dist = __fsub_rn(1.0f, sim);
        maybe_identical |= sim >= 0.99999f;
    }
    if (maybe_identical && row_equals_query_bf16(a.tiles, row, a.KB, a.dim, qp)) dist = 0.0f;
    const u64 key = make_key(dist, a.row_base + row);
    if (!(key < st.tau_exact)) return;
    st.buf[st.cnt++] = key;
    if (st.cnt == kTcSurvivorCap) {
        const u32 k = a.k;
        for (u32 x = 0; x < k; x++) {
            u32 best = x;
            for (u32 y = x + 1; y < st.cnt; y++)
                if (st.buf[y] < st.buf[best]) best = y;
            const u64 tmp = st.buf[x];
            st.buf[x] = st.buf[best];
            st.buf[best] = tmp;
        }
        st.cnt = k;
        st.tau_exact = st.buf[k - 1];
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) scan_topk_bf16_tensor_kernel(const TensorScanArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const u32 S = a.stages, K = a.k_pad, KB = a.KB;
    // [ring: S * 8 KiB (1024-byte aligned: swizzle atoms) | class maxima: k_pad*128*4 |
    //  norms + inverse norms: 2 * 2*64*4 | survivor counts: 128*4 | barriers | tmem base]
    unsigned char* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_gm = reinterpret_cast<float*>(ring + (size_t)S * kTileBytes);  // gm[class*128 + q]
    float* s_na = s_gm + (size_t)K * kTcQ;
    float* s_inv = s_na + 2 * kTcN;
    u32* s_cnt = reinterpret_cast<u32*>(s_inv + 2 * kTcN);
    u64* bars = reinterpret_cast<u64*>(s_cnt + kTcQ);
    u64* full_bar = bars;               // [S]   TMA -> MMA
    u64* empty_bar = bars + S;          // [S]   MMA -> TMA
    u64* accf_bar = bars + 2 * S;       // [2]   MMA -> epilogue
    u64* acce_bar = bars + 2 * S + 2;   // [2]   epilogue -> MMA
    u64* aready_bar = bars + 2 * S + 4; // [1]   queries are in TMEM
    u32* s_tmem = reinterpret_cast<u32*>(bars + 2 * S + 5);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (u32 i = 0; i < S; i++) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (u32 i = 0; i < 2; i++) {
            mbar_init(smem_u32(accf_bar + i), 1);
            mbar_init(smem_u32(acce_bar + i), kTcQ);
        }
        mbar_init(smem_u32(aready_bar), kTcQ);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {  // whole warp: allocate all 512 TMEM columns (1 CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem = *s_tmem;
    const u32 tmem_a = tmem + kTcAccCols;  // A operand starts after the accumulators

    // tile sequence of this CTA: its tiles in order, the first `learn` of them bound-only, then those
    // same tiles once more at the end
    const u32 my_tiles = blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const u32 learn = min(kTcLearnTiles, my_tiles);
    const u32 n_seq = my_tiles + learn;

    if (warp == 4) {
        // ================= TMA producer =================
        if (lane == 0) {
            const u64 policy = make_evict_first_policy();
            u32 slot = 0, phase = 0;
            for (u32 sq = 0; sq < n_seq; sq++) {
                const u32 li = sq < my_tiles ? sq : sq - my_tiles;
                const size_t tile = (size_t)blockIdx.x + (size_t)li * gridDim.x;
                for (u32 kb = 0; kb < KB; kb++) {
                    mbar_wait(smem_u32(empty_bar + slot), phase ^ 1);
                    mbar_expect_tx(smem_u32(full_bar + slot), kTileBytes);
                    bulk_g2s(smem_u32(ring + (size_t)slot * kTileBytes), a.tiles + (tile * KB + kb) * kTileBytes,
                             kTileBytes, smem_u32(full_bar + slot), policy);
                    if (++slot == S) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0) {
            mbar_wait(smem_u32(aready_bar), 0);
            tc_fence_after();
            u32 slot = 0, phase = 0;
            for (u32 li = 0; li < n_seq; li++) {
                const u32 buf = li & 1;
                mbar_wait(smem_u32(acce_bar + buf), ((li >> 1) & 1) ^ 1);  // epilogue drained this accumulator
                tc_fence_after();
                const u32 d_tmem = tmem + buf * kTcN;
                for (u32 kb = 0; kb < KB; kb++) {
                    mbar_wait(smem_u32(full_bar + slot), phase);
                    tc_fence_after();
                    const u64 desc = make_sw128_desc(smem_u32(ring + (size_t)slot * kTileBytes));
#pragma unroll
                    for (u32 k4 = 0; k4 < 4; k4++)  // 4 x K16 per 64-column block
                        tc_mma_ts(d_tmem, tmem_a + (kb * 4 + k4) * 8, desc + (u64)(k4 * 2), kIdescBf16M128N64,
                                  (kb | k4) != 0 ? 1u : 0u);
                    tc_commit(smem_u32(empty_bar + slot));  // slot reusable once these MMAs retire
                    if (++slot == S) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(smem_u32(accf_bar + buf));  // accumulator complete
            }
        }
    } else {
        // ================= warps 0-3: query -> TMEM, then epilogue =================
        const u32 q = tid;  // TMEM lane == query index within the CTA
        const bool active = q < a.nq_valid;
        const float* qp = a.queries + (size_t)(active ? q : 0) * a.dim;
        const u32 lane_addr = ((u32)(warp * 32) << 16);
        float qsq = 0.0f;
        for (u32 cb = 0; cb < KB * 2; cb++) {  // 32 elements -> 16 packed columns per store
            u32 r[16];
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const u32 c = cb * 32 + 2 * e;
                const unsigned short lo = (active && c < a.dim) ? f32_to_bf16_rne(qp[c]) : 0;
                const unsigned short hi = (active && c + 1 < a.dim) ? f32_to_bf16_rne(qp[c + 1]) : 0;
                const float flo = __uint_as_float((u32)lo << 16), fhi = __uint_as_float((u32)hi << 16);
                qsq = __fadd_rn(qsq, __fmul_rn(flo, flo));  // |b| over the rounded query, sequential
                qsq = __fadd_rn(qsq, __fmul_rn(fhi, fhi));
                r[e] = (u32)lo | ((u32)hi << 16);
            }
            tmem_st16(tmem_a + lane_addr + cb * 16, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(smem_u32(aready_bar));
        const float mag_b = __fsqrt_rn(qsq);
        const u32 k = a.k;
        // The identity shortcut (src/metrics.rs:148-151) precedes every other rule.  Two vectors
        // within 1e-10 per element have sim ~ 1 unless the query itself is tiny (|b| < 1e-6): such
        // a query is never pruned and checks every row exactly.
        const bool tiny_q = !(mag_b >= 1e-6f);

        float* gm = s_gm + q;  // my class maxima: gm[class * 128]
        for (u32 g = 0; g < k; g++) gm[(size_t)g * kTcQ] = -INFINITY;
        u32* g_gm = a.g_gm + (size_t)q * k;
        SurvivorState surv;
        surv.buf = a.survivors + ((size_t)q * gridDim.x + blockIdx.x) * kTcSurvivorCap;
        surv.cnt = 0;
        surv.tau_exact = kKeyMax;
        float thr = -INFINITY;    // prune rows with t < thr, t = S / |a|

        for (u32 sq = 0; sq < n_seq; sq++) {
            const u32 buf = sq & 1;
            const u32 li = sq < my_tiles ? sq : sq - my_tiles;
            const bool learn_only = sq < learn;
            const u32 tile = blockIdx.x + li * gridDim.x;
            const u32 row0 = tile * kTcN;
            if (q < kTcN) {
                const float na = (row0 + q < a.n_rows) ? __ldg(a.norms + row0 + q) : 0.0f;
                s_na[buf * kTcN + q] = na;
                s_inv[buf * kTcN + q] = (na > 0.0f) ? __fdiv_rn(1.0f, na) : 0.0f;  // NaN/0 -> 0: handled exactly below
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps only

            mbar_wait(smem_u32(accf_bar + buf), (sq >> 1) & 1);
            tc_fence_after();
            float v[kTcN];
            tmem_ld32(tmem + lane_addr + buf * kTcN, v);
            tmem_ld32(tmem + lane_addr + buf * kTcN + 32, v + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(smem_u32(acce_bar + buf));  // scores are in registers: free the accumulator

            if (active) {
                const u32 n_here = min(kTcN, a.n_rows - row0);
                u32 g = row0 % k;  // residue class of the tile's first row
                // pass 1 (branch-free, unrolled): class maxima and a bit mask of the survivors
                const float* inv = s_inv + buf * kTcN;
                const float pass_thr = learn_only ? INFINITY : thr;
                u32 m_lo = 0, m_hi = 0;
#pragma unroll
                for (u32 j = 0; j < kTcN; j++) {
                    // rows past the end of the store have inv == 0 and score 0: they can only raise a
                    // class maximum to 0, never above a real row's, and are masked out below
                    const float t = v[j] * inv[j];
                    float* cell = gm + (size_t)g * kTcQ;
                    *cell = fmaxf(*cell, j < n_here ? t : -INFINITY);  // a NaN score leaves the maximum untouched
                    const u32 bit = (!(t < pass_thr)) ? 1u : 0u;
                    if (j < 32) m_lo |= bit << (j & 31);
                    else m_hi |= bit << (j & 31);
                    g = (g + 1 == k) ? 0 : g + 1;
                }
                if (n_here < kTcN) {  // last tile of the store: drop the padding rows
                    const u64 keep = (1ull << n_here) - 1ull;
                    m_lo &= (u32)keep;
                    m_hi &= (u32)(keep >> 32);
                }
                if (learn_only) m_lo = m_hi = 0;
                // pass 2 (rare): exact distance for the survivors
                if (m_lo | m_hi) {
#pragma unroll
                    for (u32 j = 0; j < kTcN; j++) {
                        if (((j < 32 ? m_lo : m_hi) >> (j & 31)) & 1u)
                            tc_survivor(surv, a, qp, row0 + j, v[j], s_na[buf * kTcN + j], mag_b, tiny_q);
                    }
                }
                // the bound improves like log(rows seen): refresh densely while it is young, then
                // every 8 / 32 / 128 tiles
                const u32 every = sq < 64 ? 8u : (sq < 512 ? 32u : 128u);
                const bool refresh = (sq + 1 == learn) || (sq >= learn && sq < learn + 8) || ((sq + 1) % every == 0);
                if (refresh && !tiny_q) {
                    float bound = INFINITY;
                    for (u32 c0 = 0; c0 < k; c0 += 32) {
                        u32 their_bits[32];  // 32 independent L2 reads in flight, then the merge
#pragma unroll
                        for (u32 i = 0; i < 32; i++) their_bits[i] = (c0 + i < k) ? __ldcg(g_gm + c0 + i) : 0u;
#pragma unroll
                        for (u32 i = 0; i < 32; i++) {
                            if (c0 + i < k) {
                                float mine = gm[(size_t)(c0 + i) * kTcQ];
                                // 0 = nobody has published this class yet
                                const float theirs = their_bits[i] ? f32_from_ordered(their_bits[i]) : -INFINITY;
                                if (mine > theirs) atomicMax(g_gm + c0 + i, f32_ordered(mine));
                                else if (theirs > mine) gm[(size_t)(c0 + i) * kTcQ] = mine = theirs;
                                bound = fminf(bound, mine);
                            }
                        }
                    }
                    // k distinct rows reach `bound`; keep a 1e-6 (in similarity) safety margin
                    thr = (bound == -INFINITY) ? -INFINITY : bound - 1e-6f * mag_b - 1e-6f * fabsf(bound);
                }
            }
        }
        s_cnt[q] = active ? surv.cnt : 0;
        __threadfence_block();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // every warp sorts the survivors of its queries in the (now idle) ring and writes the best K
        {
            u64* scratch = reinterpret_cast<u64*>(ring) + (size_t)warp * kTcSurvivorCap;
            for (u32 qq = warp; qq < a.nq_valid; qq += 4) {
                const u32 n = s_cnt[qq];
                u32 P = K;
                while (P < n) P <<= 1;
                const u64* src = a.survivors + ((size_t)qq * gridDim.x + blockIdx.x) * kTcSurvivorCap;
                for (u32 j = lane; j < P; j += 32) scratch[j] = j < n ? src[j] : kKeyMax;
                __syncwarp();
                warp_sort_list(scratch, P, lane);
                u64* out = a.cta_keys + ((size_t)qq * gridDim.x + blockIdx.x) * K;
                for (u32 j = lane; j < K; j += 32) out[j] = scratch[j];
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

}  // namespace bsr
